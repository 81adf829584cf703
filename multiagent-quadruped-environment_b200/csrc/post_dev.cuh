// post_dev.cuh -- the device side of Go1's post-physics bookkeeping, shared by k_post_physics (post.cu: the stand-alone launch) and by
// the epilogue of k_substeps (physics.cu: the warp that integrated an env also finishes its step, so the ~24 us bookkeeping launch hides
// in the physics kernel's tail).  Both callers give every (env, agent) a group of FOUR lanes.
//   post_physics_step (legged_robot_field.py:117-119 -> legged_robot.py:117-157): derived base quantities, gait clock
//   (go1.py:240-279), termination (legged_robot.py:159-169 + legged_robot_field.py:121-146), NPC stepping
//   (go1_sheep.py:35-64), indexed reset (go1.py:110-145, legged_robot.py:394-470) and compute_observations (go1.py:153-196).
#pragma once
#include "common.cuh"

__device__ __forceinline__ float lerp_range(const float *r, float u) { return r[0] + (r[1] - r[0]) * u; }

// _step_contact_targets (go1.py:240-279)
static __device__ void dev_gait_clock(const DevParams &p, int m, float dt_policy) {
    const float *lo = p.loc_obs + (size_t)m * MQE_LOC_OBS;
    float freq = lo[7], phase = lo[8], offset = lo[9], bound = lo[10], dur = lo[11];
    float g = fmodf(p.gait[m] + dt_policy * freq, 1.0f);
    if (g < 0.f) g += 1.f;
    p.gait[m] = g;
    float fi[4] = {g + phase + offset + bound, g + offset, g + bound, g + phase};
    for (int i = 0; i < 4; i++) {
        float r = fmodf(fi[i], 1.0f);
        if (r < 0.f) r += 1.f;
        float x = fi[i];
        if (r < dur) x = r * (0.5f / dur);
        else if (r > dur) x = 0.5f + (r - dur) * (0.5f / (1.f - dur));
        p.clock[m * 4 + i] = sinf(6.28318530717958647692f * x);
    }
}

// ---- warp-cooperative helpers (one warp per agent / per env) ----
// reset_idx for one env by one warp: _reset_dofs, _reset_root_states, _reset_buffers (go1.py:110-145, legged_robot.py:394-470)
static __device__ void dev_env_reset_warp(const DevParams &p, int e, int lane, int GS) {
    const int A = p.A, P = p.P, G = p.G;
    const uint32_t ge = (uint32_t)(e + p.env_off), ep = p.episode[e];
    float *dof = p.dof + (size_t)e * (12 * A + p.D) * 2;
    for (int t = lane; t < A * 12; t += GS) {
        const int a = t / 12, j = t % 12, m = e * A + a;
        const float u = rng_uniform(p.seed, ge, ep, RNG_DOF, a * 12 + j);
        dof[(12 * a + j) * 2] = p.model->q_default[j] * (p.dof_lo + (p.dof_hi - p.dof_lo) * u);
        dof[(12 * a + j) * 2 + 1] = 0.f;
        p.last_actions[m * 12 + j] = 0.f; p.last_dof_vel[m * 12 + j] = 0.f;
    }
    for (int t = lane; t < A * 13; t += GS) {
        const int a = t / 13, i = t % 13, m = e * A + a;
        float v = p.base_init[m * 13 + i];
        if (i < 3) v += p.agent_origins[m * 3 + i];
        if (p.has_bpos && i < 2) v += lerp_range(i == 0 ? p.bpos_x : p.bpos_y, rng_uniform(p.seed, ge, ep, RNG_BASE_POS, a * 2 + i));
        if (i >= 7) v = p.bvel_lo + (p.bvel_hi - p.bvel_lo) * rng_uniform(p.seed, ge, ep, RNG_BASE_VEL, a * 6 + (i - 7));
        p.root[((size_t)e * G + a) * 13 + i] = v;
    }
    for (int a = lane; a < A; a += GS) p.gait[e * A + a] = 0.f;
    for (int k = lane; k < p.D; k += GS) { dof[(12 * A + k) * 2] = p.npc_dof_default[k]; dof[(12 * A + k) * 2 + 1] = 0.f; }
    for (int n = lane; n < P; n += GS) {               // one lane per NPC (quaternion needs all three angles)
        float *rs = p.root + ((size_t)e * G + A + n) * 13;
        for (int i = 0; i < 13; i++) rs[i] = p.npc_init[(e * P + n) * 13 + i];
        for (int i = 0; i < 3; i++) rs[i] += p.env_origins[e * 3 + i];
        if (p.has_npos) {
            rs[0] += lerp_range(p.npos_x, rng_uniform(p.seed, ge, ep, RNG_NPC_POS, n * 2));
            rs[1] += lerp_range(p.npos_y, rng_uniform(p.seed, ge, ep, RNG_NPC_POS, n * 2 + 1));
        }
        if (p.has_nrpy) {
            float r = lerp_range(p.nrpy_r, rng_uniform(p.seed, ge, ep, RNG_NPC_RPY, n * 3));
            float pp = lerp_range(p.nrpy_p, rng_uniform(p.seed, ge, ep, RNG_NPC_RPY, n * 3 + 1));
            float y = lerp_range(p.nrpy_y, rng_uniform(p.seed, ge, ep, RNG_NPC_RPY, n * 3 + 2));
            quat_from_euler_xyz(r, pp, y, rs + 3);
        }
    }
    if (lane == 0) {
        p.hist_dirty[e] = 1;          // history_locomotion_obs[env_ids] = 0, applied by the next k_policy_frame
        p.ep_len[e] = 0;
        p.reset_buf[e] = 1;
        p.episode[e] = ep + 1;
    }
}

// compute_observations for one agent by one warp (go1.py:153-196): one obs entry per lane and trip
static __device__ void dev_agent_observations_warp(const DevParams &p, int e, int a, int lane, int GS) {
    const int A = p.A, G = p.G;
    const int m = e * A + a;
    float *ob = p.obs + (size_t)m * MQE_OBS_FLOATS;
    const float *rs = p.root + ((size_t)e * G + a) * 13;
    const float *dof = p.dof + ((size_t)e * (12 * A + p.D) + 12 * a) * 2;
    const float *bq = p.quat_alias ? rs + 3 : p.base_quat + m * 4;
    const float q4[4] = {bq[0], bq[1], bq[2], bq[3]};
    float rpy[3];
    get_euler_xyz(q4, rpy);
    for (int i = lane; i < MQE_OBS_FLOATS; i += GS) {
        float v;
        if (i < MQE_OBS_BASE_QUAT) v = rs[i] - p.env_origins[e * 3 + i];
        else if (i < MQE_OBS_DOF_POS) { const int k = i - MQE_OBS_BASE_QUAT; v = k == 0 ? q4[0] : (k == 1 ? q4[1] : (k == 2 ? q4[2] : q4[3])); }
        else if (i < MQE_OBS_DOF_VEL) v = dof[(i - MQE_OBS_DOF_POS) * 2] - p.model->q_default[i - MQE_OBS_DOF_POS];
        else if (i < MQE_OBS_LIN_VEL) v = dof[(i - MQE_OBS_DOF_VEL) * 2 + 1] * 0.05f;
        else if (i < MQE_OBS_ANG_VEL) v = p.base_lin_vel[m * 3 + i - MQE_OBS_LIN_VEL] * 2.0f;
        else if (i < MQE_OBS_LAST_ACTION) v = p.base_ang_vel[m * 3 + i - MQE_OBS_ANG_VEL] * 0.25f;
        else if (i < MQE_OBS_LAST_LAST_ACTION) v = p.actions[m * 12 + i - MQE_OBS_LAST_ACTION];
        else if (i < MQE_OBS_PROJ_GRAVITY) v = p.last_actions[m * 12 + i - MQE_OBS_LAST_LAST_ACTION];
        else if (i < MQE_OBS_CLOCK) v = p.proj_grav[m * 3 + i - MQE_OBS_PROJ_GRAVITY];
        else if (i < MQE_OBS_BASE_RPY) v = p.clock[m * 4 + i - MQE_OBS_CLOCK];
        else { const int k = i - MQE_OBS_BASE_RPY; v = k == 0 ? rpy[0] : (k == 1 ? rpy[1] : rpy[2]); }
        ob[i] = v;
    }
}

// Go1Sheep._step_npc by one warp: every lane forms the flock statistics in the scalar code's order, lane n then moves sheep n
static __device__ void dev_sheep_step_warp(const DevParams &p, int e, uint32_t step_count, int lane, int GS, unsigned gmask) {
    const int A = p.A, P = p.P, G = p.G;
    float *root = p.root + (size_t)e * G * 13;
    float avg[3] = {0.f, 0.f, 0.f}, var[2] = {0.f, 0.f};
    for (int n = 0; n < P; n++) for (int i = 0; i < 3; i++) avg[i] += root[(A + n) * 13 + i] / (float)P;
    for (int n = 0; n < P; n++) for (int i = 0; i < 2; i++) { float t = root[(A + n) * 13 + i] - avg[i]; var[i] += t * t / (float)P; }
    __syncwarp(gmask);                                     // every lane of the group has read the pre-step positions
    if (lane == 0) { p.sheep_stats[e * 3] = avg[0]; p.sheep_stats[e * 3 + 1] = avg[1]; p.sheep_stats[e * 3 + 2] = var[0] + var[1]; }
    const uint32_t ge = (uint32_t)(e + p.env_off);
    for (int n = lane; n < P; n += GS) {                   // a sheep only writes its own row; agents' rows are read-only here
        float *rs = root + (A + n) * 13, dv[3];
        for (int i = 0; i < 3; i++) dv[i] = p.sheep_rand * rng_normal(p.seed, ge, step_count, RNG_SHEEP, n * 3 + i) * 2.f;
        if (P != 1) {
            float rel[3] = {avg[0] - rs[0], avg[1] - rs[1], avg[2] - rs[2]};
            float nn = sqrtf(rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]);
            for (int i = 0; i < 3; i++) dv[i] += p.sheep_rand * rel[i] / nn / 1.5f;
        }
        for (int a = 0; a < A; a++) {
            float rel[3] = {rs[0] - root[a * 13], rs[1] - root[a * 13 + 1], rs[2] - root[a * 13 + 2]};
            float sq[3] = {rel[0] * rel[0], rel[1] * rel[1], rel[2] * rel[2]};
            float dis = sqrtf(sq[0] * sq[0] + sq[1] * sq[1] + sq[2] * sq[2]);     // torch.norm(relative_pos ** 2)
            if (dis > 9.f) continue;
            float den = powf(dis, 1.4f);
            for (int i = 0; i < 3; i++) dv[i] += p.sheep_scale * rel[i] / den;
        }
        dv[2] = 0.f;
        for (int i = 0; i < 3; i++) rs[7 + i] += dv[i];
        for (int i = 0; i < 2; i++) rs[7 + i] = fminf(fmaxf(rs[7 + i], -2.f), 2.f);
        rs[2] = fminf(fmaxf(rs[2], 0.f), 0.3f);
        rs[3] = 0.f; rs[4] = 0.f;
    }
}


// ---- the four stages of one env's post-physics step; lanes gl = 0..3 of an agent's group (mask gmask) ----
// stage 1, every agent group: base-frame velocities, projected gravity, gait clock, pushes; returns the agent's termination flags
// (1 roll, 2 pitch, 4 z low, 8 z high, 16 base contact, 32 non-finite state)
static __device__ int dev_post_agent_derive(const DevParams &p, int e, int a, int gl, unsigned gmask, unsigned step_count) {
    const int A = p.A, G = p.G;
    const float PI = 3.14159265358979323846f;
    const float dt_policy = p.dt * (float)p.decimation;
    const int m = e * A + a;
    const float *rs = p.root + ((size_t)e * G + a) * 13;
    const float q4[4] = {rs[3], rs[4], rs[5], rs[6]};                    // same addresses in the lanes of a group: broadcast loads
    const V3 lv = quat_rotate_inverse(q4, mk(rs[7], rs[8], rs[9]));
    const V3 av = quat_rotate_inverse(q4, mk(rs[10], rs[11], rs[12]));
    const V3 pg = quat_rotate_inverse(q4, mk(0.f, 0.f, -1.f));
    p.base_quat[m * 4 + gl] = gl == 0 ? q4[0] : (gl == 1 ? q4[1] : (gl == 2 ? q4[2] : q4[3]));   // value select: no local array
    if (gl < 3) {
        p.base_lin_vel[m * 3 + gl] = comp(lv, gl);
        p.base_ang_vel[m * 3 + gl] = comp(av, gl);
        p.proj_grav[m * 3 + gl] = comp(pg, gl);
    }
    if (p.control_type == 0 && gl == 0) dev_gait_clock(p, m, dt_policy);     // _step_contact_targets runs for control_type 'C' only (go1.py:241)
    // _push_robots (go1.py:237-238, legged_robot.py:472-477): common_step_counter % push_interval == 0 -> every robot's base
    // velocity x, y is redrawn; it takes effect in the next physics step (the derived base quantities above are pre-push)
    const bool push = p.push_interval > 0 && ((step_count + 1u) % (unsigned)p.push_interval) == 0u;
    int f = 0;
    const float *cf = p.contact + ((size_t)e * p.NB + a * MQE_NUM_BODIES) * 3;          // body 0 = base
    if (sqrtf(cf[0] * cf[0] + cf[1] * cf[1] + cf[2] * cf[2]) > 1.f) f |= 16;
    float rpy[3];
    get_euler_xyz(q4, rpy);
    if (rpy[0] > PI) rpy[0] -= 2.f * PI;
    if (rpy[1] > PI) rpy[1] -= 2.f * PI;
    const float z = rs[2] - p.agent_origins[m * 3 + 2];
    if (fabsf(rpy[0]) > p.term_roll) f |= 1;
    if (fabsf(rpy[1]) > p.term_pitch) f |= 2;
    if (z < p.term_zlow) f |= 4;
    if (z > p.term_zhigh) f |= 8;
    // safety net outside the reference's semantics: a non-finite or exploding state can never terminate on its own (every
    // comparison with NaN is false), so it is reset here.  Never taken in the parity tests or in 3000-step soak runs.
    const float chk = rs[0] + rs[1] + rs[2] + q4[0] + q4[1] + q4[2] + q4[3] + lv.x + lv.y + lv.z + av.x + av.y + av.z;
    if (!(fabsf(chk) < 1e6f)) f |= 32;
    __syncwarp(gmask);                                                   // all lanes of the group have read the pre-push velocity
    if (push && gl < 2)
        p.root[((size_t)e * G + a) * 13 + 7 + gl] = (2.f * rng_uniform(p.seed, (uint32_t)(p.env_off + e), step_count, RNG_PUSH, 2 * a + gl) - 1.f) * p.max_push_vel;
    return f;
}

// stage 2, ONE lane per env: episode length, time-out, the termination causes the task enabled; f = OR of the agents' flags.  Returns reset.
static __device__ int dev_post_env_decide(const DevParams &p, int e, int f, unsigned step_count, long long ep_loaded = -1) {
    const int A = p.A;
    int reset = 0;
    long long ep = (ep_loaded >= 0 ? ep_loaded : p.ep_len[e]) + 1;      // ep_loaded: the caller issued the load earlier, with its other loads
    p.ep_len[e] = ep;
    if (p.term_mask & 16) { p.collide_buf[e] = (unsigned char)((f >> 4) & 1); reset |= (f >> 4) & 1; }
    int to = ep > (long long)p.max_ep_len;
    p.timeout_buf[e] = (unsigned char)to;
    reset |= to;
    if (p.term_mask & 1) { p.r_term[e] = (unsigned char)(f & 1); reset |= f & 1; }
    if (p.term_mask & 2) { p.p_term[e] = (unsigned char)((f >> 1) & 1); reset |= (f >> 1) & 1; }
    if (p.term_mask & 4) { p.zl_term[e] = (unsigned char)((f >> 2) & 1); reset |= (f >> 2) & 1; }
    if (p.term_mask & 8) { p.zh_term[e] = (unsigned char)((f >> 3) & 1); reset |= (f >> 3) & 1; }
    if (f & 32) {                                    // blown-up env: also clear what a normal reset keeps (actuator / action histories)
        reset = 1;
        atomicAdd(p.stats + 4, 1);
        for (int i = e * A * 12; i < (e + 1) * A * 12; i++) { p.err1[i] = p.err2[i] = p.vel1[i] = p.vel2[i] = 0.f; p.loc_last[i] = p.loc_last2[i] = 0.f; p.actions[i] = 0.f; }
    }
    p.reset_buf[e] = (unsigned char)reset;
    if (p.result_done) p.result_done[(long long)((step_count + 1u) & 1u) * p.result_half + e] = (unsigned char)reset;     // the learner's copy
    // legged_robot.py:164-169 binds reset_buf and collide_buf to ONE tensor and ORs every later cause in place, so the
    // reference's collide_buf equals the full reset mask whenever base contacts terminate
    if (p.term_mask & 16) p.collide_buf[e] = (unsigned char)reset;
    return reset;
}

// stage 3, the env's first agent group: NPC step, then the indexed reset
static __device__ __forceinline__ void dev_post_env_npc_reset(const DevParams &p, int e, int reset, int gl, unsigned gmask, unsigned step_count) {
    if (p.P && p.npc_ctrl == MQE_NPC_SHEEP) { dev_sheep_step_warp(p, e, step_count, gl, 4, gmask); __syncwarp(gmask); }
    if (reset) dev_env_reset_warp(p, e, gl, 4);
}

// stage 4, every agent group, after the env's reset is visible: observation row, then the last_* copies
static __device__ void dev_post_agent_finish(const DevParams &p, int e, int a, int gl, unsigned gmask) {
    const int A = p.A, G = p.G;
    const int m = e * A + a;
    dev_agent_observations_warp(p, e, a, gl, 4);
    __syncwarp(gmask);                               // the row read last_actions before they are overwritten below
    const float *rs = p.root + ((size_t)e * G + a) * 13;
    const float *dof = p.dof + ((size_t)e * (12 * A + p.D) + 12 * a) * 2;
    for (int j = gl; j < 12; j += 4) { p.last_actions[m * 12 + j] = p.actions[m * 12 + j]; p.last_dof_vel[m * 12 + j] = dof[j * 2 + 1]; }
    for (int i = gl; i < 6; i += 4) p.last_root_vel[m * 6 + i] = rs[7 + i];
}
