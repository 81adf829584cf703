// post.cu -- everything Go1.step() does after the decimation loop, one warp per (environment, agent):
//   post_physics_step (legged_robot_field.py:117-119 -> legged_robot.py:117-157): derived base quantities, gait clock
//   (go1.py:240-279), termination (legged_robot.py:159-169 + legged_robot_field.py:121-146), NPC stepping
//   (go1_sheep.py:35-64), indexed reset (go1.py:110-145, legged_robot.py:394-470) and compute_observations
//   (go1.py:153-196), plus Go1.reset() (go1.py:147-151).
// The reference spends ~100 tiny torch launches, one nonzero() sync and 4 .cpu() syncs here; this is one launch and
// no host round trip (reset envs are handled in place from the reset mask, never gathered into an index list).
#include "common.cuh"

__device__ __forceinline__ float lerp_range(const float *r, float u) { return r[0] + (r[1] - r[0]) * u; }

// _step_contact_targets (go1.py:240-279)
__device__ void dev_gait_clock(const DevParams &p, int m, float dt_policy) {
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
__device__ void dev_env_reset_warp(const DevParams &p, int e, int lane, int GS) {
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
__device__ void dev_agent_observations_warp(const DevParams &p, int e, int a, int lane, int GS) {
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
__device__ void dev_sheep_step_warp(const DevParams &p, int e, uint32_t step_count, int lane, int GS, unsigned gmask) {
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

// A group of POST_GS lanes per (env, agent); a block handles POST_GROUPS / A whole envs.  The scalar part of an agent's work (base-frame
// velocities, Euler angles, gait clock: a few hundred dependent instructions full of atan2f / sinf / fmodf) is executed once per WARP
// instruction whatever the lane count, so a whole warp per agent pays it 8192 times (measured 29 us) and a thread per agent leaves the
// machine empty (r1: 25 us); four lanes per agent issue it 1024 times and still spread the 71-float observation row, the reset and the
// sheep step over lanes.  Env-level decisions (time-out, reset, NPC step) run on the env's first agent group between two block barriers.
#define POST_GS 4
#define POST_THREADS 256
#define POST_GROUPS (POST_THREADS / POST_GS)
__global__ void __launch_bounds__(POST_THREADS) k_post_physics(DevParams p, unsigned int step_count) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ int s_flags[POST_GROUPS], s_reset[POST_GROUPS];
    if (step_count == 0xffffffffu) step_count = (unsigned int)p.ctr[1];     // graph replay: device counter
    const int A = p.A, P = p.P, G = p.G;
    const int envs_per_block = POST_GROUPS / A;
    const int grp = threadIdx.x / POST_GS, gl = threadIdx.x % POST_GS;       // agent slot of the block, lane inside the group
    const unsigned gmask = ((1u << POST_GS) - 1u) << ((threadIdx.x & 31) & ~(POST_GS - 1));
    const int el = grp / A, a = grp % A;
    const int e = blockIdx.x * envs_per_block + el;
    const bool live = el < envs_per_block && e < p.N;
    const float PI = 3.14159265358979323846f;
    const float dt_policy = p.dt * (float)p.decimation;
    if (threadIdx.x < POST_GROUPS) { s_flags[threadIdx.x] = 0; s_reset[threadIdx.x] = 0; }
    __syncthreads();
    if (live) {
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
        if (f && gl == 0) atomicOr(&s_flags[el], f);
    }
    __syncthreads();
    if (live && a == 0 && gl == 0) {
        const int f = s_flags[el];
        int reset = 0;
        long long ep = p.ep_len[e] + 1;
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
        s_reset[el] = reset;
    }
    __syncthreads();
    if (live && a == 0) {                                // the env's first agent group: NPC step, then the indexed reset
        if (P && p.npc_ctrl == MQE_NPC_SHEEP) { dev_sheep_step_warp(p, e, step_count, gl, POST_GS, gmask); __syncwarp(gmask); }
        if (s_reset[el]) dev_env_reset_warp(p, e, gl, POST_GS);
    }
    __syncthreads();                                     // reset wrote state / last_actions of every agent of the env
    if (live) {
        const int m = e * A + a;
        dev_agent_observations_warp(p, e, a, gl, POST_GS);
        __syncwarp(gmask);                               // the row read last_actions before they are overwritten below
        const float *rs = p.root + ((size_t)e * G + a) * 13;
        const float *dof = p.dof + ((size_t)e * (12 * A + p.D) + 12 * a) * 2;
        for (int j = gl; j < 12; j += POST_GS) { p.last_actions[m * 12 + j] = p.actions[m * 12 + j]; p.last_dof_vel[m * 12 + j] = dof[j * 2 + 1]; }
        for (int i = gl; i < 6; i += POST_GS) p.last_root_vel[m * 6 + i] = rs[7 + i];
    }
    // the last block to finish advances the step counter (every block read it at its start, so nobody still needs it)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&p.ctr[2], 1) == (int)gridDim.x - 1) { p.ctr[2] = 0; p.ctr[1] += 1; }
    }
}

// Go1.reset(): reset_idx(arange(N)) then compute_observations (go1.py:147-151); no physics step.  One warp per env.
__global__ void __launch_bounds__(128) k_reset_all(DevParams p) {
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= p.N) return;
    dev_env_reset_warp(p, e, lane, 32);
    __syncwarp();
    for (int a = 0; a < p.A; a++) dev_agent_observations_warp(p, e, a, lane, 32);
    if (p.result_done && lane == 0) p.result_done[(long long)(p.ctr[1] & 1) * p.result_half + e] = 1;
}

// gym.set_actor_root_state_tensor_indexed / set_dof_state_tensor_indexed for hosts that stage state elsewhere
__global__ void k_set_root_indexed(DevParams p, const float *__restrict__ src, const int *__restrict__ ids, int n) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 13) return;
    int actor = ids[t / 13];
    if (actor < 0 || actor >= p.N * p.G) return;
    p.root[(size_t)actor * 13 + t % 13] = src[(size_t)actor * 13 + t % 13];
}
__global__ void k_set_dof_indexed(DevParams p, const float *__restrict__ src, const int *__restrict__ ids, int n) {
    // actor -> its DOF range inside the env: agents own 12 DOFs each, the NPC block owns p.D
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 24) return;
    int actor = ids[t / 24], k = t % 24;
    if (actor < 0 || actor >= p.N * p.G) return;
    int e = actor / p.G, g = actor % p.G, per_env = 12 * p.A + p.D, first, cnt;
    if (g < p.A) { first = 12 * g; cnt = 12; }
    else { int per = p.P ? p.D / p.P : 0; first = 12 * p.A + (g - p.A) * per; cnt = per; }
    if (k >= 2 * cnt) return;
    size_t off = ((size_t)e * per_env + first) * 2 + k;
    p.dof[off] = src[off];
}

extern "C" cudaError_t mqe_launch_post(const DevParams &p, unsigned int step_count, cudaStream_t st) {
    const int envs_per_block = POST_GROUPS / p.A;
    return launch_heavy(k_post_physics, dim3((p.N + envs_per_block - 1) / envs_per_block), dim3(POST_THREADS), 0, st, p, step_count);
}
extern "C" cudaError_t mqe_launch_reset_all(const DevParams &p, cudaStream_t st) {
    k_reset_all<<<(p.N + 3) / 4, 128, 0, st>>>(p);
    return cudaGetLastError();
}
extern "C" cudaError_t mqe_launch_set_root_indexed(const DevParams &p, const float *src, const int *ids, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_set_root_indexed<<<(n * 13 + 255) / 256, 256, 0, st>>>(p, src, ids, n);
    return cudaGetLastError();
}
extern "C" cudaError_t mqe_launch_set_dof_indexed(const DevParams &p, const float *src, const int *ids, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_set_dof_indexed<<<(n * 24 + 255) / 256, 256, 0, st>>>(p, src, ids, n);
    return cudaGetLastError();
}
