// api.cu -- the C ABI of libmqe_b200.so (include/mqe_b200.h): engine handle, engine-owned device buffers, and the
// step / reset entry points that stand where the Isaac Gym tensor API stood in the reference
// (mqe/envs/go1/go1.py:35-62, mqe/envs/base/legged_robot.py:117-157, 394-470, 549-595).
// No CPU path exists behind these calls: without an sm_100 device mqe_sim_create fails with MQE_ERR_NO_DEVICE.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_fp16.h>

#include "kernels.cuh"

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(MQE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                \
    } while (0)

struct BufInfo { void *ptr; int64_t shape[4]; int32_t elem; };

struct MqeSim {
    DevParams p;
    int device = 0;
    cudaStream_t stream = nullptr;
    int M = 0, actrl = 0, maxpair = MQE_MAX_PAIR;
    std::vector<void *> allocs;
    BufInfo bufs[MQE_BUF_COUNT];
    PolicyWeightsDev pw;
    PolicyScratch ps;
    PolicyTcWeights tcw = {};
    unsigned int step_count = 0;
    int head = MQE_HIST_FRAMES - 1;      // slot of the newest frame; the next frame goes to (head + 1) % 30
    long long launches = 0;
    float *d_actions_stage = nullptr;
    float *h_actions = nullptr, *h_obs = nullptr;
    unsigned char *h_reset = nullptr;
    float *tmp_ring = nullptr;
    unsigned short *tmp_hi = nullptr, *tmp_lo = nullptr;
    int tmp_rows = 0;
    bool tail_fp32 = false;              // MQE_TC_TAIL=0: keep layers 1.. on the CUDA-core path (debug cross-check)
    // CUDA graph of one whole policy step (tensor-core modes): captured once per action scale, replayed every step
    struct StepGraph { float scale[3]; cudaGraphExec_t exec; int launches; };
    std::vector<StepGraph> graphs;
    cudaStream_t cap_stream = nullptr;
    bool use_graph = false;
    long long plain_steps = 0;
    cudaStream_t aux_stream = nullptr;   // forked policy: adaptation branch
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_stage[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // mqe_sim_stage_timing: step start, policy done, physics done, bookkeeping done, join done
    int stage_timing = 0;
    bool fork_policy = false;
    bool fused_policy = false;           // one layer-0 launch + one fused tail kernel (default in the tensor-core modes)
    bool incremental = false;            // incremental layer 0: the 29 known frames of the next step are contracted behind this step's physics
    int zold_head = -1;                  // ring slot (of the NEXT frame) the partial sums in ps.Zold were computed for; -1: none
    int early_tiles = 0;                 // row tiles of the background pass that start beside the fused tail (MQE_L0_EARLY_TILES)
    int balance = 0;                     // k_balance_tasks before every k_substeps of a step (MQE_BALANCE; default: grids of more than one round)
    int *d_task_order = nullptr;
    cudaEvent_t ev_bal0 = nullptr, ev_bal1 = nullptr;
    int balance_at_end = 1;              // k_balance_tasks at the end of the step (default) or at the start of the next one, beside the policy (MQE_BALANCE_WHEN=start)
    int bal_pending = 0;                 // this step's k_balance_tasks has been forked and not joined yet
    cudaStream_t bal_stream = nullptr;   // its own side stream: on aux_stream the background layer-0 node would inherit an edge from it and the
                                         // instantiated graph then launches that pass BEFORE k_substeps (+57 us on go1gate)
    int fuse_post = 1;                   // mqe_sim_step: post-physics stages run in the epilogue of k_substeps (MQE_FUSE_POST=0: separate launch)
    int bg_early = 0;                    // set by step_plain around policy_impl: this call may fork the early part
    cudaEvent_t ev_early = nullptr;
    WrapParams wrap = {};                // fused task-wrapper gather (mqe_sim_set_wrapper); kind 0 = off
    // what the learner reads after a step, packed (MQE_BUF_STEP_RESULT): wrapper obs | reward | done
    unsigned char *d_result = nullptr, *h_result = nullptr;
    MqeStepResultLayout rl = {};
    GatherParams gather = {};            // peer exchange of the step result (mqe_sim_gather_*); world 0 = off
    void *gather_peers[MQE_MAX_RANKS] = {};
    int gather_pending_world = 0;
    unsigned long long gather_seq = 0;   // exchanges enqueued so far (host mirror of ctr[5])
    struct Pinned { char *ptr; size_t bytes; };
    std::vector<Pinned> pinned;          // host ranges registered with mqe_sim_pin_host: mqe_sim_step_host copies straight to / from them
    size_t action_bytes() const {        // [N][A_ctrl][3] commands ('C') or [N][12A] joint actions ('P' / 'V' / 'T')
        return (p.control_type == 0 ? (size_t)p.N * actrl * 3 : (size_t)p.N * p.A * 12) * sizeof(float);
    }
    bool is_pinned(const void *ptr, size_t bytes) const {
        for (const auto &r : pinned)
            if ((const char *)ptr >= r.ptr && (const char *)ptr + bytes <= r.ptr + r.bytes) return true;
        return false;
    }
};

template <typename T>
static cudaError_t dalloc(MqeSim *s, T **out, size_t n, bool zero = true) {
    void *ptr = nullptr;
    size_t bytes = (n ? n : 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&ptr, bytes);
    if (e != cudaSuccess) return e;
    s->allocs.push_back(ptr);
    if (zero) { e = cudaMemsetAsync(ptr, 0, bytes, s->stream); if (e != cudaSuccess) return e; }
    *out = (T *)ptr;
    return cudaSuccess;
}
template <typename T>
static cudaError_t dupload(MqeSim *s, const T **out, const T *h, size_t n) {
    T *d = nullptr;
    cudaError_t e = dalloc(s, &d, n, false);
    if (e != cudaSuccess) return e;
    *out = d;
    // pageable source: cudaMemcpyAsync stages it before returning
    return cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, s->stream);
}
static void set_buf(MqeSim *s, int which, void *ptr, int elem, int64_t a, int64_t b = 0, int64_t c = 0, int64_t d = 0) {
    s->bufs[which].ptr = ptr; s->bufs[which].elem = elem;
    s->bufs[which].shape[0] = a; s->bufs[which].shape[1] = b; s->bufs[which].shape[2] = c; s->bufs[which].shape[3] = d;
}

// (Re)allocate MQE_BUF_STEP_RESULT for the current wrapper shape: [obs f32 N*Aw*D | reward f32 N*Aw | done u8 N], every region 16-byte aligned
static cudaError_t alloc_step_result(MqeSim *s, int Aw, int D) {
    const int N = s->p.N;
    auto up16 = [](int64_t v) { return (v + 15) & ~(int64_t)15; };
    MqeStepResultLayout &L = s->rl;
    L.num_envs = N; L.Aw = Aw; L.D = D; L.reserved = 0;
    L.obs_off = 0; L.obs_bytes = (int64_t)N * Aw * D * 4;
    L.reward_off = up16(L.obs_off + L.obs_bytes); L.reward_bytes = (int64_t)N * Aw * 4;
    L.done_off = up16(L.reward_off + L.reward_bytes); L.done_bytes = N;
    L.total_bytes = up16(L.done_off + L.done_bytes);
    cudaError_t e = dalloc(s, &s->d_result, (size_t)(2 * L.total_bytes));      // double buffered: half = policy steps done & 1
    if (e != cudaSuccess) return e;
    s->p.result_half = L.total_bytes;
    if (s->h_result) { cudaFreeHost(s->h_result); s->h_result = nullptr; }
    if ((e = cudaMallocHost(&s->h_result, (size_t)L.total_bytes)) != cudaSuccess) return e;
    s->p.result_done = s->d_result + L.done_off;
    e = cudaMemsetAsync(s->p.result_done, 1, N, s->stream);            // reset_buf starts as ones (base_task.py:77)
    if (e == cudaSuccess) e = cudaMemsetAsync(s->p.result_done + L.total_bytes, 1, N, s->stream);
    set_buf(s, MQE_BUF_STEP_RESULT, s->d_result, 1, 2, L.total_bytes);
    return e;
}
static void drop_graphs(MqeSim *s) {
    for (auto &g : s->graphs) cudaGraphExecDestroy(g.exec);
    s->graphs.clear();
}

extern "C" {

static int exchange_impl(MqeSim *s);
const char *mqe_last_error(void) { return g_err.c_str(); }
int mqe_abi_version(void) { return MQE_ABI_VERSION; }
int mqe_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mqe_sim_destroy(MqeSim *s) {
    if (!s) return MQE_OK;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    for (auto &g : s->graphs) cudaGraphExecDestroy(g.exec);
    for (auto &r : s->pinned) cudaHostUnregister(r.ptr);
    if (s->cap_stream) cudaStreamDestroy(s->cap_stream);
    if (s->aux_stream) cudaStreamDestroy(s->aux_stream);
    for (auto &e : s->ev_stage) if (e) cudaEventDestroy(e);
    if (s->ev_bal0) cudaEventDestroy(s->ev_bal0);
    if (s->ev_bal1) cudaEventDestroy(s->ev_bal1);
    if (s->bal_stream) cudaStreamDestroy(s->bal_stream);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->ev_early) cudaEventDestroy(s->ev_early);
    for (void *ptr : s->allocs) cudaFree(ptr);
    if (s->tcw.blob) cudaFree(s->tcw.blob);
    if (s->h_actions) cudaFreeHost(s->h_actions);
    if (s->h_obs) cudaFreeHost(s->h_obs);
    if (s->h_reset) cudaFreeHost(s->h_reset);
    if (s->h_result) cudaFreeHost(s->h_result);
    for (int r = 0; r < s->gather.world; r++)
        if (r != s->gather.rank && s->gather_peers[r]) cudaIpcCloseMemHandle(s->gather_peers[r]);
    if (s->tmp_ring) cudaFree(s->tmp_ring);
    if (s->tmp_hi) cudaFree(s->tmp_hi);
    if (s->tmp_lo) cudaFree(s->tmp_lo);
    delete s;
    return MQE_OK;
}

static int create_impl(const MqeSimDesc *d, int device, void *stream, MqeSim *s) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MQE_ERR_NO_DEVICE, "no CUDA device visible; libmqe_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(MQE_ERR_INVALID, "device ordinal out of range");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(MQE_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", kernels are built for sm_100a only");
    CK(cudaSetDevice(device));
    s->device = device;
    s->stream = (cudaStream_t)stream;
    const int N = d->num_envs, A = d->num_agents, P = d->num_npcs, D = d->npc_dofs, G = A + P, M = N * A;
    if (N <= 0 || A <= 0 || A > 4 || P < 0 || P > 12) return fail(MQE_ERR_INVALID, "num_envs/num_agents/num_npcs out of range (A <= 4, P <= 12)");
    if (d->defender && !(A == 3 && P >= 1)) return fail(MQE_ERR_INVALID, "defender needs 3 agents and the ball");
    s->M = M;
    s->actrl = d->defender ? A - 1 : A;
    s->maxpair = (d->max_pair_contacts > 0 && d->max_pair_contacts < MQE_MAX_PAIR) ? d->max_pair_contacts : MQE_MAX_PAIR;
    DevParams &p = s->p;
    memset(&p, 0, sizeof p);
    p.N = N; p.A = A; p.P = P; p.D = D; p.G = G;
    p.Pd = (d->npc_kind == MQE_NPC_RIGID || d->npc_kind == MQE_NPC_SEESAW || d->npc_kind == MQE_NPC_BOX) ? P : 0;   // NPCs that own a lane of the substep kernel
    if (d->npc_kind == MQE_NPC_SEESAW && (P != 1 || D != 1)) return fail(MQE_ERR_UNSUPPORTED, "seesaw: exactly one NPC with one DOF");
    if (d->npc_kind == MQE_NPC_BOX && P != 1) return fail(MQE_ERR_UNSUPPORTED, "box: exactly one NPC");
    for (int i = 0; i < 16; i++) p.geom[i] = d->npc_geom[i];
    p.env_off = d->env_id_offset;
    p.npc_kind = d->npc_kind; p.npc_ctrl = d->npc_ctrl;
    p.decimation = d->decimation; p.iters = d->solver_iters; p.max_ep_len = d->max_episode_length;
    p.term_mask = d->term_mask; p.quat_alias = d->quat_alias; p.defender = d->defender; p.command_vel = d->command_vel;
    p.policy_mode = d->policy_mode;
    const int lanes = 4 * A + p.Pd;
    if (lanes > 32) return fail(MQE_ERR_UNSUPPORTED, "4*num_agents + dynamic npcs must fit one warp");
    p.E = 32 / lanes;
    p.NB = MQE_NUM_BODIES * A + P;
    p.dt = d->sim_dt; p.gz = d->gravity_z; p.mu = d->friction; p.coff = d->contact_offset; p.vdep = d->max_depen_vel;
    p.erp = d->erp; p.cfm = d->cfm; p.floor_z = d->floor_z; p.wall_top = d->wall_top_z; p.limit_margin = d->limit_margin;
    p.term_roll = d->term_roll; p.term_pitch = d->term_pitch; p.term_zlow = d->term_zlow; p.term_zhigh = d->term_zhigh;
    for (int i = 0; i < 3; i++) { p.act_scale[i] = d->act_scale[i]; p.cmd_scale[i] = d->cmd_scale[i]; }
    p.action_scale = d->action_scale; p.hip_scale = d->hip_scale; p.clip_actions = d->clip_actions;
    p.dof_lo = d->dof_ratio_lo; p.dof_hi = d->dof_ratio_hi; p.bvel_lo = d->base_vel_lo; p.bvel_hi = d->base_vel_hi;
    p.has_bpos = d->has_base_pos_range; p.has_npos = d->has_npc_pos_range; p.has_nrpy = d->has_npc_rpy_range;
    for (int i = 0; i < 2; i++) {
        p.bpos_x[i] = d->base_pos_x[i]; p.bpos_y[i] = d->base_pos_y[i]; p.npos_x[i] = d->npc_pos_x[i]; p.npos_y[i] = d->npc_pos_y[i];
        p.nrpy_r[i] = d->npc_rpy_r[i]; p.nrpy_p[i] = d->npc_rpy_p[i]; p.nrpy_y[i] = d->npc_rpy_y[i];
    }
    p.npc_mass = d->npc_mass; p.npc_inertia = d->npc_inertia; p.npc_radius = d->npc_radius; p.npc_halflen = d->npc_halflen;
    p.sheep_scale = d->sheep_scale; p.sheep_rand = d->sheep_randomness; p.gate_x = d->gate_x;
    p.seed = d->seed;
    if (d->control_type < 0 || d->control_type > 3) return fail(MQE_ERR_UNSUPPORTED, "control_type must be 0 (C), 1 (P), 2 (T) or 3 (V)");
    p.control_type = d->control_type; p.kp = d->stiffness; p.kd = d->damping;
    p.soft_limit = d->soft_dof_pos_limit > 0.f ? d->soft_dof_pos_limit : 1.f;
    p.push_interval = d->push_interval > 0 ? d->push_interval : 0; p.max_push_vel = d->max_push_vel_xy;
    p.sdf_nx = d->sdf_nx; p.sdf_ny = d->sdf_ny; p.sdf_cell = d->sdf_cell;
    if (!d->h_sdf || !d->h_env_origins || !d->h_agent_origins || !d->h_base_init_state) return fail(MQE_ERR_INVALID, "descriptor host arrays missing");
    if (P && !d->h_npc_init_state) return fail(MQE_ERR_INVALID, "h_npc_init_state missing");
    if (mqe_substeps_smem_bytes(N, A, p.Pd, p.E, s->maxpair) > (size_t)prop.sharedMemPerBlockOptin)
        return fail(MQE_ERR_UNSUPPORTED, "substep kernel working set exceeds shared memory for this A/P");

    // ---- constants ----
    CK(dupload(s, &p.sdf, d->h_sdf, (size_t)d->sdf_nx * d->sdf_ny));
    CK(dupload(s, &p.env_origins, d->h_env_origins, (size_t)N * 3));
    CK(dupload(s, &p.agent_origins, d->h_agent_origins, (size_t)M * 3));
    CK(dupload(s, &p.base_init, d->h_base_init_state, (size_t)M * 13));
    if (P) CK(dupload(s, &p.npc_init, d->h_npc_init_state, (size_t)N * P * 13));
    if (d->h_env_friction) CK(dupload(s, &p.mu_env, d->h_env_friction, (size_t)N));
    if (d->h_base_added_mass) CK(dupload(s, &p.base_mass_add, d->h_base_added_mass, (size_t)M));
    if (d->h_base_com_shift) CK(dupload(s, &p.base_com_shift, d->h_base_com_shift, (size_t)M * 3));
    if (d->h_motor_strength) CK(dupload(s, &p.motor_strength, d->h_motor_strength, (size_t)M * 12));
    {
        std::vector<float> nd(D > 0 ? D : 1, 0.f);
        if (D && d->h_npc_dof_default) memcpy(nd.data(), d->h_npc_dof_default, D * sizeof(float));
        CK(dupload(s, &p.npc_dof_default, nd.data(), nd.size()));
        CK(cudaStreamSynchronize(s->stream));
    }
    CK(dupload(s, &p.model, &d->model, 1));
    CK(dupload(s, &p.loc_default, d->loc_obs_default, MQE_LOC_OBS));
    const MqeWeights &w = d->weights;
    for (const float *const *pp = (const float *const *)&w; pp < (const float *const *)(&w + 1); pp++)
        if (!*pp) return fail(MQE_ERR_INVALID, "weights pointer missing");
    {
        std::vector<float> aw(1316, 0.f);
        memcpy(aw.data(), w.act_w0, 192 * 4); memcpy(aw.data() + 192, w.act_b0, 32 * 4);
        for (int o = 0; o < 32; o++)                                   // W1 transposed: [in][out], see k_substeps P1
            for (int i = 0; i < 32; i++) aw[224 + i * 32 + o] = w.act_w1[o * 32 + i];
        memcpy(aw.data() + 1248, w.act_b1, 32 * 4);
        memcpy(aw.data() + 1280, w.act_w2, 32 * 4); aw[1312] = w.act_b2[0];
        CK(dupload(s, &p.act_w, aw.data(), aw.size()));
        // CTA header of k_substeps: model, actuator weights, and which lane of a robot's quad owns which probe / capsule (leg
        // links belong to their leg, base colliders are dealt round-robin (probes) or to leg 0 (capsules); lists keep table order)
        {
            const size_t mf = sizeof(MqeRobotModel) / 4;
            const size_t ACTW = 1380;                                    // physics.cu ACTW_FLOATS
            std::vector<float> hdr(mf + ACTW + 80, 0.f);
            memcpy(hdr.data(), &d->model, sizeof(MqeRobotModel));
            { const char *e = getenv("MQE_ACT_MMA"); p.act_mma = (e && e[0] == '0') ? 0 : 1; }
            if (!p.act_mma) memcpy(hdr.data() + mf, aw.data(), 1316 * sizeof(float));
            else {
                // fp16 hi / lo fragment table of the actuator net for mma.sync (k_substeps P1): B operands in the per-lane register
                // order of m16n8k8 (layer 0, K = 6 padded to 8) and m16n8k16 (layer 1), then b0, b1, W2, b2 as floats
                auto h16 = [](float v) { return (uint32_t)__half_as_ushort(__float2half_rn(v)); };
                auto hi_lo = [&](float e0, float e1, uint32_t &hi, uint32_t &lo) {
                    const uint32_t h0 = h16(e0), h1 = h16(e1);
                    hi = h0 | (h1 << 16);
                    const float r0 = e0 - __half2float(__ushort_as_half((unsigned short)h0)), r1 = e1 - __half2float(__ushort_as_half((unsigned short)h1));
                    lo = h16(r0) | (h16(r1) << 16);
                };
                uint32_t *fr = reinterpret_cast<uint32_t *>(hdr.data() + mf);
                auto W0 = [&](int n, int k) { return k < 6 ? w.act_w0[n * 6 + k] : 0.f; };
                auto W1 = [&](int n, int k) { return w.act_w1[n * 32 + k]; };
                for (int lane = 0; lane < 32; lane++) {
                    const int g = lane >> 2, t = lane & 3;
                    for (int j = 0; j < 4; j++) {
                        uint32_t hi, lo;
                        hi_lo(W0(8 * j + g, 2 * t), W0(8 * j + g, 2 * t + 1), hi, lo);
                        fr[(0 * 32 + lane) * 4 + j] = hi; fr[(1 * 32 + lane) * 4 + j] = lo;
                        for (int s2 = 0; s2 < 2; s2++) {
                            uint32_t h0, l0, h1, l1;
                            hi_lo(W1(8 * j + g, 16 * s2 + 2 * t), W1(8 * j + g, 16 * s2 + 2 * t + 1), h0, l0);
                            hi_lo(W1(8 * j + g, 16 * s2 + 2 * t + 8), W1(8 * j + g, 16 * s2 + 2 * t + 9), h1, l1);
                            uint32_t *q = fr + ((2 + j * 2 + s2) * 32 + lane) * 4;
                            q[0] = h0; q[1] = h1; q[2] = l0; q[3] = l1;
                        }
                    }
                }
                float *ft = hdr.data() + mf + 1280;
                memcpy(ft, w.act_b0, 32 * 4); memcpy(ft + 32, w.act_b1, 32 * 4); memcpy(ft + 64, w.act_w2, 32 * 4); ft[96] = w.act_b2[0];
            }
            int *tbl = reinterpret_cast<int *>(hdr.data() + mf + ACTW);
            for (int lg = 0; lg < 4; lg++) {
                int n = 0, nbase = 0;
                for (int pi = 0; pi < d->model.n_probes; pi++) {
                    const int link = (int)d->model.probes[pi][0];
                    const bool mine = link == 0 ? ((nbase++ & 3) == lg) : ((link - 1) / 3 == lg);
                    if (mine && n < 9) tbl[lg * 10 + 1 + n++] = pi;
                }
                tbl[lg * 10] = n;
                n = 0;
                for (int ci = 0; ci < d->model.n_caps; ci++) {
                    const int link = (int)d->model.caps[ci][0];
                    const bool mine = link == 0 ? (lg == 0) : ((link - 1) / 3 == lg);
                    if (mine && n < 9) tbl[40 + lg * 10 + 1 + n++] = ci;
                }
                tbl[40 + lg * 10] = n;
            }
            CK(dupload(s, &p.substep_hdr, hdr.data(), hdr.size()));
            CK(cudaStreamSynchronize(s->stream));
        }
        // layer 0 of both networks, age-blocked and padded: [768][30][80]
        const int RR = MQE_HIST_FRAMES * MQE_HIST_PAD;
        std::vector<float> w0((size_t)768 * RR, 0.f), b0(768), wl(512 * 2);
        for (int n = 0; n < 768; n++) {
            const float *src = n < 256 ? w.adapt_w0 + (size_t)n * 2100 : w.body_w0 + (size_t)(n - 256) * 2102;
            for (int b = 0; b < MQE_HIST_FRAMES; b++)
                memcpy(&w0[(size_t)n * RR + b * MQE_HIST_PAD], src + b * MQE_LOC_OBS, MQE_LOC_OBS * sizeof(float));
            b0[n] = n < 256 ? w.adapt_b0[n] : w.body_b0[n - 256];
            if (n >= 256) { wl[(n - 256) * 2] = src[2100]; wl[(n - 256) * 2 + 1] = src[2101]; }
        }
        CK(dupload(s, &s->pw.w0cat, w0.data(), w0.size()));
        CK(dupload(s, &s->pw.b0cat, b0.data(), b0.size()));
        CK(dupload(s, &s->pw.wlat, wl.data(), wl.size()));
        CK(dupload(s, &s->pw.aw1, w.adapt_w1, 128 * 256)); CK(dupload(s, &s->pw.ab1, w.adapt_b1, 128));
        CK(dupload(s, &s->pw.aw2, w.adapt_w2, 2 * 128)); CK(dupload(s, &s->pw.ab2, w.adapt_b2, 2));
        CK(dupload(s, &s->pw.bw1, w.body_w1, 256 * 512)); CK(dupload(s, &s->pw.bb1, w.body_b1, 256));
        CK(dupload(s, &s->pw.bw2, w.body_w2, 128 * 256)); CK(dupload(s, &s->pw.bb2, w.body_b2, 128));
        CK(dupload(s, &s->pw.bw3, w.body_w3, 12 * 128)); CK(dupload(s, &s->pw.bb3, w.body_b3, 12));
        CK(cudaStreamSynchronize(s->stream));     // host staging vectors go out of scope
    }
    { const char *e = getenv("MQE_TC_TAIL"); s->tail_fp32 = e && e[0] == '0'; }
    { const char *e = getenv("MQE_POLICY_FORK"); s->fork_policy = (p.policy_mode != MQE_POLICY_FP32) && !s->tail_fp32 && !(e && e[0] == '0'); }
    { const char *e = getenv("MQE_POLICY_FUSED"); s->fused_policy = (p.policy_mode != MQE_POLICY_FP32) && !s->tail_fp32 && !(e && e[0] == '0'); }
    if (s->fused_policy) s->fork_policy = false;
    { const char *e = getenv("MQE_FUSE_POST"); s->fuse_post = !(e && e[0] == '0'); }
    { const char *e = getenv("MQE_POLICY_INCR"); s->incremental = s->fused_policy && !(e && e[0] == '0'); }
    { const char *e = getenv("MQE_L0_EARLY_TILES"); s->early_tiles = s->incremental ? (e ? atoi(e) : 0) : 0; }     // measured: 0.400 ms/step with 0, 0.406 with 8 or 14, 0.431 with 28 (the early CTAs slow the tail they run beside)
    if (s->incremental) CK(cudaEventCreateWithFlags(&s->ev_early, cudaEventDisableTiming));
    if (s->fork_policy || s->incremental) {
        CK(cudaStreamCreateWithFlags(&s->aux_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
    }
    { const char *e = getenv("MQE_GRAPH"); s->use_graph = (p.policy_mode != MQE_POLICY_FP32) && !(e && e[0] == '0'); }
    if (p.policy_mode != MQE_POLICY_FP32) {
        int rc = mqe_policy_tc_prepare(&w, M, &s->tcw, s->stream);
        if (rc != 0) return fail(MQE_ERR_CUDA, "tensor-core policy weight preparation failed");
    }
    // ---- state ----
    CK(dalloc(s, &p.root, (size_t)N * G * 13)); CK(dalloc(s, &p.dof, (size_t)N * (12 * A + D) * 2));
    CK(dalloc(s, &p.contact, (size_t)N * p.NB * 3));
    CK(dalloc(s, &p.torques, (size_t)M * 12)); CK(dalloc(s, &p.actions, (size_t)M * 12)); CK(dalloc(s, &p.last_actions, (size_t)M * 12));
    CK(dalloc(s, &p.loc_last, (size_t)M * 12)); CK(dalloc(s, &p.loc_last2, (size_t)M * 12)); CK(dalloc(s, &p.loc_obs, (size_t)M * MQE_LOC_OBS));
    CK(dalloc(s, &p.err1, (size_t)M * 12)); CK(dalloc(s, &p.err2, (size_t)M * 12)); CK(dalloc(s, &p.vel1, (size_t)M * 12)); CK(dalloc(s, &p.vel2, (size_t)M * 12));
    CK(dalloc(s, &p.gait, (size_t)M)); CK(dalloc(s, &p.clock, (size_t)M * 4));
    CK(dalloc(s, &p.base_quat, (size_t)M * 4)); CK(dalloc(s, &p.base_lin_vel, (size_t)M * 3)); CK(dalloc(s, &p.base_ang_vel, (size_t)M * 3));
    CK(dalloc(s, &p.proj_grav, (size_t)M * 3)); CK(dalloc(s, &p.obs, (size_t)M * MQE_OBS_FLOATS)); CK(dalloc(s, &p.commands, (size_t)M * 3));
    CK(dalloc(s, &p.last_dof_vel, (size_t)M * 12)); CK(dalloc(s, &p.last_root_vel, (size_t)M * 6)); CK(dalloc(s, &p.sheep_stats, (size_t)N * 3));
    CK(dalloc(s, &p.ep_len, (size_t)N));
    CK(dalloc(s, &p.reset_buf, (size_t)N)); CK(dalloc(s, &p.timeout_buf, (size_t)N)); CK(dalloc(s, &p.collide_buf, (size_t)N));
    CK(dalloc(s, &p.r_term, (size_t)N)); CK(dalloc(s, &p.p_term, (size_t)N)); CK(dalloc(s, &p.zl_term, (size_t)N)); CK(dalloc(s, &p.zh_term, (size_t)N));
    CK(dalloc(s, &p.episode, (size_t)N)); CK(dalloc(s, &p.hist_dirty, (size_t)N)); CK(dalloc(s, &p.stats, (size_t)8));
    CK(dalloc(s, &p.ctr, (size_t)8));
    CK(dalloc(s, &p.warp_trace, (size_t)((N + p.E - 1) / p.E) * MQE_TRACE_COLS));
    {
        const int ntasks = (N + p.E - 1) / p.E;
        CK(dalloc(s, &p.task_cost, (size_t)2 * ntasks));
        CK(dalloc(s, &s->d_task_order, (size_t)ntasks, false));
        std::vector<int> ident(ntasks);
        for (int i = 0; i < ntasks; i++) ident[i] = i;
        CK(cudaMemcpy(s->d_task_order, ident.data(), sizeof(int) * ntasks, cudaMemcpyHostToDevice));
        // MQE_BALANCE: 0 off, 1 on; default: on when the substep grid needs more than one round of CTAs (then the order also decides which
        // CTAs start first).  Needs the side stream.
        const char *e = getenv("MQE_BALANCE");
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
        const bool multi_round = ntasks > sms * 8;                 // at most 8 warps per SM are resident (255 registers per thread)
        s->balance = e ? atoi(e) : (multi_round ? 1 : 0);          // 1: grouped (multi-round grids), 2: spread (experiment for one-wave grids)
        if (p.control_type != 0) s->balance = 0;
        { const char *w = getenv("MQE_BALANCE_WHEN"); s->balance_at_end = !(w && w[0] == 's'); }
        if (s->balance) {
            CK(cudaStreamCreateWithFlags(&s->bal_stream, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&s->ev_bal0, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s->ev_bal1, cudaEventDisableTiming));
            if (s->balance != 3) p.task_order = s->d_task_order;     // 3: experiment -- sort and join, but keep the identity order
        }
    }
    { const char *e = getenv("MQE_CTA_SYNC"); p.cta_sync = e ? atoi(e) : 1; }
    { const char *e = getenv("MQE_TRACE"); p.trace = (e && e[0] == '1') ? 1 : 0; }
    if (d->lag_enabled) {
        if (d->lag_timesteps < 0 || d->lag_timesteps > 63) return fail(MQE_ERR_INVALID, "lag_timesteps out of range (0..63)");
        p.lag_n = d->lag_timesteps + 1;
        CK(dalloc(s, &p.lag_ring, (size_t)M * p.lag_n * 12));
    }
    CK(dalloc(s, &p.row_scratch, mqe_substeps_row_scratch_floats(N, A), false));
    CK(dalloc(s, &p.prow_scratch, mqe_substeps_prow_scratch_floats(N, s->maxpair), false));
    {
        const int Gd = A + (d->npc_kind == MQE_NPC_RIGID ? P : 0), nc = d->model.n_caps;    // capsule groups only
        int total = 0;
        for (int X = 0; X < Gd; X++)
            for (int Y = X + 1; Y < Gd; Y++) total += (X < A ? nc : 1) * (Y < A ? nc : 1);
        p.max_cand = total > 0 ? total : 1;
        CK(dalloc(s, &p.cand_scratch, (size_t)N * p.max_cand, false));
    }
    CK(dalloc(s, &p.pdesc_scratch, mqe_substeps_pdesc_scratch_floats(N, s->maxpair), false));
    const size_t ring = (size_t)M * MQE_HIST_FRAMES * MQE_HIST_PAD;
    if (p.policy_mode == MQE_POLICY_FP32) CK(dalloc(s, &p.hist_f32, ring));      // tensor-core modes: the bf16 hi / lo planes ARE the ring
    const size_t ring_tc = (size_t)((M + 127) / 128) * 128 * MQE_HIST_FRAMES * MQE_HIST_PAD;     // pre-tiled planes, rows padded to 128
    if (p.policy_mode != MQE_POLICY_FP32) { CK(dalloc(s, &p.hist_hi, ring_tc)); CK(dalloc(s, &p.hist_lo, ring_tc)); }
    // Z: whole row tiles (the fused tail reads it tile-major)
    CK(dalloc(s, &s->ps.Z, (size_t)((M + 127) / 128) * 128 * 768)); CK(dalloc(s, &s->ps.T1, (size_t)M * 128)); CK(dalloc(s, &s->ps.T2, (size_t)M * 256));
    CK(dalloc(s, &s->ps.T3, (size_t)M * 128)); CK(dalloc(s, &s->ps.latent, (size_t)M * 2)); CK(dalloc(s, &s->ps.act, (size_t)M * 12));
    if (s->incremental) CK(dalloc(s, &s->ps.Zold, (size_t)((M + 127) / 128) * 128 * 768));      // whole 128-row tiles (tile-major layout, policy_tc.cu zold_index)
    CK(dalloc(s, &s->d_actions_stage, s->action_bytes() / sizeof(float)));
    CK(cudaMemsetAsync(p.reset_buf, 1, N, s->stream));                      // base_task.py:77: reset_buf starts as ones
    // _prepare_locomotion_policy: locomotion_obs = default command frame (go1.py:393-394); actors at their start poses
    {
        std::vector<float> lo((size_t)M * MQE_LOC_OBS), root((size_t)N * G * 13, 0.f), dof((size_t)N * (12 * A + D) * 2, 0.f);
        for (int m = 0; m < M; m++) memcpy(&lo[(size_t)m * MQE_LOC_OBS], d->loc_obs_default, MQE_LOC_OBS * sizeof(float));
        for (int e = 0; e < N; e++) {
            for (int a = 0; a < A; a++) {
                float *r = &root[((size_t)e * G + a) * 13];
                for (int i = 0; i < 13; i++) r[i] = d->h_base_init_state[(e * A + a) * 13 + i];
                for (int i = 0; i < 3; i++) r[i] += d->h_agent_origins[(e * A + a) * 3 + i];
                for (int j = 0; j < 12; j++) dof[((size_t)e * (12 * A + D) + 12 * a + j) * 2] = d->model.q_default[j];
            }
            for (int n = 0; n < P; n++) {
                float *r = &root[((size_t)e * G + A + n) * 13];
                for (int i = 0; i < 13; i++) r[i] = d->h_npc_init_state[(e * P + n) * 13 + i];
                for (int i = 0; i < 3; i++) r[i] += d->h_env_origins[e * 3 + i];
            }
        }
        // _init_buffers (legged_robot.py:570, 620-622): base_quat is the spawn quaternion, base_lin_vel / base_ang_vel / projected_gravity
        // are derived from the spawn state, so the observation of the very first reset() already carries gravity (0, 0, -1) in the base
        // frame (reset_idx does not recompute them; post_physics_step does, after the first physics step)
        std::vector<float> bq((size_t)M * 4), blv((size_t)M * 3), bav((size_t)M * 3), pg((size_t)M * 3);
        for (int m = 0; m < M; m++) {
            const float *r = d->h_base_init_state + (size_t)m * 13, *q = r + 3;
            auto rot_inv = [&](const float *v, float *o) {       // isaacgym.torch_utils.quat_rotate_inverse
                const float w = q[3], k = 2.f * w * w - 1.f, dq = q[0] * v[0] + q[1] * v[1] + q[2] * v[2];
                const float c[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
                for (int i = 0; i < 3; i++) o[i] = k * v[i] - 2.f * w * c[i] + 2.f * dq * q[i];
            };
            const float g[3] = {0.f, 0.f, -1.f};
            for (int i = 0; i < 4; i++) bq[(size_t)m * 4 + i] = q[i];
            rot_inv(r + 7, &blv[(size_t)m * 3]); rot_inv(r + 10, &bav[(size_t)m * 3]); rot_inv(g, &pg[(size_t)m * 3]);
        }
        CK(cudaMemcpyAsync(p.base_quat, bq.data(), bq.size() * 4, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(p.base_lin_vel, blv.data(), blv.size() * 4, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(p.base_ang_vel, bav.data(), bav.size() * 4, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(p.proj_grav, pg.data(), pg.size() * 4, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(p.loc_obs, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(p.root, root.data(), root.size() * 4, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(p.dof, dof.data(), dof.size() * 4, cudaMemcpyHostToDevice, s->stream));
        CK(cudaStreamSynchronize(s->stream));
    }
    CK(cudaMallocHost(&s->h_actions, s->action_bytes()));
    CK(cudaMallocHost(&s->h_obs, (size_t)M * MQE_OBS_FLOATS * sizeof(float)));
    CK(cudaMallocHost(&s->h_reset, (size_t)N));
    CK(alloc_step_result(s, 0, 0));
    CK(mqe_substeps_configure(p, s->maxpair));               // per-device function attribute of k_substeps

    set_buf(s, MQE_BUF_ROOT_STATES, p.root, 4, N, G, 13);
    set_buf(s, MQE_BUF_DOF_STATES, p.dof, 4, N, 12 * A + D, 2);
    set_buf(s, MQE_BUF_CONTACT_FORCES, p.contact, 4, N, p.NB, 3);
    set_buf(s, MQE_BUF_TORQUES, p.torques, 4, N, 12 * A);
    set_buf(s, MQE_BUF_ACTIONS, p.actions, 4, N, 12 * A);
    set_buf(s, MQE_BUF_LAST_ACTIONS, p.last_actions, 4, N, 12 * A);
    set_buf(s, MQE_BUF_OBS, p.obs, 4, M, MQE_OBS_FLOATS);
    set_buf(s, MQE_BUF_BASE_LIN_VEL, p.base_lin_vel, 4, M, 3);
    set_buf(s, MQE_BUF_BASE_ANG_VEL, p.base_ang_vel, 4, M, 3);
    set_buf(s, MQE_BUF_PROJ_GRAVITY, p.proj_grav, 4, M, 3);
    set_buf(s, MQE_BUF_RESET, p.reset_buf, 1, N);
    set_buf(s, MQE_BUF_TIMEOUT, p.timeout_buf, 1, N);
    set_buf(s, MQE_BUF_COLLIDE, p.collide_buf, 1, N);
    set_buf(s, MQE_BUF_ROLL_TERM, p.r_term, 1, N);
    set_buf(s, MQE_BUF_PITCH_TERM, p.p_term, 1, N);
    set_buf(s, MQE_BUF_ZLOW_TERM, p.zl_term, 1, N);
    set_buf(s, MQE_BUF_ZHIGH_TERM, p.zh_term, 1, N);
    set_buf(s, MQE_BUF_EPISODE_LENGTH, p.ep_len, 8, N);
    set_buf(s, MQE_BUF_COMMANDS, p.commands, 4, M, 3);
    set_buf(s, MQE_BUF_LOC_OBS, p.loc_obs, 4, M, MQE_LOC_OBS);
    set_buf(s, MQE_BUF_LOC_ACTION, p.loc_last, 4, M, 12);
    set_buf(s, MQE_BUF_GAIT, p.gait, 4, M);
    set_buf(s, MQE_BUF_HISTORY, p.hist_f32, 4, M, MQE_HIST_FRAMES, MQE_HIST_PAD);      // null outside MQE_POLICY_FP32
    set_buf(s, MQE_BUF_HISTORY_HI, p.hist_hi, 2, (M + 127) / 128, MQE_HIST_FRAMES, MQE_HIST_PAD / 8, 128 * 8);
    set_buf(s, MQE_BUF_HISTORY_LO, p.hist_lo, 2, (M + 127) / 128, MQE_HIST_FRAMES, MQE_HIST_PAD / 8, 128 * 8);
    set_buf(s, MQE_BUF_SHEEP_STATS, p.sheep_stats, 4, N, 3);
    set_buf(s, MQE_BUF_STATS, p.stats, 4, 8);
    set_buf(s, MQE_BUF_CLOCK, p.clock, 4, M, 4);
    set_buf(s, MQE_BUF_WARP_TRACE, p.warp_trace, 8, (N + p.E - 1) / p.E, MQE_TRACE_COLS);
    return MQE_OK;
}

int mqe_sim_create(const MqeSimDesc *desc, int device, void *stream, MqeSim **out) {
    if (!desc || !out) return fail(MQE_ERR_INVALID, "null argument");
    if (desc->abi_version != MQE_ABI_VERSION) return fail(MQE_ERR_INVALID, "descriptor abi_version mismatch");
    MqeSim *s = new MqeSim();
    int rc = create_impl(desc, device, stream, s);
    if (rc != MQE_OK) { std::string keep = g_err; mqe_sim_destroy(s); g_err = keep; *out = nullptr; return rc; }
    *out = s;
    return MQE_OK;
}

int mqe_sim_set_stream(MqeSim *s, void *stream) {
    if (!s) return fail(MQE_ERR_INVALID, "null handle");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    s->stream = (cudaStream_t)stream;
    return MQE_OK;
}

int mqe_sim_set_action_scale(MqeSim *s, const float scale[3]) {
    if (!s || !scale) return fail(MQE_ERR_INVALID, "null argument");
    for (int i = 0; i < 3; i++) s->p.act_scale[i] = scale[i];
    return MQE_OK;
}

int mqe_sim_set_wrapper(MqeSim *s, const MqeWrapperDesc *d) {
    if (!s || !d) return fail(MQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    drop_graphs(s);                                              // the step graph changes shape
    if (s->gather.world) return fail(MQE_ERR_INVALID, "mqe_sim_set_wrapper must precede mqe_sim_gather_init");
    const DevParams &p = s->p;
    WrapParams w = {};
    w.kind = d->kind;
    if (d->kind == MQE_WRAP_NONE) { s->wrap = w; CK(alloc_step_result(s, 0, 0)); return MQE_OK; }
    int gate_cols = 0;
    if (d->kind == MQE_WRAP_SHEEP) {
        if (p.npc_ctrl != MQE_NPC_SHEEP || p.P < 1 || p.P > 12) return fail(MQE_ERR_INVALID, "sheep wrapper needs 1..12 sheep");
        w.Aw = p.A; w.D = 14 + 2 * p.P + p.A; gate_cols = 2;
    } else if (d->kind == MQE_WRAP_SEESAW) {
        w.Aw = p.A; w.D = 12 + p.A;
    } else if (d->kind == MQE_WRAP_FOOTBALL_DEFENDER) {
        if (p.A < 2 || p.P < 1) return fail(MQE_ERR_INVALID, "football wrapper needs two agents and the ball");
        w.Aw = 2; w.D = 20; gate_cols = 3;
    } else if (d->kind == MQE_WRAP_PUSHBOX) {
        if (p.P < 1) return fail(MQE_ERR_INVALID, "pushbox wrapper needs the box");
        w.Aw = p.A; w.D = 20 + p.A; gate_cols = 2;
    } else if (d->kind == MQE_WRAP_WRESTLING || d->kind == MQE_WRAP_BRIDGE || d->kind == MQE_WRAP_ROTATION) {
        if (p.A != 2) return fail(MQE_ERR_INVALID, "duel wrappers need exactly two agents");
        w.Aw = 2; w.D = 12;
    } else return fail(MQE_ERR_INVALID, "unknown wrapper kind");
    if (w.Aw > 4) return fail(MQE_ERR_INVALID, "wrapper: at most 4 agents");
    for (int i = 0; i < 8; i++) w.scale[i] = d->scale[i];
    if (gate_cols) {
        if (!d->h_gate) return fail(MQE_ERR_INVALID, "h_gate missing");
        CK(dupload(s, &w.gate, d->h_gate, (size_t)p.N * gate_cols));
        CK(cudaStreamSynchronize(s->stream));
    }
    CK(alloc_step_result(s, w.Aw, w.D));                         // obs and reward live inside the packed step result
    w.obs = reinterpret_cast<float *>(s->d_result + s->rl.obs_off); w.reward = reinterpret_cast<float *>(s->d_result + s->rl.reward_off);
    CK(dalloc(s, &w.sums, (size_t)16)); CK(dalloc(s, &w.last, (size_t)p.N * 4));
    CK(dalloc(s, &w.delayed_reset, (size_t)p.N)); CK(dalloc(s, &w.has_last, (size_t)p.N));
    s->wrap = w;
    set_buf(s, MQE_BUF_WRAP_SUMS, w.sums, 8, 16);
    return MQE_OK;
}

int mqe_sim_wrapper_reset(MqeSim *s) {
    if (!s) return fail(MQE_ERR_INVALID, "null handle");
    if (s->wrap.kind == MQE_WRAP_NONE) return fail(MQE_ERR_INVALID, "no task wrapper set");
    CK(cudaSetDevice(s->device));
    CK(mqe_launch_task_gather(s->p, s->wrap, 1, s->stream));
    s->launches += 1;
    return MQE_OK;
}

int mqe_sim_get_buffer(MqeSim *s, int which, void **d_ptr, int64_t shape[4], int32_t *elem_size) {
    if (!s || which < 0 || which >= MQE_BUF_COUNT) return fail(MQE_ERR_INVALID, "bad buffer id");
    if (!s->bufs[which].ptr && which >= MQE_BUF_SUBSTEP_TORQUES && which <= MQE_BUF_SUBSTEP_EXCEED) {
        // post_decimation_step logs (legged_robot.py:112-115): first request switches them on for every following step
        CK(cudaSetDevice(s->device));
        CK(cudaStreamSynchronize(s->stream));
        DevParams &p = s->p;
        const size_t n = (size_t)p.N * p.decimation * 12 * p.A;
        CK(dalloc(s, &p.sub_tau, n)); CK(dalloc(s, &p.sub_qd, n)); CK(dalloc(s, &p.sub_exceed, n));
        set_buf(s, MQE_BUF_SUBSTEP_TORQUES, p.sub_tau, 4, p.N, p.decimation, 12 * p.A);
        set_buf(s, MQE_BUF_SUBSTEP_DOF_VEL, p.sub_qd, 4, p.N, p.decimation, 12 * p.A);
        set_buf(s, MQE_BUF_SUBSTEP_EXCEED, p.sub_exceed, 1, p.N, p.decimation, 12 * p.A);
        drop_graphs(s);                                            // kernel arguments changed
    }
    if (!s->bufs[which].ptr) return fail(MQE_ERR_INVALID, "buffer not allocated (task-wrapper buffers exist after mqe_sim_set_wrapper; MQE_BUF_HISTORY only with MQE_POLICY_FP32, the bf16 planes otherwise)");
    if (d_ptr) *d_ptr = s->bufs[which].ptr;
    if (shape) for (int i = 0; i < 4; i++) shape[i] = s->bufs[which].shape[i];
    if (elem_size) *elem_size = s->bufs[which].elem;
    return MQE_OK;
}

int mqe_sim_history_head(MqeSim *s) { return s ? s->head : -1; }

int mqe_sim_reset(MqeSim *s) {
    if (!s) return fail(MQE_ERR_INVALID, "null handle");
    CK(cudaSetDevice(s->device));
    CK(mqe_launch_reset_all(s->p, s->stream));
    s->launches += 1;
    if (s->wrap.kind != MQE_WRAP_NONE) { CK(mqe_launch_task_gather(s->p, s->wrap, 1, s->stream)); s->launches += 1; }
    if (s->gather.world) s->gather_seq++;
    return exchange_impl(s);                                      // ranks reset together: the global observation is exchanged as after a step
}

// finish: the simulation step (policy_impl) lets the fused tail kernel do k_policy_finish's work; returns 1 in *finished if it did
static int run_network(MqeSim *s, const float *ring, const unsigned short *hi, const unsigned short *lo, int head, int rows, float *latent, float *act,
                       int finish = 0, int *finished = nullptr, bool incr = false) {
    PolicyScratch ps = s->ps;
    ps.latent = latent; ps.act = act;
    if (finished) *finished = 0;
    if (s->incremental && incr) {                        // the 29 older frames are already in ps.Zold: only the new frame's K = 80 GEMM, then the tail
        int nf = 0;
        const int early = s->bg_early ? s->early_tiles : 0;
        const int head_next = head < 0 ? -2 : (head + 1) % MQE_HIST_FRAMES;
        CK(mqe_launch_policy_tc_incremental(s->tcw, s->pw, ps, s->p, hi, lo, head, rows, s->p.policy_mode == MQE_POLICY_BF16X3 ? 3 : 1, s->p.ctr, finish, s->stream, &nf,
                                            early, head_next, s->aux_stream, s->ev_early, s->bal_pending ? s->ev_bal1 : (cudaEvent_t) nullptr));
        s->bal_pending = 0;
        s->launches += nf;
        if (finished) *finished = finish;
        return MQE_OK;
    }
    if (s->fused_policy) {
        int nf = 0;
        CK(mqe_launch_policy_tc_fused(s->tcw, s->pw, ps, s->p, hi, lo, head, rows, s->p.policy_mode == MQE_POLICY_BF16X3 ? 3 : 1, s->p.ctr, finish, s->stream, &nf));
        s->launches += nf;
        if (finished) *finished = finish;
        return MQE_OK;
    }
    if (s->fork_policy) {
        int nf = 0;
        CK(mqe_launch_policy_tc_forked(s->tcw, s->pw, ps, hi, lo, head, rows, s->p.policy_mode == MQE_POLICY_BF16X3 ? 3 : 1, s->p.ctr, s->stream, s->aux_stream,
                                       s->ev_fork, s->ev_join, &nf));
        s->launches += nf;
        return MQE_OK;
    }
    if (s->p.policy_mode == MQE_POLICY_FP32)
        CK(mqe_launch_policy_l0_fp32(s->pw, ps, ring, head, rows, s->stream));
    else
        CK(mqe_launch_policy_l0_tc(s->tcw, s->pw.b0cat, hi, lo, head, rows, s->p.policy_mode == MQE_POLICY_BF16X3 ? 3 : 1, ps.Z, s->tail_fp32 ? 0 : 1, s->p.ctr, s->stream));
    int n = 1;
    if (s->p.policy_mode == MQE_POLICY_FP32 || s->tail_fp32)
        CK(mqe_launch_policy_tail(s->pw, ps, rows, s->stream, &n));
    else
        CK(mqe_launch_policy_tail_tc(s->tcw, s->pw, ps, rows, s->p.policy_mode == MQE_POLICY_BF16X3 ? 3 : 1, s->stream, &n));
    s->launches += n;
    return MQE_OK;
}

// device_ctr: kernels take the ring slot from the device counter (graph capture) instead of the host mirror
static int policy_impl(MqeSim *s, const float *d_actions, bool device_ctr) {
    if (!s || !d_actions) return fail(MQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(s->device));
    const int slot = (s->head + 1) % MQE_HIST_FRAMES;
    CK(mqe_launch_policy_frame(s->p, d_actions, device_ctr ? -1 : slot, s->stream));
    int finished = 0;
    if (s->incremental && !device_ctr && s->zold_head != slot) {     // no partial sums for this slot yet (first step, or after an explicit call): prime them now
        CK(mqe_launch_policy_l0_old(s->tcw, s->p.hist_hi, s->p.hist_lo, slot, s->M, s->p.policy_mode == MQE_POLICY_BF16X3 ? 3 : 1, s->ps.Zold, s->p.ctr, 0, -1, s->stream));
        s->launches += 1;
    }
    int rc = run_network(s, s->p.hist_f32, s->p.hist_hi, s->p.hist_lo, device_ctr ? -1 : slot, s->M, s->ps.latent, s->ps.act, 1, &finished, true);
    if (rc != MQE_OK) return rc;
    if (!finished) { CK(mqe_launch_policy_finish(s->p, s->ps.act, s->stream)); s->launches += 1; }
    s->launches += 1;                                    // k_policy_frame
    return MQE_OK;
}
// incremental layer 0: contract the 29 frames the NEXT step already knows (every slot but the one its frame will go to)
static int l0_old_impl(MqeSim *s, bool device_ctr, cudaStream_t st) {
    const int next_slot = (s->head + 2) % MQE_HIST_FRAMES;            // s->head still names the previous step's slot here
    const int first = s->bg_early ? s->early_tiles : 0;                // those row tiles were started beside the tail
    CK(mqe_launch_policy_l0_old(s->tcw, s->p.hist_hi, s->p.hist_lo, device_ctr ? -2 : next_slot, s->M, s->p.policy_mode == MQE_POLICY_BF16X3 ? 3 : 1,
                                s->ps.Zold, s->p.ctr, first, -1, st));
    s->launches += 1;
    s->zold_head = next_slot;
    return MQE_OK;
}
int mqe_sim_policy(MqeSim *s, const float *d_actions) {      // stand-alone call: the 29-frame pass runs in line (policy_impl primes it)
    if (s) s->bg_early = 0;
    int rc = policy_impl(s, d_actions, false);
    if (rc == MQE_OK) s->head = (s->head + 1) % MQE_HIST_FRAMES;
    return rc;
}

static int substeps_impl(MqeSim *s, int count, bool zero_stats) {
    if (!s || count <= 0) return fail(MQE_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(s->device));
    if (zero_stats) CK(cudaMemsetAsync(s->p.stats, 0, 5 * sizeof(int), s->stream));   // inside mqe_sim_step k_policy_finish did it
    CK(mqe_launch_substeps(s->p, count, s->maxpair, s->stream));
    s->launches += 1;
    return MQE_OK;
}
int mqe_sim_substeps(MqeSim *s, int count) { return substeps_impl(s, count, true); }

static int post_impl(MqeSim *s, bool device_ctr) {
    if (!s) return fail(MQE_ERR_INVALID, "null handle");
    CK(cudaSetDevice(s->device));
    CK(mqe_launch_post(s->p, device_ctr ? 0xffffffffu : s->step_count, s->stream));
    s->launches += 1;
    return MQE_OK;
}
int mqe_sim_post_physics(MqeSim *s) {
    int rc = post_impl(s, false);
    if (rc == MQE_OK) s->step_count++;
    return rc;
}

static int exchange_impl(MqeSim *s) {                    // peer exchange of the step result: last kernel of a step / reset
    if (!s->gather.world) return MQE_OK;
    static const bool debug_skip = getenv("MQE_DEBUG_NO_EXCHANGE") != nullptr;      // TIMING EXPERIMENT ONLY: what the exchange costs a sharded step
    if (debug_skip) return MQE_OK;
    CK(mqe_launch_gather_exchange(s->p, s->gather, s->stream));
    s->launches += 1;
    return MQE_OK;
}
static int step_plain(MqeSim *s, const float *d_actions, bool device_ctr) {
    int rc;
    // inside a capture (device_ctr) the mark has to become an event-record NODE that the host can query after the replay
#define STAGE_MARK(i) if (s->stage_timing) CK(device_ctr ? cudaEventRecordWithFlags(s->ev_stage[i], s->stream, cudaEventRecordExternal) : cudaEventRecord(s->ev_stage[i], s->stream))
    STAGE_MARK(0);
    if (s->balance && !s->balance_at_end) {              // next launch's task order, on the side stream beside the policy kernels
        CK(cudaEventRecord(s->ev_bal0, s->stream));
        CK(cudaStreamWaitEvent(s->bal_stream, s->ev_bal0, 0));
        CK(mqe_launch_balance_tasks(s->p, s->d_task_order, s->balance == 2, s->maxpair, s->bal_stream));
        s->launches += 1;
        CK(cudaEventRecord(s->ev_bal1, s->bal_stream));
        s->bal_pending = 1;
    }
    s->bg_early = (s->incremental && s->p.control_type == 0 && s->early_tiles > 0) ? 1 : 0;
    if (s->p.control_type == 0) rc = policy_impl(s, d_actions, device_ctr);
    else {                                               // 'P' / 'V' / 'T': the caller's joint actions are the actions (go1.py:43-45)
        CK(mqe_launch_joint_actions(s->p, d_actions, s->stream));
        s->launches += 1;
        rc = MQE_OK;
    }
    if (rc != MQE_OK) return rc;
    static const bool debug_no_bg = getenv("MQE_DEBUG_NO_BG") != nullptr;     // TIMING EXPERIMENT ONLY (results are wrong): how much of the step the background pass still costs
    const bool bg = s->incremental && s->p.control_type == 0 && !debug_no_bg;
    STAGE_MARK(1);
    if (bg) CK(cudaEventRecord(s->ev_fork, s->stream));   // fork point: the policy of this step is done, the ring holds its frame
    // no memset between kernels: k_policy_finish / k_joint_actions zeroed the statistics.  With fuse_post the physics kernel also
    // finishes the step (post_dev.cuh stages in its epilogue) and no k_post_physics launch follows.
    if (s->bal_pending) { CK(cudaStreamWaitEvent(s->stream, s->ev_bal1, 0)); s->bal_pending = 0; }     // policy path without the incremental launcher
    if (s->fuse_post) {                                  // the device step counter ctr[1] always equals s->step_count (every post pass bumps it)
        DevParams q = s->p;
        q.fuse_post = 1;
        CK(mqe_launch_substeps(q, s->p.decimation, s->maxpair, s->stream));
        s->launches += 1;
        rc = MQE_OK;
    } else rc = substeps_impl(s, s->p.decimation, false);
    if (rc != MQE_OK) return rc;
    STAGE_MARK(2);
    if (bg) {
        // next step's 29-frame layer-0 pass on the side stream, low priority, dependent on the fork point only.  It is enqueued AFTER
        // k_substeps on purpose: the one-wave physics grid must get its SMs first, the background pass takes what that grid leaves idle.
        CK(cudaStreamWaitEvent(s->aux_stream, s->ev_fork, 0));
        rc = l0_old_impl(s, device_ctr, s->aux_stream);
        if (rc != MQE_OK) return rc;
        CK(cudaEventRecord(s->ev_join, s->aux_stream));
    }
    if (!s->fuse_post) {
        rc = post_impl(s, device_ctr);
        if (rc != MQE_OK) return rc;
    }
    if (s->balance && s->balance_at_end) {
        // The NEXT step's task order, from the durations this step's k_substeps just wrote (its bookkeeping has advanced ctr[1]), on its own
        // side stream: it runs beside the task gather / the tail of the background layer-0 pass and is joined with that pass at the end
        // of the step, so that no kernel of the next step waits for it.
        CK(cudaEventRecord(s->ev_bal0, s->stream));
        CK(cudaStreamWaitEvent(s->bal_stream, s->ev_bal0, 0));
        CK(mqe_launch_balance_tasks(s->p, s->d_task_order, s->balance == 2, s->maxpair, s->bal_stream));
        s->launches += 1;
        CK(cudaEventRecord(s->ev_bal1, s->bal_stream));
        s->bal_pending = 1;
    }
    if (s->wrap.kind != MQE_WRAP_NONE) {
        CK(mqe_launch_task_gather(s->p, s->wrap, 0, s->stream));
        s->launches += 1;
    }
    rc = exchange_impl(s);
    STAGE_MARK(3);
    if (bg) CK(cudaStreamWaitEvent(s->stream, s->ev_join, 0));     // join: the ring must not move under the background pass
    if (s->bal_pending) { CK(cudaStreamWaitEvent(s->stream, s->ev_bal1, 0)); s->bal_pending = 0; }
    STAGE_MARK(4);
#undef STAGE_MARK
    s->bg_early = 0;
    return rc;
}

static int step_any(MqeSim *s, const float *d_actions);
int mqe_sim_step(MqeSim *s, const float *d_actions) {
    if (!s || !d_actions) return fail(MQE_ERR_INVALID, "null argument");
    if (s->p.control_type != 0) return fail(MQE_ERR_UNSUPPORTED, "control_type P / V / T takes [N][12A] joint actions: call mqe_sim_step_joint (go1.py:43-45)");
    return step_any(s, d_actions);
}
int mqe_sim_step_joint(MqeSim *s, const float *d_joint_actions) {
    if (!s || !d_joint_actions) return fail(MQE_ERR_INVALID, "null argument");
    if (s->p.control_type == 0) return fail(MQE_ERR_UNSUPPORTED, "control_type C takes [N][A][3] commands for the walk policy: call mqe_sim_step");
    return step_any(s, d_joint_actions);
}
static int step_any(MqeSim *s, const float *d_actions) {
    CK(cudaSetDevice(s->device));
    int rc;
    if (!s->use_graph || s->plain_steps < 2) {           // first steps run plainly (lazy function attributes, module load)
        rc = step_plain(s, d_actions, false);
        s->plain_steps++;
    } else {
        const size_t na = s->action_bytes();
        if (d_actions != s->d_actions_stage) CK(cudaMemcpyAsync(s->d_actions_stage, d_actions, na, cudaMemcpyDeviceToDevice, s->stream));
        MqeSim::StepGraph *g = nullptr;
        for (auto &c : s->graphs)
            if (c.scale[0] == s->p.act_scale[0] && c.scale[1] == s->p.act_scale[1] && c.scale[2] == s->p.act_scale[2]) g = &c;
        if (!g) {                                         // capture once per action scale; arguments are constant from here on
            if (!s->cap_stream) CK(cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking));
            cudaStream_t user = s->stream;
            const long long l0 = s->launches;
            CK(cudaStreamBeginCapture(s->cap_stream, cudaStreamCaptureModeThreadLocal));
            s->stream = s->cap_stream;
            const int keep_zold = s->zold_head;                   // capturing executes nothing: host mirrors must not move
            rc = step_plain(s, s->d_actions_stage, true);
            s->zold_head = keep_zold;
            s->stream = user;
            cudaGraph_t graph = nullptr;
            cudaError_t ce = cudaStreamEndCapture(s->cap_stream, &graph);
            if (rc != MQE_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(MQE_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
            MqeSim::StepGraph ng;
            for (int i = 0; i < 3; i++) ng.scale[i] = s->p.act_scale[i];
            ng.launches = (int)(s->launches - l0);
            s->launches = l0;
            ce = cudaGraphInstantiate(&ng.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return fail(MQE_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
            s->graphs.push_back(ng);
            g = &s->graphs.back();
        }
        if (s->incremental && s->p.control_type == 0 && s->zold_head != (s->head + 1) % MQE_HIST_FRAMES) {
            // the previous call was not a full step (e.g. mqe_sim_policy): the partial sums the graph expects do not exist yet
            const int slot = (s->head + 1) % MQE_HIST_FRAMES;
            CK(mqe_launch_policy_l0_old(s->tcw, s->p.hist_hi, s->p.hist_lo, slot, s->M, s->p.policy_mode == MQE_POLICY_BF16X3 ? 3 : 1, s->ps.Zold, s->p.ctr, 0, -1, s->stream));
            s->launches += 1;
        }
        CK(cudaGraphLaunch(g->exec, s->stream));
        s->launches += g->launches;
        rc = MQE_OK;
    }
    if (rc == MQE_OK) {
        s->head = (s->head + 1) % MQE_HIST_FRAMES; s->step_count++;
        if (s->gather.world) s->gather_seq++;
        if (s->incremental && s->p.control_type == 0) s->zold_head = (s->head + 1) % MQE_HIST_FRAMES;     // every step leaves the next step's partial sums behind
    }
    return rc;
}

int mqe_sim_pin_host(MqeSim *s, void *ptr, size_t bytes) {
    if (!s || !ptr || !bytes) return fail(MQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(s->device));
    if (s->is_pinned(ptr, bytes)) return MQE_OK;
    CK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    s->pinned.push_back({(char *)ptr, bytes});
    return MQE_OK;
}
int mqe_sim_unpin_host(MqeSim *s, void *ptr) {
    if (!s || !ptr) return fail(MQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(s->device));
    for (size_t i = 0; i < s->pinned.size(); i++)
        if (s->pinned[i].ptr == (char *)ptr) {
            CK(cudaStreamSynchronize(s->stream));
            CK(cudaHostUnregister(ptr));
            s->pinned.erase(s->pinned.begin() + i);
            return MQE_OK;
        }
    return fail(MQE_ERR_INVALID, "range was not pinned by this handle");
}

int mqe_sim_step_result_layout(MqeSim *s, MqeStepResultLayout *out) {
    if (!s || !out) return fail(MQE_ERR_INVALID, "null argument");
    *out = s->rl;
    return MQE_OK;
}

int mqe_sim_result_parity(MqeSim *s) { return s ? (int)(s->step_count & 1u) : -1; }

int mqe_sim_step_host_result(MqeSim *s, const float *h_actions, void *h_result) {
    if (!s || !h_actions || !h_result) return fail(MQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(s->device));
    const size_t na = s->action_bytes(), nr = (size_t)s->rl.total_bytes;
    const bool pa = s->is_pinned(h_actions, na), pr = s->is_pinned(h_result, nr);
    if (!pa) memcpy(s->h_actions, h_actions, na);
    CK(cudaMemcpyAsync(s->d_actions_stage, pa ? h_actions : s->h_actions, na, cudaMemcpyHostToDevice, s->stream));
    int rc = step_any(s, s->d_actions_stage);
    if (rc != MQE_OK) return rc;
    CK(cudaMemcpyAsync(pr ? h_result : (void *)s->h_result, s->d_result + (size_t)(s->step_count & 1u) * nr, nr, cudaMemcpyDeviceToHost, s->stream));     // ONE copy: obs | reward | done
    CK(cudaStreamSynchronize(s->stream));
    if (!pr) memcpy(h_result, s->h_result, nr);
    return MQE_OK;
}

int mqe_sim_step_host(MqeSim *s, const float *h_actions, float *h_obs, uint8_t *h_reset) {
    if (!s || !h_actions) return fail(MQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(s->device));
    const size_t na = s->action_bytes();
    const size_t no = (size_t)s->M * MQE_OBS_FLOATS * sizeof(float);
    // buffers the caller pinned (mqe_sim_pin_host) are DMA sources / targets themselves; anything else goes through the
    // handle's own pinned staging buffers with one extra host copy each way
    const bool pa = s->is_pinned(h_actions, na), po = h_obs && s->is_pinned(h_obs, no), pr = h_reset && s->is_pinned(h_reset, s->p.N);
    if (!pa) memcpy(s->h_actions, h_actions, na);
    CK(cudaMemcpyAsync(s->d_actions_stage, pa ? h_actions : s->h_actions, na, cudaMemcpyHostToDevice, s->stream));
    int rc = step_any(s, s->d_actions_stage);
    if (rc != MQE_OK) return rc;
    if (h_obs) CK(cudaMemcpyAsync(po ? h_obs : s->h_obs, s->p.obs, no, cudaMemcpyDeviceToHost, s->stream));
    if (h_reset) CK(cudaMemcpyAsync(pr ? h_reset : s->h_reset, s->p.reset_buf, s->p.N, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    if (h_obs && !po) memcpy(h_obs, s->h_obs, no);
    if (h_reset && !pr) memcpy(h_reset, s->h_reset, s->p.N);
    return MQE_OK;
}

// ---- peer exchange of the step result (gather.cu) ----
int mqe_sim_gather_init(MqeSim *s, int rank, int world, void *ipc_handle_out) {
    if (!s || !ipc_handle_out) return fail(MQE_ERR_INVALID, "null argument");
    if (world < 1 || world > MQE_MAX_RANKS || rank < 0 || rank >= world) return fail(MQE_ERR_INVALID, "rank / world out of range (world <= 16)");
    if (s->gather.world) return fail(MQE_ERR_INVALID, "gather already initialised");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    const MqeStepResultLayout &L = s->rl;
    if ((L.obs_bytes | L.reward_bytes | L.done_bytes) & 15) return fail(MQE_ERR_UNSUPPORTED, "peer exchange needs 16-byte multiples per field (num_envs per rank % 16 == 0): use an NCCL gather");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    drop_graphs(s);
    GatherParams g = {};
    g.rank = rank; g.world = world;
    g.seg_src[0] = L.obs_off; g.seg_bytes[0] = L.obs_bytes; g.seg_dst[0] = 0;
    g.seg_src[1] = L.reward_off; g.seg_bytes[1] = L.reward_bytes; g.seg_dst[1] = (long long)world * L.obs_bytes;
    g.seg_src[2] = L.done_off; g.seg_bytes[2] = L.done_bytes; g.seg_dst[2] = (long long)world * (L.obs_bytes + L.reward_bytes);
    g.parity_bytes = (long long)world * (L.obs_bytes + L.reward_bytes + L.done_bytes);
    g.flags_off = 2 * g.parity_bytes;
    g.src = s->d_result;
    const size_t total = (size_t)g.flags_off + 2 * MQE_MAX_RANKS * sizeof(unsigned int);
    unsigned char *buf = nullptr;
    CK(dalloc(s, &buf, total));                                   // a cudaMalloc allocation of its own: exportable
    CK(dalloc(s, &g.blocks_done, (size_t)1));
    CK(cudaStreamSynchronize(s->stream));
    g.peer[rank] = buf;
    s->gather_peers[rank] = buf;
    CK(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(ipc_handle_out), buf));
    s->gather = g;
    s->gather.world = 0;                                          // armed by mqe_sim_gather_connect
    s->gather.rank = rank;
    s->gather_pending_world = world;
    return MQE_OK;
}
int mqe_sim_gather_connect(MqeSim *s, const void *ipc_handles) {
    if (!s || !ipc_handles) return fail(MQE_ERR_INVALID, "null argument");
    if (!s->gather_pending_world) return fail(MQE_ERR_INVALID, "mqe_sim_gather_init first");
    CK(cudaSetDevice(s->device));
    const int world = s->gather_pending_world, rank = s->gather.rank;
    const cudaIpcMemHandle_t *h = reinterpret_cast<const cudaIpcMemHandle_t *>(ipc_handles);
    for (int r = 0; r < world; r++) {
        if (r == rank) continue;
        void *ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, h[r], cudaIpcMemLazyEnablePeerAccess));
        s->gather_peers[r] = ptr;
        s->gather.peer[r] = (unsigned char *)ptr;
    }
    s->gather.world = world;
    s->gather_pending_world = 0;
    drop_graphs(s);
    return MQE_OK;
}
int mqe_sim_gather_parity(MqeSim *s) { return (s && s->gather.world) ? (int)(s->gather_seq & 1ull) : -1; }
int mqe_sim_gather_view(MqeSim *s, int parity, void **d_ptr, MqeStepResultLayout *gl) {
    if (!s || !s->gather.world) return fail(MQE_ERR_INVALID, "peer exchange not connected");
    if (parity < 0 || parity > 1) return fail(MQE_ERR_INVALID, "parity must be 0 or 1");
    const GatherParams &g = s->gather;
    if (d_ptr) *d_ptr = g.peer[g.rank] + (long long)parity * g.parity_bytes;
    if (gl) {
        const MqeStepResultLayout &L = s->rl;
        gl->num_envs = L.num_envs * g.world; gl->Aw = L.Aw; gl->D = L.D; gl->reserved = 0;
        gl->obs_off = g.seg_dst[0]; gl->obs_bytes = L.obs_bytes * g.world;
        gl->reward_off = g.seg_dst[1]; gl->reward_bytes = L.reward_bytes * g.world;
        gl->done_off = g.seg_dst[2]; gl->done_bytes = L.done_bytes * g.world;
        gl->total_bytes = g.parity_bytes;
    }
    return MQE_OK;
}

int mqe_sim_set_root_indexed(MqeSim *s, const float *d_root_states, const int32_t *d_actor_ids, int n) {
    if (!s || !d_root_states || !d_actor_ids) return fail(MQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(s->device));
    if (d_root_states == s->p.root) return MQE_OK;         // the wrapped view IS the engine state
    CK(mqe_launch_set_root_indexed(s->p, d_root_states, d_actor_ids, n, s->stream));
    s->launches += 1;
    return MQE_OK;
}
int mqe_sim_set_dof_indexed(MqeSim *s, const float *d_dof_states, const int32_t *d_actor_ids, int n) {
    if (!s || !d_dof_states || !d_actor_ids) return fail(MQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(s->device));
    if (d_dof_states == s->p.dof) return MQE_OK;
    CK(mqe_launch_set_dof_indexed(s->p, d_dof_states, d_actor_ids, n, s->stream));
    s->launches += 1;
    return MQE_OK;
}

int mqe_policy_forward(MqeSim *s, const float *d_history, int rows, float *d_latent, float *d_action) {
    if (!s || !d_history || !d_action || rows <= 0) return fail(MQE_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(s->device));
    if (rows > s->M) return fail(MQE_ERR_INVALID, "rows exceeds num_envs*num_agents (scratch is sized for the simulation)");
    const size_t ring = (size_t)rows * MQE_HIST_FRAMES * MQE_HIST_PAD;
    if (rows > s->tmp_rows) {
        CK(cudaStreamSynchronize(s->stream));
        if (s->tmp_ring) cudaFree(s->tmp_ring);
        if (s->tmp_hi) cudaFree(s->tmp_hi);
        if (s->tmp_lo) cudaFree(s->tmp_lo);
        s->tmp_ring = nullptr; s->tmp_hi = s->tmp_lo = nullptr; s->tmp_rows = 0;
        if (s->p.policy_mode == MQE_POLICY_FP32) CK(cudaMalloc(&s->tmp_ring, ring * sizeof(float)));
        if (s->p.policy_mode != MQE_POLICY_FP32) {
            const size_t rtc = (size_t)((rows + 127) / 128) * 128 * MQE_HIST_FRAMES * MQE_HIST_PAD * 2;
            CK(cudaMalloc(&s->tmp_hi, rtc)); CK(cudaMalloc(&s->tmp_lo, rtc));
            CK(cudaMemsetAsync(s->tmp_hi, 0, rtc, s->stream)); CK(cudaMemsetAsync(s->tmp_lo, 0, rtc, s->stream));
        }
        s->tmp_rows = rows;
    }
    CK(mqe_launch_history_to_ring(d_history, s->tmp_ring, s->tmp_hi, s->tmp_lo, rows, s->stream));
    s->launches += 1;
    float *lat = d_latent ? d_latent : s->ps.latent;
    return run_network(s, s->tmp_ring, s->tmp_hi, s->tmp_lo, MQE_HIST_FRAMES - 1, rows, lat, d_action);
}

int mqe_actuator_forward(MqeSim *s, const float *d_x, int rows, float *d_torque) {
    if (!s || !d_x || !d_torque || rows <= 0) return fail(MQE_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(s->device));
    CK(mqe_launch_actuator(s->p.act_w, d_x, rows, d_torque, s->stream));
    s->launches += 1;
    return MQE_OK;
}

int mqe_sim_synchronize(MqeSim *s) {
    if (!s) return fail(MQE_ERR_INVALID, "null handle");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    return MQE_OK;
}
int64_t mqe_sim_launch_count(MqeSim *s) { return s ? s->launches : 0; }
int mqe_sim_stage_timing(MqeSim *s, int enable) {
    if (!s) return fail(MQE_ERR_INVALID, "null handle");
    CK(cudaSetDevice(s->device));
    if (enable && !s->ev_stage[0]) for (auto &e : s->ev_stage) CK(cudaEventCreate(&e));
    if ((enable != 0) != (s->stage_timing != 0)) { CK(cudaStreamSynchronize(s->stream)); drop_graphs(s); }      // the marks are nodes of the step graph
    s->stage_timing = enable ? 1 : 0;
    return MQE_OK;
}
int mqe_sim_stage_ms(MqeSim *s, float *ms4) {
    if (!s || !ms4) return fail(MQE_ERR_INVALID, "null argument");
    if (!s->stage_timing) return fail(MQE_ERR_UNSUPPORTED, "mqe_sim_stage_timing(sim, 1) first");
    CK(cudaSetDevice(s->device));
    CK(cudaEventSynchronize(s->ev_stage[4]));
    for (int i = 0; i < 4; i++) CK(cudaEventElapsedTime(ms4 + i, s->ev_stage[i], s->ev_stage[i + 1]));
    return MQE_OK;
}

}  // extern "C"
