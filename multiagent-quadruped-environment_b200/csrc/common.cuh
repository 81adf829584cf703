// common.cuh -- device-side parameter block, math helpers and the counter-based RNG shared by the kernels.
// sm_100a only.  See DESIGN.md for the data layout and the algorithm; reference citations are relative to
// /root/reference (ziyanx02/multiagent-quadruped-environment).
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <stdint.h>

#include "../../include/mqe_b200.h"

#define MQE_MAX_LOCAL 8          // world contacts kept per robot / npc per substep
#define MQE_MAX_LIMIT 4          // active joint-limit rows per robot
#define MQE_MAX_ROWS (MQE_MAX_LIMIT + 3 * MQE_MAX_LOCAL)
#define MQE_MAX_PAIR 16          // dynamic-vs-dynamic contacts per env
#define MQE_ROBOT_BOUND 0.60f
#define MQE_HIST_PAD 80          // one 70-float frame padded to 5 x 16 for the tensor-core K loop
#define MQE_NV 18
#define MQE_TRACE_COLS 20       // start ns, end ns, pair contacts, widest row count, cycles of P1..P5, integrate, prologue, epilogue

struct DevParams {
    int N, A, P, D, G;            // envs (local), agents, npcs, npc dofs per env, actors per env (A+P)
    int Pd;                       // NPCs simulated as free rigid bodies (P for sheep / ball, 0 otherwise)
    int env_off;
    int npc_kind, npc_ctrl;
    int decimation, iters, max_ep_len, term_mask, quat_alias, defender, command_vel, policy_mode;
    int E;                        // envs per warp in the substep kernel
    int NB;                       // rigid bodies per env (17A+P)
    float dt, gz, mu, coff, vdep, erp, cfm, floor_z, wall_top, limit_margin;
    float term_roll, term_pitch, term_zlow, term_zhigh;
    float act_scale[3], cmd_scale[3];
    float action_scale, hip_scale, clip_actions;
    float dof_lo, dof_hi, bvel_lo, bvel_hi;
    int has_bpos, has_npos, has_nrpy;
    float bpos_x[2], bpos_y[2], npos_x[2], npos_y[2], nrpy_r[2], nrpy_p[2], nrpy_y[2];
    float npc_mass, npc_inertia, npc_radius, npc_halflen;
    float sheep_scale, sheep_rand, gate_x;
    float *lag_ring; int lag_n;                // action lag (go1.py:337-339): [N*A][lag_n][12] scaled actions, lag_n = lag_timesteps + 1; nullptr = off
    int control_type; float kp, kd;            // cfg.control: 0 actuator net, 1 PD position, 2 torque, 3 PD velocity
    float *sub_tau, *sub_qd;                   // post_decimation_step logs [N][decimation][12A] (legged_robot.py:112-115); nullptr until asked for
    unsigned char *sub_exceed; float soft_limit;
    unsigned char *result_done;                // done flags inside half 0 of MQE_BUF_STEP_RESULT (written by k_post_physics / k_reset_all)
    long long result_half;                     // bytes between the two halves of MQE_BUF_STEP_RESULT; half in use = policy steps done & 1
    int push_interval; float max_push_vel;     // domain_rand.push_robots
    const float *base_mass_add;                // [N*A] mass added to the base link (domain_rand.randomize_base_mass) or nullptr
    const float *mu_env;                       // [N] per-env friction (domain_rand.randomize_friction) or nullptr
    const float *base_com_shift;               // [N*A][3] base-link COM shift (domain_rand.randomize_com) or nullptr
    const float *motor_strength;               // [N][12A] action factor of control types P / V / T (domain_rand.randomize_motor) or nullptr
    float geom[16];               // MQE_NPC_SEESAW geometry (MqeSimDesc.npc_geom)
    unsigned long long seed;
    int sdf_nx, sdf_ny;
    float sdf_cell;
    // device pointers ------------------------------------------------------------------
    const float *sdf, *env_origins, *agent_origins, *base_init, *npc_init, *npc_dof_default;
    const MqeRobotModel *model;   // global copy (kernels stage it in shared memory)
    const float *substep_hdr;     // [model | act_w | per-leg probe / capsule lists]: the CTA header of k_substeps, one bulk copy
    const float *act_w;           // packed actuator weights: W0[32][6] b0[32] W1[32][32] b1[32] W2[32] b2
    const float *loc_default;     // [70]
    float *root, *dof, *contact, *torques, *actions, *last_actions;
    float *loc_last, *loc_last2, *loc_obs, *err1, *err2, *vel1, *vel2, *gait, *clock;
    float *base_quat, *base_lin_vel, *base_ang_vel, *proj_grav, *obs, *commands;
    float *last_dof_vel, *last_root_vel, *sheep_stats;
    long long *ep_len;
    unsigned char *reset_buf, *timeout_buf, *collide_buf, *r_term, *p_term, *zl_term, *zh_term;
    unsigned int *episode;
    unsigned char *hist_dirty;    // [N] history must be zeroed before the next frame is appended (go1.py:141-145)
    float *row_scratch, *prow_scratch, *pdesc_scratch;
    int *cand_scratch; int max_cand;     // k_substeps: capsule-pair candidates per env (upper bound: all capsule pairs of all group pairs)
      // k_substeps: local / pair constraint rows that do not fit in shared memory
    long long *warp_trace;        // [ceil(N/E)][MQE_TRACE_COLS] k_substeps per-warp trace
    const int *task_order;        // [ceil(N/E)] which env group the i-th warp of the k_substeps grid integrates (null: identity); k_balance_tasks
    int *task_cost;               // [2][ceil(N/E)] duration [ns] of every env group's warp in the last two launches (half = step parity)
    int fuse_post;                // k_substeps finishes the step itself (post_dev.cuh stages in its epilogue); set by mqe_sim_step only
    int act_mma;                  // actuator network of k_substeps on mma.sync (fp16 hi / lo split; default) or as FFMA2 chains (MQE_ACT_MMA=0)
    int cta_sync;                 // MQE_CTA_SYNC: 0 none, 1 per substep, 2 also per phase
    int trace;                    // MQE_TRACE=1: also accumulate per-phase cycles into the trace rows
    int *stats;                   // [8]
    int *ctr;                     // device-side step counters ([3] _compute_torques calls so far, [4] CTAs of k_substeps finished: action lag; [5] peer exchanges done;
                                  // [7] ring slot of the current step's frame);
                                  // [0] ring slot that receives the next frame, [1] policy steps done
                                  // (sheep RNG key), [2] scratch (blocks of k_post_physics finished); let a captured CUDA graph of
                                  // the whole step be replayed with constant kernel arguments
    // history ring for the policy (policy.cu)
    float *hist_f32;              // [M][30][80] fp32 ring: slot s holds one padded 70-float frame
    unsigned short *hist_hi, *hist_lo;  // bf16 split ring, blocked layout (tensor-core modes)
};

// task-wrapper gather (wrapper.cu): what mqe/envs/wrappers/go1_{sheep,seesaw,football}_wrapper.py compute after every env step
struct WrapParams {
    int kind, D, Aw;              // MqeWrapperKind, observation floats per agent, agents the wrapper reports (football defender: 2 of 3)
    float scale[8];               // reward scales in the order of MqeWrapperDesc
    const float *gate;            // sheep: [N][2] gate position (env-relative); football defender: [N][3] gate position (world)
    float *obs, *reward;          // [N][Aw][D], [N][Aw] inside half 0 of MQE_BUF_STEP_RESULT (+ DevParams.result_half for half 1)
    double *sums;                 // [16]: running sums of the reward terms, [8] = steps
    float *last;                  // sheep: [N][2] last flock centre; seesaw: [N][Aw] last x
    unsigned char *delayed_reset; // sheep: reset_buf of the previous step (go1_sheep_wrapper.py:116)
    int *has_last;                // [N]
};

// per-step exchange of the packed step result between the ranks of one node over NVLink peer memory (gather.cu)
#define MQE_MAX_RANKS 16
#define MQE_STAT_GATHER_TIMEOUT 5           // index into MQE_BUF_STATS: a peer's flag did not arrive within the bounded wait
struct GatherParams {
    int rank, world;                        // world 0: exchange off
    unsigned char *peer[MQE_MAX_RANKS];     // receive buffer of every rank (peer[rank] is this rank's own allocation)
    long long parity_bytes;                 // one parity half: [obs region | reward region | done region], each world x local bytes
    long long flags_off;                    // u32 flags[2][MQE_MAX_RANKS] behind the two halves
    long long seg_src[3], seg_bytes[3], seg_dst[3];   // local field offset in MQE_BUF_STEP_RESULT, its bytes (x16), region offset in a half
    const unsigned char *src;               // this rank's MQE_BUF_STEP_RESULT (half 0)
    int *blocks_done;                       // device scratch
};

// ---------------------------------------------------------------------------------------------- programmatic dependent launch
// Every kernel of the step is launched with programmatic stream serialization: it may be scheduled while its predecessor
// is still running and does its own set-up (barrier init, TMEM allocation, constant staging) in that shadow; pdl_wait()
// blocks until the predecessor grid has completed and its writes are visible, so it must precede the first access to
// anything another kernel of the step produces or consumes.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#ifdef __CUDACC__
// MQE_PDL = 0: plain stream order everywhere; 1 (default): only the short policy-tail kernels overlap their set-up with
// the predecessor (measured: letting the 211 KB-per-CTA substep grid or the 384-CTA layer-0 grid become resident early costs
// more than the launch gaps it hides); 2: every kernel of the step.
static inline int mqe_pdl_level() {
    static const int level = [] { const char *e = getenv("MQE_PDL"); return e ? atoi(e) : 1; }();
    return level;
}
// Launch priorities (captured into the step graph as kernel-node attributes): the kernels on the critical path of a step get the
// device's greatest priority, the incremental layer-0 pass for the NEXT step (launch_background) the least, so the block scheduler
// places its CTAs only on SMs the step's own grids do not want -- the idle tail of k_substeps.
static inline void mqe_priority_range(int *least, int *greatest) {
    static int lo = 0, hi = 0, init = 0;
    if (!init) { if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) { cudaGetLastError(); lo = hi = 0; } init = 1; }
    *least = lo; *greatest = hi;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_if(bool allow, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    int least, greatest;
    mqe_priority_range(&least, &greatest);
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributePriority;
    attr[0].val.priority = greatest;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = allow ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>      // work for the next step that may only use what the current step leaves idle
static inline cudaError_t launch_background(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    int least, greatest;
    mqe_priority_range(&least, &greatest);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributePriority;
    attr[0].val.priority = least;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>      // short kernels of the policy tail
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    return launch_pdl_if(mqe_pdl_level() >= 1, kernel, grid, block, smem, st, args...);
}
template <typename... KArgs, typename... Args>      // heavy grids: frame, layer 0, substeps, post
static inline cudaError_t launch_heavy(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    return launch_pdl_if(mqe_pdl_level() >= 2, kernel, grid, block, smem, st, args...);
}
#endif

// ---------------------------------------------------------------------------------------------- mbarrier / bulk copy
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), done = 0;
    for (unsigned long long spin = 0; !done; spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (spin > (1ull << 28)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#endif

// ---------------------------------------------------------------------------------------------- vec3
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct M3 { V3 c0, c1, c2; };   // columns
__device__ __forceinline__ V3 mul(const M3 &R, V3 v) { return v.x * R.c0 + v.y * R.c1 + v.z * R.c2; }
__device__ __forceinline__ M3 quat_to_mat(float x, float y, float z, float w) {
    M3 R;
    R.c0 = mk(1 - 2 * (y * y + z * z), 2 * (x * y + z * w), 2 * (x * z - y * w));
    R.c1 = mk(2 * (x * y - z * w), 1 - 2 * (x * x + z * z), 2 * (y * z + x * w));
    R.c2 = mk(2 * (x * z + y * w), 2 * (y * z - x * w), 1 - 2 * (x * x + y * y));
    return R;
}
__device__ __forceinline__ M3 rot_x(const M3 &R, float c, float s) { M3 o; o.c0 = R.c0; o.c1 = c * R.c1 + s * R.c2; o.c2 = c * R.c2 - s * R.c1; return o; }
__device__ __forceinline__ M3 rot_y(const M3 &R, float c, float s) { M3 o; o.c1 = R.c1; o.c0 = c * R.c0 - s * R.c2; o.c2 = s * R.c0 + c * R.c2; return o; }

// isaacgym.torch_utils.quat_rotate_inverse
__device__ __forceinline__ V3 quat_rotate_inverse(const float *q, V3 v) {
    float w = q[3];
    V3 u = mk(q[0], q[1], q[2]);
    V3 c = cross(u, v);
    float d = dot(u, v), k = 2.f * w * w - 1.f;
    return k * v - (2.f * w) * c + (2.f * d) * u;
}
__device__ __forceinline__ void get_euler_xyz(const float *q, float *rpy) {
    const float PI = 3.14159265358979323846f, TWO_PI = 6.28318530717958647692f;
    float x = q[0], y = q[1], z = q[2], w = q[3];
    float roll = atan2f(2.f * (w * x + y * z), w * w - x * x - y * y + z * z);
    float sinp = 2.f * (w * y - z * x);
    float pitch = fabsf(sinp) >= 1.f ? copysignf(PI * 0.5f, sinp) : asinf(sinp);
    float yaw = atan2f(2.f * (w * z + x * y), w * w + x * x - y * y - z * z);
    roll = fmodf(roll, TWO_PI);   if (roll < 0.f) roll += TWO_PI;
    pitch = fmodf(pitch, TWO_PI); if (pitch < 0.f) pitch += TWO_PI;
    yaw = fmodf(yaw, TWO_PI);     if (yaw < 0.f) yaw += TWO_PI;
    rpy[0] = roll; rpy[1] = pitch; rpy[2] = yaw;
}
__device__ __forceinline__ void quat_from_euler_xyz(float roll, float pitch, float yaw, float *q) {
    float sy, cy, sr, cr, sp, cp;
    sincosf(yaw * 0.5f, &sy, &cy); sincosf(roll * 0.5f, &sr, &cr); sincosf(pitch * 0.5f, &sp, &cp);
    q[3] = cy * cr * cp + sy * sr * sp;
    q[0] = cy * sr * cp - sy * cr * sp;
    q[1] = cy * cr * sp + sy * sr * cp;
    q[2] = sy * cr * cp - cy * sr * sp;
}

// ---------------------------------------------------------------------------------------------- counter RNG
// Bit-identical to oracle/mqe_oracle.c: keyed (seed, global env, counter, stream, index).
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t rng_u32(unsigned long long seed, uint32_t env, uint32_t counter, uint32_t stream, uint32_t idx) {
    uint32_t h = mix32((uint32_t)seed ^ 0x9E3779B9U);
    h = mix32(h ^ (uint32_t)(seed >> 32));
    h = mix32(h ^ env);
    h = mix32(h ^ counter);
    h = mix32(h ^ (stream * 0x10001U + idx * 0x9E3779B1U));
    return h;
}
__device__ __forceinline__ float rng_uniform(unsigned long long seed, uint32_t env, uint32_t counter, uint32_t stream, uint32_t idx) {
    return (float)(rng_u32(seed, env, counter, stream, idx) >> 8) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float rng_normal(unsigned long long seed, uint32_t env, uint32_t counter, uint32_t stream, uint32_t idx) {
    float u1 = ((float)(rng_u32(seed, env, counter, stream, 2 * idx) >> 8) + 1.f) * (1.0f / 16777216.0f);
    float u2 = rng_uniform(seed, env, counter, stream, 2 * idx + 1);
    return sqrtf(-2.f * logf(u1)) * cosf(6.28318530717958647692f * u2);
}
enum { RNG_DOF = 0, RNG_BASE_POS = 1, RNG_BASE_VEL = 2, RNG_NPC_POS = 3, RNG_NPC_RPY = 4, RNG_SHEEP = 5, RNG_PUSH = 6 };

// ---------------------------------------------------------------------------------------------- static world
struct SdfSample { float sdf, gx, gy; };
__device__ __forceinline__ SdfSample sdf_sample(const DevParams &p, float x, float y) {
    float fx = x / p.sdf_cell, fy = y / p.sdf_cell;
    float mx = (float)(p.sdf_nx - 1) - 1e-3f, my = (float)(p.sdf_ny - 1) - 1e-3f;
    fx = fminf(fmaxf(fx, 0.f), mx);
    fy = fminf(fmaxf(fy, 0.f), my);
    int i = (int)fx, j = (int)fy;
    float tx = fx - (float)i, ty = fy - (float)j;
    const float *S = p.sdf + (size_t)i * p.sdf_ny + j;
    float s00 = __ldg(S), s01 = __ldg(S + 1), s10 = __ldg(S + p.sdf_ny), s11 = __ldg(S + p.sdf_ny + 1);
    SdfSample r;
    r.sdf = (1.f - tx) * (1.f - ty) * s00 + tx * (1.f - ty) * s10 + (1.f - tx) * ty * s01 + tx * ty * s11;
    r.gx = ((1.f - ty) * (s10 - s00) + ty * (s11 - s01)) / p.sdf_cell;
    r.gy = ((1.f - tx) * (s01 - s00) + tx * (s11 - s10)) / p.sdf_cell;
    return r;
}
