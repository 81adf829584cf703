// policy.cu -- preprocess_action (go1.py:64-108): command -> 70-float walk-these-ways frame -> 30-frame history ->
// adaptation module + actor body (go1.py:400-407) -> clipped joint-position actions.
//
// History is a RING, not the reference's shift-and-concat (go1.py:102 rewrites all 8.4 KB per agent per step):
// slot `head` of [M][30][80] receives this step's frame; the first layer contracts slot s against the weight block
// of its age, block(s) = (s - head - 1) mod 30 (0 = oldest ... 29 = newest), so nothing ever moves.
// Frames are padded 70 -> 80 floats (zeros) so every slot is 16-byte aligned and a multiple of the MMA K step.
//
// Two arithmetic paths produce the same numbers to fp32 rounding:
//   MQE_POLICY_FP32   : CUDA-core SGEMM chain (this file, k_linear)         -- the exact-arithmetic reference path
//   MQE_POLICY_BF16X3 : tcgen05 tensor cores, bf16 hi/lo split, 3 MMAs       -- policy_tc.cu
#include "common.cuh"
#include "kernels.cuh"

#define FRAME_PAD MQE_HIST_PAD
#define RING_ROW (MQE_HIST_FRAMES * FRAME_PAD)

// element (row m, slot s, column i) of the pre-tiled bf16 planes: [m/128][30][10][128][8]  (policy_tc.cu)
__device__ __forceinline__ size_t tc_ring_index(int m, int s, int i) {
    return ((((size_t)(m >> 7) * MQE_HIST_FRAMES + s) * (FRAME_PAD / 8) + (i >> 3)) * 128 + (m & 127)) * 8 + (i & 7);
}
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }

// defender command (go1_football_defender.py:56-80) from the previous step's observations
__device__ void dev_defender_command(const DevParams &p, int e, float *cmd) {
    const int A = p.A, G = p.G;
    const float PI = 3.14159265358979323846f;
    const float *dp = p.root + ((size_t)e * G + 2) * 13;
    const float *bp = p.root + ((size_t)e * G + A) * 13;
    float gx = p.env_origins[e * 3] + p.gate_x, gy = p.env_origins[e * 3 + 1];
    float tx = 0.6f * bp[0] + 0.4f * gx, ty = 0.6f * bp[1] + 0.4f * gy;
    float yaw = p.obs[(size_t)(e * A + 2) * MQE_OBS_FLOATS + MQE_OBS_BASE_RPY + 2];
    float yaw_to_gate = PI + atanf((gy - dp[1]) / (gx - dp[0]));
    float yc = fminf(fmaxf(yaw_to_gate - yaw, -0.3f), 0.3f) / 0.3f;
    float tdg = sqrtf((tx - gx) * (tx - gx) + (ty - gy) * (ty - gy));
    float ddg = sqrtf((dp[0] - gx) * (dp[0] - gx) + (dp[1] - gy) * (dp[1] - gy));
    float xc = fminf(fmaxf(tdg - ddg, -0.5f), 0.5f);
    float yy = gy + (ty - gy) * (dp[0] - gx) / (tx - gx) - dp[1];
    yy = fminf(fmaxf(yy, -0.5f), 0.5f);
    cmd[0] = xc; cmd[1] = -yy; cmd[2] = yc;
}

// One warp per agent row: wrapper scaling + clip (wrappers/*.py step(), go1.py:38), command -> frame, ring append.
// d_actions: [N][A_ctrl][3].  Also maintains the bf16 hi/lo ring when p.hist_hi != nullptr.
__global__ void __launch_bounds__(256) k_policy_frame(DevParams p, const float *__restrict__ d_actions, int head) {
    pdl_launch_dependents();
    pdl_wait();
    if (head < 0) head = p.ctr[0];                           // graph replay: the slot comes from the device counter
    if (blockIdx.x == 0 && threadIdx.x == 0) p.ctr[7] = head;   // slot of THIS step's frame: what the background layer-0 passes key on (ctr[0] moves mid-step)
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int M = p.N * p.A;
    if (m >= M) return;
    const int A = p.A, e = m / A, a = m % A, actrl = p.defender ? A - 1 : A;
    float cmd[3];
    if (a < actrl) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            float x = d_actions[((size_t)e * actrl + a) * 3 + i];
            x = fminf(fmaxf(x, -1.f), 1.f) * p.act_scale[i];
            cmd[i] = x;
        }
    } else dev_defender_command(p, e, cmd);
    if (!p.defender)
#pragma unroll
        for (int i = 0; i < 3; i++) cmd[i] = fminf(fmaxf(cmd[i], -1.f), 1.f);
    if (lane < 3) p.commands[m * 3 + lane] = cmd[lane];
    const float *ob = p.obs + (size_t)m * MQE_OBS_FLOATS;
    float *lo = p.loc_obs + (size_t)m * MQE_LOC_OBS;
    float *ring = p.hist_f32 ? p.hist_f32 + (size_t)m * RING_ROW : nullptr;     // fp32 ring: MQE_POLICY_FP32 only (the planes are the ring otherwise)
    if (p.hist_dirty[e]) {                                   // _reset_buffers zeroed this row's history
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ring)
            for (int i = lane; i < RING_ROW / 4; i += 32)
                if (i / (FRAME_PAD / 4) != head) reinterpret_cast<float4 *>(ring)[i] = z;
        if (p.hist_hi) {                                     // pre-tiled planes: one 16-byte chunk per (slot, k-chunk)
            uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
            for (int i = lane; i < MQE_HIST_FRAMES * (FRAME_PAD / 8); i += 32) {
                int s = i / (FRAME_PAD / 8), c = i % (FRAME_PAD / 8);
                if (s == head) continue;
                size_t o = tc_ring_index(m, s, c * 8);
                *reinterpret_cast<uint4 *>(p.hist_hi + o) = z4;
                *reinterpret_cast<uint4 *>(p.hist_lo + o) = z4;
            }
        }
    }
    for (int i = lane; i < FRAME_PAD; i += 32) {
        float v = 0.f;
        if (i < 3) v = ob[MQE_OBS_PROJ_GRAVITY + i];
        else if (i < 6) v = p.command_vel ? cmd[i - 3] * p.cmd_scale[i - 3] : lo[i];
        else if (i < 18) v = lo[i];
        else if (i < 30) v = ob[MQE_OBS_DOF_POS + i - 18];
        else if (i < 42) v = ob[MQE_OBS_DOF_VEL + i - 30];
        else if (i < 54) v = p.loc_last[m * 12 + i - 42];
        else if (i < 66) v = p.loc_last2[m * 12 + i - 54];
        else if (i < 70) v = ob[MQE_OBS_CLOCK + i - 66];
        if (i < MQE_LOC_OBS) lo[i] = v;
        if (ring) ring[head * FRAME_PAD + i] = v;
        if (p.hist_hi) {
            unsigned int b = __float_as_uint(v);
            unsigned int hi = (b + 0x7fffu + ((b >> 16) & 1u)) & 0xffff0000u;       // round-to-nearest-even bf16
            float res = v - __uint_as_float(hi);
            unsigned int rb = __float_as_uint(res);
            unsigned int lo16 = (rb + 0x7fffu + ((rb >> 16) & 1u)) >> 16;
            size_t o = tc_ring_index(m, head, i);
            p.hist_hi[o] = (unsigned short)(hi >> 16);
            p.hist_lo[o] = (unsigned short)lo16;
        }
    }
}

// after the network: last_locomotion_action(s) shift, clip to +-clip_actions (go1.py:40-41, 104-106)
__global__ void k_policy_finish(DevParams p, const float *__restrict__ act) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int M = p.N * p.A;
    if (t < M * 12) {
        float a = act[t];
        p.loc_last2[t] = p.loc_last[t];
        p.loc_last[t] = a;
        p.actions[t] = fminf(fmaxf(a, -p.clip_actions), p.clip_actions);
    }
    if (t < p.N) p.hist_dirty[t] = 0;
    if (t < 5) p.stats[t] = 0;                           // [5..7] are sticky (MQE_STAT_GATHER_TIMEOUT)
    if (t == 0) p.ctr[0] = (p.ctr[0] + 1) % MQE_HIST_FRAMES;  // nobody reads the slot counter after this point of the step                               // contact statistics of the step that follows (k_substeps accumulates)
}

// Go1.step() for control_type 'P' / 'V' / 'T' (go1.py:43-45 -> pre_physics_step, legged_robot.py:108-110): the caller's joint actions,
// clipped to +-clip_actions, ARE `actions`; the walk policy is never evaluated
__global__ void k_joint_actions(DevParams p, const float *__restrict__ joint_actions) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < p.N * p.A * 12) p.actions[t] = fminf(fmaxf(joint_actions[t], -p.clip_actions), p.clip_actions);
    if (t < p.N) p.hist_dirty[t] = 0;
    if (t < 5) p.stats[t] = 0;                           // [5..7] are sticky (MQE_STAT_GATHER_TIMEOUT)
}
extern "C" cudaError_t mqe_launch_joint_actions(const DevParams &p, const float *joint_actions, cudaStream_t st) {
    const int n = p.N * p.A * 12;
    return launch_pdl(k_joint_actions, dim3((n + 255) / 256), dim3(256), 0, st, p, joint_actions);
}

// ---------------------------------------------------------------------------------------------- fp32 SGEMM chain
// Y[m][n] = act(sum_k X[m][k] * W[n][wk(k)] + bias[n]),  act = ELU for n < elu_cols.  When head >= 0, X is the ring
// and wk rotates whole 80-float blocks by age; K % 16 == 0, ldx % 4 == 0, ldw % 4 == 0.
#define LBM 64
#define LBN 64
#define LBK 16
__global__ void __launch_bounds__(256) k_linear(const float *__restrict__ X, int ldx, const float *__restrict__ W, int ldw,
                                                const float *__restrict__ bias, float *__restrict__ Y, int ldy,
                                                int M, int N, int K, int elu_cols, int head) {
    __shared__ float Xs[LBK][LBM + 4], Ws[LBK][LBN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * LBM, n0 = blockIdx.x * LBN;
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += LBK) {
        int wk0 = k0;
        if (head >= 0) {
            int s = k0 / FRAME_PAD, b = s - head - 1;
            if (b < 0) b += MQE_HIST_FRAMES;
            wk0 = b * FRAME_PAD + (k0 - s * FRAME_PAD);
        }
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), wv = xv;
        if (m0 + lr < M) xv = *reinterpret_cast<const float4 *>(X + (size_t)(m0 + lr) * ldx + k0 + lk);
        if (n0 + lr < N) wv = *reinterpret_cast<const float4 *>(W + (size_t)(n0 + lr) * ldw + wk0 + lk);
        Xs[lk][lr] = xv.x; Xs[lk + 1][lr] = xv.y; Xs[lk + 2][lr] = xv.z; Xs[lk + 3][lr] = xv.w;
        Ws[lk][lr] = wv.x; Ws[lk + 1][lr] = wv.y; Ws[lk + 2][lr] = wv.z; Ws[lk + 3][lr] = wv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < LBK; kk++) {
            float4 a4 = *reinterpret_cast<const float4 *>(&Xs[kk][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4 *>(&Ws[kk][tx * 4]);
            float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (n < elu_cols) v = elu1(v);
            Y[(size_t)m * ldy + n] = v;
        }
    }
}

// body layer 0 tail: Z[m][256+n] = ELU(Z[m][256+n] + Wlat[n][0] lat0 + Wlat[n][1] lat1)   (the cat(h, latent) columns)
__global__ void k_body_latent(float *__restrict__ Z, const float *__restrict__ latent, const float *__restrict__ Wlat, int M) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M * 512) return;
    const int m = t >> 9, n = t & 511;
    float v = Z[(size_t)m * 768 + 256 + n] + Wlat[n * 2] * latent[m * 2] + Wlat[n * 2 + 1] * latent[m * 2 + 1];
    Z[(size_t)m * 768 + 256 + n] = elu1(v);
}

// [rows][2100] history (oldest frame first, the reference layout) -> padded ring with head = 29
__global__ void k_history_to_ring(const float *__restrict__ hist, float *__restrict__ ring, unsigned short *__restrict__ hi,
                                  unsigned short *__restrict__ lo16, int rows) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)rows * RING_ROW) return;
    const int r = (int)(t / RING_ROW), k = (int)(t % RING_ROW), s = k / FRAME_PAD, i = k % FRAME_PAD;
    float v = i < MQE_LOC_OBS ? hist[(size_t)r * 2100 + s * MQE_LOC_OBS + i] : 0.f;
    if (ring) ring[t] = v;
    if (hi) {
        unsigned int b = __float_as_uint(v);
        unsigned int h = (b + 0x7fffu + ((b >> 16) & 1u)) & 0xffff0000u;
        float res = v - __uint_as_float(h);
        unsigned int rb = __float_as_uint(res);
        size_t o = tc_ring_index(r, s, i);
        hi[o] = (unsigned short)(h >> 16);
        lo16[o] = (unsigned short)((rb + 0x7fffu + ((rb >> 16) & 1u)) >> 16);
    }
}

static inline dim3 lin_grid(int M, int N) { return dim3((N + LBN - 1) / LBN, (M + LBM - 1) / LBM); }

extern "C" cudaError_t mqe_launch_policy_l0_fp32(const PolicyWeightsDev &w, const PolicyScratch &s, const float *ring, int head, int M, cudaStream_t st) {
    k_linear<<<lin_grid(M, 768), 256, 0, st>>>(ring, RING_ROW, w.w0cat, RING_ROW, w.b0cat, s.Z, 768, M, 768, RING_ROW, 256, head);
    return cudaGetLastError();
}
// layers 1.. of both networks from Z = [ELU(adapt.0) | body.0 pre-activation without the latent columns]
extern "C" cudaError_t mqe_launch_policy_tail(const PolicyWeightsDev &w, const PolicyScratch &s, int M, cudaStream_t st, int *launches) {
    k_linear<<<lin_grid(M, 128), 256, 0, st>>>(s.Z, 768, w.aw1, 256, w.ab1, s.T1, 128, M, 128, 256, 128, -1);
    k_linear<<<lin_grid(M, 2), 256, 0, st>>>(s.T1, 128, w.aw2, 128, w.ab2, s.latent, 2, M, 2, 128, 0, -1);
    k_body_latent<<<(M * 512 + 255) / 256, 256, 0, st>>>(s.Z, s.latent, w.wlat, M);
    k_linear<<<lin_grid(M, 256), 256, 0, st>>>(s.Z + 256, 768, w.bw1, 512, w.bb1, s.T2, 256, M, 256, 512, 256, -1);
    k_linear<<<lin_grid(M, 128), 256, 0, st>>>(s.T2, 256, w.bw2, 256, w.bb2, s.T3, 128, M, 128, 256, 128, -1);
    k_linear<<<lin_grid(M, 12), 256, 0, st>>>(s.T3, 128, w.bw3, 128, w.bb3, s.act, 12, M, 12, 128, 0, -1);
    *launches += 6;
    return cudaGetLastError();
}
extern "C" cudaError_t mqe_launch_policy_frame(const DevParams &p, const float *d_actions, int head, cudaStream_t st) {
    const int M = p.N * p.A;
    return launch_heavy(k_policy_frame, dim3((M * 32 + 255) / 256), dim3(256), 0, st, p, d_actions, head);
}
extern "C" cudaError_t mqe_launch_policy_finish(const DevParams &p, const float *act, cudaStream_t st) {
    const int M = p.N * p.A, n = M * 12 > p.N ? M * 12 : p.N;
    return launch_pdl(k_policy_finish, dim3((n + 255) / 256), dim3(256), 0, st, p, act);
}
extern "C" cudaError_t mqe_launch_history_to_ring(const float *hist, float *ring, unsigned short *hi, unsigned short *lo, int rows, cudaStream_t st) {
    size_t n = (size_t)rows * RING_ROW;
    k_history_to_ring<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hist, ring, hi, lo, rows);
    return cudaGetLastError();
}
