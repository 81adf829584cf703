// wrapper.cu -- the task wrappers' per-step observation / reward gather as ONE kernel inside the step graph.
//
// Replaces the ~40 small torch launches (and one host sync per reward term) that mqe/envs/wrappers/go1_sheep_wrapper.py:54-118,
// go1_seesaw_wrapper.py:48-120 and go1_football_wrapper.py:57-91 issue after every Go1.step().  The Python
// wrappers in mqe_b200/envs/wrappers.py keep the same arithmetic in torch: they are pinned by the reference's own code on the CPU
// (tests/test_wrappers_golden.py) and are what this kernel is tested against on the GPU (tests/test_gpu_parity.py).
#include "common.cuh"
#include "kernels.cuh"

#define WRAP_THREADS 128

__device__ __forceinline__ void base_info(const DevParams &p, int m, float *o) {     // (pos, rpy) of agent row m: go1.py:153-196 obs struct
    const float *r = p.obs + (size_t)m * MQE_OBS_FLOATS;
    o[0] = r[MQE_OBS_BASE_POS]; o[1] = r[MQE_OBS_BASE_POS + 1]; o[2] = r[MQE_OBS_BASE_POS + 2];
    o[3] = r[MQE_OBS_BASE_RPY]; o[4] = r[MQE_OBS_BASE_RPY + 1]; o[5] = r[MQE_OBS_BASE_RPY + 2];
}

// mode 0: step (obs + reward + running sums); mode 1: wrapper reset() (obs only; sheep forgets its last flock centre)
__global__ void __launch_bounds__(WRAP_THREADS) k_task_gather(DevParams p, WrapParams w, int mode) {
    pdl_launch_dependents();
    pdl_wait();                                         // launched with the PDL attribute at MQE_PDL=2: k_post_physics must have finished
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int A = p.A, P = p.P, Aw = w.Aw, D = w.D;
    // half of the double-buffered step result this step writes: k_post_physics has already advanced ctr[1] (reset: the current half)
    const long long half_f = (long long)(p.ctr[1] & 1) * (p.result_half >> 2);
    w.obs += half_f; w.reward += half_f;
    float term[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (e < p.N) {
        float bi[4][6];
        for (int a = 0; a < Aw; a++) base_info(p, e * A + a, bi[a]);
        // ---- observation: one-hot id, (pos, rpy) self, (pos, rpy) of the agent at the mirrored index, then task columns ----
        for (int a = 0; a < Aw; a++) {
            float *o = w.obs + ((size_t)e * Aw + a) * D;
            for (int k = 0; k < Aw; k++) o[k] = k == a ? 1.f : 0.f;
            for (int k = 0; k < 6; k++) { o[Aw + k] = bi[a][k]; o[Aw + 6 + k] = bi[Aw - 1 - a][k]; }
        }
        const float *eo = p.env_origins + (size_t)e * 3;
        float reward = 0.f;
        if (w.kind == MQE_WRAP_SHEEP) {
            const float gx = w.gate[e * 2], gy = w.gate[e * 2 + 1];
            float sx[12], sy[12];
            for (int n = 0; n < P; n++) {
                const float *r = p.root + ((size_t)e * p.G + A + n) * 13;
                sx[n] = r[0] - eo[0]; sy[n] = r[1] - eo[1];
            }
            for (int a = 0; a < Aw; a++) {
                float *o = w.obs + ((size_t)e * Aw + a) * D + Aw + 12;
                o[0] = gx; o[1] = gy;
                for (int n = 0; n < P; n++) { o[2 + 2 * n] = sx[n]; o[3 + 2 * n] = sy[n]; }
            }
            if (mode == 0) {
                const float s_success = w.scale[0], s_contact = w.scale[1], s_move = w.scale[2], s_mixed = w.scale[3], s_lin = w.scale[4], s_exp = w.scale[5];
                if (s_success != 0.f) {                                    // the count itself, not scaled (go1_sheep_wrapper.py:73-77)
                    int cnt = 0;
                    for (int n = 0; n < P; n++) cnt += (sx[n] - gx) > 0.f;
                    reward = (float)cnt; term[0] = (float)cnt;
                }
                if (s_contact != 0.f) { const float c = s_contact * (float)p.collide_buf[e]; reward += c; term[1] = c; }
                const float ax = p.sheep_stats[e * 3], ay = p.sheep_stats[e * 3 + 1], var = p.sheep_stats[e * 3 + 2];
                if (s_move != 0.f) {
                    if (w.has_last[e]) {
                        float xm = ax - w.last[e * 2];
                        if (w.delayed_reset[e]) xm = 0.f;
                        const float r = s_move * xm;
                        reward += r; term[2] = r;
                    }
                    w.last[e * 2] = ax; w.last[e * 2 + 1] = ay;
                    w.has_last[e] = 1;
                }
                if (s_mixed != 0.f) {
                    float sum = 0.f;
                    for (int n = 0; n < P; n++) {
                        const float dx = sx[n] - gx, dy = sy[n] - gy;
                        float mval = expf(-sqrtf(dx * dx + dy * dy) / 2.f) * s_mixed;
                        if (sx[n] >= gx) mval = s_mixed;
                        sum += mval;
                    }
                    reward += sum; term[3] = sum;
                }
                if (s_lin != 0.f || s_exp != 0.f) {
                    const float r = s_lin * (var - 1.f) + s_exp * expf(var / 2.f - 1.f);
                    reward += r; term[4] = r;
                }
                w.delayed_reset[e] = p.reset_buf[e];
            } else {
                w.has_last[e] = 0;
            }
        } else if (w.kind == MQE_WRAP_SEESAW && mode == 0) {
            const float s_x = w.scale[0], s_h = w.scale[1], s_y = w.scale[2], s_contact = w.scale[3], s_dist = w.scale[4], s_success = w.scale[5], s_fall = w.scale[6];
            if (s_x != 0.f) {
                float xr = 0.f;
                for (int a = 0; a < Aw; a++) {
                    const float x = bi[a][0];
                    if (w.has_last[e]) xr += x - w.last[e * Aw + a];
                    w.last[e * Aw + a] = x;
                }
                w.has_last[e] = 1;
                if (p.reset_buf[e]) xr = 0.f;
                xr *= s_x;
                reward += xr; term[0] = xr;
            }
            if (s_h != 0.f) { float z = 0.f; for (int a = 0; a < Aw; a++) z += bi[a][2]; const float r = s_h * (z - 0.56f); reward += r; term[1] = r; }
            if (s_y != 0.f) { float y2 = 0.f; for (int a = 0; a < Aw; a++) y2 += bi[a][1] * bi[a][1]; const float r = s_y * (y2 - 0.5f); reward += r; term[2] = r; }
            if (s_contact != 0.f) { const float c = s_contact * (float)p.collide_buf[e]; reward += c; term[3] = c; }
            if (s_dist != 0.f) {
                const float dx = bi[0][0] - bi[Aw - 1][0], dy = bi[0][1] - bi[Aw - 1][1], d2 = dx * dx + dy * dy;
                if (d2 < 0.25f) { const float r = s_dist / fmaxf(d2, 1e-12f); reward += r; term[4] = r; }
            }
            if (s_success != 0.f) {
                int cnt = 0;
                for (int a = 0; a < Aw; a++) cnt += (bi[a][0] > 7.7f) && (bi[a][2] > 1.3f);
                const float r = s_success * (float)cnt;
                reward += r; term[5] = r;
            }
            if (s_fall != 0.f && (p.r_term[e] | p.p_term[e])) { reward += s_fall; term[6] = s_fall; }
        } else if (w.kind == MQE_WRAP_FOOTBALL_DEFENDER) {
            const float *r = p.root + ((size_t)e * p.G + A) * 13;
            const float bx = r[0] - eo[0], by = r[1] - eo[1], bz = r[2] - eo[2];
            for (int a = 0; a < Aw; a++) {
                float *o = w.obs + ((size_t)e * Aw + a) * D + Aw + 12;
                o[0] = bx; o[1] = by; o[2] = bz; o[3] = r[7]; o[4] = r[8]; o[5] = r[9];
            }
            if (mode == 0) {
                const float s_goal = w.scale[0], s_dist = w.scale[1];
                const float gx = w.gate[e * 3], gy = w.gate[e * 3 + 1];     // the reference compares the env-relative ball x with the WORLD gate x (:77)
                if (s_goal != 0.f && bx > gx) { reward += s_goal; term[0] = s_goal; }
                if (s_dist != 0.f) {
                    const float dx = bx - gx, dy = by - gy;
                    const float rr = s_dist * expf(-sqrtf(dx * dx + dy * dy) / 3.f);
                    reward += rr; term[1] = rr;
                }
            }
        }
        if (mode == 0)
            for (int a = 0; a < Aw; a++) w.reward[(size_t)e * Aw + a] = reward;
    }
    if (mode != 0) return;
    // ---- running sums of the reward terms (reward_buffer[...] of the reference, accumulated without a host sync) ----
    __shared__ float red[8][WRAP_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        float v = term[t];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[t][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double s = 0.0;
        for (int k = 0; k < WRAP_THREADS / 32; k++) s += (double)red[threadIdx.x][k];
        if (s != 0.0) atomicAdd(w.sums + threadIdx.x, s);
    }
    if (blockIdx.x == 0 && threadIdx.x == 8) atomicAdd(w.sums + 8, 1.0);       // step count
}

extern "C" cudaError_t mqe_launch_task_gather(const DevParams &p, const WrapParams &w, int mode, cudaStream_t st) {
    return launch_heavy(k_task_gather, dim3((p.N + WRAP_THREADS - 1) / WRAP_THREADS), dim3(WRAP_THREADS), 0, st, p, w, mode);
}
