// wrapper.cu -- the task wrappers' per-step observation / reward gather as ONE kernel inside the step graph.
//
// Replaces the ~40 small torch launches (and one host sync per reward term) that mqe/envs/wrappers/go1_sheep_wrapper.py:54-118,
// go1_seesaw_wrapper.py:48-120 and go1_football_wrapper.py:57-91 issue after every Go1.step().  The Python
// wrappers in mqe_b200/envs/wrappers.py keep the same arithmetic in torch: they are pinned by the reference's own code on the CPU
// (tests/test_wrappers_golden.py) and are what this kernel is tested against on the GPU (tests/test_gpu_parity.py).
#include "common.cuh"
#include "kernels.cuh"

#define WRAP_THREADS 256                                   // one WARP per env: 8 envs per block (small blocks find room on SMs the background layer-0 pass occupies)
#define WRAP_STAGE 84                                       // per-warp staging: 24 base-info floats (<= 4 agents x (pos, rpy)), up to 48 task columns, 12 scalars of the reward terms

// mode 0: step (obs + reward + running sums); mode 1: wrapper reset() (obs only; sheep forgets its last flock centre).
// One warp per env.  The lanes gather what the observation is made of into a per-warp staging row (coalesced loads, one round trip instead
// of a thread walking ~60 dependent accesses: as one thread per env this kernel took 14 us alone and 30-60 us beside the background
// layer-0 pass), every lane then writes its share of the observation rows, and lane 0 evaluates the reward terms from the staging row
// with the arithmetic of the torch wrappers (go1_sheep_wrapper.py:54-118, go1_seesaw_wrapper.py:48-120, go1_football_wrapper.py:57-91,
// go1_pushbox_wrapper.py, go1_wrestling_wrapper.py, go1_bridge_wrapper.py, go1_rotation_wrapper.py).
__global__ void __launch_bounds__(WRAP_THREADS) k_task_gather(DevParams p, WrapParams w, int mode) {
    pdl_launch_dependents();
    pdl_wait();                                         // launched with the PDL attribute at MQE_PDL=2: the producer of the state must have finished
    __shared__ float stage[WRAP_THREADS / 32][WRAP_STAGE];
    __shared__ float red[8][WRAP_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * (WRAP_THREADS / 32) + warp;
    const int A = p.A, P = p.P, Aw = w.Aw, D = w.D;
    // half of the double-buffered step result this step writes: the bookkeeping has already advanced ctr[1] (reset: the current half)
    const long long half_f = (long long)(p.ctr[1] & 1) * (p.result_half >> 2);
    w.obs += half_f; w.reward += half_f;
    float term[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (e < p.N) {
        float *sg = stage[warp];
        const float *eo = p.env_origins + (size_t)e * 3;
        // ---- staging: (pos, rpy) of every wrapper agent (go1.py:153-196 obs struct), then the task columns ----
        if (lane < Aw * 6) {
            const int a = lane / 6, k = lane % 6;
            const float *r = p.obs + (size_t)(e * A + a) * MQE_OBS_FLOATS;
            sg[lane] = r[k < 3 ? MQE_OBS_BASE_POS + k : MQE_OBS_BASE_RPY + k - 3];
        }
        int ncol = 0;
        if (w.kind == MQE_WRAP_SHEEP) {
            ncol = 2 + 2 * P;
            if (lane < 2) sg[24 + lane] = w.gate[e * 2 + lane];
            if (lane < P) {
                const float *r = p.root + ((size_t)e * p.G + A + lane) * 13;
                sg[26 + 2 * lane] = r[0] - eo[0]; sg[27 + 2 * lane] = r[1] - eo[1];
            }
        } else if (w.kind == MQE_WRAP_FOOTBALL_DEFENDER) {
            ncol = 6;
            if (lane < 6) {
                const float *r = p.root + ((size_t)e * p.G + A) * 13;
                sg[24 + lane] = lane < 3 ? r[lane] - eo[lane] : r[7 + lane - 3];
            }
        } else if (w.kind == MQE_WRAP_PUSHBOX) {           // gate xy | box xy (env-relative) | box quaternion (go1_pushbox_wrapper.py)
            ncol = 8;
            if (lane < 8) {
                const float *r = p.root + ((size_t)e * p.G + A) * 13;
                sg[24 + lane] = lane < 2 ? w.gate[e * 2 + lane] : (lane < 4 ? r[lane - 2] - eo[lane - 2] : r[3 + lane - 4]);
            }
        }
        // every scalar the reward terms read, fetched in the SAME round trip by the upper lanes (sg[72..]); lane 0 then works from shared memory
        if (mode == 0) {
            float *ax = sg + 72;
            if (w.kind == MQE_WRAP_SHEEP) {
                if (lane == 16) ax[0] = (float)p.collide_buf[e];
                else if (lane >= 17 && lane < 20) ax[1 + lane - 17] = p.sheep_stats[e * 3 + lane - 17];
                else if (lane == 20) ax[4] = (float)w.has_last[e];
                else if (lane == 21 || lane == 22) ax[5 + lane - 21] = w.last[e * 2 + lane - 21];
                else if (lane == 23) ax[7] = (float)w.delayed_reset[e];
                else if (lane == 24) ax[8] = (float)p.reset_buf[e];
            } else if (w.kind == MQE_WRAP_SEESAW) {
                if (lane == 16) ax[0] = (float)p.collide_buf[e];
                else if (lane == 17) ax[1] = (float)p.reset_buf[e];
                else if (lane == 18) ax[2] = (float)(p.r_term[e] | p.p_term[e]);
                else if (lane == 19) ax[3] = (float)w.has_last[e];
                else if (lane >= 20 && lane < 20 + Aw) ax[4 + lane - 20] = w.last[e * Aw + lane - 20];
            } else if (w.kind == MQE_WRAP_FOOTBALL_DEFENDER) {
                if (lane == 16 || lane == 17) ax[lane - 16] = w.gate[e * 3 + lane - 16];
            } else if (w.kind == MQE_WRAP_PUSHBOX) {
                if (lane == 16) ax[0] = (float)w.has_last[e];
                else if (lane == 17) ax[1] = w.last[e];
                else if (lane == 18) ax[2] = (float)p.reset_buf[e];
            } else if (w.kind == MQE_WRAP_BRIDGE || w.kind == MQE_WRAP_ROTATION) {
                if (lane >= 16 && lane < 20) ax[lane - 16] = w.last[e * 4 + lane - 16];     // what the wrapper's reset() stored (see mode 1 below)
            }
        }
        __syncwarp();
        const bool duel = w.kind == MQE_WRAP_WRESTLING || w.kind == MQE_WRAP_BRIDGE || w.kind == MQE_WRAP_ROTATION;
        if (duel) {
            // (pos, rpy) self | other, no ids; agent 1 lives in a mirrored world: y and pitch negated (wrestling :32-36, rotation :45-49), or x
            // reflected about the midpoint of the two start positions and pitch negated (bridge :33-41; span = |x0 + x1| at the wrapper's reset())
            const float span = w.kind == MQE_WRAP_BRIDGE ? (mode == 0 ? sg[72 + 3] : fabsf(sg[6] + sg[0])) : 0.f;
            for (int idx = lane; idx < 24; idx += 32) {
                const int a = idx / 12, col = idx % 12;
                float v = col < 6 ? sg[a * 6 + col] : sg[(1 - a) * 6 + col - 6];
                if (a == 1) {
                    if (w.kind == MQE_WRAP_BRIDGE) { if (col == 0 || col == 6) v = span - v; else if (col == 4 || col == 10) v = -v; }
                    else if (col % 3 == 1) v = -v;
                }
                w.obs[((size_t)e * 2 + a) * 12 + col] = v;
            }
        }
        // ---- observation rows: one-hot id, (pos, rpy) self, (pos, rpy) of the agent at the mirrored index, then the task columns ----
        for (int idx = lane; !duel && idx < Aw * D; idx += 32) {
            const int a = idx / D, col = idx % D;
            float v;
            if (col < Aw) v = col == a ? 1.f : 0.f;
            else if (col < Aw + 6) v = sg[a * 6 + col - Aw];
            else if (col < Aw + 12) v = sg[(Aw - 1 - a) * 6 + col - Aw - 6];
            else if (col - Aw - 12 < ncol) v = sg[24 + col - Aw - 12];
            else continue;
            w.obs[((size_t)e * Aw + a) * D + col] = v;
        }
        // ---- reward terms: lane 0, scalar, in the wrappers' order ----
        if (lane == 0) {
            float reward = 0.f;
#define BI(a, k) sg[(a) * 6 + (k)]
            if (w.kind == MQE_WRAP_SHEEP) {
                const float gx = sg[24], gy = sg[25];
                const float *sx = sg + 26;                  // sx[n] = sx[2 n], sy[n] = sx[2 n + 1]
                if (mode == 0) {
                    const float s_success = w.scale[0], s_contact = w.scale[1], s_move = w.scale[2], s_mixed = w.scale[3], s_lin = w.scale[4], s_exp = w.scale[5];
                    if (s_success != 0.f) {                                    // the count itself, not scaled (go1_sheep_wrapper.py:73-77)
                        int cnt = 0;
                        for (int n = 0; n < P; n++) cnt += (sx[2 * n] - gx) > 0.f;
                        reward = (float)cnt; term[0] = (float)cnt;
                    }
                    const float *aux = sg + 72;
                    if (s_contact != 0.f) { const float c = s_contact * aux[0]; reward += c; term[1] = c; }
                    const float ax = aux[1], ay = aux[2], var = aux[3];
                    if (s_move != 0.f) {
                        if (aux[4] != 0.f) {
                            float xm = ax - aux[5];
                            if (aux[7] != 0.f) xm = 0.f;
                            const float r = s_move * xm;
                            reward += r; term[2] = r;
                        }
                        w.last[e * 2] = ax; w.last[e * 2 + 1] = ay;
                        w.has_last[e] = 1;
                    }
                    if (s_mixed != 0.f) {
                        float sum = 0.f;
                        for (int n = 0; n < P; n++) {
                            const float dx = sx[2 * n] - gx, dy = sx[2 * n + 1] - gy;
                            float mval = expf(-sqrtf(dx * dx + dy * dy) / 2.f) * s_mixed;
                            if (sx[2 * n] >= gx) mval = s_mixed;
                            sum += mval;
                        }
                        reward += sum; term[3] = sum;
                    }
                    if (s_lin != 0.f || s_exp != 0.f) {
                        const float r = s_lin * (var - 1.f) + s_exp * expf(var / 2.f - 1.f);
                        reward += r; term[4] = r;
                    }
                    w.delayed_reset[e] = (unsigned char)(aux[8] != 0.f);
                } else {
                    w.has_last[e] = 0;
                }
            } else if (w.kind == MQE_WRAP_SEESAW && mode == 0) {
                const float s_x = w.scale[0], s_h = w.scale[1], s_y = w.scale[2], s_contact = w.scale[3], s_dist = w.scale[4], s_success = w.scale[5], s_fall = w.scale[6];
                const float *aux = sg + 72;
                if (s_x != 0.f) {
                    float xr = 0.f;
                    for (int a = 0; a < Aw; a++) {
                        const float x = BI(a, 0);
                        if (aux[3] != 0.f) xr += x - aux[4 + a];
                        w.last[e * Aw + a] = x;
                    }
                    w.has_last[e] = 1;
                    if (aux[1] != 0.f) xr = 0.f;
                    xr *= s_x;
                    reward += xr; term[0] = xr;
                }
                if (s_h != 0.f) { float z = 0.f; for (int a = 0; a < Aw; a++) z += BI(a, 2); const float r = s_h * (z - 0.56f); reward += r; term[1] = r; }
                if (s_y != 0.f) { float y2 = 0.f; for (int a = 0; a < Aw; a++) y2 += BI(a, 1) * BI(a, 1); const float r = s_y * (y2 - 0.5f); reward += r; term[2] = r; }
                if (s_contact != 0.f) { const float c = s_contact * aux[0]; reward += c; term[3] = c; }
                if (s_dist != 0.f) {
                    const float dx = BI(0, 0) - BI(Aw - 1, 0), dy = BI(0, 1) - BI(Aw - 1, 1), d2 = dx * dx + dy * dy;
                    if (d2 < 0.25f) { const float r = s_dist / fmaxf(d2, 1e-12f); reward += r; term[4] = r; }
                }
                if (s_success != 0.f) {
                    int cnt = 0;
                    for (int a = 0; a < Aw; a++) cnt += (BI(a, 0) > 7.7f) && (BI(a, 2) > 1.3f);
                    const float r = s_success * (float)cnt;
                    reward += r; term[5] = r;
                }
                if (s_fall != 0.f && aux[2] != 0.f) { reward += s_fall; term[6] = s_fall; }
            } else if (w.kind == MQE_WRAP_FOOTBALL_DEFENDER && mode == 0) {
                const float bx = sg[24], by = sg[25];
                const float s_goal = w.scale[0], s_dist = w.scale[1];
                const float gx = sg[72], gy = sg[73];                       // the reference compares the env-relative ball x with the WORLD gate x (:77)
                if (s_goal != 0.f && bx > gx) { reward += s_goal; term[0] = s_goal; }
                if (s_dist != 0.f) {
                    const float dx = bx - gx, dy = by - gy;
                    const float rr = s_dist * expf(-sqrtf(dx * dx + dy * dy) / 3.f);
                    reward += rr; term[1] = rr;
                }
            }
            else if (w.kind == MQE_WRAP_PUSHBOX) {
                // reward = scale * (box x - box x of the previous step), zero on the step of a reset; the wrapper's reset() forgets the last
                // position (`last_box_pos = None`), every step stores it whatever the scale
                if (mode == 0) {
                    const float *aux = sg + 72;
                    const float bx = sg[26], s_move = w.scale[0];
                    if (s_move != 0.f && aux[0] != 0.f) {
                        float xm = bx - aux[1];
                        if (aux[2] != 0.f) xm = 0.f;
                        const float r = s_move * xm;
                        reward += r; term[0] = r;
                    }
                    w.last[e] = bx; w.has_last[e] = 1;
                } else w.has_last[e] = 0;
            }
            else if (w.kind == MQE_WRAP_WRESTLING && mode == 0) {
                // flipped: |pitch| > 0.9 pi or |roll| >= 0.4 pi with both angles wrapped to (-pi, pi] (go1_wrestling_wrapper.py:66-72)
                const float PI = 3.14159265358979323846f;
                bool fl[2];
                for (int a = 0; a < 2; a++) {
                    float r = BI(a, 3), pt = BI(a, 4);
                    if (r > PI) r -= 2.f * PI;
                    if (pt > PI) pt -= 2.f * PI;
                    fl[a] = fabsf(pt) > PI * 0.9f || fabsf(r) >= PI * 0.4f;
                }
                if (w.scale[0] != 0.f) { const float r = fl[1] ? w.scale[0] : 0.f; reward += r; term[0] = r; }
                if (w.scale[1] != 0.f) { const float r = fl[0] ? w.scale[1] : 0.f; reward -= r; term[1] = r; }
            } else if (w.kind == MQE_WRAP_BRIDGE) {
                if (mode == 0) {                             // success: the opponent fell off (z < 0.5), punishment: agent 0 did, target: agent 0 is past the opponent's start x
                    const float *aux = sg + 72;
                    if (w.scale[0] != 0.f) { const float r = BI(1, 2) < 0.5f ? w.scale[0] : 0.f; reward += r; term[0] = r; }
                    if (w.scale[1] != 0.f) { const float r = BI(0, 2) < 0.5f ? w.scale[1] : 0.f; reward -= r; term[1] = r; }
                    if (w.scale[2] != 0.f) { const float r = BI(0, 0) > aux[2] ? w.scale[2] : 0.f; reward += r; term[2] = r; }
                } else { w.last[e * 4 + 2] = BI(1, 0); w.last[e * 4 + 3] = fabsf(BI(1, 0) + BI(0, 0)); }     // target_pos = flip(base_pos) at reset() (:28-31)
            } else if (w.kind == MQE_WRAP_ROTATION) {
                const float T = w.scale[3];
                if (mode == 0) {
                    const float *aux = sg + 72;
                    if (w.scale[0] != 0.f) { const float r = BI(0, 0) > T ? w.scale[0] : 0.f; reward += r; term[0] = r; }
                    if (w.scale[1] != 0.f) { const float r = BI(1, 0) > T ? w.scale[1] : 0.f; reward -= r; term[1] = r; }
                    if (w.scale[2] != 0.f) {                 // the reference subtracts the target x from BOTH x and y here (:77), from x only in reset() (:38-39)
                        const float dx = BI(0, 0) - T, dy = BI(0, 1) - T, dis = sqrtf(dx * dx + dy * dy);
                        const float r = dis < aux[0] ? w.scale[2] : 0.f;
                        reward += r; term[2] = r;
                        w.last[e * 4] = dis;
                    }
                } else { const float dx = BI(0, 0) - T, dy = BI(0, 1); w.last[e * 4] = sqrtf(dx * dx + dy * dy); }
            }
#undef BI
            if (mode == 0)
                for (int a = 0; a < Aw; a++) w.reward[(size_t)e * Aw + a] = (duel && a > 0) ? 0.f : reward;
        }
    }
    if (mode != 0) return;
    // ---- running sums of the reward terms (reward_buffer[...] of the reference, accumulated without a host sync) ----
    if (lane == 0)
#pragma unroll
        for (int t = 0; t < 8; t++) red[t][warp] = term[t];
    __syncthreads();
    if (threadIdx.x < 8) {
        double s = 0.0;
        for (int k = 0; k < WRAP_THREADS / 32; k++) s += (double)red[threadIdx.x][k];
        if (s != 0.0) atomicAdd(w.sums + threadIdx.x, s);
    }
    if (blockIdx.x == 0 && threadIdx.x == 8) atomicAdd(w.sums + 8, 1.0);       // step count
}

extern "C" cudaError_t mqe_launch_task_gather(const DevParams &p, const WrapParams &w, int mode, cudaStream_t st) {
    const int envs_per_block = WRAP_THREADS / 32;
    return launch_heavy(k_task_gather, dim3((p.N + envs_per_block - 1) / envs_per_block), dim3(WRAP_THREADS), 0, st, p, w, mode);
}
