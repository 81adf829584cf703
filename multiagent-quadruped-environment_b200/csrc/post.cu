// post.cu -- everything Go1.step() does after the decimation loop, one warp per (environment, agent):
//   post_physics_step (legged_robot_field.py:117-119 -> legged_robot.py:117-157): derived base quantities, gait clock
//   (go1.py:240-279), termination (legged_robot.py:159-169 + legged_robot_field.py:121-146), NPC stepping
//   (go1_sheep.py:35-64), indexed reset (go1.py:110-145, legged_robot.py:394-470) and compute_observations
//   (go1.py:153-196), plus Go1.reset() (go1.py:147-151).
// The reference spends ~100 tiny torch launches, one nonzero() sync and 4 .cpu() syncs here; this is one launch and
// no host round trip (reset envs are handled in place from the reset mask, never gathered into an index list).
#include "post_dev.cuh"

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
    const int A = p.A;
    const int envs_per_block = POST_GROUPS / A;
    const int grp = threadIdx.x / POST_GS, gl = threadIdx.x % POST_GS;       // agent slot of the block, lane inside the group
    const unsigned gmask = ((1u << POST_GS) - 1u) << ((threadIdx.x & 31) & ~(POST_GS - 1));
    const int el = grp / A, a = grp % A;
    const int e = blockIdx.x * envs_per_block + el;
    const bool live = el < envs_per_block && e < p.N;
    if (threadIdx.x < POST_GROUPS) { s_flags[threadIdx.x] = 0; s_reset[threadIdx.x] = 0; }
    __syncthreads();
    if (live) {
        const int f = dev_post_agent_derive(p, e, a, gl, gmask, step_count);
        if (f && gl == 0) atomicOr(&s_flags[el], f);
    }
    __syncthreads();
    if (live && a == 0 && gl == 0) s_reset[el] = dev_post_env_decide(p, e, s_flags[el], step_count);
    __syncthreads();
    if (live && a == 0) dev_post_env_npc_reset(p, e, s_reset[el], gl, gmask, step_count);
    __syncthreads();                                     // reset wrote state / last_actions of every agent of the env
    if (live) dev_post_agent_finish(p, e, a, gl, gmask);
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
