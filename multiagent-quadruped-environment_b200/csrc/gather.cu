// gather.cu -- the per-step exchange of what the learner reads (task-wrapper observation, reward, done flags) between the ranks of
// one node, as the LAST KERNEL OF THE STEP GRAPH: plain stores into every peer's receive buffer over NVLink / NVSwitch peer memory,
// one system-scope flag per (rank, parity), and a bounded wait for the peers' flags.  Replaces the two un-captured NCCL all-gathers
// of round 1 (the data the reference's learner reads: openrl_ws/utils.py:53-67).
//
// Layout of a rank's receive buffer (cudaMalloc + cudaIpcGetMemHandle, mapped by every peer):
//     half 0 | half 1 | u32 flags[2][MQE_MAX_RANKS]
//     half   = obs region [world][obs bytes] | reward region [world][reward bytes] | done region [world][done bytes]
// so after an exchange every rank holds the GLOBAL tensors contiguously, in rank (= global env) order.
// Exchange number q = 1, 2, ... (device counter ctr[5]) uses half q & 1 and flag value q.  Double buffering is enough: a peer can
// only start exchange q + 2 after it has seen this rank's flag q + 1, which this rank publishes after -- in stream order -- everything
// that read half q & 1.
#include "common.cuh"
#include "kernels.cuh"

#define GX_THREADS 256
#define GX_MAX_BLOCKS 64

__device__ __forceinline__ void st_release_sys(unsigned int *addr, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *addr) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}

__global__ void __launch_bounds__(GX_THREADS) k_gather_exchange(DevParams p, GatherParams g) {
    pdl_launch_dependents();
    pdl_wait();
    const unsigned int seq = (unsigned int)p.ctr[5] + 1u;       // every block reads it before the last block advances it
    const int parity = (int)(seq & 1u);
    __shared__ int s_last;
    // ---- push: this rank's three fields into slot `rank` of the three regions of every peer's half (16-byte stores) ----
    for (int k = 0; k < 3; k++) {
        const long long n16 = g.seg_bytes[k] >> 4;
        const uint4 *src = reinterpret_cast<const uint4 *>(g.src + (long long)(p.ctr[1] & 1) * p.result_half + g.seg_src[k]);   // the half just written
        for (int r = 0; r < g.world; r++) {
            const int peer = (g.rank + r) % g.world;             // start with the own buffer, spread the peers over time
            uint4 *dst = reinterpret_cast<uint4 *>(g.peer[peer] + (long long)parity * g.parity_bytes + g.seg_dst[k] + (long long)g.rank * g.seg_bytes[k]);
            for (long long i = (long long)blockIdx.x * GX_THREADS + threadIdx.x; i < n16; i += (long long)gridDim.x * GX_THREADS) dst[i] = src[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(g.blocks_done, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    // ---- last block: publish this rank's flag to every peer, then wait for every peer's flag in the local buffer ----
    __threadfence_system();
    if ((int)threadIdx.x < g.world) {
        const int r = threadIdx.x;
        unsigned int *theirs = reinterpret_cast<unsigned int *>(g.peer[r] + g.flags_off) + parity * MQE_MAX_RANKS + g.rank;
        st_release_sys(theirs, seq);
        const unsigned int *mine = reinterpret_cast<const unsigned int *>(g.peer[g.rank] + g.flags_off) + parity * MQE_MAX_RANKS + r;
        long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        // flags only grow; >= also accepts a peer that is already one exchange ahead on this parity (cannot happen, see header)
        while ((int)(ld_acquire_sys(mine) - seq) < 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 2000000000ll) { atomicExch(p.stats + MQE_STAT_GATHER_TIMEOUT, 1 + r); break; }   // bounded: never hangs the device
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { *g.blocks_done = 0; p.ctr[5] = (int)seq; }
}

extern "C" cudaError_t mqe_launch_gather_exchange(const DevParams &p, const GatherParams &g, cudaStream_t st) {
    long long n16 = 0;
    for (int k = 0; k < 3; k++) n16 += g.seg_bytes[k] >> 4;
    int blocks = (int)((n16 + GX_THREADS - 1) / GX_THREADS);
    blocks = blocks < 1 ? 1 : (blocks > GX_MAX_BLOCKS ? GX_MAX_BLOCKS : blocks);
    return launch_heavy(k_gather_exchange, dim3(blocks), dim3(GX_THREADS), 0, st, p, g);
}
