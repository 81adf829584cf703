// policy_tc.cu -- layer 0 of the walk-these-ways networks on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// The only genuinely dense contraction on the Go1.step() path (go1.py:400-407): for every agent row the 30 x 70
// observation history against adapt.0 (256 x 2100) and body.0[:, :2100] (512 x 2100), fused into one
// [M x 2400(padded)] x [2400 x 768] GEMM = 89 % of the policy FLOPs.  fp32 parity is kept by splitting every operand
// into bf16 hi + lo and issuing three MMAs (hi*hi + hi*lo + lo*hi) into the same fp32 TMEM accumulator
// ("bf16x3"; the dropped lo*lo term is < 2^-16 relative).  passes == 1 runs plain bf16.
//
// Data layout (HBM):  both operands are stored pre-tiled in the tcgen05 no-swizzle K-major canonical layout, so one
// stage of the pipeline is four contiguous 20 KB bulk copies (cp.async.bulk -> UBLKCP), no tensor maps:
//     ring  : [M/128 row tiles][30 slots ][10 k-chunks][128 rows][8 bf16]     (hi and lo planes)
//     weight: [6 col tiles    ][30 blocks][10 k-chunks][128 rows][8 bf16]
// A core matrix (8 rows x 16 B) is contiguous (128 B); SBO = 128 B between row groups, LBO = 2048 B between k-chunks.
// The ring never shifts: slot s is contracted with weight block (s - head - 1) mod 30 (its age).
//
// CTA = 128 rows x 128 columns, 6 warps: warp 0 producer (bulk copies, mbarrier expect_tx), warp 1 MMA issuer
// (one elected lane; tcgen05.commit releases stages), warps 2..5 epilogue (tcgen05.ld 32x32b, bias + ELU, store Z).
#include <vector>

#include "kernels.cuh"

extern "C" cudaError_t mqe_launch_policy_tail_only(const PolicyTcWeights &w, const PolicyWeightsDev &pw, const PolicyScratch &s, const DevParams &p, int M,
                                                   int passes, int finish, cudaStream_t st);

#define TC_TILE_ELEMS (10 * 128 * 8)
#define TC_TILE_BYTES (TC_TILE_ELEMS * 2)
#define TC_STAGES 2
#define TC_STAGE_BYTES (4 * TC_TILE_BYTES)
#define TC_SMEM_BYTES (TC_STAGES * TC_STAGE_BYTES + 128)
#define TC_TMEM_COLS 128
#define TC_LBO 2048u
#define TC_SBO 128u

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(TC_LBO >> 4) << 16) | ((uint64_t)(TC_SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ELU(alpha = 1) without the ~30-instruction expm1f: degree-7 Taylor for -0.25 < x <= 0 (|err| < 2e-8 relative),
// ex2.approx-based exp(x) - 1 below that (result magnitude >= 0.22, so the absolute 1e-7 is <= 5e-7 relative).
__device__ __forceinline__ float elu1_tc(float x) {
    // ELU on the tensor-core path: exp through ex2.approx.  exp(t) - 1 loses RELATIVE accuracy for |t| << 1 but stays within 2.5e-7
    // ABSOLUTE (the value is about to be split into bf16 hi + lo and summed 128..512 wide against a 2e-4 parity bound), and costs 6
    // instructions where the expm1-grade polynomial + select used before cost 15 -- the A-operand producers of the fused tail are
    // bound by exactly this arithmetic (tools/tail_trace.py: 18.6 of 43 us).  The fp32 cross-check path (policy.cu) keeps expm1f.
    const float t = fminf(x, 0.f);
    const float neg = __expf(t) - 1.f;
    return x > 0.f ? x : neg;
}
// v = float(hi) + float(lo) + O(2^-17 |v|), both halves rounded to nearest even: two packed conversions per PAIR of values
// (cvt.rn.bf16x2.f32; same bits as the integer round-to-nearest-even k_policy_frame uses, for every finite value)
__device__ __forceinline__ uint32_t bf16x2_rn(float lo_half, float hi_half) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_half), "f"(lo_half));
    return r;
}
__device__ __forceinline__ void split_bf16x8(const float *v, uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        h[i] = bf16x2_rn(v[2 * i], v[2 * i + 1]);
        const float r0 = v[2 * i] - __uint_as_float(h[i] << 16), r1 = v[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u);
        l[i] = bf16x2_rn(r0, r1);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// element offset of (row m, column k) in an activation plane with `kchunks` = K/8 chunks per row tile
__device__ __forceinline__ size_t plane_index(int m, int k, int kchunks) {
    return (((size_t)(m >> 7) * kchunks + (k >> 3)) * 128 + (m & 127)) * 8 + (k & 7);
}


// float offset of (row r of row tile mt, 4-column group g4 of column tile nt) in Zold: [mt][6 column tiles][32 groups][128 rows][4]
__device__ __forceinline__ size_t zold_index(int mt, int nt, int g4, int r) {
    return ((((size_t)mt * 6 + nt) * 32 + g4) * 128 + r) * 4;
}
__global__ void __launch_bounds__(320, 1)
k_policy_l0_tc(const unsigned short *__restrict__ a_hi, const unsigned short *__restrict__ a_lo,
               const unsigned short *__restrict__ b_hi, const unsigned short *__restrict__ b_lo,
               const float *__restrict__ b0cat, float *__restrict__ Z, unsigned short *__restrict__ a0_hi,
               unsigned short *__restrict__ a0_lo, int M, int head, int passes, const int *__restrict__ ctr, int ntile0,
               int mode, float *__restrict__ Zold, const unsigned char *__restrict__ dirty, int agents, int mtile0, int ztiled, int ntpc) {
    // mode 0: all 30 slots.  Incremental layer 0 (29 of the 30 history frames of the NEXT step are known as soon as this step's frame is in
    // the ring): mode 1 = the 29 slots other than `head` (the slot the next frame will go to), raw accumulators -> Zold, launched at low
    // priority behind the physics of the step so that it fills the idle tail of k_substeps; mode 2 = slot `head` alone (K = 80), plus
    // Zold (dropped for rows whose history was just reset: `dirty`), then the normal epilogue.
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t *empty = full + TC_STAGES;
    uint64_t *accum = empty + TC_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ntpc > 1 (mode 2 only): this CTA does `ntpc` neighbouring column tiles of its row tile one after the other -- the new frame's A tile is
    // loaded once, every tile has its own 128 TMEM columns, and TMEM allocation / barrier set-up are paid once instead of per 128 x 128 tile
    if (mode != 2) ntpc = 1;
    const int ntile_first = blockIdx.x * ntpc + ntile0, mtile = blockIdx.y + mtile0;     // ntile0 / mtile0: first column / row tile of this launch
    const uint32_t tmem_cols = ntpc == 1 ? (uint32_t)TC_TMEM_COLS : 512u;
    pdl_launch_dependents();

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                                              // set-up above ran in the predecessor's shadow
    if (head == -2) head = (ctr[7] + 1) % MQE_HIST_FRAMES;  // graph replay, background pass: the slot after the one k_policy_frame wrote this step
    else if (head < 0) head = ctr[0];                        // graph replay: newest slot = the one k_policy_frame just wrote (mode 1: will write next)
    const uint32_t stage_tx = passes == 3 ? TC_STAGE_BYTES : 2 * TC_TILE_BYTES;
    const int n_iter = mode == 0 ? MQE_HIST_FRAMES : (mode == 1 ? MQE_HIST_FRAMES - 1 : ntpc);     // mode 2: one iteration per column tile

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < n_iter; i++) {
                const int it = mode == 0 ? i : (mode == 1 ? (i < head ? i : i + 1) : head);     // ring slot of this k-iteration
                const int st = i & 1, ph = (i >> 1) & 1;
                const bool load_a = mode != 2 || i == 0;             // mode 2: the A tile (the new frame) stays in stage 0 for every column tile
                const int ntile = mode == 2 ? ntile_first + i : ntile_first;
                mbar_wait(&empty[st], ph ^ 1);
                mbar_expect_tx(&full[st], load_a ? stage_tx : stage_tx / 2);
                int blk = it - head - 1;
                if (blk < 0) blk += MQE_HIST_FRAMES;
                unsigned char *sb = smem + st * TC_STAGE_BYTES;
                const size_t ao = ((size_t)mtile * MQE_HIST_FRAMES + it) * TC_TILE_ELEMS;
                const size_t bo = ((size_t)ntile * MQE_HIST_FRAMES + blk) * TC_TILE_ELEMS;
                if (load_a) bulk_g2s(sb, a_hi + ao, TC_TILE_BYTES, &full[st]);
                bulk_g2s(sb + 2 * TC_TILE_BYTES, b_hi + bo, TC_TILE_BYTES, &full[st]);
                if (passes == 3) {
                    if (load_a) bulk_g2s(sb + TC_TILE_BYTES, a_lo + ao, TC_TILE_BYTES, &full[st]);
                    bulk_g2s(sb + 3 * TC_TILE_BYTES, b_lo + bo, TC_TILE_BYTES, &full[st]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B bf16 K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            for (int it = 0; it < n_iter; it++) {
                const int st = it & 1, ph = (it >> 1) & 1;
                mbar_wait(&full[st], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sb = smem_u32(smem + st * TC_STAGE_BYTES);
                const uint32_t sa = mode == 2 ? smem_u32(smem) : sb;                 // mode 2: A lives in stage 0
                const uint32_t d_tmem = mode == 2 ? tmem + 128u * (uint32_t)it : tmem;  // mode 2: one accumulator per column tile
                const uint64_t dAh = umma_desc(sa), dAl = umma_desc(sa + TC_TILE_BYTES);
                const uint64_t dBh = umma_desc(sb + 2 * TC_TILE_BYTES), dBl = umma_desc(sb + 3 * TC_TILE_BYTES);
#pragma unroll
                for (int j = 0; j < 5; j++) umma_f16(d_tmem, dAh + j * 256, dBh + j * 256, idesc, ((mode == 2 ? 0 : it) | j) ? 1u : 0u);
                if (passes == 3) {
#pragma unroll
                    for (int j = 0; j < 5; j++) umma_f16(d_tmem, dAh + j * 256, dBl + j * 256, idesc, 1u);
#pragma unroll
                    for (int j = 0; j < 5; j++) umma_f16(d_tmem, dAl + j * 256, dBh + j * 256, idesc, 1u);
                }
                umma_commit(&empty[st]);             // implies tcgen05.fence::before_thread_sync
            }
            umma_commit(accum);
        }
        __syncwarp();
    } else {
        mbar_wait(accum, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;             // two epilogue warps per quarter, two 32-column chunks each
        const int row = mtile * 128 + q * 32 + lane;
#pragma unroll 1
        for (int tt = 0; tt < 2 * ntpc; tt++) {
            const int t = tt >> 1, c = 2 * half + (tt & 1);     // column tile of this CTA, 32-column chunk inside it
            const int ntile = ntile_first + t;
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 128 + c * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int n0 = ntile * 128 + c * 32;
            if (mode == 1) {                                // raw partial sums of the 29 known frames
                // Zold is private to these two passes, so it is stored the way its threads touch it: [row tile][column tile][4-column group]
                // [128 rows][4] -- the 32 lanes of a warp (32 consecutive rows) then write / read 512 contiguous bytes per instruction
                // instead of one half-used 32-byte sector per lane (row-major [M][768])
                {
                    float *dst = Zold + zold_index(mtile, ntile, c * 8, q * 32 + lane);
#pragma unroll
                    for (int i = 0; i < 32; i += 4) *reinterpret_cast<uint4 *>(dst + (size_t)(i / 4) * 512) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
                continue;
            }
            if (mode == 2 && row < M && !dirty[row / agents]) {      // + the 29 older frames (a row whose history was reset keeps only the new frame)
                const float *src = Zold + zold_index(mtile, ntile, c * 8, q * 32 + lane);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 z = *reinterpret_cast<const float4 *>(src + (size_t)(i / 4) * 512);
                    v[i] = __float_as_uint(__uint_as_float(v[i]) + z.x); v[i + 1] = __float_as_uint(__uint_as_float(v[i + 1]) + z.y);
                    v[i + 2] = __float_as_uint(__uint_as_float(v[i + 2]) + z.z); v[i + 3] = __float_as_uint(__uint_as_float(v[i + 3]) + z.w);
                }
            }
            if (n0 < 256 && a0_hi) {                        // adapt.0: bias + ELU, emitted as bf16 hi/lo planes for adapt.2
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; i++) f[i] = elu1_tc(__uint_as_float(v[i]) + __ldg(b0cat + n0 + i));
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    uint4 hi, lo;
                    split_bf16x8(f + 8 * g, hi, lo);
                    const size_t o = plane_index(row, n0 + 8 * g, 32);
                    *reinterpret_cast<uint4 *>(a0_hi + o) = hi;
                    *reinterpret_cast<uint4 *>(a0_lo + o) = lo;
                }
            } else if (row < M || ztiled) {
                // ztiled (the fused tail follows): Z in the tile-major layout of Zold, so that this store and the tail's loads are coalesced
                float *dst = ztiled ? Z + zold_index(mtile, ntile, c * 8, q * 32 + lane) : Z + (size_t)row * 768 + n0;
                const size_t gstride = ztiled ? 512 : 4;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float4 o;
                    float t0 = __uint_as_float(v[i]) + __ldg(b0cat + n0 + i), t1 = __uint_as_float(v[i + 1]) + __ldg(b0cat + n0 + i + 1);
                    float t2 = __uint_as_float(v[i + 2]) + __ldg(b0cat + n0 + i + 2), t3 = __uint_as_float(v[i + 3]) + __ldg(b0cat + n0 + i + 3);
                    if (n0 < 256) { t0 = elu1_tc(t0); t1 = elu1_tc(t1); t2 = elu1_tc(t2); t3 = elu1_tc(t3); }
                    o.x = t0; o.y = t1; o.z = t2; o.w = t3;
                    *reinterpret_cast<float4 *>(dst + (size_t)(i / 4) * gstride) = o;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------- generic tensor-core layer
// Y = act(A W^T + bias) with BOTH operands arriving as pre-tiled bf16 hi/lo planes by bulk copy (the layer-0 recipe):
//   A planes: [M/128 row tiles][K/8 k-chunks][128 rows][8 bf16]    written by the previous layer's epilogue
//   W planes: [N/NCTA col tiles][K/64][8 k-chunks][NCTA rows][8 bf16]
// Output either fp32 row-major (the two network heads: latent, action) or bias + ELU'd hi/lo planes for the next layer.
// Activations never exist in fp32 in memory between layers; the only fp32 intermediate is body.0's pre-activation,
// which has to wait for the latent (k_body_latent_planes).
#define LT_BK 64
#define LT_STAGES 3
#define LT_A_PLANE (128 * LT_BK * 2)                       // 16 KB
__host__ __device__ constexpr int lt_stage_bytes(int ncta) { return 2 * LT_A_PLANE + 2 * ncta * LT_BK * 2; }
__host__ __device__ constexpr int lt_smem_bytes(int ncta) { return LT_STAGES * lt_stage_bytes(ncta) + 128; }

#define LT_THREADS 320                                       // producer warp, MMA warp, 8 epilogue warps
// Fused network heads.  HEAD > 0: the layer that follows (adapt.4: 128 -> 2, body.6: 128 -> 12) is so small that the epilogue
// evaluates it in fp32 straight from the accumulator row each thread already holds, instead of writing planes for a 16-column
// tensor-core launch.  LATENT: with the latent in hand the same CTA also finishes body.0 for its 128 rows,
// planes(ELU(Z_body + W_lat latent)) (go1.py:404-406), which used to be a kernel of its own.
struct HeadArgs {
    const float *hw, *hb;          // [HEAD][128] row-major fp32, [HEAD]
    float *hy;                     // [M][HEAD]
    const float *Z, *wlat;         // LATENT: layer-0 output [M][768], latent columns of body.0 [512][2]
    unsigned short *l_hi, *l_lo;   // LATENT: body.0 activation planes (64 k-chunks)
};
template <int NCTA, int HEAD = 0, bool LATENT = false>
__global__ void __launch_bounds__(LT_THREADS, 1)
k_linear_tc(const unsigned short *__restrict__ a_hi, const unsigned short *__restrict__ a_lo, int K,
            const unsigned short *__restrict__ w_hi, const unsigned short *__restrict__ w_lo, const float *__restrict__ bias,
            float *__restrict__ Y, int ldy, int n_valid,                       // fp32 output (heads) when Y != nullptr
            unsigned short *__restrict__ o_hi, unsigned short *__restrict__ o_lo, int out_kchunks,   // plane output otherwise
            int M, int elu, int passes, HeadArgs ha) {
    extern __shared__ __align__(1024) unsigned char smem[];
    static_assert(HEAD == 0 || NCTA == 128, "fused heads read a full 128-wide activation row");
    constexpr int STAGE = lt_stage_bytes(NCTA);
    constexpr int W_PLANE = NCTA * LT_BK * 2;
    constexpr uint32_t TCOLS = NCTA < 32 ? 32 : NCTA;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + LT_STAGES * STAGE);
    uint64_t *empty = full + LT_STAGES;
    uint64_t *accum = empty + LT_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum + 1);
    __shared__ float s_bias[NCTA];
    __shared__ float s_hw[HEAD > 0 ? HEAD * 128 : 1], s_part[HEAD > 0 ? HEAD * 128 : 1], s_lat[LATENT ? 256 : 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntile = blockIdx.x, mtile = blockIdx.y;
    pdl_launch_dependents();
    const int nk = K / LT_BK;
    for (int i = threadIdx.x; i < NCTA; i += blockDim.x) s_bias[i] = (ntile * NCTA + i < n_valid) ? bias[ntile * NCTA + i] : 0.f;   // constants
    if (HEAD > 0)
        for (int i = threadIdx.x; i < HEAD * 128; i += blockDim.x) s_hw[i] = ha.hw[i];

    if (threadIdx.x == 0) {
        for (int i = 0; i < LT_STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                                              // set-up above ran in the predecessor's shadow

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nk; it++) {
                const int st = it % LT_STAGES, ph = (it / LT_STAGES) & 1;
                mbar_wait(&empty[st], ph ^ 1);
                mbar_expect_tx(&full[st], passes == 3 ? 2 * (LT_A_PLANE + W_PLANE) : (LT_A_PLANE + W_PLANE));
                unsigned char *sb = smem + st * STAGE;
                const size_t ao = ((size_t)mtile * nk + it) * (128 * LT_BK);
                const size_t wo = ((size_t)ntile * nk + it) * (NCTA * LT_BK);
                bulk_g2s(sb, a_hi + ao, LT_A_PLANE, &full[st]);
                bulk_g2s(sb + 2 * LT_A_PLANE, w_hi + wo, W_PLANE, &full[st]);
                if (passes == 3) {
                    bulk_g2s(sb + LT_A_PLANE, a_lo + ao, LT_A_PLANE, &full[st]);
                    bulk_g2s(sb + 2 * LT_A_PLANE + W_PLANE, w_lo + wo, W_PLANE, &full[st]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NCTA >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t hiA = ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
            const uint64_t hiB = ((uint64_t)((NCTA * 16u) >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
            for (int it = 0; it < nk; it++) {
                const int st = it % LT_STAGES, ph = (it / LT_STAGES) & 1;
                mbar_wait(&full[st], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sb = smem_u32(smem + st * STAGE);
                const uint64_t dAh = hiA | ((sb & 0x3FFFFu) >> 4), dAl = hiA | (((sb + LT_A_PLANE) & 0x3FFFFu) >> 4);
                const uint64_t dBh = hiB | (((sb + 2 * LT_A_PLANE) & 0x3FFFFu) >> 4), dBl = hiB | (((sb + 2 * LT_A_PLANE + W_PLANE) & 0x3FFFFu) >> 4);
                constexpr uint32_t aStep = (2 * 2048) >> 4, bStep = (2 * NCTA * 16) >> 4;      // two 8-wide k-chunks per MMA
#pragma unroll
                for (int j = 0; j < LT_BK / 16; j++) umma_f16(tmem, dAh + j * aStep, dBh + j * bStep, idesc, (it | j) ? 1u : 0u);
                if (passes == 3) {
#pragma unroll
                    for (int j = 0; j < LT_BK / 16; j++) umma_f16(tmem, dAh + j * aStep, dBl + j * bStep, idesc, 1u);
#pragma unroll
                    for (int j = 0; j < LT_BK / 16; j++) umma_f16(tmem, dAl + j * aStep, dBh + j * bStep, idesc, 1u);
                }
                umma_commit(&empty[st]);
            }
            umma_commit(accum);
        }
        __syncwarp();
    } else {
        mbar_wait(accum, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                              // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;                    // the two epilogue warps of a quarter split the columns
        const int orow = mtile * 128 + q * 32 + lane;
        constexpr int NCH = NCTA / 16, CH0 = (NCH + 1) / 2;
        float hp[HEAD > 0 ? HEAD : 1];
#pragma unroll
        for (int j = 0; j < (HEAD > 0 ? HEAD : 1); j++) hp[j] = 0.f;
#pragma unroll 1
        for (int c = half ? CH0 : 0; c < (half ? NCH : CH0); c++) {
            uint32_t v[16];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int n0 = ntile * NCTA + c * 16;
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                float t = __uint_as_float(v[i]) + s_bias[c * 16 + i];
                f[i] = elu ? elu1_tc(t) : t;
            }
            if (HEAD > 0) {
#pragma unroll
                for (int j = 0; j < HEAD; j++)
#pragma unroll
                    for (int i = 0; i < 16; i++) hp[j] = fmaf(f[i], s_hw[j * 128 + c * 16 + i], hp[j]);
            } else if (Y) {
                if (orow < M)
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        if (n0 + i < n_valid) Y[(size_t)orow * ldy + n0 + i] = f[i];
            } else {                                         // rows >= M of the last tile are written too (planes are padded)
                uint4 hi, lo;
                split_bf16x8(f, hi, lo);
                size_t o = plane_index(orow, n0, out_kchunks);
                *reinterpret_cast<uint4 *>(o_hi + o) = hi; *reinterpret_cast<uint4 *>(o_lo + o) = lo;
                split_bf16x8(f + 8, hi, lo);
                o = plane_index(orow, n0 + 8, out_kchunks);
                *reinterpret_cast<uint4 *>(o_hi + o) = hi; *reinterpret_cast<uint4 *>(o_lo + o) = lo;
            }
        }
        if (HEAD > 0) {                                      // the two warps of a lane quarter each hold half of the row's dot products
            const int r = q * 32 + lane;
            if (half == 1)
#pragma unroll
                for (int j = 0; j < HEAD; j++) s_part[r * HEAD + j] = hp[j];
            asm volatile("bar.sync 2, 256;" ::: "memory");   // the 8 epilogue warps
            if (half == 0) {
#pragma unroll
                for (int j = 0; j < HEAD; j++) {
                    const float o = (hp[j] + s_part[r * HEAD + j]) + __ldg(ha.hb + j);
                    if (orow < M) ha.hy[(size_t)orow * HEAD + j] = o;
                    if (LATENT && j < 2) s_lat[r * 2 + j] = o;
                }
            }
            if (LATENT) {
                asm volatile("bar.sync 2, 256;" ::: "memory");
                // one (row, 8-column chunk) per thread and trip; consecutive threads take consecutive rows so plane stores coalesce
                for (int t = (int)threadIdx.x - 64; t < 128 * 64; t += 256) {
                    const int chunk = t >> 7, m = t & 127, row = mtile * 128 + m;
                    float v[8];
                    if (row < M) {
                        const float *z = ha.Z + (size_t)row * 768 + 256 + chunk * 8;
                        const float4 a = *reinterpret_cast<const float4 *>(z), b = *reinterpret_cast<const float4 *>(z + 4);
                        const float l0 = s_lat[m * 2], l1 = s_lat[m * 2 + 1];
                        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int n = chunk * 8 + i;
                            v[i] = elu1_tc(v[i] + __ldg(ha.wlat + n * 2) * l0 + __ldg(ha.wlat + n * 2 + 1) * l1);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; i++) v[i] = 0.f;
                    }
                    uint4 hi, lo;
                    split_bf16x8(v, hi, lo);
                    const size_t o = plane_index(row, chunk * 8, 64);
                    *reinterpret_cast<uint4 *>(ha.l_hi + o) = hi;
                    *reinterpret_cast<uint4 *>(ha.l_lo + o) = lo;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS) : "memory");
    }
}

// body.0 tail: planes(ELU(Zbody + Wlat latent)) -- the two latent columns of cat(h, latent) (go1.py:404-406).
// One thread per (row, 8-column chunk); consecutive threads take consecutive rows so plane stores coalesce.
__global__ void k_body_latent_planes(const float *__restrict__ Z, const float *__restrict__ latent, const float *__restrict__ wlat,
                                     unsigned short *__restrict__ o_hi, unsigned short *__restrict__ o_lo, int M, int Mpad) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Mpad * 64) return;
    const int chunk = t / Mpad, m = t % Mpad;          // 64 chunks of 8 columns
    float v[8];
    if (m < M) {
        const float *z = Z + (size_t)m * 768 + 256 + chunk * 8;
        float4 a = *reinterpret_cast<const float4 *>(z), b = *reinterpret_cast<const float4 *>(z + 4);
        const float l0 = latent[(size_t)m * 2], l1 = latent[(size_t)m * 2 + 1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int n = chunk * 8 + i;
            v[i] = elu1_tc(v[i] + __ldg(wlat + n * 2) * l0 + __ldg(wlat + n * 2 + 1) * l1);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = 0.f;
    }
    uint4 hi, lo;
    split_bf16x8(v, hi, lo);
    const size_t o = plane_index(m, chunk * 8, 64);
    *reinterpret_cast<uint4 *>(o_hi + o) = hi;
    *reinterpret_cast<uint4 *>(o_lo + o) = lo;
}

// ---------------------------------------------------------------------------------------------- fused policy tail
// Everything behind layer 0 for one 128-row tile in ONE kernel (SURVEY 7 hard part 4):
//     adapt.2 (256 -> 128, ELU) -> adapt.4 (128 -> 2) = latent                                  (go1.py:400-403)
//     body.0 finish: ELU(Z_body + W_lat latent) -> body.2 (512 -> 256, ELU) -> body.4 (256 -> 128, ELU) -> body.6 (128 -> 12)   (go1.py:404-407)
//     last_locomotion_action(s) shift + clip to +-clip_actions                                  (go1.py:40-41, 104-106)
// Activations never leave the SM: the A operand of body.2 is produced chunk by chunk (64 columns) by the epilogue warps straight into the
// pipeline stage from Z and the latent, the A operand of body.4 from body.2's TMEM accumulator; only the weights stream in by bulk copy.
// TMEM: adapt.2 accumulator cols [0,128), body.2 [128,384), body.4 [384,512).  Sixteen stage fills through a 2-stage ring:
//     fills 0..3   adapt.2  : A planes (hi, lo) + W planes by bulk copy                         64 KB
//     fills 4..11  body.2   : W planes of both 128-column tiles by bulk copy (64 KB), A chunk written by the epilogue warps (32 KB)
//     fills 12..15 body.4   : W planes by bulk copy (32 KB), A chunk written by the epilogue warps from TMEM (32 KB)
// Barriers per stage: full_w (weights landed, tx count), full_a (8 epilogue warps wrote the A chunk), empty (MMAs that read it retired).
#define FT_STAGE_BYTES (96 * 1024)
#define FT_EPI_WARPS 16                                   // epilogue / A-operand producer warps: four per TMEM lane quarter
#define FT_THREADS (64 + 32 * FT_EPI_WARPS)
struct TailArgs {
    const unsigned short *a0_hi, *a0_lo;                 // adapt.0 activation planes (K = 256), from layer 0's epilogue
    const float *Z;                                      // layer-0 output, TILE-MAJOR (zold_index; k_policy_l0_tc with ztiled = 1); columns 256.. = body.0 pre-activation without the latent
    const unsigned short *wa_hi, *wa_lo, *wb1_hi, *wb1_lo, *wb2_hi, *wb2_lo;   // tail weights, pre-tiled (tile_layer, 128-row tiles)
    const float *ba1, *bb1, *bb2;                        // biases of adapt.2, body.2, body.4
    const float *aw2, *ab2, *bw3, *bb3, *wlat;           // heads (fp32): adapt.4 [2][128], body.6 [12][128]; latent columns of body.0 [512][2]
    float *latent, *act;                                 // [M][2], [M][12]
    int M, passes, finish;                               // finish: also do k_policy_finish's work (simulation step; not for mqe_policy_forward)
};
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__global__ void __launch_bounds__(FT_THREADS, 1) k_policy_tail(TailArgs a, DevParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full_w = reinterpret_cast<uint64_t *>(smem + 2 * FT_STAGE_BYTES);
    uint64_t *full_a = full_w + 2, *empty = full_a + 2, *acc_done = empty + 2;     // acc_done[3]: adapt.2, body.2, body.4
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_done + 3);
    float *sf = reinterpret_cast<float *>(smem + 2 * FT_STAGE_BYTES + 128);
    float *s_ba1 = sf, *s_bb1 = sf + 128, *s_bb2 = sf + 384, *s_hwA = sf + 512, *s_hwB = sf + 768, *s_lat = sf + 2304, *s_part = sf + 2560;   // .. + 4608: [128 rows][3 parts][12]
    float *s_wlat = sf + 7168;                              // [512][2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mtile = blockIdx.x, M = a.M;
    // MQE_TRACE=1: the first epilogue thread stamps the global timer at the phase boundaries into row `mtile` of the warp trace
    // (read it after a stand-alone mqe_sim_policy call: k_substeps reuses the buffer); tools/tail_trace.py
    long long *const ttr = (p.trace && p.warp_trace && threadIdx.x == 64) ? p.warp_trace + (size_t)mtile * MQE_TRACE_COLS : nullptr;
#define TAIL_MARK(k) if (ttr) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); ttr[k] = t_; }
    TAIL_MARK(0);
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) { mbar_init(&full_w[i], 1); mbar_init(&full_a[i], FT_EPI_WARPS); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 3; i++) mbar_init(&acc_done[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                                              // set-up above ran in the predecessor's shadow
    const int passes = a.passes;
    const uint32_t PL = 128 * LT_BK * 2;                     // one 128-row x 64-k plane: 16 KB

    if (warp == 0) {
        if (lane == 0) {
            for (int f = 0; f < 16; f++) {
                const int st = f & 1;
                mbar_wait(&empty[st], ((f >> 1) & 1) ^ 1);
                unsigned char *sb = smem + st * FT_STAGE_BYTES;
                if (f < 4) {                                 // adapt.2: A chunk f of the adapt.0 planes + W chunk f
                    mbar_expect_tx(&full_w[st], passes == 3 ? 4 * PL : 2 * PL);
                    const size_t ao = ((size_t)mtile * 4 + f) * (128 * LT_BK), wo = (size_t)f * (128 * LT_BK);
                    bulk_g2s(sb, a.a0_hi + ao, PL, &full_w[st]);
                    bulk_g2s(sb + 2 * PL, a.wa_hi + wo, PL, &full_w[st]);
                    if (passes == 3) { bulk_g2s(sb + PL, a.a0_lo + ao, PL, &full_w[st]); bulk_g2s(sb + 3 * PL, a.wa_lo + wo, PL, &full_w[st]); }
                } else if (f < 12) {                         // body.2: W chunk j of both 128-column tiles (tile_layer: [n-tile][k-chunk][..])
                    const int j = f - 4;
                    mbar_expect_tx(&full_w[st], passes == 3 ? 4 * PL : 2 * PL);
                    for (int nt = 0; nt < 2; nt++) {
                        const size_t wo = ((size_t)nt * 8 + j) * (128 * LT_BK);
                        bulk_g2s(sb + (2 + 2 * nt) * PL, a.wb1_hi + wo, PL, &full_w[st]);
                        if (passes == 3) bulk_g2s(sb + (3 + 2 * nt) * PL, a.wb1_lo + wo, PL, &full_w[st]);
                    }
                } else {                                     // body.4: W chunk j
                    const int j = f - 12;
                    mbar_expect_tx(&full_w[st], passes == 3 ? 2 * PL : PL);
                    const size_t wo = (size_t)j * (128 * LT_BK);
                    bulk_g2s(sb + 2 * PL, a.wb2_hi + wo, PL, &full_w[st]);
                    if (passes == 3) bulk_g2s(sb + 3 * PL, a.wb2_lo + wo, PL, &full_w[st]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);     // D fp32, A/B bf16 K-major, N = 128, M = 128
            const uint64_t hiD = ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);      // 128-row planes: LBO 2 KB, SBO 128 B
            constexpr uint32_t kStep = (2 * 2048) >> 4;      // two 8-wide k-chunks per MMA
            for (int f = 0; f < 16; f++) {
                const int st = f & 1;
                mbar_wait(&full_w[st], (f >> 1) & 1);
                if (f >= 4) mbar_wait(&full_a[st], ((f - 4) >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sb = smem_u32(smem + st * FT_STAGE_BYTES);
                auto D = [&](uint32_t off) { return hiD | (uint64_t)(((sb + off) & 0x3FFFFu) >> 4); };
                const uint64_t dAh = D(0), dAl = D(PL);
                const int ntiles = (f >= 4 && f < 12) ? 2 : 1;
                const int first = (f == 0 || f == 4 || f == 12);
                for (int nt = 0; nt < ntiles; nt++) {
                    const uint64_t dBh = D((2 + 2 * nt) * PL), dBl = D((3 + 2 * nt) * PL);
                    const uint32_t d_tmem = tmem + (f < 4 ? 0u : (f < 12 ? 128u + 128u * nt : 384u));
#pragma unroll
                    for (int j = 0; j < LT_BK / 16; j++) umma_f16(d_tmem, dAh + j * kStep, dBh + j * kStep, idesc, (first && j == 0) ? 0u : 1u);
                    if (passes == 3) {
#pragma unroll
                        for (int j = 0; j < LT_BK / 16; j++) umma_f16(d_tmem, dAh + j * kStep, dBl + j * kStep, idesc, 1u);
#pragma unroll
                        for (int j = 0; j < LT_BK / 16; j++) umma_f16(d_tmem, dAl + j * kStep, dBh + j * kStep, idesc, 1u);
                    }
                }
                umma_commit(&empty[st]);
                if (f == 3) umma_commit(&acc_done[0]);
                if (f == 11) umma_commit(&acc_done[1]);
                if (f == 15) umma_commit(&acc_done[2]);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3, part = (warp - 2) >> 2;       // TMEM lane quarter this warp may read; the four warps of a quarter split the columns
        const int r = q * 32 + lane, row = mtile * 128 + r;    // this thread's row of the tile
        const int et = (int)threadIdx.x - 64;                 // 0..511 among the epilogue threads
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        // biases, head weights and the latent columns of body.0 into shared memory -- by the epilogue warps only, while the producer and the
        // MMA warp already stream adapt.2 (these constants used to sit in front of the first bulk copy: 2 us of the kernel)
        for (int i = et; i < 128; i += 32 * FT_EPI_WARPS) { s_ba1[i] = a.ba1[i]; s_bb2[i] = a.bb2[i]; }
        for (int i = et; i < 256; i += 32 * FT_EPI_WARPS) { s_bb1[i] = a.bb1[i]; s_hwA[i] = a.aw2[i]; }
        for (int i = et; i < 1536; i += 32 * FT_EPI_WARPS) s_hwB[i] = a.bw3[i];
        for (int i = et; i < 1024; i += 32 * FT_EPI_WARPS) s_wlat[i] = a.wlat[i];
        asm volatile("bar.sync 2, %0;" ::"n"(32 * FT_EPI_WARPS) : "memory");
        // ---- adapt.2 epilogue: bias + ELU, head adapt.4 in fp32 from the accumulator row -> latent ----
        TAIL_MARK(1);
        mbar_wait(&acc_done[0], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        TAIL_MARK(2);
        {
            float h0 = 0.f, h1 = 0.f;
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
                uint32_t v[16];
                const int col = part * 32 + c * 16;
                tmem_ld16(trow + (uint32_t)col, v);
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float t = elu1_tc(__uint_as_float(v[i]) + s_ba1[col + i]);
                    h0 = fmaf(t, s_hwA[col + i], h0); h1 = fmaf(t, s_hwA[128 + col + i], h1);
                }
            }
            if (part) { s_part[(r * 3 + part - 1) * 2] = h0; s_part[(r * 3 + part - 1) * 2 + 1] = h1; }
            asm volatile("bar.sync 2, %0;" ::"n"(32 * FT_EPI_WARPS) : "memory");
            if (part == 0) {
                const float l0 = (((h0 + s_part[(r * 3) * 2]) + s_part[(r * 3 + 1) * 2]) + s_part[(r * 3 + 2) * 2]) + __ldg(a.ab2);
                const float l1 = (((h1 + s_part[(r * 3) * 2 + 1]) + s_part[(r * 3 + 1) * 2 + 1]) + s_part[(r * 3 + 2) * 2 + 1]) + __ldg(a.ab2 + 1);
                s_lat[r * 2] = l0; s_lat[r * 2 + 1] = l1;
                if (row < M) { a.latent[(size_t)row * 2] = l0; a.latent[(size_t)row * 2 + 1] = l1; }
            }
            asm volatile("bar.sync 2, %0;" ::"n"(32 * FT_EPI_WARPS) : "memory");
        }
        TAIL_MARK(3);
        // ---- body.2 A operand: ELU(Z_body + W_lat latent), one 64-column chunk per fill, written in the canonical K-major layout ----
        // this thread's part of Z (row m, 16 columns of every 64-column chunk) is fetched one chunk ahead of its use
        const int m = et & 127, grow = mtile * 128 + m, kc0 = (et >> 7) * 2;
        // Z arrives tile-major (k_policy_l0_tc, ztiled): body column n = 256 + j * 64 + kc0 * 8 + 4 i sits in column tile n / 128, group (n % 128) / 4
        auto zptr = [&](int j_, int i_) {
            const int n = 256 + j_ * 64 + kc0 * 8 + 4 * i_;
            return reinterpret_cast<const float4 *>(a.Z + zold_index(mtile, n >> 7, (n & 127) >> 2, m));
        };
        float4 zq[4];
#pragma unroll
        for (int i = 0; i < 4; i++) zq[i] = *zptr(0, i);
        const float l0 = s_lat[m * 2], l1 = s_lat[m * 2 + 1];
#pragma unroll 1
        for (int j = 0; j < 8; j++) {
            const int f = 4 + j, st = f & 1;
            float4 zc[4];
#pragma unroll
            for (int i = 0; i < 4; i++) zc[i] = zq[i];
            if (j + 1 < 8)
#pragma unroll
                for (int i = 0; i < 4; i++) zq[i] = *zptr(j + 1, i);
            mbar_wait(&empty[st], ((f >> 1) & 1) ^ 1);
            unsigned char *sb = smem + st * FT_STAGE_BYTES;
#pragma unroll
            for (int cc = 0; cc < 2; cc++) {
                const int kc = kc0 + cc;                      // 8-wide k-chunk inside the 64-column chunk
                float v[8];
                if (grow < M) {
                    const float4 z0 = zc[2 * cc], z1 = zc[2 * cc + 1];
                    v[0] = z0.x; v[1] = z0.y; v[2] = z0.z; v[3] = z0.w; v[4] = z1.x; v[5] = z1.y; v[6] = z1.z; v[7] = z1.w;
                    // latent columns of body.0 for these 8 outputs: four 16-byte broadcast loads instead of sixteen scalar ones (the
                    // producers stalled on the short scoreboard here: ncu warm-cache source view)
                    const float4 *w4 = reinterpret_cast<const float4 *>(s_wlat + (j * 64 + kc * 8) * 2);
#pragma unroll
                    for (int i2 = 0; i2 < 4; i2++) {
                        const float4 w = w4[i2];              // (w0, w1) of output 2 i2, (w0, w1) of output 2 i2 + 1
                        v[2 * i2] = elu1_tc(v[2 * i2] + w.x * l0 + w.y * l1);
                        v[2 * i2 + 1] = elu1_tc(v[2 * i2 + 1] + w.z * l0 + w.w * l1);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[i] = 0.f;
                }
                uint4 hi, lo;
                split_bf16x8(v, hi, lo);
                *reinterpret_cast<uint4 *>(sb + ((size_t)kc * 128 + m) * 16) = hi;
                *reinterpret_cast<uint4 *>(sb + PL + ((size_t)kc * 128 + m) * 16) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_a[st])) : "memory");
        }
        // ---- body.4 A operand: ELU(body.2 accumulator + bias) from TMEM, 64 columns per fill ----
        TAIL_MARK(4);
        mbar_wait(&acc_done[1], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        TAIL_MARK(5);
#pragma unroll 1
        for (int j = 0; j < 4; j++) {
            const int f = 12 + j, st = f & 1;
            mbar_wait(&empty[st], ((f >> 1) & 1) ^ 1);
            unsigned char *sb = smem + st * FT_STAGE_BYTES;
            {
                uint32_t v[16];
                const int col = j * 64 + part * 16;           // column of body.2's output = k index of body.4
                tmem_ld16(trow + 128u + (uint32_t)col, v);
                float fv[16];
#pragma unroll
                for (int i4 = 0; i4 < 4; i4++) {
                    const float4 b = *reinterpret_cast<const float4 *>(s_bb1 + col + 4 * i4);
                    fv[4 * i4] = elu1_tc(__uint_as_float(v[4 * i4]) + b.x); fv[4 * i4 + 1] = elu1_tc(__uint_as_float(v[4 * i4 + 1]) + b.y);
                    fv[4 * i4 + 2] = elu1_tc(__uint_as_float(v[4 * i4 + 2]) + b.z); fv[4 * i4 + 3] = elu1_tc(__uint_as_float(v[4 * i4 + 3]) + b.w);
                }
                uint4 hi, lo;
                const int kc = part * 2;
                split_bf16x8(fv, hi, lo);
                *reinterpret_cast<uint4 *>(sb + ((size_t)kc * 128 + r) * 16) = hi;
                *reinterpret_cast<uint4 *>(sb + PL + ((size_t)kc * 128 + r) * 16) = lo;
                split_bf16x8(fv + 8, hi, lo);
                *reinterpret_cast<uint4 *>(sb + ((size_t)(kc + 1) * 128 + r) * 16) = hi;
                *reinterpret_cast<uint4 *>(sb + PL + ((size_t)(kc + 1) * 128 + r) * 16) = lo;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_a[st])) : "memory");
        }
        // ---- body.4 epilogue: bias + ELU, head body.6 in fp32 -> action; then the shift / clip of k_policy_finish ----
        TAIL_MARK(6);
        mbar_wait(&acc_done[2], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        TAIL_MARK(7);
        {
            float hp[12];
#pragma unroll
            for (int o = 0; o < 12; o++) hp[o] = 0.f;
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
                uint32_t v[16];
                const int col = part * 32 + c * 16;
                tmem_ld16(trow + 384u + (uint32_t)col, v);
                float t[16];
#pragma unroll
                for (int i4 = 0; i4 < 4; i4++) {
                    const float4 b = *reinterpret_cast<const float4 *>(s_bb2 + col + 4 * i4);
                    t[4 * i4] = elu1_tc(__uint_as_float(v[4 * i4]) + b.x); t[4 * i4 + 1] = elu1_tc(__uint_as_float(v[4 * i4 + 1]) + b.y);
                    t[4 * i4 + 2] = elu1_tc(__uint_as_float(v[4 * i4 + 2]) + b.z); t[4 * i4 + 3] = elu1_tc(__uint_as_float(v[4 * i4 + 3]) + b.w);
                }
#pragma unroll
                for (int o = 0; o < 12; o++) {                // weights four at a time (same per-output summation order as one at a time)
                    const float4 *w4 = reinterpret_cast<const float4 *>(s_hwB + o * 128 + col);
#pragma unroll
                    for (int i4 = 0; i4 < 4; i4++) {
                        const float4 w = w4[i4];
                        hp[o] = fmaf(t[4 * i4], w.x, hp[o]); hp[o] = fmaf(t[4 * i4 + 1], w.y, hp[o]);
                        hp[o] = fmaf(t[4 * i4 + 2], w.z, hp[o]); hp[o] = fmaf(t[4 * i4 + 3], w.w, hp[o]);
                    }
                }
            }
            if (part)
#pragma unroll
                for (int o = 0; o < 12; o++) s_part[(r * 3 + part - 1) * 12 + o] = hp[o];
            asm volatile("bar.sync 2, %0;" ::"n"(32 * FT_EPI_WARPS) : "memory");
            if (part == 0 && row < M) {
#pragma unroll
                for (int o = 0; o < 12; o++) {
                    const float act = (((hp[o] + s_part[(r * 3) * 12 + o]) + s_part[(r * 3 + 1) * 12 + o]) + s_part[(r * 3 + 2) * 12 + o]) + __ldg(a.bb3 + o);
                    a.act[(size_t)row * 12 + o] = act;
                    if (a.finish) {                          // go1.py:104-106 + :40-41
                        const size_t t = (size_t)row * 12 + o;
                        p.loc_last2[t] = p.loc_last[t];
                        p.loc_last[t] = act;
                        p.actions[t] = fminf(fmaxf(act, -p.clip_actions), p.clip_actions);
                    }
                }
            }
        }
        TAIL_MARK(8);
        if (a.finish && blockIdx.x == 0) {                    // the rest of k_policy_finish: nothing reads these before the next kernel
            for (int t = et; t < p.N; t += 32 * FT_EPI_WARPS) p.hist_dirty[t] = 0;
            if (et < 5) p.stats[et] = 0;
            if (et == 0) p.ctr[0] = (p.ctr[0] + 1) % MQE_HIST_FRAMES;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}
#define FT_SMEM_BYTES (2 * FT_STAGE_BYTES + 128 + (2560 + 4608 + 1024) * 4)

// ---------------------------------------------------------------------------------------------- host side
static inline unsigned short f2bf_rne(float v) {
    uint32_t b;
    memcpy(&b, &v, 4);
    return (unsigned short)((b + 0x7fffu + ((b >> 16) & 1u)) >> 16);
}
static inline float bf2f(unsigned short h) {
    uint32_t b = (uint32_t)h << 16;
    float v;
    memcpy(&v, &b, 4);
    return v;
}

// [N][K] row-major fp32 -> bf16 hi/lo planes tiled [n-tile][k-chunk 64][8 sub-chunks][ncta rows][8], rows padded with zeros
static void tile_layer(const float *W, int N, int K, int ncta, std::vector<unsigned short> &hi, std::vector<unsigned short> &lo) {
    const int ntiles = (N + ncta - 1) / ncta, nk = K / LT_BK;
    hi.assign((size_t)ntiles * nk * ncta * LT_BK, 0);
    lo.assign(hi.size(), 0);
    for (int n = 0; n < N; n++)
        for (int k = 0; k < K; k++) {
            float v = W[(size_t)n * K + k];
            unsigned short h = f2bf_rne(v), l = f2bf_rne(v - bf2f(h));
            int nt = n / ncta, r = n % ncta, kc = k / LT_BK, c = (k % LT_BK) / 8, w = k % 8;
            size_t o = ((((size_t)nt * nk + kc) * 8 + c) * ncta + r) * 8 + w;
            hi[o] = h; lo[o] = l;
        }
}

extern "C" int mqe_policy_tc_prepare(const MqeWeights *w, int rows, PolicyTcWeights *out, cudaStream_t st) {
    const size_t n = (size_t)6 * MQE_HIST_FRAMES * TC_TILE_ELEMS;
    std::vector<unsigned short> hi(n, 0), lo(n, 0);
    for (int col = 0; col < 768; col++) {
        const float *src = col < 256 ? w->adapt_w0 + (size_t)col * 2100 : w->body_w0 + (size_t)(col - 256) * 2102;
        const int nt = col >> 7, r = col & 127;
        for (int b = 0; b < MQE_HIST_FRAMES; b++)
            for (int i = 0; i < MQE_LOC_OBS; i++) {
                float v = src[b * MQE_LOC_OBS + i];
                unsigned short h = f2bf_rne(v), l = f2bf_rne(v - bf2f(h));
                size_t o = ((((size_t)nt * MQE_HIST_FRAMES + b) * 10 + (i >> 3)) * 128 + r) * 8 + (i & 7);
                hi[o] = h; lo[o] = l;
            }
    }
    // tail layers: adapt.2 (128x256), adapt.4 (2x128 -> 16 rows), body.2 (256x512), body.4 (128x256), body.6 (12x128 -> 16 rows)
    const float *Ws[5] = {w->adapt_w1, w->adapt_w2, w->body_w1, w->body_w2, w->body_w3};
    const int Ns[5] = {128, 2, 256, 128, 12}, Ks[5] = {256, 128, 512, 256, 128}, Cs[5] = {128, 16, 128, 128, 16};
    std::vector<unsigned short> thi[5], tlo[5];
    size_t total = 2 * n;
    for (int i = 0; i < 5; i++) { tile_layer(Ws[i], Ns[i], Ks[i], Cs[i], thi[i], tlo[i]); total += 2 * thi[i].size(); }
    // activation planes between layers: adapt.0 out (256), adapt.2 out (128), body.0 out (512), body.2 out (256), body.4 out (128)
    const int mpad = (rows + 127) / 128 * 128;
    const int act_k[5] = {256, 128, 512, 256, 128};
    size_t act_off[5];
    for (int i = 0; i < 5; i++) { act_off[i] = total; total += 2 * (size_t)mpad * act_k[i]; }
    void *blob = nullptr;
    if (cudaMalloc(&blob, total * sizeof(unsigned short)) != cudaSuccess) return -1;
    if (cudaMemsetAsync(blob, 0, total * sizeof(unsigned short), st) != cudaSuccess) return -1;
    unsigned short *base = (unsigned short *)blob;
    if (cudaMemcpyAsync(base, hi.data(), n * 2, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
    if (cudaMemcpyAsync(base + n, lo.data(), n * 2, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
    size_t off = 2 * n;
    for (int i = 0; i < 5; i++) {
        const size_t m = thi[i].size();
        if (cudaMemcpyAsync(base + off, thi[i].data(), m * 2, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
        if (cudaMemcpyAsync(base + off + m, tlo[i].data(), m * 2, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
        out->t_hi[i] = base + off; out->t_lo[i] = base + off + m;
        off += 2 * m;
        out->p_hi[i] = base + act_off[i]; out->p_lo[i] = base + act_off[i] + (size_t)mpad * act_k[i];
    }
    out->rows = rows;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_policy_l0_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_linear_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, lt_smem_bytes(128)) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_linear_tc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, lt_smem_bytes(16)) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_linear_tc<128, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lt_smem_bytes(128)) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_linear_tc<128, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lt_smem_bytes(128)) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_linear_tc<128, 12, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lt_smem_bytes(128)) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_policy_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM_BYTES) != cudaSuccess) return -1;
    out->blob = blob;
    out->bytes = total * sizeof(unsigned short);
    out->l0_hi = blob;
    out->l0_lo = base + n;
    return 0;
}

// layer 0 only: Z[M][768] = act(ring x W0cat^T + b0cat), ELU on the adapt columns (< 256); the body columns stay
// pre-activation until the latent columns are added (prologue of the first tail layer).
extern "C" cudaError_t mqe_launch_policy_l0_tc(const PolicyTcWeights &w, const float *b0cat, const unsigned short *hist_hi,
                                               const unsigned short *hist_lo, int head, int rows, int passes, float *Z, int planes_out,
                                               const int *ctr, cudaStream_t st) {
    dim3 grid(6, (rows + 127) / 128);
    return launch_heavy(k_policy_l0_tc, grid, dim3(320), TC_SMEM_BYTES, st, hist_hi, hist_lo, (const unsigned short *)w.l0_hi,
                      (const unsigned short *)w.l0_lo, b0cat, Z, planes_out ? (unsigned short *)w.p_hi[0] : (unsigned short *)nullptr,
                      planes_out ? (unsigned short *)w.p_lo[0] : (unsigned short *)nullptr, rows, head, passes, ctr, 0, 0, (float *)nullptr, (const unsigned char *)nullptr, 1, 0, 0, 1);
}

// Fused policy (default): ONE layer-0 launch over all 768 columns (the six column tiles of a row tile are neighbours in launch order, so the
// history planes are fetched from HBM once and shared through L2), then ONE kernel for everything behind it (k_policy_tail).  With
// k_policy_frame in front that is three launches for preprocess_action (go1.py:64-108).
extern "C" cudaError_t mqe_launch_policy_tc_fused(const PolicyTcWeights &w, const PolicyWeightsDev &pw, const PolicyScratch &s, const DevParams &p,
                                                  const unsigned short *hist_hi, const unsigned short *hist_lo, int head, int M, int passes, const int *ctr,
                                                  int finish, cudaStream_t st, int *launches) {
    const int mt = (M + 127) / 128;
    cudaError_t e;
    if ((e = launch_heavy(k_policy_l0_tc, dim3(6, mt), dim3(320), TC_SMEM_BYTES, st, hist_hi, hist_lo, (const unsigned short *)w.l0_hi,
                          (const unsigned short *)w.l0_lo, pw.b0cat, s.Z, (unsigned short *)w.p_hi[0], (unsigned short *)w.p_lo[0], M, head, passes, ctr, 0, 0, (float *)nullptr, (const unsigned char *)nullptr, 1, 0, 1, 1)) != cudaSuccess) return e;
    if ((e = mqe_launch_policy_tail_only(w, pw, s, p, M, passes, finish, st)) != cudaSuccess) return e;
    *launches += 2;
    return cudaGetLastError();
}
extern "C" cudaError_t mqe_launch_policy_tail_only(const PolicyTcWeights &w, const PolicyWeightsDev &pw, const PolicyScratch &s, const DevParams &p, int M,
                                                   int passes, int finish, cudaStream_t st) {
    const int mt = (M + 127) / 128;
    cudaError_t e;
    TailArgs a;
    a.a0_hi = (const unsigned short *)w.p_hi[0]; a.a0_lo = (const unsigned short *)w.p_lo[0];
    a.Z = s.Z;
    a.wa_hi = (const unsigned short *)w.t_hi[0]; a.wa_lo = (const unsigned short *)w.t_lo[0];
    a.wb1_hi = (const unsigned short *)w.t_hi[2]; a.wb1_lo = (const unsigned short *)w.t_lo[2];
    a.wb2_hi = (const unsigned short *)w.t_hi[3]; a.wb2_lo = (const unsigned short *)w.t_lo[3];
    a.ba1 = pw.ab1; a.bb1 = pw.bb1; a.bb2 = pw.bb2;
    a.aw2 = pw.aw2; a.ab2 = pw.ab2; a.bw3 = pw.bw3; a.bb3 = pw.bb3; a.wlat = pw.wlat;
    a.latent = s.latent; a.act = s.act;
    a.M = M; a.passes = passes; a.finish = finish;
    // plain stream order (not PDL): a 213 KB-per-CTA grid that becomes resident early would take SMs from layer 0's last wave
    if ((e = launch_heavy(k_policy_tail, dim3(mt), dim3(FT_THREADS), FT_SMEM_BYTES, st, a, p)) != cudaSuccess) return e;
    return cudaGetLastError();
}

// Incremental layer 0.  The first layer contracts the 30-frame history; 29 of the 30 frames of step t+1 are already in the ring when step t's
// frame has been written.  mqe_launch_policy_l0_old (mode 1) contracts those 29 frames into Zold at LOW launch priority on a side stream
// behind the policy of step t, i.e. concurrently with k_substeps / k_post_physics of step t, whose one-wave grid leaves ~30 % of the SM time
// idle in its tail; step t+1 then only needs the K = 80 GEMM of its new frame (mode 2) in front of the fused tail.  preprocess_action's
// critical path drops from frame + 96 us + tail to frame + ~6 us + tail.
// row tiles [mtile0, mtile0 + mtiles) (mtiles < 0: to the end)
extern "C" cudaError_t mqe_launch_policy_l0_old(const PolicyTcWeights &w, const unsigned short *hist_hi, const unsigned short *hist_lo, int head_next,
                                                int M, int passes, float *Zold, const int *ctr, int mtile0, int mtiles, cudaStream_t st) {
    const int mt = (M + 127) / 128;
    if (mtiles < 0) mtiles = mt - mtile0;
    if (mtiles <= 0) return cudaSuccess;
    return launch_background(k_policy_l0_tc, dim3(6, mtiles), dim3(320), TC_SMEM_BYTES, st, hist_hi, hist_lo, (const unsigned short *)w.l0_hi,
                             (const unsigned short *)w.l0_lo, (const float *)nullptr, (float *)nullptr, (unsigned short *)nullptr, (unsigned short *)nullptr,
                             M, head_next, passes, ctr, 0, 1, Zold, (const unsigned char *)nullptr, 1, mtile0, 0, 1);
}
extern "C" cudaError_t mqe_launch_policy_tc_incremental(const PolicyTcWeights &w, const PolicyWeightsDev &pw, const PolicyScratch &s, const DevParams &p,
                                                        const unsigned short *hist_hi, const unsigned short *hist_lo, int head, int M, int passes,
                                                        const int *ctr, int finish, cudaStream_t st, int *launches,
                                                        int early_tiles, int head_next, cudaStream_t aux, cudaEvent_t ev, cudaEvent_t join_before_tail) {
    const int mt = (M + 127) / 128;
    cudaError_t e;
    static const int ntpc = [] { const char *e = getenv("MQE_L0_NTPC"); const int v = e ? atoi(e) : 3; return (v == 1 || v == 2 || v == 3) ? v : 3; }();     // 512 TMEM columns = at most 3 accumulators of 128
    if ((e = launch_heavy(k_policy_l0_tc, dim3(6 / ntpc, mt), dim3(320), TC_SMEM_BYTES, st, hist_hi, hist_lo, (const unsigned short *)w.l0_hi,
                          (const unsigned short *)w.l0_lo, pw.b0cat, s.Z, (unsigned short *)w.p_hi[0], (unsigned short *)w.p_lo[0], M, head, passes, ctr, 0,
                          2, s.Zold, (const unsigned char *)p.hist_dirty, p.A, 0, 1, ntpc)) != cudaSuccess) return e;
    *launches += 2;
    if (early_tiles > 0 && aux) {
        // the first row tiles of the NEXT step's 29-frame pass start right here, beside the fused tail, which keeps only 64 of the 148
        // SMs busy: as many CTAs as finish before the tail does (so k_substeps still finds every SM free).  It may only start once the
        // new-frame pass above has read Zold.  NOTE: the tail advances ctr[0] at its end, so this launch takes the next slot explicitly
        // (head_next >= 0) or, in graph replay, derives it from the not-yet-advanced counter (head_next == -2).
        if ((e = cudaEventRecord(ev, st)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(aux, ev, 0)) != cudaSuccess) return e;
        if ((e = mqe_launch_policy_l0_old(w, hist_hi, hist_lo, head_next, M, passes, s.Zold, ctr, 0, early_tiles < mt ? early_tiles : mt, aux)) != cudaSuccess) return e;
        *launches += 1;
    }
    // a side-stream kernel of this step (k_balance_tasks) is joined HERE, in front of the tail, not in front of k_substeps: an extra edge
    // into the k_substeps node makes the instantiated graph launch the background layer-0 pass first, which then takes the SMs
    if (join_before_tail && (e = cudaStreamWaitEvent(st, join_before_tail, 0)) != cudaSuccess) return e;
    if ((e = mqe_launch_policy_tail_only(w, pw, s, p, M, passes, finish, st)) != cudaSuccess) return e;
    return cudaGetLastError();
}

// Forked policy: the adaptation branch (layer-0 columns 0..255 -> adapt.2 -> adapt.4 = latent) runs on a second stream next to the body's
// layer-0 columns (256..767) and joins before body.0's latent / ELU / plane stage.  Same kernels, same arithmetic; the 64-CTA adapt
// layer then fills SMs that the 2.6-wave layer-0 grid leaves idle in its last wave instead of running after it.
extern "C" cudaError_t mqe_launch_policy_tc_forked(const PolicyTcWeights &w, const PolicyWeightsDev &pw, const PolicyScratch &s, const unsigned short *hist_hi,
                                                   const unsigned short *hist_lo, int head, int M, int passes, const int *ctr, cudaStream_t st, cudaStream_t aux,
                                                   cudaEvent_t ev_fork, cudaEvent_t ev_join, int *launches) {
    const int mt = (M + 127) / 128, mpad = mt * 128;
    auto H = [&](int i) { return (const unsigned short *)w.t_hi[i]; };
    auto L = [&](int i) { return (const unsigned short *)w.t_lo[i]; };
    auto PH = [&](int i) { return (unsigned short *)w.p_hi[i]; };
    auto PL = [&](int i) { return (unsigned short *)w.p_lo[i]; };
    float *const nof = nullptr;
    unsigned short *const nou = nullptr;
    const HeadArgs none = {};
    const HeadArgs h1 = {pw.aw2, pw.ab2, s.latent, s.Z, pw.wlat, PH(2), PL(2)};
    const HeadArgs h2 = {pw.bw3, pw.bb3, s.act, nullptr, nullptr, nullptr, nullptr};
    cudaError_t e;
    if ((e = cudaEventRecord(ev_fork, st)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(aux, ev_fork, 0)) != cudaSuccess) return e;
    if ((e = launch_pdl_if(false, k_policy_l0_tc, dim3(2, mt), dim3(320), TC_SMEM_BYTES, aux, hist_hi, hist_lo, (const unsigned short *)w.l0_hi,
                           (const unsigned short *)w.l0_lo, pw.b0cat, s.Z, PH(0), PL(0), M, head, passes, ctr, 0, 0, (float *)nullptr, (const unsigned char *)nullptr, 1, 0, 0, 1)) != cudaSuccess) return e;
    if ((e = launch_pdl(k_linear_tc<128, 2, false>, dim3(1, mt), dim3(LT_THREADS), lt_smem_bytes(128), aux, PH(0), PL(0), 256, H(0), L(0), pw.ab1, nof, 0, 128, nou, nou, 0, M, 1, passes, h1)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(ev_join, aux)) != cudaSuccess) return e;
    if ((e = launch_pdl_if(false, k_policy_l0_tc, dim3(4, mt), dim3(320), TC_SMEM_BYTES, st, hist_hi, hist_lo, (const unsigned short *)w.l0_hi,
                           (const unsigned short *)w.l0_lo, pw.b0cat, s.Z, PH(0), PL(0), M, head, passes, ctr, 2, 0, (float *)nullptr, (const unsigned char *)nullptr, 1, 0, 0, 1)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(st, ev_join, 0)) != cudaSuccess) return e;
    if ((e = launch_pdl_if(false, k_body_latent_planes, dim3((mpad * 64 + 255) / 256), dim3(256), 0, st, s.Z, s.latent, pw.wlat, PH(2), PL(2), M, mpad)) != cudaSuccess) return e;
    if ((e = launch_pdl(k_linear_tc<128>, dim3(2, mt), dim3(LT_THREADS), lt_smem_bytes(128), st, PH(2), PL(2), 512, H(2), L(2), pw.bb1, nof, 0, 256, PH(3), PL(3), 32, M, 1, passes, none)) != cudaSuccess) return e;
    if ((e = launch_pdl(k_linear_tc<128, 12, false>, dim3(1, mt), dim3(LT_THREADS), lt_smem_bytes(128), st, PH(3), PL(3), 256, H(3), L(3), pw.bb2, nof, 0, 128, nou, nou, 0, M, 1, passes, h2)) != cudaSuccess) return e;
    *launches += 6;
    return cudaGetLastError();
}

// layers 1.. on the tensor cores: operands are bf16 hi/lo planes end to end (needs layer 0 launched with planes_out = 1)
extern "C" cudaError_t mqe_launch_policy_tail_tc(const PolicyTcWeights &w, const PolicyWeightsDev &pw, const PolicyScratch &s, int M, int passes,
                                                 cudaStream_t st, int *launches) {
    const int mt = (M + 127) / 128, mpad = mt * 128;
    auto H = [&](int i) { return (const unsigned short *)w.t_hi[i]; };
    auto L = [&](int i) { return (const unsigned short *)w.t_lo[i]; };
    auto PH = [&](int i) { return (unsigned short *)w.p_hi[i]; };
    auto PL = [&](int i) { return (unsigned short *)w.p_lo[i]; };
    float *const nof = nullptr;
    unsigned short *const nou = nullptr;
    const HeadArgs none = {};
    cudaError_t e;
    // MQE_TC_FUSED_HEADS: 1 (default) the two small heads ride in the epilogue of the layer before them (4 launches);
    // 2 additionally folds body.0's latent / ELU / plane stage into the adapt kernel (3 launches; measured slower: that stage
    // then runs on the 64 CTAs of the row tiles instead of the whole chip); 0 one launch per layer (6).
    static const int fused = [] { const char *v = getenv("MQE_TC_FUSED_HEADS"); return v ? atoi(v) : 1; }();
    if (fused >= 1) {
        const HeadArgs h1 = {pw.aw2, pw.ab2, s.latent, s.Z, pw.wlat, PH(2), PL(2)};
        const HeadArgs h2 = {pw.bw3, pw.bb3, s.act, nullptr, nullptr, nullptr, nullptr};
        int n = 3;
        if (fused >= 2) {
            if ((e = launch_pdl(k_linear_tc<128, 2, true>, dim3(1, mt), dim3(LT_THREADS), lt_smem_bytes(128), st, PH(0), PL(0), 256, H(0), L(0), pw.ab1, nof, 0, 128, nou, nou, 0, M, 1, passes, h1)) != cudaSuccess) return e;
        } else {
            if ((e = launch_pdl(k_linear_tc<128, 2, false>, dim3(1, mt), dim3(LT_THREADS), lt_smem_bytes(128), st, PH(0), PL(0), 256, H(0), L(0), pw.ab1, nof, 0, 128, nou, nou, 0, M, 1, passes, h1)) != cudaSuccess) return e;
            if ((e = launch_pdl(k_body_latent_planes, dim3((mpad * 64 + 255) / 256), dim3(256), 0, st, s.Z, s.latent, pw.wlat, PH(2), PL(2), M, mpad)) != cudaSuccess) return e;
            n = 4;
        }
        if ((e = launch_pdl(k_linear_tc<128>, dim3(2, mt), dim3(LT_THREADS), lt_smem_bytes(128), st, PH(2), PL(2), 512, H(2), L(2), pw.bb1, nof, 0, 256, PH(3), PL(3), 32, M, 1, passes, none)) != cudaSuccess) return e;
        if ((e = launch_pdl(k_linear_tc<128, 12, false>, dim3(1, mt), dim3(LT_THREADS), lt_smem_bytes(128), st, PH(3), PL(3), 256, H(3), L(3), pw.bb2, nof, 0, 128, nou, nou, 0, M, 1, passes, h2)) != cudaSuccess) return e;
        *launches += n;
        return cudaGetLastError();
    }
    if ((e = launch_pdl(k_linear_tc<128>, dim3(1, mt), dim3(LT_THREADS), lt_smem_bytes(128), st, PH(0), PL(0), 256, H(0), L(0), pw.ab1, nof, 0, 128, PH(1), PL(1), 16, M, 1, passes, none)) != cudaSuccess) return e;
    if ((e = launch_pdl(k_linear_tc<16>, dim3(1, mt), dim3(LT_THREADS), lt_smem_bytes(16), st, PH(1), PL(1), 128, H(1), L(1), pw.ab2, s.latent, 2, 2, nou, nou, 0, M, 0, passes, none)) != cudaSuccess) return e;
    if ((e = launch_pdl(k_body_latent_planes, dim3((mpad * 64 + 255) / 256), dim3(256), 0, st, s.Z, s.latent, pw.wlat, PH(2), PL(2), M, mpad)) != cudaSuccess) return e;
    if ((e = launch_pdl(k_linear_tc<128>, dim3(2, mt), dim3(LT_THREADS), lt_smem_bytes(128), st, PH(2), PL(2), 512, H(2), L(2), pw.bb1, nof, 0, 256, PH(3), PL(3), 32, M, 1, passes, none)) != cudaSuccess) return e;
    if ((e = launch_pdl(k_linear_tc<128>, dim3(1, mt), dim3(LT_THREADS), lt_smem_bytes(128), st, PH(3), PL(3), 256, H(3), L(3), pw.bb2, nof, 0, 128, PH(4), PL(4), 16, M, 1, passes, none)) != cudaSuccess) return e;
    if ((e = launch_pdl(k_linear_tc<16>, dim3(1, mt), dim3(LT_THREADS), lt_smem_bytes(16), st, PH(4), PL(4), 128, H(4), L(4), pw.bb3, s.act, 12, 12, nou, nou, 0, M, 0, passes, none)) != cudaSuccess) return e;
    *launches += 6;
    return cudaGetLastError();
}
