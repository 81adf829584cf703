// physics.cu -- the fused decimation kernel: `decimation` x { actuator net -> articulated-body forward dynamics ->
// contact generation -> projected Gauss-Seidel impulse solve -> integration } for every environment, state held in
// registers / shared memory for the whole policy step.
//
// Replaces (reference) go1.py:48-58: _compute_torques (go1.py:315-354), gym.set_dof_actuation_force_tensor,
// gym.simulate, gym.fetch_results, gym.refresh_dof_state_tensor, and the net-contact-force / root-state refresh
// of legged_robot.py:122-124.  Algorithm: DESIGN.md section 4.  The CPU oracle (oracle/mqe_oracle.c) states the same
// physics with a dense mass matrix; here the mass matrix is never formed:
//
//   * one lane per (robot, leg); the four lanes of a robot reduce base-level quantities with xor-shuffles;
//   * generalized velocity is kept in block-LDL^T coordinates  w = (v_base, u_l = qd_l + G_l^T v_base)  in which
//     M^-1 = blockdiag(S^-1, H_l^-1) (S = Schur complement of the base), so a contact row is 9+9 floats and one
//     Gauss-Seidel update costs 18 FMAs;
//   * all per-robot operators of one substep live in shared memory (4 KB / robot), the persistent state in registers.
#include "post_dev.cuh"
#include "kernels.cuh"

#define ROWF 24        // floats per local row: Jb6 Jl3 Yb6 Yl3 dinv bias lambda meta + pad
#define PROWF 44       // floats per pair row: side A at 0 (18 + 2 pad), side B at 20 (18 + 2 pad), then dinv bias lambda meta at 40
// per-robot shared block (floats)
#define RS_ORIGIN 0
#define RS_SINV 3
#define RS_G 24
#define RS_HINV 96
#define RS_A 120
#define RS_P 156
#define RS_CAP 192
#define RS_CNT 318     // int scratch
#define RS_BOUND 319   // true bounding radius of the robot's capsules about its origin (this substep)
#define RS_ROWS 320
#define SROWS 19       // local rows of a robot kept in shared memory; rows past that (a robot lying on the ground) go to a
                       // per-robot global scratch with the same arithmetic -- shared memory per env sets the residency
static_assert((SROWS - MQE_MAX_LIMIT) % 3 == 0, "contact blocks start at row MQE_MAX_LIMIT and must not straddle the shared-memory rows");
#define RS_CMETA (RS_ROWS + SROWS * ROWF)
#define RS_FORCE (RS_CMETA + MQE_MAX_LOCAL * 4)
#define RS_SIZE (RS_FORCE + 52)
// per-npc shared block
#define NS_ORIGIN 0
#define NS_CAP 3       // p0 p1 r   (seesaw: cos, sin of the plank angle)
#define NS_CNT 10
#define NS_BOUND 11
#define NS_FORCE 12
#define NS_ROT 16      // box: rotation matrix, columns ex ey ez (9 floats)
#define NS_ROWS 28
#define NS_CMETA (NS_ROWS + 12 * ROWF)
#define NS_SIZE (NS_CMETA + 16)
// per-warp pair-row pool: E * spair contacts x 3 rows.  Envs claim slots in env order each substep (most substeps only one env
// of a warp has dynamic contacts, and it then keeps all of them in shared memory); what does not fit goes to a global scratch.
// per-env: int capmask[G][G] (capsules of X within reach of group Y)
#define ES_MASKSZ(G) (((G) * (G) + 3) & ~3)
#define PDESCF 12      // pair-contact descriptor (global scratch): n3, body a, body b, point3, gap, packed (X, ci, Y, cj)

#define ACTW_PLAIN 1316    // W0[32][6] b0[32] W1T[32][32] b1[32] W2[32] b2 = 1313 floats, padded (FFMA path, k_actuator)
#define ACTW_FLOATS 1380   // actuator-net region of the CTA header: the plain table, or (MMA path) the fp16 hi / lo fragment table:
                           // uint4 frag[10][32 lanes] (W0 hi, W0 lo, then {hi.b0, hi.b1, lo.b0, lo.b1} per (n-tile, k-step) of W1),
                           // then b0[32] b1[32] W2[32] b2 as floats at 1280
#define TBL_INTS 80        // per-leg probe lists [4][10] (count + 9 ids), per-leg capsule lists [4][10]

__host__ __device__ inline int physics_warp_smem_floats(int A, int P, int E, int spair, int maxpair) {
    (void)maxpair;
    return E * A * RS_SIZE + E * P * NS_SIZE + E * spair * 3 * PROWF + E * ES_MASKSZ(A + P);
}
__host__ __device__ inline int physics_cta_header_floats() { return (int)(sizeof(MqeRobotModel) / 4) + ACTW_FLOATS + TBL_INTS; }

struct SV { V3 w, v; };
struct RBI { float m; V3 h; float I[6]; };   // xx xy xz yy yz zz about O

__device__ __forceinline__ RBI rbi_from_link(const float *in, const M3 &R, V3 p, float added_mass = 0.f, V3 com_shift = V3{0.f, 0.f, 0.f}) {
    RBI o;
    float m = in[0] + added_mass;      // extra mass sits at the link's COM (PhysX changes the mass, not the inertia tensor)
    V3 c = mul(R, mk(in[1] + com_shift.x, in[2] + com_shift.y, in[3] + com_shift.z)) + p;      // randomize_com moves the COM, the tensor about it stays
    V3 t0 = in[4] * R.c0 + in[5] * R.c1 + in[6] * R.c2;
    V3 t1 = in[5] * R.c0 + in[7] * R.c1 + in[8] * R.c2;
    V3 t2 = in[6] * R.c0 + in[8] * R.c1 + in[9] * R.c2;
    float cc = dot(c, c);
    o.m = m;
    o.h = m * c;
    o.I[0] = t0.x * R.c0.x + t1.x * R.c1.x + t2.x * R.c2.x + m * (cc - c.x * c.x);
    o.I[1] = t0.x * R.c0.y + t1.x * R.c1.y + t2.x * R.c2.y - m * c.x * c.y;
    o.I[2] = t0.x * R.c0.z + t1.x * R.c1.z + t2.x * R.c2.z - m * c.x * c.z;
    o.I[3] = t0.y * R.c0.y + t1.y * R.c1.y + t2.y * R.c2.y + m * (cc - c.y * c.y);
    o.I[4] = t0.y * R.c0.z + t1.y * R.c1.z + t2.y * R.c2.z - m * c.y * c.z;
    o.I[5] = t0.z * R.c0.z + t1.z * R.c1.z + t2.z * R.c2.z + m * (cc - c.z * c.z);
    return o;
}
__device__ __forceinline__ void rbi_add(RBI &o, const RBI &a) {
    o.m += a.m; o.h = o.h + a.h;
#pragma unroll
    for (int i = 0; i < 6; i++) o.I[i] += a.I[i];
}
__device__ __forceinline__ SV rbi_mul(const RBI &I, const SV &x) {
    SV f;
    V3 hv = cross(I.h, x.v), hw = cross(I.h, x.w);
    f.w = mk(I.I[0] * x.w.x + I.I[1] * x.w.y + I.I[2] * x.w.z + hv.x,
             I.I[1] * x.w.x + I.I[3] * x.w.y + I.I[4] * x.w.z + hv.y,
             I.I[2] * x.w.x + I.I[4] * x.w.y + I.I[5] * x.w.z + hv.z);
    f.v = I.m * x.v - hw;
    return f;
}
__device__ __forceinline__ SV crm(const SV &a, const SV &b) { SV o; o.w = cross(a.w, b.w); o.v = cross(a.w, b.v) + cross(a.v, b.w); return o; }
__device__ __forceinline__ SV crf(const SV &a, const SV &f) { SV o; o.w = cross(a.w, f.w) + cross(a.v, f.v); o.v = cross(a.w, f.v); return o; }
__device__ __forceinline__ float svdot(const SV &a, const SV &b) { return dot(a.w, b.w) + dot(a.v, b.v); }
__device__ __forceinline__ SV svadd(const SV &a, const SV &b) { SV o; o.w = a.w + b.w; o.v = a.v + b.v; return o; }
__device__ __forceinline__ SV svscale(float s, const SV &a) { SV o; o.w = s * a.w; o.v = s * a.v; return o; }
__device__ __forceinline__ float svc(const SV &a, int i) { return i < 3 ? comp(a.w, i) : comp(a.v, i - 3); }

__device__ __forceinline__ float quad_sum(float x, unsigned m) {   // sum over the 4 lanes of a robot
    x += __shfl_xor_sync(m, x, 1);
    x += __shfl_xor_sync(m, x, 2);
    return x;
}
__host__ __device__ constexpr int sidx(int i, int j) { return i <= j ? (i * 6 - i * (i - 1) / 2 + (j - i)) : (j * 6 - j * (j - 1) / 2 + (i - j)); }
__host__ __device__ constexpr int lidx(int i, int j) { return i * (i + 1) / 2 + j; }

// symmetric positive definite 6x6 inverse (Cholesky), all in registers
__device__ __forceinline__ void spd6_inverse(const float *A /*21 upper*/, float *Ainv /*21*/) {
    float L[21], rinv[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int j = 0; j <= i; j++) {
            float s = A[sidx(j, i)];
#pragma unroll
            for (int k = 0; k < j; k++) s -= L[lidx(i, k)] * L[lidx(j, k)];
            if (i == j) { float d = sqrtf(fmaxf(s, 1e-20f)); L[lidx(i, i)] = d; rinv[i] = 1.f / d; }
            else L[lidx(i, j)] = s * rinv[j];
        }
    }
    float Li[21];   // inverse of L (lower)
#pragma unroll
    for (int c = 0; c < 6; c++) {
#pragma unroll
        for (int i = c; i < 6; i++) {
            float s = (i == c) ? 1.f : 0.f;
#pragma unroll
            for (int k = c; k < i; k++) s -= L[lidx(i, k)] * Li[lidx(k, c)];
            Li[lidx(i, c)] = s * rinv[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = i; j < 6; j++) {
            float s = 0.f;
#pragma unroll
            for (int k = j; k < 6; k++) s += Li[lidx(k, i)] * Li[lidx(k, j)];
            Ainv[sidx(i, j)] = s;
        }
}
__device__ __forceinline__ void sym6_mulv(const float *A, const float *x, float *y) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 6; j++) s += A[sidx(i, j)] * x[j];
        y[i] = s;
    }
}
__device__ __forceinline__ int nth_set_bit(unsigned m, int n) {      // position of the n-th (0-based) set bit
    for (int i = 0; i < n; i++) m &= m - 1u;
    return __ffs(m) - 1;
}
// value selects (never references: a reference to one of several register objects sends all of them to local memory)
__device__ __forceinline__ float sel1(int k, float a, float b, float c, float d) { return k == 0 ? a : (k == 1 ? b : (k == 2 ? c : d)); }
__device__ __forceinline__ V3 sel3(int k, V3 a, V3 b, V3 c, V3 d) { return mk(sel1(k, a.x, b.x, c.x, d.x), sel1(k, a.y, b.y, c.y, d.y), sel1(k, a.z, b.z, c.z, d.z)); }
__device__ __forceinline__ M3 selm(int k, const M3 &a, const M3 &b, const M3 &c, const M3 &d) {
    M3 o;
    o.c0 = sel3(k, a.c0, b.c0, c.c0, d.c0); o.c1 = sel3(k, a.c1, b.c1, c.c1, d.c1); o.c2 = sel3(k, a.c2, b.c2, c.c2, d.c2);
    return o;
}
// sym 3x3: 00 01 02 11 12 22
__device__ __forceinline__ void sym3_inverse(const float *H, float *Hi) {
    float c00 = H[3] * H[5] - H[4] * H[4], c01 = H[2] * H[4] - H[1] * H[5], c02 = H[1] * H[4] - H[2] * H[3];
    float det = H[0] * c00 + H[1] * c01 + H[2] * c02;
    float r = 1.f / det;
    Hi[0] = c00 * r; Hi[1] = c01 * r; Hi[2] = c02 * r;
    Hi[3] = (H[0] * H[5] - H[2] * H[2]) * r;
    Hi[4] = (H[1] * H[2] - H[0] * H[4]) * r;
    Hi[5] = (H[0] * H[3] - H[1] * H[1]) * r;
}
__device__ __forceinline__ void sym3_mulv(const float *H, const float *x, float *y) {
    y[0] = H[0] * x[0] + H[1] * x[1] + H[2] * x[2];
    y[1] = H[1] * x[0] + H[3] * x[1] + H[4] * x[2];
    y[2] = H[2] * x[0] + H[4] * x[1] + H[5] * x[2];
}

// packed fp32 pairs for FFMA2 (fma.rn.f32x2, sm_100+): two independent round-to-nearest fp32 FMAs per instruction
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(unsigned long long &d, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ void tangent_basis(V3 n, V3 &t1, V3 &t2) {
    V3 e = fabsf(n.x) < 0.9f ? mk(1.f, 0.f, 0.f) : mk(0.f, 1.f, 0.f);
    t1 = cross(n, e);
    t1 = (1.f / sqrtf(dot(t1, t1))) * t1;
    t2 = cross(n, t1);
}
__device__ __forceinline__ float contact_bias(const DevParams &p, float gap) {
    if (gap > 0.f) return gap / p.dt;
    return fmaxf(p.erp * gap / p.dt, -p.vdep);
}
// x / (1 + |x|): reciprocal by MUFU.RCP + one Newton step (d >= 1, so no special cases); <= 1 ulp from the exact quotient
__device__ __forceinline__ float softsign(float x) {
    float d = 1.f + fabsf(x), r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    r = fmaf(r, fmaf(-d, r, 1.f), r);
    return x * r;
}

// world probe of a sphere against floor slab + wall footprint; bit0: floor/top contact, bit1: wall contact
struct ProbeHit { int mask; float gap_f, gap_w; V3 nw; };
// `fix`: position of the seesaw's fixed base when the env has one (platform top is ground inside its footprint, the column
// below it is a vertical cylinder that competes with the wall footprint for the wall-like contact); seesaw.urdf.
// `far`: the caller has proved sdf(x) - r >= contact offset (wall_is_far below), so the four SDF loads are skipped; every
// output that matters is the same as with the real sample (no wall bit, not inside the footprint).
__device__ __forceinline__ ProbeHit probe_world(const DevParams &p, V3 x, float r, bool has_fix = false, V3 fix = V3{0.f, 0.f, 0.f}, bool far = false) {
    ProbeHit h;
    h.mask = 0;
    SdfSample s;
    if (far) { s.sdf = 1.0e3f; s.gx = 0.f; s.gy = 0.f; }
    else s = sdf_sample(p, x.x, x.y);
    bool inside = s.sdf < 0.f, above = x.z >= p.wall_top;
    float ground = (inside && above) ? p.wall_top : p.floor_z;
    float gw = s.sdf - r, gn = sqrtf(s.gx * s.gx + s.gy * s.gy);
    bool wall_ok = !above && gn > 1e-6f;
    h.nw = wall_ok ? mk(s.gx / gn, s.gy / gn, 0.f) : mk(0.f, 0.f, 0.f);
    if (has_fix && p.npc_kind == MQE_NPC_PLATFORM) {      // wrestling.urdf / bridge.urdf: tops of fixed boxes are ground inside their footprints
        const float *g = p.geom;
        for (int b = 0; b < (int)g[0]; b++) {
            const float *bx = g + 1 + 5 * b;
            const float px = x.x - fix.x - bx[0], py = x.y - fix.y - bx[1], top = fix.z + bx[4];
            if (fabsf(px) <= bx[2] && fabsf(py) <= bx[3] && x.z >= top - 0.15f && top > ground) ground = top;
        }
    } else if (has_fix) {
        const float *g = p.geom;
        float px = x.x - fix.x, py = x.y - fix.y;
        if (g[7] > 0.f && fabsf(px) <= g[7] && fabsf(py) <= g[8] && x.z >= fix.z) ground = fmaxf(ground, fix.z + g[9]);
        if (g[10] > 0.f && x.z < fix.z && x.z > fix.z - g[11] - r) {
            float dh = sqrtf(px * px + py * py), gc = dh - g[10] - r;
            if (dh > 1e-6f && (!wall_ok || gc < gw)) { wall_ok = true; gw = gc; h.nw = mk(px / dh, py / dh, 0.f); }
        }
    }
    h.gap_f = x.z - r - ground;
    if (h.gap_f < p.coff) h.mask |= 1;
    h.gap_w = gw;
    if (wall_ok && gw < p.coff) h.mask |= 2;
    return h;
}

// Wall cull.  The SDF grid is a Euclidean distance transform (1-Lipschitz); its bilinear interpolant has |grad| <= sqrt(2).
// With sdf_o sampled at the robot origin, a probe whose centre is xr (relative to the origin) and radius r cannot reach a wall,
// nor be inside the footprint, if sdf_o - 1.5 |xr_xy| - r - 2 cells > contact offset.
__device__ __forceinline__ bool wall_is_far(const DevParams &p, float sdf_o, V3 xr, float r) {
    return sdf_o - 1.5f * sqrtf(xr.x * xr.x + xr.y * xr.y) - r - 2.f * p.sdf_cell > p.coff;
}

// sphere (centre x, radius r) vs an oriented box (centre c, axes ex ey ez, half extents h); normal box -> sphere.
// Used for the seesaw plank and the push box; same arithmetic as the oracle's sphere_obb.
__device__ __forceinline__ bool sphere_obb(const DevParams &p, V3 c, V3 ex, V3 ey, V3 ez, const float *h, V3 x, float r, V3 &nrm, float &gap, V3 &pos) {
    V3 dx = x - c;
    float loc[3] = {dot(dx, ex), dot(dx, ey), dot(dx, ez)}, df[3], nl[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; i++) df[i] = loc[i] - fminf(fmaxf(loc[i], -h[i]), h[i]);
    float d2 = df[0] * df[0] + df[1] * df[1] + df[2] * df[2];
    if (d2 > 1e-12f) {
        float dist = sqrtf(d2);
        nl[0] = df[0] / dist; nl[1] = df[1] / dist; nl[2] = df[2] / dist;
        gap = dist - r;
    } else {
        int ax = 0;
        float best = h[0] - fabsf(loc[0]);
#pragma unroll
        for (int i = 1; i < 3; i++) { float pen = h[i] - fabsf(loc[i]); if (pen < best) { best = pen; ax = i; } }
        float sg = (ax == 0 ? loc[0] : (ax == 1 ? loc[1] : loc[2])) >= 0.f ? 1.f : -1.f;
        nl[0] = ax == 0 ? sg : 0.f; nl[1] = ax == 1 ? sg : 0.f; nl[2] = ax == 2 ? sg : 0.f;
        gap = -best - r;
    }
    if (gap >= p.coff) return false;
    nrm = nl[0] * ex + nl[1] * ey + nl[2] * ez;
    pos = x - (r + 0.5f * gap) * nrm;
    return true;
}

// sphere (centre x, radius r) vs a solid vertical cylinder (centre c, radius R, half height hh); normal cylinder -> sphere.
// The tug-of-war disc (resources/objects/cylinder.urdf); same arithmetic as the oracle's sphere_vcyl.
__device__ __forceinline__ bool sphere_vcyl(const DevParams &p, V3 c, float R, float hh, V3 x, float r, V3 &nrm, float &gap, V3 &pos) {
    const V3 dx = x - c;
    const float dr = sqrtf(dx.x * dx.x + dx.y * dx.y);
    const float ux = dr > 1e-9f ? dx.x / dr : 1.f, uy = dr > 1e-9f ? dx.y / dr : 0.f;
    const float dfr = dr - fminf(dr, R), dfz = dx.z - fminf(fmaxf(dx.z, -hh), hh);
    const float d2 = dfr * dfr + dfz * dfz;
    if (d2 > 1e-12f) {
        const float dist = sqrtf(d2);
        nrm = mk(ux * dfr / dist, uy * dfr / dist, dfz / dist);
        gap = dist - r;
    } else {
        const float penr = R - dr, penz = hh - fabsf(dx.z);
        if (penr < penz) { nrm = mk(ux, uy, 0.f); gap = -penr - r; }
        else { nrm = mk(0.f, 0.f, dx.z >= 0.f ? 1.f : -1.f); gap = -penz - r; }
    }
    if (gap >= p.coff) return false;
    pos = x - (r + 0.5f * gap) * nrm;
    return true;
}

// closest points of two segments (Ericson 5.1.9); same branch structure as the oracle
__device__ __forceinline__ void seg_seg(V3 p1, V3 q1, V3 p2, V3 q2, V3 &c1, V3 &c2) {
    V3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
    float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), s, t;
    const float EPS = 1e-12f;
    if (a <= EPS && e <= EPS) { s = t = 0.f; }
    else if (a <= EPS) { s = 0.f; t = fminf(fmaxf(f / e, 0.f), 1.f); }
    else {
        float c = dot(d1, r);
        if (e <= EPS) { t = 0.f; s = fminf(fmaxf(-c / a, 0.f), 1.f); }
        else {
            float b = dot(d1, d2), denom = a * e - b * b;
            s = denom > EPS ? fminf(fmaxf((b * f - c * e) / denom, 0.f), 1.f) : 0.f;
            t = (b * s + f) / e;
            if (t < 0.f) { t = 0.f; s = fminf(fmaxf(-c / a, 0.f), 1.f); }
            else if (t > 1.f) { t = 1.f; s = fminf(fmaxf((b - c) / a, 0.f), 1.f); }
        }
    }
    c1 = p1 + s * d1;
    c2 = p2 + t * d2;
}

// Build one side (18 floats: Jb6 Jl3 Yb6 Yl3) of a row for a robot from the operators in shared memory.
// link: 0 base, 1..12.  Returns the diagonal contribution J.Y.
__device__ __forceinline__ float robot_side_from_smem(const float *rs, int link, V3 r, V3 d, float *out) {
    V3 rxd = cross(r, d);
    float Jb[6] = {rxd.x, rxd.y, rxd.z, d.x, d.y, d.z}, Jl[3], Yb[6], Yl[3];
    const int leg = link > 0 ? (link - 1) / 3 : 0, k = link > 0 ? (link - 1) % 3 + 1 : 0;
#pragma unroll
    for (int j = 0; j < 3; j++) {      // fixed trip count, joints past the link masked: no dynamically indexed registers
        const float *a = rs + RS_A + (3 * leg + j) * 3, *pj = rs + RS_P + (3 * leg + j) * 3;
        V3 rel = mk(r.x - pj[0], r.y - pj[1], r.z - pj[2]);
        const float v = dot(mk(a[0], a[1], a[2]), cross(rel, d));
        Jl[j] = j < k ? v : 0.f;
    }
    const float *G = rs + RS_G + leg * 18;
#pragma unroll
    for (int i = 0; i < 6; i++) Jb[i] -= G[i * 3] * Jl[0] + G[i * 3 + 1] * Jl[1] + G[i * 3 + 2] * Jl[2];
    sym6_mulv(rs + RS_SINV, Jb, Yb);
    sym3_mulv(rs + RS_HINV + leg * 6, Jl, Yl);
    float dd = 0.f;
#pragma unroll
    for (int i = 0; i < 6; i++) { out[i] = Jb[i]; out[9 + i] = Yb[i]; dd += Jb[i] * Yb[i]; }
#pragma unroll
    for (int i = 0; i < 3; i++) { out[6 + i] = Jl[i]; out[15 + i] = Yl[i]; dd += Jl[i] * Yl[i]; }
    return dd;
}
__device__ float npc_side(const DevParams &p, V3 r, V3 d, float *out) {
    V3 rxd = cross(r, d);
    float iI = 1.f / p.npc_inertia, im = 1.f / p.npc_mass;
    float up = p.npc_ctrl == MQE_NPC_SHEEP ? 0.f : 1.f;      // sheep stay upright (go1_sheep.py:61 zeroes their tilt)
    out[0] = rxd.x; out[1] = rxd.y; out[2] = rxd.z; out[3] = d.x; out[4] = d.y; out[5] = d.z;
    out[6] = out[7] = out[8] = 0.f;
    out[9] = rxd.x * iI * up; out[10] = rxd.y * iI * up; out[11] = rxd.z * iI;
    out[12] = d.x * im; out[13] = d.y * im; out[14] = d.z * im;
    out[15] = out[16] = out[17] = 0.f;
    if (p.npc_kind == MQE_NPC_SEESAW) {      // one revolute DOF about the pivot (seesaw: y, revolving door: z): only w_axis responds
        const bool hz = p.geom[13] > 0.5f;
        const float r2 = hz ? p.geom[3] * p.geom[3] + p.geom[14] * p.geom[14] : p.geom[3] * p.geom[3] + p.geom[15] * p.geom[15];
        const float iIp = 1.f / (p.npc_inertia + p.npc_mass * r2);
#pragma unroll
        for (int i = 9; i < 15; i++) out[i] = 0.f;
        if (p.geom[13] > 1.5f) out[13] = d.y * im;      // prismatic y (tug disc): only v_y responds
        else if (hz) out[11] = rxd.z * iIp;
        else out[10] = rxd.y * iIp;
    }
    float dd = 0.f;
    for (int i = 0; i < 6; i++) dd += out[i] * out[9 + i];
    return dd;
}
// one side of a row for a robot from the lane's OWN operators (registers); k = link depth in the lane's leg (0 base .. 3 calf)
__device__ __forceinline__ float robot_side_regs(int k, V3 r, V3 d, V3 a1, V3 a2, V3 p1, V3 p2, V3 p3, const float *Gm, const float *Sinv,
                                                 const float *Hinv, float *out) {
    V3 rxd = cross(r, d);
    float Jb[6] = {rxd.x, rxd.y, rxd.z, d.x, d.y, d.z}, Jl[3] = {0.f, 0.f, 0.f}, Yb[6], Yl[3];
    if (k >= 1) Jl[0] = dot(a1, cross(r - p1, d));
    if (k >= 2) Jl[1] = dot(a2, cross(r - p2, d));
    if (k >= 3) Jl[2] = dot(a2, cross(r - p3, d));
#pragma unroll
    for (int i = 0; i < 6; i++) Jb[i] -= Gm[i * 3] * Jl[0] + Gm[i * 3 + 1] * Jl[1] + Gm[i * 3 + 2] * Jl[2];
    sym6_mulv(Sinv, Jb, Yb);
    sym3_mulv(Hinv, Jl, Yl);
    float dd = Jl[0] * Yl[0] + Jl[1] * Yl[1] + Jl[2] * Yl[2];
#pragma unroll
    for (int i = 0; i < 6; i++) { dd += Jb[i] * Yb[i]; out[i] = Jb[i]; out[9 + i] = Yb[i]; }
#pragma unroll
    for (int i = 0; i < 3; i++) { out[6 + i] = Jl[i]; out[15 + i] = Yl[i]; }
    return dd;
}

// ---- actuator network on the tensor cores (mma.sync, fp16 hi / lo split, fp32 accumulate) -------------------------------------------
// 12 joints x E x A robots = up to 96 rows per warp and substep through 6 -> 32 -> 32 -> 1 (unitree_go1.pt, go1.py:369-380): a batched MLP
// that cost 3.6 k warp instructions per substep as FFMA2 chains.  Every operand is split into fp16 hi + lo and three MMAs
// (hi*hi + hi*lo + lo*hi) accumulate in fp32 -- the dropped lo*lo term is < 2^-22 relative, i.e. fp32-level results.
__device__ __forceinline__ uint32_t pack_h2(float e0, float e1) {        // low half = e0 (lower column / k index), high half = e1
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
    return r;
}
__device__ __forceinline__ void split_h2(float e0, float e1, uint32_t &hi, uint32_t &lo) {
    hi = pack_h2(e0, e1);
    float h0, h1;
    asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi));
    lo = pack_h2(e0 - h0, e1 - h1);
}
__device__ __forceinline__ void mma_k8(float *c, uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void mma_k16(float *c, const uint32_t *a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x / (1 + |x|) with MUFU.RCP alone (<= 1 ulp of the quotient; the Newton step of softsign() buys nothing at the tolerances in use)
__device__ __forceinline__ float softsign_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + fabsf(x)));
    return x * r;
}

// Contact blocks.  The three rows of a contact (normal, two friction directions) are solved as ONE block per Gauss-Seidel sweep: the
// three J.w are formed together from the velocity at the start of the block, and the effect of the block's own earlier updates is added
// through the coupling terms K_ij = J_i . (M^-1 J_j^T) = J_i . Y_j (i > j), stored in the pad floats of rows 1 and 2.  Algebraically
// this IS the row-by-row sweep of the oracle (same order, same clamps); it shortens the dependency chain of the sweep from three
// dot -> clamp -> update rounds to one, and for pair contacts needs one exchange between the two groups instead of three.
__device__ __forceinline__ float side_dot(const float *ji, const float *yj) {      // J_i . Y_j for one side: Jb6 Jl3 at [0..8], Yb6 Yl3 at [9..17]
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 9; k++) s = fmaf(ji[k], yj[9 + k], s);
    return s;
}
__device__ __forceinline__ void local_block_coupling(float *r0, float *r1, float *r2) {
    r1[22] = side_dot(r1, r0);
    r2[22] = side_dot(r2, r0);
    r2[23] = side_dot(r2, r1);
}
__device__ __forceinline__ void pair_block_coupling(float *r0, float *r1, float *r2) {   // both sides: A at [0..17], B at [20..37]
    r1[18] = side_dot(r1, r0) + side_dot(r1 + 20, r0 + 20);
    r2[18] = side_dot(r2, r0) + side_dot(r2 + 20, r0 + 20);
    r2[19] = side_dot(r2, r1) + side_dot(r2 + 20, r1 + 20);
}

// ---- fused post-physics step (mqe_sim_step): the warp that integrated an env also does its bookkeeping, so the fast warps finish the
// step while the slowest ones are still solving contacts and no separate launch waits for the whole grid (post_dev.cuh; same stages,
// same four lanes per agent as k_post_physics).  Env-level decisions travel by shuffle: an env never spans two warps.  Kept out of line
// (DevParams is a __grid_constant__ parameter, so the reference costs nothing): inlined, its live ranges cost the substep loop registers.
struct PostLaneState { V3 pos, vlin, wang; float qx, qy, qz, qw, q[3], qd[3], act[3]; };
static __device__ __noinline__ void substeps_fused_post(const DevParams &p, const MqeRobotModel *md, const float *rs, const PostLaneState &st,
                                                        int env, int ag, int leg, int e_loc, int lane, bool active, bool is_robot, unsigned quad_mask) {
    const unsigned FULL = 0xffffffffu;
    const int A = p.A, E = p.E, GA = p.G;
    const int m_idx = env * A + ag;
    const V3 pos = st.pos, vlin = st.vlin, wang = st.wang;
    const float qx = st.qx, qy = st.qy, qz = st.qz, qw = st.qw;
    const float q[3] = {st.q[0], st.q[1], st.q[2]}, qd[3] = {st.qd[0], st.qd[1], st.qd[2]}, act[3] = {st.act[0], st.act[1], st.act[2]};
    __syncwarp();                                                      // root / dof / contact rows of the env are visible to its lanes
    const unsigned step_count = (unsigned)p.ctr[1];                    // read before this warp's arrival below: the increment comes last
    const bool robot_live = active && is_robot;
    const int lead = min(31, e_loc * 4 * A);                           // first lane of my env
    // Stage 1 and stage 4 of a live robot work from the registers this lane integrated (quaternion, velocities, its leg's q / qd /
    // actions) instead of re-reading the state through global memory: the stand-alone kernel's ~20 us are one chain of dependent
    // L2 round trips per agent, which the slowest warp would otherwise pay in full.  Same expressions, same rounding as post_dev.cuh.
    float la[3] = {0.f, 0.f, 0.f}, org[3] = {0.f, 0.f, 0.f}, gp[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, gait0 = 0.f, oz = 0.f;
    long long ep_pre = -1;
    if (robot_live && ag == 0 && leg == 0) ep_pre = p.ep_len[env];     // the env's decision lane: issued with the loads below, used three stages later
    if (robot_live) {                                                  // every load the fast path needs, issued together
#pragma unroll
        for (int k = 0; k < 3; k++) { la[k] = p.last_actions[m_idx * 12 + 3 * leg + k]; org[k] = p.env_origins[env * 3 + k]; }
        oz = p.agent_origins[m_idx * 3 + 2];
        if (p.control_type == 0) {
#pragma unroll
            for (int k = 0; k < 5; k++) gp[k] = p.loc_obs[(size_t)m_idx * MQE_LOC_OBS + 7 + k];
            gait0 = p.gait[m_idx];
        }
    }
    const float q4[4] = {qx, qy, qz, qw};
    V3 lv = mk(0, 0, 0), av = mk(0, 0, 0), pg = mk(0, 0, 0);
    float rpy[3] = {0.f, 0.f, 0.f}, clk = 0.f, pushed = 0.f;
    bool push = false;
    int f = 0;
    if (robot_live) {
        const float PI = 3.14159265358979323846f;
        lv = quat_rotate_inverse(q4, vlin); av = quat_rotate_inverse(q4, wang); pg = quat_rotate_inverse(q4, mk(0.f, 0.f, -1.f));
        p.base_quat[m_idx * 4 + leg] = leg == 0 ? qx : (leg == 1 ? qy : (leg == 2 ? qz : qw));
        if (leg < 3) {
            p.base_lin_vel[m_idx * 3 + leg] = comp(lv, leg);
            p.base_ang_vel[m_idx * 3 + leg] = comp(av, leg);
            p.proj_grav[m_idx * 3 + leg] = comp(pg, leg);
        }
        if (p.control_type == 0) {                                     // dev_gait_clock, foot `leg` on lane `leg`
            const float dt_policy = p.dt * (float)p.decimation;
            float g = fmodf(gait0 + dt_policy * gp[0], 1.0f);
            if (g < 0.f) g += 1.f;
            if (leg == 0) p.gait[m_idx] = g;
            const float fi = leg == 0 ? g + gp[1] + gp[2] + gp[3] : (leg == 1 ? g + gp[2] : (leg == 2 ? g + gp[3] : g + gp[1]));
            const float dur = gp[4];
            float r = fmodf(fi, 1.0f);
            if (r < 0.f) r += 1.f;
            float x = fi;
            if (r < dur) x = r * (0.5f / dur);
            else if (r > dur) x = 0.5f + (r - dur) * (0.5f / (1.f - dur));
            clk = sinf(6.28318530717958647692f * x);
            p.clock[m_idx * 4 + leg] = clk;
        } else clk = p.clock[m_idx * 4 + leg];
        push = p.push_interval > 0 && ((step_count + 1u) % (unsigned)p.push_interval) == 0u;
        const float *cfb = rs + RS_FORCE;                              // base body: what the last substep stored to p.contact
        if (sqrtf(cfb[0] * cfb[0] + cfb[1] * cfb[1] + cfb[2] * cfb[2]) > 1.f) f |= 16;
        get_euler_xyz(q4, rpy);
        float r0 = rpy[0], r1 = rpy[1];
        if (r0 > PI) r0 -= 2.f * PI;
        if (r1 > PI) r1 -= 2.f * PI;
        const float z = pos.z - oz;
        if (fabsf(r0) > p.term_roll) f |= 1;
        if (fabsf(r1) > p.term_pitch) f |= 2;
        if (z < p.term_zlow) f |= 4;
        if (z > p.term_zhigh) f |= 8;
        const float chk = pos.x + pos.y + pos.z + q4[0] + q4[1] + q4[2] + q4[3] + lv.x + lv.y + lv.z + av.x + av.y + av.z;
        if (!(fabsf(chk) < 1e6f)) f |= 32;
        if (push && leg < 2) {
            pushed = (2.f * rng_uniform(p.seed, (uint32_t)(p.env_off + env), step_count, RNG_PUSH, 2 * ag + leg) - 1.f) * p.max_push_vel;
            p.root[((size_t)env * GA + ag) * 13 + 7 + leg] = pushed;
        }
    }
    int fe = 0;
    for (int a2 = 0; a2 < A; a2++) fe |= __shfl_sync(FULL, f, min(31, lead + 4 * a2));
    int reset = 0;
    if (robot_live && ag == 0 && leg == 0) reset = dev_post_env_decide(p, env, fe, step_count, ep_pre);
    reset = __shfl_sync(FULL, reset, lead);
    if (robot_live && ag == 0) dev_post_env_npc_reset(p, env, reset, leg, quad_mask, step_count);
    __syncwarp();                                                      // the reset wrote state / last_actions of every agent of the env
    if (robot_live && reset) dev_post_agent_finish(p, env, ag, leg, quad_mask);       // rare: the row of a freshly reset env, from memory
    else if (robot_live) {
        float *ob = p.obs + (size_t)m_idx * MQE_OBS_FLOATS;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int j = 3 * leg + k;
            ob[MQE_OBS_DOF_POS + j] = q[k] - md->q_default[j];
            ob[MQE_OBS_DOF_VEL + j] = qd[k] * 0.05f;
            ob[MQE_OBS_LAST_ACTION + j] = act[k];
            ob[MQE_OBS_LAST_LAST_ACTION + j] = la[k];
            p.last_actions[m_idx * 12 + j] = act[k];
            p.last_dof_vel[m_idx * 12 + j] = qd[k];
        }
        ob[MQE_OBS_CLOCK + leg] = clk;
        if (leg == 0) {
            ob[MQE_OBS_BASE_POS] = pos.x - org[0]; ob[MQE_OBS_BASE_POS + 1] = pos.y - org[1]; ob[MQE_OBS_BASE_POS + 2] = pos.z - org[2];
            ob[MQE_OBS_BASE_QUAT] = qx; ob[MQE_OBS_BASE_QUAT + 1] = qy; ob[MQE_OBS_BASE_QUAT + 2] = qz; ob[MQE_OBS_BASE_QUAT + 3] = qw;
        } else if (leg == 1) {
            ob[MQE_OBS_LIN_VEL] = lv.x * 2.0f; ob[MQE_OBS_LIN_VEL + 1] = lv.y * 2.0f; ob[MQE_OBS_LIN_VEL + 2] = lv.z * 2.0f;
            ob[MQE_OBS_ANG_VEL] = av.x * 0.25f; ob[MQE_OBS_ANG_VEL + 1] = av.y * 0.25f; ob[MQE_OBS_ANG_VEL + 2] = av.z * 0.25f;
        } else if (leg == 2) {
            ob[MQE_OBS_PROJ_GRAVITY] = pg.x; ob[MQE_OBS_PROJ_GRAVITY + 1] = pg.y; ob[MQE_OBS_PROJ_GRAVITY + 2] = pg.z;
        } else {
            ob[MQE_OBS_BASE_RPY] = rpy[0]; ob[MQE_OBS_BASE_RPY + 1] = rpy[1]; ob[MQE_OBS_BASE_RPY + 2] = rpy[2];
        }
        // last_root_vel: the root velocity AFTER a push (legged_robot.py:146 reads root_states once _push_robots has run)
        float *lr = p.last_root_vel + m_idx * 6;
        if (leg == 0) { lr[0] = push ? pushed : vlin.x; lr[4] = wang.y; }
        else if (leg == 1) { lr[1] = push ? pushed : vlin.y; lr[5] = wang.z; }
        else if (leg == 2) lr[2] = vlin.z;
        else lr[3] = wang.x;
    }
    if (lane == 0) {
        // The last warp to arrive advances the step counter.  No fence: every warp's read of ctr[1] has returned before it gets here (its
        // value addressed the stores above), and nothing else in this kernel is ordered against the counter.
        if (atomicAdd(&p.ctr[2], 1) == (p.N + E - 1) / E - 1) { p.ctr[2] = 0; p.ctr[1] += 1; }
    }
}

__global__ void __launch_bounds__(256) k_substeps(const __grid_constant__ DevParams p, int nsub, int maxpair, int spair, int max_cand) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const unsigned FULL = 0xffffffffu;
    pdl_launch_dependents();
    // ---- CTA header: robot model + actuator weights + per-leg probe / capsule lists, one 8 KB bulk copy (api.cu builds it) ----
    __shared__ uint64_t hdr_bar;
    MqeRobotModel *md = reinterpret_cast<MqeRobotModel *>(smem);
    float *actw = smem + sizeof(MqeRobotModel) / 4;
    int *tbl = reinterpret_cast<int *>(actw + ACTW_FLOATS);
    if (threadIdx.x == 0) {
        mbar_init(&hdr_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&hdr_bar, (uint32_t)physics_cta_header_floats() * 4u);
        bulk_g2s(smem, p.substep_hdr, (uint32_t)physics_cta_header_floats() * 4u, &hdr_bar);
    }
    __syncthreads();
    mbar_wait(&hdr_bar, 0);
    pdl_wait();                                         // constants above were staged in the predecessor's shadow
    const int A = p.A, P = p.Pd, E = p.E, G = A + P;   // P: NPCs that own a lane
    const int GA = p.G;                                 // actors per env in the root-state tensor
    const bool seesaw = p.npc_kind == MQE_NPC_SEESAW;   // the NPC lane is a 1-DOF plank on a fixed base (seesaw.urdf)
    const bool box = p.npc_kind == MQE_NPC_BOX;         // the NPC lane is a free box (box.urdf)
    const bool has_fix = seesaw || p.npc_kind == MQE_NPC_PLATFORM;   // probes see a fixed NPC base (seesaw platform / column, raised boxes)
    const bool obb = seesaw || box;                     // robot probes collide with an oriented box instead of capsules
    const int Gc = obb ? A : G;                         // groups that take part in the capsule / capsule phase
    float *wbase = smem + physics_cta_header_floats() + warp * physics_warp_smem_floats(A, P, E, spair, maxpair);
    // Which env group this warp integrates.  With a task order (k_balance_tasks: groups sorted by the time their warp took in the last
    // launch, longest first) the warps of a CTA carry similar loads -- they wait less for each other at the alignment barrier -- and
    // the longest CTAs of a multi-round grid start first.  Results do not depend on the order: an env never looks at another one.
    const int task = blockIdx.x * nwarps + warp, ntasks = (p.N + E - 1) / E;
    if (task >= ntasks) return;
    const int first_env = (p.task_order ? p.task_order[task] : task) * E;
    long long t_start;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    // optional per-phase cycle trace (MQE_TRACE=1): lane 0 accumulates clock deltas straight into the warp's trace row
    long long *const tr = p.warp_trace + (size_t)(first_env / E) * MQE_TRACE_COLS;
    unsigned t_phase = 0;
    if (p.trace && lane == 0) {
        for (int i = 4; i < MQE_TRACE_COLS; i++) tr[i] = 0;
        t_phase = (unsigned)clock();
    }
#define PHASE_MARK(k)                                                                   \
    if (p.trace && lane == 0) {                                                         \
        const unsigned now_ = (unsigned)clock();                                        \
        tr[4 + (k)] += (long long)(now_ - t_phase);                                     \
        t_phase = now_;                                                                 \
    }

    // ---- lane roles ----
    const int nrl = 4 * A * E;
    const bool is_robot = lane < nrl;
    const bool is_npc = !is_robot && lane < nrl + P * E;
    int e_loc = 0, ag = 0, leg = 0, pn = 0;
    if (is_robot) { e_loc = lane / (4 * A); ag = (lane % (4 * A)) / 4; leg = lane & 3; }
    else if (is_npc) { e_loc = (lane - nrl) / P; pn = (lane - nrl) % P; }
    const int env = first_env + e_loc;
    const bool active = (is_robot || is_npc) && env < p.N;
    const int grp = is_robot ? ag : A + pn;                              // group index inside the env
    float *rs = wbase + (e_loc * A + ag) * RS_SIZE;                      // my robot block
    float *ns = wbase + E * A * RS_SIZE + (e_loc * P + pn) * NS_SIZE;    // my npc block
    float *const pool = wbase + E * A * RS_SIZE + E * P * NS_SIZE;
    int *const capmask = reinterpret_cast<int *>(pool + E * spair * 3 * PROWF) + e_loc * ES_MASKSZ(G);
    int my_start = e_loc * spair, my_smem = spair;                      // this env's slice of the pool (contacts), reassigned per substep
    // row stores: shared memory first, global scratch for the overflow (generic pointers; same layout in both)
    float *const grows = p.row_scratch + (size_t)(env * A + ag) * ((MQE_MAX_ROWS - SROWS) * ROWF);
    float *const gprows = p.prow_scratch + (size_t)env * ((size_t)maxpair * 3 * PROWF);
    float *const pdesc = p.pdesc_scratch + (size_t)env * ((size_t)maxpair * PDESCF);
    auto lrow = [&](int i) -> float * { return i < SROWS ? rs + RS_ROWS + i * ROWF : grows + (i - SROWS) * ROWF; };
    auto prow = [&](int i) -> float * { return i < 3 * my_smem ? pool + (3 * my_start + i) * PROWF : gprows + (i - 3 * my_smem) * PROWF; };
    const unsigned quad_mask = is_robot ? (0xFu << (lane & ~3)) : (1u << lane);
    unsigned env_mask = 0;
    {
        unsigned rm = (4 * A >= 32) ? FULL : ((1u << (4 * A)) - 1u);
        env_mask = rm << (e_loc * 4 * A);
        if (P) env_mask |= ((1u << P) - 1u) << (nrl + e_loc * P);
        if (!(is_robot || is_npc)) env_mask = 1u << lane;
    }
    const int rank_in_env = is_robot ? (lane - e_loc * 4 * A) : (4 * A + pn);
    const int lanes_per_env = 4 * A + P;

    // ---- persistent state in registers ----
    V3 pos = mk(0, 0, 0), vlin = mk(0, 0, 0), wang = mk(0, 0, 0);
    float qx = 0, qy = 0, qz = 0, qw = 1;
    float q[3] = {0, 0, 0}, qd[3] = {0, 0, 0}, act[3] = {0, 0, 0};
    float e1[3], e2[3], v1[3], v2[3], tau[3] = {0, 0, 0};
    const int m_idx = env * A + ag;                                      // agent row
    if (active) {
        const float *r = p.root + ((size_t)env * GA + grp) * 13;
        pos = mk(r[0], r[1], r[2]); qx = r[3]; qy = r[4]; qz = r[5]; qw = r[6];
        vlin = mk(r[7], r[8], r[9]); wang = mk(r[10], r[11], r[12]);
    }
    if (active && is_robot) {
        const float *d = p.dof + ((size_t)env * (12 * A + p.D) + 12 * ag + 3 * leg) * 2;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            q[k] = d[2 * k]; qd[k] = d[2 * k + 1];
            int j = m_idx * 12 + 3 * leg + k;
            act[k] = p.actions[j];
            e1[k] = p.err1[j]; e2[k] = p.err2[j]; v1[k] = p.vel1[j]; v2[k] = p.vel2[j];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 3; k++) { e1[k] = e2[k] = v1[k] = v2[k] = 0.f; }
    }
    V3 fixb = mk(0, 0, 0);                                               // seesaw: position of the fixed base (platform)
    if (has_fix && env < p.N) {
        const float *r = p.root + ((size_t)env * GA + A) * 13;
        fixb = mk(r[0], r[1], r[2]);
        if (seesaw && is_npc) { const float *d = p.dof + ((size_t)env * (12 * A + p.D) + 12 * A) * 2; q[0] = d[0]; qd[0] = d[1]; }
    }
    const float mu_e = (p.mu_env && env < p.N) ? p.mu_env[env] : p.mu;    // domain_rand.randomize_friction: one coefficient per env
    const int lag_c0 = p.lag_ring ? p.ctr[3] : 0;                          // _compute_torques calls before this launch
    int stat_local = 0, stat_lim = 0, stat_pair = 0, stat_rows = 0;

    PHASE_MARK(6);
    // warps of the CTA that own envs (the others returned above); they re-align at phase boundaries so that the seven warps of
    // an SM walk the 240 KB instruction stream together instead of each missing the instruction cache on its own
    const int live_threads = 32 * min(nwarps, (p.N + E - 1) / E - (int)blockIdx.x * nwarps);
#define CTA_ALIGN(level)                                                                 \
    if (p.cta_sync >= (level)) asm volatile("bar.sync 1, %0;" ::"r"(live_threads) : "memory");
    for (int sub = 0; sub < nsub; sub++) {
        const bool last = (sub == nsub - 1);
        if (p.cta_sync != 3 || (sub & 1) == 0) { CTA_ALIGN(1); }      // MQE_CTA_SYNC=3: re-align every other substep only (experiment)
        // ================================================================ P1: actuator network (go1.py:315-354, 369-380)
        if (active && is_robot && p.control_type != 0) {
            // LeggedRobot._compute_torques (legged_robot.py:384-392): 'P' PD towards action * scale + default pose, 'T' scaled torques,
            // 'V' PD on the joint velocity with the velocity of the previous POLICY step as the derivative reference (last_dof_vel)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int j = 3 * leg + k;
                const float a = (p.motor_strength ? p.motor_strength[m_idx * 12 + j] : 1.f) * act[k] * p.action_scale;     // legged_robot_field.py:180-183
                float t = a;
                if (p.control_type == 1) t = p.kp * (a + md->q_default[j] - q[k]) - p.kd * qd[k];
                else if (p.control_type == 3) t = p.kp * (a - qd[k]) - p.kd * (qd[k] - p.last_dof_vel[m_idx * 12 + j]) / p.dt;
                const float lim = md->tau_limit[j];
                tau[k] = fminf(fmaxf(t, -lim), lim);
            }
        } else if (p.act_mma) {
            // ---- tensor-core path.  Rows of the batched MLP: (joint k, lane L) -> m-tile (k, L / 16), row L % 16.
            float x[3][6];
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int i = 0; i < 6; i++) x[k][i] = 0.f;
            if (active && is_robot) {
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    int j = 3 * leg + k;
                    float a = act[k] * p.action_scale;
                    if (k == 0) a *= p.hip_scale;
                    if (p.lag_ring) {      // lag_buffer = lag_buffer[1:] + [actions_scaled]; target = lag_buffer[0] (go1.py:337-339), as a ring
                        float *ring = p.lag_ring + (size_t)m_idx * p.lag_n * 12 + j;
                        const int c = lag_c0 + sub;
                        ring[(c % p.lag_n) * 12] = a;
                        a = ring[((c + 1) % p.lag_n) * 12];
                    }
                    float err = q[k] - (a + md->q_default[j]);
                    x[k][0] = err; x[k][1] = e1[k]; x[k][2] = e2[k]; x[k][3] = qd[k]; x[k][4] = v1[k]; x[k][5] = v2[k];
                    e2[k] = e1[k]; e1[k] = err; v2[k] = v1[k]; v1[k] = qd[k];
                }
            }
            // stage the inputs as fp16 hi / lo rows [96][8] in row storage that is dead until P3 (robots 0 and 1 of the warp)
            uint32_t *xh = reinterpret_cast<uint32_t *>(wbase + RS_ROWS);
            uint32_t *xl = reinterpret_cast<uint32_t *>(wbase + RS_SIZE + RS_ROWS);
            float *outs = wbase + RS_SIZE + RS_ROWS + 384;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                uint4 h4, l4;
                split_h2(x[k][0], x[k][1], h4.x, l4.x); split_h2(x[k][2], x[k][3], h4.y, l4.y); split_h2(x[k][4], x[k][5], h4.z, l4.z);
                h4.w = 0u; l4.w = 0u;
                reinterpret_cast<uint4 *>(xh)[k * 32 + lane] = h4;
                reinterpret_cast<uint4 *>(xl)[k * 32 + lane] = l4;
            }
            __syncwarp();
            const int fg = lane >> 2, ft = lane & 3;
            const uint4 *frag = reinterpret_cast<const uint4 *>(actw);
            const float *fb0 = actw + 1280, *fb1 = actw + 1312, *fw2 = actw + 1344;
            const uint4 w0h = frag[lane], w0l = frag[32 + lane];
            const int mtiles = 2 * 3;
#pragma unroll 1
            for (int mt = 0; mt < mtiles; mt++) {
                const int k = mt >> 1, hf = mt & 1;
                if (16 * hf >= nrl) continue;                        // no robot lane in this half of the warp
                const int r0 = (k * 32 + 16 * hf + fg) * 4 + ft;
                const uint32_t ah0 = xh[r0], ah1 = xh[r0 + 32], al0 = xl[r0], al1 = xl[r0 + 32];
                float c[4][4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float2 b = *reinterpret_cast<const float2 *>(fb0 + 8 * j + 2 * ft);
                    c[j][0] = c[j][2] = b.x; c[j][1] = c[j][3] = b.y;
                }
                const uint32_t w0hj[4] = {w0h.x, w0h.y, w0h.z, w0h.w}, w0lj[4] = {w0l.x, w0l.y, w0l.z, w0l.w};
#pragma unroll
                for (int j = 0; j < 4; j++) { mma_k8(c[j], ah0, ah1, w0hj[j]); mma_k8(c[j], ah0, ah1, w0lj[j]); mma_k8(c[j], al0, al1, w0hj[j]); }
                uint32_t a2h[2][4], a2l[2][4];
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int i = 0; i < 4; i++) c[j][i] = softsign_fast(c[j][i]);
#pragma unroll
                for (int s2 = 0; s2 < 2; s2++) {                     // C fragments of n-tiles 2s, 2s+1 ARE the A fragment of k-step s
                    split_h2(c[2 * s2][0], c[2 * s2][1], a2h[s2][0], a2l[s2][0]); split_h2(c[2 * s2][2], c[2 * s2][3], a2h[s2][1], a2l[s2][1]);
                    split_h2(c[2 * s2 + 1][0], c[2 * s2 + 1][1], a2h[s2][2], a2l[s2][2]); split_h2(c[2 * s2 + 1][2], c[2 * s2 + 1][3], a2h[s2][3], a2l[s2][3]);
                }
                float vlo = 0.f, vhi = 0.f;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float2 b = *reinterpret_cast<const float2 *>(fb1 + 8 * j + 2 * ft);
                    float d[4] = {b.x, b.y, b.x, b.y};
#pragma unroll
                    for (int s2 = 0; s2 < 2; s2++) {
                        const uint4 w = frag[(2 + j * 2 + s2) * 32 + lane];      // hi.b0 hi.b1 lo.b0 lo.b1
                        mma_k16(d, a2h[s2], w.x, w.y); mma_k16(d, a2h[s2], w.z, w.w); mma_k16(d, a2l[s2], w.x, w.y);
                    }
                    const float2 w2 = *reinterpret_cast<const float2 *>(fw2 + 8 * j + 2 * ft);
                    vlo = fmaf(w2.y, softsign_fast(d[1]), fmaf(w2.x, softsign_fast(d[0]), vlo));
                    vhi = fmaf(w2.y, softsign_fast(d[3]), fmaf(w2.x, softsign_fast(d[2]), vhi));
                }
                vlo += __shfl_xor_sync(FULL, vlo, 1); vhi += __shfl_xor_sync(FULL, vhi, 1);
                vlo += __shfl_xor_sync(FULL, vlo, 2); vhi += __shfl_xor_sync(FULL, vhi, 2);
                if (ft == 0) { outs[k * 32 + 16 * hf + fg] = vlo + actw[1376]; outs[k * 32 + 16 * hf + fg + 8] = vhi + actw[1376]; }
            }
            __syncwarp();
            if (active && is_robot) {
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const float lim = md->tau_limit[3 * leg + k];
                    tau[k] = fminf(fmaxf(outs[k * 32 + lane], -lim), lim);
                }
            }
            __syncwarp();                                            // the staging area becomes row storage again in P3
        } else if (active && is_robot) {
            float x[3][6];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                int j = 3 * leg + k;
                float a = act[k] * p.action_scale;
                if (k == 0) a *= p.hip_scale;
                if (p.lag_ring) {      // lag_buffer = lag_buffer[1:] + [actions_scaled]; target = lag_buffer[0] (go1.py:337-339), as a ring
                    float *ring = p.lag_ring + (size_t)m_idx * p.lag_n * 12 + j;
                    const int c = lag_c0 + sub;
                    ring[(c % p.lag_n) * 12] = a;
                    a = ring[((c + 1) % p.lag_n) * 12];
                }
                float err = q[k] - (a + md->q_default[j]);
                x[k][0] = err; x[k][1] = e1[k]; x[k][2] = e2[k]; x[k][3] = qd[k]; x[k][4] = v1[k]; x[k][5] = v2[k];
                e2[k] = e1[k]; e1[k] = err; v2[k] = v1[k]; v1[k] = qd[k];
            }
            // Rank-1 formulation: for every hidden unit i of layer 1, h_i (3 joints) updates all 32 layer-2 pre-activations at
            // once.  96 independent accumulators instead of 3 dependent chains (the kernel is latency-bound), a 32-trip loop
            // whose body stays in the instruction cache, and packed FFMA2 (two fp32 FMAs per instruction on sm_100).
            // Packed weights: W0[32][6] b0[32] W1T[32 in][32 out] b1[32] W2[32] b2.
            const float *W0 = actw, *b0 = actw + 192, *W1T = actw + 224, *b1 = actw + 1248, *W2 = actw + 1280;
            unsigned long long acc[3][16];
#pragma unroll
            for (int o2 = 0; o2 < 16; o2++) {
                const unsigned long long bb = pack2(b1[2 * o2], b1[2 * o2 + 1]);
                acc[0][o2] = bb; acc[1][o2] = bb; acc[2][o2] = bb;
            }
#pragma unroll 2
            for (int i = 0; i < 32; i++) {
                const float w0 = W0[i * 6], w1 = W0[i * 6 + 1], w2 = W0[i * 6 + 2], w3 = W0[i * 6 + 3], w4 = W0[i * 6 + 4], w5 = W0[i * 6 + 5], b = b0[i];
                unsigned long long hh[3];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const float h = softsign(b + w0 * x[k][0] + w1 * x[k][1] + w2 * x[k][2] + w3 * x[k][3] + w4 * x[k][4] + w5 * x[k][5]);
                    hh[k] = pack2(h, h);
                }
                const ulonglong2 *wr = reinterpret_cast<const ulonglong2 *>(W1T + i * 32);
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    const ulonglong2 w = wr[g];                      // (W1[4g][i], W1[4g+1][i]), (W1[4g+2][i], W1[4g+3][i])
#pragma unroll
                    for (int k = 0; k < 3; k++) { ffma2(acc[k][2 * g], w.x, hh[k]); ffma2(acc[k][2 * g + 1], w.y, hh[k]); }
                }
            }
            float out[3] = {actw[1312], actw[1312], actw[1312]};
#pragma unroll
            for (int o2 = 0; o2 < 16; o2++) {
                const float wa = W2[2 * o2], wb = W2[2 * o2 + 1];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    float lo, hi;
                    unpack2(acc[k][o2], lo, hi);
                    out[k] = fmaf(wa, softsign(lo), out[k]);
                    out[k] = fmaf(wb, softsign(hi), out[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float lim = md->tau_limit[3 * leg + k];
                tau[k] = fminf(fmaxf(out[k], -lim), lim);
            }
        }

        PHASE_MARK(0);
        CTA_ALIGN(2);
        // ================================================================ P2: kinematics + dynamics (robot lanes), npc prediction
        float vb[6] = {0, 0, 0, 0, 0, 0}, u[3] = {0, 0, 0};   // solve coordinates
        float Gm[18];                                           // my leg's G (6x3 row-major)
        M3 Rb, R1, R2, R3;
        V3 p1 = mk(0, 0, 0), p2 = p1, p3 = p1, a1 = p1, a2 = p1;
        float Sinv[21], Hinv[6];
        if (is_robot) {   // executed by whole quads (inactive envs compute on zeros; quad shuffles need all four lanes)
            Rb = quat_to_mat(qx, qy, qz, qw);
            float s1, c1, s2, c2, s3, c3;
            sincosf(q[0], &s1, &c1); sincosf(q[1], &s2, &c2); sincosf(q[2], &s3, &c3);
            const float *off = &md->leg_offsets[leg][0][0];
            p1 = mul(Rb, mk(off[0], off[1], off[2]));
            a1 = Rb.c0;
            R1 = rot_x(Rb, c1, s1);
            p2 = p1 + mul(R1, mk(off[3], off[4], off[5]));
            a2 = R1.c1;
            R2 = rot_y(R1, c2, s2);
            p3 = p2 + mul(R2, mk(off[6], off[7], off[8]));
            R3 = rot_y(R2, c3, s3);
            SV S1, S2, S3;
            S1.w = a1; S1.v = cross(p1, a1);
            S2.w = a2; S2.v = cross(p2, a2);
            S3.w = a2; S3.v = cross(p3, a2);
            V3 com_sh = mk(0.f, 0.f, 0.f);
            if (p.base_com_shift && active) com_sh = mk(p.base_com_shift[m_idx * 3], p.base_com_shift[m_idx * 3 + 1], p.base_com_shift[m_idx * 3 + 2]);
            RBI I0 = rbi_from_link(md->base_inertial, Rb, mk(0, 0, 0), (p.base_mass_add && active) ? p.base_mass_add[m_idx] : 0.f, com_sh);
            RBI I1 = rbi_from_link(md->leg_inertial[leg][0], R1, p1);
            RBI I2 = rbi_from_link(md->leg_inertial[leg][1], R2, p2);
            RBI I3 = rbi_from_link(md->leg_inertial[leg][2], R3, p3);
            // velocities, bias accelerations
            SV V0; V0.w = wang; V0.v = vlin;
            SV A0; A0.w = mk(0, 0, 0); A0.v = mk(0, 0, -p.gz);
            SV sq1 = svscale(qd[0], S1), sq2 = svscale(qd[1], S2), sq3 = svscale(qd[2], S3);
            SV V1 = svadd(V0, sq1), V2 = svadd(V1, sq2), V3_ = svadd(V2, sq3);
            SV A1 = svadd(A0, crm(V0, sq1)), A2 = svadd(A1, crm(V1, sq2)), A3 = svadd(A2, crm(V2, sq3));
            SV f0 = svadd(rbi_mul(I0, A0), crf(V0, rbi_mul(I0, V0)));
            SV f1 = svadd(rbi_mul(I1, A1), crf(V1, rbi_mul(I1, V1)));
            SV f2 = svadd(rbi_mul(I2, A2), crf(V2, rbi_mul(I2, V2)));
            SV f3 = svadd(rbi_mul(I3, A3), crf(V3_, rbi_mul(I3, V3_)));
            f2 = svadd(f2, f3); f1 = svadd(f1, f2);
            float cl[3] = {svdot(S1, f1), svdot(S2, f2), svdot(S3, f3)};
            // composite inertias, F columns, leg block H
            RBI I2c = I2; rbi_add(I2c, I3);
            RBI I1c = I1; rbi_add(I1c, I2c);
            SV F1 = rbi_mul(I1c, S1), F2 = rbi_mul(I2c, S2), F3 = rbi_mul(I3, S3);
            float H[6] = {svdot(S1, F1), svdot(S1, F2), svdot(S1, F3), svdot(S2, F2), svdot(S2, F3), svdot(S3, F3)};
            sym3_inverse(H, Hinv);
            float Fm[18];
#pragma unroll
            for (int i = 0; i < 6; i++) { Fm[i * 3] = svc(F1, i); Fm[i * 3 + 1] = svc(F2, i); Fm[i * 3 + 2] = svc(F3, i); }
#pragma unroll
            for (int i = 0; i < 6; i++) {
                float x0 = Fm[i * 3], x1 = Fm[i * 3 + 1], x2 = Fm[i * 3 + 2];
                Gm[i * 3] = x0 * Hinv[0] + x1 * Hinv[1] + x2 * Hinv[2];
                Gm[i * 3 + 1] = x0 * Hinv[1] + x1 * Hinv[3] + x2 * Hinv[4];
                Gm[i * 3 + 2] = x0 * Hinv[2] + x1 * Hinv[4] + x2 * Hinv[5];
            }
            // base-level reductions over the quad
            RBI Ic = I1c;
            Ic.m = quad_sum(Ic.m, quad_mask);
            Ic.h = mk(quad_sum(Ic.h.x, quad_mask), quad_sum(Ic.h.y, quad_mask), quad_sum(Ic.h.z, quad_mask));
#pragma unroll
            for (int i = 0; i < 6; i++) Ic.I[i] = quad_sum(Ic.I[i], quad_mask);
            rbi_add(Ic, I0);
            float Ssch[21];
            {
                const float hx = Ic.h.x, hy = Ic.h.y, hz = Ic.h.z;
                float M6[21];
                M6[sidx(0, 0)] = Ic.I[0]; M6[sidx(0, 1)] = Ic.I[1]; M6[sidx(0, 2)] = Ic.I[2];
                M6[sidx(1, 1)] = Ic.I[3]; M6[sidx(1, 2)] = Ic.I[4]; M6[sidx(2, 2)] = Ic.I[5];
                M6[sidx(0, 3)] = 0.f; M6[sidx(0, 4)] = -hz; M6[sidx(0, 5)] = hy;
                M6[sidx(1, 3)] = hz;  M6[sidx(1, 4)] = 0.f; M6[sidx(1, 5)] = -hx;
                M6[sidx(2, 3)] = -hy; M6[sidx(2, 4)] = hx;  M6[sidx(2, 5)] = 0.f;
                M6[sidx(3, 3)] = Ic.m; M6[sidx(3, 4)] = 0.f; M6[sidx(3, 5)] = 0.f;
                M6[sidx(4, 4)] = Ic.m; M6[sidx(4, 5)] = 0.f; M6[sidx(5, 5)] = Ic.m;
#pragma unroll
                for (int i = 0; i < 6; i++)
#pragma unroll
                    for (int j = i; j < 6; j++) {
                        float gf = Gm[i * 3] * Fm[j * 3] + Gm[i * 3 + 1] * Fm[j * 3 + 1] + Gm[i * 3 + 2] * Fm[j * 3 + 2];
                        Ssch[sidx(i, j)] = M6[sidx(i, j)] - quad_sum(gf, quad_mask);
                    }
            }
            spd6_inverse(Ssch, Sinv);
            // right-hand sides
            float rl[3] = {tau[0] - cl[0], tau[1] - cl[1], tau[2] - cl[2]};
            float rb[6];
#pragma unroll
            for (int i = 0; i < 6; i++) {
                float legpart = svc(f1, i) + Gm[i * 3] * rl[0] + Gm[i * 3 + 1] * rl[1] + Gm[i * 3 + 2] * rl[2];
                rb[i] = -(svc(f0, i) + quad_sum(legpart, quad_mask));
            }
            float ab[6], au[3];
            sym6_mulv(Sinv, rb, ab);
            sym3_mulv(Hinv, rl, au);
            // to solve coordinates and predict
            V3 wxv = cross(wang, vlin);
            vb[0] = wang.x + p.dt * ab[0]; vb[1] = wang.y + p.dt * ab[1]; vb[2] = wang.z + p.dt * ab[2];
            vb[3] = vlin.x + p.dt * (ab[3] + wxv.x); vb[4] = vlin.y + p.dt * (ab[4] + wxv.y); vb[5] = vlin.z + p.dt * (ab[5] + wxv.z);
            // u = qd + G^T v_base must be formed with the SAME base velocity the classical-acceleration term w x v
            // was added to, otherwise qd = u - G^T v_base picks up a spurious -dt G_lin^T (w x v)
            float v0[6] = {wang.x, wang.y, wang.z, vlin.x + p.dt * wxv.x, vlin.y + p.dt * wxv.y, vlin.z + p.dt * wxv.z};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float gt = 0.f;
#pragma unroll
                for (int i = 0; i < 6; i++) gt += Gm[i * 3 + k] * v0[i];
                u[k] = qd[k] + gt + p.dt * au[k];
            }
            // publish operators for row builders of other lanes (pair contacts) and the capsule endpoints
            if (leg == 0) {
                rs[RS_ORIGIN] = pos.x; rs[RS_ORIGIN + 1] = pos.y; rs[RS_ORIGIN + 2] = pos.z;
#pragma unroll
                for (int i = 0; i < 21; i++) rs[RS_SINV + i] = Sinv[i];
                ((int *)rs)[RS_CNT] = 0;
                for (int i = 0; i < 51; i++) rs[RS_FORCE + i] = 0.f;
            }
        } else if (is_npc && seesaw) {
            // passive revolute joint: y hinge (seesaw plank, gravity torque r_x m g) or z hinge (revolving door, no gravity torque);
            // I_pivot = I_axis + m r_perp^2
            float sn, cs;
            sincosf(q[0], &sn, &cs);
            const bool hz = p.geom[13] > 0.5f;
            const float r2 = hz ? p.geom[3] * p.geom[3] + p.geom[14] * p.geom[14] : p.geom[3] * p.geom[3] + p.geom[15] * p.geom[15];
            const float Ip = p.npc_inertia + p.npc_mass * r2;
            const float tau_g = hz ? 0.f : p.geom[3] * cs * p.npc_mass * (-p.gz);
            const bool pz = p.geom[13] > 1.5f;                              // prismatic y joint (tug disc): q is a displacement
            if (pz) vb[4] = qd[0];
            else vb[hz ? 2 : 1] = qd[0] + p.dt * tau_g / Ip;
            ns[NS_ORIGIN] = pos.x + p.geom[0]; ns[NS_ORIGIN + 1] = pos.y + p.geom[1] + (pz ? q[0] : 0.f); ns[NS_ORIGIN + 2] = pos.z + p.geom[2];
            ns[NS_CAP] = cs; ns[NS_CAP + 1] = sn;
            ((int *)ns)[NS_CNT] = 0;
            ns[NS_BOUND] = 0.f;
            ns[NS_FORCE] = ns[NS_FORCE + 1] = ns[NS_FORCE + 2] = 0.f;
        } else if (is_npc) {
            vb[0] = wang.x; vb[1] = wang.y; vb[2] = wang.z; vb[3] = vlin.x; vb[4] = vlin.y; vb[5] = vlin.z + p.dt * p.gz;
            Rb = quat_to_mat(qx, qy, qz, qw);
            V3 ax = p.npc_halflen * Rb.c2;
            ns[NS_ORIGIN] = pos.x; ns[NS_ORIGIN + 1] = pos.y; ns[NS_ORIGIN + 2] = pos.z;
            ns[NS_CAP] = pos.x - ax.x; ns[NS_CAP + 1] = pos.y - ax.y; ns[NS_CAP + 2] = pos.z - ax.z;
            ns[NS_CAP + 3] = pos.x + ax.x; ns[NS_CAP + 4] = pos.y + ax.y; ns[NS_CAP + 5] = pos.z + ax.z; ns[NS_CAP + 6] = p.npc_radius;
            ((int *)ns)[NS_CNT] = 0;
            ns[NS_BOUND] = p.npc_radius + p.npc_halflen;
            if (box) {
                ns[NS_ROT] = Rb.c0.x; ns[NS_ROT + 1] = Rb.c0.y; ns[NS_ROT + 2] = Rb.c0.z; ns[NS_ROT + 3] = Rb.c1.x; ns[NS_ROT + 4] = Rb.c1.y;
                ns[NS_ROT + 5] = Rb.c1.z; ns[NS_ROT + 6] = Rb.c2.x; ns[NS_ROT + 7] = Rb.c2.y; ns[NS_ROT + 8] = Rb.c2.z;
            }
            ns[NS_FORCE] = ns[NS_FORCE + 1] = ns[NS_FORCE + 2] = 0.f;
        }
        my_start = e_loc * spair; my_smem = spair;
        __syncwarp();

        PHASE_MARK(1);
        CTA_ALIGN(2);
        // ================================================================ P3: rows.  joint limits, then world contacts
        // Row slots: a robot's joint-limit rows sit at [0, nlim), its contact blocks ALWAYS start at row CROW0 = MQE_MAX_LIMIT, so a block never
        // straddles the shared / global boundary (SROWS = CROW0 + 3 * 5) and one pointer select addresses all three of its rows; NPC rows start at 0.
        int nrows = 0, nlim = 0;
        const int crow0 = is_robot ? MQE_MAX_LIMIT : 0;
        if (is_robot) {
            // ---- joint limits: canonical order (dof, lower/upper), cap MQE_MAX_LIMIT ----
            unsigned lm = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                int j = 3 * leg + k;
                if (q[k] - md->q_lower[j] < p.limit_margin) lm |= 1u << (2 * j);
                if (md->q_upper[j] - q[k] < p.limit_margin) lm |= 1u << (2 * j + 1);
            }
            if (!active) lm = 0;
            unsigned lall = lm;
            lall |= __shfl_xor_sync(quad_mask, lall, 1);
            lall |= __shfl_xor_sync(quad_mask, lall, 2);
            nlim = min(__popc(lall), MQE_MAX_LIMIT);
            while (lm) {
                int b = __ffs(lm) - 1;
                lm &= lm - 1;
                int slot = __popc(lall & ((1u << b) - 1u));
                if (slot >= MQE_MAX_LIMIT) break;
                int k = (b >> 1) - 3 * leg, side = b & 1;
                float sg = side ? -1.f : 1.f;
                float gap = side ? md->q_upper[3 * leg + k] - q[k] : q[k] - md->q_lower[3 * leg + k];
                float Jb[6], Yb[6];
#pragma unroll
                for (int i = 0; i < 6; i++) Jb[i] = -sg * sel1(k, Gm[i * 3], Gm[i * 3 + 1], Gm[i * 3 + 2], 0.f);   // value selects keep Gm / Hinv in registers
                sym6_mulv(Sinv, Jb, Yb);
                const float hk[3] = {sel1(k, Hinv[0], Hinv[1], Hinv[2], 0.f), sel1(k, Hinv[1], Hinv[3], Hinv[4], 0.f), sel1(k, Hinv[2], Hinv[4], Hinv[5], 0.f)};
                float dd = sel1(k, hk[0], hk[1], hk[2], 0.f);
#pragma unroll
                for (int i = 0; i < 6; i++) dd += Jb[i] * Yb[i];
                float *row = lrow(slot);
#pragma unroll
                for (int i = 0; i < 6; i++) { row[i] = Jb[i]; row[9 + i] = Yb[i]; }
#pragma unroll
                for (int i = 0; i < 3; i++) { row[6 + i] = (i == k) ? sg : 0.f; row[15 + i] = sg * hk[i]; }
                row[18] = 1.f / (dd + p.cfm); row[19] = contact_bias(p, gap); row[20] = 0.f;
                row[21] = __int_as_float(leg | (0 << 4) | (slot << 8));
            }
            PHASE_MARK(8);
            // ---- world contacts: pass 1 flags ----
            const float sdf_o = active ? sdf_sample(p, pos.x, pos.y).sdf : 0.f;     // same address in all four lanes of the quad
            unsigned long long cm = 0ull;
            const int *pl_ = tbl + leg * 10;
            if (active) {
                const int np_ = pl_[0];
#pragma unroll 1          // measured: two probes in flight (unroll 2) cost more in instruction fetch than the overlap gives (+0.6 % step)
                for (int t = 0; t < np_; t++) {
                    const int pi = pl_[1 + t];
                    const float *pr = md->probes[pi];
                    const int link = (int)pr[0];
                    const int k = link == 0 ? 0 : link - 3 * leg;
                    const M3 Rl = selm(k, Rb, R1, R2, R3);
                    V3 pl = sel3(k, mk(0, 0, 0), p1, p2, p3);
                    V3 xr = pl + mul(Rl, mk(pr[2], pr[3], pr[4]));
                    ProbeHit h = probe_world(p, pos + xr, pr[5], has_fix, fixb, wall_is_far(p, sdf_o, xr, pr[5]));
                    cm |= (unsigned long long)h.mask << (2 * pi);
                }
            }
            unsigned long long call = cm;
            call |= __shfl_xor_sync(quad_mask, call, 1);
            call |= __shfl_xor_sync(quad_mask, call, 2);
            int ncon = min(__popcll(call), MQE_MAX_LOCAL);
            nrows = crow0 + 3 * ncon;
            PHASE_MARK(9);
            // ---- pass 2: build rows for my contacts ----
            while (cm) {
                int b = __ffsll((long long)cm) - 1;
                cm &= cm - 1ull;
                int slot = __popcll(call & ((1ull << b) - 1ull));
                if (slot >= MQE_MAX_LOCAL) break;
                int pi = b >> 1, kind = b & 1;
                const float *pr = md->probes[pi];
                int link = (int)pr[0], body = (int)pr[1];
                int k = link == 0 ? 0 : link - 3 * leg;
                const M3 Rl = selm(k, Rb, R1, R2, R3);
                V3 pl = sel3(k, mk(0, 0, 0), p1, p2, p3);
                V3 xr = pl + mul(Rl, mk(pr[2], pr[3], pr[4]));            // probe centre rel. O
                ProbeHit h = probe_world(p, pos + xr, pr[5], has_fix, fixb, wall_is_far(p, sdf_o, xr, pr[5]));
                V3 n = kind ? h.nw : mk(0, 0, 1);
                float gap = kind ? h.gap_w : h.gap_f;
                V3 r = xr - (pr[5] + 0.5f * gap) * n;                     // contact point rel. O
                V3 t1, t2;
                tangent_basis(n, t1, t2);
                float *cmeta = rs + RS_CMETA + slot * 4;
                cmeta[0] = n.x; cmeta[1] = n.y; cmeta[2] = n.z; cmeta[3] = __int_as_float(body);
                int r0 = crow0 + 3 * slot;
#pragma unroll 1          // rolled: the body (~850 instructions) then fits the 32 KB L1.5 instruction cache, unrolled it does not
                for (int dch = 0; dch < 3; dch++) {
                    V3 d = dch == 0 ? n : (dch == 1 ? t1 : t2);
                    V3 rxd = cross(r, d);
                    float Jb[6] = {rxd.x, rxd.y, rxd.z, d.x, d.y, d.z}, Jl[3] = {0, 0, 0}, Yb[6], Yl[3];
                    if (k >= 1) Jl[0] = dot(a1, cross(r - p1, d));
                    if (k >= 2) Jl[1] = dot(a2, cross(r - p2, d));
                    if (k >= 3) Jl[2] = dot(a2, cross(r - p3, d));
#pragma unroll
                    for (int i = 0; i < 6; i++) Jb[i] -= Gm[i * 3] * Jl[0] + Gm[i * 3 + 1] * Jl[1] + Gm[i * 3 + 2] * Jl[2];
                    sym6_mulv(Sinv, Jb, Yb);
                    sym3_mulv(Hinv, Jl, Yl);
                    float dd = Jl[0] * Yl[0] + Jl[1] * Yl[1] + Jl[2] * Yl[2];
#pragma unroll
                    for (int i = 0; i < 6; i++) dd += Jb[i] * Yb[i];
                    float *row = lrow(r0 + dch);
#pragma unroll
                    for (int i = 0; i < 6; i++) { row[i] = Jb[i]; row[9 + i] = Yb[i]; }
#pragma unroll
                    for (int i = 0; i < 3; i++) { row[6 + i] = Jl[i]; row[15 + i] = Yl[i]; }
                    row[18] = 1.f / (dd + p.cfm);
                    row[19] = dch == 0 ? contact_bias(p, gap) : 0.f;
                    row[20] = 0.f;
                    row[21] = __int_as_float(leg | ((dch ? 1 : 0) << 4) | (r0 << 8));
                }
                { float *R = lrow(r0); local_block_coupling(R, R + ROWF, R + 2 * ROWF); }
            }
            if (leg == 0 && active) { stat_local += ncon; stat_lim += nlim; }
        } else if (is_npc && active && seesaw) {
            int ncon = 0;
            const float cs = ns[NS_CAP], sn = ns[NS_CAP + 1];
            for (int en = 0; en < 2 && !(p.geom[13] > 0.5f); en++) {
                const float xe = p.geom[3] + (en == 0 ? -1.f : 1.f) * p.geom[4];
                V3 r = mk(xe * cs, 0.f, -xe * sn);                       // plank end rel. the pivot
                float gap = ns[NS_ORIGIN + 2] + r.z - p.geom[6] - p.floor_z;
                if (gap >= p.coff) continue;
                V3 n = mk(0, 0, 1), t1, t2;
                r.z -= p.geom[6] + 0.5f * gap;
                tangent_basis(n, t1, t2);
                float *cmeta = ns + NS_CMETA + ncon * 4;
                cmeta[0] = n.x; cmeta[1] = n.y; cmeta[2] = n.z; cmeta[3] = 0.f;
#pragma unroll 1
                for (int dch = 0; dch < 3; dch++) {
                    V3 d = dch == 0 ? n : (dch == 1 ? t1 : t2);
                    float *row = ns + NS_ROWS + (3 * ncon + dch) * ROWF;
                    float dd = npc_side(p, r, d, row);
                    row[18] = 1.f / (dd + p.cfm);
                    row[19] = dch == 0 ? contact_bias(p, gap) : 0.f;
                    row[20] = 0.f;
                    row[21] = __int_as_float(0 | ((dch ? 1 : 0) << 4) | ((3 * ncon) << 8));
                }
                local_block_coupling(ns + NS_ROWS + 3 * ncon * ROWF, ns + NS_ROWS + (3 * ncon + 1) * ROWF, ns + NS_ROWS + (3 * ncon + 2) * ROWF);
                ncon++;
            }
            nrows = 3 * ncon;
            stat_local += ncon;
        } else if (is_npc && active) {
            int ends = box ? 8 : (p.npc_halflen > 0.f ? 2 : 1), ncon = 0;    // box: its 8 corners as point probes
            const float prad = box ? 0.f : p.npc_radius;
            for (int en = 0; en < ends && ncon < 4; en++) {
                V3 xr = ((en == 0 ? -1.f : 1.f) * p.npc_halflen) * Rb.c2;
                if (box) xr = (((en & 1) ? 1.f : -1.f) * p.geom[4]) * Rb.c0 + (((en & 2) ? 1.f : -1.f) * p.geom[5]) * Rb.c1 + (((en & 4) ? 1.f : -1.f) * p.geom[6]) * Rb.c2;
                ProbeHit h = probe_world(p, pos + xr, prad);
                for (int kind = 0; kind < 2 && ncon < 4; kind++) {
                    if (!(h.mask & (1 << kind))) continue;
                    V3 n = kind ? h.nw : mk(0, 0, 1);
                    float gap = kind ? h.gap_w : h.gap_f;
                    V3 r = xr - (prad + 0.5f * gap) * n, t1, t2;
                    tangent_basis(n, t1, t2);
                    float *cmeta = ns + NS_CMETA + ncon * 4;
                    cmeta[0] = n.x; cmeta[1] = n.y; cmeta[2] = n.z; cmeta[3] = 0.f;
#pragma unroll 1
                    for (int dch = 0; dch < 3; dch++) {
                        V3 d = dch == 0 ? n : (dch == 1 ? t1 : t2);
                        float *row = ns + NS_ROWS + (3 * ncon + dch) * ROWF;
                        float dd = npc_side(p, r, d, row);
                        row[18] = 1.f / (dd + p.cfm);
                        row[19] = dch == 0 ? contact_bias(p, gap) : 0.f;
                        row[20] = 0.f;
                        row[21] = __int_as_float(0 | ((dch ? 1 : 0) << 4) | ((3 * ncon) << 8));
                    }
                    local_block_coupling(ns + NS_ROWS + 3 * ncon * ROWF, ns + NS_ROWS + (3 * ncon + 1) * ROWF, ns + NS_ROWS + (3 * ncon + 2) * ROWF);
                    ncon++;
                }
            }
            nrows = 3 * ncon;
            stat_local += ncon;
        }
        __syncwarp();

        PHASE_MARK(2);
        // ================================================================ P3b: dynamic-vs-dynamic pairs (capsule / capsule)
        int npair = 0;
        bool pairs_pending = false;                      // capsule contacts recorded in pdesc, rows not built yet
        bool need_narrow = false;                        // this env's capsule masks are live: run the narrow phase for it
        unsigned t_sub = 0;                              // sub-phase marks of P3b (MQE_TRACE=1): lane 0's clock at warp-uniform points
        if (p.trace && lane == 0) t_sub = (unsigned)clock();
        if (G > 1 && (is_robot || is_npc)) {
            // broadphase over group pairs
            bool close = false;
            int ngp = Gc * (Gc - 1) / 2;
            for (int t = rank_in_env; t < ngp; t += lanes_per_env) {
                int X = 0, rem = t;
                while (rem >= Gc - 1 - X) { rem -= Gc - 1 - X; X++; }
                int Y = X + 1 + rem;
                const float *ox = X < A ? wbase + (e_loc * A + X) * RS_SIZE + RS_ORIGIN : wbase + E * A * RS_SIZE + (e_loc * P + X - A) * NS_SIZE + NS_ORIGIN;
                const float *oy = Y < A ? wbase + (e_loc * A + Y) * RS_SIZE + RS_ORIGIN : wbase + E * A * RS_SIZE + (e_loc * P + Y - A) * NS_SIZE + NS_ORIGIN;
                float bx = X < A ? MQE_ROBOT_BOUND : p.npc_radius + p.npc_halflen, by = Y < A ? MQE_ROBOT_BOUND : p.npc_radius + p.npc_halflen;
                float lim = bx + by + p.coff, dx = ox[0] - oy[0], dy = ox[1] - oy[1], dz = ox[2] - oy[2];
                const bool c = dx * dx + dy * dy + dz * dz <= lim * lim;
                capmask[X * G + Y] = capmask[Y * G + X] = c ? 1 : 0;       // pre-flag: only pairs in range get their capsule masks computed
                close |= c;
            }
            if (env >= p.N) close = false;
            bool any_close = __ballot_sync(env_mask, close) != 0u;
#define SUB_MARK(k)                                                                                        \
    if (p.trace && lane == 0) {                                                                            \
        const unsigned now_ = (unsigned)clock();                                                           \
        tr[4 + (k)] += (long long)(now_ - t_sub);                                                          \
        t_sub = now_;                                                                                      \
    }
            if (any_close) {
                // publish the operators other lanes need to build pair rows, and the capsule end points (only now:
                // most substeps have no dynamic pair in range)
                if (is_robot) {
#pragma unroll
                    for (int i = 0; i < 18; i++) rs[RS_G + leg * 18 + i] = Gm[i];
#pragma unroll
                    for (int i = 0; i < 6; i++) rs[RS_HINV + leg * 6 + i] = Hinv[i];
                    float *pa = rs + RS_A + leg * 9, *pp = rs + RS_P + leg * 9;
                    pa[0] = a1.x; pa[1] = a1.y; pa[2] = a1.z; pa[3] = a2.x; pa[4] = a2.y; pa[5] = a2.z; pa[6] = a2.x; pa[7] = a2.y; pa[8] = a2.z;
                    pp[0] = p1.x; pp[1] = p1.y; pp[2] = p1.z; pp[3] = p2.x; pp[4] = p2.y; pp[5] = p2.z; pp[6] = p3.x; pp[7] = p3.y; pp[8] = p3.z;
                    const int *cl_ = tbl + 40 + leg * 10;
                    float bound = 0.f;
                    for (int t = 0; t < cl_[0]; t++) {
                        const int ci = cl_[1 + t];
                        const float *cp = md->caps[ci];
                        const int link = (int)cp[0];
                        const int k = link == 0 ? 0 : link - 3 * leg;
                        const M3 Rl = selm(k, Rb, R1, R2, R3);
                        V3 pl = sel3(k, mk(0, 0, 0), p1, p2, p3);
                        V3 l0 = pl + mul(Rl, mk(cp[2], cp[3], cp[4])), l1 = pl + mul(Rl, mk(cp[5], cp[6], cp[7]));
                        V3 w0 = pos + l0, w1 = pos + l1;
                        float *o = rs + RS_CAP + ci * 7;
                        o[0] = w0.x; o[1] = w0.y; o[2] = w0.z; o[3] = w1.x; o[4] = w1.y; o[5] = w1.z; o[6] = cp[8];
                        bound = fmaxf(bound, sqrtf(fmaxf(dot(l0, l0), dot(l1, l1))) + cp[8]);
                    }
                    bound = fmaxf(bound, __shfl_xor_sync(quad_mask, bound, 1));
                    bound = fmaxf(bound, __shfl_xor_sync(quad_mask, bound, 2));
                    if (leg == 0) rs[RS_BOUND] = bound * 1.0001f + 1e-5f;
                }
                __syncwarp(env_mask);
                // capsule culling: bit ci of capmask[X][Y] = capsule ci of group X reaches into the true bounding sphere of
                // group Y (+ contact offset).  A pair (X,ci,Y,cj) can only touch if both bits are set, so the narrow phase
                // below skips everything else -- exact, it never drops a pair the full enumeration would accept.
                bool live_any = false;
                for (int Y = 0; Y < Gc; Y++) {
                    if (Y == grp || grp >= Gc) continue;
                    if (!capmask[grp * G + Y]) continue;                     // out of broadphase range: the mask stays 0 (same for the whole quad)
                    const float *by_ = Y < A ? wbase + (e_loc * A + Y) * RS_SIZE : wbase + E * A * RS_SIZE + (e_loc * P + Y - A) * NS_SIZE;
                    const float *oy = by_ + (Y < A ? RS_ORIGIN : NS_ORIGIN);
                    const float reach = by_[Y < A ? RS_BOUND : NS_BOUND] + p.coff;
                    const V3 cy = mk(oy[0], oy[1], oy[2]);
                    int m = 0;
                    if (is_robot) {
                        const int *cl_ = tbl + 40 + leg * 10;
                        for (int t = 0; t < cl_[0]; t++) {
                            const int ci = cl_[1 + t];
                            const float *o = rs + RS_CAP + ci * 7;
                            V3 a0 = mk(o[0], o[1], o[2]), d = mk(o[3], o[4], o[5]) - a0, w = cy - a0;
                            float dd = dot(d, d), tpar = dd > 1e-12f ? fminf(fmaxf(dot(w, d) / dd, 0.f), 1.f) : 0.f;
                            V3 c = w - tpar * d;
                            float lim = reach + o[6];
                            if (dot(c, c) <= lim * lim) m |= 1 << ci;
                        }
                        m |= __shfl_xor_sync(quad_mask, m, 1);
                        m |= __shfl_xor_sync(quad_mask, m, 2);
                    } else {
                        V3 a0 = mk(ns[NS_CAP], ns[NS_CAP + 1], ns[NS_CAP + 2]), d = mk(ns[NS_CAP + 3], ns[NS_CAP + 4], ns[NS_CAP + 5]) - a0, w = cy - a0;
                        float dd = dot(d, d), tpar = dd > 1e-12f ? fminf(fmaxf(dot(w, d) / dd, 0.f), 1.f) : 0.f;
                        V3 c = w - tpar * d;
                        float lim = reach + ns[NS_CAP + 6];
                        if (dot(c, c) <= lim * lim) m = 1;
                    }
                    if (is_npc || leg == 0) capmask[grp * G + Y] = m;
                    live_any |= (m != 0);
                }
                if (env >= p.N) live_any = false;
                __syncwarp(env_mask);
                need_narrow = __ballot_sync(env_mask, live_any) != 0u;
            }
        }
        __syncwarp();
        SUB_MARK(11);                                    // broadphase + publish + capsule masks
        if (G > 1) {
            // ---- narrow phase, WARP-wide.  An env whose capsule masks are live borrows all 32 lanes of its warp (most substeps at most one
            // env of a warp has robots within reach of each other).  Candidates are enumerated from the masks -- capsule ci of X (bit set in
            // mask[X][Y]) against capsule cj of Y (bit set in mask[Y][X]), groups X < Y, ci then cj ascending: the oracle's loop order
            // restricted to pairs that can touch, so contact slots come out in the same canonical order (ballot compaction keeps lane =
            // item order).  Stage 1: a bounding-sphere test per capsule pair (it can only pass pairs the exact test might accept) compacts
            // the candidates into a per-env list; stage 2: the exact segment / segment test.  Hits only record a descriptor; rows are built
            // afterwards, also warp-wide.
            unsigned pendn = __ballot_sync(FULL, need_narrow && is_robot && rank_in_env == 0);
            while (pendn) {
                const int src = __ffs(pendn) - 1;
                pendn &= pendn - 1u;
                const int h_eloc = __shfl_sync(FULL, e_loc, src), h_env = __shfl_sync(FULL, env, src);
                const int *cm = reinterpret_cast<const int *>(pool + E * spair * 3 * PROWF) + h_eloc * ES_MASKSZ(G);
                int *cand = p.cand_scratch + (size_t)h_env * max_cand;
                float *h_pdesc = p.pdesc_scratch + (size_t)h_env * ((size_t)maxpair * PDESCF);
                auto gblock = [&](int Z) -> const float * { return Z < A ? wbase + (h_eloc * A + Z) * RS_SIZE : wbase + E * A * RS_SIZE + (h_eloc * P + Z - A) * NS_SIZE; };
                int ncand = 0;
                for (int X = 0; X < min(A, Gc - 1); X++)                        // pairs with a robot on the X side (robot-robot, robot-NPC)
                    for (int Y = X + 1; Y < Gc; Y++) {
                        const unsigned mXY = (unsigned)cm[X * G + Y], mYX = (unsigned)cm[Y * G + X];
                        if (!mXY || !mYX) continue;                                  // uniform over the warp
                        const float *bx_ = gblock(X), *by_ = gblock(Y);
                        const float *ox = bx_ + (X < A ? RS_ORIGIN : NS_ORIGIN), *oy = by_ + (Y < A ? RS_ORIGIN : NS_ORIGIN);
                        const float bx = X < A ? MQE_ROBOT_BOUND : p.npc_radius + p.npc_halflen, by = Y < A ? MQE_ROBOT_BOUND : p.npc_radius + p.npc_halflen;
                        const float lim = bx + by + p.coff, dx = ox[0] - oy[0], dy = ox[1] - oy[1], dz = ox[2] - oy[2];
                        if (!(dx * dx + dy * dy + dz * dz <= lim * lim)) continue;
                        const int nY = __popc(mYX), total = __popc(mXY) * nY;
                        for (int t0 = 0; t0 < total; t0 += 32) {
                            const int t = t0 + lane;
                            bool ok = false;
                            int ci = 0, cj = 0;
                            if (t < total) {
                                ci = __fns(mXY, 0, t / nY + 1); cj = __fns(mYX, 0, t % nY + 1);     // the (t / nY)-th capsule of X against the (t % nY)-th of Y
                                const float *ca = bx_ + (X < A ? RS_CAP + ci * 7 : NS_CAP), *cb = by_ + (Y < A ? RS_CAP + cj * 7 : NS_CAP);
                                const V3 a0 = mk(ca[0], ca[1], ca[2]), a1 = mk(ca[3], ca[4], ca[5]), b0 = mk(cb[0], cb[1], cb[2]), b1 = mk(cb[3], cb[4], cb[5]);
                                const V3 dA = a1 - a0, dB = b1 - b0, dc = 0.5f * (a0 + a1) - 0.5f * (b0 + b1);
                                const float reachA = 0.5f * sqrtf(dot(dA, dA)) + ca[6] + p.coff;
                                const float reach = (reachA + 0.5f * sqrtf(dot(dB, dB)) + cb[6]) * 1.0001f + 1e-6f;
                                ok = dot(dc, dc) <= reach * reach;
                            }
                            const unsigned ob = __ballot_sync(FULL, ok);
                            if (ok) cand[ncand + __popc(ob & ((1u << lane) - 1u))] = X | (ci << 8) | (Y << 16) | (cj << 24);
                            ncand += __popc(ob);
                        }
                    }
                // NPC-NPC pairs have one capsule each: a pair whose two masks are set IS the candidate (lexicographic (X, Y) order, after
                // every pair with a robot on the X side: still the canonical order)
                {
                    const int Pn = Gc - A, npp = Pn * (Pn - 1) / 2;
                    for (int t0 = 0; t0 < npp; t0 += 32) {
                        const int t = t0 + lane;
                        bool ok = false;
                        int X = 0, Y = 0;
                        if (t < npp) {
                            int i = 0, rem = t;
                            while (rem >= Pn - 1 - i) { rem -= Pn - 1 - i; i++; }
                            X = A + i; Y = X + 1 + rem;
                            ok = cm[X * G + Y] && cm[Y * G + X];
                        }
                        const unsigned ob = __ballot_sync(FULL, ok);
                        if (ok) cand[ncand + __popc(ob & ((1u << lane) - 1u))] = X | (Y << 16);
                        ncand += __popc(ob);
                    }
                }
                __syncwarp();
                // Stage 2: exact segment / segment test on the candidates
                int np = 0;
                for (int t0 = 0; t0 < ncand; t0 += 32) {
                    const int c = t0 + lane;
                    bool hit = false;
                    V3 cn = mk(0, 0, 0), cpos = mk(0, 0, 0);
                    float cgap = 0.f;
                    int X = 0, Y = 0, ci = 0, cj = 0;
                    if (c < ncand) {
                        const unsigned ent = (unsigned)cand[c];
                        X = ent & 0xff; ci = (ent >> 8) & 0xff; Y = (ent >> 16) & 0xff; cj = ent >> 24;
                        const float *bx_ = gblock(X), *by_ = gblock(Y);
                        const float *ca = bx_ + (X < A ? RS_CAP + ci * 7 : NS_CAP), *cb = by_ + (Y < A ? RS_CAP + cj * 7 : NS_CAP);
                        V3 c1, c2;
                        seg_seg(mk(ca[0], ca[1], ca[2]), mk(ca[3], ca[4], ca[5]), mk(cb[0], cb[1], cb[2]), mk(cb[3], cb[4], cb[5]), c1, c2);
                        V3 dv = c1 - c2;
                        float dist = sqrtf(dot(dv, dv));
                        cgap = dist - ca[6] - cb[6];
                        if (cgap < p.coff && dist >= 1e-9f) {
                            hit = true;
                            cn = (1.f / dist) * dv;
                            cpos = c2 + (cb[6] + 0.5f * cgap) * cn;
                        }
                    }
                    const unsigned hb = __ballot_sync(FULL, hit);
                    const int slot = np + __popc(hb & ((1u << lane) - 1u));
                    np += __popc(hb);
                    if (hit && slot < maxpair) {
                        float *ds = h_pdesc + slot * PDESCF;
                        const int rba = X < A ? X * MQE_NUM_BODIES + (int)md->caps[ci][1] : A * MQE_NUM_BODIES + (X - A);
                        const int rbb = Y < A ? Y * MQE_NUM_BODIES + (int)md->caps[cj][1] : A * MQE_NUM_BODIES + (Y - A);
                        ds[0] = cn.x; ds[1] = cn.y; ds[2] = cn.z; ds[3] = __int_as_float(rba); ds[4] = __int_as_float(rbb);
                        ds[5] = cpos.x; ds[6] = cpos.y; ds[7] = cpos.z; ds[8] = cgap;
                        ds[9] = __int_as_float(X | (ci << 8) | (Y << 16) | (cj << 24));
                    }
                }
                np = min(np, maxpair);
                if ((is_robot || is_npc) && e_loc == h_eloc) { npair = np; pairs_pending = np > 0; }
            }
        }
        __syncwarp();
        SUB_MARK(12);                                    // narrow phase
        // claim pool slots in env order (capsule path; the oriented-box path keeps its static slice)
        if (!obb && G > 1) {
            if (__ballot_sync(FULL, npair > 0)) {
                int start = 0;
                for (int e2 = 0; e2 < E; e2++) {
                    const int n = __shfl_sync(FULL, npair, e2 * 4 * A);
                    if (e2 < e_loc) start += n;
                }
                my_start = min(start, E * spair);
                my_smem = min(npair, E * spair - my_start);
            }
        }
        if (G > 1) {
            // ---- capsule-pair rows, WARP-wide: the env that found dynamic contacts borrows the lanes of the other envs of its warp (most substeps
            // at most one env of a warp has any), one (contact, direction, side) per lane and round; the two sides of a row sit in
            // neighbouring lanes and exchange their diagonal terms with one shuffle.  The operators it reads were published to shared
            // memory above; the descriptors live in the env's global scratch.
            unsigned pend = __ballot_sync(FULL, pairs_pending && is_robot && rank_in_env == 0);
            while (pend) {
                const int src = __ffs(pend) - 1;
                pend &= pend - 1u;
                const int h_eloc = __shfl_sync(FULL, e_loc, src), h_np = __shfl_sync(FULL, npair, src), h_env = __shfl_sync(FULL, env, src);
                const int h_start = __shfl_sync(FULL, my_start, src), h_smem = __shfl_sync(FULL, my_smem, src);
                const float *h_pdesc = p.pdesc_scratch + (size_t)h_env * ((size_t)maxpair * PDESCF);
                float *const h_gp = p.prow_scratch + (size_t)h_env * ((size_t)maxpair * 3 * PROWF);
                auto hrow = [&](int i) -> float * { return i < 3 * h_smem ? pool + (3 * h_start + i) * PROWF : h_gp + (i - 3 * h_smem) * PROWF; };
                const int total = 6 * h_np;
                for (int t0 = 0; t0 < total; t0 += 32) {
                    const int t = t0 + lane;
                    const bool on = t < total;
                    const int slot = t / 6, rem = t - 6 * slot, dch = rem >> 1, side = rem & 1;
                    float dd = 0.f, cgap = 0.f;
                    float *row = nullptr;
                    int X = 0, Y = 0, lega = 0, legb = 0;
                    if (on) {
                        const float *ds = h_pdesc + slot * PDESCF;
                        const V3 cn = mk(ds[0], ds[1], ds[2]), cpos = mk(ds[5], ds[6], ds[7]);
                        cgap = ds[8];
                        const unsigned ent = (unsigned)__float_as_int(ds[9]);
                        X = ent & 0xff; Y = (ent >> 16) & 0xff;
                        const int ci = (ent >> 8) & 0xff, cj = ent >> 24;
                        V3 t1, t2;
                        tangent_basis(cn, t1, t2);
                        const int la = X < A ? (int)md->caps[ci][0] : 0, lb = Y < A ? (int)md->caps[cj][0] : 0;
                        lega = la > 0 ? (la - 1) / 3 : 0; legb = lb > 0 ? (lb - 1) / 3 : 0;
                        const int Z = side ? Y : X, lz = side ? lb : la;                      // this lane's side of the row
                        const float *bz = Z < A ? wbase + (h_eloc * A + Z) * RS_SIZE : wbase + E * A * RS_SIZE + (h_eloc * P + Z - A) * NS_SIZE;
                        const float *oz = bz + (Z < A ? RS_ORIGIN : NS_ORIGIN);
                        const V3 rz = cpos - mk(oz[0], oz[1], oz[2]);
                        V3 d = dch == 0 ? cn : (dch == 1 ? t1 : t2);
                        if (side) d = -d;
                        row = hrow(3 * slot + dch);
                        dd = Z < A ? robot_side_from_smem(bz, lz, rz, d, row + 20 * side) : npc_side(p, rz, d, row + 20 * side);
                    }
                    const float dd_other = __shfl_xor_sync(FULL, dd, 1);
                    if (on && side == 0) {
                        row[18] = row[19] = 0.f;
                        row[40] = 1.f / ((dd + dd_other) + p.cfm);
                        row[41] = dch == 0 ? contact_bias(p, cgap) : 0.f;
                        row[42] = 0.f;
                        row[43] = __int_as_float(X | (lega << 4) | (Y << 8) | (legb << 12) | ((dch ? 1 : 0) << 16) | ((3 * slot) << 20));
                    } else if (on) {
                        row[38] = row[39] = 0.f;
                    }
                }
                __syncwarp();                                  // both sides of every row are in place
                for (int c = lane; c < h_np; c += 32) pair_block_coupling(hrow(3 * c), hrow(3 * c + 1), hrow(3 * c + 2));
            }
        }
        SUB_MARK(13);                                    // capsule-pair rows
        if (G > 1 && (is_robot || is_npc)) {
            if (obb) {
                // robot probes on the plank / the push box: canonical order = robot ascending, probe-table order (two-pass compaction)
                const float *nsS = wbase + E * A * RS_SIZE + (e_loc * P) * NS_SIZE;
                const V3 pivot = mk(nsS[NS_ORIGIN], nsS[NS_ORIGIN + 1], nsS[NS_ORIGIN + 2]);   // group origin: seesaw pivot / box COM
                V3 oex, oey, oez, oc;
                const float oh[3] = {p.geom[4], p.geom[5], p.geom[6]};
                const bool vcyl = seesaw && p.geom[13] > 1.5f;                    // tug disc: vertical cylinder, radius geom[4], half height geom[5]
                if (vcyl) {
                    oex = mk(1.f, 0.f, 0.f); oey = mk(0.f, 1.f, 0.f); oez = mk(0.f, 0.f, 1.f);
                    oc = pivot + mk(0.f, 0.f, p.geom[15]);
                } else if (seesaw) {
                    const float cs = nsS[NS_CAP], sn = nsS[NS_CAP + 1];
                    if (p.geom[13] > 0.5f) { oex = mk(cs, sn, 0.f); oey = mk(-sn, cs, 0.f); oez = mk(0.f, 0.f, 1.f); }
                    else { oex = mk(cs, 0.f, -sn); oey = mk(0.f, 1.f, 0.f); oez = mk(sn, 0.f, cs); }
                    oc = pivot + p.geom[3] * oex + p.geom[14] * oey + p.geom[15] * oez;
                } else {
                    oex = mk(nsS[NS_ROT], nsS[NS_ROT + 1], nsS[NS_ROT + 2]); oey = mk(nsS[NS_ROT + 3], nsS[NS_ROT + 4], nsS[NS_ROT + 5]);
                    oez = mk(nsS[NS_ROT + 6], nsS[NS_ROT + 7], nsS[NS_ROT + 8]);
                    oc = pivot;
                }
                unsigned sm = 0;
                const int *pl_ = tbl + leg * 10;
                if (is_robot && active)
                    for (int t = 0; t < pl_[0]; t++) {
                        const int pi = pl_[1 + t];
                        const float *pr = md->probes[pi];
                        const int link = (int)pr[0];
                        const int k = link == 0 ? 0 : link - 3 * leg;
                        const M3 Rl = selm(k, Rb, R1, R2, R3);
                        V3 pl = sel3(k, mk(0, 0, 0), p1, p2, p3);
                        V3 nn, pp;
                        float gg;
                        const V3 xs = pos + pl + mul(Rl, mk(pr[2], pr[3], pr[4]));
                        if (vcyl ? sphere_vcyl(p, oc, oh[0], oh[1], xs, pr[5], nn, gg, pp) : sphere_obb(p, oc, oex, oey, oez, oh, xs, pr[5], nn, gg, pp)) sm |= 1u << pi;
                    }
                unsigned sall = sm;
                if (is_robot) { sall |= __shfl_xor_sync(quad_mask, sall, 1); sall |= __shfl_xor_sync(quad_mask, sall, 2); }
                int base_slot = npair, total = 0;
                for (int X = 0; X < A; X++) {
                    int cnt = __popc(__shfl_sync(env_mask, sall, e_loc * 4 * A + 4 * X));
                    if (is_robot && X < ag) base_slot += cnt;
                    total += cnt;
                }
                while (sm) {
                    const int b = __ffs(sm) - 1;
                    sm &= sm - 1u;
                    const int slot = base_slot + __popc(sall & ((1u << b) - 1u));
                    if (slot >= maxpair) break;
                    const float *pr = md->probes[b];
                    const int link = (int)pr[0], body = (int)pr[1];
                    const int k = link == 0 ? 0 : link - 3 * leg;
                    const M3 Rl = selm(k, Rb, R1, R2, R3);
                    V3 pl = sel3(k, mk(0, 0, 0), p1, p2, p3);
                    V3 cn, cpos, t1, t2;
                    float cgap;
                    const V3 xs = pos + pl + mul(Rl, mk(pr[2], pr[3], pr[4]));
                    if (vcyl) sphere_vcyl(p, oc, oh[0], oh[1], xs, pr[5], cn, cgap, cpos);
                    else sphere_obb(p, oc, oex, oey, oez, oh, xs, pr[5], cn, cgap, cpos);
                    tangent_basis(cn, t1, t2);
                    const V3 ra = cpos - pos, rb_ = cpos - pivot;
                    float *cmeta = pdesc + slot * PDESCF;
                    cmeta[0] = cn.x; cmeta[1] = cn.y; cmeta[2] = cn.z;
                    cmeta[3] = __int_as_float(ag * MQE_NUM_BODIES + body); cmeta[4] = __int_as_float(A * MQE_NUM_BODIES);
#pragma unroll 1
                    for (int dch = 0; dch < 3; dch++) {
                        V3 d = dch == 0 ? cn : (dch == 1 ? t1 : t2);
                        float *row = prow(3 * slot + dch);
                        float dd = robot_side_regs(k, ra, d, a1, a2, p1, p2, p3, Gm, Sinv, Hinv, row);
                        dd += npc_side(p, rb_, -d, row + 20);
                        row[18] = row[19] = row[38] = row[39] = 0.f;
                        row[40] = 1.f / (dd + p.cfm);
                        row[41] = dch == 0 ? contact_bias(p, cgap) : 0.f;
                        row[42] = 0.f;
                        row[43] = __int_as_float(ag | (leg << 4) | (A << 8) | (0 << 12) | ((dch ? 1 : 0) << 16) | ((3 * slot) << 20));
                    }
                }
                npair = min(npair + total, maxpair);
            }
            if (obb && npair > 0) {                            // uniform over the env's lanes (capsule contacts were done in the warp-wide pass; recomputing them is idempotent)
                __syncwarp(env_mask);                              // rows of a contact were written by the lane that owns the probe
                for (int c = rank_in_env; c < npair; c += lanes_per_env) { float *R = prow(3 * c); pair_block_coupling(R, R + PROWF, R + 2 * PROWF); }   // a contact's rows are contiguous in either store
            }
            if (rank_in_env == 0 && env < p.N) stat_pair += npair;
        }
        __syncwarp();
        if (active && (leg == 0 || is_npc)) stat_rows = max(stat_rows, nlim + nrows - crow0);

        PHASE_MARK(3);
        CTA_ALIGN(2);
        // ================================================================ P4: projected Gauss-Seidel
        {
            const int quad_base = lane & ~3;
            // every lane of a quad keeps all four legs' u (12 floats): a row's leg term J_l . u_leg is then local arithmetic
            // instead of a shuffle on the critical path of the sweep
            float ua[4][3];
#pragma unroll
            for (int L = 0; L < 4; L++)
#pragma unroll
                for (int k = 0; k < 3; k++) {           // full-mask shuffle executed by every lane (other lanes read a neighbour's zeros): measured
                    const float t = __shfl_sync(FULL, u[k], quad_base + L);      // 7x faster here than the quad-mask shuffle under `is_robot ?`
                    ua[L][k] = is_robot ? t : 0.f;
                }
            // row i lives at rowS + i*ROWF (shared memory; generic pointer) or, past SROWS (robots only), at rowG + i*ROWF
            float *const rowS = is_robot ? rs + RS_ROWS : ns + NS_ROWS;
            float *const rowG = grows - SROWS * ROWF;
            PHASE_MARK(10);
            // warp-uniform: almost no warp holds a pair contact, and a uniform branch keeps the compiler from hoisting the pair sweep's
            // leg-velocity selects (36 instructions) into every sweep of every warp
            const bool warp_pairs = __any_sync(FULL, npair > 0);
            for (int it = 0; it < p.iters; it++) {
                // ---- joint-limit rows: single rows, kind 0 (lambda >= 0) ----
                for (int i = 0; i < nlim; i++) {
                    float *row = rowS + i * ROWF;
                    const float4 r0 = *reinterpret_cast<const float4 *>(row), r1 = *reinterpret_cast<const float4 *>(row + 4), r2 = *reinterpret_cast<const float4 *>(row + 8);
                    const float4 r3 = *reinterpret_cast<const float4 *>(row + 12), r4 = *reinterpret_cast<const float4 *>(row + 16), r5 = *reinterpret_cast<const float4 *>(row + 20);
                    const int rleg = __float_as_int(r5.y) & 15;
                    const float pb = (fmaf(r0.x, vb[0], r0.y * vb[1]) + fmaf(r0.z, vb[2], r0.w * vb[3])) + fmaf(r1.x, vb[4], r1.y * vb[5]);
                    float pl = 0.f;
#pragma unroll
                    for (int L = 0; L < 4; L++) {
                        const float t = fmaf(r1.z, ua[L][0], fmaf(r1.w, ua[L][1], r2.x * ua[L][2]));
                        pl = rleg == L ? t : pl;
                    }
                    const float urel = (r4.w + pl) + pb;            // bias + J w
                    const float lam_old = r5.x, lam = fmaxf(lam_old - urel * r4.z, 0.f), dl = lam - lam_old;
                    row[20] = lam;                               // all four lanes of the quad store the SAME value (benign same-value race)
                    vb[0] += r2.y * dl; vb[1] += r2.z * dl; vb[2] += r2.w * dl; vb[3] += r3.x * dl; vb[4] += r3.y * dl; vb[5] += r3.z * dl;
#pragma unroll
                    for (int L = 0; L < 4; L++) {                       // branch-free: only the row's leg sees a non-zero dl
                        const float dlL = rleg == L ? dl : 0.f;
                        ua[L][0] = fmaf(r3.w, dlL, ua[L][0]); ua[L][1] = fmaf(r4.x, dlL, ua[L][1]); ua[L][2] = fmaf(r4.y, dlL, ua[L][2]);
                    }
                }
                // ---- contact blocks: normal + two friction rows solved together (see local_block_coupling) ----
                for (int c0 = crow0; c0 < nrows; c0 += 3) {
                    float *ra = (c0 < SROWS ? rowS : rowG) + c0 * ROWF, *rb = ra + ROWF, *rc = ra + 2 * ROWF;     // block-aligned boundary: see crow0
                    // J parts of the three rows: Jb0..5 Jl0..2
                    const float4 a0 = *reinterpret_cast<const float4 *>(ra), a1 = *reinterpret_cast<const float4 *>(ra + 4), a2 = *reinterpret_cast<const float4 *>(ra + 8);
                    const float4 b0 = *reinterpret_cast<const float4 *>(rb), b1 = *reinterpret_cast<const float4 *>(rb + 4), b2 = *reinterpret_cast<const float4 *>(rb + 8);
                    const float4 c0v = *reinterpret_cast<const float4 *>(rc), c1 = *reinterpret_cast<const float4 *>(rc + 4), c2 = *reinterpret_cast<const float4 *>(rc + 8);
                    const float4 a4 = *reinterpret_cast<const float4 *>(ra + 16), a5 = *reinterpret_cast<const float4 *>(ra + 20);   // Yl1 Yl2 dinv bias | lam meta - -
                    const float4 b4 = *reinterpret_cast<const float4 *>(rb + 16), b5 = *reinterpret_cast<const float4 *>(rb + 20);   //                 | lam meta K10 -
                    const float4 c4 = *reinterpret_cast<const float4 *>(rc + 16), c5 = *reinterpret_cast<const float4 *>(rc + 20);   //                 | lam meta K20 K21
                    const int rleg = __float_as_int(a5.y) & 15;
                    // the block's leg velocity, selected once
                    const float ul0 = sel1(rleg, ua[0][0], ua[1][0], ua[2][0], ua[3][0]), ul1 = sel1(rleg, ua[0][1], ua[1][1], ua[2][1], ua[3][1]),
                                ul2 = sel1(rleg, ua[0][2], ua[1][2], ua[2][2], ua[3][2]);
#define BLOCK_DOT(q0, q1, q2)                                                                                              \
    (((fmaf(q0.x, vb[0], q0.y * vb[1]) + fmaf(q0.z, vb[2], q0.w * vb[3])) + fmaf(q1.x, vb[4], q1.y * vb[5])) +            \
     fmaf(q1.z, ul0, fmaf(q1.w, ul1, q2.x * ul2)))
                    const float j0 = BLOCK_DOT(a0, a1, a2), j1 = BLOCK_DOT(b0, b1, b2), j2 = BLOCK_DOT(c0v, c1, c2);
                    // sequential solve inside the block (row order n, t1, t2 as in the oracle)
                    const float l0o = a5.x, l1o = b5.x, l2o = c5.x;
                    const float l0 = fmaxf(l0o - (a4.w + j0) * a4.z, 0.f), d0 = l0 - l0o;
                    const float lim = mu_e * l0;
                    const float l1 = fminf(fmaxf(l1o - (b4.w + fmaf(b5.z, d0, j1)) * b4.z, -lim), lim), d1 = l1 - l1o;
                    const float l2 = fminf(fmaxf(l2o - (c4.w + fmaf(c5.w, d1, fmaf(c5.z, d0, j2))) * c4.z, -lim), lim), d2 = l2 - l2o;
                    ra[20] = l0; rb[20] = l1; rc[20] = l2;         // all four lanes of the quad store the SAME values
                    // velocity update: w += Y0 d0 + Y1 d1 + Y2 d2
                    const float4 a3 = *reinterpret_cast<const float4 *>(ra + 12), b3 = *reinterpret_cast<const float4 *>(rb + 12), c3 = *reinterpret_cast<const float4 *>(rc + 12);
                    vb[0] = fmaf(c2.y, d2, fmaf(b2.y, d1, fmaf(a2.y, d0, vb[0]))); vb[1] = fmaf(c2.z, d2, fmaf(b2.z, d1, fmaf(a2.z, d0, vb[1])));
                    vb[2] = fmaf(c2.w, d2, fmaf(b2.w, d1, fmaf(a2.w, d0, vb[2]))); vb[3] = fmaf(c3.x, d2, fmaf(b3.x, d1, fmaf(a3.x, d0, vb[3])));
                    vb[4] = fmaf(c3.y, d2, fmaf(b3.y, d1, fmaf(a3.y, d0, vb[4]))); vb[5] = fmaf(c3.z, d2, fmaf(b3.z, d1, fmaf(a3.z, d0, vb[5])));
                    const float du0 = fmaf(c3.w, d2, fmaf(b3.w, d1, a3.w * d0)), du1 = fmaf(c4.x, d2, fmaf(b4.x, d1, a4.x * d0)),
                                du2 = fmaf(c4.y, d2, fmaf(b4.y, d1, a4.y * d0));
#pragma unroll
                    for (int L = 0; L < 4; L++) {                       // branch-free: only the block's leg moves (1 * du + ua is exact)
                        const float mine = rleg == L ? 1.f : 0.f;
                        ua[L][0] = fmaf(mine, du0, ua[L][0]); ua[L][1] = fmaf(mine, du1, ua[L][1]); ua[L][2] = fmaf(mine, du2, ua[L][2]);
                    }
                }
                if (p.trace && lane == 0) t_sub = (unsigned)clock();
                if (warp_pairs && npair > 0) {   // npair is uniform over the env's lanes
#pragma unroll
                    for (int k = 0; k < 3; k++) u[k] = leg == 0 ? ua[0][k] : (leg == 1 ? ua[1][k] : (leg == 2 ? ua[2][k] : ua[3][k]));
                    // Pair contacts, one block per contact.  A lane reads only its own group's side of the three rows; the lane that
                    // owns the contact's leg on each side forms that side's three J.w, ONE round of shuffles combines the sides, every
                    // lane of the env runs the same 3-row solve, and the multipliers are carried to the next sweep through the rows.
                    for (int c = 0; c < npair; c++) {
                        float *R0 = prow(3 * c), *R1 = R0 + PROWF, *R2 = R0 + 2 * PROWF;      // contiguous in either store (the shared / global boundary is contact-aligned)
                        const float4 t0 = *reinterpret_cast<const float4 *>(R0 + 40), t1 = *reinterpret_cast<const float4 *>(R1 + 40),
                                     t2 = *reinterpret_cast<const float4 *>(R2 + 40);                    // dinv, bias, lambda, meta
                        const float k10 = R1[18], k20 = R2[18], k21 = R2[19];
                        const int meta = __float_as_int(t0.w);
                        const int ga = meta & 15, la = (meta >> 4) & 15, gb = (meta >> 8) & 15, lb = (meta >> 12) & 15;
                        const bool inA = grp == ga, inB = grp == gb, in = inA || inB;
                        const int so = inB ? 20 : 0;
                        float jw0 = 0.f, jw1 = 0.f, jw2 = 0.f;
                        if (in) {
#define SIDE_DOT(R)                                                                                                         \
    {                                                                                                                       \
        const float4 q0 = *reinterpret_cast<const float4 *>((R) + so), q1 = *reinterpret_cast<const float4 *>((R) + so + 4); \
        const float q2x = (R)[so + 8];                                                                                      \
        jw_ = ((fmaf(q0.x, vb[0], q0.y * vb[1]) + fmaf(q0.z, vb[2], q0.w * vb[3])) + fmaf(q1.x, vb[4], q1.y * vb[5])) +     \
              fmaf(q1.z, u[0], fmaf(q1.w, u[1], q2x * u[2]));                                                               \
    }
                            float jw_;
                            SIDE_DOT(R0) jw0 = jw_;
                            SIDE_DOT(R1) jw1 = jw_;
                            SIDE_DOT(R2) jw2 = jw_;
                        }
                        const int srcA = ga < A ? e_loc * 4 * A + 4 * ga + la : nrl + e_loc * P + (ga - A);
                        const int srcB = gb < A ? e_loc * 4 * A + 4 * gb + lb : nrl + e_loc * P + (gb - A);
                        // (exchanging the six J.w through shared memory instead measured the same, 207 k vs 209 k cycles for the slowest warp:
                        // a pair block costs what its ~160 instructions cost, ~6 cycles each in a warp that runs alone)
                        const float s0 = __shfl_sync(env_mask, jw0, srcA) + __shfl_sync(env_mask, jw0, srcB);
                        const float s1 = __shfl_sync(env_mask, jw1, srcA) + __shfl_sync(env_mask, jw1, srcB);
                        const float s2 = __shfl_sync(env_mask, jw2, srcA) + __shfl_sync(env_mask, jw2, srcB);
                        const float l0 = fmaxf(t0.z - (t0.y + s0) * t0.x, 0.f), d0 = l0 - t0.z;
                        const float lim = mu_e * l0;
                        const float l1 = fminf(fmaxf(t1.z - (t1.y + fmaf(k10, d0, s1)) * t1.x, -lim), lim), d1 = l1 - t1.z;
                        const float l2 = fminf(fmaxf(t2.z - (t2.y + fmaf(k21, d1, fmaf(k20, d0, s2))) * t2.x, -lim), lim), d2 = l2 - t2.z;
                        if (rank_in_env == 0) { R0[42] = l0; R1[42] = l1; R2[42] = l2; }
                        if (in) {
                            // Y parts: Yb0..5 at [9..14], Yl0..2 at [15..17] of this lane's side
                            const float dll0 = (is_robot && leg == (inB ? lb : la)) ? 1.f : 0.f;
#define SIDE_UPD(R, d)                                                                                                      \
    {                                                                                                                       \
        const float4 y2 = *reinterpret_cast<const float4 *>((R) + so + 8), y3 = *reinterpret_cast<const float4 *>((R) + so + 12); \
        const float4 y4 = *reinterpret_cast<const float4 *>((R) + so + 16);                                                 \
        vb[0] = fmaf(y2.y, d, vb[0]); vb[1] = fmaf(y2.z, d, vb[1]); vb[2] = fmaf(y2.w, d, vb[2]);                           \
        vb[3] = fmaf(y3.x, d, vb[3]); vb[4] = fmaf(y3.y, d, vb[4]); vb[5] = fmaf(y3.z, d, vb[5]);                           \
        const float dl_ = dll0 * (d);                                                                                       \
        u[0] = fmaf(y3.w, dl_, u[0]); u[1] = fmaf(y4.x, dl_, u[1]); u[2] = fmaf(y4.y, dl_, u[2]);                           \
    }
                            SIDE_UPD(R0, d0) SIDE_UPD(R1, d1) SIDE_UPD(R2, d2)
                        }
                    }
                    __syncwarp(env_mask);
                    if (is_robot) {
#pragma unroll
                        for (int L = 0; L < 4; L++)
#pragma unroll
                            for (int k = 0; k < 3; k++) ua[L][k] = __shfl_sync(quad_mask, u[k], quad_base + L);
                    }
                }
                if (p.trace) { __syncwarp(); SUB_MARK(14); }     // pair sweeps (lane 0 waits here for the envs of its warp that have pairs)
            }
#pragma unroll
            for (int k = 0; k < 3; k++) u[k] = leg == 0 ? ua[0][k] : (leg == 1 ? ua[1][k] : (leg == 2 ? ua[2][k] : ua[3][k]));
        }

        PHASE_MARK(4);
        CTA_ALIGN(3);
        // ================================================================ P5: contact force report (last substep), integrate
        if (last) {
            float idt = 1.f / p.dt;
            __syncwarp();                                        // multipliers written by other lanes of the quad are read below
            if (is_robot && active) {
                int ncon = (nrows - crow0) / 3;
                for (int c = leg; c < ncon; c += 4) {
                    const float *cmeta = rs + RS_CMETA + c * 4;
                    V3 n = mk(cmeta[0], cmeta[1], cmeta[2]), t1, t2;
                    tangent_basis(n, t1, t2);
                    int body = __float_as_int(cmeta[3]);
                    const float *Rf = lrow(crow0 + 3 * c);
                    V3 f = (Rf[20] * idt) * n + (Rf[ROWF + 20] * idt) * t1 + (Rf[2 * ROWF + 20] * idt) * t2;
                    atomicAdd(rs + RS_FORCE + body * 3, f.x); atomicAdd(rs + RS_FORCE + body * 3 + 1, f.y); atomicAdd(rs + RS_FORCE + body * 3 + 2, f.z);
                }
            } else if (is_npc && active) {
                for (int c = 0; c < nrows / 3; c++) {
                    const float *cmeta = ns + NS_CMETA + c * 4;
                    V3 n = mk(cmeta[0], cmeta[1], cmeta[2]), t1, t2;
                    tangent_basis(n, t1, t2);
                    const float *row = ns + NS_ROWS + 3 * c * ROWF;
                    V3 f = (row[20] * idt) * n + (row[ROWF + 20] * idt) * t1 + (row[2 * ROWF + 20] * idt) * t2;
                    ns[NS_FORCE] += f.x; ns[NS_FORCE + 1] += f.y; ns[NS_FORCE + 2] += f.z;
                }
            }
            __syncwarp();
            if (npair > 0 && active) {
                for (int c = rank_in_env; c < npair; c += lanes_per_env) {
                    const float *cmeta = pdesc + c * PDESCF;
                    V3 n = mk(cmeta[0], cmeta[1], cmeta[2]), t1, t2;
                    tangent_basis(n, t1, t2);
                    int rba = __float_as_int(cmeta[3]), rbb = __float_as_int(cmeta[4]);
                    const float *Rf = prow(3 * c);
                    V3 f = (Rf[42] * idt) * n + (Rf[PROWF + 42] * idt) * t1 + (Rf[2 * PROWF + 42] * idt) * t2;
                    for (int side = 0; side < 2; side++) {
                        int rb = side ? rbb : rba;
                        float sg = side ? -1.f : 1.f;
                        float *dst = rb < A * MQE_NUM_BODIES
                                         ? wbase + (e_loc * A + rb / MQE_NUM_BODIES) * RS_SIZE + RS_FORCE + (rb % MQE_NUM_BODIES) * 3
                                         : wbase + E * A * RS_SIZE + (e_loc * P + rb - A * MQE_NUM_BODIES) * NS_SIZE + NS_FORCE;
                        atomicAdd(dst, sg * f.x); atomicAdd(dst + 1, sg * f.y); atomicAdd(dst + 2, sg * f.z);
                    }
                }
            }
            __syncwarp();
            if (active) {
                float *cf = p.contact + (size_t)env * p.NB * 3;
                if (is_robot) for (int i = leg; i < 51; i += 4) cf[ag * 51 + i] = rs[RS_FORCE + i];
                else for (int i = 0; i < 3; i++) cf[(A * MQE_NUM_BODIES + pn) * 3 + i] = ns[NS_FORCE + i];
            }
        }
        if (is_robot) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float gt = 0.f;
#pragma unroll
                for (int i = 0; i < 6; i++) gt += Gm[i * 3 + k] * vb[i];
                float v = u[k] - gt, lim = md->qd_limit[3 * leg + k];
                v = fminf(fmaxf(v, -lim), lim);
                qd[k] = v;
                q[k] += p.dt * v;
            }
            if (p.sub_tau && active) {      // post_decimation_step (legged_robot.py:112-115): torques applied, then the refreshed dof state
                const size_t o = ((size_t)env * nsub + sub) * (12 * A) + 12 * ag + 3 * leg;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const int j = 3 * leg + k;
                    const float mid = 0.5f * (md->q_lower[j] + md->q_upper[j]), half = 0.5f * (md->q_upper[j] - md->q_lower[j]) * p.soft_limit;
                    p.sub_tau[o + k] = tau[k]; p.sub_qd[o + k] = qd[k];
                    p.sub_exceed[o + k] = (unsigned char)((q[k] < mid - half) | (q[k] > mid + half));
                }
            }
        }
        if (is_npc && seesaw) {
            const float lim = p.geom[12];                                    // URDF joint velocity limit
            qd[0] = fminf(fmaxf(p.geom[13] > 1.5f ? vb[4] : (p.geom[13] > 0.5f ? vb[2] : vb[1]), -lim), lim);
            q[0] += p.dt * qd[0];
        } else if (is_robot || is_npc) {
            wang = mk(vb[0], vb[1], vb[2]); vlin = mk(vb[3], vb[4], vb[5]);
            pos = pos + p.dt * vlin;
            float wn = sqrtf(dot(wang, wang)), th = wn * p.dt, dx, dy, dz, dw;
            if (th < 1e-8f) { dx = wang.x * p.dt * 0.5f; dy = wang.y * p.dt * 0.5f; dz = wang.z * p.dt * 0.5f; dw = 1.f; }
            else { float sn, cs; sincosf(th * 0.5f, &sn, &cs); float s = sn / wn; dx = wang.x * s; dy = wang.y * s; dz = wang.z * s; dw = cs; }
            float nx = dw * qx + dx * qw + dy * qz - dz * qy;
            float ny = dw * qy - dx * qz + dy * qw + dz * qx;
            float nz = dw * qz + dx * qy - dy * qx + dz * qw;
            float nw = dw * qw - dx * qx - dy * qy - dz * qz;
            float inv = 1.f / sqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
            qx = nx * inv; qy = ny * inv; qz = nz * inv; qw = nw * inv;
        }
        __syncwarp();
        PHASE_MARK(5);
    }

    // ---- write back ----
    if (active) {
        if (is_npc && seesaw) {
            float *d = p.dof + ((size_t)env * (12 * A + p.D) + 12 * A) * 2;
            d[0] = q[0]; d[1] = qd[0];
        } else if (is_npc || leg == 0) {
            float *r = p.root + ((size_t)env * GA + grp) * 13;
            r[0] = pos.x; r[1] = pos.y; r[2] = pos.z; r[3] = qx; r[4] = qy; r[5] = qz; r[6] = qw;
            r[7] = vlin.x; r[8] = vlin.y; r[9] = vlin.z; r[10] = wang.x; r[11] = wang.y; r[12] = wang.z;
        }
        if (is_robot) {
            float *d = p.dof + ((size_t)env * (12 * A + p.D) + 12 * ag + 3 * leg) * 2;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                d[2 * k] = q[k]; d[2 * k + 1] = qd[k];
                int j = m_idx * 12 + 3 * leg + k;
                p.err1[j] = e1[k]; p.err2[j] = e2[k]; p.vel1[j] = v1[k]; p.vel2[j] = v2[k];
                p.torques[j] = tau[k];
            }
        }
        if (is_npc || leg == 0) {
            atomicAdd(p.stats + 0, stat_local); atomicAdd(p.stats + 1, stat_lim);
            atomicMax(p.stats + 3, stat_rows);
        }
        if (rank_in_env == 0) atomicAdd(p.stats + 2, stat_pair);
    }
    if (p.fuse_post) {                                                     // the env's post-physics step, by the warp that integrated it
        PostLaneState st;
        st.pos = pos; st.vlin = vlin; st.wang = wang; st.qx = qx; st.qy = qy; st.qz = qz; st.qw = qw;
#pragma unroll
        for (int k = 0; k < 3; k++) { st.q[k] = q[k]; st.qd[k] = qd[k]; st.act[k] = act[k]; }
        substeps_fused_post(p, md, rs, st, env, ag, leg, e_loc, lane, active, is_robot, quad_mask);
        PHASE_MARK(15);
    }
    if (p.lag_ring && threadIdx.x == 0) {                                  // the last CTA to finish advances the call counter
        __threadfence();
        if (atomicAdd(&p.ctr[4], 1) == (int)gridDim.x - 1) { p.ctr[4] = 0; p.ctr[3] = (p.ctr[3] + nsub) % p.lag_n; }   // kept reduced: never overflows
    }
    // per-warp trace (MQE_BUF_WARP_TRACE): start / end on the global timer [ns], pair contacts and widest row count of the warp
    {
        int wp = stat_pair, wr = stat_rows;
#pragma unroll
        for (int o = 16; o; o >>= 1) { wp += __shfl_xor_sync(FULL, wp, o); wr = max(wr, __shfl_xor_sync(FULL, wr, o)); }
        if (lane == 0) {
            long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            PHASE_MARK(7);
            tr[0] = t_start; tr[1] = t_end; tr[2] = wp; tr[3] = wr;
            if (p.task_cost) p.task_cost[(p.ctr[1] & 1) * ntasks + first_env / E] = (int)min(t_end - t_start, (long long)0x7fffffff);
        }
    }
}

// stand-alone actuator network (parity tests against unitree_go1.pt): x [rows][6] -> torque [rows], unclipped
__global__ void k_actuator(const float *__restrict__ aw, const float *__restrict__ x, int rows, float *__restrict__ out) {
    __shared__ float w[ACTW_PLAIN];
    for (int i = threadIdx.x; i < 1313; i += blockDim.x) w[i] = aw[i];
    __syncthreads();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float xi[6], h0[32];
    for (int i = 0; i < 6; i++) xi[i] = x[(size_t)r * 6 + i];
    for (int o = 0; o < 32; o++) {
        float s = w[192 + o];
        for (int i = 0; i < 6; i++) s += w[o * 6 + i] * xi[i];
        h0[o] = softsign(s);
    }
    float y = w[1312];
    for (int o = 0; o < 32; o++) {
        float s = w[1248 + o];
        for (int i = 0; i < 32; i++) s += w[224 + i * 32 + o] * h0[i];      // W1 is stored transposed ([in][out])
        y += w[1280 + o] * softsign(s);
    }
    out[r] = y;
}

// host-side launchers (api.cu)
// Launch plan.  Shared memory per warp sets how many warps an SM holds (registers allow 8); the kernel is latency-bound, so
// residency is throughput.  spair (pair contacts whose rows stay in shared memory) is the largest value that does not cost a
// resident warp; warps per CTA are then balanced so the grid fills whole waves of one CTA per SM (C2: 1024 warps of 4 envs
// = 147 CTAs x 7 warps = one wave on 148 SMs).
struct SubstepPlan { int warps, spair, grid; size_t smem; };
static SubstepPlan substeps_plan(int N, int A, int Pd, int E, int maxpair) {
    static int sm_count[64] = {0};                    // per device ordinal (one process may own engines on several devices)
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
        if (sm_count[dev] == 0) {
            int n = 0;
            if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
            sm_count[dev] = n;
        }
        sms = sm_count[dev];
    }
    const size_t budget = 227 * 1024, hdr = (size_t)physics_cta_header_floats() * 4;
    auto fit = [&](int sp) { int w = (int)((budget - hdr) / ((size_t)physics_warp_smem_floats(A, Pd, E, sp, maxpair) * 4)); return w > 8 ? 8 : w; };
    SubstepPlan pl;
    const int spmin = maxpair < 2 ? maxpair : 2;
    pl.spair = maxpair;
    while (pl.spair > spmin && fit(pl.spair) < fit(spmin)) pl.spair--;
    const int wmax = fit(pl.spair);
    pl.warps = wmax;
    if (wmax >= 1) {
        const int tasks = (N + E - 1) / E, waves = (tasks + sms * wmax - 1) / (sms * wmax);
        int w = (tasks + sms * waves - 1) / (sms * waves);
        pl.warps = w < 1 ? 1 : (w > wmax ? wmax : w);
        { static const char *e = getenv("MQE_SUBSTEP_WARPS"); if (e && atoi(e) >= 1 && atoi(e) <= wmax) pl.warps = atoi(e); }   // experiment: CTA granularity
        pl.grid = (tasks + pl.warps - 1) / pl.warps;
    } else pl.grid = 0;
    pl.smem = (size_t)(physics_cta_header_floats() + (pl.warps < 1 ? 1 : pl.warps) * physics_warp_smem_floats(A, Pd, E, pl.spair, maxpair)) * sizeof(float);
    return pl;
}
extern "C" size_t mqe_substeps_smem_bytes(int N, int A, int Pd, int E, int maxpair) { return substeps_plan(N, A, Pd, E, maxpair).smem; }
extern "C" size_t mqe_substeps_row_scratch_floats(int N, int A) { return (size_t)N * A * (MQE_MAX_ROWS - SROWS) * ROWF; }
extern "C" size_t mqe_substeps_prow_scratch_floats(int N, int maxpair) { return (size_t)N * maxpair * 3 * PROWF; }
extern "C" size_t mqe_substeps_pdesc_scratch_floats(int N, int maxpair) { return (size_t)N * maxpair * PDESCF; }
extern "C" cudaError_t mqe_launch_actuator(const float *act_w, const float *x, int rows, float *out, cudaStream_t st) {
    k_actuator<<<(rows + 127) / 128, 128, 0, st>>>(act_w, x, rows, out);
    return cudaGetLastError();
}
// cudaFuncSetAttribute applies to the CURRENT device: called once per engine at creation (api.cu), for that engine's device
extern "C" cudaError_t mqe_substeps_configure(const DevParams &p, int maxpair) {
    const SubstepPlan pl = substeps_plan(p.N, p.A, p.Pd, p.E, maxpair);
    if (pl.warps < 1) return cudaErrorInvalidConfiguration;
    static size_t configured[64] = {0};                // largest size asked for so far, per device ordinal
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || pl.smem > configured[dev]) {
        e = cudaFuncSetAttribute(k_substeps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = pl.smem;
    }
    return cudaSuccess;
}
// Task order for the next k_substeps launch: env groups ordered by the duration of their warp in the previous launch, longest first.
// One CTA, counting sort over 1024 duration buckets in shared memory (a few microseconds for 8192 groups; a bitonic sort of the same keys took
// 15 us for 1024 groups and 45 us for 4096, which showed in front of the policy tail).  Groups of one bucket come out in arbitrary order: the
// order never changes results, only which warp works on which env group.  Runs on a side stream beside the policy kernels.
__global__ void __launch_bounds__(1024) k_balance_tasks(const int *__restrict__ cost2, const int *__restrict__ ctr, int ntasks, int *__restrict__ order, int spread_warps) {
    __shared__ int hist[1024], offs[1024], s_max;
    const int *cost = cost2 + ((ctr[1] + 1) & 1) * ntasks;     // the half the PREVIOUS step's launch wrote (its bookkeeping has advanced ctr[1] since)
    const int t = threadIdx.x;
    hist[t] = 0;
    if (t == 0) s_max = 1;
    __syncthreads();
    int m = 1;
    for (int i = t; i < ntasks; i += 1024) m = max(m, cost[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((t & 31) == 0) atomicMax(&s_max, m);
    __syncthreads();
    const float scale = 1023.f / (float)s_max;
    for (int i = t; i < ntasks; i += 1024) atomicAdd(&hist[1023 - min(1023, (int)((float)max(cost[i], 0) * scale))], 1);     // bucket 0 = longest
    __syncthreads();
    int v = hist[t];                                          // exclusive scan of the 1024 buckets (Hillis-Steele, inclusive, then shifted)
    offs[t] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const int add = t >= d ? offs[t - d] : 0;
        __syncthreads();
        offs[t] += add;
        __syncthreads();
    }
    const int excl = offs[t] - v;
    __syncthreads();
    offs[t] = excl;
    __syncthreads();
    const int W = spread_warps, C = W > 0 ? ntasks / W : 0;   // spread: C full CTAs of W warps
    for (int i = t; i < ntasks; i += 1024) {
        const int b = 1023 - min(1023, (int)((float)max(cost[i], 0) * scale));
        const int r = atomicAdd(&offs[b], 1);                 // rank of group i, longest first
        if (W <= 0 || r >= C * W) order[r] = i;               // grouped: CTA c takes ranks [c * warps, ..): similar loads together, longest CTAs first
        else {
            // spread (one-wave grids): every full CTA takes ONE group of each tier of the ranking (tier k = ranks [k * C, (k + 1) * C)), tiers dealt
            // in alternating direction, so the slowest warps sit on different SMs, each beside the fastest ones
            const int k = r / C, pos = r % C, c = (k & 1) ? C - 1 - pos : pos;
            order[c * W + k] = i;
        }
    }
}
extern "C" cudaError_t mqe_launch_balance_tasks(const DevParams &p, int *order, int spread, int maxpair, cudaStream_t st) {
    const int ntasks = (p.N + p.E - 1) / p.E;
    const int spread_warps = spread ? substeps_plan(p.N, p.A, p.Pd, p.E, maxpair).warps : 0;
    // greatest priority, like every kernel the step waits for
    return launch_pdl_if(false, k_balance_tasks, dim3(1), dim3(1024), 0, st, (const int *)p.task_cost, (const int *)p.ctr, ntasks, order, spread_warps);
}
extern "C" cudaError_t mqe_launch_substeps(const DevParams &p, int nsub, int maxpair, cudaStream_t st) {
    const SubstepPlan pl = substeps_plan(p.N, p.A, p.Pd, p.E, maxpair);
    if (pl.warps < 1) return cudaErrorInvalidConfiguration;
    return launch_heavy(k_substeps, dim3(pl.grid), dim3(pl.warps * 32), pl.smem, st, p, nsub, maxpair, pl.spair, p.max_cand);
}
