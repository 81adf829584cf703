/*
 * mqe_oracle.c -- CPU restatement of the Go1.step() hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product (libmqe_b200.so) never links, imports or falls back to it.
 *
 * PARITY STATUS
 *   - bookkeeping (go1.py, legged_robot.py, legged_robot_field.py, the npc modules): restated line by line from the
 *     reference sources cited at each function; PINNED by golden vectors that the reference's own methods produced in
 *     this container (tests/golden/bookkeeping_*.npz via tools/gen_bookkeeping_golden.py, replayed by
 *     tests/test_oracle_bookkeeping.py), by the TorchScript known-answer vectors (tests/golden/mlp_kat.npz) for the two
 *     networks and by tests/golden/terrain_*.npz for BarrierTrack.
 *   - rigid-body physics (gym.simulate, go1.py:52-56): the arithmetic lives in NVIDIA Isaac Gym Preview 4
 *     (closed PhysX binary, un-vendored, unversioned in setup.py:11).  It cannot be run or inspected here
 *     and the reference ships no tests or golden trajectories => PARITY UNPINNED against PhysX.  What is
 *     restated instead is this project's own algorithm, written here in the most
 *     literal dense form (dense CRBA mass matrix, dense Cholesky inverse, dense Jacobian rows), while the
 *     CUDA kernels use a structured/sparse formulation; agreement of the two is the parity check, physics
 *     invariants (tests/test_oracle_physics.py) are the sanity check.
 *
 * Build: see oracle/Makefile (REAL=double -> liboracle_f64.so, REAL=float -> liboracle_f32.so).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mqe_b200.h"

#ifndef REAL
#define REAL double
#endif
typedef REAL real;

#define NV 18                 /* generalized velocity of one robot: ang3, lin3, 12 joints */
#define MAX_LOCAL_CONTACTS 8  /* per robot / npc, world contacts                          */
#define MAX_LIMIT_ROWS 4
#define MAX_PAIR_CONTACTS 16  /* per env, dynamic-vs-dynamic                              */
#define ROBOT_BOUND 0.60      /* broadphase radius around the base origin [m]             */
#define PI_R ((real)3.14159265358979323846)

#if defined(_OPENMP)
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------ small math */
static inline void v3set(real *o, real x, real y, real z) { o[0] = x; o[1] = y; o[2] = z; }
static inline void v3cpy(real *o, const real *a) { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
static inline void v3add(real *o, const real *a, const real *b) { o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; }
static inline void v3sub(real *o, const real *a, const real *b) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline void v3axpy(real *o, real s, const real *a) { o[0] += s * a[0]; o[1] += s * a[1]; o[2] += s * a[2]; }
static inline real v3dot(const real *a, const real *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void v3cross(real *o, const real *a, const real *b) {
    real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline void m3mulv(real *o, const real *R, const real *v) {
    real x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    real y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    real z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline void m3mul(real *o, const real *A, const real *B) {
    real t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
    memcpy(o, t, sizeof t);
}
static void quat_to_mat(real *R, const real *q) { /* xyzw */
    real x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
/* isaacgym.torch_utils.quat_rotate_inverse */
static void quat_rotate_inverse(real *o, const real *q, const real *v) {
    real w = q[3], u[3] = {q[0], q[1], q[2]}, c[3];
    v3cross(c, u, v);
    real d = v3dot(u, v), k = 2 * w * w - 1;
    for (int i = 0; i < 3; i++) o[i] = v[i] * k - 2 * w * c[i] + 2 * u[i] * d;
}
/* isaacgym.torch_utils.get_euler_xyz: each angle returned modulo 2*pi */
static void get_euler_xyz(real *rpy, const real *q) {
    real x = q[0], y = q[1], z = q[2], w = q[3];
    real sinr = 2 * (w * x + y * z), cosr = w * w - x * x - y * y + z * z;
    real roll = atan2(sinr, cosr);
    real sinp = 2 * (w * y - z * x);
    real pitch = (fabs(sinp) >= 1) ? copysign(PI_R / 2, sinp) : asin(sinp);
    real siny = 2 * (w * z + x * y), cosy = w * w + x * x - y * y - z * z;
    real yaw = atan2(siny, cosy);
    real two_pi = 2 * PI_R;
    rpy[0] = fmod(roll, two_pi);  if (rpy[0] < 0) rpy[0] += two_pi;
    rpy[1] = fmod(pitch, two_pi); if (rpy[1] < 0) rpy[1] += two_pi;
    rpy[2] = fmod(yaw, two_pi);   if (rpy[2] < 0) rpy[2] += two_pi;
}
static void quat_from_euler_xyz(real *q, real roll, real pitch, real yaw) {
    real cy = cos(yaw * 0.5), sy = sin(yaw * 0.5), cr = cos(roll * 0.5), sr = sin(roll * 0.5);
    real cp = cos(pitch * 0.5), sp = sin(pitch * 0.5);
    q[3] = cy * cr * cp + sy * sr * sp;
    q[0] = cy * sr * cp - sy * cr * sp;
    q[1] = cy * cr * sp + sy * sr * cp;
    q[2] = sy * cr * cp - cy * sr * sp;
}

/* ------------------------------------------------------------------------------------------------ counter RNG */
/* Replaces the reference's torch.rand streams (legged_robot.py:403-462) by a counter-based generator keyed
 * (seed, global env, episode, stream, index) so sharded runs draw identical values. */
static inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
static inline uint32_t rng_u32(uint64_t seed, uint32_t env, uint32_t counter, uint32_t stream, uint32_t idx) {
    uint32_t h = mix32((uint32_t)seed ^ 0x9E3779B9U);
    h = mix32(h ^ (uint32_t)(seed >> 32));
    h = mix32(h ^ env);
    h = mix32(h ^ counter);
    h = mix32(h ^ (stream * 0x10001U + idx * 0x9E3779B1U));
    return h;
}
static inline real rng_uniform(uint64_t seed, uint32_t env, uint32_t counter, uint32_t stream, uint32_t idx) {
    return (real)(rng_u32(seed, env, counter, stream, idx) >> 8) * (real)(1.0 / 16777216.0);
}
static inline real rng_normal(uint64_t seed, uint32_t env, uint32_t counter, uint32_t stream, uint32_t idx) {
    real u1 = ((real)(rng_u32(seed, env, counter, stream, 2 * idx) >> 8) + 1) * (real)(1.0 / 16777216.0);
    real u2 = rng_uniform(seed, env, counter, stream, 2 * idx + 1);
    return sqrt(-2 * log(u1)) * cos(2 * PI_R * u2);
}
enum { RNG_DOF = 0, RNG_BASE_POS = 1, RNG_BASE_VEL = 2, RNG_NPC_POS = 3, RNG_NPC_RPY = 4, RNG_SHEEP = 5, RNG_PUSH = 6 };

/* ------------------------------------------------------------------------------------------------ spatial algebra
 * Everything of one robot is expressed in ONE frame: world axes, origin O = current base origin taken as an
 * instantaneous inertial point.  Motion vectors are (ang, lin-of-point-at-O), forces (moment about O, force). */
typedef struct { real w[3], v[3]; } sv;
typedef struct { real m, h[3], I[6]; } rbi; /* mass, first moment m*c, rotational inertia about O: xx xy xz yy yz zz */

static void rbi_from_link_mc(rbi *o, const float *in10, const real *R, const real *p, real added_mass, const real *com_shift);
static void rbi_from_link_m(rbi *o, const float *in10, const real *R, const real *p, real added_mass) { rbi_from_link_mc(o, in10, R, p, added_mass, NULL); }
static void rbi_from_link(rbi *o, const float *in10, const real *R, const real *p) { rbi_from_link_mc(o, in10, R, p, 0, NULL); }
static void rbi_from_link_mc(rbi *o, const float *in10, const real *R, const real *p, real added_mass, const real *com_shift) {
    /* in10: mass, com, I about com in link frame; R link->world, p link origin rel. O; added_mass sits at the COM;
     * com_shift (domain_rand.randomize_com, legged_robot_field.py:321-332) moves the COM, the tensor about it stays */
    real m = in10[0] + added_mass, cl[3] = {in10[1], in10[2], in10[3]}, c[3];
    if (com_shift) { cl[0] += com_shift[0]; cl[1] += com_shift[1]; cl[2] += com_shift[2]; }
    m3mulv(c, R, cl);
    v3add(c, c, p);
    real Il[9] = {in10[4], in10[5], in10[6], in10[5], in10[7], in10[8], in10[6], in10[8], in10[9]}, T[9], Iw[9], Rt[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rt[i * 3 + j] = R[j * 3 + i];
    m3mul(T, R, Il);
    m3mul(Iw, T, Rt);
    real cc = v3dot(c, c);
    o->m = m;
    for (int i = 0; i < 3; i++) o->h[i] = m * c[i];
    o->I[0] = Iw[0] + m * (cc - c[0] * c[0]);
    o->I[1] = Iw[1] - m * c[0] * c[1];
    o->I[2] = Iw[2] - m * c[0] * c[2];
    o->I[3] = Iw[4] + m * (cc - c[1] * c[1]);
    o->I[4] = Iw[5] - m * c[1] * c[2];
    o->I[5] = Iw[8] + m * (cc - c[2] * c[2]);
}
static void rbi_add(rbi *o, const rbi *a) {
    o->m += a->m;
    for (int i = 0; i < 3; i++) o->h[i] += a->h[i];
    for (int i = 0; i < 6; i++) o->I[i] += a->I[i];
}
static void rbi_mul(sv *f, const rbi *I, const sv *x) { /* f = I x */
    real hv[3], hw[3];
    v3cross(hv, I->h, x->v);
    v3cross(hw, I->h, x->w);
    f->w[0] = I->I[0] * x->w[0] + I->I[1] * x->w[1] + I->I[2] * x->w[2] + hv[0];
    f->w[1] = I->I[1] * x->w[0] + I->I[3] * x->w[1] + I->I[4] * x->w[2] + hv[1];
    f->w[2] = I->I[2] * x->w[0] + I->I[4] * x->w[1] + I->I[5] * x->w[2] + hv[2];
    for (int i = 0; i < 3; i++) f->v[i] = I->m * x->v[i] - hw[i];
}
static void crm(sv *o, const sv *a, const sv *b) { /* a x b (motion) */
    real t1[3], t2[3], t3[3];
    v3cross(t1, a->w, b->w);
    v3cross(t2, a->w, b->v);
    v3cross(t3, a->v, b->w);
    v3cpy(o->w, t1);
    v3add(o->v, t2, t3);
}
static void crf(sv *o, const sv *a, const sv *f) { /* a x* f (force) */
    real t1[3], t2[3], t3[3];
    v3cross(t1, a->w, f->w);
    v3cross(t2, a->v, f->v);
    v3cross(t3, a->w, f->v);
    v3add(o->w, t1, t2);
    v3cpy(o->v, t3);
}
static inline real svdot(const sv *a, const sv *b) { return v3dot(a->w, b->w) + v3dot(a->v, b->v); }

/* ------------------------------------------------------------------------------------------------ robot kinematics + dynamics */
typedef struct {
    real Rb[9];
    real R[13][9];      /* link rotations (0 = base, 1+3*leg+k)        */
    real p[13][3];      /* link frame origins rel. O (joint positions) */
    real a[12][3];      /* joint axes, world                           */
    sv S[12];
    rbi Il[13];
    sv vel[13];
    real M[NV][NV];
    real c[NV];
    real Minv[NV][NV];
} RobotDyn;

static void robot_kinematics_mc(RobotDyn *d, const MqeRobotModel *md, const real *quat, const real *q, real base_added_mass, const real *com_shift);
static void robot_kinematics_m(RobotDyn *d, const MqeRobotModel *md, const real *quat, const real *q, real base_added_mass) { robot_kinematics_mc(d, md, quat, q, base_added_mass, NULL); }
static void robot_kinematics(RobotDyn *d, const MqeRobotModel *md, const real *quat, const real *q) { robot_kinematics_mc(d, md, quat, q, 0, NULL); }
static void robot_kinematics_mc(RobotDyn *d, const MqeRobotModel *md, const real *quat, const real *q, real base_added_mass, const real *com_shift) {
    quat_to_mat(d->Rb, quat);
    memcpy(d->R[0], d->Rb, sizeof d->Rb);
    v3set(d->p[0], 0, 0, 0);
    rbi_from_link_mc(&d->Il[0], md->base_inertial, d->Rb, d->p[0], base_added_mass, com_shift);
    for (int l = 0; l < 4; l++) {
        const real *Rp = d->Rb;
        const real *pp = d->p[0];
        for (int k = 0; k < 3; k++) {
            int j = 3 * l + k, li = 1 + j;
            real off[3] = {md->leg_offsets[l][k][0], md->leg_offsets[l][k][1], md->leg_offsets[l][k][2]}, t[3];
            m3mulv(t, Rp, off);
            v3add(d->p[li], pp, t);
            real ax[3] = {k == 0 ? 1.0 : 0.0, k == 0 ? 0.0 : 1.0, 0.0};
            m3mulv(d->a[j], Rp, ax);
            real c = cos(q[j]), s = sin(q[j]), Rj[9];
            if (k == 0) { real t9[9] = {1, 0, 0, 0, c, -s, 0, s, c}; memcpy(Rj, t9, sizeof t9); }
            else        { real t9[9] = {c, 0, s, 0, 1, 0, -s, 0, c}; memcpy(Rj, t9, sizeof t9); }
            m3mul(d->R[li], Rp, Rj);
            v3cpy(d->S[j].w, d->a[j]);
            v3cross(d->S[j].v, d->p[li], d->a[j]);
            rbi_from_link(&d->Il[li], md->leg_inertial[l][k], d->R[li], d->p[li]);
            Rp = d->R[li];
            pp = d->p[li];
        }
    }
}

/* v: generalized velocity [wx wy wz vx vy vz qd0..11]; fills d->vel, d->M (CRBA), d->c (RNEA bias incl. gravity) */
static void robot_dynamics(RobotDyn *d, const real *v, real gz) {
    /* velocities */
    v3cpy(d->vel[0].w, v); v3cpy(d->vel[0].v, v + 3);
    sv acc[13], f[13];
    v3set(acc[0].w, 0, 0, 0); v3set(acc[0].v, 0, 0, -gz); /* a0 = -a_gravity */
    for (int l = 0; l < 4; l++)
        for (int k = 0; k < 3; k++) {
            int j = 3 * l + k, li = 1 + j, pi = (k == 0) ? 0 : li - 1;
            sv sq = d->S[j];
            for (int i = 0; i < 3; i++) { sq.w[i] *= v[6 + j]; sq.v[i] *= v[6 + j]; }
            for (int i = 0; i < 3; i++) { d->vel[li].w[i] = d->vel[pi].w[i] + sq.w[i]; d->vel[li].v[i] = d->vel[pi].v[i] + sq.v[i]; }
            sv cx;
            crm(&cx, &d->vel[pi], &sq); /* Sdot*qd = v_parent x (S qd) */
            for (int i = 0; i < 3; i++) { acc[li].w[i] = acc[pi].w[i] + cx.w[i]; acc[li].v[i] = acc[pi].v[i] + cx.v[i]; }
        }
    for (int b = 0; b < 13; b++) {
        sv Ia, Iv, vf;
        rbi_mul(&Ia, &d->Il[b], &acc[b]);
        rbi_mul(&Iv, &d->Il[b], &d->vel[b]);
        crf(&vf, &d->vel[b], &Iv);
        for (int i = 0; i < 3; i++) { f[b].w[i] = Ia.w[i] + vf.w[i]; f[b].v[i] = Ia.v[i] + vf.v[i]; }
    }
    rbi Ic[13];
    memcpy(Ic, d->Il, sizeof Ic);
    for (int l = 0; l < 4; l++)
        for (int k = 2; k >= 0; k--) {
            int li = 1 + 3 * l + k, pi = (k == 0) ? 0 : li - 1;
            for (int i = 0; i < 3; i++) { f[pi].w[i] += f[li].w[i]; f[pi].v[i] += f[li].v[i]; }
            rbi_add(&Ic[pi], &Ic[li]);
        }
    for (int i = 0; i < 3; i++) { d->c[i] = f[0].w[i]; d->c[3 + i] = f[0].v[i]; }
    for (int j = 0; j < 12; j++) d->c[6 + j] = svdot(&d->S[j], &f[1 + j]);
    /* mass matrix */
    memset(d->M, 0, sizeof d->M);
    {
        const rbi *I0 = &Ic[0];
        real I3[9] = {I0->I[0], I0->I[1], I0->I[2], I0->I[1], I0->I[3], I0->I[4], I0->I[2], I0->I[4], I0->I[5]};
        real hx[9] = {0, -I0->h[2], I0->h[1], I0->h[2], 0, -I0->h[0], -I0->h[1], I0->h[0], 0};
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                d->M[i][j] = I3[i * 3 + j];
                d->M[i][3 + j] = hx[i * 3 + j];
                d->M[3 + i][j] = hx[j * 3 + i];
                d->M[3 + i][3 + j] = (i == j) ? I0->m : 0;
            }
    }
    for (int l = 0; l < 4; l++)
        for (int k = 0; k < 3; k++) {
            int j = 3 * l + k;
            sv F;
            rbi_mul(&F, &Ic[1 + j], &d->S[j]);
            d->M[6 + j][6 + j] = svdot(&d->S[j], &F);
            for (int k2 = 0; k2 < k; k2++) {
                int j2 = 3 * l + k2;
                real m = svdot(&d->S[j2], &F);
                d->M[6 + j][6 + j2] = m;
                d->M[6 + j2][6 + j] = m;
            }
            for (int i = 0; i < 3; i++) {
                d->M[i][6 + j] = d->M[6 + j][i] = F.w[i];
                d->M[3 + i][6 + j] = d->M[6 + j][3 + i] = F.v[i];
            }
        }
}

/* dense SPD inverse via Cholesky; returns 0 on success */
static int spd_inverse(int n, const real *A, real *Ainv) {
    real L[NV * NV];
    memset(L, 0, sizeof L);
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            real s = A[i * n + j];
            for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
            if (i == j) {
                if (s <= 0) return -1;
                L[i * n + i] = sqrt(s);
            } else L[i * n + j] = s / L[j * n + j];
        }
    for (int c = 0; c < n; c++) {
        real y[NV];
        for (int i = 0; i < n; i++) {
            real s = (i == c) ? 1 : 0;
            for (int k = 0; k < i; k++) s -= L[i * n + k] * y[k];
            y[i] = s / L[i * n + i];
        }
        for (int i = n - 1; i >= 0; i--) {
            real s = y[i];
            for (int k = i + 1; k < n; k++) s -= L[k * n + i] * Ainv[k * n + c];
            Ainv[i * n + c] = s / L[i * n + i];
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ networks */
static inline real elu(real x) { return x > 0 ? x : (real)expm1((double)x); }
static inline real softsign(real x) { return x / (1 + fabs(x)); }

static void linear(const float *W, const float *b, int nout, int nin, const real *x, real *y) {
    for (int o = 0; o < nout; o++) {
        const float *w = W + (size_t)o * nin;
        if (sizeof(real) == sizeof(float)) {      /* fp32 build: 8 partial sums so the compiler can vectorise */
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            int i = 0;
            for (; i + 8 <= nin; i += 8)
                for (int k = 0; k < 8; k++) acc[k] += w[i + k] * (float)x[i + k];
            float s = ((acc[0] + acc[4]) + (acc[1] + acc[5])) + ((acc[2] + acc[6]) + (acc[3] + acc[7]));
            for (; i < nin; i++) s += w[i] * (float)x[i];
            y[o] = (real)(s + b[o]);
        } else {
            double s = b[o];
            for (int i = 0; i < nin; i++) s += (double)w[i] * (double)x[i];
            y[o] = (real)s;
        }
    }
}
/* go1.py:400-407: latent = AM(h); action = Body(cat(h, latent)) */
static void policy_forward(const MqeWeights *w, const real *hist2100, real *latent2, real *action12) {
    real h0[256], h1[128], x2[2102], b0[512], b1[256], b2[128];
    linear(w->adapt_w0, w->adapt_b0, 256, 2100, hist2100, h0);
    for (int i = 0; i < 256; i++) h0[i] = elu(h0[i]);
    linear(w->adapt_w1, w->adapt_b1, 128, 256, h0, h1);
    for (int i = 0; i < 128; i++) h1[i] = elu(h1[i]);
    linear(w->adapt_w2, w->adapt_b2, 2, 128, h1, latent2);
    memcpy(x2, hist2100, 2100 * sizeof(real));
    x2[2100] = latent2[0]; x2[2101] = latent2[1];
    linear(w->body_w0, w->body_b0, 512, 2102, x2, b0);
    for (int i = 0; i < 512; i++) b0[i] = elu(b0[i]);
    linear(w->body_w1, w->body_b1, 256, 512, b0, b1);
    for (int i = 0; i < 256; i++) b1[i] = elu(b1[i]);
    linear(w->body_w2, w->body_b2, 128, 256, b1, b2);
    for (int i = 0; i < 128; i++) b2[i] = elu(b2[i]);
    linear(w->body_w3, w->body_b3, 12, 128, b2, action12);
}
/* go1.py:369-380, unitree_go1.pt: 6 -> 32 -> 32 -> 1, softsign */
static real actuator_forward(const MqeWeights *w, const real *x6) {
    real h0[32], h1[32], y;
    linear(w->act_w0, w->act_b0, 32, 6, x6, h0);
    for (int i = 0; i < 32; i++) h0[i] = softsign(h0[i]);
    linear(w->act_w1, w->act_b1, 32, 32, h0, h1);
    for (int i = 0; i < 32; i++) h1[i] = softsign(h1[i]);
    linear(w->act_w2, w->act_b2, 1, 32, h1, &y);
    return y;
}

/* ------------------------------------------------------------------------------------------------ oracle state */
typedef struct {
    MqeSimDesc d;
    int N, A, P, D, M, NB; /* envs, agents, npcs, npc dofs/env, N*A, rigid bodies per env */
    float *sdf;
    real *env_origins, *agent_origins, *base_init, *npc_init, *npc_dof_default;
    float *wbuf[20];
    /* state (names follow the reference attributes) */
    real *root;        /* all_root_states [N][A+P][13]            */
    real *dof;         /* all_dof_states  [N][12A+D][2]           */
    real *contact;     /* contact_forces  [N][NB][3]              */
    real *torques;     /* [N][12A]                                */
    real *actions, *last_actions;           /* [N][12A]                         */
    real *loc_last, *loc_last2;             /* last_locomotion_action(s) [M][12] */
    real *loc_obs;                          /* [M][70]                          */
    real *hist;                             /* history_locomotion_obs [M][2100] */
    real *err1, *err2, *vel1, *vel2;        /* actuator histories [M][12]       */
    real *gait, *clock;                     /* [M], [M][4]                      */
    real *base_quat, *base_lin_vel, *base_ang_vel, *proj_grav; /* [M][4|3]     */
    real *obs;                              /* [M][71]                          */
    real *commands;                         /* [M][3]                           */
    real *last_dof_vel, *last_root_vel;
    real *sub_tau, *sub_qd; uint8_t *sub_exceed;   /* post_decimation_step logs [N][decimation][12A] (legged_robot.py:112-115) */
    real *sheep_stats;                      /* [N][3]                           */
    float *mu_env;                          /* [N] per-env friction or NULL (domain_rand.randomize_friction) */
    real *lag_ring; int lag_n; uint32_t lag_calls;   /* action lag (go1.py:337-339, 363): [M][lag_n][12] scaled actions; calls so far */
    float *base_mass_add;                   /* [M] mass added to the base link or NULL (domain_rand.randomize_base_mass) */
    float *base_com_shift;                  /* [M][3] base-link COM shift or NULL (domain_rand.randomize_com) */
    float *motor_strength;                  /* [M][12] action factor for control types P / V / T or NULL (domain_rand.randomize_motor) */
    int64_t *ep_len;
    uint8_t *reset_buf, *timeout_buf, *collide_buf, *r_term, *p_term, *zl_term, *zh_term;
    uint32_t *episode;                      /* reset counter per env (RNG key)  */
    uint32_t step_count;
    int32_t stats[8];
    int reset_state;   /* 1 (default): reset_idx re-initialises DOF / root state from the counter RNG; 0 (golden replays): only _reset_buffers */
} Oracle;

static void *xcalloc(size_t n, size_t s) { void *p = calloc(n ? n : 1, s); if (!p) abort(); return p; }
static float *dupf(const float *src, size_t n) {
    float *p = (float *)xcalloc(n, sizeof(float));
    if (src) memcpy(p, src, n * sizeof(float));
    return p;
}
static real *dupr(const float *src, size_t n) {
    real *p = (real *)xcalloc(n, sizeof(real));
    if (src) for (size_t i = 0; i < n; i++) p[i] = src[i];
    return p;
}

Oracle *orc_create(const MqeSimDesc *desc) {
    Oracle *o = (Oracle *)xcalloc(1, sizeof(Oracle));
    o->d = *desc;
    int N = o->N = desc->num_envs, A = o->A = desc->num_agents, P = o->P = desc->num_npcs;
    o->D = desc->npc_dofs;
    o->M = N * A;
    o->NB = MQE_NUM_BODIES * A + P;
    int M = o->M;
    o->sdf = dupf(desc->h_sdf, (size_t)desc->sdf_nx * desc->sdf_ny);
    o->env_origins = dupr(desc->h_env_origins, (size_t)N * 3);
    o->agent_origins = dupr(desc->h_agent_origins, (size_t)M * 3);
    o->base_init = dupr(desc->h_base_init_state, (size_t)M * 13);
    o->npc_init = dupr(desc->h_npc_init_state, (size_t)N * P * 13);
    o->npc_dof_default = dupr(desc->h_npc_dof_default, (size_t)(o->D ? o->D : 1));
    o->mu_env = desc->h_env_friction ? dupf(desc->h_env_friction, (size_t)N) : NULL;
    if (desc->lag_enabled) { o->lag_n = desc->lag_timesteps + 1; o->lag_ring = (real *)xcalloc((size_t)M * o->lag_n * 12, sizeof(real)); }
    o->base_mass_add = desc->h_base_added_mass ? dupf(desc->h_base_added_mass, (size_t)M) : NULL;
    o->base_com_shift = desc->h_base_com_shift ? dupf(desc->h_base_com_shift, (size_t)M * 3) : NULL;
    o->motor_strength = desc->h_motor_strength ? dupf(desc->h_motor_strength, (size_t)M * 12) : NULL;
    /* private copies of the weights */
    const float **src = (const float **)&desc->weights;
    const size_t sz[20] = {256 * 2100, 256, 128 * 256, 128, 2 * 128, 2, 512 * 2102, 512, 256 * 512, 256,
                           128 * 256, 128, 12 * 128, 12, 32 * 6, 32, 32 * 32, 32, 32, 1};
    const float **dst = (const float **)&o->d.weights;
    for (int i = 0; i < 20; i++) { o->wbuf[i] = dupf(src[i], sz[i]); dst[i] = o->wbuf[i]; }
    o->d.h_sdf = NULL;
#define RA(n) (real *)xcalloc((size_t)(n), sizeof(real))
    o->root = RA(N * (A + P) * 13); o->dof = RA(N * (12 * A + o->D) * 2); o->contact = RA(N * o->NB * 3);
    o->torques = RA(M * 12); o->actions = RA(M * 12); o->last_actions = RA(M * 12);
    o->loc_last = RA(M * 12); o->loc_last2 = RA(M * 12); o->loc_obs = RA(M * 70); o->hist = RA((size_t)M * 2100);
    o->err1 = RA(M * 12); o->err2 = RA(M * 12); o->vel1 = RA(M * 12); o->vel2 = RA(M * 12);
    o->gait = RA(M); o->clock = RA(M * 4);
    o->base_quat = RA(M * 4); o->base_lin_vel = RA(M * 3); o->base_ang_vel = RA(M * 3); o->proj_grav = RA(M * 3);
    o->obs = RA(M * MQE_OBS_FLOATS); o->commands = RA(M * 3);
    o->last_dof_vel = RA(M * 12); o->last_root_vel = RA(M * 6); o->sheep_stats = RA(N * 3);
    o->sub_tau = RA((size_t)M * 12 * desc->decimation); o->sub_qd = RA((size_t)M * 12 * desc->decimation);
    o->sub_exceed = (uint8_t *)xcalloc((size_t)M * 12 * desc->decimation, 1);
#undef RA
    o->ep_len = (int64_t *)xcalloc(N, sizeof(int64_t));
    o->reset_buf = (uint8_t *)xcalloc(N, 1); o->timeout_buf = (uint8_t *)xcalloc(N, 1);
    o->collide_buf = (uint8_t *)xcalloc(N, 1); o->r_term = (uint8_t *)xcalloc(N, 1);
    o->p_term = (uint8_t *)xcalloc(N, 1); o->zl_term = (uint8_t *)xcalloc(N, 1); o->zh_term = (uint8_t *)xcalloc(N, 1);
    o->episode = (uint32_t *)xcalloc(N, sizeof(uint32_t));
    memset(o->reset_buf, 1, N); /* base_task.py:77 */
    o->reset_state = 1;
    /* _prepare_locomotion_policy: locomotion_obs = default command frame repeated (go1.py:393-394) */
    for (int m = 0; m < M; m++)
        for (int i = 0; i < 70; i++) o->loc_obs[m * 70 + i] = desc->loc_obs_default[i];
    /* actors are created at their start poses (legged_robot.py:864-874); reset() overwrites them */
    for (int e = 0; e < N; e++) {
        for (int a = 0; a < A; a++) {
            real *r = o->root + ((size_t)e * (A + P) + a) * 13;
            for (int i = 0; i < 13; i++) r[i] = o->base_init[(e * A + a) * 13 + i];
            for (int i = 0; i < 3; i++) r[i] += o->agent_origins[(e * A + a) * 3 + i];
            for (int j = 0; j < 12; j++) o->dof[((size_t)e * (12 * A + o->D) + 12 * a + j) * 2] = desc->model.q_default[j];
            /* _init_buffers (legged_robot.py:570, 620-622): base_quat = spawn quaternion; base_lin_vel / base_ang_vel / projected_gravity
             * derived from the spawn state.  reset_idx does not recompute them, so the first reset()'s observation carries these. */
            int m = e * A + a;
            const real g[3] = {0, 0, -1};
            for (int i = 0; i < 4; i++) o->base_quat[m * 4 + i] = r[3 + i];
            quat_rotate_inverse(o->base_lin_vel + m * 3, r + 3, r + 7);
            quat_rotate_inverse(o->base_ang_vel + m * 3, r + 3, r + 10);
            quat_rotate_inverse(o->proj_grav + m * 3, r + 3, g);
        }
        for (int p = 0; p < P; p++) {
            real *r = o->root + ((size_t)e * (A + P) + A + p) * 13;
            for (int i = 0; i < 13; i++) r[i] = o->npc_init[(e * P + p) * 13 + i];
            for (int i = 0; i < 3; i++) r[i] += o->env_origins[e * 3 + i];
        }
    }
    return o;
}

void orc_destroy(Oracle *o) {
    if (!o) return;
    free(o->sdf); free(o->env_origins); free(o->agent_origins); free(o->base_init); free(o->npc_init); free(o->npc_dof_default); free(o->mu_env); free(o->base_mass_add); free(o->base_com_shift); free(o->motor_strength); free(o->lag_ring);
    for (int i = 0; i < 20; i++) free(o->wbuf[i]);
    free(o->root); free(o->dof); free(o->contact); free(o->torques); free(o->actions); free(o->last_actions);
    free(o->loc_last); free(o->loc_last2); free(o->loc_obs); free(o->hist); free(o->err1); free(o->err2); free(o->vel1); free(o->vel2);
    free(o->gait); free(o->clock); free(o->base_quat); free(o->base_lin_vel); free(o->base_ang_vel); free(o->proj_grav);
    free(o->obs); free(o->commands); free(o->last_dof_vel); free(o->last_root_vel); free(o->sheep_stats);
    free(o->sub_tau); free(o->sub_qd); free(o->sub_exceed);
    free(o->ep_len); free(o->reset_buf); free(o->timeout_buf); free(o->collide_buf); free(o->r_term); free(o->p_term);
    free(o->zl_term); free(o->zh_term); free(o->episode);
    free(o);
}

/* ------------------------------------------------------------------------------------------------ static world */
typedef struct { real sdf, gx, gy; } SdfSample;

static SdfSample sdf_sample(const Oracle *o, real x, real y) {
    const MqeSimDesc *d = &o->d;
    real fx = x / d->sdf_cell, fy = y / d->sdf_cell;
    real mx = (real)(d->sdf_nx - 1) - (real)1e-3, my = (real)(d->sdf_ny - 1) - (real)1e-3;
    if (fx < 0) fx = 0; if (fx > mx) fx = mx;
    if (fy < 0) fy = 0; if (fy > my) fy = my;
    int i = (int)fx, j = (int)fy;
    real tx = fx - i, ty = fy - j;
    const float *S = o->sdf;
    int ny = d->sdf_ny;
    real s00 = S[i * ny + j], s10 = S[(i + 1) * ny + j], s01 = S[i * ny + j + 1], s11 = S[(i + 1) * ny + j + 1];
    SdfSample r;
    r.sdf = (1 - tx) * (1 - ty) * s00 + tx * (1 - ty) * s10 + (1 - tx) * ty * s01 + tx * ty * s11;
    r.gx = ((1 - ty) * (s10 - s00) + ty * (s11 - s01)) / d->sdf_cell;
    r.gy = ((1 - tx) * (s01 - s00) + tx * (s11 - s10)) / d->sdf_cell;
    return r;
}

/* ------------------------------------------------------------------------------------------------ contacts / rows */
typedef struct {
    int ga, la, rba;   /* dynamic group (0..A-1 robots, A.. npcs), link within robot, rigid body for reporting */
    int gb, lb, rbb;   /* gb = -1: static world                                                                */
    real pos[3];       /* world contact point                                                                  */
    real n[3];         /* from B to A                                                                          */
    real gap;
} Contact;

typedef struct {
    int ga, gb;                 /* groups (gb=-1 none)                      */
    real Ja[NV], Jb[NV];        /* dense over the group's dofs              */
    real Ya[NV], Yb[NV];        /* Minv J^T                                 */
    real dinv, bias, lambda;
    int kind;                   /* 0 unilateral (lambda>=0), 1 friction (|lambda| <= mu * lambda[normal_row]) */
    int normal_row;
    int contact;                /* index into contact list or -1            */
    real dir[3];
} Row;

static void tangent_basis(const real *n, real *t1, real *t2) {
    real e[3] = {0, 0, 0};
    if (fabs(n[0]) < (real)0.9) e[0] = 1; else e[1] = 1;
    v3cross(t1, n, e);
    real inv = 1 / sqrt(v3dot(t1, t1));
    for (int i = 0; i < 3; i++) t1[i] *= inv;
    v3cross(t2, n, t1);
}

/* world probe of a sphere (centre x world, radius r) against floor slab + wall footprint.
 * Emits up to 2 contacts (floor/top, wall side). */
static int probe_world(const Oracle *o, const real *x, real r, const real *fix, Contact *out) {
    const MqeSimDesc *d = &o->d;
    int n = 0;
    SdfSample s = sdf_sample(o, x[0], x[1]);
    int inside = s.sdf < 0;
    int above_top = x[2] >= d->wall_top_z;
    real ground = (inside && above_top) ? d->wall_top_z : d->floor_z;
    /* wall-like candidate: BarrierTrack footprint ... */
    real gw = s.sdf - r, gn = sqrt(s.gx * s.gx + s.gy * s.gy), wn[2] = {0, 0};
    int wall_ok = !above_top && gn > (real)1e-6;
    if (wall_ok) { wn[0] = s.gx / gn; wn[1] = s.gy / gn; }
    if (fix && d->npc_kind == MQE_NPC_PLATFORM) {   /* wrestling.urdf / bridge.urdf: tops of fixed boxes are ground inside their footprints */
        const float *g = d->npc_geom;
        for (int b = 0; b < (int)g[0]; b++) {
            const float *bx = g + 1 + 5 * b;
            real px = x[0] - fix[0] - bx[0], py = x[1] - fix[1] - bx[1], top = fix[2] + bx[4];
            if (fabs(px) <= bx[2] && fabs(py) <= bx[3] && x[2] >= top - (real)0.15 && top > ground) ground = top;
        }
    } else if (fix) { /* seesaw.urdf statics: platform top is ground inside its footprint, the column is a vertical cylinder */
        const float *g = d->npc_geom;
        real px = x[0] - fix[0], py = x[1] - fix[1];
        if (g[7] > 0 && fabs(px) <= g[7] && fabs(py) <= g[8] && x[2] >= fix[2]) { real top = fix[2] + g[9]; if (top > ground) ground = top; }
        if (g[10] > 0 && x[2] < fix[2] && x[2] > fix[2] - g[11] - r) {
            real dh = sqrt(px * px + py * py), gc = dh - g[10] - r;
            if (dh > (real)1e-6 && (!wall_ok || gc < gw)) { wall_ok = 1; gw = gc; wn[0] = px / dh; wn[1] = py / dh; }
        }
    }
    real gap = x[2] - r - ground;
    if (gap < d->contact_offset) {
        Contact *c = &out[n++];
        v3set(c->n, 0, 0, 1);
        c->gap = gap;
        v3set(c->pos, x[0], x[1], x[2] - r - gap * (real)0.5);
    }
    if (wall_ok && gw < d->contact_offset) {
        Contact *c = &out[n++];
        v3set(c->n, wn[0], wn[1], 0);
        c->gap = gw;
        for (int i = 0; i < 3; i++) c->pos[i] = x[i] - c->n[i] * (r + gw * (real)0.5);
    }
    return n;
}

/* sphere (centre x, radius r) vs an oriented box (centre c, axes ex ey ez, half extents h).  Returns 1 and fills the
 * normal (box -> sphere), gap and contact point when gap < contact_offset.  Used for the seesaw plank and the push box. */
static int sphere_obb(const MqeSimDesc *d, const real *c, const real *ex, const real *ey, const real *ez, const real *h,
                      const real *x, real r, real *nrm, real *gap_out, real *pos) {
    real dx[3] = {x[0] - c[0], x[1] - c[1], x[2] - c[2]};
    real loc[3] = {v3dot(dx, ex), v3dot(dx, ey), v3dot(dx, ez)}, q[3], df[3], nl[3] = {0, 0, 0}, gap;
    for (int i = 0; i < 3; i++) { q[i] = loc[i] < -h[i] ? -h[i] : (loc[i] > h[i] ? h[i] : loc[i]); df[i] = loc[i] - q[i]; }
    real d2 = df[0] * df[0] + df[1] * df[1] + df[2] * df[2];
    if (d2 > (real)1e-12) {
        real dist = sqrt(d2);
        for (int i = 0; i < 3; i++) nl[i] = df[i] / dist;
        gap = dist - r;
    } else {            /* centre inside the box: push out along the axis of least penetration */
        int ax = 0;
        real best = h[0] - fabs(loc[0]);
        for (int i = 1; i < 3; i++) { real pen = h[i] - fabs(loc[i]); if (pen < best) { best = pen; ax = i; } }
        nl[ax] = loc[ax] >= 0 ? 1 : -1;
        gap = -best - r;
    }
    if (gap >= d->contact_offset) return 0;
    for (int i = 0; i < 3; i++) nrm[i] = nl[0] * ex[i] + nl[1] * ey[i] + nl[2] * ez[i];
    for (int i = 0; i < 3; i++) pos[i] = x[i] - nrm[i] * (r + gap * (real)0.5);
    *gap_out = gap;
    return 1;
}

/* sphere (centre x, radius r) vs a solid vertical cylinder (centre c, radius R, half height hh); normal cylinder -> sphere.
 * The tug-of-war disc: resources/objects/cylinder.urdf (collision cylinder r 1.2, length 0.5). */
static int sphere_vcyl(const MqeSimDesc *d, const real *c, real R, real hh, const real *x, real r, real *nrm, real *gap_out, real *pos) {
    real dx[3] = {x[0] - c[0], x[1] - c[1], x[2] - c[2]};
    real dr = sqrt(dx[0] * dx[0] + dx[1] * dx[1]);
    real ux = dr > (real)1e-9 ? dx[0] / dr : 1, uy = dr > (real)1e-9 ? dx[1] / dr : 0;
    real qz = dx[2] < -hh ? -hh : (dx[2] > hh ? hh : dx[2]);
    real dfr = dr - (dr < R ? dr : R), dfz = dx[2] - qz, gap;
    real d2 = dfr * dfr + dfz * dfz;
    if (d2 > (real)1e-12) {
        real dist = sqrt(d2);
        nrm[0] = ux * dfr / dist; nrm[1] = uy * dfr / dist; nrm[2] = dfz / dist;
        gap = dist - r;
    } else {
        real penr = R - dr, penz = hh - fabs(dx[2]);
        if (penr < penz) { nrm[0] = ux; nrm[1] = uy; nrm[2] = 0; gap = -penr - r; }
        else { nrm[0] = 0; nrm[1] = 0; nrm[2] = dx[2] >= 0 ? 1 : -1; gap = -penz - r; }
    }
    if (gap >= d->contact_offset) return 0;
    for (int i = 0; i < 3; i++) pos[i] = x[i] - nrm[i] * (r + gap * (real)0.5);
    *gap_out = gap;
    return 1;
}

/* closest points of two segments (Ericson, Real-Time Collision Detection 5.1.9) */
static void seg_seg(const real *p1, const real *q1, const real *p2, const real *q2, real *c1, real *c2) {
    real d1[3], d2[3], r[3];
    v3sub(d1, q1, p1); v3sub(d2, q2, p2); v3sub(r, p1, p2);
    real a = v3dot(d1, d1), e = v3dot(d2, d2), f = v3dot(d2, r), s, t;
    const real EPS = (real)1e-12;
    if (a <= EPS && e <= EPS) { s = t = 0; }
    else if (a <= EPS) { s = 0; t = f / e; t = t < 0 ? 0 : (t > 1 ? 1 : t); }
    else {
        real c = v3dot(d1, r);
        if (e <= EPS) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
        else {
            real b = v3dot(d1, d2), denom = a * e - b * b;
            if (denom > EPS) { s = (b * f - c * e) / denom; s = s < 0 ? 0 : (s > 1 ? 1 : s); } else s = 0;
            t = (b * s + f) / e;
            if (t < 0) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
            else if (t > 1) { t = 1; s = (b - c) / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
        }
    }
    for (int i = 0; i < 3; i++) { c1[i] = p1[i] + d1[i] * s; c2[i] = p2[i] + d2[i] * t; }
}

/* per-env scratch */
typedef struct {
    RobotDyn rd[4];
    real origin[16][3];            /* group origin in world (base origin / npc com) */
    real vel[16][NV];              /* generalized velocities being solved           */
    int ndof[16];
    real npc_minv[2];              /* 1/I, 1/m                                      */
} EnvScratch;

static void point_jacobian_robot(const RobotDyn *rd, int link, const real *r, const real *dir, real *J) {
    /* row of d/dv of (dir . velocity of the point at r (rel. O) fixed in `link`) */
    real rxd[3];
    v3cross(rxd, r, dir);
    memset(J, 0, NV * sizeof(real));
    for (int i = 0; i < 3; i++) { J[i] = rxd[i]; J[3 + i] = dir[i]; }
    if (link > 0) {
        int l = (link - 1) / 3, k = (link - 1) % 3;
        for (int k2 = 0; k2 <= k; k2++) {
            int j = 3 * l + k2;
            real rel[3], t[3];
            v3sub(rel, r, rd->p[1 + j]);
            v3cross(t, rel, dir);
            J[6 + j] = v3dot(rd->a[j], t);
        }
    }
}

static void group_jacobian(const Oracle *o, const EnvScratch *es, int g, int link, const real *pos, const real *dir, real *J, real *Y) {
    real r[3];
    v3sub(r, pos, es->origin[g]);
    if (g < o->A) {
        point_jacobian_robot(&es->rd[g], link, r, dir, J);
        for (int i = 0; i < NV; i++) {
            real s = 0;
            for (int k = 0; k < NV; k++) s += es->rd[g].Minv[i][k] * J[k];
            Y[i] = s;
        }
    } else { /* rigid npc with isotropic inertia about its com */
        real rxd[3];
        v3cross(rxd, r, dir);
        memset(J, 0, NV * sizeof(real)); memset(Y, 0, NV * sizeof(real));
        /* sheep are kept upright: go1_sheep.py:61 zeroes their tilt every policy step, so contacts get no roll/pitch
         * response; the ball is a free sphere */
        real up = (o->d.npc_ctrl == MQE_NPC_SHEEP) ? 0 : 1;
        for (int i = 0; i < 3; i++) { J[i] = rxd[i]; J[3 + i] = dir[i]; Y[i] = rxd[i] * es->npc_minv[0] * (i < 2 ? up : 1); Y[3 + i] = dir[i] * es->npc_minv[1]; }
        if (o->d.npc_kind == MQE_NPC_SEESAW) {   /* one revolute DOF about the pivot (= group origin): only w_axis responds */
            int ax = o->d.npc_geom[13] > 0.5f ? 2 : 1;   /* seesaw: y, revolving door: z */
            memset(Y, 0, NV * sizeof(real));
            if (o->d.npc_geom[13] > 1.5f) Y[4] = dir[1] * es->npc_minv[1];   /* prismatic y (tug disc): only v_y responds */
            else Y[ax] = rxd[ax] * es->npc_minv[0];
        }
    }
}

static real contact_bias(const MqeSimDesc *d, real gap) {
    real dt = d->sim_dt;
    if (gap > 0) return gap / dt;
    real b = d->erp * gap / dt;
    return b < -d->max_depen_vel ? -d->max_depen_vel : b;
}

static int add_contact_rows(const Oracle *o, const EnvScratch *es, const Contact *c, int ci, Row *rows, int nr) {
    real t1[3], t2[3];
    tangent_basis(c->n, t1, t2);
    const real *dirs[3] = {c->n, t1, t2};
    for (int k = 0; k < 3; k++) {
        Row *r = &rows[nr + k];
        memset(r, 0, sizeof *r);
        r->ga = c->ga; r->gb = c->gb; r->contact = ci;
        v3cpy(r->dir, dirs[k]);
        group_jacobian(o, es, c->ga, c->la, c->pos, dirs[k], r->Ja, r->Ya);
        real dd = 0;
        for (int i = 0; i < NV; i++) dd += r->Ja[i] * r->Ya[i];
        if (c->gb >= 0) {
            real neg[3] = {-dirs[k][0], -dirs[k][1], -dirs[k][2]};
            group_jacobian(o, es, c->gb, c->lb, c->pos, neg, r->Jb, r->Yb);
            for (int i = 0; i < NV; i++) dd += r->Jb[i] * r->Yb[i];
        }
        r->dinv = 1 / (dd + o->d.cfm);
        r->kind = (k == 0) ? 0 : 1;
        r->normal_row = nr;
        r->bias = (k == 0) ? contact_bias(&o->d, c->gap) : 0;
    }
    return nr + 3;
}

static void solve_row(Row *r, Row *rows, EnvScratch *es, real mu) {
    real u = r->bias;
    for (int i = 0; i < NV; i++) u += r->Ja[i] * es->vel[r->ga][i];
    if (r->gb >= 0) for (int i = 0; i < NV; i++) u += r->Jb[i] * es->vel[r->gb][i];
    real lam = r->lambda - u * r->dinv;
    if (r->kind == 0) { if (lam < 0) lam = 0; }
    else {
        real lim = mu * rows[r->normal_row].lambda;
        if (lam > lim) lam = lim;
        if (lam < -lim) lam = -lim;
    }
    real dl = lam - r->lambda;
    r->lambda = lam;
    for (int i = 0; i < NV; i++) es->vel[r->ga][i] += r->Ya[i] * dl;
    if (r->gb >= 0) for (int i = 0; i < NV; i++) es->vel[r->gb][i] += r->Yb[i] * dl;
}

/* contacts of every robot probe (robot X ascending, probe-table order) with the oriented box of NPC group A
 * (ex == NULL: a vertical cylinder of radius h[0] and half height h[1] instead) */
static int obb_probe_contacts(const Oracle *o, const EnvScratch *es, const MqeRobotModel *md, const real *c, const real *ex, const real *ey,
                              const real *ez, const real *h, Contact *contacts, int *nc_io, Row *rows, int *nr_io, int npair, int max_pair) {
    const MqeSimDesc *d = &o->d;
    int A = o->A, nc = *nc_io, nr = *nr_io;
    for (int X = 0; X < A; X++)
        for (int pi = 0; pi < md->n_probes; pi++) {
            const float *pr = md->probes[pi];
            int link = (int)pr[0], body = (int)pr[1];
            real loc[3] = {pr[2], pr[3], pr[4]}, x[3], nrm[3], pos[3], gap;
            m3mulv(x, es->rd[X].R[link], loc);
            v3add(x, x, es->rd[X].p[link]);
            v3add(x, x, es->origin[X]);
            if (npair >= max_pair) continue;
            if (ex ? !sphere_obb(d, c, ex, ey, ez, h, x, pr[5], nrm, &gap, pos) : !sphere_vcyl(d, c, h[0], h[1], x, pr[5], nrm, &gap, pos)) continue;
            Contact *ct = &contacts[nc];
            ct->ga = X; ct->la = link; ct->rba = X * MQE_NUM_BODIES + body; ct->gb = A; ct->lb = 0; ct->rbb = A * MQE_NUM_BODIES;
            v3cpy(ct->n, nrm); v3cpy(ct->pos, pos); ct->gap = gap;
            nr = add_contact_rows(o, es, ct, nc, rows, nr);
            nc++; npair++;
        }
    *nc_io = nc; *nr_io = nr;
    return npair;
}

/* ------------------------------------------------------------------------------------------------ one physics substep of one env
 * Replaces gym.set_dof_actuation_force_tensor + gym.simulate + gym.refresh_dof_state_tensor (go1.py:52-56). */
static void env_substep(Oracle *o, int e, const real *tau /* [12A] */, int32_t *stats) {
    const MqeSimDesc *d = &o->d;
    const MqeRobotModel *md = &d->model;
    int A = o->A, P = o->P, G = A + P;
    real dt = d->sim_dt;
    EnvScratch es;
    memset(&es, 0, sizeof es);
    real *root = o->root + (size_t)e * G * 13;
    real *dof = o->dof + (size_t)e * (12 * A + o->D) * 2;
    real q[4][12];
    Row rows[(MAX_LOCAL_CONTACTS * 3 + MAX_LIMIT_ROWS) * 16 + MAX_PAIR_CONTACTS * 3];
    Contact contacts[MAX_LOCAL_CONTACTS * 16 + MAX_PAIR_CONTACTS];
    int nr = 0, nc = 0;

    /* 1. unconstrained velocities */
    for (int a = 0; a < A; a++) {
        real *rs = root + a * 13;
        RobotDyn *rd = &es.rd[a];
        real v[NV], rhs[NV], acc[NV];
        for (int j = 0; j < 12; j++) { q[a][j] = dof[(12 * a + j) * 2]; v[6 + j] = dof[(12 * a + j) * 2 + 1]; }
        for (int i = 0; i < 3; i++) { v[i] = rs[10 + i]; v[3 + i] = rs[7 + i]; es.origin[a][i] = rs[i]; }
        real csh[3] = {0, 0, 0};
        if (o->base_com_shift) for (int i = 0; i < 3; i++) csh[i] = o->base_com_shift[(e * A + a) * 3 + i];
        robot_kinematics_mc(rd, md, rs + 3, q[a], o->base_mass_add ? (real)o->base_mass_add[e * A + a] : 0, o->base_com_shift ? csh : NULL);
        robot_dynamics(rd, v, d->gravity_z);
        if (spd_inverse(NV, &rd->M[0][0], &rd->Minv[0][0]) != 0) { fprintf(stderr, "oracle: mass matrix not SPD (env %d)\n", e); abort(); }
        for (int i = 0; i < NV; i++) rhs[i] = -rd->c[i];
        for (int j = 0; j < 12; j++) rhs[6 + j] += tau[12 * a + j];
        for (int i = 0; i < NV; i++) { real s = 0; for (int k = 0; k < NV; k++) s += rd->Minv[i][k] * rhs[k]; acc[i] = s; }
        real wxv[3];
        v3cross(wxv, v, v + 3); /* spatial -> classical acceleration of the base origin */
        for (int i = 0; i < NV; i++) es.vel[a][i] = v[i] + dt * acc[i];
        for (int i = 0; i < 3; i++) es.vel[a][3 + i] += dt * wxv[i];
        es.ndof[a] = NV;
    }
    const int rigid = d->npc_kind == MQE_NPC_RIGID || d->npc_kind == MQE_NPC_BOX;   /* free 6-DOF NPC bodies */
    if (P && rigid) {
        es.npc_minv[0] = 1 / d->npc_inertia; es.npc_minv[1] = 1 / d->npc_mass;
        for (int p = 0; p < P; p++) {
            real *rs = root + (A + p) * 13;
            int g = A + p;
            for (int i = 0; i < 3; i++) { es.vel[g][i] = rs[10 + i]; es.vel[g][3 + i] = rs[7 + i]; es.origin[g][i] = rs[i]; }
            es.vel[g][5] += dt * d->gravity_z;
            es.ndof[g] = 6;
        }
    }

    const int seesaw = P && d->npc_kind == MQE_NPC_SEESAW;      /* hinged box on a fixed base: seesaw plank (y) or revolving door (z) */
    const int hz = seesaw && d->npc_geom[13] > 0.5f;            /* hinge axis z (also set for the prismatic disc: no floor-end rows) */
    const int pz = seesaw && d->npc_geom[13] > 1.5f;            /* prismatic y joint: the tug disc (resources/objects/cylinder.urdf) */
    real ss_c = 1, ss_s = 0;
    real h_ex[3] = {1, 0, 0}, h_ey[3] = {0, 1, 0}, h_ez[3] = {0, 0, 1}, h_c[3] = {0, 0, 0};
    const real *ss_fix = NULL;
    if (P && d->npc_kind == MQE_NPC_PLATFORM) ss_fix = root + A * 13;   /* fixed boxes: only their position matters */
    if (seesaw) {   /* resources/objects/seesaw.urdf, rotation_door.urdf: fixed base, box on a passive revolute joint */
        const float *gm = d->npc_geom;
        const real *rs = root + A * 13;
        real th = dof[(12 * A) * 2], thd = dof[(12 * A) * 2 + 1];
        real cx = gm[3], cy = gm[14], cz = gm[15];
        real r2 = hz ? cx * cx + cy * cy : cx * cx + cz * cz;      /* squared distance of the box centre (= COM) from the axis */
        real Ip = d->npc_inertia + d->npc_mass * r2;
        ss_fix = rs;
        ss_c = cos(th); ss_s = sin(th);
        for (int i = 0; i < 3; i++) es.origin[A][i] = rs[i] + gm[i];
        if (pz) es.origin[A][1] += th;                           /* th is the displacement along y */
        if (hz) { v3set(h_ex, ss_c, ss_s, 0); v3set(h_ey, -ss_s, ss_c, 0); v3set(h_ez, 0, 0, 1); }
        else    { v3set(h_ex, ss_c, 0, -ss_s); v3set(h_ey, 0, 1, 0); v3set(h_ez, ss_s, 0, ss_c); }
        for (int i = 0; i < 3; i++) h_c[i] = es.origin[A][i] + cx * h_ex[i] + cy * h_ey[i] + cz * h_ez[i];
        es.npc_minv[0] = 1 / Ip; es.npc_minv[1] = pz ? 1 / d->npc_mass : 0;
        real tau_g = hz ? 0 : gm[3] * ss_c * d->npc_mass * (-d->gravity_z);       /* r_x m g about +y; none about a vertical axis */
        if (pz) { es.vel[A][4] = thd; v3set(h_c, es.origin[A][0], es.origin[A][1], es.origin[A][2] + gm[15]); }
        else es.vel[A][hz ? 2 : 1] = thd + dt * tau_g / Ip;
        es.ndof[A] = 6;
    }

    /* 2. rows: per robot joint limits, then world contacts (probe order); per npc world contacts; then pairs */
    for (int a = 0; a < A; a++) {
        RobotDyn *rd = &es.rd[a];
        int nl = 0;
        for (int j = 0; j < 12 && nl < MAX_LIMIT_ROWS; j++) {
            real glo = q[a][j] - md->q_lower[j], ghi = md->q_upper[j] - q[a][j];
            for (int side = 0; side < 2 && nl < MAX_LIMIT_ROWS; side++) {
                real gap = side ? ghi : glo;
                if (gap >= d->limit_margin) continue;
                Row *r = &rows[nr++];
                memset(r, 0, sizeof *r);
                r->ga = a; r->gb = -1; r->contact = -1;
                r->Ja[6 + j] = side ? -1 : 1;
                for (int i = 0; i < NV; i++) r->Ya[i] = rd->Minv[i][6 + j] * r->Ja[6 + j];
                r->dinv = 1 / (rd->Minv[6 + j][6 + j] + d->cfm);
                r->bias = contact_bias(d, gap);
                r->kind = 0; r->normal_row = nr - 1;
                nl++;
            }
        }
        int nloc = 0;
        for (int pi = 0; pi < md->n_probes && nloc < MAX_LOCAL_CONTACTS; pi++) {
            const float *pr = md->probes[pi];
            int link = (int)pr[0], body = (int)pr[1];
            real loc[3] = {pr[2], pr[3], pr[4]}, x[3];
            m3mulv(x, rd->R[link], loc);
            v3add(x, x, rd->p[link]);
            v3add(x, x, es.origin[a]);
            Contact cand[2];
            int k = probe_world(o, x, pr[5], ss_fix, cand);
            for (int i = 0; i < k && nloc < MAX_LOCAL_CONTACTS; i++) {
                Contact *c = &contacts[nc];
                *c = cand[i];
                c->ga = a; c->la = link; c->rba = a * MQE_NUM_BODIES + body; c->gb = -1; c->lb = 0; c->rbb = -1;
                nr = add_contact_rows(o, &es, c, nc, rows, nr);
                nc++; nloc++;
            }
        }
        stats[0] += nloc; stats[1] += nl;
    }
    real npcR[16][9];
    for (int p = 0; p < P && rigid; p++) {
        int g = A + p, nloc = 0;
        real *rs = root + g * 13;
        quat_to_mat(npcR[g], rs + 3);
        const int box = d->npc_kind == MQE_NPC_BOX;
        int ends = box ? 8 : (d->npc_halflen > 0 ? 2 : 1);             /* box: its 8 corners as point probes */
        for (int en = 0; en < ends && nloc < 4; en++) {
            real loc[3] = {0, 0, (en == 0 ? -1 : 1) * d->npc_halflen}, x[3];
            if (box) for (int i = 0; i < 3; i++) loc[i] = ((en >> i) & 1 ? 1 : -1) * d->npc_geom[4 + i];
            m3mulv(x, npcR[g], loc);
            v3add(x, x, es.origin[g]);
            Contact cand[2];
            int k = probe_world(o, x, box ? 0 : d->npc_radius, NULL, cand);
            for (int i = 0; i < k && nloc < 4; i++) {
                Contact *c = &contacts[nc];
                *c = cand[i];
                c->ga = g; c->la = 0; c->rba = A * MQE_NUM_BODIES + p; c->gb = -1; c->lb = 0; c->rbb = -1;
                nr = add_contact_rows(o, &es, c, nc, rows, nr);
                nc++; nloc++;
            }
        }
        stats[0] += nloc;
    }
    /* pairs: groups X<Y, capsules i of X, j of Y */
    int npair = 0;
    const int max_pair = (d->max_pair_contacts > 0 && d->max_pair_contacts < MAX_PAIR_CONTACTS) ? d->max_pair_contacts : MAX_PAIR_CONTACTS;
    for (int X = 0; X < G; X++)
        for (int Y = X + 1; Y < G; Y++) {
            if (X >= A && d->npc_kind != MQE_NPC_RIGID) continue;
            if (Y >= A && d->npc_kind != MQE_NPC_RIGID) continue;
            real dd[3];
            v3sub(dd, es.origin[X], es.origin[Y]);
            real bx = X < A ? (real)ROBOT_BOUND : d->npc_radius + d->npc_halflen;
            real by = Y < A ? (real)ROBOT_BOUND : d->npc_radius + d->npc_halflen;
            real lim = bx + by + d->contact_offset;
            if (v3dot(dd, dd) > lim * lim) continue;
            int nx = X < A ? md->n_caps : 1, ny = Y < A ? md->n_caps : 1;
            for (int i = 0; i < nx; i++)
                for (int j = 0; j < ny; j++) {
                    real a0[3], a1[3], b0[3], b1[3], ra, rb;
                    int la = 0, lb = 0, rba, rbb;
                    if (X < A) {
                        const float *cp = md->caps[i];
                        la = (int)cp[0]; rba = X * MQE_NUM_BODIES + (int)cp[1]; ra = cp[8];
                        real l0[3] = {cp[2], cp[3], cp[4]}, l1[3] = {cp[5], cp[6], cp[7]};
                        m3mulv(a0, es.rd[X].R[la], l0); v3add(a0, a0, es.rd[X].p[la]); v3add(a0, a0, es.origin[X]);
                        m3mulv(a1, es.rd[X].R[la], l1); v3add(a1, a1, es.rd[X].p[la]); v3add(a1, a1, es.origin[X]);
                    } else {
                        real l0[3] = {0, 0, -d->npc_halflen}, l1[3] = {0, 0, d->npc_halflen};
                        m3mulv(a0, npcR[X], l0); v3add(a0, a0, es.origin[X]);
                        m3mulv(a1, npcR[X], l1); v3add(a1, a1, es.origin[X]);
                        ra = d->npc_radius; rba = A * MQE_NUM_BODIES + (X - A);
                    }
                    if (Y < A) {
                        const float *cp = md->caps[j];
                        lb = (int)cp[0]; rbb = Y * MQE_NUM_BODIES + (int)cp[1]; rb = cp[8];
                        real l0[3] = {cp[2], cp[3], cp[4]}, l1[3] = {cp[5], cp[6], cp[7]};
                        m3mulv(b0, es.rd[Y].R[lb], l0); v3add(b0, b0, es.rd[Y].p[lb]); v3add(b0, b0, es.origin[Y]);
                        m3mulv(b1, es.rd[Y].R[lb], l1); v3add(b1, b1, es.rd[Y].p[lb]); v3add(b1, b1, es.origin[Y]);
                    } else {
                        real l0[3] = {0, 0, -d->npc_halflen}, l1[3] = {0, 0, d->npc_halflen};
                        m3mulv(b0, npcR[Y], l0); v3add(b0, b0, es.origin[Y]);
                        m3mulv(b1, npcR[Y], l1); v3add(b1, b1, es.origin[Y]);
                        rb = d->npc_radius; rbb = A * MQE_NUM_BODIES + (Y - A);
                    }
                    real ca[3], cb[3], dv[3];
                    seg_seg(a0, a1, b0, b1, ca, cb);
                    v3sub(dv, ca, cb);
                    real dist = sqrt(v3dot(dv, dv));
                    real gap = dist - ra - rb;
                    if (gap >= d->contact_offset || dist < (real)1e-9 || npair >= max_pair) continue;
                    Contact *c = &contacts[nc];
                    c->ga = X; c->la = la; c->rba = rba; c->gb = Y; c->lb = lb; c->rbb = rbb;
                    for (int k = 0; k < 3; k++) { c->n[k] = dv[k] / dist; c->pos[k] = cb[k] + c->n[k] * (rb + gap * (real)0.5); }
                    c->gap = gap;
                    nr = add_contact_rows(o, &es, c, nc, rows, nr);
                    nc++; npair++;
                }
        }
    if (seesaw) {
        const float *gm = d->npc_geom;
        /* plank ends resting on the floor (local rows of the seesaw group) */
        int nloc = 0;
        for (int en = 0; en < 2 && !hz; en++) {
            real xe = gm[3] + (en == 0 ? -1 : 1) * gm[4];
            real x[3] = {es.origin[A][0] + xe * ss_c, es.origin[A][1], es.origin[A][2] - xe * ss_s};
            real gap = x[2] - gm[6] - d->floor_z;
            if (gap >= d->contact_offset) continue;
            Contact *c = &contacts[nc];
            v3set(c->n, 0, 0, 1);
            c->gap = gap;
            v3set(c->pos, x[0], x[1], x[2] - gm[6] - gap * (real)0.5);
            c->ga = A; c->la = 0; c->rba = A * MQE_NUM_BODIES; c->gb = -1; c->lb = 0; c->rbb = -1;
            nr = add_contact_rows(o, &es, c, nc, rows, nr);
            nc++; nloc++;
        }
        stats[0] += nloc;
        /* robot probes on the plank: robot X ascending, probe table order */
        {
            real h[3] = {gm[4], gm[5], gm[6]};
            npair = obb_probe_contacts(o, &es, md, h_c, pz ? NULL : h_ex, h_ey, h_ez, h, contacts, &nc, rows, &nr, npair, max_pair);
        }
    }
    if (P && d->npc_kind == MQE_NPC_BOX) {   /* robot probes on the push box (resources/objects/box.urdf) */
        const real *R = npcR[A];
        real ex[3] = {R[0], R[3], R[6]}, ey[3] = {R[1], R[4], R[7]}, ez[3] = {R[2], R[5], R[8]};
        real h[3] = {d->npc_geom[4], d->npc_geom[5], d->npc_geom[6]};
        npair = obb_probe_contacts(o, &es, md, es.origin[A], ex, ey, ez, h, contacts, &nc, rows, &nr, npair, max_pair);
    }
    stats[2] += npair;
    if (nr > stats[3]) stats[3] = nr;

    /* 3. projected Gauss-Seidel sweeps */
    for (int it = 0; it < d->solver_iters; it++)
        for (int i = 0; i < nr; i++) solve_row(&rows[i], rows, &es, o->mu_env ? (real)o->mu_env[e] : (real)d->friction);

    /* 4. contact force report (impulse / dt, world frame) */
    real *cf = o->contact + (size_t)e * o->NB * 3;
    memset(cf, 0, (size_t)o->NB * 3 * sizeof(real));
    for (int i = 0; i < nr; i++) {
        Row *r = &rows[i];
        if (r->contact < 0) continue;
        const Contact *c = &contacts[r->contact];
        for (int k = 0; k < 3; k++) {
            cf[c->rba * 3 + k] += r->dir[k] * r->lambda / dt;
            if (c->rbb >= 0) cf[c->rbb * 3 + k] -= r->dir[k] * r->lambda / dt;
        }
    }

    if (seesaw) {
        real lim = d->npc_geom[12], thd = es.vel[A][pz ? 4 : (hz ? 2 : 1)];
        thd = thd > lim ? lim : (thd < -lim ? -lim : thd);         /* URDF joint velocity limit */
        dof[(12 * A) * 2] += dt * thd;
        dof[(12 * A) * 2 + 1] = thd;
    }
    /* 5. integrate (semi-implicit Euler; exponential map for orientation) */
    for (int g = 0; g < G; g++) {
        if (g >= A && !rigid) continue;
        real *rs = root + g * 13;
        real *v = es.vel[g];
        if (g < A)
            for (int j = 0; j < 12; j++) {
                real lim = md->qd_limit[j];
                if (v[6 + j] > lim) v[6 + j] = lim;
                if (v[6 + j] < -lim) v[6 + j] = -lim;
                dof[(12 * g + j) * 2] += dt * v[6 + j];
                dof[(12 * g + j) * 2 + 1] = v[6 + j];
            }
        for (int i = 0; i < 3; i++) { rs[i] += dt * v[3 + i]; rs[7 + i] = v[3 + i]; rs[10 + i] = v[i]; }
        real wn = sqrt(v3dot(v, v)), th = wn * dt, dq[4];
        if (th < (real)1e-8) { dq[0] = v[0] * dt * (real)0.5; dq[1] = v[1] * dt * (real)0.5; dq[2] = v[2] * dt * (real)0.5; dq[3] = 1; }
        else { real s = sin(th * (real)0.5) / wn; dq[0] = v[0] * s; dq[1] = v[1] * s; dq[2] = v[2] * s; dq[3] = cos(th * (real)0.5); }
        real *Q = rs + 3, x = Q[0], y = Q[1], z = Q[2], w = Q[3], nq[4];
        nq[0] = dq[3] * x + dq[0] * w + dq[1] * z - dq[2] * y;
        nq[1] = dq[3] * y - dq[0] * z + dq[1] * w + dq[2] * x;
        nq[2] = dq[3] * z + dq[0] * y - dq[1] * x + dq[2] * w;
        nq[3] = dq[3] * w - dq[0] * x - dq[1] * y - dq[2] * z;
        real inv = 1 / sqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
        for (int i = 0; i < 4; i++) Q[i] = nq[i] * inv;
    }
}

/* ------------------------------------------------------------------------------------------------ _compute_torques (go1.py:315-354) */
static void env_torques(Oracle *o, int e, real *tau, uint32_t call) {
    const MqeSimDesc *d = &o->d;
    int A = o->A;
    real *dof = o->dof + (size_t)e * (12 * A + o->D) * 2;
    for (int a = 0; a < A; a++) {
        int m = e * A + a;
        for (int j = 0; j < 12; j++) {
            real act = o->actions[m * 12 + j] * d->action_scale;
            if (d->control_type != 0 && o->motor_strength) act *= o->motor_strength[m * 12 + j];     /* legged_robot_field.py:180-183 */
            if (d->control_type != 0) {   /* LeggedRobot._compute_torques (legged_robot.py:384-392): 'P' / 'V' / 'T', no hip scale, no histories */
                real qj = dof[(12 * a + j) * 2], qdj = dof[(12 * a + j) * 2 + 1];
                real t = act;
                if (d->control_type == 1) t = d->stiffness * (act + d->model.q_default[j] - qj) - d->damping * qdj;
                else if (d->control_type == 3) t = d->stiffness * (act - qdj) - d->damping * (qdj - o->last_dof_vel[m * 12 + j]) / d->sim_dt;
                real lim = d->model.tau_limit[j];
                t = t > lim ? lim : (t < -lim ? -lim : t);
                tau[12 * a + j] = t;
                o->torques[m * 12 + j] = t;
                continue;
            }
            if (j % 3 == 0) act *= d->hip_scale;
            if (o->lag_ring) {   /* lag_buffer = lag_buffer[1:] + [actions_scaled]; target = lag_buffer[0] (go1.py:337-339), as a ring */
                real *ring = o->lag_ring + (size_t)m * o->lag_n * 12 + j;
                ring[(call % o->lag_n) * 12] = act;
                act = ring[((call + 1) % o->lag_n) * 12];
            }
            real target = act + d->model.q_default[j];
            real err = dof[(12 * a + j) * 2] - target, vel = dof[(12 * a + j) * 2 + 1];
            real x[6] = {err, o->err1[m * 12 + j], o->err2[m * 12 + j], vel, o->vel1[m * 12 + j], o->vel2[m * 12 + j]};
            real t = actuator_forward(&d->weights, x);
            o->err2[m * 12 + j] = o->err1[m * 12 + j]; o->err1[m * 12 + j] = err;
            o->vel2[m * 12 + j] = o->vel1[m * 12 + j]; o->vel1[m * 12 + j] = vel;
            real lim = d->model.tau_limit[j];
            t = t > lim ? lim : (t < -lim ? -lim : t);
            tau[12 * a + j] = t;
            o->torques[m * 12 + j] = t;
        }
    }
}

/* ------------------------------------------------------------------------------------------------ compute_observations (go1.py:153-196) */
static void env_observations(Oracle *o, int e) {
    const MqeSimDesc *d = &o->d;
    int A = o->A, G = o->A + o->P;
    for (int a = 0; a < A; a++) {
        int m = e * A + a;
        real *ob = o->obs + (size_t)m * MQE_OBS_FLOATS;
        const real *rs = o->root + ((size_t)e * G + a) * 13;
        const real *dof = o->dof + ((size_t)e * (12 * A + o->D) + 12 * a) * 2;
        const real *bq = d->quat_alias ? rs + 3 : o->base_quat + m * 4;
        for (int i = 0; i < 3; i++) ob[MQE_OBS_BASE_POS + i] = rs[i] - o->env_origins[e * 3 + i];
        for (int i = 0; i < 4; i++) ob[MQE_OBS_BASE_QUAT + i] = bq[i];
        for (int j = 0; j < 12; j++) {
            ob[MQE_OBS_DOF_POS + j] = dof[j * 2] - d->model.q_default[j];
            ob[MQE_OBS_DOF_VEL + j] = dof[j * 2 + 1] * (real)0.05;
            ob[MQE_OBS_LAST_ACTION + j] = o->actions[m * 12 + j];
            ob[MQE_OBS_LAST_LAST_ACTION + j] = o->last_actions[m * 12 + j];
        }
        for (int i = 0; i < 3; i++) {
            ob[MQE_OBS_LIN_VEL + i] = o->base_lin_vel[m * 3 + i] * (real)2.0;
            ob[MQE_OBS_ANG_VEL + i] = o->base_ang_vel[m * 3 + i] * (real)0.25;
            ob[MQE_OBS_PROJ_GRAVITY + i] = o->proj_grav[m * 3 + i];
        }
        for (int i = 0; i < 4; i++) ob[MQE_OBS_CLOCK + i] = o->clock[m * 4 + i];
        get_euler_xyz(ob + MQE_OBS_BASE_RPY, bq);
    }
}

/* ------------------------------------------------------------------------------------------------ reset_idx for one env
 * go1.py:110-145, legged_robot.py:394-470, 647-652 */
static void env_reset(Oracle *o, int e) {
    const MqeSimDesc *d = &o->d;
    int A = o->A, P = o->P, G = A + P;
    uint32_t ge = (uint32_t)(e + d->env_id_offset), ep = o->episode[e];
    real *dof = o->dof + (size_t)e * (12 * A + o->D) * 2;
    for (int a = 0; a < A && !o->reset_state; a++) {      /* golden replays: the RNG-free part only (_reset_buffers) */
        int m = e * A + a;
        for (int j = 0; j < 12; j++) { o->last_actions[m * 12 + j] = 0; o->last_dof_vel[m * 12 + j] = 0; }
        o->gait[m] = 0;
        memset(o->hist + (size_t)m * 2100, 0, 2100 * sizeof(real));
    }
    if (!o->reset_state) { o->ep_len[e] = 0; o->reset_buf[e] = 1; o->episode[e] = ep + 1; return; }
    for (int a = 0; a < A; a++) {
        int m = e * A + a;
        for (int j = 0; j < 12; j++) {
            real u = rng_uniform(d->seed, ge, ep, RNG_DOF, a * 12 + j);
            dof[(12 * a + j) * 2] = d->model.q_default[j] * (d->dof_ratio_lo + (d->dof_ratio_hi - d->dof_ratio_lo) * u);
            dof[(12 * a + j) * 2 + 1] = 0;
        }
        real *rs = o->root + ((size_t)e * G + a) * 13;
        for (int i = 0; i < 13; i++) rs[i] = o->base_init[m * 13 + i];
        for (int i = 0; i < 3; i++) rs[i] += o->agent_origins[m * 3 + i];
        if (d->has_base_pos_range) {
            rs[0] += d->base_pos_x[0] + (d->base_pos_x[1] - d->base_pos_x[0]) * rng_uniform(d->seed, ge, ep, RNG_BASE_POS, a * 2);
            rs[1] += d->base_pos_y[0] + (d->base_pos_y[1] - d->base_pos_y[0]) * rng_uniform(d->seed, ge, ep, RNG_BASE_POS, a * 2 + 1);
        }
        for (int i = 0; i < 6; i++)
            rs[7 + i] = d->base_vel_lo + (d->base_vel_hi - d->base_vel_lo) * rng_uniform(d->seed, ge, ep, RNG_BASE_VEL, a * 6 + i);
        /* _reset_buffers */
        for (int j = 0; j < 12; j++) { o->last_actions[m * 12 + j] = 0; o->last_dof_vel[m * 12 + j] = 0; }
        o->gait[m] = 0;
        memset(o->hist + (size_t)m * 2100, 0, 2100 * sizeof(real));
    }
    for (int k = 0; k < o->D; k++) { dof[(12 * A + k) * 2] = o->npc_dof_default[k]; dof[(12 * A + k) * 2 + 1] = 0; }
    for (int p = 0; p < P; p++) {
        real *rs = o->root + ((size_t)e * G + A + p) * 13;
        for (int i = 0; i < 13; i++) rs[i] = o->npc_init[(e * P + p) * 13 + i];
        for (int i = 0; i < 3; i++) rs[i] += o->env_origins[e * 3 + i];
        if (d->has_npc_pos_range) {
            rs[0] += d->npc_pos_x[0] + (d->npc_pos_x[1] - d->npc_pos_x[0]) * rng_uniform(d->seed, ge, ep, RNG_NPC_POS, p * 2);
            rs[1] += d->npc_pos_y[0] + (d->npc_pos_y[1] - d->npc_pos_y[0]) * rng_uniform(d->seed, ge, ep, RNG_NPC_POS, p * 2 + 1);
        }
        if (d->has_npc_rpy_range) {
            real r = d->npc_rpy_r[0] + (d->npc_rpy_r[1] - d->npc_rpy_r[0]) * rng_uniform(d->seed, ge, ep, RNG_NPC_RPY, p * 3);
            real pp = d->npc_rpy_p[0] + (d->npc_rpy_p[1] - d->npc_rpy_p[0]) * rng_uniform(d->seed, ge, ep, RNG_NPC_RPY, p * 3 + 1);
            real y = d->npc_rpy_y[0] + (d->npc_rpy_y[1] - d->npc_rpy_y[0]) * rng_uniform(d->seed, ge, ep, RNG_NPC_RPY, p * 3 + 2);
            quat_from_euler_xyz(rs + 3, r, pp, y);
        }
    }
    o->ep_len[e] = 0;
    o->reset_buf[e] = 1;
    o->episode[e] = ep + 1;
}

/* Go1.reset (go1.py:147-151) */
void orc_reset(Oracle *o) {
    for (int e = 0; e < o->N; e++) env_reset(o, e);
    for (int e = 0; e < o->N; e++) env_observations(o, e);
}

/* defender command (go1_football_defender.py:56-80), from the observations of the previous step */
static void defender_command(const Oracle *o, int e, real *cmd) {
    const MqeSimDesc *d = &o->d;
    int A = o->A, G = o->A + o->P;
    const real *dp = o->root + ((size_t)e * G + 2) * 13;   /* agent index 2 */
    const real *bp = o->root + ((size_t)e * G + A) * 13;   /* ball          */
    real gate[3] = {o->env_origins[e * 3] + d->gate_x, o->env_origins[e * 3 + 1], o->env_origins[e * 3 + 2]};
    real tx = (real)0.6 * bp[0] + (real)0.4 * gate[0], ty = (real)0.6 * bp[1] + (real)0.4 * gate[1];
    real yaw = o->obs[(size_t)(e * A + 2) * MQE_OBS_FLOATS + MQE_OBS_BASE_RPY + 2];
    real yaw_to_gate = PI_R + atan((gate[1] - dp[1]) / (gate[0] - dp[0]));
    real yc = yaw_to_gate - yaw;
    yc = (yc < (real)-0.3 ? (real)-0.3 : (yc > (real)0.3 ? (real)0.3 : yc)) / (real)0.3;
    real tdg = sqrt((tx - gate[0]) * (tx - gate[0]) + (ty - gate[1]) * (ty - gate[1]));
    real ddg = sqrt((dp[0] - gate[0]) * (dp[0] - gate[0]) + (dp[1] - gate[1]) * (dp[1] - gate[1]));
    real xc = tdg - ddg;
    xc = xc < (real)-0.5 ? (real)-0.5 : (xc > (real)0.5 ? (real)0.5 : xc);
    real yy = gate[1] + (ty - gate[1]) * (dp[0] - gate[0]) / (tx - gate[0]) - dp[1];
    yy = yy < (real)-0.5 ? (real)-0.5 : (yy > (real)0.5 ? (real)0.5 : yy);
    cmd[0] = xc; cmd[1] = -yy; cmd[2] = yc;
}

/* ------------------------------------------------------------------------------------------------ step phases */
/* wrapper scaling + Go1.step clip + preprocess_action (wrappers step(), go1.py:35-41, 64-108) */
void orc_policy(Oracle *o, const float *actions /* [N][A_ctrl][3] */) {
    const MqeSimDesc *d = &o->d;
    int A = o->A, actrl = d->defender ? A - 1 : A;
#pragma omp parallel for schedule(static)
    for (int m = 0; m < o->M; m++) {
        int e = m / A, a = m % A;
        real cmd[3];
        if (a < actrl) {
            for (int i = 0; i < 3; i++) {
                real x = actions[((size_t)e * actrl + a) * 3 + i];
                x = x < -1 ? -1 : (x > 1 ? 1 : x);          /* wrapper clip */
                x *= d->act_scale[i];
                cmd[i] = x;
            }
        } else defender_command(o, e, cmd);
        if (!d->defender) for (int i = 0; i < 3; i++) cmd[i] = cmd[i] < -1 ? -1 : (cmd[i] > 1 ? 1 : cmd[i]); /* go1.py:38 */
        real *lo = o->loc_obs + (size_t)m * 70;
        const real *ob = o->obs + (size_t)m * MQE_OBS_FLOATS;
        for (int i = 0; i < 3; i++) { o->commands[m * 3 + i] = cmd[i]; if (d->command_vel) lo[3 + i] = cmd[i] * d->cmd_scale[i]; }
        for (int i = 0; i < 3; i++) lo[i] = ob[MQE_OBS_PROJ_GRAVITY + i];
        for (int j = 0; j < 12; j++) {
            lo[18 + j] = ob[MQE_OBS_DOF_POS + j];
            lo[30 + j] = ob[MQE_OBS_DOF_VEL + j];
            lo[42 + j] = o->loc_last[m * 12 + j];
            lo[54 + j] = o->loc_last2[m * 12 + j];
        }
        for (int i = 0; i < 4; i++) lo[66 + i] = ob[MQE_OBS_CLOCK + i];
        real *h = o->hist + (size_t)m * 2100;
        memmove(h, h + 70, 2030 * sizeof(real));             /* cat(history[:, 70:], obs) */
        memcpy(h + 2030, lo, 70 * sizeof(real));
        real latent[2], act[12];
        policy_forward(&d->weights, h, latent, act);
        for (int j = 0; j < 12; j++) {
            o->loc_last2[m * 12 + j] = o->loc_last[m * 12 + j];
            o->loc_last[m * 12 + j] = act[j];
            real c = act[j];
            c = c > d->clip_actions ? d->clip_actions : (c < -d->clip_actions ? -d->clip_actions : c);
            o->actions[m * 12 + j] = c;
        }
    }
}

void orc_substeps(Oracle *o, int count) {
    int32_t tot[8] = {0};
#pragma omp parallel for schedule(dynamic, 8)
    for (int e = 0; e < o->N; e++) {
        int32_t st[8] = {0};
        real tau[48];
        for (int s = 0; s < count; s++) {
            env_torques(o, e, tau, o->lag_calls + (uint32_t)s);
            env_substep(o, e, tau, st);
            if (s < o->d.decimation) {   /* post_decimation_step (legged_robot.py:112-115) */
                const real soft = o->d.soft_dof_pos_limit > 0 ? o->d.soft_dof_pos_limit : 1;
                const real *dof = o->dof + (size_t)e * (12 * o->A + o->D) * 2;
                for (int k = 0; k < 12 * o->A; k++) {
                    size_t ix = ((size_t)e * o->d.decimation + s) * (12 * o->A) + k;
                    int j = k % 12;
                    real mid = (real)0.5 * ((real)o->d.model.q_lower[j] + (real)o->d.model.q_upper[j]);
                    real half = (real)0.5 * ((real)o->d.model.q_upper[j] - (real)o->d.model.q_lower[j]) * soft;
                    o->sub_tau[ix] = tau[k]; o->sub_qd[ix] = dof[k * 2 + 1];
                    o->sub_exceed[ix] = (uint8_t)((dof[k * 2] < mid - half) | (dof[k * 2] > mid + half));
                }
            }
        }
#pragma omp critical
        { tot[0] += st[0]; tot[1] += st[1]; tot[2] += st[2]; if (st[3] > tot[3]) tot[3] = st[3]; }
    }
    memcpy(o->stats, tot, sizeof tot);
    o->lag_calls += (uint32_t)count;
}

/* _step_contact_targets (go1.py:240-279) */
static void gait_clock(Oracle *o, int m, real dt_policy) {
    const real *lo = o->loc_obs + (size_t)m * 70;
    real freq = lo[7], phase = lo[8], offset = lo[9], bound = lo[10], dur = lo[11];
    real g = fmod(o->gait[m] + dt_policy * freq, (real)1.0);
    if (g < 0) g += 1;
    o->gait[m] = g;
    real fi[4] = {g + phase + offset + bound, g + offset, g + bound, g + phase};
    for (int i = 0; i < 4; i++) {
        real r = fmod(fi[i], (real)1.0);
        if (r < 0) r += 1;
        real x = fi[i];
        if (r < dur) x = r * ((real)0.5 / dur);
        else if (r > dur) x = (real)0.5 + (r - dur) * ((real)0.5 / (1 - dur));
        o->clock[m * 4 + i] = sin(2 * PI_R * x);
    }
}

/* Go1Sheep._step_npc (go1_sheep.py:35-64) */
static void sheep_step(Oracle *o, int e) {
    const MqeSimDesc *d = &o->d;
    int A = o->A, P = o->P, G = A + P;
    real *root = o->root + (size_t)e * G * 13;
    real avg[3] = {0, 0, 0}, var[2] = {0, 0};
    for (int p = 0; p < P; p++) for (int i = 0; i < 3; i++) avg[i] += root[(A + p) * 13 + i] / P;
    for (int p = 0; p < P; p++) for (int i = 0; i < 2; i++) { real t = root[(A + p) * 13 + i] - avg[i]; var[i] += t * t / P; }
    o->sheep_stats[e * 3] = avg[0]; o->sheep_stats[e * 3 + 1] = avg[1]; o->sheep_stats[e * 3 + 2] = var[0] + var[1];
    uint32_t ge = (uint32_t)(e + d->env_id_offset);
    for (int p = 0; p < P; p++) {
        real *rs = root + (A + p) * 13, dv[3];
        for (int i = 0; i < 3; i++) dv[i] = d->sheep_randomness * rng_normal(d->seed, ge, o->step_count, RNG_SHEEP, p * 3 + i) * 2;
        if (P != 1) {
            real rel[3] = {avg[0] - rs[0], avg[1] - rs[1], avg[2] - rs[2]};
            real nn = sqrt(v3dot(rel, rel));
            for (int i = 0; i < 3; i++) dv[i] += d->sheep_randomness * rel[i] / nn / (real)1.5;
        }
        for (int a = 0; a < A; a++) {
            real rel[3] = {rs[0] - root[a * 13], rs[1] - root[a * 13 + 1], rs[2] - root[a * 13 + 2]};
            real sq[3] = {rel[0] * rel[0], rel[1] * rel[1], rel[2] * rel[2]};
            real dis = sqrt(v3dot(sq, sq));                 /* torch.norm(relative_pos ** 2) */
            if (dis > 9) continue;
            real den = pow(dis, (real)1.4);
            for (int i = 0; i < 3; i++) dv[i] += d->sheep_scale * rel[i] / den;
        }
        dv[2] = 0;
        for (int i = 0; i < 3; i++) rs[7 + i] += dv[i];
        for (int i = 0; i < 2; i++) rs[7 + i] = rs[7 + i] < -2 ? -2 : (rs[7 + i] > 2 ? 2 : rs[7 + i]);
        rs[2] = rs[2] < 0 ? 0 : (rs[2] > (real)0.3 ? (real)0.3 : rs[2]);
        rs[3] = 0; rs[4] = 0;
    }
}

/* post_physics_step (legged_robot_field.py:117-119 -> legged_robot.py:117-157) */
void orc_post_physics(Oracle *o) {
    const MqeSimDesc *d = &o->d;
    int A = o->A, P = o->P, G = A + P;
    real dt_policy = d->sim_dt * d->decimation;
    real grav[3] = {0, 0, -1};
    for (int e = 0; e < o->N; e++) {
        o->ep_len[e] += 1;
        int collide = 0, rt = 0, pt = 0, zl = 0, zh = 0;
        for (int a = 0; a < A; a++) {
            int m = e * A + a;
            const real *rs = o->root + ((size_t)e * G + a) * 13;
            for (int i = 0; i < 4; i++) o->base_quat[m * 4 + i] = rs[3 + i];
            quat_rotate_inverse(o->base_lin_vel + m * 3, rs + 3, rs + 7);
            quat_rotate_inverse(o->base_ang_vel + m * 3, rs + 3, rs + 10);
            quat_rotate_inverse(o->proj_grav + m * 3, rs + 3, grav);
            if (d->control_type == 0) gait_clock(o, m, dt_policy);   /* _step_contact_targets: control_type 'C' only (go1.py:241) */
            /* _push_robots (go1.py:237-238, legged_robot.py:472-477) */
            if (d->push_interval > 0 && ((o->step_count + 1u) % (uint32_t)d->push_interval) == 0u) {
                real *rw = o->root + ((size_t)e * G + a) * 13;
                for (int k = 0; k < 2; k++)
                    rw[7 + k] = (2 * (real)rng_uniform(d->seed, (uint32_t)(d->env_id_offset + e), o->step_count, RNG_PUSH, 2 * a + k) - 1) * d->max_push_vel_xy;
            }
            /* check_termination: legged_robot.py:159-169 + legged_robot_field.py:121-146 */
            const real *cf = o->contact + ((size_t)e * o->NB + a * MQE_NUM_BODIES) * 3; /* base = body 0 */
            if (sqrt(v3dot(cf, cf)) > 1) collide = 1;
            real rpy[3];
            get_euler_xyz(rpy, rs + 3);
            if (rpy[0] > PI_R) rpy[0] -= 2 * PI_R;
            if (rpy[1] > PI_R) rpy[1] -= 2 * PI_R;
            real z = rs[2] - o->agent_origins[m * 3 + 2];
            if (fabs(rpy[0]) > d->term_roll) rt = 1;
            if (fabs(rpy[1]) > d->term_pitch) pt = 1;
            if (z < d->term_zlow) zl = 1;
            if (z > d->term_zhigh) zh = 1;
        }
        int reset = 0;
        if (d->term_mask & 16) { o->collide_buf[e] = (uint8_t)collide; reset |= collide; }
        o->timeout_buf[e] = o->ep_len[e] > d->max_episode_length;
        reset |= o->timeout_buf[e];
        if (d->term_mask & 1) { o->r_term[e] = (uint8_t)rt; reset |= rt; }
        if (d->term_mask & 2) { o->p_term[e] = (uint8_t)pt; reset |= pt; }
        if (d->term_mask & 4) { o->zl_term[e] = (uint8_t)zl; reset |= zl; }
        if (d->term_mask & 8) { o->zh_term[e] = (uint8_t)zh; reset |= zh; }
        o->reset_buf[e] = (uint8_t)reset;
        /* legged_robot.py:164-169: `self.reset_buf = self.collide_buf` binds ONE tensor to both names and every later
         * `reset_buf |= ...` is in place, so the reference's collide_buf ends up equal to the full reset mask */
        if (d->term_mask & 16) o->collide_buf[e] = (uint8_t)reset;
        if (P && d->npc_ctrl == MQE_NPC_SHEEP) sheep_step(o, e);
        if (reset) env_reset(o, e);
        env_observations(o, e);
        for (int a = 0; a < A; a++) {
            int m = e * A + a;
            const real *rs = o->root + ((size_t)e * G + a) * 13;
            const real *dof = o->dof + ((size_t)e * (12 * A + o->D) + 12 * a) * 2;
            for (int j = 0; j < 12; j++) { o->last_actions[m * 12 + j] = o->actions[m * 12 + j]; o->last_dof_vel[m * 12 + j] = dof[j * 2 + 1]; }
            for (int i = 0; i < 6; i++) o->last_root_vel[m * 6 + i] = rs[7 + i];
        }
    }
    o->step_count++;
}

void orc_step(Oracle *o, const float *actions) {
    orc_policy(o, actions);
    orc_substeps(o, o->d.decimation);
    orc_post_physics(o);
}
/* Go1.step() for control_type 'P' / 'V' / 'T' (go1.py:43-45 -> pre_physics_step, legged_robot.py:108-110): [N][12A] joint actions */
void orc_step_joint(Oracle *o, const float *joint_actions) {
    const real c = o->d.clip_actions;
    for (int i = 0; i < o->M * 12; i++) { real a = joint_actions[i]; o->actions[i] = a > c ? c : (a < -c ? -c : a); }
    orc_substeps(o, o->d.decimation);
    orc_post_physics(o);
}

/* ------------------------------------------------------------------------------------------------ accessors */
static void to_f32(float *dst, const real *src, size_t n) { for (size_t i = 0; i < n; i++) dst[i] = (float)src[i]; }
static void from_f32(real *dst, const float *src, size_t n) { for (size_t i = 0; i < n; i++) dst[i] = src[i]; }

/* returns element count; copies as float32 (or raw for integer buffers) when out != NULL */
int64_t orc_get(Oracle *o, int which, void *out) {
    int N = o->N, A = o->A, P = o->P, M = o->M;
    const real *src = NULL; size_t n = 0;
    switch (which) {
    case MQE_BUF_ROOT_STATES: src = o->root; n = (size_t)N * (A + P) * 13; break;
    case MQE_BUF_DOF_STATES: src = o->dof; n = (size_t)N * (12 * A + o->D) * 2; break;
    case MQE_BUF_CONTACT_FORCES: src = o->contact; n = (size_t)N * o->NB * 3; break;
    case MQE_BUF_TORQUES: src = o->torques; n = (size_t)M * 12; break;
    case MQE_BUF_ACTIONS: src = o->actions; n = (size_t)M * 12; break;
    case MQE_BUF_LAST_ACTIONS: src = o->last_actions; n = (size_t)M * 12; break;
    case MQE_BUF_OBS: src = o->obs; n = (size_t)M * MQE_OBS_FLOATS; break;
    case MQE_BUF_BASE_LIN_VEL: src = o->base_lin_vel; n = (size_t)M * 3; break;
    case MQE_BUF_BASE_ANG_VEL: src = o->base_ang_vel; n = (size_t)M * 3; break;
    case MQE_BUF_PROJ_GRAVITY: src = o->proj_grav; n = (size_t)M * 3; break;
    case MQE_BUF_COMMANDS: src = o->commands; n = (size_t)M * 3; break;
    case MQE_BUF_LOC_OBS: src = o->loc_obs; n = (size_t)M * 70; break;
    case MQE_BUF_LOC_ACTION: src = o->loc_last; n = (size_t)M * 12; break;
    case MQE_BUF_GAIT: src = o->gait; n = (size_t)M; break;
    case MQE_BUF_CLOCK: src = o->clock; n = (size_t)M * 4; break;
    case MQE_BUF_HISTORY: src = o->hist; n = (size_t)M * 2100; break;
    case MQE_BUF_SHEEP_STATS: src = o->sheep_stats; n = (size_t)N * 3; break;
    case MQE_BUF_RESET: if (out) memcpy(out, o->reset_buf, N); return N;
    case MQE_BUF_TIMEOUT: if (out) memcpy(out, o->timeout_buf, N); return N;
    case MQE_BUF_COLLIDE: if (out) memcpy(out, o->collide_buf, N); return N;
    case MQE_BUF_ROLL_TERM: if (out) memcpy(out, o->r_term, N); return N;
    case MQE_BUF_PITCH_TERM: if (out) memcpy(out, o->p_term, N); return N;
    case MQE_BUF_ZLOW_TERM: if (out) memcpy(out, o->zl_term, N); return N;
    case MQE_BUF_ZHIGH_TERM: if (out) memcpy(out, o->zh_term, N); return N;
    case MQE_BUF_EPISODE_LENGTH: if (out) memcpy(out, o->ep_len, N * sizeof(int64_t)); return N;
    case MQE_BUF_STATS: if (out) memcpy(out, o->stats, sizeof o->stats); return 8;
    case MQE_BUF_SUBSTEP_TORQUES: src = o->sub_tau; n = (size_t)M * 12 * o->d.decimation; break;
    case MQE_BUF_SUBSTEP_DOF_VEL: src = o->sub_qd; n = (size_t)M * 12 * o->d.decimation; break;
    case MQE_BUF_SUBSTEP_EXCEED: if (out) memcpy(out, o->sub_exceed, (size_t)M * 12 * o->d.decimation); return (int64_t)M * 12 * o->d.decimation;
    default: return -1;
    }
    if (out) to_f32((float *)out, src, n);
    return (int64_t)n;
}

/* overwrite a float state buffer (tests inject states) */
int64_t orc_set(Oracle *o, int which, const float *in) {
    int N = o->N, A = o->A, P = o->P, M = o->M;
    switch (which) {
    case MQE_BUF_ROOT_STATES: from_f32(o->root, in, (size_t)N * (A + P) * 13); return 0;
    case MQE_BUF_DOF_STATES: from_f32(o->dof, in, (size_t)N * (12 * A + o->D) * 2); return 0;
    case MQE_BUF_ACTIONS: from_f32(o->actions, in, (size_t)M * 12); return 0;
    case MQE_BUF_HISTORY: from_f32(o->hist, in, (size_t)M * 2100); return 0;
    case MQE_BUF_CONTACT_FORCES: from_f32(o->contact, in, (size_t)N * o->NB * 3); return 0;
    case MQE_BUF_EPISODE_LENGTH: for (int e = 0; e < N; e++) o->ep_len[e] = (int64_t)in[e]; return 0;
    default: return -1;
    }
}

int orc_real_size(void) { return (int)sizeof(real); }

/* ---- test hooks for the golden replays (tests/test_oracle_bookkeeping.py) ---- */
void orc_set_reset_state(Oracle *o, int on) { o->reset_state = on; }
/* _compute_torques for every env, no physics (go1.py:315-354) */
void orc_torques(Oracle *o) {
    for (int e = 0; e < o->N; e++) { real tau[48]; env_torques(o, e, tau, o->lag_calls); }
    o->lag_calls += 1;
}
/* base_quat <- root quaternion, then compute_observations: the state Go1.reset() leaves behind */
void orc_observe(Oracle *o) {
    int A = o->A, G = o->A + o->P;
    for (int e = 0; e < o->N; e++) {
        for (int a = 0; a < A; a++)
            for (int i = 0; i < 4; i++) o->base_quat[(e * A + a) * 4 + i] = o->root[((size_t)e * G + a) * 13 + 3 + i];
        env_observations(o, e);
    }
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU-baseline legs ask for the host's cores explicitly */
void orc_set_threads(int n) {
#if defined(_OPENMP)
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------ unit entry points */
void orc_policy_forward(const MqeWeights *w, const float *hist, int rows, float *latent, float *action) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; r++) {
        real h[2100], l[2], a[12];
        for (int i = 0; i < 2100; i++) h[i] = hist[(size_t)r * 2100 + i];
        policy_forward(w, h, l, a);
        if (latent) { latent[r * 2] = (float)l[0]; latent[r * 2 + 1] = (float)l[1]; }
        for (int j = 0; j < 12; j++) action[r * 12 + j] = (float)a[j];
    }
}
void orc_actuator_forward(const MqeWeights *w, const float *x, int rows, float *torque) {
    for (int r = 0; r < rows; r++) {
        real xi[6];
        for (int i = 0; i < 6; i++) xi[i] = x[r * 6 + i];
        torque[r] = (float)actuator_forward(w, xi);
    }
}
/* mass matrix, bias and unconstrained acceleration of one free robot: quat xyzw, q[12], v[18] = (w, vlin, qd), tau[12] */
void orc_robot_dynamics(const MqeRobotModel *md, real gz, const double *quat, const double *q, const double *v, const double *tau,
                        double *M_out /* 18x18 */, double *c_out /* 18 */, double *acc_out /* 18, spatial */) {
    RobotDyn rd;
    real qq[4], qj[12], vv[NV];
    for (int i = 0; i < 4; i++) qq[i] = quat[i];
    for (int i = 0; i < 12; i++) qj[i] = q[i];
    for (int i = 0; i < NV; i++) vv[i] = v[i];
    robot_kinematics(&rd, md, qq, qj);
    robot_dynamics(&rd, vv, gz);
    spd_inverse(NV, &rd.M[0][0], &rd.Minv[0][0]);
    for (int i = 0; i < NV; i++) {
        c_out[i] = rd.c[i];
        for (int j = 0; j < NV; j++) M_out[i * NV + j] = rd.M[i][j];
        real s = 0;
        for (int k = 0; k < NV; k++) s += rd.Minv[i][k] * ((k >= 6 ? tau[k - 6] : 0) - rd.c[k]);
        acc_out[i] = s;
    }
}
/* forward kinematics: world positions (rel. base origin) of the 13 link frames and the 4 feet */
void orc_robot_fk(const MqeRobotModel *md, const double *quat, const double *q, double *link_pos /* 13x3 */, double *foot_pos /* 4x3 */) {
    RobotDyn rd;
    real qq[4], qj[12];
    for (int i = 0; i < 4; i++) qq[i] = quat[i];
    for (int i = 0; i < 12; i++) qj[i] = q[i];
    robot_kinematics(&rd, md, qq, qj);
    for (int b = 0; b < 13; b++) for (int i = 0; i < 3; i++) link_pos[b * 3 + i] = rd.p[b][i];
    for (int l = 0; l < 4; l++) {
        real off[3] = {md->leg_offsets[l][3][0], md->leg_offsets[l][3][1], md->leg_offsets[l][3][2]}, t[3];
        m3mulv(t, rd.R[3 + 3 * l], off);
        for (int i = 0; i < 3; i++) foot_pos[l * 3 + i] = rd.p[3 + 3 * l][i] + t[i];
    }
}
