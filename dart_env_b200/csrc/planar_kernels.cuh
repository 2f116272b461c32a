// planar_kernels.cuh — hand-written sm_100a device code for one DART time step of a planar
// skeleton, one world per thread, the whole world state in registers.
//
// What it computes is DART's World::step as restated by oracle/dart_oracle.c:
//   K1 forward kinematics            (BodyNode::updateTransform/updateVelocity/updatePartialAcceleration)
//   K2 ABA backward pass             (updateArtInertia [implicit damping/spring] + updateBiasForce)
//   K3 ABA forward pass + dq += dt*ddq  (updateAccelerationFD, integrateVelocities)
//   K4 capsule vs static box         (ODE dCollideCapsuleBox: dClosestLineBoxPoints + sphere/point)
//   K5 contact + joint-limit rows, A = J M^-1 J^T (+CFM) by impulse passes with the PLAIN
//      articulated inertia (ContactConstraint / JointLimitConstraint / applyUnitImpulse)
//   K6 boxed LCP: Dantzig pivoting with ODE's two-stage friction bounds, or fixed-sweep PGS
//   K7 dq += M^-1 J^T x ; q += dt*dq  (computeImpulseForwardDynamics, integratePositions)
// replacing the pydart2 call at gym/envs/dart/dart_env.py:174-175 of the reference.
//
// Layout: the skeleton TOPOLOGY (parents, joint types, shape->body map) is a compile-time
// policy `T`, so every per-body loop is unrolled and per-body quantities live in registers;
// the PARAMETERS (masses, anchors, limits...) are a __grid_constant__ kernel argument, i.e.
// constant-bank operands.  Spatial vectors are planar 3-vectors [angular; lin_x; lin_y] in
// WORLD axes taken at each body's own origin, so parent<->child transforms are pure shifts.
// Only the LCP (variable row count) uses thread-local memory.
#pragma once
#ifdef DARTB_HOST_EMU
// tools/host_emu: the SAME device code compiled as plain C++ for CPU-side numerics debugging and
// the no-GPU regression tests.  Never part of libdartb.so; the product has no CPU path.
#include "../../tools/host_emu/cuda_shims.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include <type_traits>

#include "../../include/dartb.h"
#include "planar_model.h"

#ifndef DEVI
#define DEVI __device__ __forceinline__
#endif
#ifdef DARTB_HOST_EMU
extern long g_emu_counters[8];  // [0] exact LCP calls [1] solved by lcp_small<4> [2] by lcp_small<8> [3] Dantzig [4] BPP iterations
extern long g_emu_hist[32];     // LCP row-count histogram (all lcp_exact calls)
#define EMU_COUNT(i, v) (g_emu_counters[i] += (v))
#define EMU_HIST(n) (g_emu_hist[(n) < 31 ? (n) : 31]++)
#else
#define EMU_COUNT(i, v) ((void)0)
#define EMU_HIST(n) ((void)0)
#endif

// max of v over the lanes of this warp that are currently executing this code
#ifdef DARTB_HOST_EMU
static inline int warp_max_active(int v) { return v; }
#else
DEVI int warp_max_active(int v) { return __reduce_max_sync(__activemask(), v); }
#endif

// ------------------------------------------------------------------------ static loops
template <int I, int N, class F>
DEVI void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}
template <int I, class F>
DEVI void static_rfor(F&& f) {  // I-1, ..., 0
    if constexpr (I > 0) {
        f(std::integral_constant<int, I - 1>{});
        static_rfor<I - 1>(f);
    }
}

// ------------------------------------------------------------------------ topologies
// (after the weld merge of lower.h; signature strings must match lower_model()'s)
template <int NB_, int NS_, int NL_>
struct TopoBase {
    static constexpr int NB = NB_, NS = NS_, NL = NL_;
    static constexpr int NSA = NS_ > 0 ? NS_ : 1;
    static constexpr int NR = 2 * NS_ + NL_;  // max LCP rows: (normal, tangent) per capsule + limits
};
struct TopoHopper : TopoBase<6, 4, 3> {  // hopper_capsule.skel: x, y, rot, thigh, shin, foot
    static constexpr const char* name = "hopper6";
    static constexpr const char* sig = "P-1,P0,R1,R2,R3,R4,S2,S3,S4,S5,";
    __host__ __device__ static constexpr int parent(int i) { constexpr int p[NB] = {-1, 0, 1, 2, 3, 4}; return p[i]; }
    __host__ __device__ static constexpr int jtype(int i) { constexpr int t[NB] = {2, 2, 1, 1, 1, 1}; return t[i]; }
    __host__ __device__ static constexpr int sbody(int s) { constexpr int b[NSA] = {2, 3, 4, 5}; return b[s]; }
};
struct TopoWalker : TopoBase<9, 7, 6> {  // walker2d.skel: root(3) + two 3-link legs
    static constexpr const char* name = "walker9";
    static constexpr const char* sig = "P-1,P0,R1,R2,R3,R4,R2,R6,R7,S2,S3,S4,S5,S6,S7,S8,";
    __host__ __device__ static constexpr int parent(int i) { constexpr int p[NB] = {-1, 0, 1, 2, 3, 4, 2, 6, 7}; return p[i]; }
    __host__ __device__ static constexpr int jtype(int i) { constexpr int t[NB] = {2, 2, 1, 1, 1, 1, 1, 1, 1}; return t[i]; }
    __host__ __device__ static constexpr int sbody(int s) { constexpr int b[NSA] = {2, 3, 4, 5, 6, 7, 8}; return b[s]; }
};
struct TopoCheetah : TopoBase<9, 8, 6> {  // half_cheetah.skel, head welded into the torso
    static constexpr const char* name = "cheetah9";
    static constexpr const char* sig = "P-1,P0,R1,R2,R3,R4,R2,R6,R7,S2,S2,S3,S4,S5,S6,S7,S8,";
    __host__ __device__ static constexpr int parent(int i) { constexpr int p[NB] = {-1, 0, 1, 2, 3, 4, 2, 6, 7}; return p[i]; }
    __host__ __device__ static constexpr int jtype(int i) { constexpr int t[NB] = {2, 2, 1, 1, 1, 1, 1, 1, 1}; return t[i]; }
    __host__ __device__ static constexpr int sbody(int s) { constexpr int b[NSA] = {2, 2, 3, 4, 5, 6, 7, 8}; return b[s]; }
};
struct TopoSnake : TopoBase<9, 0, 6> {  // snake_7link.skel: 9-chain, never touches the ground
    static constexpr const char* name = "snake9";
    static constexpr const char* sig = "P-1,P0,R1,R2,R3,R4,R5,R6,R7,";
    __host__ __device__ static constexpr int parent(int i) { constexpr int p[NB] = {-1, 0, 1, 2, 3, 4, 5, 6, 7}; return p[i]; }
    __host__ __device__ static constexpr int jtype(int i) { constexpr int t[NB] = {2, 2, 1, 1, 1, 1, 1, 1, 1}; return t[i]; }
    __host__ __device__ static constexpr int sbody(int) { return 0; }
};

template <class T>
__host__ __device__ constexpr bool topo_has_child(int i) {
    for (int k = 0; k < T::NB; k++) if (T::parent(k) == i) return true;
    return false;
}
template <class T>
__host__ __device__ constexpr bool topo_is_ancestor(int j, int b) {  // j == b or j above b
    for (int k = b; k >= 0; k = T::parent(k)) if (k == j) return true;
    return false;
}

// ------------------------------------------------------------------------ scalar helpers
template <typename R> struct Num;
template <> struct Num<float> {
#ifdef DARTB_LIBM_SINCOS
    static DEVI void sincos_(float x, float* s, float* c) { sincosf(x, s, c); }
#else
    // Branch-free: Cody-Waite reduction by pi/2 in three fma steps (pi/2 split into three floats), then the minimax
    // polynomials of sin and cos on [-pi/4, pi/4] and quadrant selects.  libm's sincosf carries a Payne-Hanek slow path for
    // |x| > 105615 that joint angles never reach (137 static branches per inlined copy; the hottest source line of the
    // headline capture, profiles/r2_hopper_quad_v2).  Max error 1.5 ulp / 7e-8 absolute for |x| <= 1e5 (libm: 2 ulp); the
    // same instructions on the CPU build, so the emulation is bit-identical here.
    static DEVI void sincos_(float x, float* s, float* c) {
        const float k = rintf(x * 6.366197467e-01f);
        float r = fmaf(k, -1.5707963705062866f, x);
        r = fmaf(k, 4.371138828673793e-08f, r);
        r = fmaf(k, 1.7151245100058819e-15f, r);
        const int q = (int)k;
        const float r2 = r * r;
        float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
        ps = fmaf(ps, r2, -1.6666654611e-1f);
        const float sn = fmaf(ps * r2, r, r);
        float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
        pc = fmaf(pc, r2, 4.166664568298827e-2f);
        pc = fmaf(pc, r2, -0.5f);
        const float cs = fmaf(pc, r2, 1.0f);
        const bool sw = (q & 1) != 0;
        const float so = sw ? cs : sn, co = sw ? sn : cs;
        *s = (q & 2) ? -so : so;
        *c = ((q + 1) & 2) ? -co : co;
    }
#endif
    static DEVI float sqrt_(float x) { return sqrtf(x); }
    // sqrt of the diagonal scales in the LCP's tolerance tests (tw = tol * (|b| + sqrt(A_ii) * S)): one MUFU-based
    // approximation (max rel. error 2^-22) instead of the IEEE sequence with its slow-path branch — sqrtf was the hottest
    // source line of the HalfCheetah capture (profiles/r2_cheetah16k_v2: 4.2 % of the samples)
#ifdef DARTB_HOST_EMU
    static DEVI float sqrt_tol_(float x) { return sqrtf(x); }
#else
    static DEVI float sqrt_tol_(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif
    // MUFU.RSQ, 2 ulp: inner solves of the LCP only (arguments >= inert(): the flush-to-zero form drops rsqrtf's
    // denormal pre-/post-scaling, four instructions per call)
#ifdef DARTB_HOST_EMU
    static DEVI float rsqrt_(float x) { return rsqrtf(x); }
#else
    static DEVI float rsqrt_(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif
    // 1/x of the projected articulated inertias (12 per DART step): one MUFU.RCP (max rel. error
    // 2^-23) instead of the IEEE sequence with its slow-path branch
#ifdef DARTB_HOST_EMU
    static DEVI float rcp_(float x) { return 1.0f / x; }
#else
    static DEVI float rcp_(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif
    static DEVI float abs_(float x) { return fabsf(x); }
    // c ? a : b as ONE select instruction (the compiler turns nested ternaries into branch trees)
#ifdef DARTB_HOST_EMU
    static DEVI float sel_(bool c, float a, float b) { return c ? a : b; }
#else
    static DEVI float sel_(bool c, float a, float b) {
        float r;
        asm("{\n.reg .pred p;\nsetp.ne.s32 p, %3, 0;\nselp.f32 %0, %1, %2, p;\n}" : "=f"(r) : "f"(a), "f"(b), "r"((int)c));
        return r;
    }
#endif
    static DEVI float min_(float a, float b) { return fminf(a, b); }
    static DEVI float max_(float a, float b) { return fmaxf(a, b); }
    static DEVI float inf() { return __int_as_float(0x7f800000); }
    static DEVI float inert() { return 1e-14f; }
    static DEVI float mindist() { return 1e-6f; }   // ODE dCollideCapsuleBox, dSINGLE build
    static DEVI float lcp_tol() { return 7.6e-6f; }   // 64 eps: complementarity tests of the pivoting LCP
};
template <> struct Num<double> {
    static DEVI void sincos_(double x, double* s, double* c) { sincos(x, s, c); }
    static DEVI double sqrt_(double x) { return sqrt(x); }
    static DEVI double sqrt_tol_(double x) { return sqrt(x); }
    static DEVI double rsqrt_(double x) { return 1.0 / sqrt(x); }
    static DEVI double rcp_(double x) { return 1.0 / x; }
    static DEVI double abs_(double x) { return fabs(x); }
#ifdef DARTB_HOST_EMU
    static DEVI double sel_(bool c, double a, double b) { return c ? a : b; }
#else
    static DEVI double sel_(bool c, double a, double b) {
        double r;
        asm("{\n.reg .pred p;\nsetp.ne.s32 p, %3, 0;\nselp.f64 %0, %1, %2, p;\n}" : "=d"(r) : "d"(a), "d"(b), "r"((int)c));
        return r;
    }
#endif
    static DEVI double min_(double a, double b) { return fmin(a, b); }
    static DEVI double max_(double a, double b) { return fmax(a, b); }
    static DEVI double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static DEVI double inert() { return 1e-14; }
    static DEVI double mindist() { return 1e-15; }  // ODE dCollideCapsuleBox, dDOUBLE build
    static DEVI double lcp_tol() { return 1.4e-14; }
};

// DART constants (ContactConstraint.cpp / JointLimitConstraint.cpp), see oracle/dart_oracle.c
#define DK_CONTACT_ERP 0.01
#define DK_CONTACT_MAX_ERV 1e-3
#define DK_CONTACT_CFM 1e-5
#define DK_LIMIT_CFM 1e-9
#define DK_FRICTION_THRESHOLD 1e-3
#define DK_CONTACT_EPS 1e-6

// exact-LCP form of the per-thread kernels: 0 = register block pivoting per size class (lcp_small<4/6/8>),
// 1 = thread-local block pivoting (lcp_bpp_local), 2 = compact rolled tableau (lcp_ppt_loop)
#ifndef DARTB_LCP_FORM
#define DARTB_LCP_FORM 0
#endif
// K4 collision + contact rows of the per-thread kernels: 1 = one rolled loop over the capsules, 0 = unrolled per capsule
#ifndef DARTB_ROLLED_COLLIDE
#define DARTB_ROLLED_COLLIDE 1
#endif
// quad form: deal the narrow phase out over the four lanes (capsule s on lane s % 4) instead of repeating it on each
#ifndef DARTB_QUAD_SPLIT_COLLIDE
#define DARTB_QUAD_SPLIT_COLLIDE 1
#endif

// ------------------------------------------------------------------------ Philox4x32-10
// identical to oracle/dart_oracle.c::orc_reset_uniform so reset noise is bit-identical
DEVI void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0];
        uint32_t h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = h1 ^ c[1] ^ k0, n1 = l1, n2 = h0 ^ c[3] ^ k1, n3 = l0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
DEVI float reset_uniform(uint64_t seed, int64_t world, uint32_t episode, int i) {
    uint32_t c[4] = {(uint32_t)world, (uint32_t)((uint64_t)world >> 32), episode, (uint32_t)(i >> 2)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t bits = c[i & 3];
    float u = __fmul_rn((float)(bits >> 8), 1.0f / 16777216.0f);
    return __fadd_rn(__fmul_rn(u, 2.0f), -1.0f);
}

// ------------------------------------------------------------------------ K4: segment vs box (2-D)
// ODE dClosestLineBoxPoints restricted to the plane (the out-of-plane axis has v = 0, region 0).
// Exact minimiser of the convex piecewise-quadratic distance along p1->p2; ties -> t = 0 (p1).
#if defined(DARTB_NOINLINE_CLOSEST) && !defined(DARTB_HOST_EMU)
#define CLOSEST_DECL __device__ __noinline__
#else
#define CLOSEST_DECL DEVI
#endif
template <typename R>
CLOSEST_DECL void closest_segment_box2(R p1x, R p1y, R p2x, R p2y, R cx, R cy, R hx, R hy, R& lx, R& ly, R& ddx, R& ddy) {
    R s[2] = {p1x - cx, p1y - cy}, v[2] = {p2x - p1x, p2y - p1y}, sg[2], v2[2], ta[2];
    const R h[2] = {hx, hy};
    const R dvx = v[0], dvy = v[1];
    int reg[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        if (v[i] < 0) { s[i] = -s[i]; v[i] = -v[i]; sg[i] = (R)-1; } else sg[i] = (R)1;
        v2[i] = v[i] * v[i];
        if (v[i] > (R)1e-19) {
            if (s[i] < -h[i]) { reg[i] = -1; ta[i] = (-h[i] - s[i]) / v[i]; }
            else { reg[i] = (s[i] > h[i]) ? 1 : 0; ta[i] = (h[i] - s[i]) / v[i]; }
        } else { reg[i] = 0; ta[i] = (R)2; }
    }
    R t = 0, dd = 0;
#pragma unroll
    for (int i = 0; i < 2; i++) dd -= (reg[i] ? v2[i] : (R)0) * ta[i];
    if (dd < 0) {
        bool found = false;
        for (int it = 0; it < 8; it++) {
            R nt = 1;
#pragma unroll
            for (int i = 0; i < 2; i++) if (ta[i] > t && ta[i] < (R)1 && ta[i] < nt) nt = ta[i];
            R nd = 0;
#pragma unroll
            for (int i = 0; i < 2; i++) nd += (reg[i] ? v2[i] : (R)0) * (nt - ta[i]);
            if (nd >= 0) {
                R mm = (nd - dd) / (nt - t);
                t -= dd / mm;
                found = true;
                break;
            }
#pragma unroll
            for (int i = 0; i < 2; i++) if (ta[i] == nt) { ta[i] = (h[i] - s[i]) / v[i]; reg[i]++; }
            t = nt;
            dd = nd;
            if (!(t < (R)1)) break;
        }
        if (!found) t = 1;
    }
    lx = p1x + t * dvx;
    ly = p1y + t * dvy;
    // (line point) - (box point), formed in box coordinates so it is EXACTLY zero when the line
    // point is inside the box (ODE forms both points in world coordinates and tests d < 1e-15)
    const R q0 = sg[0] * (s[0] + t * v[0]), q1 = sg[1] * (s[1] + t * v[1]);
    ddx = q0 - (q0 < -hx ? -hx : (q0 > hx ? hx : q0));
    ddy = q1 - (q1 < -hy ? -hy : (q1 > hy ? hy : q1));
}

// ------------------------------------------------------------------------ K6: boxed LCP
// Dense Cholesky solve of the nC x nC system A[idx,idx] y = rhs (thread-local scratch L).
template <typename R, int NR>
DEVI bool chol_solve_sub(int n, const R* A, const int* idx, int nC, const R* rhs, R* y, R* L) {
    for (int a = 0; a < nC; a++) {
        const int ia = idx[a];
        for (int b = 0; b <= a; b++) {
            R s = A[ia * n + idx[b]];
            for (int k = 0; k < b; k++) s -= L[a * NR + k] * L[b * NR + k];
            EMU_COUNT(7, b);
            if (a == b) {
                if (!(s > 0)) return false;
                L[a * NR + a] = Num<R>::sqrt_(s);
            } else L[a * NR + b] = s / L[b * NR + b];
        }
    }
    for (int a = 0; a < nC; a++) {
        R s = rhs[a];
        for (int k = 0; k < a; k++) s -= L[a * NR + k] * y[k];
        y[a] = s / L[a * NR + a];
    }
    for (int a = nC - 1; a >= 0; a--) {
        R s = y[a];
        for (int k = a + 1; k < nC; k++) s -= L[k * NR + a] * y[k];
        y[a] = s / L[a * NR + a];
    }
    return true;
}

// ODE dSolveLCP (Dantzig driving-index loop) as restated in oracle/dart_oracle.c:
// non-friction rows first; when the first friction row is reached its bounds become
// +-|mu * x[findex]| from the frictionless normal impulses and stay fixed.
template <typename R, int NR>
DEVI void lcp_dantzig(int n, const R* A, R* x, const R* b, R* lo, R* hi, const int* fidx) {
    R w[NR], dx[NR], dw[NR], L[NR * NR], rhs[NR], sol[NR];
    int idx[NR];
    uint32_t inC = 0, inN = 0, st = 0;
    for (int i = 0; i < n; i++) { x[i] = 0; w[i] = 0; }
    const R INF = Num<R>::inf();
    bool failed = false;
    for (int phase = 0; phase < 2 && !failed; phase++) {
        if (phase == 1) {
            bool any = false;
            for (int k = 0; k < n; k++)
                if (fidx[k] >= 0) {
                    any = true;
                    R wfk = x[fidx[k]];
                    if (wfk == 0) { hi[k] = 0; lo[k] = 0; }
                    else { hi[k] = Num<R>::abs_(hi[k] * wfk); lo[k] = -hi[k]; }
                }
            if (!any) break;
        }
        for (int i = 0; i < n && !failed; i++) {
            if ((fidx[i] >= 0) != (phase == 1)) continue;
            const uint32_t bi = 1u << i;
            if (!(A[i * n + i] > Num<R>::inert())) { x[i] = 0; w[i] = 0; lo[i] = 0; hi[i] = 0; inN |= bi; continue; }
            R wi = -b[i];
            const uint32_t live = inC | inN;
            for (int j = 0; j < n; j++) if (live >> j & 1) wi += A[i * n + j] * x[j];
            w[i] = wi;
            if (lo[i] == 0 && wi >= 0) { inN |= bi; st &= ~bi; continue; }
            if (hi[i] == 0 && wi <= 0) { inN |= bi; st |= bi; continue; }
            if (wi == 0) { inC |= bi; continue; }
            bool placed = false;
            for (int guard = 0; guard < 10 * n + 50; guard++) {
                const R dirf = (w[i] <= 0) ? (R)1 : (R)-1;
                int nC = 0;
                for (int j = 0; j < n; j++) if (inC >> j & 1) idx[nC++] = j;
                for (int r = 0; r < nC; r++) rhs[r] = -dirf * A[idx[r] * n + i];
                if (nC > 0 && !chol_solve_sub<R, NR>(n, A, idx, nC, rhs, sol, L)) { failed = true; break; }
                for (int r = 0; r < nC; r++) dx[idx[r]] = sol[r];
                for (int j = 0; j < n; j++) {
                    if (!((inN >> j & 1) || j == i)) continue;
                    R s = A[j * n + i] * dirf;
                    for (int r = 0; r < nC; r++) s += A[j * n + idx[r]] * sol[r];
                    dw[j] = s;
                }
                int cmd = 1, si = 0;
                R s = -w[i] / dw[i];
                if (dirf > 0) {
                    if (hi[i] < INF) { R s2 = (hi[i] - x[i]) * dirf; if (s2 < s) { s = s2; cmd = 3; } }
                } else {
                    if (lo[i] > -INF) { R s2 = (lo[i] - x[i]) * dirf; if (s2 < s) { s = s2; cmd = 2; } }
                }
                for (int k = 0; k < n; k++) {
                    if (!(inN >> k & 1)) continue;
                    const bool stk = st >> k & 1;
                    if ((!stk && dw[k] < 0) || (stk && dw[k] > 0)) {
                        if (lo[k] == 0 && hi[k] == 0) continue;
                        R s2 = -w[k] / dw[k];
                        if (s2 < s) { s = s2; cmd = 4; si = k; }
                    }
                }
                for (int r = 0; r < nC; r++) {
                    const int k = idx[r];
                    if (sol[r] < 0 && lo[k] > -INF) { R s2 = (lo[k] - x[k]) / sol[r]; if (s2 < s) { s = s2; cmd = 5; si = k; } }
                    if (sol[r] > 0 && hi[k] < INF) { R s2 = (hi[k] - x[k]) / sol[r]; if (s2 < s) { s = s2; cmd = 6; si = k; } }
                }
                if (!(s > 0)) {
                    if (s != s) { failed = true; break; }
                    s = 0;  // rounding produced a tiny negative step: take a zero step and switch
                }
                for (int r = 0; r < nC; r++) x[idx[r]] += s * sol[r];
                x[i] += s * dirf;
                for (int k = 0; k < n; k++) if (inN >> k & 1) w[k] += s * dw[k];
                w[i] += s * dw[i];
                const uint32_t bs = 1u << si;
                switch (cmd) {
                    case 1: w[i] = 0; inC |= bi; break;
                    case 2: x[i] = lo[i]; st &= ~bi; inN |= bi; break;
                    case 3: x[i] = hi[i]; st |= bi; inN |= bi; break;
                    case 4: w[si] = 0; inN &= ~bs; inC |= bs; break;
                    case 5: x[si] = lo[si]; st &= ~bs; inC &= ~bs; inN |= bs; break;
                    default: x[si] = hi[si]; st |= bs; inC &= ~bs; inN |= bs; break;
                }
                if (cmd <= 3) { placed = true; break; }
            }
            if (!placed && !failed) { inN |= bi; }
        }
    }
    if (failed) {  // ODE: "LCP internal error": keep what was solved, finite values only
        for (int i = 0; i < n; i++) if (!(x[i] == x[i]) || Num<R>::abs_(x[i]) == INF) x[i] = 0;
    }
}

// ------------------------------------------------------------------------ K6 fast path: small n in registers
// Exact boxed LCP for n <= NM rows by block principal pivoting (Judice & Pires) with a masked
// dense Cholesky, everything in registers and fully unrolled (no thread-local memory, no
// data-dependent loops except the pivoting iterations).  The solution of the two-stage problem
// (frictionless normals first, then friction bounds +-mu*x_n fixed: ODE dSolveLCP semantics) is
// unique for the positive-definite A = J M^-1 J^T (1 + CFM), so any exact method returns what
// Dantzig returns; ncu showed the Dantzig loop (thread-local arrays, 3-5 active lanes) to be 50% of
// the kernel's instructions.  Returns false if it did not converge (caller falls back to Dantzig).
template <typename R, int NM>
DEVI bool lcp_small(int n, const R* Ag, R* xg, const R* bg, const R* log_, const R* hig, const int* fidxg,
                    const uint8_t* hin = nullptr, uint8_t* sout = nullptr) {
    R A[NM][NM], b[NM], lo[NM], hi[NM], x[NM], mu[NM];
    int fi[NM];
    unsigned st = 0;  // 2 bits per row: 0 free, 1 at lo, 2 at hi, 3 permanently bound at x = 0
    const R INF = Num<R>::inf();
#pragma unroll
    for (int i = 0; i < NM; i++) {
        const bool on = i < n;
        b[i] = on ? bg[i] : (R)0; lo[i] = on ? log_[i] : (R)0; hi[i] = on ? hig[i] : (R)0; fi[i] = on ? fidxg[i] : -1;
        mu[i] = hi[i];
        x[i] = 0;
#pragma unroll
        for (int j = 0; j < NM; j++) A[i][j] = (on && j < n) ? Ag[i * n + j] : (i == j ? (R)1 : (R)0);
        // initial active set = the solution of the decoupled (diagonal) problem: a unilateral row is
        // free iff its own b asks for an impulse of the admissible sign; usually already correct, so
        // the first pivoting iteration only verifies it.
        // (`hin[i]`, the set the row ended in at the previous DART step, is used for the FRICTION rows
        // in stage 2 only: stick/slide persists between steps, while for the unilateral rows the
        // decoupled guess above measured better than the previous step's set under random actions)
        // (flat selects: padding / inert rows and friction rows -> 3; lower-unilateral -> free or at lo; upper-unilateral ->
        //  free or at hi; any other box starts free)
        const bool fixed = !on || !(A[i][i] > Num<R>::inert()) || fi[i] >= 0;
        const bool lower = lo[i] == 0 && hi[i] == INF, upper = hi[i] == 0 && lo[i] == -INF;
        const unsigned sl = b[i] > 0 ? 0u : 1u, su = b[i] < 0 ? 0u : 2u;
        const unsigned s = fixed ? 3u : (lower ? sl : (upper ? su : 0u));
        st |= s << (2 * i);
    }
    bool ok = true;
    // rounding-aware feasibility tests: |A_ij x_j| <= sqrt(A_ii) sqrt(A_jj) |x_j| (A is PSD), so
    // sqrt(A_ii) * sum_j sqrt(A_jj)|x_j| bounds the magnitude of the terms of w_i.  Without the
    // tolerance a degenerate row (x at its bound AND w = 0) flips between sets forever in fp32.
    R sd[NM];
#pragma unroll
    for (int i = 0; i < NM; i++) sd[i] = Num<R>::sqrt_tol_(A[i][i]);
#pragma unroll 1
    for (int stage = 0; stage < 2; stage++) {
        if (stage == 1) {
            bool any = false;
#pragma unroll
            for (int i = 0; i < NM; i++) {
                // (flat selects again: one predicate per row instead of a branch tree)
                const bool fr = fi[i] >= 0 && i < n && A[i][i] > Num<R>::inert();
                R xn = 0;
#pragma unroll
                for (int j = 0; j < NM; j++) xn = Num<R>::sel_(j == fi[i], x[j], xn);
                const R h = Num<R>::abs_(mu[i] * xn);
                hi[i] = Num<R>::sel_(fr, h, hi[i]); lo[i] = Num<R>::sel_(fr, -h, lo[i]);
                const unsigned hh = hin ? hin[i] : 3u;
                const unsigned ns = h == 0 ? 3u : (hh < 3u ? hh : 0u);   // hinted set, else free (sticking)
                st = fr ? ((st & ~(3u << (2 * i))) | (ns << (2 * i))) : st;
                any = any || (fr && h != 0);
            }
            if (!any) break;
        }
        int best = NM + 1, tries = 3;
        bool done = false;
#pragma unroll 1
        for (int it = 0; it < 6 + 3 * NM && !done; it++) {
            // masked system: free rows keep A, bound rows become identity with rhs = bound value
            R L[NM][NM], y[NM];
#pragma unroll
            for (int i = 0; i < NM; i++) {
                const unsigned si = (st >> (2 * i)) & 3u;
                const R xb = Num<R>::sel_(si == 1, lo[i], Num<R>::sel_(si == 2, hi[i], (R)0));
                x[i] = Num<R>::sel_(si != 0, xb, x[i]);
            }
            // (this whole iteration body is written without data-dependent branches: selects and masked products only.
            //  A lone warp per scheduler pays ~20 cycles for every branch; r2: 87 branches in 930 instructions before)
#pragma unroll
            for (int i = 0; i < NM; i++) {
                const bool fr = ((st >> (2 * i)) & 3u) == 0;
                R r = b[i];
#pragma unroll
                for (int j = 0; j < NM; j++) {
                    const bool fj = ((st >> (2 * j)) & 3u) == 0;
                    r -= A[i][j] * (fj ? (R)0 : x[j]);
                }
                y[i] = fr ? r : (R)0;
            }
            // Cholesky of the masked matrix (bound rows/cols = identity)
            bool pd = true;
#pragma unroll
            for (int i = 0; i < NM; i++) {
                const bool fr = ((st >> (2 * i)) & 3u) == 0;
#pragma unroll
                for (int j = 0; j <= i; j++) {
                    const bool fj = ((st >> (2 * j)) & 3u) == 0;
                    R s = (fr && fj) ? A[i][j] : (i == j ? (R)1 : (R)0);
#pragma unroll
                    for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
                    if (i == j) { const bool pos = s > 0; pd = pd && pos; L[i][i] = Num<R>::rsqrt_(pos ? s : (R)1); }
                    else L[i][j] = s * L[j][j];
                }
            }
            if (!pd) { ok = false; break; }
#pragma unroll
            for (int i = 0; i < NM; i++) {
                R s = y[i];
#pragma unroll
                for (int k = 0; k < i; k++) s -= L[i][k] * y[k];
                y[i] = s * L[i][i];
            }
#pragma unroll
            for (int i = NM - 1; i >= 0; i--) {
                R s = y[i];
#pragma unroll
                for (int k = i + 1; k < NM; k++) s -= L[k][i] * y[k];
                y[i] = s * L[i][i];
            }
#pragma unroll
            for (int i = 0; i < NM; i++) x[i] = (((st >> (2 * i)) & 3u) == 0) ? y[i] : x[i];
            // infeasibilities (with rounding tolerances)
            unsigned bad = 0;  // bit i set: row i must change set
            int nbad = 0;
            unsigned nst = st;
            R xs = 0, S = 0;
#pragma unroll
            for (int i = 0; i < NM; i++) { const R ax = Num<R>::abs_(x[i]); xs = ax > xs ? ax : xs; S += sd[i] * ax; }
            const R tx = Num<R>::lcp_tol() * xs;
#pragma unroll
            for (int i = 0; i < NM; i++) {
                const unsigned si = (st >> (2 * i)) & 3u;
                // a free row leaves its box / a bound row's w has the wrong sign (w is formed for every row: a few wasted
                // products instead of a branch per row)
                const bool below = x[i] < lo[i] - tx, above = x[i] > hi[i] + tx;
                R w = -b[i];
#pragma unroll
                for (int j = 0; j < NM; j++) w += A[i][j] * x[j];
                const R tw = Num<R>::lcp_tol() * (Num<R>::abs_(b[i]) + sd[i] * S);
                const bool wbad = ((si == 1 && w < -tw) || (si == 2 && w > tw)) && lo[i] < hi[i];
                const bool badi = si == 0 ? (below || above) : wbad;     // (si == 3 makes wbad false)
                const unsigned ns = si == 0 ? (below ? 1u : 2u) : 0u;     // the set the row moves to when it is infeasible
                bad |= (badi ? 1u : 0u) << i;
                nbad += badi ? 1 : 0;
                nst = badi ? ((nst & ~(3u << (2 * i))) | (ns << (2 * i))) : nst;
            }
            EMU_COUNT(4, 1);
            if (nbad == 0) { done = true; break; }
            if (nbad < best) { best = nbad; tries = 3; st = nst; }
            else if (tries > 0) { tries--; st = nst; }
            else {  // Murty: flip only the highest-index infeasible row (finite for P-matrices)
                const int k = 31 - __clz(bad);
                st = (st & ~(3u << (2 * k))) | (nst & (3u << (2 * k)));
            }
        }
        if (!done) { ok = false; }
        if (!ok) break;
    }
    if (!ok) { EMU_COUNT(6, 1); return false; }
#pragma unroll
    for (int i = 0; i < NM; i++) if (i < n) { xg[i] = x[i]; if (sout) sout[i] = (uint8_t)((st >> (2 * i)) & 3u); }
    return true;
}

// ------------------------------------------------------------------------ K6 fast path, tableau form
// The same block-principal-pivoting iteration as lcp_small (same sets visited, same rounding-aware
// tests), with the linear algebra kept as a PRINCIPAL PIVOT TRANSFORM of A instead of a fresh masked
// Cholesky per iteration: with F the free rows and B the bound ones, the tableau T maps
// z = [b_F ; x_B] to y = [x_F ; (w + b)_B].  Moving one row between F and B is one symmetric
// exchange step on T ((NM-1)^2 FMAs, pivot T_kk > 0 for the positive-definite A), so an iteration
// costs (rows that changed set) exchanges + one NM x NM product, and stage 2 (friction bounds fixed
// from the frictionless normals, ODE dSolveLCP semantics) CONTINUES from stage 1's tableau.
// ncu (profiles/r1_hopper_v3_stages.md): the Cholesky form was 28 % of the kernel's instructions.
template <int K, typename R, int NM>
DEVI bool ppt_exchange(R (&T)[NM][NM]) {
    const R d = T[K][K];
    if (!(d > 0)) return false;
    const R p = Num<R>::rcp_(d);
    R rk[NM];
#pragma unroll
    for (int j = 0; j < NM; j++) rk[j] = T[K][j] * p;
#pragma unroll
    for (int i = 0; i < NM; i++) {
        if (i == K) continue;
        const R c = T[i][K];
#pragma unroll
        for (int j = 0; j < NM; j++) if (j != K) T[i][j] -= c * rk[j];
        T[i][K] = c * p;
    }
#pragma unroll
    for (int j = 0; j < NM; j++) if (j != K) T[K][j] = -rk[j];
    T[K][K] = p;
    return true;
}

template <typename R, int NM>
DEVI bool lcp_ppt(int n, const R* Ag, R* xg, const R* bg, const R* log_, const R* hig, const int* fidxg,
                  const uint8_t* hin = nullptr, uint8_t* sout = nullptr) {
    R T[NM][NM], b[NM], lo[NM], hi[NM], x[NM], mu[NM], sd[NM];
    int fi[NM];
    unsigned cur = 0;   // set the tableau currently represents; 2 bits per row: 0 free, 1 at lo, 2 at hi, 3 fixed at 0
    unsigned st = 0;    // set to evaluate next
    const R INF = Num<R>::inf();
#pragma unroll
    for (int i = 0; i < NM; i++) {
        const bool on = i < n;
        b[i] = on ? bg[i] : (R)0; lo[i] = on ? log_[i] : (R)0; hi[i] = on ? hig[i] : (R)0; fi[i] = on ? fidxg[i] : -1;
        mu[i] = hi[i];
        x[i] = 0;
#pragma unroll
        for (int j = 0; j < NM; j++) T[i][j] = (on && j < n) ? Ag[i * n + j] : (i == j ? (R)1 : (R)0);
        sd[i] = Num<R>::sqrt_tol_(T[i][i]);
        // initial set = the solution of the decoupled (diagonal) problem (see lcp_small)
        unsigned s = 0, c = 3;
        if (!on || !(T[i][i] > Num<R>::inert())) s = 3;            // padding / inert row
        else if (fi[i] >= 0) s = 3;                                  // friction rows wait for stage 2
        else if (lo[i] == 0 && hi[i] == INF) { s = b[i] > 0 ? 0u : 1u; c = 1; }
        else if (hi[i] == 0 && lo[i] == -INF) { s = b[i] < 0 ? 0u : 2u; c = 2; }
        else c = 0;                                                  // two-sided non-friction row: starts free
        st |= s << (2 * i);
        cur |= (c == 0 ? 3u : c) << (2 * i);                         // the tableau starts as A itself: every row bound
    }
    bool ok = true;
#pragma unroll 1
    for (int stage = 0; stage < 2; stage++) {
        if (stage == 1) {
            bool any = false;
#pragma unroll
            for (int i = 0; i < NM; i++) {
                if (fi[i] >= 0 && i < n && T[i][i] > Num<R>::inert()) {   // still bound: T_ii is A_ii's Schur complement > 0
                    R xn = 0;
#pragma unroll
                    for (int j = 0; j < NM; j++) if (j == fi[i]) xn = x[j];
                    const R h = Num<R>::abs_(mu[i] * xn);
                    hi[i] = h; lo[i] = -h;
                    st &= ~(3u << (2 * i));
                    const unsigned hh = hin ? hin[i] : 3u;
                    if (h == 0) st |= 3u << (2 * i);
                    else { any = true; if (hh < 3u) st |= hh << (2 * i); }   // hinted set, else free (sticking)
                }
            }
            if (!any) break;
        }
        int best = NM + 1, tries = 3;
        bool done = false;
#pragma unroll 1
        for (int it = 0; it < 6 + 3 * NM && !done; it++) {
            // bring the tableau to the set `st`: one exchange per row whose free/bound status differs
            unsigned flip = 0;
#pragma unroll
            for (int i = 0; i < NM; i++)
                if ((((st >> (2 * i)) & 3u) == 0) != (((cur >> (2 * i)) & 3u) == 0)) flip |= 1u << i;
            bool pd = true;
            static_for<0, NM>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                if ((flip >> k) & 1u) pd = ppt_exchange<k, R, NM>(T) && pd;
            });
            if (!pd) { ok = false; break; }
            cur = st;
            R z[NM];
#pragma unroll
            for (int i = 0; i < NM; i++) {
                const unsigned si = (st >> (2 * i)) & 3u;
                z[i] = si == 0 ? b[i] : (si == 1 ? lo[i] : (si == 2 ? hi[i] : (R)0));
            }
            unsigned nst = st, bad = 0;
            int nbad = 0;
            R y[NM], xs = 0, S = 0;
#pragma unroll
            for (int i = 0; i < NM; i++) {
                R s = 0;
#pragma unroll
                for (int j = 0; j < NM; j++) s += T[i][j] * z[j];
                y[i] = s;
                x[i] = ((st >> (2 * i)) & 3u) == 0 ? s : z[i];
                const R ax = Num<R>::abs_(x[i]);
                xs = ax > xs ? ax : xs;
                S += sd[i] * ax;
            }
            const R tx = Num<R>::lcp_tol() * xs;
#pragma unroll
            for (int i = 0; i < NM; i++) {
                const unsigned si = (st >> (2 * i)) & 3u;
                if (si == 3) continue;
                if (si == 0) {
                    if (x[i] < lo[i] - tx) { bad |= 1u << i; nbad++; nst = (nst & ~(3u << (2 * i))) | (1u << (2 * i)); }
                    else if (x[i] > hi[i] + tx) { bad |= 1u << i; nbad++; nst = (nst & ~(3u << (2 * i))) | (2u << (2 * i)); }
                } else {
                    const R w = y[i] - b[i];
                    const R tw = Num<R>::lcp_tol() * (Num<R>::abs_(b[i]) + sd[i] * S);
                    if ((si == 1 && w < -tw) || (si == 2 && w > tw)) {
                        if (lo[i] < hi[i]) { bad |= 1u << i; nbad++; nst = nst & ~(3u << (2 * i)); }
                    }
                }
            }
            EMU_COUNT(4, 1);
            if (nbad == 0) { done = true; break; }
            if (nbad < best) { best = nbad; tries = 3; st = nst; }
            else if (tries > 0) { tries--; st = nst; }
            else {  // Murty: flip only the highest-index infeasible row (finite for P-matrices)
                const int k = 31 - __clz(bad);
                st = (st & ~(3u << (2 * k))) | (nst & (3u << (2 * k)));
            }
        }
        if (!done) { ok = false; }
        if (!ok) break;
        // one round of iterative refinement against A itself (still in thread-local memory): the
        // tableau's A_FF^-1 is a Gauss-Jordan inverse, and for the nearly rank-deficient A of several
        // contacts on one body (regularised by the CFM only) its fp32 residual is ~100x Cholesky's
        {
            R r[NM];
#pragma unroll
            for (int i = 0; i < NM; i++) {
                R s = 0;
                if (i < n && ((st >> (2 * i)) & 3u) == 0) {
                    s = b[i];
#pragma unroll
                    for (int j = 0; j < NM; j++) if (j < n) s -= Ag[i * n + j] * x[j];
                }
                r[i] = s;
            }
#pragma unroll
            for (int i = 0; i < NM; i++) {
                R s = 0;
#pragma unroll
                for (int j = 0; j < NM; j++) s += T[i][j] * r[j];
                if (((st >> (2 * i)) & 3u) == 0) x[i] += s;
            }
        }
    }
    if (!ok) { EMU_COUNT(6, 1); return false; }
#ifdef DARTB_HOST_EMU
    if (getenv("EMU_LCP_DEBUG")) {
        printf("ppt<%d,%d> n=%d st=", NM, (int)sizeof(R), n);
        for (int i = 0; i < n; i++) printf("%u", (st >> (2 * i)) & 3u);
        printf(" x=");
        for (int i = 0; i < n; i++) printf(" %.6g", (double)x[i]);
        printf("\n");
    }
#endif
#pragma unroll
    for (int i = 0; i < NM; i++) if (i < n) { xg[i] = x[i]; if (sout) sout[i] = (uint8_t)((st >> (2 * i)) & 3u); }
    return true;
}

// The same block-principal-pivoting iteration for any n <= NR, as loops over thread-local arrays
// (used above the register sizes; still ~n^3/6 + 3n^2 MACs per iteration and 2-3 iterations per
// stage, against one fresh factorisation per pivot in the Dantzig loop).
template <typename R, int NR>
DEVI bool lcp_bpp_local(int n, const R* A, R* x, const R* b, const R* lo_in, const R* hi_in, const int* fidx) {
    R L[NR * NR], y[NR], lo[NR], hi[NR];
    int idx[NR];
    uint64_t st = 0;
    for (int i = 0; i < n; i++) {
        lo[i] = lo_in[i]; hi[i] = hi_in[i];
        x[i] = 0;
        uint64_t s = 0;
        if (!(A[i * n + i] > Num<R>::inert()) || fidx[i] >= 0) s = 3;
        else if (lo[i] == 0 && hi[i] == Num<R>::inf()) s = b[i] > 0 ? 0 : 1;
        else if (hi[i] == 0 && lo[i] == -Num<R>::inf()) s = b[i] < 0 ? 0 : 2;
        st |= s << (2 * i);
    }
    for (int stage = 0; stage < 2; stage++) {
        if (stage == 1) {
            bool any = false;
            for (int i = 0; i < n; i++)
                if (fidx[i] >= 0 && A[i * n + i] > Num<R>::inert()) {
                    const R h = Num<R>::abs_(hi[i] * x[fidx[i]]);
                    hi[i] = h; lo[i] = -h;
                    st &= ~((uint64_t)3 << (2 * i));
                    if (h == 0) st |= (uint64_t)3 << (2 * i); else any = true;
                }
            if (!any) break;
        }
        int best = n + 1, tries = 3;
        bool done = false;
        for (int it = 0; it < 3 * n + 6 && !done; it++) {
            // compact list of free rows; bound rows take their bound value
            int nF = 0;
            for (int i = 0; i < n; i++) {
                const unsigned si = (unsigned)(st >> (2 * i)) & 3u;
                if (si == 0) idx[nF++] = i;
                else x[i] = si == 1 ? lo[i] : (si == 2 ? hi[i] : (R)0);
            }
            // rhs_F = b_F - A_FB x_B ; Cholesky of A_FF ; solve
            for (int a = 0; a < nF; a++) {
                const int ia = idx[a];
                R r = b[ia];
                for (int j = 0; j < n; j++) if (((st >> (2 * j)) & 3u) != 0) r -= A[ia * n + j] * x[j];
                y[a] = r;
                for (int c = 0; c <= a; c++) {
                    R s = A[ia * n + idx[c]];
#pragma unroll 4
                    for (int k = 0; k < c; k++) s -= L[a * NR + k] * L[c * NR + k];
                    EMU_COUNT(7, c);
                    if (a == c) { if (!(s > 0)) return false; L[a * NR + a] = Num<R>::rsqrt_(s); }
                    else L[a * NR + c] = s * L[c * NR + c];
                }
            }
            for (int a = 0; a < nF; a++) {
                R s = y[a];
#pragma unroll 4
                for (int k = 0; k < a; k++) s -= L[a * NR + k] * y[k];
                y[a] = s * L[a * NR + a];
            }
            for (int a = nF - 1; a >= 0; a--) {
                R s = y[a];
#pragma unroll 4
                for (int k = a + 1; k < nF; k++) s -= L[k * NR + a] * y[k];
                y[a] = s * L[a * NR + a];
            }
            EMU_COUNT(7, nF * nF);
            for (int a = 0; a < nF; a++) x[idx[a]] = y[a];
            uint64_t nst = st;
            int nbad = 0, last = -1;
            R xs = 0, S = 0;
            for (int i = 0; i < n; i++) { const R ax = Num<R>::abs_(x[i]); xs = ax > xs ? ax : xs; S += Num<R>::sqrt_tol_(A[i * n + i]) * ax; }
            const R tx = Num<R>::lcp_tol() * xs;
            for (int i = 0; i < n; i++) {
                const unsigned si = (unsigned)(st >> (2 * i)) & 3u;
                if (si == 3) continue;
                const uint64_t clr = ~((uint64_t)3 << (2 * i));
                if (si == 0) {
                    if (x[i] < lo[i] - tx) { nbad++; last = i; nst = (nst & clr) | ((uint64_t)1 << (2 * i)); }
                    else if (x[i] > hi[i] + tx) { nbad++; last = i; nst = (nst & clr) | ((uint64_t)2 << (2 * i)); }
                } else {
                    R w = -b[i];
#pragma unroll 4
                    for (int j = 0; j < n; j++) w += A[i * n + j] * x[j];
                    EMU_COUNT(7, n);
                    const R tw = Num<R>::lcp_tol() * (Num<R>::abs_(b[i]) + Num<R>::sqrt_tol_(A[i * n + i]) * S);
                    if (((si == 1 && w < -tw) || (si == 2 && w > tw)) && lo[i] < hi[i]) { nbad++; last = i; nst = nst & clr; }
                }
            }
            EMU_COUNT(4, 1);
            if (nbad == 0) { done = true; break; }
            if (nbad < best) { best = nbad; tries = 3; st = nst; }
            else if (tries > 0) { tries--; st = nst; }
            else { const uint64_t m2 = (uint64_t)3 << (2 * last); st = (st & ~m2) | (nst & m2); }
        }
        if (!done) return false;
    }
    return true;
}

// ------------------------------------------------------------------------ K6 compact path: tableau in loops
// lcp_ppt's iteration (same sets visited, same rounding-aware tests, same two-stage friction bounds and
// refinement) written as rolled loops over a thread-local tableau of stride n.  ~400 SASS instructions
// for any n <= NR, against ~5 k per register size class of lcp_small: the large-batch per-thread kernels
// are bound by instruction fetch (ncu r1: no_inst 47-55 % with the unrolled forms), and a solve that
// stays inside the L0/L1.5 instruction caches costs less than one that executes fewer, colder instructions.
template <typename R>
DEVI unsigned set2(uint64_t v, int i) { return (unsigned)(v >> (2 * i)) & 3u; }
DEVI uint64_t put2(uint64_t v, int i, unsigned s) { return (v & ~((uint64_t)3 << (2 * i))) | ((uint64_t)s << (2 * i)); }

template <typename R, int NR>
DEVI bool lcp_ppt_loop(int n, const R* A, R* x, const R* b, const R* lo_in, const R* hi_in, const int* fidx,
                       const uint8_t* hin = nullptr, uint8_t* sout = nullptr) {
    R T[NR * NR], lo[NR], hi[NR], sd[NR], z[NR], y[NR];
    uint64_t cur = 0, st = 0;   // 2 bits per row: 0 free, 1 at lo, 2 at hi, 3 fixed at 0 (cur: what the tableau represents)
    const R INF = Num<R>::inf();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        lo[i] = lo_in[i]; hi[i] = hi_in[i]; x[i] = 0;
#pragma unroll 1
        for (int j = 0; j < n; j++) T[i * n + j] = A[i * n + j];
        const R d = A[i * n + i];
        sd[i] = Num<R>::sqrt_tol_(d);
        unsigned s = 0, c = 3;
        if (!(d > Num<R>::inert())) s = 3;                            // inert row
        else if (fidx[i] >= 0) s = 3;                                  // friction rows wait for stage 2
        else if (lo[i] == 0 && hi[i] == INF) { s = b[i] > 0 ? 0u : 1u; c = 1; }
        else if (hi[i] == 0 && lo[i] == -INF) { s = b[i] < 0 ? 0u : 2u; c = 2; }
        st = put2(st, i, s); cur = put2(cur, i, c);                    // the tableau starts as A itself: every row bound
    }
#pragma unroll 1
    for (int stage = 0; stage < 2; stage++) {
        if (stage == 1) {
            bool any = false;
#pragma unroll 1
            for (int i = 0; i < n; i++) {
                if (fidx[i] >= 0 && T[i * n + i] > Num<R>::inert()) {   // still bound: T_ii is A_ii's Schur complement > 0
                    const R h = Num<R>::abs_(hi_in[i] * x[fidx[i]]);
                    hi[i] = h; lo[i] = -h;
                    const unsigned hh = hin ? hin[i] : 3u;
                    unsigned s = 3;
                    if (h != 0) { any = true; s = hh < 3u ? hh : 0u; }   // hinted set, else free (sticking)
                    st = put2(st, i, s);
                }
            }
            if (!any) break;
        }
        int best = n + 1, tries = 3;
        bool done = false;
#pragma unroll 1
        for (int it = 0; it < 6 + 3 * n && !done; it++) {
            // bring the tableau to the set `st`: one principal exchange per row whose free/bound status differs
#pragma unroll 1
            for (int k = 0; k < n; k++) {
                if ((set2<R>(st, k) == 0) == (set2<R>(cur, k) == 0)) continue;
                const R d = T[k * n + k];
                if (!(d > 0)) return false;
                const R p = Num<R>::rcp_(d);
#pragma unroll 1
                for (int i = 0; i < n; i++) {
                    if (i == k) continue;
                    const R c = T[i * n + k];
#pragma unroll 1
                    for (int j = 0; j < n; j++) if (j != k) T[i * n + j] -= c * (T[k * n + j] * p);
                    T[i * n + k] = c * p;
                }
#pragma unroll 1
                for (int j = 0; j < n; j++) if (j != k) T[k * n + j] = -(T[k * n + j] * p);
                T[k * n + k] = p;
            }
            cur = st;
#pragma unroll 1
            for (int i = 0; i < n; i++) {
                const unsigned si = set2<R>(st, i);
                z[i] = si == 0 ? b[i] : (si == 1 ? lo[i] : (si == 2 ? hi[i] : (R)0));
            }
            R xs = 0, S = 0;
#pragma unroll 1
            for (int i = 0; i < n; i++) {
                R s = 0;
#pragma unroll 1
                for (int j = 0; j < n; j++) s += T[i * n + j] * z[j];
                y[i] = s;
                x[i] = set2<R>(st, i) == 0 ? s : z[i];
                const R ax = Num<R>::abs_(x[i]);
                xs = ax > xs ? ax : xs;
                S += sd[i] * ax;
            }
            const R tx = Num<R>::lcp_tol() * xs;
            uint64_t nst = st;
            int nbad = 0, last = -1;
#pragma unroll 1
            for (int i = 0; i < n; i++) {
                const unsigned si = set2<R>(st, i);
                if (si == 3) continue;
                if (si == 0) {
                    if (x[i] < lo[i] - tx) { nbad++; last = i; nst = put2(nst, i, 1); }
                    else if (x[i] > hi[i] + tx) { nbad++; last = i; nst = put2(nst, i, 2); }
                } else {
                    const R w = y[i] - b[i];
                    const R tw = Num<R>::lcp_tol() * (Num<R>::abs_(b[i]) + sd[i] * S);
                    if (((si == 1 && w < -tw) || (si == 2 && w > tw)) && lo[i] < hi[i]) { nbad++; last = i; nst = put2(nst, i, 0); }
                }
            }
            EMU_COUNT(4, 1);
            if (nbad == 0) { done = true; break; }
            if (nbad < best) { best = nbad; tries = 3; st = nst; }
            else if (tries > 0) { tries--; st = nst; }
            else st = put2(st, last, set2<R>(nst, last));   // Murty: flip only the highest-index infeasible row
        }
        if (!done) return false;
        // one round of iterative refinement against A itself (see lcp_ppt)
#pragma unroll 1
        for (int i = 0; i < n; i++) {
            R s = 0;
            if (set2<R>(st, i) == 0) {
                s = b[i];
#pragma unroll 1
                for (int j = 0; j < n; j++) s -= A[i * n + j] * x[j];
            }
            z[i] = s;
        }
#pragma unroll 1
        for (int i = 0; i < n; i++) {
            if (set2<R>(st, i) != 0) continue;
            R s = 0;
#pragma unroll 1
            for (int j = 0; j < n; j++) s += T[i * n + j] * z[j];
            y[i] = s;
        }
#pragma unroll 1
        for (int i = 0; i < n; i++) if (set2<R>(st, i) == 0) x[i] += y[i];
    }
    if (sout) for (int i = 0; i < n; i++) sout[i] = (uint8_t)set2<R>(st, i);
    return true;
}

// exact LCP dispatch.  The size class is chosen per WARP (max n over the lanes that have rows), so
// a warp executes ONE code path instead of one per distinct n: register block pivoting for
// n <= 4 / 6 / 8, the thread-local block pivoting above that, Dantzig only if pivoting fails.
template <typename R, int NR>
DEVI void lcp_exact(int n, const R* A, R* x, const R* b, R* lo, R* hi, const int* fidx, const uint8_t* hin = nullptr,
                    uint8_t* sout = nullptr) {
    const int nmax = warp_max_active(n);
    bool ok = false;
    EMU_COUNT(0, 1);
    EMU_HIST(n);
    // Measured on B200 (gpurun A/B, 4096 Hopper worlds): the tableau form (lcp_ppt) executes fewer
    // instructions but its per-row exchange branches cost more fetch stalls than they save for a lone
    // warp per SM (58.6 vs 53.8 us / env step), so the branch-free masked-Cholesky form stays the default.
    // (Also measured and rejected: stage 1 on a permuted leading block at half the register size — the
    // permutation through thread-local index arrays and the second instantiation cost more than the smaller
    // Cholesky saves: Hopper 65536 worlds 138 vs 118 us, HalfCheetah 16384 worlds 644 vs 560 us.)
#if defined(DARTB_LCP_PPT)
#define LCP_REG lcp_ppt
#else
#define LCP_REG lcp_small
#endif
#if DARTB_LCP_FORM == 2
    // (n > 8: six or more capsules on the ground, A rank-deficient up to the CFM: the Gauss-Jordan tableau
    // loses ~1e-3 there in fp32, the Cholesky-based pivoting below does not)
    if (n <= 8) { ok = lcp_ppt_loop<R, NR>(n, A, x, b, lo, hi, fidx, hin, sout); if (ok) EMU_COUNT(1, 1); }
    (void)nmax;
#elif DARTB_LCP_FORM == 1
    (void)nmax;
#else
    if (nmax <= 4) { ok = LCP_REG<R, 4>(n, A, x, b, lo, hi, fidx, hin, sout); if (ok) EMU_COUNT(1, 1); }
    else if (nmax <= 6 && NR > 4) { ok = LCP_REG<R, 6>(n, A, x, b, lo, hi, fidx, hin, sout); if (ok) EMU_COUNT(1, 1); }
    else if (NR > 6) {
        if (n <= 8) { ok = LCP_REG<R, 8>(n, A, x, b, lo, hi, fidx, hin, sout); if (ok) EMU_COUNT(2, 1); }
    }
#endif
#undef LCP_REG
    if (!ok) {
        if (sout) for (int i = 0; i < n; i++) sout[i] = 3;
        ok = lcp_bpp_local<R, NR>(n, A, x, b, lo, hi, fidx);
        if (ok) EMU_COUNT(5, 1);
    }
    if (!ok) { EMU_COUNT(3, 1); lcp_dantzig<R, NR>(n, A, x, b, lo, hi, fidx); }
}

// fixed-sweep PGS (oracle/dart_oracle.c::orc_solve_lcp_pgs; DART PGSLCPSolver shape)
template <typename R>
DEVI void lcp_pgs(int n, const R* A, R* x, const R* b, const R* lo, const R* hi, const int* fidx, int iters) {
    for (int i = 0; i < n; i++) x[i] = 0;
    for (int it = 0; it < iters; it++)
        for (int i = 0; i < n; i++) {
            const R aii = A[i * n + i];
            if (aii < (R)1e-9) { x[i] = 0; continue; }
            R s = b[i];
            for (int j = 0; j < n; j++) if (j != i) s -= A[i * n + j] * x[j];
            s /= aii;
            R l = lo[i], h = hi[i];
            if (fidx[i] >= 0) { h = hi[i] * x[fidx[i]]; l = -h; }
            if (s > h) s = h;
            if (s < l) s = l;
            x[i] = s;
        }
}

// The same sweeps with the whole problem in REGISTERS for n <= NM (fully unrolled rows, rolled sweep loop whose body
// stays in the L0 instruction cache).  r2 measurement (gpurun_out/r2a_sweep.log): the thread-local form above made the
// PGS path SLOWER than the exact solver (Walker2d 16384 worlds: 390 vs 165 us per env step; every A[i*n+j] is a
// dependent L1 round trip for a lone warp).  Same iterates as lcp_pgs up to the rounding of s * (1 / a_ii) vs s / a_ii.
template <typename R, int NM>
DEVI void pgs_small(int n, const R* Ag, R* xg, const R* bg, const R* log_, const R* hig, const int* fidxg, int iters) {
    R A[NM][NM], b[NM], lo[NM], hi[NM], x[NM], inv[NM];
    bool isf[NM];
    bool adjacent = true;
#pragma unroll
    for (int i = 0; i < NM; i++) {
        const bool on = i < n;
        const int fi = on ? fidxg[i] : -1;
        isf[i] = fi >= 0;
        if (fi >= 0 && fi != i - 1) adjacent = false;
        b[i] = on ? bg[i] : (R)0; lo[i] = on ? log_[i] : (R)0; hi[i] = on ? hig[i] : (R)0;
        x[i] = 0;
        R aii = 0;
#pragma unroll
        for (int j = 0; j < NM; j++) { A[i][j] = (on && j < n) ? Ag[i * n + j] : (R)0; if (j == i) aii = A[i][j]; }
        const bool live = on && !(aii < (R)1e-9);
        inv[i] = live ? Num<R>::rcp_(aii) : (R)0;   // (one MUFU.RCP: the IEEE division carries a slow-path branch per row)
        if (!live) { lo[i] = 0; hi[i] = 0; }     // padding / inert row (lcp_pgs: aii < 1e-9): the clamp keeps x = 0
    }
    if (!adjacent) { lcp_pgs<R>(n, Ag, xg, bg, log_, hig, fidxg, iters); return; }   // (not produced by this kernel's row layout)
    // A Gauss-Seidel sweep is ONE dependent chain through the rows.  Each row first sums everything that does not depend
    // on the row just before it (those x are older: the scheduler overlaps that part with the previous rows), and only
    // then adds the newest term: the chain per row is FMA -> FMUL -> min -> max instead of the whole dot product.
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NM; i++) {
            const int p = (i + NM - 1) % NM;      // the row updated just before row i
            R s = b[i];
            // oldest values first (rows after i: still the previous sweep's), then this sweep's in the order they appeared
#pragma unroll
            for (int j = i + 1; j < NM; j++) if (j != p) s -= A[i][j] * x[j];
#pragma unroll
            for (int j = 0; j < i; j++) if (j != p) s -= A[i][j] * x[j];
            if (p != i) s -= A[i][p] * x[p];
            s *= inv[i];
            // friction bounds +-mu x_n: a friction row directly follows its normal row (checked at load), so the sweep
            // body has no data-dependent branch (a lone warp pays ~20 cycles per branch)
            const R hf = hi[i] * (i > 0 ? x[i > 0 ? i - 1 : 0] : (R)0);
            const R h = isf[i] ? hf : hi[i], l = isf[i] ? -hf : lo[i];
            x[i] = Num<R>::max_(Num<R>::min_(s, h), l);     // (FMNMX; same result as the compare-and-select clamp for finite s)
        }
    }
#pragma unroll
    for (int i = 0; i < NM; i++) if (i < n) xg[i] = x[i];
}
// PGS dispatch: the register form per WARP size class (one code path per warp), thread-local sweeps above 8 rows
template <typename R, int NR>
DEVI void lcp_pgs_dispatch(int n, const R* A, R* x, const R* b, const R* lo, const R* hi, const int* fidx, int iters) {
    const int nmax = warp_max_active(n);
    if (nmax <= 4) { pgs_small<R, 4>(n, A, x, b, lo, hi, fidx, iters); return; }
    if constexpr (NR > 4) {
        if (nmax <= 6) { pgs_small<R, 6>(n, A, x, b, lo, hi, fidx, iters); return; }
    }
    if constexpr (NR > 6) {
        if (nmax <= 8) { pgs_small<R, 8>(n, A, x, b, lo, hi, fidx, iters); return; }
    }
    if constexpr (NR > 8) lcp_pgs<R>(n, A, x, b, lo, hi, fidx, iters);
}

// ------------------------------------------------------------------------ per-thread contact record
template <typename R>
struct ContactSink {      // where the LAST sub-step's contacts go (dartb_get_contacts); may be null
    int32_t* count;       // [n]
    int32_t* body;        // [n, maxc]
    float* data;          // [n, maxc, 10]
    int maxc;
};

// ------------------------------------------------------------------------ kernel arguments
template <typename R>
struct StepArgs {
    int n;
    R* q;                // [nd][n]
    R* dq;               // [nd][n]
    uint32_t* episode;   // [n] reset counter (Philox stream position)
    int32_t* elapsed;    // [n] env steps since reset (TimeLimit)
    uint8_t* truncated;  // [n]
    uint64_t* hint;      // [n] LCP active-set warm start (2 bits per constraint slot), all ones = none
    const float* action; // [n, n_act]
    float* obs;          // [n, n_obs]
    float* reward;       // [n]
    uint8_t* done;       // [n]
    double* reward64;    // gym return types (sync_vector_env.py:44-47): when set, rewards go HERE as float64 and `done`
                         // receives plain 0/1 bools (the truncation flag only in `truncated`)
    R* aux;              // [3][n] per-world task state (DARTB_TASK_REACHER2D target: world x, y, z) or null
    R* wpar;             // [4 nb + ns][n] per-world dynamics parameters (dartb_set_body_params: mass, cx, cy, izz per planar
                         // body, friction per capsule) or null = the model's; read by the loop kernels only
    float rand_mass, rand_mu;   // DARTB_OPT_RANDOMIZE_*: half ranges of the per-reset redraw of wpar's mass / friction rows
    const uint8_t* mask; // reset mask (k_reset) or null
    int auto_reset, lcp_mode, pgs_iters, max_episode_steps;
    int wpw;             // worlds per warp in k_env_step (1..32): lanes >= wpw idle, see dartb.cu::wpw_for
    int tma;             // k_env_step_coop: 1 = stage the CTA tile (lane table, q, dq) with TMA bulk copies and store the obs tile
                         // with one, 2 = the actions too; 0 = plain loads / stores (set by the launcher, see inst.cu)
    uint64_t seed;
    int64_t world_offset;
    const uint64_t* seeds;   // [n] per-world seeds (VectorEnv.seed(list), sync_vector_env.py:50-57) or null
    // fused observation all-gather (dartb_set_obs_peers): every observation row is also stored into these buffers
    // (peer GPUs' memory over NVLink, or local) at float offset obs_peer_off + its offset in `obs`
    float* obs_peer[DARTB_MAX_PEERS];
    int n_obs_peers;
    long long obs_peer_off;
    ContactSink<R> sink;
};
template <typename R>
DEVI void store_obs(const StepArgs<R>& a, size_t idx, float v) {
    a.obs[idx] = v;
    for (int p = 0; p < a.n_obs_peers; p++) a.obs_peer[p][a.obs_peer_off + (long long)idx] = v;
}
// Philox key of world w's reset draws: (seed, global world id), or (its own seed, 0) after a per-world seeding, so that
// world i seeded s_i draws what a single env seeded s_i draws
template <typename R>
DEVI uint64_t reset_seed(const StepArgs<R>& a, int w) { return a.seeds ? a.seeds[w] : a.seed; }
template <typename R>
DEVI int64_t reset_world(const StepArgs<R>& a, int w) { return a.seeds ? (int64_t)0 : a.world_offset + w; }

// ------------------------------------------------------------------------ kinematics only
// positions/orientations for the task layer (height of a body COM)
template <class T, typename R>
DEVI void fk_positions(const PModel<R>& M, const R (&q)[T::NB], R (&cs)[T::NB], R (&sn)[T::NB], R (&px)[T::NB],
                       R (&py)[T::NB]) {
    R th[T::NB];
    static_for<0, T::NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int par = T::parent(i);
        R cp = 1, sp = 0, ppx = 0, ppy = 0, thp = 0;
        if constexpr (par >= 0) { cp = cs[par]; sp = sn[par]; ppx = px[par]; ppy = py[par]; thp = th[par]; }
        const R arx = cp * M.ax[i] - sp * M.ay[i], ary = sp * M.ax[i] + cp * M.ay[i];
        if constexpr (T::jtype(i) == PM_REV) {
            th[i] = thp + M.sgn[i] * q[i];
            Num<R>::sincos_(th[i], &sn[i], &cs[i]);
            px[i] = ppx + arx; py[i] = ppy + ary;
        } else {
            th[i] = thp; cs[i] = cp; sn[i] = sp;
            const R uwx = cp * M.ux[i] - sp * M.uy[i], uwy = sp * M.ux[i] + cp * M.uy[i];
            px[i] = ppx + arx + uwx * q[i]; py[i] = ppy + ary + uwy * q[i];
        }
    });
}

#include "warp_group.cuh"   // group collectives + the row-per-lane LCP used by the quad form (G = 4) below

// ------------------------------------------------------------------------ one DART time step
// q, dq: in/out.  tau: generalized forces.  (eft, efx, efy): external spatial force per planar
// body [torque about the body origin; fx; fy] in world axes (only read when FEXT).
// FLUID: compute the snake fluid force (snake_7link.py:35-47) from the pre-step state instead.
// G = 1: one world per thread.  G = 4 (the QUAD form): the four lanes of a group hold the SAME world — K1-K4 and the row
// assembly run redundantly and stay bit-identical across the group — and share the constraint phase: row r lives on lane
// r % 4 (its impulse pass, its row of A = J M^-1 J^T, its row of the LCP tableau: GroupLcp<4>), dq += M^-1 J^T x is a
// group sum.  A batch of 16384 worlds is then 2048 warps instead of 512 (3.5 per scheduler instead of < 1) and a warp
// runs the union of 8 worlds' branches instead of 32.
template <class T, typename R, bool FEXT, bool FLUID, int G = 1>
DEVI void substep(const PModel<R>& M, R (&q)[T::NB], R (&dq)[T::NB], const R (&tau)[T::NB], const R (&eft)[T::NB],
                  const R (&efx)[T::NB], const R (&efy)[T::NB], R fluid_offset, R fluid_coef, int lcp_mode,
                  int pgs_iters, const ContactSink<R>* sink, int world, uint64_t& hint) {
    // `hint`: 2 bits per constraint SLOT (contact s -> slots 2s, 2s+1; limit of dof i -> 2*NS + i):
    // the set (0 free, 1 at lo, 2 at hi, 3 unknown) the slot's row ended in at the previous step.
    constexpr int NB = T::NB, NS = T::NS, NR = T::NR;
    const R dt = M.dt;
    // ---------------- K1: forward kinematics, velocities, partial accelerations
    R th[NB], cs[NB], sn[NB], px[NB], py[NB], rx[NB], ry[NB], wz[NB], vx[NB], vy[NB], ex[NB], ey[NB], uwx[NB], uwy[NB];
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int par = T::parent(i);
        R cp = 1, sp = 0, ppx = 0, ppy = 0, wp = 0, vpx = 0, vpy = 0, thp = 0;
        if constexpr (par >= 0) { cp = cs[par]; sp = sn[par]; ppx = px[par]; ppy = py[par]; wp = wz[par]; vpx = vx[par]; vpy = vy[par]; thp = th[par]; }
        const R arx = cp * M.ax[i] - sp * M.ay[i], ary = sp * M.ax[i] + cp * M.ay[i];
        if constexpr (T::jtype(i) == PM_REV) {
            th[i] = thp + M.sgn[i] * q[i];
            Num<R>::sincos_(th[i], &sn[i], &cs[i]);
            rx[i] = arx; ry[i] = ary;
            const R sd = M.sgn[i] * dq[i];
            wz[i] = wp + sd;
            vx[i] = vpx - wp * ary; vy[i] = vpy + wp * arx;
            ex[i] = sd * vy[i]; ey[i] = -sd * vx[i];
            uwx[i] = 0; uwy[i] = 0;
        } else {
            th[i] = thp; cs[i] = cp; sn[i] = sp;
            uwx[i] = cp * M.ux[i] - sp * M.uy[i]; uwy[i] = sp * M.ux[i] + cp * M.uy[i];
            rx[i] = arx + uwx[i] * q[i]; ry[i] = ary + uwy[i] * q[i];
            wz[i] = wp;
            vx[i] = vpx - wp * ry[i] + uwx[i] * dq[i]; vy[i] = vpy + wp * rx[i] + uwy[i] * dq[i];
            ex[i] = -wp * uwy[i] * dq[i]; ey[i] = wp * uwx[i] * dq[i];
        }
        px[i] = ppx + rx[i]; py[i] = ppy + ry[i];
    });

    // ---------------- K2: bias forces + implicit articulated inertia (leaves -> root)
    R U0[NB], U1[NB], U2[NB], Di[NB], uu[NB];
    {
        R aJ[NB], ahx[NB], ahy[NB], ama[NB], amb[NB], amc[NB], apt[NB], apx[NB], apy[NB];
        static_for<0, NB>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            aJ[i] = 0; ahx[i] = 0; ahy[i] = 0; ama[i] = 0; amb[i] = 0; amc[i] = 0; apt[i] = 0; apx[i] = 0; apy[i] = 0;
        });
        static_rfor<NB>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr int par = T::parent(i);
            const R m = M.mass[i];
            const R dx = cs[i] * M.cx[i] - sn[i] * M.cy[i], dy = sn[i] * M.cx[i] + cs[i] * M.cy[i];
            R J = M.izz[i] + m * (dx * dx + dy * dy), hx = -m * dy, hy = m * dx, ma = m, mb = 0, mc = m;
            const R Px = m * (vx[i] - wz[i] * dy), Py = m * (vy[i] + wz[i] * dx);
            R pt = vx[i] * Py - vy[i] * Px - (dx * m * M.gy - dy * m * M.gx);
            R pfx = -wz[i] * Py - m * M.gx, pfy = wz[i] * Px - m * M.gy;
            if constexpr (FEXT) { pt -= eft[i]; pfx -= efx[i]; pfy -= efy[i]; }
            if constexpr (FLUID) {
                // bn.com_spatial_velocity(), norm_dir = R*ez, add_ext_force at the body origin
                const R nx = cs[i] * M.fnx[i] - sn[i] * M.fny[i], ny = sn[i] * M.fnx[i] + cs[i] * M.fny[i];
                // COM velocity in BODY coordinates: pydart2's no-argument com_spatial_velocity() is DART's
                // BodyNode::getCOMSpatialVelocity(), expressed in the body frame; the reference dots it with the
                // WORLD-frame norm_dir component by component (snake_7link.py:37-45), and so do we
                const R vwx = vx[i] - wz[i] * dy, vwy = vy[i] + wz[i] * dx;  // COM velocity, world axes
                const R vcx = cs[i] * vwx + sn[i] * vwy, vcy = cs[i] * vwy - sn[i] * vwx;
                const R crx = -wz[i] * ny, cry = wz[i] * nx;                 // omega x n (omega is the same in both frames)
                const R dp = (vcx + crx * fluid_offset) * nx + (vcy + cry * fluid_offset) * ny;
                const R dn = (vcx - crx * fluid_offset) * nx + (vcy - cry * fluid_offset) * ny;
                R ffx = 0, ffy = 0;
                if (dp > 0) { ffx = -fluid_coef * dp * nx; ffy = -fluid_coef * dp * ny; }
                if (dn < 0) { ffx = -fluid_coef * dn * nx; ffy = -fluid_coef * dn * ny; }
                const R oxw = cs[i] * M.ox[i] - sn[i] * M.oy[i], oyw = sn[i] * M.ox[i] + cs[i] * M.oy[i];
                pt -= oxw * ffy - oyw * ffx; pfx -= ffx; pfy -= ffy;
            }
            if constexpr (topo_has_child<T>(i)) {
                J += aJ[i]; hx += ahx[i]; hy += ahy[i]; ma += ama[i]; mb += amb[i]; mc += amc[i];
                pt += apt[i]; pfx += apx[i]; pfy += apy[i];
            }
            const R t0 = hx * ex[i] + hy * ey[i], t1 = ma * ex[i] + mb * ey[i], t2 = mb * ex[i] + mc * ey[i];
            R D, u;
            if constexpr (T::jtype(i) == PM_REV) {
                const R s = M.sgn[i];
                U0[i] = s * J; U1[i] = s * hx; U2[i] = s * hy;
                D = J;
                u = tau[i] - s * (pt + t0);
            } else {
                U0[i] = hx * uwx[i] + hy * uwy[i]; U1[i] = ma * uwx[i] + mb * uwy[i]; U2[i] = mb * uwx[i] + mc * uwy[i];
                D = uwx[i] * U1[i] + uwy[i] * U2[i];
                u = tau[i] - (uwx[i] * (pfx + t1) + uwy[i] * (pfy + t2));
            }
            u += -M.kspring[i] * (q[i] - M.rest[i] + dt * dq[i]) - M.damping[i] * dq[i];
            D += dt * M.damping[i] + dt * dt * M.kspring[i];
            const R di = Num<R>::rcp_(D);
            Di[i] = di; uu[i] = u;
            if constexpr (par >= 0) {
                const R g = u * di;
                const R pa0 = pt + t0 + U0[i] * g, pa1 = pfx + t1 + U1[i] * g, pa2 = pfy + t2 + U2[i] * g;
                const R P00 = J - U0[i] * U0[i] * di, P01 = hx - U0[i] * U1[i] * di, P02 = hy - U0[i] * U2[i] * di;
                const R P11 = ma - U1[i] * U1[i] * di, P12 = mb - U1[i] * U2[i] * di, P22 = mc - U2[i] * U2[i] * di;
                const R kx = -ry[i], ky = rx[i];
                const R nhx = P01 + P11 * kx + P12 * ky, nhy = P02 + P12 * kx + P22 * ky;
                aJ[par] += P00 + kx * (P01 + nhx) + ky * (P02 + nhy);
                ahx[par] += nhx; ahy[par] += nhy; ama[par] += P11; amb[par] += P12; amc[par] += P22;
                apt[par] += pa0 + kx * pa1 + ky * pa2; apx[par] += pa1; apy[par] += pa2;
            }
        });
    }
    // ---------------- K3: accelerations (root -> leaves), dq += dt * ddq
    {
        R a0[NB], a1[NB], a2[NB];
        static_for<0, NB>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr int par = T::parent(i);
            R p0 = 0, p1 = 0, p2 = 0;
            if constexpr (par >= 0) { p0 = a0[par]; p1 = a1[par] - a0[par] * ry[i]; p2 = a2[par] + a0[par] * rx[i]; }
            const R dd = Di[i] * (uu[i] - (U0[i] * p0 + U1[i] * p1 + U2[i] * p2));
            if constexpr (T::jtype(i) == PM_REV) { a0[i] = p0 + M.sgn[i] * dd; a1[i] = p1 + ex[i]; a2[i] = p2 + ey[i]; }
            else { a0[i] = p0; a1[i] = p1 + ex[i] + uwx[i] * dd; a2[i] = p2 + ey[i] + uwy[i] * dd; }
            dq[i] += dt * dd;
        });
    }

    // ---------------- K4/K5: collide + constraint rows
    int n = 0, nc = 0;
    R Jr[NR * NB], bb[NR], lo[NR], hi[NR];
    int fidx[NR];
    uint8_t rslot[NR];
    R cpx[T::NSA], cpy[T::NSA], cnx[T::NSA], cny[T::NSA], cdep[T::NSA];
    int crow[T::NSA], cshape[T::NSA];
    const R INF = Num<R>::inf();
#if DARTB_ROLLED_COLLIDE
    // K4 as ONE rolled loop over the capsules (the unrolled form below carries NS copies of the closest-point walk
    // and of the row assembly: 4.5 k of the 10.8 k distinct SASS instructions a HalfCheetah DART step executes, and
    // the per-thread kernels are bound by instruction fetch: ncu r2, 71 % of the no_inst samples sit on 128-byte line
    // starts).  The capsule frames come from the unrolled kinematics through small thread-local arrays; the Jacobian
    // entries of non-ancestor bodies are masked with the capsule's ancestor set instead of being compiled out.
    if constexpr (NS > 0) {
        if (M.has_ground) {
            const R inv_dt = (R)1 / dt;
            R capx[NS], capy[NS], capdx[NS], capdy[NS];
            static_for<0, NS>([&](auto sc) {
                constexpr int s = decltype(sc)::value;
                constexpr int b = T::sbody(s);
                capx[s] = px[b] + cs[b] * M.scx[s] - sn[b] * M.scy[s]; capy[s] = py[b] + sn[b] * M.scx[s] + cs[b] * M.scy[s];
                capdx[s] = cs[b] * M.sdx[s] - sn[b] * M.sdy[s]; capdy[s] = sn[b] * M.sdx[s] + cs[b] * M.sdy[s];
            });
            // narrow phase of capsule s -> contact point / normal / depth (ODE dCollideCapsuleBox in the plane)
            auto detect = [&](int s, R& Px, R& Py, R& nx, R& ny, R& depth) -> bool {
                const R ccx = capx[s], ccy = capy[s], adx = capdx[s], ady = capdy[s];
                const R hl = M.shalf[s], rad = M.srad[s];
                const R ex_ = hl * Num<R>::abs_(adx) + rad + (R)1e-5, ey_ = hl * Num<R>::abs_(ady) + rad + (R)1e-5;
                const bool near_ = Num<R>::abs_(ccx - M.gcx) <= M.ghx + ex_ && Num<R>::abs_(ccy - M.gcy) <= M.ghy + ey_;
                R lx = 0, ly = 0, ddx = 0, ddy = 0, d = INF;
                if (near_) {
                    closest_segment_box2<R>(ccx + hl * adx, ccy + hl * ady, ccx - hl * adx, ccy - hl * ady, M.gcx, M.gcy,
                                            M.ghx, M.ghy, lx, ly, ddx, ddy);
                    d = Num<R>::sqrt_(ddx * ddx + ddy * ddy);
                }
                if (d > rad) return false;
                if (!(d < Num<R>::mindist())) {
                    const R id = Num<R>::rcp_(d);
                    nx = ddx * id; ny = ddy * id;
                    depth = rad - d;
                    const R k = (R)0.5 * (-rad - d);
                    Px = lx + nx * k; Py = ly + ny * k;
                } else {
                    nx = M.gupx; ny = M.gupy;
                    depth = rad + (M.ghup - ((lx - M.gcx) * nx + (ly - M.gcy) * ny));
                    Px = lx; Py = ly;
                }
                return true;
            };
            // contact rows of capsule s: (normal, in-plane tangent)
            auto assemble = [&](int s, R Px, R Py, R nx, R ny, R depth) {
                const R mu = M.smu[s];
                const bool fric = mu > (R)DK_FRICTION_THRESHOLD;
                const R tx = -ny, ty = nx;
                const int r0 = n;
                unsigned anc = 0;   // ancestors-or-self of this capsule's body
                static_for<0, NS>([&](auto sc) {
                    constexpr int s2 = decltype(sc)::value;
                    constexpr unsigned m2 = [] { unsigned m = 0; for (int j = 0; j < T::NB; j++) if (topo_is_ancestor<T>(j, T::sbody(s2))) m |= 1u << j; return m; }();
                    if (s == s2) anc = m2;
                });
                R vn = 0, vt = 0;
                static_for<0, NB>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    R ax_, ay_;
                    if constexpr (T::jtype(j) == PM_REV) { ax_ = -M.sgn[j] * (Py - py[j]); ay_ = M.sgn[j] * (Px - px[j]); }
                    else { ax_ = uwx[j]; ay_ = uwy[j]; }
                    const bool on = (anc >> j) & 1u;
                    const R jn = on ? ax_ * nx + ay_ * ny : (R)0, jt = on ? ax_ * tx + ay_ * ty : (R)0;
                    if (on) { vn += jn * dq[j]; vt += jt * dq[j]; }
                    Jr[r0 * NB + j] = jn;
                    if (fric) Jr[(r0 + 1) * NB + j] = jt;
                });
                R bounce = depth;
                if (bounce < 0) bounce = 0;
                else { bounce *= inv_dt * (R)DK_CONTACT_ERP; if (bounce > (R)DK_CONTACT_MAX_ERV) bounce = (R)DK_CONTACT_MAX_ERV; }
                bb[r0] = -vn + bounce; lo[r0] = 0; hi[r0] = INF; fidx[r0] = -1; rslot[r0] = 2 * s;
                n = r0 + 1;
                if (fric) { bb[r0 + 1] = -vt; lo[r0 + 1] = -mu; hi[r0 + 1] = mu; fidx[r0 + 1] = r0; rslot[r0 + 1] = 2 * s + 1; n = r0 + 2; }
                cpx[nc] = Px; cpy[nc] = Py; cnx[nc] = nx; cny[nc] = ny; cdep[nc] = depth; crow[nc] = r0 | (fric ? 0x100 : 0);
                cshape[nc] = s;
                nc++;
            };
            if constexpr (G > 1 && DARTB_QUAD_SPLIT_COLLIDE && NS <= G) {   // (two capsules per lane measured slower: r2_experiments.md section 8)
                // quad form: the narrow phase is dealt out over the lanes (capsule s on lane s % G), the contacts are
                // then gathered in capsule order so that every lane assembles the same rows
                constexpr int KS = (NS + G - 1) / G;
                const int l = (threadIdx.x & 31) % G;
                R rPx[KS], rPy[KS], rnx[KS], rny[KS], rdep[KS];
                int rhas[KS];
#pragma unroll
                for (int k = 0; k < KS; k++) {
                    const int s = l + k * G;
                    rPx[k] = 0; rPy[k] = 0; rnx[k] = 0; rny[k] = 0; rdep[k] = 0; rhas[k] = 0;
                    if (s < NS) rhas[k] = detect(s, rPx[k], rPy[k], rnx[k], rny[k], rdep[k]) ? 1 : 0;
                }
#pragma unroll 1
                for (int s = 0; s < NS; s++) {
                    const int o = s % G, k = s / G;
                    R sPx = rPx[0], sPy = rPy[0], snx = rnx[0], sny = rny[0], sdep = rdep[0];
                    int shas = rhas[0];
#pragma unroll
                    for (int k2 = 1; k2 < KS; k2++) if (k == k2) { sPx = rPx[k2]; sPy = rPy[k2]; snx = rnx[k2]; sny = rny[k2]; sdep = rdep[k2]; shas = rhas[k2]; }
                    const int has = gshfl<G>(shas, o);
                    const R Px = gshfl<G>(sPx, o), Py = gshfl<G>(sPy, o), nx = gshfl<G>(snx, o), ny = gshfl<G>(sny, o), depth = gshfl<G>(sdep, o);
                    if (has) assemble(s, Px, Py, nx, ny, depth);
                }
            } else {
#pragma unroll 1
                for (int s = 0; s < NS; s++) {
                    R Px, Py, nx, ny, depth;
                    if (detect(s, Px, Py, nx, ny, depth)) assemble(s, Px, Py, nx, ny, depth);
                }
            }
        }
    }
#else
    if constexpr (NS > 0) {
        if (M.has_ground) {
            const R inv_dt = (R)1 / dt;
            static_for<0, NS>([&](auto sc) {
                constexpr int s = decltype(sc)::value;
                constexpr int b = T::sbody(s);
                const R ccx = px[b] + cs[b] * M.scx[s] - sn[b] * M.scy[s], ccy = py[b] + sn[b] * M.scx[s] + cs[b] * M.scy[s];
                const R adx = cs[b] * M.sdx[s] - sn[b] * M.sdy[s], ady = sn[b] * M.sdx[s] + cs[b] * M.sdy[s];
                const R hl = M.shalf[s], rad = M.srad[s];
                // broad phase: capsule AABB (segment box inflated by r) against the ground box AABB.
                // Disjoint by more than 10 um => distance - r > 0 => no contact (exact reject); only
                // overlapping pairs pay for the ODE closest-point walk.
                const R ex_ = hl * Num<R>::abs_(adx) + rad + (R)1e-5, ey_ = hl * Num<R>::abs_(ady) + rad + (R)1e-5;
                const bool near_ = Num<R>::abs_(ccx - M.gcx) <= M.ghx + ex_ && Num<R>::abs_(ccy - M.gcy) <= M.ghy + ey_;
                R lx = 0, ly = 0, ddx = 0, ddy = 0, d = INF;
                if (near_) {
                    closest_segment_box2<R>(ccx + hl * adx, ccy + hl * ady, ccx - hl * adx, ccy - hl * ady, M.gcx, M.gcy,
                                            M.ghx, M.ghy, lx, ly, ddx, ddy);
                    d = Num<R>::sqrt_(ddx * ddx + ddy * ddy);
                }
                if (!(d > rad)) {
                    R nx, ny, depth, Px, Py;
                    if (!(d < Num<R>::mindist())) {  // ODE dCollideCapsuleBox: pl == pb up to mindist
                        const R id = Num<R>::rcp_(d);
                        nx = ddx * id; ny = ddy * id;
                        depth = rad - d;
                        const R k = (R)0.5 * (-rad - d);
                        Px = lx + nx * k; Py = ly + ny * k;
                    } else {  // axis inside the box: push out through the box's local +y face
                        nx = M.gupx; ny = M.gupy;
                        depth = rad + (M.ghup - ((lx - M.gcx) * nx + (ly - M.gcy) * ny));
                        Px = lx; Py = ly;
                    }
                    const R mu = M.smu[s];
                    const bool fric = mu > (R)DK_FRICTION_THRESHOLD;
                    const R tx = -ny, ty = nx;  // DART tangent t1 = z x n (in-plane); t2 is out of plane (inert)
                    const int r0 = n;
                    R vn = 0, vt = 0;
                    static_for<0, NB>([&](auto jc) {
                        constexpr int j = decltype(jc)::value;
                        R jn = 0, jt = 0;
                        if constexpr (topo_is_ancestor<T>(j, b)) {
                            R ax_, ay_;
                            if constexpr (T::jtype(j) == PM_REV) { ax_ = -M.sgn[j] * (Py - py[j]); ay_ = M.sgn[j] * (Px - px[j]); }
                            else { ax_ = uwx[j]; ay_ = uwy[j]; }
                            jn = ax_ * nx + ay_ * ny; jt = ax_ * tx + ay_ * ty;
                            vn += jn * dq[j]; vt += jt * dq[j];
                        }
                        Jr[r0 * NB + j] = jn;
                        if (fric) Jr[(r0 + 1) * NB + j] = jt;
                    });
                    R bounce = depth;
                    if (bounce < 0) bounce = 0;
                    else { bounce *= inv_dt * (R)DK_CONTACT_ERP; if (bounce > (R)DK_CONTACT_MAX_ERV) bounce = (R)DK_CONTACT_MAX_ERV; }
                    bb[r0] = -vn + bounce; lo[r0] = 0; hi[r0] = INF; fidx[r0] = -1; rslot[r0] = 2 * s;
                    n = r0 + 1;
                    if (fric) { bb[r0 + 1] = -vt; lo[r0 + 1] = -mu; hi[r0 + 1] = mu; fidx[r0 + 1] = r0; rslot[r0 + 1] = 2 * s + 1; n = r0 + 2; }
                    cpx[nc] = Px; cpy[nc] = Py; cnx[nc] = nx; cny[nc] = ny; cdep[nc] = depth; crow[nc] = r0 | (fric ? 0x100 : 0);
                    cshape[nc] = s;
                    nc++;
                }
            });
        }
    }
#endif
    const int n_contact_rows = n;
    // joint-limit rows: q BEFORE this step's integration, dq AFTER the unconstrained update
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        if (M.limited[i]) {
            int act = 0;
            if (q[i] - M.qlo[i] <= 0) act = -1;
            else if (q[i] - M.qhi[i] >= 0) act = 1;
            if (act != 0 && n < NR) {
                static_for<0, NB>([&](auto jc) { constexpr int j = decltype(jc)::value; Jr[n * NB + j] = (j == i) ? (R)1 : (R)0; });
                bb[n] = -dq[i];
                if (act < 0) { lo[n] = 0; hi[n] = INF; } else { lo[n] = -INF; hi[n] = 0; }
                fidx[n] = -1;
                rslot[n] = 2 * NS + i;
                n++;
            }
        }
    });

    uint64_t new_hint = ~(uint64_t)0;
    // quad form: the shared constraint phase is entered by the WHOLE warp (its collectives need every lane) whenever any
    // of its worlds has 1..8 rows; a world with more rows (a fallen walker) sits that phase out (nq = 0) and then takes the
    // per-thread path on its four lanes, so a world's arithmetic never depends on which worlds share its warp
    const int nq = (G > 1 && n <= 8) ? n : 0;
    int nmaxq = 0;
    bool quad = false;
    if constexpr (G > 1) { nmaxq = coop_warp_max(nq); quad = nmaxq > 0; }
    if (n > 0 || quad) {
        // plain (non-implicit) articulated inertia for the impulse passes
        R V0[NB], V1[NB], V2[NB], Ei[NB];
        {
            R aJ[NB], ahx[NB], ahy[NB], ama[NB], amb[NB], amc[NB];
            static_for<0, NB>([&](auto ic) { constexpr int i = decltype(ic)::value; aJ[i] = 0; ahx[i] = 0; ahy[i] = 0; ama[i] = 0; amb[i] = 0; amc[i] = 0; });
            static_rfor<NB>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                constexpr int par = T::parent(i);
                const R m = M.mass[i];
                const R dx = cs[i] * M.cx[i] - sn[i] * M.cy[i], dy = sn[i] * M.cx[i] + cs[i] * M.cy[i];
                R J = M.izz[i] + m * (dx * dx + dy * dy), hx = -m * dy, hy = m * dx, ma = m, mb = 0, mc = m;
                if constexpr (topo_has_child<T>(i)) { J += aJ[i]; hx += ahx[i]; hy += ahy[i]; ma += ama[i]; mb += amb[i]; mc += amc[i]; }
                R D;
                if constexpr (T::jtype(i) == PM_REV) { const R s = M.sgn[i]; V0[i] = s * J; V1[i] = s * hx; V2[i] = s * hy; D = J; }
                else {
                    V0[i] = hx * uwx[i] + hy * uwy[i]; V1[i] = ma * uwx[i] + mb * uwy[i]; V2[i] = mb * uwx[i] + mc * uwy[i];
                    D = uwx[i] * V1[i] + uwy[i] * V2[i];
                }
                const R di = Num<R>::rcp_(D);
                Ei[i] = di;
                if constexpr (par >= 0) {
                    const R P00 = J - V0[i] * V0[i] * di, P01 = hx - V0[i] * V1[i] * di, P02 = hy - V0[i] * V2[i] * di;
                    const R P11 = ma - V1[i] * V1[i] * di, P12 = mb - V1[i] * V2[i] * di, P22 = mc - V2[i] * V2[i] * di;
                    const R kx = -ry[i], ky = rx[i];
                    const R nhx = P01 + P11 * kx + P12 * ky, nhy = P02 + P12 * kx + P22 * ky;
                    aJ[par] += P00 + kx * (P01 + nhx) + ky * (P02 + nhy);
                    ahx[par] += nhx; ahy[par] += nhy; ama[par] += P11; amb[par] += P12; amc[par] += P22;
                }
            });
        }
        R A[NR * NR], x[NR];
        if constexpr (G > 1) {
          if (quad) {
            const int l = (threadIdx.x & 31) % G, gbase = (threadIdx.x & 31) - l;
            auto quad_tail = [&](auto ncc) {
                constexpr int NCx = decltype(ncc)::value, RPL = (NCx + G - 1) / G;
                R MJq[RPL][NB], Aq[RPL][NCx], bq[RPL], loq[RPL], hiq[RPL], xq[RPL];
                int fiq[RPL];
                unsigned hinq[RPL], stq[RPL];
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    const int r = l + h * G;
                    const bool valid = r < nq;
                    const int rr = valid ? r : 0;
                    // this lane's impulse pass (DART: applyUnitImpulse + getVelocityChange)
                    R rh[NB], ur[NB], apt[NB], apx[NB], apy[NB], ddr[NB];
                    static_for<0, NB>([&](auto ic) { constexpr int i = decltype(ic)::value; rh[i] = valid ? Jr[rr * NB + i] : (R)0; apt[i] = 0; apx[i] = 0; apy[i] = 0; });
                    static_rfor<NB>([&](auto ic) {
                        constexpr int i = decltype(ic)::value;
                        constexpr int par = T::parent(i);
                        R pt = 0, pfx = 0, pfy = 0;
                        if constexpr (topo_has_child<T>(i)) { pt = apt[i]; pfx = apx[i]; pfy = apy[i]; }
                        R u;
                        if constexpr (T::jtype(i) == PM_REV) u = rh[i] - M.sgn[i] * pt;
                        else u = rh[i] - (uwx[i] * pfx + uwy[i] * pfy);
                        ur[i] = u;
                        if constexpr (par >= 0) {
                            const R g = u * Ei[i];
                            const R pa0 = pt + V0[i] * g, pa1 = pfx + V1[i] * g, pa2 = pfy + V2[i] * g;
                            apt[par] += pa0 - ry[i] * pa1 + rx[i] * pa2; apx[par] += pa1; apy[par] += pa2;
                        }
                    });
                    R a0[NB], a1[NB], a2[NB];
                    static_for<0, NB>([&](auto ic) {
                        constexpr int i = decltype(ic)::value;
                        constexpr int par = T::parent(i);
                        R p0 = 0, p1 = 0, p2 = 0;
                        if constexpr (par >= 0) { p0 = a0[par]; p1 = a1[par] - a0[par] * ry[i]; p2 = a2[par] + a0[par] * rx[i]; }
                        const R dd = Ei[i] * (ur[i] - (V0[i] * p0 + V1[i] * p1 + V2[i] * p2));
                        if constexpr (T::jtype(i) == PM_REV) { a0[i] = p0 + M.sgn[i] * dd; a1[i] = p1; a2[i] = p2; }
                        else { a0[i] = p0; a1[i] = p1 + uwx[i] * dd; a2[i] = p2 + uwy[i] * dd; }
                        ddr[i] = dd;
                        MJq[h][i] = dd;
                    });
                    // its row of A = J M^-1 J^T (+ CFM on the diagonal): every J_s is known to every lane
#pragma unroll
                    for (int sidx = 0; sidx < NCx; sidx++) {
                        R v = 0;
                        if (sidx < nq) {
                            static_for<0, NB>([&](auto jc) { constexpr int j = decltype(jc)::value; v += Jr[sidx * NB + j] * ddr[j]; });
                            if (sidx == r) v *= (R)1 + (r < n_contact_rows ? (R)DK_CONTACT_CFM : (R)DK_LIMIT_CFM);
                        }
                        Aq[h][sidx] = valid ? v : (R)0;
                    }
                    bq[h] = valid ? bb[rr] : (R)0; loq[h] = valid ? lo[rr] : (R)0; hiq[h] = valid ? hi[rr] : (R)0;
                    fiq[h] = valid ? fidx[rr] : -1;
                    hinq[h] = valid ? (unsigned)((hint >> (2 * rslot[rr])) & 3u) : 3u;
                    xq[h] = 0; stq[h] = 3u;
                }
                bool ok = lcp_mode != 1;
                if (lcp_mode != 1) ok = GroupLcp<G, R, NCx>::solve(l, gbase, nq, nmaxq, Aq, bq, loq, hiq, fiq, hinq, xq, stq);
                // PGS mode, or a group whose pivoting did not converge (never observed): gather A and run the per-thread
                // solver on every lane of the group (identical inputs, identical results)
                if (__any_sync(COOP_FULL, !ok)) {
#pragma unroll
                    for (int rr2 = 0; rr2 < NCx; rr2++)
#pragma unroll
                        for (int sidx = 0; sidx < NCx; sidx++) {
                            const R v = gshfl<G>(Aq[rr2 / G][sidx], rr2 % G);
                            if (rr2 < nq && sidx < nq) A[rr2 * nq + sidx] = v;
                        }
                    if (!ok) {
                        if (lcp_mode == 1) lcp_pgs_dispatch<R, NR>(nq, A, x, bb, lo, hi, fidx, pgs_iters);
                        else lcp_exact<R, NR>(nq, A, x, bb, lo, hi, fidx);
#pragma unroll
                        for (int h = 0; h < RPL; h++) { const int r = l + h * G; xq[h] = r < nq ? x[r] : (R)0; stq[h] = 3u; }
                    }
                }
                // K7: dq += sum_r (M^-1 J_r^T) x_r, each lane's rows first, then over the group (same bits on every lane)
                static_for<0, NB>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    R v = 0;
#pragma unroll
                    for (int h = 0; h < RPL; h++) if (l + h * G < nq) v += MJq[h][i] * xq[h];
                    dq[i] += group_sum<G>(v);
                });
                // warm-start sets for the next step, and x of every row (contact read-back)
                uint64_t clr = 0, setb = 0;
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    const int r = l + h * G;
                    if (r < nq) { clr |= (uint64_t)3 << (2 * rslot[r]); setb |= (uint64_t)(stq[h] & 3u) << (2 * rslot[r]); }
                }
                const unsigned clo = group_or<G>((unsigned)clr), chi = group_or<G>((unsigned)(clr >> 32));
                const unsigned slo = group_or<G>((unsigned)setb), shi = group_or<G>((unsigned)(setb >> 32));
                if (nq > 0) new_hint = (~(((uint64_t)chi << 32) | clo)) | (((uint64_t)shi << 32) | slo);
#pragma unroll
                for (int rr2 = 0; rr2 < NCx; rr2++) { const R v = gshfl<G>(xq[rr2 / G], rr2 % G); if (rr2 < nq) x[rr2] = v; }
            };
            if (nmaxq <= 4) quad_tail(std::integral_constant<int, 4>{});
            else quad_tail(std::integral_constant<int, 8>{});
          }
        }
        if (n > nq) {   // G == 1, or a quad world with more than 8 rows
        // M^-1 J^T, one impulse pass per row (DART: applyUnitImpulse + getVelocityChange), and row r of
        // A = J M^-1 J^T formed right away from the pass result still in registers (lower triangle,
        // mirrored: A is symmetric)
        R MJ[NR * NB];
        for (int r = 0; r < n; r++) {
            R rh[NB], ur[NB], apt[NB], apx[NB], apy[NB], ddr[NB];
            static_for<0, NB>([&](auto ic) { constexpr int i = decltype(ic)::value; rh[i] = Jr[r * NB + i]; apt[i] = 0; apx[i] = 0; apy[i] = 0; });
            static_rfor<NB>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                constexpr int par = T::parent(i);
                R pt = 0, pfx = 0, pfy = 0;
                if constexpr (topo_has_child<T>(i)) { pt = apt[i]; pfx = apx[i]; pfy = apy[i]; }
                R u;
                if constexpr (T::jtype(i) == PM_REV) u = rh[i] - M.sgn[i] * pt;
                else u = rh[i] - (uwx[i] * pfx + uwy[i] * pfy);
                ur[i] = u;
                if constexpr (par >= 0) {
                    const R g = u * Ei[i];
                    const R pa0 = pt + V0[i] * g, pa1 = pfx + V1[i] * g, pa2 = pfy + V2[i] * g;
                    apt[par] += pa0 - ry[i] * pa1 + rx[i] * pa2; apx[par] += pa1; apy[par] += pa2;
                }
            });
            R a0[NB], a1[NB], a2[NB];
            static_for<0, NB>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                constexpr int par = T::parent(i);
                R p0 = 0, p1 = 0, p2 = 0;
                if constexpr (par >= 0) { p0 = a0[par]; p1 = a1[par] - a0[par] * ry[i]; p2 = a2[par] + a0[par] * rx[i]; }
                const R dd = Ei[i] * (ur[i] - (V0[i] * p0 + V1[i] * p1 + V2[i] * p2));
                if constexpr (T::jtype(i) == PM_REV) { a0[i] = p0 + M.sgn[i] * dd; a1[i] = p1; a2[i] = p2; }
                else { a0[i] = p0; a1[i] = p1 + uwx[i] * dd; a2[i] = p2 + uwy[i] * dd; }
                ddr[i] = dd;
                MJ[r * NB + i] = dd;
            });
            for (int s = 0; s <= r; s++) {
                R v = 0;
                static_for<0, NB>([&](auto jc) { constexpr int j = decltype(jc)::value; v += Jr[s * NB + j] * ddr[j]; });
                A[r * n + s] = v;
                A[s * n + r] = v;
            }
        }
        for (int r = 0; r < n; r++) A[r * n + r] *= (R)1 + (r < n_contact_rows ? (R)DK_CONTACT_CFM : (R)DK_LIMIT_CFM);
        if (lcp_mode == 1) lcp_pgs_dispatch<R, NR>(n, A, x, bb, lo, hi, fidx, pgs_iters);
        else {
            uint8_t hrow[NR], srow[NR];
            for (int r = 0; r < n; r++) hrow[r] = (uint8_t)((hint >> (2 * rslot[r])) & 3u);
            lcp_exact<R, NR>(n, A, x, bb, lo, hi, fidx, hrow, srow);
            for (int r = 0; r < n; r++)
                new_hint = (new_hint & ~((uint64_t)3 << (2 * rslot[r]))) | ((uint64_t)srow[r] << (2 * rslot[r]));
        }
        // ---------------- K7: apply impulses
        for (int r = 0; r < n; r++) {
            const R xr = x[r];
#pragma unroll
            for (int j = 0; j < NB; j++) dq[j] += MJ[r * NB + j] * xr;
        }
        }   // per-thread constraint phase
        if (sink && sink->data) {
            const R inv_dt = (R)1 / dt;
            for (int c = 0; c < nc && c < sink->maxc; c++) {
                const int r0 = crow[c] & 0xff;
                const R xn = x[r0], xt = (crow[c] & 0x100) ? x[r0 + 1] : (R)0;
                const R fx = (cnx[c] * xn - cny[c] * xt) * inv_dt, fy = (cny[c] * xn + cnx[c] * xt) * inv_dt;
                float* o = sink->data + ((size_t)world * sink->maxc + c) * 10;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    o[k] = (float)(M.e1[k] * cpx[c] + M.e2[k] * cpy[c] + M.en[k] * M.hz);
                    o[3 + k] = (float)(M.e1[k] * cnx[c] + M.e2[k] * cny[c]);
                    o[7 + k] = (float)(M.e1[k] * fx + M.e2[k] * fy);
                }
                o[6] = (float)cdep[c];
            }
        }
    }
    if (sink) {
        if (sink->count) sink->count[world] = nc;
        if (sink->body)
            for (int c = 0; c < sink->maxc; c++) sink->body[(size_t)world * sink->maxc + c] = c < nc ? M.sorig[cshape[c]] : -1;
    }
    hint = new_hint;
    // ---------------- integrate positions
    static_for<0, NB>([&](auto ic) { constexpr int i = decltype(ic)::value; q[i] += dt * dq[i]; });
}
