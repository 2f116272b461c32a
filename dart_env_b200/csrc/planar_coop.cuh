// planar_coop.cuh — the LANE-COOPERATIVE form of one DART time step: G = 8 or 16 lanes of a warp
// step ONE world together (lane i owns body/dof i, capsule i and LCP row i), so a 4096-world batch
// spreads over every warp scheduler of the B200 instead of 128 lone warps.
//
// Why (profiles/r1_hopper_v3_*.md): with one world per thread the headline batch is 128 warps on
// 592 schedulers, each executing ~4.5 k dependent instructions per DART step at IPC ~0.2 — pure
// single-warp latency.  The same arithmetic restated so that it parallelises along the chain:
//   K1  forward kinematics      = inclusive ANCESTOR PREFIX SUMS of planar increments (angles,
//       origin offsets, velocities, velocity-product accelerations): log2(depth) rounds of
//       warp shuffles by pointer jumping (hop 1, 2, 4, 8 ancestors), sincos in parallel
//   K2  bias forces             = recursive Newton-Euler with ddq = 0: per-body wrench, then
//       SUBTREE sums (masked all-gather over the group) projected on each joint axis
//   K2' joint-space inertia     = composite-rigid-body: subtree sums of the body inertias taken about
//       each receiving joint's own origin (no cancellation), M_ij = S_j . (Ic_i S_i)
//   K3  (M + dt D + dt^2 K) ddq = tau - c ...: DART's implicit joint damping / spring adds to the
//       diagonal pivot of each joint, i.e. to diag(M); solved by the sparse L^T L factorisation
//       (Featherstone's LTL: no fill-in for the tree), unrolled at compile time
//   K4  capsule s vs ground box on lane s (the ODE walk of planar_kernels.cuh, one shape per lane)
//   K5  rows compacted through shared memory; row r lives on lane r: J_r, M^-1 J_r^T = L^-1 L^-T J_r^T,
//       A[r][:] = J_s . (M^-1 J_r^T) with the fp32 rows J_s fetched by shuffles: DART's impulse tests
//   K6  boxed LCP: the block-principal-pivoting iteration of planar_kernels.cuh::lcp_ppt with the
//       tableau distributed one row per lane (exchange = one shuffle round + one FMA per column)
//   K7  dq += sum_r (M^-1 J_r^T) x_r (group sum) ; q += dt dq
// Mathematically the same step as planar_kernels.cuh::substep (ABA and M^-1 are the same linear
// operator); tests hold both against the fp64 oracle.
#pragma once
#include "../../include/dartb.h"
#include "planar_kernels.cuh"


// ------------------------------------------------------------------------ compile-time topology facts
template <class T>
struct Coop {
    static constexpr int NB = T::NB, NS = T::NS;
    static constexpr int G = NB <= 8 ? 8 : 16;          // lanes per world
    static constexpr int WPW = 32 / G;                  // worlds per warp
    static constexpr int RPL = (T::NR + G - 1) / G;     // LCP rows per lane
    static constexpr int NC = RPL * G;                  // tableau columns
    static constexpr int OREF = NB > 2 ? 2 : 0;         // body whose origin is the reference point O
    static_assert(NS <= G && NB <= G && NC <= 32, "group too small for this topology");
    __host__ __device__ static constexpr int depth(int i) { int d = 0; for (int k = i; k >= 0; k = T::parent(k)) d++; return d; }
    __host__ __device__ static constexpr int maxdepth() { int m = 0; for (int i = 0; i < NB; i++) m = depth(i) > m ? depth(i) : m; return m; }
    static constexpr int ROUNDS = maxdepth() <= 2 ? 1 : (maxdepth() <= 4 ? 2 : (maxdepth() <= 8 ? 3 : 4));
    __host__ __device__ static constexpr int hop(int i, int k) { int a = i; for (int s = 0; s < k && a >= 0; s++) a = T::parent(a); return a; }
    __host__ __device__ static constexpr unsigned ancmask(int i) { unsigned m = 0; for (int k = i; k >= 0; k = T::parent(k)) m |= 1u << k; return m; }
    __host__ __device__ static constexpr unsigned descmask(int i) { unsigned m = 0; for (int k = 0; k < NB; k++) if (topo_is_ancestor<T>(i, k)) m |= 1u << k; return m; }
    __host__ __device__ static constexpr bool anc(int j, int i) { return topo_is_ancestor<T>(j, i); }   // j == i or j above i
};

// per-lane constants (lane l of a group: body l, dof l, capsule l, actuator of dof l).  A table of
// G of these is built on the HOST (coop_lane_init) and kept in device memory; each lane loads its
// entry once per kernel with a few vector loads instead of a per-field select chain.
template <class T, typename R>
struct alignas(16) CoopLane {
    int l;
    int isb, iss;
    int par, hop2, hop4, hop8;
    unsigned anc, desc;
    int jt, limited;
    R sgn, ax, ay, ux, uy, mass, cx, cy, izz, damp, ksp, rest, qlo, qhi, ox, oy, fnx, fny, qinit, dqinit;
    int sb, sorig;
    unsigned sanc;
    R scx, scy, sdx, sdy, shalf, srad, smu;
    int act, pen_dof;          // task layer: actuator index driving this dof (-1: none); limit-penalty dof flag
    R ascale, alo, ahi;
};

template <class T, typename R>
inline void coop_lane_init(const PModel<R>& M, const PTask<R>* K, int l, CoopLane<T, R>& c) {   // HOST: builds table entry l
    using C = Coop<T>;
    c = CoopLane<T, R>();
    c.l = l; c.isb = l < C::NB; c.iss = l < C::NS;
    c.par = -1; c.hop2 = -1; c.hop4 = -1; c.hop8 = -1; c.sorig = -1; c.act = -1;
    if (l < C::NB) {
        const int i = l;
        c.par = T::parent(i); c.hop2 = C::hop(i, 2); c.hop4 = C::hop(i, 4); c.hop8 = C::hop(i, 8);
        c.anc = C::ancmask(i); c.desc = C::descmask(i); c.jt = T::jtype(i); c.limited = M.limited[i];
        c.sgn = M.sgn[i]; c.ax = M.ax[i]; c.ay = M.ay[i]; c.ux = M.ux[i]; c.uy = M.uy[i];
        c.mass = M.mass[i]; c.cx = M.cx[i]; c.cy = M.cy[i]; c.izz = M.izz[i];
        c.damp = M.damping[i]; c.ksp = M.kspring[i]; c.rest = M.rest[i]; c.qlo = M.qlo[i]; c.qhi = M.qhi[i];
        c.ox = M.ox[i]; c.oy = M.oy[i]; c.fnx = M.fnx[i]; c.fny = M.fny[i]; c.qinit = M.qinit[i]; c.dqinit = M.dqinit[i];
        if (K) { c.act = K->dof_act[i]; c.ascale = K->dof_scale[i]; c.alo = K->dof_lo[i]; c.ahi = K->dof_hi[i]; c.pen_dof = K->limit_pen_dof == i; }
    }
    if (l < C::NS) {
        const int s = l;
        c.sb = T::sbody(s); c.sanc = C::ancmask(T::sbody(s)); c.sorig = M.sorig[s];
        c.scx = M.scx[s]; c.scy = M.scy[s]; c.sdx = M.sdx[s]; c.sdy = M.sdy[s];
        c.shalf = M.shalf[s]; c.srad = M.srad[s]; c.smu = M.smu[s];
    }
}
template <class T, typename R>
inline void coop_build_table(const PModel<R>& M, const PTask<R>* K, CoopLane<T, R>* out /*[Coop<T>::G]*/) {
    for (int l = 0; l < Coop<T>::G; l++) coop_lane_init<T, R>(M, K, l, out[l]);
}

#include "warp_group.cuh"   // group collectives + GroupLcp (shared with the quad form of the per-thread kernels)
template <class T, typename R, int NCx>
using CoopLcp = GroupLcp<Coop<T>::G, R, NCx>;


// inclusive sum over the ancestors of each body (pointer jumping: parent, 2nd, 4th, 8th ancestor)
template <class T, int K, typename R>
DEVI void anc_prefix(const CoopLane<T, R>& c, R (&x)[K]) {
    using C = Coop<T>;
    R t[K];
#pragma unroll
    for (int k = 0; k < K; k++) t[k] = gshfl<C::G>(x[k], c.par < 0 ? c.l : c.par);
#pragma unroll
    for (int k = 0; k < K; k++) if (c.par >= 0) x[k] += t[k];
    if constexpr (C::ROUNDS >= 2) {
#pragma unroll
        for (int k = 0; k < K; k++) t[k] = gshfl<C::G>(x[k], c.hop2 < 0 ? c.l : c.hop2);
#pragma unroll
        for (int k = 0; k < K; k++) if (c.hop2 >= 0) x[k] += t[k];
    }
    if constexpr (C::ROUNDS >= 3) {
#pragma unroll
        for (int k = 0; k < K; k++) t[k] = gshfl<C::G>(x[k], c.hop4 < 0 ? c.l : c.hop4);
#pragma unroll
        for (int k = 0; k < K; k++) if (c.hop4 >= 0) x[k] += t[k];
    }
    if constexpr (C::ROUNDS >= 4) {
#pragma unroll
        for (int k = 0; k < K; k++) t[k] = gshfl<C::G>(x[k], c.hop8 < 0 ? c.l : c.hop8);
#pragma unroll
        for (int k = 0; k < K; k++) if (c.hop8 >= 0) x[k] += t[k];
    }
}

// 1/sqrt(x) for the mass-matrix core.  FAST (the fp32 engine's fp64 core): MUFU.RSQ seed (2^-22) + one
// Newton step (rel. error ~1e-13, far below the fp32 kinematics it serves) instead of the
// ~100-instruction IEEE sqrt + divide sequence; the fp64 validation engine keeps the exact one.
template <bool FAST, typename RM>
DEVI RM mass_rsqrt(RM x) {
    if constexpr (FAST && std::is_same<RM, double>::value) {
        const double h = 0.5 * x;
        double y = (double)Num<float>::rsqrt_((float)x);
        y = y * (1.5 - h * y * y);
        return y;
    } else return Num<RM>::rsqrt_(x);
}

// ------------------------------------------------------------------------ sparse L^T L (Featherstone LTL)
// A (lower triangle, entries between related bodies only) -> L with M = L^T L; Li[k] = 1 / L_kk.
template <class T, typename R, bool FAST = false>
DEVI void ltl_factor(R (&A)[T::NB][T::NB], R (&Li)[T::NB]) {
    using C = Coop<T>;
    static_rfor<T::NB>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        Li[k] = mass_rsqrt<FAST, R>(A[k][k]);
        static_for<0, k>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if constexpr (C::anc(i, k)) A[k][i] *= Li[k];
        });
        static_for<0, k>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if constexpr (C::anc(i, k)) {
                static_for<0, i + 1>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    if constexpr (C::anc(j, k)) A[i][j] -= A[k][i] * A[k][j];
                });
            }
        });
    });
}
// Two independent factorisations statement by statement, so the two dependent chains (rsqrt -> scale ->
// update per column) interleave in the instruction stream of a warp that has nothing else to issue.
template <class T, typename R, bool FAST = false>
DEVI void ltl_factor2(R (&A)[T::NB][T::NB], R (&La)[T::NB], R (&B)[T::NB][T::NB], R (&Lb)[T::NB]) {
    using C = Coop<T>;
    static_rfor<T::NB>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        La[k] = mass_rsqrt<FAST, R>(A[k][k]);
        Lb[k] = mass_rsqrt<FAST, R>(B[k][k]);
        static_for<0, k>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if constexpr (C::anc(i, k)) { A[k][i] *= La[k]; B[k][i] *= Lb[k]; }
        });
        static_for<0, k>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if constexpr (C::anc(i, k)) {
                static_for<0, i + 1>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    if constexpr (C::anc(j, k)) { A[i][j] -= A[k][i] * A[k][j]; B[i][j] -= B[k][i] * B[k][j]; }
                });
            }
        });
    });
}
// x <- L^-T x
template <class T, typename R>
DEVI void ltl_solve_t(const R (&L)[T::NB][T::NB], const R (&Li)[T::NB], R (&x)[T::NB]) {
    using C = Coop<T>;
    static_rfor<T::NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        x[i] *= Li[i];
        static_for<0, i>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            if constexpr (C::anc(j, i)) x[j] -= L[i][j] * x[i];
        });
    });
}
// x <- L^-1 x
template <class T, typename R>
DEVI void ltl_solve(const R (&L)[T::NB][T::NB], const R (&Li)[T::NB], R (&x)[T::NB]) {
    using C = Coop<T>;
    static_for<0, T::NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        static_for<0, i>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            if constexpr (C::anc(j, i)) x[i] -= L[i][j] * x[j];
        });
        x[i] *= Li[i];
    });
}

// ------------------------------------------------------------------------ shared scratch (per group)
// constraint-row records: written by the lane that found the constraint, read by the lane that owns
// the row (the compaction of "which capsules touch / which limits are active" into rows 0..n-1)
template <class T, typename R>
struct CoopRows {
    static constexpr int NC = Coop<T>::NC;
    int kind[NC];        // 1 contact normal, 2 contact tangent, 3 limit (lower), 4 limit (upper)
    int src[NC];         // capsule index / dof index
    unsigned anc[NC];    // ancestors-or-self of the contact body
    R w0[NC], w1[NC], w2[NC];   // unit wrench about O: [moment; direction]
    R bias[NC], mu[NC];
};
template <typename R> struct RowIO {
    template <class S> static DEVI void put(S* s, int r, R w0, R w1, R w2, R bias, R mu) { s->w0[r] = w0; s->w1[r] = w1; s->w2[r] = w2; s->bias[r] = bias; s->mu[r] = mu; }
    template <class S> static DEVI void get(const S* s, int r, R& w0, R& w1, R& w2, R& bias, R& mu) { w0 = s->w0[r]; w1 = s->w1[r]; w2 = s->w2[r]; bias = s->bias[r]; mu = s->mu[r]; }
};


// fixed-sweep projected Gauss-Seidel, rows on lanes (lcp_pgs of planar_kernels.cuh; DART PGSLCPSolver shape)
template <class T, typename R, int NCx>
DEVI void coop_pgs(int l, int n, int nmax, const R (&A)[(NCx + Coop<T>::G - 1) / Coop<T>::G][NCx],
                   const R (&b)[(NCx + Coop<T>::G - 1) / Coop<T>::G], const R (&lo)[(NCx + Coop<T>::G - 1) / Coop<T>::G],
                   const R (&hi)[(NCx + Coop<T>::G - 1) / Coop<T>::G], const int (&fi)[(NCx + Coop<T>::G - 1) / Coop<T>::G], int iters,
                   R (&x)[(NCx + Coop<T>::G - 1) / Coop<T>::G]) {
    using C = Coop<T>;
    constexpr int G = C::G, NC = NCx, RPL = (NCx + G - 1) / G;
    R xa[NC];
#pragma unroll
    for (int cc = 0; cc < NC; cc++) xa[cc] = 0;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NC; i++) {
            // the owner of row i updates x_i from the freshest values, then everyone learns it
            const int h = i / G;
            R xi = 0;
            if (l == i % G && i < n) {
                R aii = 0, s = 0;
#pragma unroll
                for (int hh = 0; hh < RPL; hh++) if (hh == h) {
                    s = b[hh];
#pragma unroll
                    for (int cc = 0; cc < NC; cc++) { if (cc == i) aii = A[hh][cc]; else if (cc < n) s -= A[hh][cc] * xa[cc]; }
                    if (aii < (R)1e-9) xi = 0;
                    else {
                        s /= aii;
                        R lw = lo[hh], hg = hi[hh];
                        if (fi[hh] >= 0) {
                            R xn = 0;
#pragma unroll
                            for (int cc = 0; cc < NC; cc++) { if (cc == fi[hh]) xn = xa[cc]; }
                            hg = hi[hh] * xn; lw = -hg;
                        }
                        if (s > hg) s = hg;
                        if (s < lw) s = lw;
                        xi = s;
                    }
                }
            }
            xa[i] = gshfl<G>(xi, i % G);
        }
    }
#pragma unroll
    for (int h = 0; h < RPL; h++) {
        x[h] = 0;
#pragma unroll
        for (int cc = 0; cc < NC; cc++) { if (cc == l + h * G) x[h] = xa[cc]; }
    }
}

// ------------------------------------------------------------------------ kinematics of the group
// positions only (task layer: height of a body COM); returns this lane's (cs, sn, px, py)
template <class T, typename R>
DEVI void coop_fk_positions(const CoopLane<T, R>& c, R q, R& cs, R& sn, R& px, R& py) {
    using C = Coop<T>;
    const bool rev = c.isb && c.jt == PM_REV, pri = c.isb && c.jt == PM_PRI;
    R a1[1] = {rev ? c.sgn * q : (R)0};
    anc_prefix<T, 1, R>(c, a1);
    Num<R>::sincos_(a1[0], &sn, &cs);
    const int ps = c.par < 0 ? c.l : c.par;
    R cp = gshfl<C::G>(cs, ps), sp = gshfl<C::G>(sn, ps);
    if (c.par < 0) { cp = 1; sp = 0; }
    const R arx = cp * c.ax - sp * c.ay, ary = sp * c.ax + cp * c.ay;
    const R uwx = pri ? cp * c.ux - sp * c.uy : (R)0, uwy = pri ? sp * c.ux + cp * c.uy : (R)0;
    R p2[2] = {c.isb ? arx + uwx * q : (R)0, c.isb ? ary + uwy * q : (R)0};
    anc_prefix<T, 2, R>(c, p2);
    px = p2[0]; py = p2[1];
}

// ------------------------------------------------------------------------ K5-K7 for one column class
template <typename R>
struct CoopContact {    // what lane l found for capsule l / dof l
    bool hasc, fric;
    R Px, Py, nx, ny, depth, mu;
    int lact;
    unsigned cm, fm, lm;   // group masks: capsules in contact, with friction, dofs at a limit
    R Ox, Oy;
};

template <class T, typename R, typename RM, bool FLUID, int NCx>
DEVI void coop_constraints(const PModel<R>& M, const CoopLane<T, R>& c, int gbase, R& dq, const CoopContact<R>& ct, int n, int nmax,
                           const RM (&Mf)[T::NB][T::NB], const RM (&Li)[T::NB], const R (&Pgx)[T::NB], const R (&Pgy)[T::NB],
                           const R (&Ugx)[T::NB], const R (&Ugy)[T::NB], int lcp_mode, int pgs_iters, const ContactSink<R>* sink,
                           int world, uint32_t& hint, CoopRows<T, R>* rows) {
    using C = Coop<T>;
    constexpr int NB = C::NB, G = C::G, NC = NCx, RPL = (NCx + G - 1) / G;
    const R dt = M.dt;
    const int l = c.l;
    const R INF = Num<R>::inf();
    const bool hasc = ct.hasc, fric = ct.fric;
    const R cPx = ct.Px, cPy = ct.Py, cnx = ct.nx, cny = ct.ny, cdepth = ct.depth, cmu = ct.mu, Ox = ct.Ox, Oy = ct.Oy;
    const int lact = ct.lact;
    const unsigned cm = ct.cm, fm = ct.fm, lm = ct.lm;
    const unsigned below = (1u << l) - 1u;
    const int ncr = __popc(cm) + __popc(fm);
    int my_rown = -1;
    {
        // ---------------- K5: compact the constraints into rows 0..n-1 through shared memory
        if (hasc) {
            const int r0 = __popc(cm & below) + __popc(fm & below);
            my_rown = r0;
            const R qx = cPx - Ox, qy = cPy - Oy;   // contact point relative to O
            R bounce = cdepth;
            if (bounce < 0) bounce = 0;
            else { bounce *= ((R)1 / dt) * (R)DK_CONTACT_ERP; if (bounce > (R)DK_CONTACT_MAX_ERV) bounce = (R)DK_CONTACT_MAX_ERV; }
            rows->kind[r0] = 1; rows->src[r0] = l; rows->anc[r0] = c.sanc;
            RowIO<R>::put(rows, r0, qx * cny - qy * cnx, cnx, cny, bounce, (R)0);
            if (fric) {
                const R tx = -cny, ty = cnx;   // DART tangent t1 = z x n (in-plane); t2 is out of plane (inert)
                rows->kind[r0 + 1] = 2; rows->src[r0 + 1] = l; rows->anc[r0 + 1] = c.sanc;
                RowIO<R>::put(rows, r0 + 1, qx * ty - qy * tx, tx, ty, (R)0, cmu);
            }
        }
        if (lact != 0) {
            const int r = ncr + __popc(lm & below);
            rows->kind[r] = lact < 0 ? 3 : 4; rows->src[r] = l; rows->anc[r] = 0;
            RowIO<R>::put(rows, r, (R)0, (R)0, (R)0, (R)0, (R)0);
        }
        __syncwarp();
        R dqg[NB];
        static_for<0, NB>([&](auto ic) { constexpr int i = decltype(ic)::value; dqg[i] = gshfl<G>(dq, i); });
        // Rows 13+ only occur with six or more capsules on the ground at once (a fallen walker / cheetah): A is
        // then rank-deficient up to the CFM and the fp32 tableau loses ~1e-4, so the two large column classes
        // (never reached by the in-scope tasks' normal operation) run the LCP in the mass-matrix precision.
        using RL = typename std::conditional<(NCx >= 16), RM, R>::type;
        const RL INFL = Num<RL>::inf();
        R Jf[RPL][NB];      // J_r (this lane's rows)
        RM MJ[RPL][NB];     // M^-1 J_r^T
        RL bb[RPL], lo[RPL], hi[RPL];
        int fi[RPL], kind[RPL], rsrc[RPL];
        unsigned hin[RPL];
#pragma unroll
        for (int h = 0; h < RPL; h++) {
            const int r = l + h * G;
            kind[h] = 0; rsrc[h] = 0; fi[h] = -1; bb[h] = 0; lo[h] = 0; hi[h] = 0; hin[h] = 3u;
            static_for<0, NB>([&](auto jc) { constexpr int j = decltype(jc)::value; Jf[h][j] = 0; });
            if (r < n) {
                kind[h] = rows->kind[r]; rsrc[h] = rows->src[r];
                const unsigned anc = rows->anc[r];
                R w0, w1, w2, bias, mu;
                RowIO<R>::get(rows, r, w0, w1, w2, bias, mu);
                R vn = 0;
                static_for<0, NB>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    R Jv;
                    if (kind[h] <= 2) {
                        if constexpr (T::jtype(j) == PM_REV) Jv = M.sgn[j] * (w0 + Pgy[j] * w1 - Pgx[j] * w2);
                        else Jv = Ugx[j] * w1 + Ugy[j] * w2;
                        if (!((anc >> j) & 1u)) Jv = 0;
                    } else Jv = (j == rsrc[h]) ? (R)1 : (R)0;
                    Jf[h][j] = Jv;
                    vn += Jv * dqg[j];
                });
                bb[h] = (RL)(bias - vn);
                if (kind[h] == 1 || kind[h] == 3) { lo[h] = 0; hi[h] = INFL; }
                else if (kind[h] == 4) { lo[h] = -INFL; hi[h] = 0; }
                else { lo[h] = -(RL)mu; hi[h] = (RL)mu; fi[h] = r - 1; hin[h] = (hint >> (2 * rsrc[h])) & 3u; }
            }
            static_for<0, NB>([&](auto jc) { constexpr int j = decltype(jc)::value; MJ[h][j] = (RM)Jf[h][j]; });
            ltl_solve_t<T, RM>(Mf, Li, MJ[h]);
            ltl_solve<T, RM>(Mf, Li, MJ[h]);      // M^-1 J_r^T = L^-1 L^-T J_r^T  (DART: applyUnitImpulse + getVelocityChange)
        }
        // A[r][s] = J_s . (M^-1 J_r^T) (+ CFM on the diagonal): the fp32 rows J_s travel by shuffles, the products
        // accumulate in RM
        RL A[RPL][NC];
#pragma unroll
        for (int s = 0; s < NC; s++) {
            R Js[NB];
            static_for<0, NB>([&](auto jc) { constexpr int j = decltype(jc)::value; Js[j] = gshfl<G>(Jf[s / G][j], s % G); });
#pragma unroll
            for (int h = 0; h < RPL; h++) {
                RM v = 0;
                static_for<0, NB>([&](auto jc) { constexpr int j = decltype(jc)::value; v += MJ[h][j] * (RM)Js[j]; });
                if (s == l + h * G) v *= (RM)1 + (kind[h] <= 2 ? (RM)DK_CONTACT_CFM : (RM)DK_LIMIT_CFM);
                A[h][s] = (RL)v;
            }
        }
        // ---------------- K6
        RL x[RPL];
        unsigned st[RPL];
        if (lcp_mode == 1) {
            coop_pgs<T, RL, NCx>(l, n, nmax, A, bb, lo, hi, fi, pgs_iters, x);
#pragma unroll
            for (int h = 0; h < RPL; h++) st[h] = 3u;
        } else {
            const bool ok = CoopLcp<T, RL, NCx>::solve(l, gbase, n, nmax, A, bb, lo, hi, fi, hin, x, st);
            // not converged (never observed): many PGS sweeps rather than leaving the rows unsolved
            const bool anybad = __any_sync(COOP_FULL, !ok);
            if (anybad) {
                RL x2[RPL], lo2[RPL], hi2[RPL];
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    // PGS bounds of a friction row are mu * x_normal: restore mu (solve() overwrote hi with |mu x_n|)
                    R mu = 0;
                    if (kind[h] == 2) { R w0, w1, w2, bias; RowIO<R>::get(rows, l + h * G, w0, w1, w2, bias, mu); }
                    lo2[h] = kind[h] == 2 ? -(RL)mu : lo[h]; hi2[h] = kind[h] == 2 ? (RL)mu : hi[h];
                }
                coop_pgs<T, RL, NCx>(l, n, nmax, A, bb, lo2, hi2, fi, 200, x2);
#pragma unroll
                for (int h = 0; h < RPL; h++) if (!ok) { x[h] = x2[h]; st[h] = 3u; }
            }
        }
        // ---------------- K7: dq += sum_r (M^-1 J_r^T) x_r: each row's contribution rounded to R, summed over the group
        static_for<0, NB>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            R v = (l < n) ? (R)(MJ[0][i] * (RM)x[0]) : (R)0;
#pragma unroll
            for (int h = 1; h < RPL; h++) if (l + h * G < n) v = coop_add_rn(v, (R)(MJ[h][i] * (RM)x[h]));
            const R d = group_sum<G>(v);
            if (l == i) dq += d;
        });
        // stick/slide sets of the friction rows for the next step
        {
            unsigned clr = 0, set = 0;
#pragma unroll
            for (int h = 0; h < RPL; h++)
                if (kind[h] == 2) { clr |= 3u << (2 * rsrc[h]); set |= (st[h] & 3u) << (2 * rsrc[h]); }
            clr = group_or<G>(clr); set = group_or<G>(set);
            hint = (0xffffffffu & ~clr) | set;
        }
        // contact read-back (last DART step of an env step)
        if (sink) {
            // impulses of this capsule's rows
            const int rn = my_rown < 0 ? 0 : my_rown, rt = rn + 1 < NC ? rn + 1 : rn;
            R xn = gshfl<G>(x[0], rn % G), xt = gshfl<G>(x[0], rt % G);
            if constexpr (RPL > 1) {
                const R xn1 = gshfl<G>(x[1], rn % G), xt1 = gshfl<G>(x[1], rt % G);
                if (rn / G == 1) xn = xn1;
                if (rt / G == 1) xt = xt1;
            }
            if (hasc && sink->data) {
                const int ci = __popc(cm & below);
                if (ci < sink->maxc) {
                    if (!fric) xt = 0;
                    const R inv_dt = (R)1 / dt;
                    const R fx = (cnx * xn - cny * xt) * inv_dt, fy = (cny * xn + cnx * xt) * inv_dt;
                    float* o = sink->data + ((size_t)world * sink->maxc + ci) * 10;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        o[k] = (float)(M.e1[k] * cPx + M.e2[k] * cPy + M.en[k] * M.hz);
                        o[3 + k] = (float)(M.e1[k] * cnx + M.e2[k] * cny);
                        o[7 + k] = (float)(M.e1[k] * fx + M.e2[k] * fy);
                    }
                    o[6] = (float)cdepth;
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------ one DART time step, G lanes per world
// q, dq, tau: this lane's dof.  hint: 2 bits per capsule (stick/slide set of its friction row at the
// previous step; all ones = unknown), identical in every lane of the group.
template <class T, typename R, bool FLUID>
DEVI void coop_substep(const PModel<R>& M, const CoopLane<T, R>& c, int gbase, R& q, R& dq, const R tau, R fluid_offset,
                       R fluid_coef, int lcp_mode, int pgs_iters, const ContactSink<R>* sink, bool wactive, int world,
                       uint32_t& hint, CoopRows<T, R>* rows) {
    using C = Coop<T>;
    constexpr int NB = C::NB, NS = C::NS, G = C::G, RPL = C::RPL, NC = C::NC;
#ifdef DARTB_COOP_MASS_NATIVE
    using RM = R;
#else
    using RM = double;
#endif
    constexpr bool FASTM = std::is_same<R, float>::value;
    const R dt = M.dt;
    const int l = c.l;
    const bool rev = c.isb && c.jt == PM_REV, pri = c.isb && c.jt == PM_PRI;
    const R INF = Num<R>::inf();

    // ---------------- K1: kinematics by ancestor prefix sums
    R a2[2] = {rev ? c.sgn * q : (R)0, rev ? c.sgn * dq : (R)0};
    anc_prefix<T, 2, R>(c, a2);
    const R th = a2[0], wz = a2[1];
    R sn, cs;
    Num<R>::sincos_(th, &sn, &cs);
    const int ps = c.par < 0 ? l : c.par;
    R cp = gshfl<G>(cs, ps), sp = gshfl<G>(sn, ps), wp = gshfl<G>(wz, ps);
    if (c.par < 0) { cp = 1; sp = 0; wp = 0; }
    const R arx = cp * c.ax - sp * c.ay, ary = sp * c.ax + cp * c.ay;
    const R uwx = pri ? cp * c.ux - sp * c.uy : (R)0, uwy = pri ? sp * c.ux + cp * c.uy : (R)0;
    const R rx = arx + uwx * q, ry = ary + uwy * q;    // parent origin -> this origin, world axes
    R k6[6] = {rx, ry,
               -wp * ry + uwx * dq, wp * rx + uwy * dq,                                   // origin velocity increment
               -wp * wp * rx - (R)2 * wp * uwy * dq, -wp * wp * ry + (R)2 * wp * uwx * dq};  // origin acceleration, ddq = 0
    if (!c.isb) { k6[0] = 0; k6[1] = 0; k6[2] = 0; k6[3] = 0; k6[4] = 0; k6[5] = 0; }
    anc_prefix<T, 6, R>(c, k6);
    const R px = k6[0], py = k6[1], vx = k6[2], vy = k6[3], a0x = k6[4], a0y = k6[5];
    // everything dynamic is taken relative to O = origin of the root rotational body (precision)
    const R Ox = gshfl<G>(px, C::OREF), Oy = gshfl<G>(py, C::OREF);
    const R Px = px - Ox, Py = py - Oy;
    const R dcx = cs * c.cx - sn * c.cy, dcy = sn * c.cx + cs * c.cy;   // origin -> COM
    const R Dx = Px + dcx, Dy = Py + dcy;

    // ---------------- K2: per-body wrench for ddq = 0 (Newton-Euler), external forces
    const R m = c.mass;
    const R acx = a0x - wz * wz * dcx, acy = a0y - wz * wz * dcy;
    R Fx = m * (acx - M.gx), Fy = m * (acy - M.gy);
    R nF = 0;   // extra moment about O that does not come from (Fx, Fy) acting at the COM
    if constexpr (FLUID) {
        // snake_7link.py:35-47: bn.com_spatial_velocity(), norm_dir = R*ez, add_ext_force at the body origin
        const R nx = cs * c.fnx - sn * c.fny, ny = sn * c.fnx + cs * c.fny;
        const R vwx = vx - wz * dcy, vwy = vy + wz * dcx;
        const R vcx = cs * vwx + sn * vwy, vcy = cs * vwy - sn * vwx;   // body coordinates (see planar_kernels.cuh::substep)
        const R crx = -wz * ny, cry = wz * nx;
        const R dp = (vcx + crx * fluid_offset) * nx + (vcy + cry * fluid_offset) * ny;
        const R dn = (vcx - crx * fluid_offset) * nx + (vcy - cry * fluid_offset) * ny;
        R ffx = 0, ffy = 0;
        if (dp > 0) { ffx = -fluid_coef * dp * nx; ffy = -fluid_coef * dp * ny; }
        if (dn < 0) { ffx = -fluid_coef * dn * nx; ffy = -fluid_coef * dn * ny; }
        const R oxw = cs * c.ox - sn * c.oy, oyw = sn * c.ox + cs * c.oy;
        // the applied force acts at (P + o); fold it into an equivalent (force at COM, extra moment)
        Fx -= ffx; Fy -= ffy;
        nF = -((Px + oxw - Dx) * ffy - (Py + oyw - Dy) * ffx);
    }
    // ---------------- K2/K2': subtree sums about each receiving joint's own origin
    // The joint-space inertia of these skeletons is ill-conditioned (hopper: cond ~ 6e3 after diagonal
    // scaling: torso and hip rotations move the leg almost alike), so entries rounded to fp32 cost ~1e-4
    // in ddq where the articulated-body recursion loses ~1e-6.  The composite inertias, M, its factor and
    // the solves are therefore carried in RM = fp64 (exact for the fp32 kinematics they are built from,
    // i.e. the mass matrix of a 1e-7-perturbed pose); everything else stays in R.
    RM Jc = 0, hxc = 0, hyc = 0, mc = 0;
    R nc_ = 0, fxc = 0, fyc = 0;   // the bias wrench needs no more than R (measured: no loss on the goldens)
    R Pgx[NB], Pgy[NB];
    {
        const R g4 = c.isb ? Fx : (R)0, g5 = c.isb ? Fy : (R)0;
        static_for<0, NB>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            const R Dx_j = gshfl<G>(Dx, j), Dy_j = gshfl<G>(Dy, j), Fx_j = gshfl<G>(g4, j), Fy_j = gshfl<G>(g5, j);
            const R nF_j = FLUID ? gshfl<G>(nF, j) : (R)0;
            Pgx[j] = gshfl<G>(Px, j); Pgy[j] = gshfl<G>(Py, j);
            const bool in = (c.desc >> j) & 1u;             // body j hangs below this joint
            // its COM seen from this joint's origin: the difference is formed in RM, where it is EXACT, so all
            // rows of M describe one and the same (1e-7-perturbed) geometry
            const RM ex = (RM)Dx_j - (RM)Px, ey = (RM)Dy_j - (RM)Py;
            const R exf = (R)ex, eyf = (R)ey;
            const RM mj = in ? (RM)M.mass[j] : (RM)0, izj = in ? (RM)M.izz[j] : (RM)0;
            Jc += izj + mj * (ex * ex + ey * ey); hxc -= mj * ey; hyc += mj * ex; mc += mj;
            const R fxm = in ? Fx_j : (R)0, fym = in ? Fy_j : (R)0;
            nc_ += exf * fym - eyf * fxm + (in ? nF_j : (R)0); fxc += fxm; fyc += fym;
        });
    }
    // prismatic axes of the ancestors (static joint types: only prismatic bodies are fetched)
    R Ugx[NB], Ugy[NB];
    static_for<0, NB>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (T::jtype(j) == PM_PRI) { Ugx[j] = gshfl<G>(uwx, j); Ugy[j] = gshfl<G>(uwy, j); }
        else { Ugx[j] = 0; Ugy[j] = 0; }
    });
    // bias force and F_i = Ic_i S_i, both about this joint's origin
    RM F0, F1, F2;
    R cb;
    if (rev) { cb = c.sgn * nc_; F0 = (RM)c.sgn * Jc; F1 = (RM)c.sgn * hxc; F2 = (RM)c.sgn * hyc; }
    else { cb = uwx * fxc + uwy * fyc; F0 = hxc * (RM)uwx + hyc * (RM)uwy; F1 = mc * (RM)uwx; F2 = mc * (RM)uwy; }
    // own row of M: M_ij = S_j(at this origin) . F_i for the ancestors j
    RM Mrow[NB];
    static_for<0, NB>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (T::jtype(j) == PM_REV) Mrow[j] = (RM)M.sgn[j] * (F0 + ((RM)Pgy[j] - (RM)Py) * F1 - ((RM)Pgx[j] - (RM)Px) * F2);
        else Mrow[j] = (RM)Ugx[j] * F1 + (RM)Ugy[j] * F2;
    });
    // every lane assembles the full (tree-sparse) M from the rows
    RM Mf[NB][NB];
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        static_for<0, i + 1>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            if constexpr (C::anc(j, i)) Mf[i][j] = gshfl<G>(Mrow[j], i);
        });
    });
    // ---------------- K3: (M + dt D + dt^2 K) ddq = tau - c - K (q - rest + dt dq) - D dq
    RM Lp[NB];   // 1 / diag of the factor of the plain M (filled here when NS > 0, else on demand below)
    {
        const R rhs = c.isb ? tau - cb - c.ksp * (q - c.rest + dt * dq) - c.damp * dq : (R)0;
        RM xg[NB], Mt[NB][NB], Li[NB];
        static_for<0, NB>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            xg[i] = (RM)gshfl<G>(rhs, i);
            static_for<0, i + 1>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                if constexpr (C::anc(j, i)) Mt[i][j] = Mf[i][j];
            });
            Mt[i][i] += (RM)dt * (RM)M.damping[i] + (RM)dt * (RM)dt * (RM)M.kspring[i];
        });
        // skeletons with capsules nearly always have constraint rows somewhere in the warp: factor the plain M
        // (needed for the impulse tests) together with the implicit one instead of after the collision phase
        if constexpr (NS > 0) { ltl_factor2<T, RM, FASTM>(Mt, Li, Mf, Lp); }
        else ltl_factor<T, RM, FASTM>(Mt, Li);
        ltl_solve_t<T, RM>(Mt, Li, xg);
        ltl_solve<T, RM>(Mt, Li, xg);
        RM ddq = 0;
        static_for<0, NB>([&](auto ic) { constexpr int i = decltype(ic)::value; if (l == i) ddq = xg[i]; });
        dq += dt * (R)ddq;
    }

    // ---------------- K4: capsule l against the ground box
    bool hasc = false, fric = false;
    R cPx = 0, cPy = 0, cnx = 0, cny = 0, cdepth = 0, cmu = 0;
    if constexpr (NS > 0) {
        const R bpx = gshfl<G>(px, c.sb), bpy = gshfl<G>(py, c.sb), bcs = gshfl<G>(cs, c.sb), bsn = gshfl<G>(sn, c.sb);
        if (M.has_ground && c.iss && wactive) {
            const R ccx = bpx + bcs * c.scx - bsn * c.scy, ccy = bpy + bsn * c.scx + bcs * c.scy;
            const R adx = bcs * c.sdx - bsn * c.sdy, ady = bsn * c.sdx + bcs * c.sdy;
            const R hl = c.shalf, rad = c.srad;
            const R ex_ = hl * Num<R>::abs_(adx) + rad + (R)1e-5, ey_ = hl * Num<R>::abs_(ady) + rad + (R)1e-5;
            const bool near_ = Num<R>::abs_(ccx - M.gcx) <= M.ghx + ex_ && Num<R>::abs_(ccy - M.gcy) <= M.ghy + ey_;
            R lx = 0, ly = 0, ddx = 0, ddy = 0, d = INF;
            if (near_) {
                closest_segment_box2<R>(ccx + hl * adx, ccy + hl * ady, ccx - hl * adx, ccy - hl * ady, M.gcx, M.gcy, M.ghx,
                                        M.ghy, lx, ly, ddx, ddy);
                d = Num<R>::sqrt_(ddx * ddx + ddy * ddy);
            }
            if (!(d > rad)) {
                hasc = true;
                if (!(d < Num<R>::mindist())) {
                    const R id = Num<R>::rcp_(d);
                    cnx = ddx * id; cny = ddy * id;
                    cdepth = rad - d;
                    const R k = (R)0.5 * (-rad - d);
                    cPx = lx + cnx * k; cPy = ly + cny * k;
                } else {
                    cnx = M.gupx; cny = M.gupy;
                    cdepth = rad + (M.ghup - ((lx - M.gcx) * cnx + (ly - M.gcy) * cny));
                    cPx = lx; cPy = ly;
                }
                cmu = c.smu;
                fric = cmu > (R)DK_FRICTION_THRESHOLD;
            }
        }
    }
    // joint limits: q BEFORE this step's integration
    int lact = 0;
    if (c.isb && c.limited && wactive) { if (q - c.qlo <= 0) lact = -1; else if (q - c.qhi >= 0) lact = 1; }
    const unsigned cm = group_ballot<G>(hasc, gbase), fm = group_ballot<G>(hasc && fric, gbase), lm = group_ballot<G>(lact != 0, gbase);
    const unsigned below = (1u << l) - 1u;
    const int ncont = __popc(cm), ncr = ncont + __popc(fm);
    const int n = ncr + __popc(lm);
    const int nmax = coop_warp_max(n);
    if (nmax > 0) {
        // plain M = L^T L for the impulse tests (DART uses the non-implicit articulated inertia there)
        using RC = RM;
        if constexpr (NS == 0) ltl_factor<T, RC, FASTM>(Mf, Lp);
        CoopContact<R> ct;
        ct.hasc = hasc; ct.fric = fric; ct.Px = cPx; ct.Py = cPy; ct.nx = cnx; ct.ny = cny; ct.depth = cdepth; ct.mu = cmu;
        ct.lact = lact; ct.cm = cm; ct.fm = fm; ct.lm = lm; ct.Ox = Ox; ct.Oy = Oy;
        // the size class is chosen per warp, so a warp runs one code path
        if (nmax <= 4 && C::NC >= 4) coop_constraints<T, R, RC, FLUID, 4>(M, c, gbase, dq, ct, n, nmax, Mf, Lp, Pgx, Pgy, Ugx, Ugy, lcp_mode, pgs_iters, sink, world, hint, rows);
        else if (nmax <= 8 && C::NC >= 8) coop_constraints<T, R, RC, FLUID, (C::NC >= 8 ? 8 : C::NC)>(M, c, gbase, dq, ct, n, nmax, Mf, Lp, Pgx, Pgy, Ugx, Ugy, lcp_mode, pgs_iters, sink, world, hint, rows);
        else if (nmax <= 16 && C::NC >= 16) coop_constraints<T, R, RC, FLUID, (C::NC >= 16 ? 16 : C::NC)>(M, c, gbase, dq, ct, n, nmax, Mf, Lp, Pgx, Pgy, Ugx, Ugy, lcp_mode, pgs_iters, sink, world, hint, rows);
        else coop_constraints<T, R, RC, FLUID, C::NC>(M, c, gbase, dq, ct, n, nmax, Mf, Lp, Pgx, Pgy, Ugx, Ugy, lcp_mode, pgs_iters, sink, world, hint, rows);
    } else {
        hint = 0xffffffffu;
    }
    if (sink && wactive) {
        if (sink->count && l == 0) sink->count[world] = ncont;
        if (sink->body) {
            if (hasc) { const int ci = __popc(cm & below); if (ci < sink->maxc) sink->body[(size_t)world * sink->maxc + ci] = c.sorig; }
            if (l < sink->maxc && l >= ncont) sink->body[(size_t)world * sink->maxc + l] = -1;
            if constexpr (G < 16) { if (l + G < sink->maxc && l + G >= ncont) sink->body[(size_t)world * sink->maxc + l + G] = -1; }
        }
    }
    // ---------------- integrate positions
    q += dt * dq;
}

// ------------------------------------------------------------------------ TMA 1-D bulk copies + mbarrier (sm_100a)
// cp.async.bulk (SASS UBLKCP) moves a contiguous, 16-byte aligned run between global and shared memory without
// touching the register file; completion of loads is counted in bytes on an mbarrier (SASS SYNCS), stores are tracked
// by bulk groups.  Used by k_env_step_coop to stage a CTA's tile (lane table, q / dq rows, actions) with ONE elected
// thread instead of ~20 LDGs per thread, and to write its observation tile with one bulk store.
#ifndef DARTB_HOST_EMU
DEVI uint32_t tma_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVI void tma_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DEVI void tma_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(bytes) : "memory");
}
DEVI void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tma_smem_u32(dst)), "l"(src), "r"(bytes), "r"(tma_smem_u32(bar)) : "memory");
}
DEVI void tma_mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(tma_smem_u32(bar)), "r"(phase) : "memory");
}
DEVI void tma_store_1d(void* dst, const void* src, uint32_t bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the async proxy
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(tma_smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem may be reused / the CTA may exit
}
#endif

// ======================================================================== kernels (G lanes per world)
#ifdef DARTB_HOST_EMU
#define COOP_GLOBAL
#define COOP_GRID_CONSTANT
#define COOP_SHARED_BYTES(name) unsigned char* name = (unsigned char*)simt::shared_ptr()
#else
#ifndef DARTB_COOP_MIN_BLOCKS
#define DARTB_COOP_MIN_BLOCKS 1
#endif
#define COOP_GLOBAL __global__ __launch_bounds__(128, DARTB_COOP_MIN_BLOCKS)
#define COOP_GRID_CONSTANT __grid_constant__
#define COOP_SHARED_BYTES(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// shared memory per block: one CoopRows per world of the block, then the obs staging of each warp
template <class T, typename R>
__host__ __device__ constexpr size_t coop_shared_bytes(int warps, int n_obs) {
    return (size_t)warps * Coop<T>::WPW * sizeof(CoopRows<T, R>) + (size_t)warps * Coop<T>::WPW * n_obs * sizeof(float);
}
// + the TMA staging area of k_env_step_coop: [mbarrier | lane table | q rows | dq rows | actions], 16-byte aligned pieces
template <class T, typename R>
__host__ __device__ constexpr size_t coop_stage_bytes(int warps, int n_act) {
    return 16 + sizeof(CoopLane<T, R>) * Coop<T>::G + 2 * (size_t)T::NB * warps * Coop<T>::WPW * sizeof(R) +
           (((size_t)warps * Coop<T>::WPW * n_act * sizeof(float) + 15) / 16) * 16;
}
template <class T, typename R>
__host__ __device__ constexpr size_t coop_stage_offset(int warps, int n_obs) {
    return ((coop_shared_bytes<T, R>(warps, n_obs) + 15) / 16) * 16;
}

// exactly `skel.set_forces(tau); world.step()` (dart_env.py:174-175), no external forces
template <class T, typename R>
COOP_GLOBAL void k_substep_coop(const COOP_GRID_CONSTANT PModel<R> M, const CoopLane<T, R>* tab, int n, R* qs, R* dqs,
                                const R* tau_in /*[n,nd]*/, int lcp_mode, int pgs_iters, const COOP_GRID_CONSTANT ContactSink<R> sink) {
    using C = Coop<T>;
    constexpr int G = C::G, NB = C::NB;
    COOP_SHARED_BYTES(smraw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int l = lane % G, gi = lane / G, gbase = lane - l;
    const int w = (blockIdx.x * nwarps + warp) * C::WPW + gi;
    const bool wactive = w < n;
    CoopRows<T, R>* rows = reinterpret_cast<CoopRows<T, R>*>(smraw) + (warp * C::WPW + gi);
    const bool mine = wactive && l < NB;
    const CoopLane<T, R> c = tab[l];
    R q = mine ? qs[(size_t)l * n + w] : (R)0, dq = mine ? dqs[(size_t)l * n + w] : (R)0;
    if (!mine) q = c.qinit;
    const R tau = (mine && tau_in) ? tau_in[(size_t)w * NB + l] : (R)0;
    uint32_t hint = 0xffffffffu;   // the literal World.step() drop-in is stateless
    coop_substep<T, R, false>(M, c, gbase, q, dq, tau, (R)0, (R)0, lcp_mode, pgs_iters, &sink, wactive, w, hint, rows);
    if (mine) { qs[(size_t)l * n + w] = q; dqs[(size_t)l * n + w] = dq; }
}

// one launch per env.step(): action -> frame_skip DART steps -> obs / reward / done -> masked auto-reset
template <class T, typename R, bool FLUID>
COOP_GLOBAL void k_env_step_coop(const COOP_GRID_CONSTANT PModel<R> M, const COOP_GRID_CONSTANT PTask<R> K,
                                 const COOP_GRID_CONSTANT StepArgs<R> a, const CoopLane<T, R>* tab) {
    using C = Coop<T>;
    constexpr int G = C::G, NB = C::NB, WPW = C::WPW;
    COOP_SHARED_BYTES(smraw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int l = lane % G, gi = lane / G, gbase = lane - l;
    const int wb = (blockIdx.x * nwarps + warp) * WPW;     // first world of this warp
    const int w = wb + gi;
    const bool wactive = w < a.n;
    CoopRows<T, R>* rows = reinterpret_cast<CoopRows<T, R>*>(smraw) + (warp * WPW + gi);
    float* sobs = reinterpret_cast<float*>(smraw + (size_t)nwarps * WPW * sizeof(CoopRows<T, R>)) + (size_t)warp * WPW * K.n_obs;
    // every global load of the step is issued up front and none depends on another (one DRAM round trip)
    const bool mine = wactive && l < NB;
    CoopLane<T, R> c;
    R q, dq, araw;
#ifndef DARTB_HOST_EMU
    if (a.tma) {
        // TMA staging (a.tma is set by the launcher only for full, 16-byte aligned tiles): one elected thread issues
        // 2*NB + 2 bulk copies for the whole CTA and everybody waits on one mbarrier
        const int WPB = nwarps * WPW, wb0 = blockIdx.x * WPB;
        unsigned char* st = smraw + coop_stage_offset<T, R>(nwarps, K.n_obs);
        uint64_t* bar = reinterpret_cast<uint64_t*>(st);
        CoopLane<T, R>* s_tab = reinterpret_cast<CoopLane<T, R>*>(st + 16);
        R* s_q = reinterpret_cast<R*>(st + 16 + sizeof(CoopLane<T, R>) * G);
        R* s_dq = s_q + NB * WPB;
        float* s_act = reinterpret_cast<float*>(s_dq + NB * WPB);
        if (threadIdx.x == 0) tma_mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t tab_b = sizeof(CoopLane<T, R>) * G, row_b = WPB * sizeof(R), act_b = WPB * K.n_act * sizeof(float);
            tma_mbar_expect_tx(bar, tab_b + 2 * NB * row_b + (a.tma > 1 ? act_b : 0u));
            tma_load_1d(s_tab, tab, tab_b, bar);
#pragma unroll
            for (int i = 0; i < NB; i++) {
                tma_load_1d(s_q + i * WPB, a.q + (size_t)i * a.n + wb0, row_b, bar);
                tma_load_1d(s_dq + i * WPB, a.dq + (size_t)i * a.n + wb0, row_b, bar);
            }
            if (a.tma > 1) tma_load_1d(s_act, a.action + (size_t)wb0 * K.n_act, act_b, bar);
        }
        const int wl = warp * WPW + gi;    // world index inside the CTA
        R araw_g = 0;
        if (a.tma <= 1) araw_g = (wactive && l < K.n_act) ? (R)a.action[(size_t)w * K.n_act + l] : (R)0;
        tma_mbar_wait(bar, 0);
        c = s_tab[l];
        q = mine ? s_q[l * WPB + wl] : (R)0; dq = mine ? s_dq[l * WPB + wl] : (R)0;
        araw = a.tma > 1 ? ((wactive && l < K.n_act) ? (R)s_act[wl * K.n_act + l] : (R)0) : araw_g;
    } else
#endif
    {
        c = tab[l];
        q = mine ? a.q[(size_t)l * a.n + w] : (R)0; dq = mine ? a.dq[(size_t)l * a.n + w] : (R)0;
        araw = (wactive && l < K.n_act) ? (R)a.action[(size_t)w * K.n_act + l] : (R)0;
    }
    const int el_in = (wactive && l == 0 && a.max_episode_steps > 0) ? a.elapsed[w] : 0;
    const uint32_t ep_in = (wactive && l == 0) ? a.episode[w] : 0u;
    uint32_t hint = wactive ? (uint32_t)a.hint[w] : 0xffffffffu;
    if (!mine) q = c.qinit;
    // this dof's actuator (hopper.py:24-32: clamp, scale, scatter)
    const int act = c.act, pen_dof = c.pen_dof;
    const R ascale = c.ascale, alo = c.alo, ahi = c.ahi;
    // control cost uses the RAW action, summed in action order
    R a2 = 0, tau = 0;
    static_for<0, G>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const R v = gshfl<G>(araw, j);
        if (j < K.n_act) a2 += v * v;
        if (act == j) { R t = v > ahi ? ahi : v; t = t < alo ? alo : t; tau = t * ascale; }
    });
    if (!mine) tau = 0;
    const R posbefore = gshfl<G>(q, 0);
    for (int f = 0; f < K.frame_skip; f++) {
        const ContactSink<R>* sk = (f == K.frame_skip - 1 && (a.sink.count || a.sink.body || a.sink.data)) ? &a.sink : nullptr;
        coop_substep<T, R, FLUID>(M, c, gbase, q, dq, tau, K.fluid_offset, K.fluid_coef, a.lcp_mode, a.pgs_iters, sk, wactive, w, hint, rows);
    }
    // reward / done (hopper.py:36-65, walker2d.py:22-65, half_cheetah.py:40-77, snake_7link.py:68-87)
    const R q0 = gshfl<G>(q, 0), ang = gshfl<G>(q, NB > 2 ? 2 : 0);
    R r = (q0 - posbefore) * K.inv_dt_env * K.vel_weight;
    r += K.alive_bonus;
    r -= K.ctrl_cost * a2;
    {
        R pen = 0;
        if (pen_dof) {
            if ((c.qlo - q) > -K.limit_pen_margin) pen += (R)1.5;
            if ((c.qhi - q) < K.limit_pen_margin) pen += (R)1.5;
        }
        const R pg = gshfl<G>(pen, K.limit_pen_dof >= 0 ? K.limit_pen_dof : 0);
        if (K.limit_pen_dof >= 0) r -= K.limit_pen_weight * pg;
    }
    r -= K.dev_cost * Num<R>::abs_(ang);
    bool lane_ok = true;
    if (c.isb) {
        if (l >= 2 && !(Num<R>::abs_(q) < K.state_bound)) lane_ok = false;
        if (l < 2 && !(Num<R>::abs_(q) < Num<R>::inf())) lane_ok = false;
        if (!(Num<R>::abs_(dq) < K.state_bound)) lane_ok = false;
    }
    bool ok = group_ballot<G>(!lane_ok, gbase) == 0;
    if (K.zero_reward_on_blowup && !ok) r = 0;
    R hgt = 0;
    if (K.height_body >= 0) {
        R cs, sn, px, py;
        coop_fk_positions<T, R>(c, q, cs, sn, px, py);
        const R X = px + cs * K.hcx - sn * K.hcy, Y = py + sn * K.hcx + cs * K.hcy;
        hgt = gshfl<G>(K.wy1 * X + K.wy2 * Y + K.wy0, K.height_body);
        ok = ok && (hgt > K.height_lo) && (hgt < K.height_hi);
    }
    ok = ok && (Num<R>::abs_(ang) < K.ang_max);
    bool done = !ok;
    bool trunc = false;
    // per-world counters: lane 0 of the group reads, everyone learns the value, lane 0 writes back
    const int el = gshfl<G>(el_in + 1, 0);
    const uint32_t ep = gshfl<G>(ep_in, 0);
    if (wactive && a.max_episode_steps > 0) {
        if (el >= a.max_episode_steps) { trunc = !done; done = true; }
        if (l == 0) a.elapsed[w] = (done && a.auto_reset) ? 0 : el;
    }
    const bool do_reset = wactive && done && a.auto_reset;
    if (do_reset) {
        if (c.isb) {   // reset_model(): q0 + U(+-noise), dq0 + U(+-noise), fp32 arithmetic (bit-identical to the oracle)
            const float noise = (float)K.reset_noise;
            const float ua = __fmul_rn(reset_uniform(reset_seed(a, w), reset_world(a, w), ep, l), noise);
            const float ub = __fmul_rn(reset_uniform(reset_seed(a, w), reset_world(a, w), ep, NB + l), noise);
            q = (R)__fadd_rn((float)c.qinit, ua);
            dq = (R)__fadd_rn((float)c.dqinit, ub);
        }
        if (l == 0) a.episode[w] = ep + 1;
        hint = 0xffffffffu;
    }
    if (wactive && l == 0) a.hint[w] = 0xffffffff00000000ull | (uint64_t)hint;
    // obs (of the reset state for auto-reset worlds: gym/vector/sync_vector_env.py:76-79)
    const bool height_obs = K.obs_mode == DARTB_OBS_HEIGHT_Q2_DQ;
    if (height_obs && K.height_body >= 0 && __any_sync(COOP_FULL, do_reset)) {
        R cs, sn, px, py;
        coop_fk_positions<T, R>(c, q, cs, sn, px, py);
        const R X = px + cs * K.hcx - sn * K.hcy, Y = py + sn * K.hcx + cs * K.hcy;
        const R h2 = gshfl<G>(K.wy1 * X + K.wy2 * Y + K.wy0, K.height_body);
        if (do_reset) hgt = h2;
    }
    float* so = sobs + gi * K.n_obs;
    if (c.isb) {
        if (l >= 2) so[l - 1] = (float)q;
        R v = dq;
        if (K.dq_clip > 0) v = v > K.dq_clip ? K.dq_clip : (v < -K.dq_clip ? -K.dq_clip : v);
        so[NB - 1 + l] = (float)v;
        if (height_obs) { if (l == 0) so[0] = (float)hgt; }
        else if (l == 1) so[0] = (float)q;
    }
#ifndef DARTB_HOST_EMU
    if (a.tma) {   // one bulk store of the CTA's [WPB, n_obs] observation tile (full tiles only, see the launcher)
        __syncthreads();
        const int WPB = nwarps * WPW;
        if (threadIdx.x == 0)
            tma_store_1d(a.obs + (size_t)blockIdx.x * WPB * K.n_obs, smraw + (size_t)nwarps * WPW * sizeof(CoopRows<T, R>), WPB * K.n_obs * sizeof(float));
    } else
#endif
    {
        __syncwarp();
        const int cnt = a.n - wb < WPW ? a.n - wb : WPW;   // worlds of this warp that exist
        if (cnt > 0) for (int k = lane; k < cnt * K.n_obs; k += 32) store_obs(a, (size_t)wb * K.n_obs + k, sobs[k]);
    }
    if (mine) { a.q[(size_t)l * a.n + w] = q; a.dq[(size_t)l * a.n + w] = dq; }
    if (wactive && l == 0) {
        if (a.reward64) { a.reward64[w] = (double)r; a.done[w] = done ? 1 : 0; }
        else {
            a.reward[w] = (float)r;
            a.done[w] = (uint8_t)((done ? 1 : 0) | (trunc ? 2 : 0));  // bit 0 done, bit 1 TimeLimit.truncated
        }
        if (a.truncated) a.truncated[w] = trunc ? 1 : 0;
    }
}
