// warp_group.cuh — collectives of a GROUP of G lanes of a warp (G = 4, 8, 16) and the boxed-LCP solver whose tableau
// is distributed one row per lane of such a group.  Shared by the lane-cooperative kernels (planar_coop.cuh: G = 8 / 16
// lanes per world, one body per lane) and by the quad form of the per-thread kernels (planar_kernels.cuh::substep with
// G = 4: the whole world on every lane, constraint rows dealt out over the four lanes).
#pragma once
#ifdef DARTB_HOST_EMU
#include "../../tools/host_emu/simt.h"
#endif

#define COOP_FULL 0xffffffffu

// products / sums that must not be contracted into FMAs (bit-identical results across code paths)
#ifdef DARTB_HOST_EMU
static inline double coop_mul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline float coop_mul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline double coop_add_rn(double a, double b) { volatile double r = a + b; return r; }
static inline float coop_add_rn(float a, float b) { volatile float r = a + b; return r; }
#else
DEVI double coop_mul_rn(double a, double b) { return __dmul_rn(a, b); }
DEVI float coop_mul_rn(float a, float b) { return __fmul_rn(a, b); }
DEVI double coop_add_rn(double a, double b) { return __dadd_rn(a, b); }
DEVI float coop_add_rn(float a, float b) { return __fadd_rn(a, b); }
#endif

// ------------------------------------------------------------------------ group collectives
template <int G, typename V>
DEVI V gshfl(V v, int src) { return __shfl_sync(COOP_FULL, v, src, G); }

DEVI int coop_warp_max(int v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { const int o = __shfl_xor_sync(COOP_FULL, v, m); v = o > v ? o : v; }
    return v;
}
template <int G, typename R>
DEVI R group_sum(R v) {
#pragma unroll
    for (int m = G / 2; m >= 1; m >>= 1) v = coop_add_rn(v, __shfl_xor_sync(COOP_FULL, v, m));
    return v;
}
template <int G, typename R>
DEVI R group_max(R v) {
#pragma unroll
    for (int m = G / 2; m >= 1; m >>= 1) { const R o = __shfl_xor_sync(COOP_FULL, v, m); v = o > v ? o : v; }
    return v;
}
template <int G>
DEVI unsigned group_or(unsigned v) {
#pragma unroll
    for (int m = G / 2; m >= 1; m >>= 1) v |= __shfl_xor_sync(COOP_FULL, v, m);
    return v;
}
// the group's slice of a warp ballot, as bits 0..G-1
template <int G>
DEVI unsigned group_ballot(bool p, int gbase) { return (__ballot_sync(COOP_FULL, p) >> gbase) & ((G == 32) ? 0xffffffffu : ((1u << G) - 1u)); }

// ------------------------------------------------------------------------ K6: distributed boxed LCP
// Row r = l + h*G lives on lane l (slot h).  Same iteration as lcp_ppt (sets visited, tolerances,
// two-stage friction bounds, iterative refinement); `live` lanes belong to groups still solving.
// NCx = column class (>= the warp's largest row count): 4, 8, 16 or 32; RPL rows per lane follow.
template <int G_, typename R, int NCx>
struct GroupLcp {
    static constexpr int G = G_, NC = NCx, RPL = (NCx + G - 1) / G;

    // gather a per-row quantity v[h] of every row c < nmax into out[c]
    static DEVI void gather_rows(const R (&v)[RPL], R (&out)[NC], int nmax) {
#pragma unroll
        for (int cc = 0; cc < NC; cc++) {
            out[cc] = gshfl<G>(v[cc / G], cc % G);
        }
    }

    // one principal exchange on pivot row k (per group; k < 0: nothing to do for this group)
    static DEVI bool exchange(R (&Tb)[RPL][NC], int k, int l, int nmax) {
        const bool has = k >= 0;
        const int ko = has ? (k % G) : l, kh = has ? (k / G) : 0;
        R rk[NC];
        R d = 1;
#pragma unroll
        for (int cc = 0; cc < NC; cc++) {
            R mine = Tb[0][cc];
#pragma unroll
            for (int h2 = 1; h2 < RPL; h2++) if (kh == h2) mine = Tb[h2][cc];
            rk[cc] = gshfl<G>(mine, ko);
            if (cc == k) d = rk[cc];
        }
        const bool pd = !has || d > 0;
        const R p = Num<R>::rcp_(d);
#pragma unroll
        for (int h = 0; h < RPL; h++) {
            const bool prow = has && (l + h * G == k);
            R ck = 0;
#pragma unroll
            for (int cc = 0; cc < NC; cc++) { if (cc == k) ck = Tb[h][cc]; }
#pragma unroll
            for (int cc = 0; cc < NC; cc++) {
                const R rp = rk[cc] * p;
                const R other = (cc == k) ? ck * p : Tb[h][cc] - ck * rp;
                const R piv = (cc == k) ? p : -rp;
                if (has && pd) Tb[h][cc] = prow ? piv : other;
            }
        }
        return pd;
    }

    // Solve.  Per row slot h: A row, b, lo, hi (mu in hi for friction rows), fi (normal row of a friction
    // row, else -1), valid.  hin: hinted set of a friction row (3 = none).  Out: x, st.
    static DEVI bool solve(int l, int gbase, int n, int nmax, const R (&A)[RPL][NC], const R (&b)[RPL],
                           R (&lo)[RPL], R (&hi)[RPL], const int (&fi)[RPL], const unsigned (&hin)[RPL], R (&x)[RPL],
                           unsigned (&st)[RPL]) {
        const R INF = Num<R>::inf();
        R Tb[RPL][NC], sd[RPL], mu[RPL];
        unsigned cur[RPL];
        bool valid[RPL], fricrow[RPL];
#pragma unroll
        for (int h = 0; h < RPL; h++) {
            const int r = l + h * G;
            valid[h] = r < n;
            R diag = 1;
#pragma unroll
            for (int cc = 0; cc < NC; cc++) {
                Tb[h][cc] = (valid[h] && cc < n) ? A[h][cc] : (cc == r ? (R)1 : (R)0);
                if (cc == r) diag = Tb[h][cc];
            }
            sd[h] = Num<R>::sqrt_tol_(diag);
            mu[h] = hi[h];
            x[h] = 0;
            fricrow[h] = valid[h] && fi[h] >= 0;
            unsigned s = 0, cu = 3;
            if (!valid[h] || !(diag > Num<R>::inert())) s = 3;
            else if (fi[h] >= 0) s = 3;
            else if (lo[h] == 0 && hi[h] == INF) { s = b[h] > 0 ? 0u : 1u; cu = 1; }
            else if (hi[h] == 0 && lo[h] == -INF) { s = b[h] < 0 ? 0u : 2u; cu = 2; }
            st[h] = s; cur[h] = cu;
            if (!valid[h]) { lo[h] = 0; hi[h] = 0; }
        }
        bool gok = true;       // this group's solve has not failed
        bool gfin = n == 0;    // this group is finished (nothing more to do in any stage)
#pragma unroll 1
        for (int stage = 0; stage < 2; stage++) {
            if (stage == 1) {
                bool mine = false;
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    // x of the normal row this friction row hangs on
                    const int f = fricrow[h] ? fi[h] : l;
                    R xn = gshfl<G>(x[0], f % G);
#pragma unroll
                    for (int h2 = 1; h2 < RPL; h2++) { const R xh = gshfl<G>(x[h2], f % G); if (f / G == h2) xn = xh; }
                    R diag = 0;
#pragma unroll
                    for (int cc = 0; cc < NC; cc++) { if (cc == l + h * G) diag = Tb[h][cc]; }
                    if (fricrow[h] && diag > Num<R>::inert() && !gfin && gok) {
                        const R hh = Num<R>::abs_(mu[h] * xn);
                        hi[h] = hh; lo[h] = -hh;
                        if (hh == 0) st[h] = 3;
                        else { mine = true; st[h] = hin[h] < 3u ? hin[h] : 0u; }
                    }
                }
                const bool any = group_ballot<G>(mine, gbase) != 0;
                if (!any) gfin = true;
            }
            if (!__any_sync(COOP_FULL, !gfin && gok)) { if (stage == 0) continue; else break; }
            int best = NC + 1, tries = 3;
            bool done = gfin || !gok;
#pragma unroll 1
            for (int it = 0; it < 6 + 3 * nmax; it++) {
                if (!__any_sync(COOP_FULL, !done)) break;
                // rows whose free/bound status differs between the tableau and the set to evaluate
                unsigned flip = 0;
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    const bool f = !done && ((st[h] == 0) != (cur[h] == 0));
                    flip |= group_ballot<G>(f, gbase) << (h * G);
                }
#pragma unroll 1
                while (__any_sync(COOP_FULL, flip != 0)) {
                    const int k = flip ? (__ffs(flip) - 1) : -1;
                    const bool pd = exchange(Tb, k, l, nmax);
                    if (!pd) { gok = false; done = true; flip = 0; }
                    flip &= flip - 1;
                }
                R z[RPL], zg[NC], y[RPL];
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    cur[h] = done ? cur[h] : st[h];
                    // (selects, not branch trees: a warp with one or two resident neighbours pays for every branch)
                    z[h] = Num<R>::sel_(st[h] == 0, b[h], Num<R>::sel_(st[h] == 1, lo[h], Num<R>::sel_(st[h] == 2, hi[h], (R)0)));
                }
                gather_rows(z, zg, nmax);
                R axm = 0, Ssum = 0;
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    R s = 0;
#pragma unroll
                    for (int cc = 0; cc < NC; cc++) { s += Tb[h][cc] * zg[cc]; }
                    y[h] = s;
                    const R xv = Num<R>::sel_(st[h] == 0, s, z[h]);
                    x[h] = Num<R>::sel_(done, x[h], xv);
                    const R ax = valid[h] ? Num<R>::abs_(x[h]) : (R)0;
                    axm = ax > axm ? ax : axm;
                    Ssum += sd[h] * ax;
                }
                const R xs = group_max<G>(axm), S = group_sum<G>(Ssum);
                const R tx = Num<R>::lcp_tol() * xs;
                unsigned badm = 0, nst[RPL];
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    const bool below = x[h] < lo[h] - tx, above = x[h] > hi[h] + tx;
                    const R w = y[h] - b[h];
                    const R tw = Num<R>::lcp_tol() * (Num<R>::abs_(b[h]) + sd[h] * S);
                    const bool wbad = ((st[h] == 1 && w < -tw) || (st[h] == 2 && w > tw)) && lo[h] < hi[h];
                    const bool bad = st[h] == 0 ? (below || above) : wbad;      // (st == 3 makes wbad false)
                    const unsigned ns = st[h] == 0 ? (below ? 1u : 2u) : 0u;
                    nst[h] = bad ? ns : st[h];
                    badm |= group_ballot<G>(bad && !done, gbase) << (h * G);
                }
                if (done) continue;
                const int nbad = __popc(badm);
                if (nbad == 0) { done = true; continue; }
                if (nbad < best) {
                    best = nbad; tries = 3;
#pragma unroll
                    for (int h = 0; h < RPL; h++) st[h] = nst[h];
                } else if (tries > 0) {
                    tries--;
#pragma unroll
                    for (int h = 0; h < RPL; h++) st[h] = nst[h];
                } else {  // Murty: flip only the highest-index infeasible row
                    const int kk = 31 - __clz(badm);
#pragma unroll
                    for (int h = 0; h < RPL; h++) if (l + h * G == kk) st[h] = nst[h];
                }
            }
            // converged iff every row's tableau matches st and the last check found nothing
            {
                // a group that ran out of iterations has done == false here
                if (!done) gok = false;
            }
            // one round of iterative refinement against A (see lcp_ppt)
            if (__any_sync(COOP_FULL, gok && !gfin)) {
                R xg[NC], rr[RPL], rg[NC];
                gather_rows(x, xg, nmax);
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    R s = 0;
                    if (h * G < nmax && valid[h] && st[h] == 0) {
                        s = b[h];
#pragma unroll
                        for (int cc = 0; cc < NC; cc++) { if (cc < n) s -= A[h][cc] * xg[cc]; }
                    }
                    rr[h] = s;
                }
                gather_rows(rr, rg, nmax);
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    R s = 0;
#pragma unroll
                    for (int cc = 0; cc < NC; cc++) { s += Tb[h][cc] * rg[cc]; }
                    if (gok && !gfin && valid[h] && st[h] == 0) x[h] += s;
                }
            }
        }
        return gok;
    }
};
