// planar_loop.cuh — the same DART time step as planar_kernels.cuh::substep, written as RUNTIME
// loops over bodies / shapes / rows with the topology read from the model (constant bank).
//
// Why it exists (profiles/r1_*): the fully unrolled per-topology stepper executes ~7.4k distinct
// instructions per sub-step (118 KB of SASS), far more than the 32 KB L1.5 instruction cache, and
// ncu shows stall_no_inst as its top stall reason (40%, 75% inside the straight-line ABA code).
// This variant keeps the instruction footprint at a few thousand instructions: per-body
// quantities live in thread-local arrays (hardware-interleaved, so a warp's accesses to one
// slot are one 128 B line) and every pass is a short loop body that stays cache resident.
// It is also topology-generic: any planar prismatic/revolute skeleton that fits LOOP_MAXB runs
// without a dedicated instantiation.
//
// The arithmetic is statement-for-statement the unrolled version's (same formulas, same order),
// so both variants agree to rounding; tests pin each against the oracle.
#pragma once
#include "planar_kernels.cuh"

#define LOOP_MAXB PM_MAXB
#define LOOP_MAXS PM_MAXS
#define LOOP_MAXR 32 /* rows: (n,t) per capsule + active limits + Coulomb-friction rows (<= 32: 64-bit state word) */

template <typename R>
DEVI void fk_positions_loop(const PModel<R>& M, const R* q, R* cs, R* sn, R* px, R* py) {
    R th[LOOP_MAXB];
    const int nb = M.nb;
#pragma unroll 1
    for (int i = 0; i < nb; i++) {
        const int par = M.parent[i];
        R cp = 1, sp = 0, ppx = 0, ppy = 0, thp = 0;
        if (par >= 0) { cp = cs[par]; sp = sn[par]; ppx = px[par]; ppy = py[par]; thp = th[par]; }
        const R arx = cp * M.ax[i] - sp * M.ay[i], ary = sp * M.ax[i] + cp * M.ay[i];
        if (M.jtype[i] == PM_REV) {
            th[i] = thp + M.sgn[i] * q[i];
            Num<R>::sincos_(th[i], &sn[i], &cs[i]);
            px[i] = ppx + arx; py[i] = ppy + ary;
        } else {
            th[i] = thp; cs[i] = cp; sn[i] = sp;
            const R uwx = cp * M.ux[i] - sp * M.uy[i], uwy = sp * M.ux[i] + cp * M.uy[i];
            px[i] = ppx + arx + uwx * q[i]; py[i] = ppy + ary + uwy * q[i];
        }
    }
}

// One DART time step; q, dq in/out (length M.nb).  FEXT / FLUID as in planar_kernels.cuh, but
// runtime flags (warp-uniform branches) so there is a single copy of the code.
template <typename R>
DEVI void substep_loop(const PModel<R>& M, R* q, R* dq, const R* tau, const bool FEXT, const R* eft, const R* efx,
                       const R* efy, const bool FLUID, R fluid_offset, R fluid_coef, int lcp_mode, int pgs_iters,
                       const ContactSink<R>* sink, int world, const R* wp = nullptr, size_t wn = 0) {
    // wp: this world's column of the per-world dynamics parameters (StepArgs::wpar + world, row stride wn) or null
    constexpr int MB = LOOP_MAXB, NR = LOOP_MAXR;
    const int nb = M.nb;
    auto mass_of = [&](int i) { return wp ? wp[(size_t)i * wn] : M.mass[i]; };
    auto cx_of = [&](int i) { return wp ? wp[(size_t)(nb + i) * wn] : M.cx[i]; };
    auto cy_of = [&](int i) { return wp ? wp[(size_t)(2 * nb + i) * wn] : M.cy[i]; };
    auto izz_of = [&](int i) { return wp ? wp[(size_t)(3 * nb + i) * wn] : M.izz[i]; };
    const R dt = M.dt;
    R th[MB], cs[MB], sn[MB], px[MB], py[MB], rx[MB], ry[MB], wz[MB], vx[MB], vy[MB], ex[MB], ey[MB];
    R s0[MB], s1[MB], s2[MB];  // joint motion subspace in world axes at the body origin: [sgn; uw]
    // ---------------- K1
#pragma unroll 1
    for (int i = 0; i < nb; i++) {
        const int par = M.parent[i];
        R cp = 1, sp = 0, ppx = 0, ppy = 0, wp = 0, vpx = 0, vpy = 0, thp = 0;
        if (par >= 0) { cp = cs[par]; sp = sn[par]; ppx = px[par]; ppy = py[par]; wp = wz[par]; vpx = vx[par]; vpy = vy[par]; thp = th[par]; }
        const R arx = cp * M.ax[i] - sp * M.ay[i], ary = sp * M.ax[i] + cp * M.ay[i];
        if (M.jtype[i] == PM_REV) {
            const R t = thp + M.sgn[i] * q[i];
            th[i] = t;
            Num<R>::sincos_(t, &sn[i], &cs[i]);
            rx[i] = arx; ry[i] = ary;
            const R sd = M.sgn[i] * dq[i];
            wz[i] = wp + sd;
            const R vxi = vpx - wp * ary, vyi = vpy + wp * arx;
            vx[i] = vxi; vy[i] = vyi;
            ex[i] = sd * vyi; ey[i] = -sd * vxi;
            s0[i] = M.sgn[i]; s1[i] = 0; s2[i] = 0;
            px[i] = ppx + arx; py[i] = ppy + ary;
        } else {
            th[i] = thp; cs[i] = cp; sn[i] = sp;
            const R uwx = cp * M.ux[i] - sp * M.uy[i], uwy = sp * M.ux[i] + cp * M.uy[i];
            const R rxi = arx + uwx * q[i], ryi = ary + uwy * q[i];
            rx[i] = rxi; ry[i] = ryi;
            wz[i] = wp;
            vx[i] = vpx - wp * ryi + uwx * dq[i]; vy[i] = vpy + wp * rxi + uwy * dq[i];
            ex[i] = -wp * uwy * dq[i]; ey[i] = wp * uwx * dq[i];
            s0[i] = 0; s1[i] = uwx; s2[i] = uwy;
            px[i] = ppx + rxi; py[i] = ppy + ryi;
        }
    }
    // ---------------- K2
    R U0[MB], U1[MB], U2[MB], Di[MB], uu[MB];
    R aJ[MB], ahx[MB], ahy[MB], ama[MB], amb[MB], amc[MB], apt[MB], apx[MB], apy[MB];
#pragma unroll 1
    for (int i = 0; i < nb; i++) { aJ[i] = 0; ahx[i] = 0; ahy[i] = 0; ama[i] = 0; amb[i] = 0; amc[i] = 0; apt[i] = 0; apx[i] = 0; apy[i] = 0; }
#pragma unroll 1
    for (int i = nb - 1; i >= 0; i--) {
        const int par = M.parent[i];
        const R m = mass_of(i), cxi = cx_of(i), cyi = cy_of(i);
        const R dx = cs[i] * cxi - sn[i] * cyi, dy = sn[i] * cxi + cs[i] * cyi;
        const R J = izz_of(i) + m * (dx * dx + dy * dy) + aJ[i], hx = -m * dy + ahx[i], hy = m * dx + ahy[i];
        const R ma = m + ama[i], mb = amb[i], mc = m + amc[i];
        const R Px = m * (vx[i] - wz[i] * dy), Py = m * (vy[i] + wz[i] * dx);
        R pt = vx[i] * Py - vy[i] * Px - (dx * m * M.gy - dy * m * M.gx);
        R pfx = -wz[i] * Py - m * M.gx, pfy = wz[i] * Px - m * M.gy;
        if (FEXT) { pt -= eft[i]; pfx -= efx[i]; pfy -= efy[i]; }
        if (FLUID) {
            const R nx = cs[i] * M.fnx[i] - sn[i] * M.fny[i], ny = sn[i] * M.fnx[i] + cs[i] * M.fny[i];
            const R vwx = vx[i] - wz[i] * dy, vwy = vy[i] + wz[i] * dx;
            const R vcx = cs[i] * vwx + sn[i] * vwy, vcy = cs[i] * vwy - sn[i] * vwx;   // body coordinates (see planar_kernels.cuh)
            const R crx = -wz[i] * ny, cry = wz[i] * nx;
            const R dp = (vcx + crx * fluid_offset) * nx + (vcy + cry * fluid_offset) * ny;
            const R dn = (vcx - crx * fluid_offset) * nx + (vcy - cry * fluid_offset) * ny;
            R ffx = 0, ffy = 0;
            if (dp > 0) { ffx = -fluid_coef * dp * nx; ffy = -fluid_coef * dp * ny; }
            if (dn < 0) { ffx = -fluid_coef * dn * nx; ffy = -fluid_coef * dn * ny; }
            const R oxw = cs[i] * M.ox[i] - sn[i] * M.oy[i], oyw = sn[i] * M.ox[i] + cs[i] * M.oy[i];
            pt -= oxw * ffy - oyw * ffx; pfx -= ffx; pfy -= ffy;
        }
        pt += apt[i]; pfx += apx[i]; pfy += apy[i];
        const R t0 = hx * ex[i] + hy * ey[i], t1 = ma * ex[i] + mb * ey[i], t2 = mb * ex[i] + mc * ey[i];
        R u0, u1, u2, D, u;
        if (M.jtype[i] == PM_REV) {
            const R s = s0[i];
            u0 = s * J; u1 = s * hx; u2 = s * hy;
            D = J;
            u = tau[i] - s * (pt + t0);
        } else {
            const R a = s1[i], b = s2[i];
            u0 = hx * a + hy * b; u1 = ma * a + mb * b; u2 = mb * a + mc * b;
            D = a * u1 + b * u2;
            u = tau[i] - (a * (pfx + t1) + b * (pfy + t2));
        }
        u += -M.kspring[i] * (q[i] - M.rest[i] + dt * dq[i]) - M.damping[i] * dq[i];
        D += dt * M.damping[i] + dt * dt * M.kspring[i];
        const R di = Num<R>::rcp_(D);
        U0[i] = u0; U1[i] = u1; U2[i] = u2; Di[i] = di; uu[i] = u;
        if (par >= 0) {
            const R g = u * di;
            const R pa0 = pt + t0 + u0 * g, pa1 = pfx + t1 + u1 * g, pa2 = pfy + t2 + u2 * g;
            const R P00 = J - u0 * u0 * di, P01 = hx - u0 * u1 * di, P02 = hy - u0 * u2 * di;
            const R P11 = ma - u1 * u1 * di, P12 = mb - u1 * u2 * di, P22 = mc - u2 * u2 * di;
            const R kx = -ry[i], ky = rx[i];
            const R nhx = P01 + P11 * kx + P12 * ky, nhy = P02 + P12 * kx + P22 * ky;
            aJ[par] += P00 + kx * (P01 + nhx) + ky * (P02 + nhy);
            ahx[par] += nhx; ahy[par] += nhy; ama[par] += P11; amb[par] += P12; amc[par] += P22;
            apt[par] += pa0 + kx * pa1 + ky * pa2; apx[par] += pa1; apy[par] += pa2;
        }
    }
    // ---------------- K3
    {
        R a0[MB], a1[MB], a2[MB];
#pragma unroll 1
        for (int i = 0; i < nb; i++) {
            const int par = M.parent[i];
            R p0 = 0, p1 = 0, p2 = 0;
            if (par >= 0) { p0 = a0[par]; p1 = a1[par] - a0[par] * ry[i]; p2 = a2[par] + a0[par] * rx[i]; }
            const R dd = Di[i] * (uu[i] - (U0[i] * p0 + U1[i] * p1 + U2[i] * p2));
            a0[i] = p0 + s0[i] * dd; a1[i] = p1 + ex[i] + s1[i] * dd; a2[i] = p2 + ey[i] + s2[i] * dd;
            dq[i] += dt * dd;
        }
    }
    // ---------------- K4/K5
    int n = 0, nc = 0;
    R Jr[NR * MB], bb[NR], lo[NR], hi[NR];
    int fidx[NR];
    R cpx[LOOP_MAXS], cpy[LOOP_MAXS], cnx[LOOP_MAXS], cny[LOOP_MAXS], cdep[LOOP_MAXS];
    int crow[LOOP_MAXS], cshape[LOOP_MAXS];
    const R INF = Num<R>::inf();
    if (M.has_ground) {
        const R inv_dt = (R)1 / dt;
        const int ns = M.ns;
#pragma unroll 1
        for (int s = 0; s < ns; s++) {
            const int b = M.sbody[s];
            const R cb = cs[b], sb = sn[b];
            const R ccx = px[b] + cb * M.scx[s] - sb * M.scy[s], ccy = py[b] + sb * M.scx[s] + cb * M.scy[s];
            const R adx = cb * M.sdx[s] - sb * M.sdy[s], ady = sb * M.sdx[s] + cb * M.sdy[s];
            const R hl = M.shalf[s], rad = M.srad[s];
            const R ex_ = hl * Num<R>::abs_(adx) + rad + (R)1e-5, ey_ = hl * Num<R>::abs_(ady) + rad + (R)1e-5;
            const bool near_ = Num<R>::abs_(ccx - M.gcx) <= M.ghx + ex_ && Num<R>::abs_(ccy - M.gcy) <= M.ghy + ey_;
            R lx = 0, ly = 0, ddx = 0, ddy = 0, d = INF;
            if (near_) {  // broad-phase AABB reject (see planar_kernels.cuh)
                closest_segment_box2<R>(ccx + hl * adx, ccy + hl * ady, ccx - hl * adx, ccy - hl * ady, M.gcx, M.gcy, M.ghx,
                                        M.ghy, lx, ly, ddx, ddy);
                d = Num<R>::sqrt_(ddx * ddx + ddy * ddy);
            }
            if (!(d > rad) && n + 2 <= NR) {
                R nx, ny, depth, Px, Py;
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
                const R mu = wp ? wp[(size_t)(4 * nb + s) * wn] : M.smu[s];
                const bool fric = mu > (R)DK_FRICTION_THRESHOLD;
                const R tx = -ny, ty = nx;
                const int r0 = n;
                R vn = 0, vt = 0;
#pragma unroll 1
                for (int j = 0; j < nb; j++) { Jr[r0 * MB + j] = 0; if (fric) Jr[(r0 + 1) * MB + j] = 0; }
#pragma unroll 1
                for (int j = b; j >= 0; j = M.parent[j]) {  // ancestors of b (incl. b)
                    const R ax_ = s1[j] - s0[j] * (Py - py[j]), ay_ = s2[j] + s0[j] * (Px - px[j]);
                    const R jn = ax_ * nx + ay_ * ny, jt = ax_ * tx + ay_ * ty;
                    vn += jn * dq[j]; vt += jt * dq[j];
                    Jr[r0 * MB + j] = jn;
                    if (fric) Jr[(r0 + 1) * MB + j] = jt;
                }
                R bounce = depth;
                if (bounce < 0) bounce = 0;
                else { bounce *= inv_dt * (R)DK_CONTACT_ERP; if (bounce > (R)DK_CONTACT_MAX_ERV) bounce = (R)DK_CONTACT_MAX_ERV; }
                bb[r0] = -vn + bounce; lo[r0] = 0; hi[r0] = INF; fidx[r0] = -1;
                n = r0 + 1;
                if (fric) { bb[r0 + 1] = -vt; lo[r0 + 1] = -mu; hi[r0 + 1] = mu; fidx[r0 + 1] = r0; n = r0 + 2; }
                cpx[nc] = Px; cpy[nc] = Py; cnx[nc] = nx; cny[nc] = ny; cdep[nc] = depth; crow[nc] = r0 | (fric ? 0x100 : 0);
                cshape[nc] = s;
                nc++;
            }
        }
    }
    const int n_contact_rows = n;
#pragma unroll 1
    for (int i = 0; i < nb; i++) {
        if (!M.limited[i]) continue;
        int act = 0;
        if (q[i] - M.qlo[i] <= 0) act = -1;
        else if (q[i] - M.qhi[i] >= 0) act = 1;
        if (act != 0 && n < NR) {
#pragma unroll 1
            for (int j = 0; j < nb; j++) Jr[n * MB + j] = (j == i) ? (R)1 : (R)0;
            bb[n] = -dq[i];
            if (act < 0) { lo[n] = 0; hi[n] = INF; } else { lo[n] = -INF; hi[n] = 0; }
            fidx[n] = -1;
            n++;
        }
    }
    if (M.any_coulomb) {
        // JointCoulombFrictionConstraint rows (after the limit rows, like DART): b = -dq, |x| <= friction*dt
#pragma unroll 1
        for (int i = 0; i < nb; i++) {
            if (M.coulomb[i] == 0 || dq[i] == 0 || n >= NR) continue;
#pragma unroll 1
            for (int j = 0; j < nb; j++) Jr[n * MB + j] = (j == i) ? (R)1 : (R)0;
            bb[n] = -dq[i];
            hi[n] = M.coulomb[i] * dt; lo[n] = -hi[n];
            fidx[n] = -1;
            n++;
        }
    }
    if (n > 0) {
        // plain articulated inertia (reuses the accumulator arrays)
        R V0[MB], V1[MB], V2[MB], Ei[MB];
#pragma unroll 1
        for (int i = 0; i < nb; i++) { aJ[i] = 0; ahx[i] = 0; ahy[i] = 0; ama[i] = 0; amb[i] = 0; amc[i] = 0; }
#pragma unroll 1
        for (int i = nb - 1; i >= 0; i--) {
            const int par = M.parent[i];
            const R m = mass_of(i), cxi = cx_of(i), cyi = cy_of(i);
            const R dx = cs[i] * cxi - sn[i] * cyi, dy = sn[i] * cxi + cs[i] * cyi;
            const R J = izz_of(i) + m * (dx * dx + dy * dy) + aJ[i], hx = -m * dy + ahx[i], hy = m * dx + ahy[i];
            const R ma = m + ama[i], mb = amb[i], mc = m + amc[i];
            R v0, v1, v2, D;
            if (M.jtype[i] == PM_REV) { const R s = s0[i]; v0 = s * J; v1 = s * hx; v2 = s * hy; D = J; }
            else {
                const R a = s1[i], b = s2[i];
                v0 = hx * a + hy * b; v1 = ma * a + mb * b; v2 = mb * a + mc * b;
                D = a * v1 + b * v2;
            }
            const R di = Num<R>::rcp_(D);
            V0[i] = v0; V1[i] = v1; V2[i] = v2; Ei[i] = di;
            if (par >= 0) {
                const R P00 = J - v0 * v0 * di, P01 = hx - v0 * v1 * di, P02 = hy - v0 * v2 * di;
                const R P11 = ma - v1 * v1 * di, P12 = mb - v1 * v2 * di, P22 = mc - v2 * v2 * di;
                const R kx = -ry[i], ky = rx[i];
                const R nhx = P01 + P11 * kx + P12 * ky, nhy = P02 + P12 * kx + P22 * ky;
                aJ[par] += P00 + kx * (P01 + nhx) + ky * (P02 + nhy);
                ahx[par] += nhx; ahy[par] += nhy; ama[par] += P11; amb[par] += P12; amc[par] += P22;
            }
        }
        R MJ[NR * MB];
        R A[NR * NR], x[NR];
#pragma unroll 1
        for (int r = 0; r < n; r++) {
            R ur[MB], a0[MB], a1[MB], a2[MB];
#pragma unroll 1
            for (int i = 0; i < nb; i++) { apt[i] = 0; apx[i] = 0; apy[i] = 0; }
#pragma unroll 1
            for (int i = nb - 1; i >= 0; i--) {
                const int par = M.parent[i];
                const R pt = apt[i], pfx = apx[i], pfy = apy[i];
                R u;
                if (M.jtype[i] == PM_REV) u = Jr[r * MB + i] - s0[i] * pt;
                else u = Jr[r * MB + i] - (s1[i] * pfx + s2[i] * pfy);
                ur[i] = u;
                if (par >= 0) {
                    const R g = u * Ei[i];
                    const R pa0 = pt + V0[i] * g, pa1 = pfx + V1[i] * g, pa2 = pfy + V2[i] * g;
                    apt[par] += pa0 - ry[i] * pa1 + rx[i] * pa2; apx[par] += pa1; apy[par] += pa2;
                }
            }
#pragma unroll 1
            for (int i = 0; i < nb; i++) {
                const int par = M.parent[i];
                R p0 = 0, p1 = 0, p2 = 0;
                if (par >= 0) { p0 = a0[par]; p1 = a1[par] - a0[par] * ry[i]; p2 = a2[par] + a0[par] * rx[i]; }
                const R dd = Ei[i] * (ur[i] - (V0[i] * p0 + V1[i] * p1 + V2[i] * p2));
                a0[i] = p0 + s0[i] * dd; a1[i] = p1 + s1[i] * dd; a2[i] = p2 + s2[i] * dd;
                MJ[r * MB + i] = dd;
            }
            for (int s = 0; s <= r; s++) {  // row r of A = J M^-1 J^T (lower triangle, mirrored)
                R v = 0;
                for (int j = 0; j < nb; j++) v += Jr[s * MB + j] * MJ[r * MB + j];
                A[r * n + s] = v;
                A[s * n + r] = v;
            }
        }
        for (int r = 0; r < n; r++) A[r * n + r] *= (R)1 + (r < n_contact_rows ? (R)DK_CONTACT_CFM : (R)DK_LIMIT_CFM);
        if (lcp_mode == 1) lcp_pgs<R>(n, A, x, bb, lo, hi, fidx, pgs_iters);
        else lcp_exact<R, NR>(n, A, x, bb, lo, hi, fidx);
        for (int r = 0; r < n; r++) {
            const R xr = x[r];
            for (int j = 0; j < nb; j++) dq[j] += MJ[r * MB + j] * xr;
        }
        if (sink && sink->data) {
            const R inv_dt = (R)1 / dt;
            for (int c = 0; c < nc && c < sink->maxc; c++) {
                const int r0 = crow[c] & 0xff;
                const R xn = x[r0], xt = (crow[c] & 0x100) ? x[r0 + 1] : (R)0;
                const R fx = (cnx[c] * xn - cny[c] * xt) * inv_dt, fy = (cny[c] * xn + cnx[c] * xt) * inv_dt;
                float* o = sink->data + ((size_t)world * sink->maxc + c) * 10;
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
#pragma unroll 1
    for (int i = 0; i < nb; i++) q[i] += dt * dq[i];
}
