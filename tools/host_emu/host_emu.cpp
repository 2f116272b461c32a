// host_emu.cpp — runs the product's device code (planar_kernels.cuh, lower.h) on the CPU, one
// world at a time, for numerics debugging and the no-GPU regression tests
// (tests/test_kernel_host_emulation.py).  TEST TOOL: not linked into libdartb.so, not a fallback.
#define DARTB_HOST_EMU 1
#include <string>
#include <vector>

#include "../../dart_env_b200/csrc/lower.h"
#include "../../dart_env_b200/csrc/planar_kernels.cuh"
#include "../../dart_env_b200/csrc/planar_loop.cuh"
#include "simt.h"
#include "../../dart_env_b200/csrc/planar_coop.cuh"
#include "../../dart_env_b200/csrc/task_kinds.cuh"

template <class T, typename R>
static void run_substep(const PModel<R>& M, int n, const double* q_in, const double* dq_in, const double* tau_in,
                        const double* fext, int lcp_mode, int pgs_iters, double* q_out, double* dq_out, int32_t* count,
                        int32_t* body, float* data, int maxc, uint64_t* hints) {
    constexpr int NB = T::NB;
    for (int w = 0; w < n; w++) {
        R q[NB], dq[NB], tau[NB], eft[NB], efx[NB], efy[NB];
        for (int i = 0; i < NB; i++) { q[i] = (R)q_in[w * NB + i]; dq[i] = (R)dq_in[w * NB + i]; tau[i] = tau_in ? (R)tau_in[w * NB + i] : (R)0; eft[i] = efx[i] = efy[i] = 0; }
        ContactSink<R> sink;
        sink.count = count; sink.body = body; sink.data = data; sink.maxc = maxc;
        if (fext) {
            R cs[NB], sn[NB], px[NB], py[NB];
            fk_positions<T, R>(M, q, cs, sn, px, py);
            for (int k = 0; k < M.nbd; k++) {
                const double* f = fext + ((size_t)w * M.nbd + k) * 3;
                const R fx = (R)(M.e1[0] * f[0] + M.e1[1] * f[1] + M.e1[2] * f[2]);
                const R fy = (R)(M.e2[0] * f[0] + M.e2[1] * f[1] + M.e2[2] * f[2]);
                const int g = M.dgroup[k];
                const R ox = cs[g] * M.dox[k] - sn[g] * M.doy[k], oy = sn[g] * M.dox[k] + cs[g] * M.doy[k];
                eft[g] += ox * fy - oy * fx; efx[g] += fx; efy[g] += fy;
            }
            { uint64_t hint = ~(uint64_t)0; substep<T, R, true, false>(M, q, dq, tau, eft, efx, efy, (R)0, (R)0, lcp_mode, pgs_iters, &sink, w, hint); }
        } else {
            { uint64_t hint = hints ? hints[w] : ~(uint64_t)0; substep<T, R, false, false>(M, q, dq, tau, eft, efx, efy, (R)0, (R)0, lcp_mode, pgs_iters, &sink, w, hint); if (hints) hints[w] = hint; }
        }
        for (int i = 0; i < NB; i++) { q_out[w * NB + i] = (double)q[i]; dq_out[w * NB + i] = (double)dq[i]; }
    }
}

template <typename R>
static void run_substep_loop(const PModel<R>& M, int n, const double* q_in, const double* dq_in, const double* tau_in,
                             const double* fext, int lcp_mode, int pgs_iters, double* q_out, double* dq_out,
                             int32_t* count, int32_t* body, float* data, int maxc, const R* wpar = nullptr) {
    const int NB = M.nb;
    for (int w = 0; w < n; w++) {
        const R* wp = wpar ? wpar + w : nullptr;
        R q[LOOP_MAXB], dq[LOOP_MAXB], tau[LOOP_MAXB], eft[LOOP_MAXB], efx[LOOP_MAXB], efy[LOOP_MAXB];
        for (int i = 0; i < NB; i++) { q[i] = (R)q_in[w * NB + i]; dq[i] = (R)dq_in[w * NB + i]; tau[i] = tau_in ? (R)tau_in[w * NB + i] : (R)0; eft[i] = efx[i] = efy[i] = 0; }
        ContactSink<R> sink;
        sink.count = count; sink.body = body; sink.data = data; sink.maxc = maxc;
        if (fext) {
            R cs[LOOP_MAXB], sn[LOOP_MAXB], px[LOOP_MAXB], py[LOOP_MAXB];
            fk_positions_loop<R>(M, q, cs, sn, px, py);
            for (int k = 0; k < M.nbd; k++) {
                const double* f = fext + ((size_t)w * M.nbd + k) * 3;
                const R fx = (R)(M.e1[0] * f[0] + M.e1[1] * f[1] + M.e1[2] * f[2]);
                const R fy = (R)(M.e2[0] * f[0] + M.e2[1] * f[1] + M.e2[2] * f[2]);
                const int g = M.dgroup[k];
                const R ox = cs[g] * M.dox[k] - sn[g] * M.doy[k], oy = sn[g] * M.dox[k] + cs[g] * M.doy[k];
                eft[g] += ox * fy - oy * fx; efx[g] += fx; efy[g] += fy;
            }
            substep_loop<R>(M, q, dq, tau, true, eft, efx, efy, false, (R)0, (R)0, lcp_mode, pgs_iters, &sink, w, wp, (size_t)n);
        } else {
            substep_loop<R>(M, q, dq, tau, false, eft, efx, efy, false, (R)0, (R)0, lcp_mode, pgs_iters, &sink, w, wp, (size_t)n);
        }
        for (int i = 0; i < NB; i++) { q_out[w * NB + i] = (double)q[i]; dq_out[w * NB + i] = (double)dq[i]; }
    }
}

// the lane-cooperative kernel (planar_coop.cuh) under the one-warp SIMT emulator of simt.h
template <class T, typename R>
static void run_substep_coop(const PModel<R>& M, int n, const double* q_in, const double* dq_in, const double* tau_in,
                             int lcp_mode, int pgs_iters, double* q_out, double* dq_out, int32_t* count, int32_t* body,
                             float* data, int maxc) {
    constexpr int NB = T::NB;
    std::vector<R> qs((size_t)NB * n), dqs((size_t)NB * n), tau((size_t)NB * n);
    for (int w = 0; w < n; w++)
        for (int i = 0; i < NB; i++) {
            qs[(size_t)i * n + w] = (R)q_in[w * NB + i]; dqs[(size_t)i * n + w] = (R)dq_in[w * NB + i];
            tau[(size_t)w * NB + i] = tau_in ? (R)tau_in[w * NB + i] : (R)0;
        }
    ContactSink<R> sink;
    sink.count = count; sink.body = body; sink.data = data; sink.maxc = maxc;
    const int grid = (n + Coop<T>::WPW - 1) / Coop<T>::WPW;
    CoopLane<T, R> tab[Coop<T>::G];
    coop_build_table<T, R>(M, nullptr, tab);
    simt::launch(grid, coop_shared_bytes<T, R>(1, 0), [&] {
        k_substep_coop<T, R>(M, tab, n, qs.data(), dqs.data(), tau.data(), lcp_mode, pgs_iters, sink);
    });
    for (int w = 0; w < n; w++)
        for (int i = 0; i < NB; i++) { q_out[w * NB + i] = (double)qs[(size_t)i * n + w]; dq_out[w * NB + i] = (double)dqs[(size_t)i * n + w]; }
}

// the quad form of the per-thread kernels (substep<..., G = 4>: four lanes per world) under the SIMT emulator;
// mirrors csrc/kernels.cuh::k_substep_quad
template <class T, typename R, int G>
static void run_substep_quad(const PModel<R>& M, int n, const double* q_in, const double* dq_in, const double* tau_in,
                             int lcp_mode, int pgs_iters, double* q_out, double* dq_out, int32_t* count, int32_t* body,
                             float* data, int maxc) {
    constexpr int NB = T::NB;
    ContactSink<R> sink;
    sink.count = count; sink.body = body; sink.data = data; sink.maxc = maxc;
    constexpr int WPW = 32 / G;
    const int grid = (n + WPW - 1) / WPW;
    simt::launch(grid, 0, [&] {
        const int lane = threadIdx.x & 31, gi = lane / G, l = lane % G;
        const int w = blockIdx.x * WPW + gi;
        const bool active = w < n;
        const int wr = active ? w : 0;
        R q[NB], dq[NB], tau[NB], zero[NB];
        for (int i = 0; i < NB; i++) {
            q[i] = active ? (R)q_in[wr * NB + i] : M.qinit[i];
            dq[i] = active ? (R)dq_in[wr * NB + i] : (R)0;
            tau[i] = (active && tau_in) ? (R)tau_in[wr * NB + i] : (R)0;
            zero[i] = 0;
        }
        uint64_t hint = ~(uint64_t)0;
        substep<T, R, false, false, G>(M, q, dq, tau, zero, zero, zero, (R)0, (R)0, lcp_mode, pgs_iters, (active && l == 0) ? &sink : nullptr, wr, hint);
        // every lane of the group must hold the same bits: lane 1 checks against lane 0 through the outputs
        if (active && l == 0) for (int i = 0; i < NB; i++) { q_out[w * NB + i] = (double)q[i]; dq_out[w * NB + i] = (double)dq[i]; }
        __syncwarp();
        if (active && l == G - 1) for (int i = 0; i < NB; i++) if (q_out[w * NB + i] != (double)q[i] || dq_out[w * NB + i] != (double)dq[i]) { q_out[w * NB + i] = NAN; }
    });
}

long g_emu_counters[8];
long g_emu_hist[32];
extern "C" void emu_hist(long* out, int reset) { for (int i = 0; i < 32; i++) { out[i] = g_emu_hist[i]; if (reset) g_emu_hist[i] = 0; } }
extern "C" void emu_counters(long* out, int reset) { for (int i = 0; i < 8; i++) { out[i] = g_emu_counters[i]; if (reset) g_emu_counters[i] = 0; } }

static std::string g_err;

// Num<float>::sincos_ / Num<double>::sincos_ of the kernel source, element-wise (the branch-free fp32 form's accuracy test)
extern "C" void emu_sincos(int n, const float* x, float* s, float* c) {
    for (int i = 0; i < n; i++) Num<float>::sincos_(x[i], &s[i], &c[i]);
}

extern "C" const char* emu_last_error() { return g_err.c_str(); }

// the loop kernel's sub-step with per-world bodynode masses / friction coefficients ([n][n_bodies], either may be null):
// the same lower::body_param_table dartb_set_body_params uploads
extern "C" int emu_substep_body_params(const dartb_model_t* model, const dartb_task_t* task, int f64, int n, const double* q,
                                       const double* dq, const double* tau, const double* mass, const double* mu, int lcp_mode,
                                       int pgs_iters, double* q_out, double* dq_out, int32_t* count, int32_t* body, float* data,
                                       int maxc) {
    lower::Result res;
    std::string why = lower::lower_model(*model, *task, res);
    if (!why.empty()) { g_err = why; return 1; }
    std::vector<double> tab;
    why = lower::body_param_table(*model, *task, n, mass, mu, res.signature, res.m.nb, res.m.ns, tab);
    if (!why.empty()) { g_err = why; return 1; }
    if (f64) run_substep_loop<double>(res.m, n, q, dq, tau, nullptr, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc, tab.data());
    else {
        PModel<float> mf;
        lower::convert(res.m, mf);
        std::vector<float> tf(tab.begin(), tab.end());
        run_substep_loop<float>(mf, n, q, dq, tau, nullptr, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc, tf.data());
    }
    return 0;
}

// returns 0 ok; arrays are [n, nd] row-major doubles (converted to the kernel precision inside)
extern "C" int emu_substep(const dartb_model_t* model, const dartb_task_t* task, int f64, int n, const double* q,
                           const double* dq, const double* tau, const double* fext, int lcp_mode, int pgs_iters,
                           double* q_out, double* dq_out, int32_t* count, int32_t* body, float* data, int maxc, int variant,
                           uint64_t* hints) {
    lower::Result res;
    std::string why = lower::lower_model(*model, *task, res);
    if (!why.empty()) { g_err = why; return 1; }
    PModel<float> mf;
    lower::convert(res.m, mf);
    if (variant == 1) {
        if (f64) run_substep_loop<double>(res.m, n, q, dq, tau, fext, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc);
        else run_substep_loop<float>(mf, n, q, dq, tau, fext, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc);
        return 0;
    }
#define RUNC(T)                                                                                                         \
    if (variant == 2 && res.signature == T::sig) {                                                                       \
        if (fext) { g_err = "the cooperative kernel takes no external forces"; return 1; }                              \
        if (f64) run_substep_coop<T, double>(res.m, n, q, dq, tau, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc); \
        else run_substep_coop<T, float>(mf, n, q, dq, tau, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc);          \
        return 0;                                                                                                        \
    }
    RUNC(TopoHopper) RUNC(TopoWalker) RUNC(TopoCheetah) RUNC(TopoSnake)
#define RUNQ(T)                                                                                                         \
    if (variant >= 3 && variant <= 5 && res.signature == T::sig) {                                                       \
        if (fext) { g_err = "the group kernels take no external forces"; return 1; }                                    \
        if (variant == 3) { if (f64) run_substep_quad<T, double, 4>(res.m, n, q, dq, tau, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc); \
                            else run_substep_quad<T, float, 4>(mf, n, q, dq, tau, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc); }        \
        if (variant == 4) { if (f64) run_substep_quad<T, double, 2>(res.m, n, q, dq, tau, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc); \
                            else run_substep_quad<T, float, 2>(mf, n, q, dq, tau, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc); }        \
        if (variant == 5) { if (f64) run_substep_quad<T, double, 8>(res.m, n, q, dq, tau, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc); \
                            else run_substep_quad<T, float, 8>(mf, n, q, dq, tau, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc); }        \
        return 0;                                                                                                        \
    }
    RUNQ(TopoHopper) RUNQ(TopoWalker) RUNQ(TopoCheetah) RUNQ(TopoSnake)
#define RUN(T)                                                                                                          \
    if (res.signature == T::sig) {                                                                                       \
        if (f64) run_substep<T, double>(res.m, n, q, dq, tau, fext, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc, hints); \
        else run_substep<T, float>(mf, n, q, dq, tau, fext, lcp_mode, pgs_iters, q_out, dq_out, count, body, data, maxc, hints);          \
        return 0;                                                                                                        \
    }
    RUN(TopoHopper) RUN(TopoWalker) RUN(TopoCheetah) RUN(TopoSnake)
    g_err = "no topology for " + res.signature;
    return 1;
}

// task layer of the contact-free kinds (task_kinds.cuh) on ONE state per world: obs [n, n_obs], reward [n], done [n];
// with do_reset != 0 the state is first replaced by reset_model()'s draw (q, dq, aux are in/out then)
template <typename R>
static void run_task_kind(const PModel<R>& M, const PTask<R>& K, int n, double* q, double* dq, double* aux, const double* a2,
                          int do_reset, uint64_t seed, float* obs, double* rew, int32_t* done) {
    const int nb = M.nb;
    for (int w = 0; w < n; w++) {
        R qq[LOOP_MAXB], dd[LOOP_MAXB], tg[3];
        for (int i = 0; i < nb; i++) { qq[i] = (R)q[w * nb + i]; dd[i] = (R)dq[w * nb + i]; }
        for (int c = 0; c < 3; c++) tg[c] = (R)aux[w * 3 + c];
        if (do_reset) {
            reset_state_kind<R>(M, K, seed, w, 0, qq, dd, tg);
            for (int i = 0; i < nb; i++) { q[w * nb + i] = (double)qq[i]; dq[w * nb + i] = (double)dd[i]; }
            for (int c = 0; c < 3; c++) aux[w * 3 + c] = (double)tg[c];
        }
        R r; bool d;
        task_kind_eval<R>(M, K, qq, dd, tg, (R)a2[w], r, d);
        write_obs_kind<R>(M, K, qq, dd, tg, obs + (size_t)w * K.n_obs);
        rew[w] = (double)r; done[w] = d ? 1 : 0;
    }
}
extern "C" int emu_task_kind(const dartb_model_t* model, const dartb_task_t* task, int f64, int n, double* q, double* dq,
                             double* aux, const double* a2, int do_reset, uint64_t seed, float* obs, double* rew, int32_t* done) {
    lower::Result res;
    std::string why = lower::lower_model(*model, *task, res);
    if (!why.empty()) { g_err = why; return 1; }
    if (res.t.kind == DARTB_TASK_LOCOMOTION) { g_err = "not a contact-free task kind"; return 1; }
    if (f64) run_task_kind<double>(res.m, res.t, n, q, dq, aux, a2, do_reset, seed, obs, rew, done);
    else {
        PModel<float> mf; PTask<float> tf;
        lower::convert(res.m, mf); lower::convert(res.t, tf);
        run_task_kind<float>(mf, tf, n, q, dq, aux, a2, do_reset, seed, obs, rew, done);
    }
    return 0;
}

extern "C" float emu_reset_uniform(uint64_t seed, int64_t world, uint32_t episode, int i) {
    return reset_uniform(seed, world, episode, i);
}

// standalone LCP entry for tests: mode 0 = lcp_exact dispatch, 1 = lcp_small<8>, 2 = lcp_bpp_local,
// 3 = lcp_dantzig, 4 = lcp_small<4>, 5 = lcp_small<6>, 6/7/8 = lcp_ppt<4/6/8>, 9 = lcp_ppt_loop.  Returns 0 ok, 1 solver reported failure.
template <typename R>
static int run_lcp(int n, const double* A, const double* b, const double* lo, const double* hi, const int* fidx, int mode,
                   double* x) {
    constexpr int NR = 28;
    R a[NR * NR], bb[NR], l[NR], h[NR], xx[NR];
    int fi[NR];
    for (int i = 0; i < n; i++) {
        bb[i] = (R)b[i]; l[i] = (R)lo[i]; h[i] = (R)hi[i]; fi[i] = fidx[i]; xx[i] = 0;
        for (int j = 0; j < n; j++) a[i * n + j] = (R)A[i * n + j];
    }
    bool ok = true;
    if (mode == 0) lcp_exact<R, NR>(n, a, xx, bb, l, h, fi);
    else if (mode == 1) ok = lcp_small<R, 8>(n, a, xx, bb, l, h, fi);
    else if (mode == 2) ok = lcp_bpp_local<R, NR>(n, a, xx, bb, l, h, fi);
    else if (mode == 3) lcp_dantzig<R, NR>(n, a, xx, bb, l, h, fi);
    else if (mode == 4) ok = lcp_small<R, 4>(n, a, xx, bb, l, h, fi);
    else if (mode == 5) ok = lcp_small<R, 6>(n, a, xx, bb, l, h, fi);
    else if (mode == 6) ok = lcp_ppt<R, 4>(n, a, xx, bb, l, h, fi);
    else if (mode == 7) ok = lcp_ppt<R, 6>(n, a, xx, bb, l, h, fi);
    else if (mode == 8) ok = lcp_ppt<R, 8>(n, a, xx, bb, l, h, fi);
    else if (mode == 9) ok = lcp_ppt_loop<R, NR>(n, a, xx, bb, l, h, fi);
    for (int i = 0; i < n; i++) x[i] = (double)xx[i];
    return ok ? 0 : 1;
}
extern "C" int emu_lcp(int f64, int n, const double* A, const double* b, const double* lo, const double* hi, const int* fidx,
                       int mode, double* x) {
    if (n > 28) return 2;
    return f64 ? run_lcp<double>(n, A, b, lo, hi, fidx, mode, x) : run_lcp<float>(n, A, b, lo, hi, fidx, mode, x);
}
