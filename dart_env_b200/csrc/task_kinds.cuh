// task_kinds.cuh — the contact-free envs' task layers as device functions of the topology-generic
// kernels: cart_pole.py, cartpole_swingup.py, inverted_double_pendulum.py, reacher2d.py of the reference, selected by
// dartb_task_t.kind.  Included by kernels.cuh (k_env_step_loop / k_reset_loop) and, for the no-GPU regression tests,
// by tools/host_emu.
#pragma once
#include "../../include/dartb.h"
#include "planar_kernels.cuh"
#include "planar_loop.cuh"

// ---- SURVEY §8f.1 task kinds (cart_pole.py, cartpole_swingup.py, inverted_double_pendulum.py, reacher2d.py) ----
template <typename R> struct NumX;
template <> struct NumX<float> {
    static DEVI float log_(float x) { return logf(x); }
    static DEVI float cos_(float x) { return cosf(x); }
};
template <> struct NumX<double> {
    static DEVI double log_(double x) { return log(x); }
    static DEVI double cos_(double x) { return cos(x); }
};
// world position (plane coordinates X, Y) of probe k
template <typename R>
DEVI void probe_point_loop(const PModel<R>& M, const PTask<R>& K, const R* q, int k, R& X, R& Y) {
    R cs[LOOP_MAXB], sn[LOOP_MAXB], px[LOOP_MAXB], py[LOOP_MAXB];
    fk_positions_loop<R>(M, q, cs, sn, px, py);
    const int i = K.probe_body[k];
    X = px[i] + cs[i] * K.probe_x[k] - sn[i] * K.probe_y[k];
    Y = py[i] + sn[i] * K.probe_x[k] + cs[i] * K.probe_y[k];
}
// reacher2d.py:31: bodynodes[-1].com() - target, world x / y / z
template <typename R>
DEVI void reacher_vec(const PModel<R>& M, const PTask<R>& K, const R* q, const R* tg, R (&vec)[3]) {
    R X, Y;
    probe_point_loop<R>(M, K, q, 0, X, Y);
#pragma unroll
    for (int c = 0; c < 3; c++) vec[c] = M.e1[c] * X + M.e2[c] * Y + M.en[c] * K.probe_n[0] - tg[c];
}
template <typename R>
DEVI void write_obs_kind(const PModel<R>& M, const PTask<R>& K, const R* q, const R* dq, const R* tg, float* so) {
    const int nb = M.nb;
    if (K.kind == DARTB_TASK_CARTPOLE || K.kind == DARTB_TASK_SWINGUP) {
        for (int i = 0; i < nb; i++) { so[i] = (float)q[i]; so[nb + i] = (float)dq[i]; }
    } else if (K.kind == DARTB_TASK_DOUBLE_PENDULUM) {   // [q[:1], sin(q[1:]), cos(q[1:]), dq]
        so[0] = (float)q[0];
        for (int i = 1; i < nb; i++) { R s, c; Num<R>::sincos_(q[i], &s, &c); so[i] = (float)s; so[nb - 1 + i] = (float)c; }
        for (int i = 0; i < nb; i++) so[2 * nb - 1 + i] = (float)dq[i];
    } else {                                             // [cos(q), sin(q), target[0], target[2], dq, tip - target]
        for (int i = 0; i < nb; i++) { R s, c; Num<R>::sincos_(q[i], &s, &c); so[i] = (float)c; so[nb + i] = (float)s; }
        so[2 * nb] = (float)tg[0]; so[2 * nb + 1] = (float)tg[2];
        for (int i = 0; i < nb; i++) so[2 * nb + 2 + i] = (float)dq[i];
        R vec[3];
        reacher_vec<R>(M, K, q, tg, vec);
        for (int c = 0; c < 3; c++) so[3 * nb + 2 + c] = (float)vec[c];
    }
}
// reward and done of one step; a2 = sum of squares of the RAW action
template <typename R>
DEVI void task_kind_eval(const PModel<R>& M, const PTask<R>& K, const R* q, const R* dq, const R* tg, R a2, R& r, bool& done) {
    const int nb = M.nb;
    if (K.kind == DARTB_TASK_CARTPOLE) {
        bool fin = true;
        for (int i = 0; i < nb; i++) fin = fin && (Num<R>::abs_(q[i]) < Num<R>::inf()) && (Num<R>::abs_(dq[i]) < Num<R>::inf());
        r = 1;
        done = !(fin && Num<R>::abs_(q[1]) <= (R)0.2);
    } else if (K.kind == DARTB_TASK_SWINGUP) {
        r = (R)6.0 - Num<R>::abs_(q[1]) - (R)0.01 * a2 - (R)0.01 * Num<R>::abs_(q[0]);
        done = Num<R>::abs_(q[1]) > (R)(8 * 3.14159265358979323846) || Num<R>::abs_(dq[1]) > (R)25 || Num<R>::abs_(q[0]) > (R)5;
    } else if (K.kind == DARTB_TASK_DOUBLE_PENDULUM) {
        R X0, Y0, X1, Y1;
        probe_point_loop<R>(M, K, q, 0, X0, Y0);
        probe_point_loop<R>(M, K, q, 1, X1, Y1);
        // to_world()[1] of 'cart' and 'weight': world y = e1.y X + e2.y Y + en.y n
        const R base = M.e1[1] * X0 + M.e2[1] * Y0 + M.en[1] * K.probe_n[0];
        const R raw = M.e1[1] * X1 + M.e2[1] * Y1 + M.en[1] * K.probe_n[1];
        const R height = (R)2.0 * (raw - base - (R)0.02) / (R)0.6;
        const R dist_penalty = (R)0.01 * q[0] * q[0] + (height - (R)2.) * (height - (R)2.);
        const R vel_penalty = (R)1e-3 * dq[1] * dq[1] + (R)5e-3 * dq[2] * dq[2];
        r = (R)10. - dist_penalty - vel_penalty;
        done = height <= (R)1;
    } else {
        R vec[3];
        reacher_vec<R>(M, K, q, tg, vec);
        r = -Num<R>::sqrt_(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]) - a2;
        done = false;
    }
}
// reset_model() of the contact-free kinds: counter-based draws (Philox keyed by seed, global world id, episode)
template <typename R>
DEVI void reset_state_kind(const PModel<R>& M, const PTask<R>& K, uint64_t seed, int64_t gw, uint32_t ep, R* q, R* dq, R* tg) {
    const int nb = M.nb;
    for (int i = 0; i < nb; i++) q[i] = M.qinit[i] + (R)reset_uniform(seed, gw, ep, i) * K.reset_noise;
    if (K.kind == DARTB_TASK_DOUBLE_PENDULUM) {          // init_qvel + randn * 0.1 (Box-Muller)
        for (int i = 0; i < nb; i++) {
            const R u1 = (R)0.5 * ((R)reset_uniform(seed, gw, ep, nb + 2 * i) + (R)1) + (R)2.9802322e-8;   // (0, 1]
            const R u2 = (R)0.5 * ((R)reset_uniform(seed, gw, ep, nb + 2 * i + 1) + (R)1);
            dq[i] = M.dqinit[i] + K.noise_dq * Num<R>::sqrt_((R)-2 * NumX<R>::log_(u1)) * NumX<R>::cos_((R)6.283185307179586 * u2);
        }
    } else {
        for (int i = 0; i < nb; i++) dq[i] = M.dqinit[i] + (R)reset_uniform(seed, gw, ep, nb + i) * K.noise_dq;
    }
    if (K.kind == DARTB_TASK_SWINGUP)                    // qpos[1] += pi if U(0,1) > 0.5 else -pi
        q[1] += reset_uniform(seed, gw, ep, 2 * nb) > 0.0f ? (R)3.14159265358979323846 : (R)-3.14159265358979323846;
    if (K.kind == DARTB_TASK_REACHER2D) {                // target: U(+-0.2)^3 with [1] = 0, redrawn until |target| < 0.2
        R tx = 0, tz = 0;
        for (int k = 0; k < 64; k++) {
            tx = (R)0.2 * (R)reset_uniform(seed, gw, ep, 2 * nb + 2 * k);
            tz = (R)0.2 * (R)reset_uniform(seed, gw, ep, 2 * nb + 2 * k + 1);
            if (tx * tx + tz * tz < (R)0.04) break;
        }
        tg[0] = tx; tg[1] = (R)0.01; tg[2] = tz;
    }
}

