// dartb.cu — kernels + C-ABI (include/dartb.h) of the B200-native batched DART stepper.
//
// One thread = one world.  State is SoA [nd][n_worlds] in HBM (coalesced loads/stores);
// the AoS boundary arrays (action [n,n_act], obs [n,n_obs], tau [n,nd]) are staged through
// shared memory per warp so global accesses stay coalesced.  One launch per env.step():
// action -> frame_skip x DART step -> obs/reward/done -> masked auto-reset, nothing else
// touches HBM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a (see build.py).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/dartb.h"
#include "lower.h"
#include "planar_kernels.cuh"
#include "planar_loop.cuh"

// ------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));   \
    } while (0)

// ------------------------------------------------------------------------ kernel arguments
template <typename R>
struct StepArgs {
    int n;
    R* q;                // [nd][n]
    R* dq;               // [nd][n]
    uint32_t* episode;   // [n] reset counter (Philox stream position)
    int32_t* elapsed;    // [n] env steps since reset (TimeLimit)
    uint8_t* truncated;  // [n]
    const float* action; // [n, n_act]
    float* obs;          // [n, n_obs]
    float* reward;       // [n]
    uint8_t* done;       // [n]
    const uint8_t* mask; // reset mask (k_reset) or null
    int auto_reset, lcp_mode, pgs_iters, max_episode_steps;
    uint64_t seed;
    int64_t world_offset;
    ContactSink<R> sink;
};

template <class T, typename R>
DEVI void write_obs(const PModel<R>& M, const PTask<R>& K, const R (&q)[T::NB], const R (&dq)[T::NB], float* so) {
    constexpr int NB = T::NB;
    if (K.obs_mode == DARTB_OBS_HEIGHT_Q2_DQ) {
        R cs[NB], sn[NB], px[NB], py[NB];
        fk_positions<T, R>(M, q, cs, sn, px, py);
        R h = 0;
        static_for<0, NB>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if (K.height_body == i) {
                const R X = px[i] + cs[i] * K.hcx - sn[i] * K.hcy, Y = py[i] + sn[i] * K.hcx + cs[i] * K.hcy;
                h = K.wy1 * X + K.wy2 * Y + K.wy0;
            }
        });
        so[0] = (float)h;
    } else {
        so[0] = (float)q[1];
    }
    static_for<2, NB>([&](auto ic) { constexpr int i = decltype(ic)::value; so[i - 1] = (float)q[i]; });
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        R v = dq[i];
        if (K.dq_clip > 0) v = v > K.dq_clip ? K.dq_clip : (v < -K.dq_clip ? -K.dq_clip : v);
        so[NB - 1 + i] = (float)v;
    });
}

template <class T, typename R>
DEVI R body_height(const PModel<R>& M, const PTask<R>& K, const R (&q)[T::NB]) {
    constexpr int NB = T::NB;
    R cs[NB], sn[NB], px[NB], py[NB];
    fk_positions<T, R>(M, q, cs, sn, px, py);
    R h = 0;
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        if (K.height_body == i) {
            const R X = px[i] + cs[i] * K.hcx - sn[i] * K.hcy, Y = py[i] + sn[i] * K.hcx + cs[i] * K.hcy;
            h = K.wy1 * X + K.wy2 * Y + K.wy0;
        }
    });
    return h;
}

// reset_model(): q0 + U(+-noise), dq0 + U(+-noise) in fp32 arithmetic (bit-identical to the oracle)
template <class T, typename R>
DEVI void reset_state(const PModel<R>& M, const PTask<R>& K, uint64_t seed, int64_t gw, uint32_t ep, R (&q)[T::NB],
                      R (&dq)[T::NB]) {
    constexpr int NB = T::NB;
    const float noise = (float)K.reset_noise;
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        const float a = __fmul_rn(reset_uniform(seed, gw, ep, i), noise);
        const float b = __fmul_rn(reset_uniform(seed, gw, ep, NB + i), noise);
        q[i] = (R)__fadd_rn((float)M.qinit[i], a);
        dq[i] = (R)__fadd_rn((float)M.dqinit[i], b);
    });
}

// ------------------------------------------------------------------------ env.step() kernel
template <class T, typename R>
__global__ void __launch_bounds__(256)
k_env_step(const __grid_constant__ PModel<R> M, const __grid_constant__ PTask<R> K, const __grid_constant__ StepArgs<R> a) {
    constexpr int NB = T::NB;
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int wb = w - lane;                          // first world of this warp
    const int cnt = min(32, a.n - wb);                // worlds this warp owns (<= 0: idle warp)
    const bool active = w < a.n;
    const int stage = K.n_obs > K.n_act ? K.n_obs : K.n_act;
    float* sw = smem + warp * 32 * stage;

    // coalesced action load -> smem [lane][n_act]
    if (cnt > 0) for (int k = lane; k < cnt * K.n_act; k += 32) sw[k] = a.action[(size_t)wb * K.n_act + k];
    __syncwarp();

    R q[NB], dq[NB], tau[NB], zero[NB];
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        q[i] = active ? a.q[(size_t)i * a.n + w] : M.qinit[i];
        dq[i] = active ? a.dq[(size_t)i * a.n + w] : (R)0;
        zero[i] = 0;
    });
    // advance(): clamp, scale, scatter (hopper.py:24-32); control cost uses the RAW action
    R a2 = 0;
    if (active) for (int j = 0; j < K.n_act; j++) { const R v = (R)sw[lane * K.n_act + j]; a2 += v * v; }
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        R t = 0;
        if (active && K.dof_act[i] >= 0) {
            R v = (R)sw[lane * K.n_act + K.dof_act[i]];
            v = v > K.dof_hi[i] ? K.dof_hi[i] : v;
            v = v < K.dof_lo[i] ? K.dof_lo[i] : v;
            t = v * K.dof_scale[i];
        }
        tau[i] = t;
    });
    __syncwarp();

    const R posbefore = q[0];
    const ContactSink<R>* sink = &a.sink;
    for (int f = 0; f < K.frame_skip; f++) {
        const ContactSink<R>* sk = (active && f == K.frame_skip - 1) ? sink : nullptr;
        if (K.fluid_force)
            substep<T, R, false, true>(M, q, dq, tau, zero, zero, zero, K.fluid_offset, K.fluid_coef, a.lcp_mode, a.pgs_iters, sk, w);
        else
            substep<T, R, false, false>(M, q, dq, tau, zero, zero, zero, (R)0, (R)0, a.lcp_mode, a.pgs_iters, sk, w);
    }
    // reward / done (hopper.py:36-65, walker2d.py:22-65, half_cheetah.py:40-77, snake_7link.py:68-87)
    const R ang = q[2];
    R r = (q[0] - posbefore) * K.inv_dt_env * K.vel_weight;
    r += K.alive_bonus;
    r -= K.ctrl_cost * a2;
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        if (K.limit_pen_dof == i) {
            R pen = 0;
            if ((M.qlo[i] - q[i]) > -K.limit_pen_margin) pen += (R)1.5;
            if ((M.qhi[i] - q[i]) < K.limit_pen_margin) pen += (R)1.5;
            r -= K.limit_pen_weight * pen;
        }
    });
    r -= K.dev_cost * Num<R>::abs_(ang);
    bool ok = true;
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        // isfinite and |.| < bound in one comparison (NaN / inf fail it)
        if (i >= 2 && !(Num<R>::abs_(q[i]) < K.state_bound)) ok = false;
        if (i < 2 && !(Num<R>::abs_(q[i]) < Num<R>::inf())) ok = false;
        if (!(Num<R>::abs_(dq[i]) < K.state_bound)) ok = false;
    });
    if (K.zero_reward_on_blowup && !ok) r = 0;
    if (K.height_body >= 0) {
        const R h = body_height<T, R>(M, K, q);
        ok = ok && (h > K.height_lo) && (h < K.height_hi);
    }
    ok = ok && (Num<R>::abs_(ang) < K.ang_max);
    bool done = !ok;
    bool trunc = false;
    if (active && a.max_episode_steps > 0) {
        const int el = a.elapsed[w] + 1;
        if (el >= a.max_episode_steps) { trunc = !done; done = true; }
        a.elapsed[w] = (done && a.auto_reset) ? 0 : el;
    }
    if (active && done && a.auto_reset) {
        const uint32_t ep = a.episode[w];
        reset_state<T, R>(M, K, a.seed, a.world_offset + w, ep, q, dq);
        a.episode[w] = ep + 1;
    }
    // obs (of the reset state for auto-reset worlds: gym/vector/sync_vector_env.py:76-79)
    if (active) write_obs<T, R>(M, K, q, dq, sw + lane * K.n_obs);
    __syncwarp();
    if (cnt > 0) for (int k = lane; k < cnt * K.n_obs; k += 32) a.obs[(size_t)wb * K.n_obs + k] = sw[k];
    if (active) {
        static_for<0, NB>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            a.q[(size_t)i * a.n + w] = q[i];
            a.dq[(size_t)i * a.n + w] = dq[i];
        });
        a.reward[w] = (float)r;
        a.done[w] = done ? 1 : 0;
        if (a.truncated) a.truncated[w] = trunc ? 1 : 0;
    }
}

// ------------------------------------------------------------------------ reset kernel
template <class T, typename R>
__global__ void __launch_bounds__(256)
k_reset(const __grid_constant__ PModel<R> M, const __grid_constant__ PTask<R> K, const __grid_constant__ StepArgs<R> a) {
    constexpr int NB = T::NB;
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int wb = w - lane, cnt = min(32, a.n - wb);
    const bool active = w < a.n;
    float* sw = smem + warp * 32 * K.n_obs;
    R q[NB], dq[NB];
    const bool doit = active && (a.mask == nullptr || a.mask[w]);
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        q[i] = active ? a.q[(size_t)i * a.n + w] : M.qinit[i];
        dq[i] = active ? a.dq[(size_t)i * a.n + w] : (R)0;
    });
    if (doit) {
        const uint32_t ep = a.episode[w];
        reset_state<T, R>(M, K, a.seed, a.world_offset + w, ep, q, dq);
        a.episode[w] = ep + 1;
        a.elapsed[w] = 0;
        static_for<0, NB>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            a.q[(size_t)i * a.n + w] = q[i];
            a.dq[(size_t)i * a.n + w] = dq[i];
        });
        if (a.sink.count) a.sink.count[w] = 0;
    }
    if (a.obs) {
        if (active) write_obs<T, R>(M, K, q, dq, sw + lane * K.n_obs);
        __syncwarp();
        if (cnt > 0) for (int k = lane; k < cnt * K.n_obs; k += 32) a.obs[(size_t)wb * K.n_obs + k] = sw[k];
    }
}

// ------------------------------------------------------------------------ single DART step kernel
// exactly `skel.set_forces(tau); world.step()` (dart_env.py:174-175) with optional ext forces
template <class T, typename R>
__global__ void __launch_bounds__(256)
k_substep(const __grid_constant__ PModel<R> M, int n, R* qs, R* dqs, const R* tau_in /*[n,nd]*/, const R* fext /*[n,nbd,3]*/,
          int lcp_mode, int pgs_iters, const __grid_constant__ ContactSink<R> sink) {
    constexpr int NB = T::NB;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    R q[NB], dq[NB], tau[NB], eft[NB], efx[NB], efy[NB];
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        q[i] = qs[(size_t)i * n + w];
        dq[i] = dqs[(size_t)i * n + w];
        tau[i] = tau_in ? tau_in[(size_t)w * NB + i] : (R)0;
        eft[i] = 0; efx[i] = 0; efy[i] = 0;
    });
    if (fext) {
        R cs[NB], sn[NB], px[NB], py[NB];
        fk_positions<T, R>(M, q, cs, sn, px, py);
        for (int k = 0; k < M.nbd; k++) {
            const R* f = fext + ((size_t)w * M.nbd + k) * 3;
            const R fx = M.e1[0] * f[0] + M.e1[1] * f[1] + M.e1[2] * f[2];
            const R fy = M.e2[0] * f[0] + M.e2[1] * f[1] + M.e2[2] * f[2];
            const int g = M.dgroup[k];
            static_for<0, NB>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                if (g == i) {
                    const R ox = cs[i] * M.dox[k] - sn[i] * M.doy[k], oy = sn[i] * M.dox[k] + cs[i] * M.doy[k];
                    eft[i] += ox * fy - oy * fx; efx[i] += fx; efy[i] += fy;
                }
            });
        }
        substep<T, R, true, false>(M, q, dq, tau, eft, efx, efy, (R)0, (R)0, lcp_mode, pgs_iters, &sink, w);
    } else {
        substep<T, R, false, false>(M, q, dq, tau, eft, efx, efy, (R)0, (R)0, lcp_mode, pgs_iters, &sink, w);
    }
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        qs[(size_t)i * n + w] = q[i];
        dqs[(size_t)i * n + w] = dq[i];
    });
}


// ======================================================================== loop (generic) variant
// Same kernels built on planar_loop.cuh: runtime topology, small instruction footprint.
template <typename R>
DEVI R body_height_loop(const PModel<R>& M, const PTask<R>& K, const R* q) {
    R cs[LOOP_MAXB], sn[LOOP_MAXB], px[LOOP_MAXB], py[LOOP_MAXB];
    fk_positions_loop<R>(M, q, cs, sn, px, py);
    const int i = K.height_body;
    const R X = px[i] + cs[i] * K.hcx - sn[i] * K.hcy, Y = py[i] + sn[i] * K.hcx + cs[i] * K.hcy;
    return K.wy1 * X + K.wy2 * Y + K.wy0;
}
template <typename R>
DEVI void write_obs_loop(const PModel<R>& M, const PTask<R>& K, const R* q, const R* dq, float* so) {
    const int nb = M.nb;
    so[0] = (K.obs_mode == DARTB_OBS_HEIGHT_Q2_DQ) ? (float)body_height_loop<R>(M, K, q) : (float)q[1];
    for (int i = 2; i < nb; i++) so[i - 1] = (float)q[i];
    for (int i = 0; i < nb; i++) {
        R v = dq[i];
        if (K.dq_clip > 0) v = v > K.dq_clip ? K.dq_clip : (v < -K.dq_clip ? -K.dq_clip : v);
        so[nb - 1 + i] = (float)v;
    }
}
template <typename R>
DEVI void reset_state_loop(const PModel<R>& M, const PTask<R>& K, uint64_t seed, int64_t gw, uint32_t ep, R* q, R* dq) {
    const int nb = M.nb;
    const float noise = (float)K.reset_noise;
    for (int i = 0; i < nb; i++) {
        const float a = __fmul_rn(reset_uniform(seed, gw, ep, i), noise);
        const float b = __fmul_rn(reset_uniform(seed, gw, ep, nb + i), noise);
        q[i] = (R)__fadd_rn((float)M.qinit[i], a);
        dq[i] = (R)__fadd_rn((float)M.dqinit[i], b);
    }
}

template <typename R>
__global__ void __launch_bounds__(256)
k_env_step_loop(const __grid_constant__ PModel<R> M, const __grid_constant__ PTask<R> K, const __grid_constant__ StepArgs<R> a) {
    extern __shared__ float smem[];
    const int nb = M.nb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int wb = w - lane, cnt = min(32, a.n - wb);
    const bool active = w < a.n;
    const int stage = K.n_obs > K.n_act ? K.n_obs : K.n_act;
    float* sw = smem + warp * 32 * stage;
    if (cnt > 0) for (int k = lane; k < cnt * K.n_act; k += 32) sw[k] = a.action[(size_t)wb * K.n_act + k];
    __syncwarp();
    R q[LOOP_MAXB], dq[LOOP_MAXB], tau[LOOP_MAXB];
    R a2 = 0;
    if (active) for (int j = 0; j < K.n_act; j++) { const R v = (R)sw[lane * K.n_act + j]; a2 += v * v; }
    for (int i = 0; i < nb; i++) {
        q[i] = active ? a.q[(size_t)i * a.n + w] : M.qinit[i];
        dq[i] = active ? a.dq[(size_t)i * a.n + w] : (R)0;
        R t = 0;
        if (active && K.dof_act[i] >= 0) {
            R v = (R)sw[lane * K.n_act + K.dof_act[i]];
            v = v > K.dof_hi[i] ? K.dof_hi[i] : v;
            v = v < K.dof_lo[i] ? K.dof_lo[i] : v;
            t = v * K.dof_scale[i];
        }
        tau[i] = t;
    }
    __syncwarp();
    const R posbefore = q[0];
    for (int f = 0; f < K.frame_skip; f++) {
        const ContactSink<R>* sk = (active && f == K.frame_skip - 1) ? &a.sink : nullptr;
        substep_loop<R>(M, q, dq, tau, false, tau, tau, tau, K.fluid_force != 0, K.fluid_offset, K.fluid_coef, a.lcp_mode,
                        a.pgs_iters, sk, w);
    }
    const R ang = q[2];
    R r = (q[0] - posbefore) * K.inv_dt_env * K.vel_weight;
    r += K.alive_bonus;
    r -= K.ctrl_cost * a2;
    if (K.limit_pen_dof >= 0) {
        const int i = K.limit_pen_dof;
        R pen = 0;
        if ((M.qlo[i] - q[i]) > -K.limit_pen_margin) pen += (R)1.5;
        if ((M.qhi[i] - q[i]) < K.limit_pen_margin) pen += (R)1.5;
        r -= K.limit_pen_weight * pen;
    }
    r -= K.dev_cost * Num<R>::abs_(ang);
    bool ok = true;
    for (int i = 0; i < nb; i++) {
        if (i >= 2 && !(Num<R>::abs_(q[i]) < K.state_bound)) ok = false;
        if (i < 2 && !(Num<R>::abs_(q[i]) < Num<R>::inf())) ok = false;
        if (!(Num<R>::abs_(dq[i]) < K.state_bound)) ok = false;
    }
    if (K.zero_reward_on_blowup && !ok) r = 0;
    if (K.height_body >= 0) {
        const R h = body_height_loop<R>(M, K, q);
        ok = ok && (h > K.height_lo) && (h < K.height_hi);
    }
    ok = ok && (Num<R>::abs_(ang) < K.ang_max);
    bool done = !ok, trunc = false;
    if (active && a.max_episode_steps > 0) {
        const int el = a.elapsed[w] + 1;
        if (el >= a.max_episode_steps) { trunc = !done; done = true; }
        a.elapsed[w] = (done && a.auto_reset) ? 0 : el;
    }
    if (active && done && a.auto_reset) {
        const uint32_t ep = a.episode[w];
        reset_state_loop<R>(M, K, a.seed, a.world_offset + w, ep, q, dq);
        a.episode[w] = ep + 1;
    }
    if (active) write_obs_loop<R>(M, K, q, dq, sw + lane * K.n_obs);
    __syncwarp();
    if (cnt > 0) for (int k = lane; k < cnt * K.n_obs; k += 32) a.obs[(size_t)wb * K.n_obs + k] = sw[k];
    if (active) {
        for (int i = 0; i < nb; i++) { a.q[(size_t)i * a.n + w] = q[i]; a.dq[(size_t)i * a.n + w] = dq[i]; }
        a.reward[w] = (float)r;
        a.done[w] = done ? 1 : 0;
        if (a.truncated) a.truncated[w] = trunc ? 1 : 0;
    }
}

template <typename R>
__global__ void __launch_bounds__(256)
k_reset_loop(const __grid_constant__ PModel<R> M, const __grid_constant__ PTask<R> K, const __grid_constant__ StepArgs<R> a) {
    extern __shared__ float smem[];
    const int nb = M.nb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int wb = w - lane, cnt = min(32, a.n - wb);
    const bool active = w < a.n;
    float* sw = smem + warp * 32 * K.n_obs;
    R q[LOOP_MAXB], dq[LOOP_MAXB];
    const bool doit = active && (a.mask == nullptr || a.mask[w]);
    for (int i = 0; i < nb; i++) {
        q[i] = active ? a.q[(size_t)i * a.n + w] : M.qinit[i];
        dq[i] = active ? a.dq[(size_t)i * a.n + w] : (R)0;
    }
    if (doit) {
        const uint32_t ep = a.episode[w];
        reset_state_loop<R>(M, K, a.seed, a.world_offset + w, ep, q, dq);
        a.episode[w] = ep + 1;
        a.elapsed[w] = 0;
        for (int i = 0; i < nb; i++) { a.q[(size_t)i * a.n + w] = q[i]; a.dq[(size_t)i * a.n + w] = dq[i]; }
        if (a.sink.count) a.sink.count[w] = 0;
    }
    if (a.obs) {
        if (active) write_obs_loop<R>(M, K, q, dq, sw + lane * K.n_obs);
        __syncwarp();
        if (cnt > 0) for (int k = lane; k < cnt * K.n_obs; k += 32) a.obs[(size_t)wb * K.n_obs + k] = sw[k];
    }
}

template <typename R>
__global__ void __launch_bounds__(256)
k_substep_loop(const __grid_constant__ PModel<R> M, int n, R* qs, R* dqs, const R* tau_in, const R* fext, int lcp_mode,
               int pgs_iters, const __grid_constant__ ContactSink<R> sink) {
    const int nb = M.nb;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    R q[LOOP_MAXB], dq[LOOP_MAXB], tau[LOOP_MAXB], eft[LOOP_MAXB], efx[LOOP_MAXB], efy[LOOP_MAXB];
    for (int i = 0; i < nb; i++) {
        q[i] = qs[(size_t)i * n + w];
        dq[i] = dqs[(size_t)i * n + w];
        tau[i] = tau_in ? tau_in[(size_t)w * nb + i] : (R)0;
        eft[i] = 0; efx[i] = 0; efy[i] = 0;
    }
    if (fext) {
        R cs[LOOP_MAXB], sn[LOOP_MAXB], px[LOOP_MAXB], py[LOOP_MAXB];
        fk_positions_loop<R>(M, q, cs, sn, px, py);
        for (int k = 0; k < M.nbd; k++) {
            const R* f = fext + ((size_t)w * M.nbd + k) * 3;
            const R fx = M.e1[0] * f[0] + M.e1[1] * f[1] + M.e1[2] * f[2];
            const R fy = M.e2[0] * f[0] + M.e2[1] * f[1] + M.e2[2] * f[2];
            const int g = M.dgroup[k];
            const R ox = cs[g] * M.dox[k] - sn[g] * M.doy[k], oy = sn[g] * M.dox[k] + cs[g] * M.doy[k];
            eft[g] += ox * fy - oy * fx; efx[g] += fx; efy[g] += fy;
        }
        substep_loop<R>(M, q, dq, tau, true, eft, efx, efy, false, (R)0, (R)0, lcp_mode, pgs_iters, &sink, w);
    } else {
        substep_loop<R>(M, q, dq, tau, false, eft, efx, efy, false, (R)0, (R)0, lcp_mode, pgs_iters, &sink, w);
    }
    for (int i = 0; i < nb; i++) { qs[(size_t)i * n + w] = q[i]; dqs[(size_t)i * n + w] = dq[i]; }
}

// [n, nd] row-major <-> SoA [nd][n] with precision conversion
template <typename S, typename D>
__global__ void k_to_soa(int n, int nd, const S* src, D* dst) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n * nd) { const int w = k / nd, d = k % nd; dst[(size_t)d * n + w] = (D)src[k]; }
}
template <typename S, typename D>
__global__ void k_from_soa(int n, int nd, const S* src, D* dst) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n * nd) { const int w = k / nd, d = k % nd; dst[k] = (D)src[(size_t)d * n + w]; }
}
template <typename S, typename D>
__global__ void k_convert(size_t n, const S* src, D* dst) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) dst[k] = (D)src[k];
}

// ------------------------------------------------------------------------ engine
enum { TOPO_HOPPER = 0, TOPO_WALKER, TOPO_CHEETAH, TOPO_SNAKE, TOPO_COUNT };

struct dartb_engine {
    int device = 0, n = 0, nd = 0, topo = -1, max_contacts = 1, n_orig_bodies = 0;
    bool f64 = false;
    uint64_t seed = 0;
    int64_t world_offset = 0;
    PModel<float> mf; PModel<double> md;
    PTask<float> tf; PTask<double> td;
    dartb_model_t model; dartb_task_t task;   // kept so friction/options can re-lower
    void* q = nullptr; void* dq = nullptr;
    void* scratch = nullptr;                  // [n * max(nd, nbd*3)] of Real for tau / fext conversion
    uint32_t* episode = nullptr; int32_t* elapsed = nullptr; uint8_t* truncated = nullptr;
    int32_t* ccount = nullptr; int32_t* cbody = nullptr; float* cdata = nullptr;
    int lcp_mode = 0, pgs_iters = 30, max_episode_steps = 0;
    int variant_request = -1;                 // -1 auto (DARTB_VARIANT env or static if available), 0, 1
    int variant = 0;                          // 0 = unrolled static topology, 1 = loop / generic topology
    int64_t launches = 0;
    std::string kernel_name;
};

static int pick_topo(const std::string& sig, int n_limited) {
    if (sig == TopoHopper::sig && n_limited <= TopoHopper::NL) return TOPO_HOPPER;
    if (sig == TopoWalker::sig && n_limited <= TopoWalker::NL) return TOPO_WALKER;
    if (sig == TopoCheetah::sig && n_limited <= TopoCheetah::NL) return TOPO_CHEETAH;
    if (sig == TopoSnake::sig && n_limited <= TopoSnake::NL) return TOPO_SNAKE;
    return -1;
}
static const char* topo_name(int t) {
    switch (t) {
        case TOPO_HOPPER: return TopoHopper::name;
        case TOPO_WALKER: return TopoWalker::name;
        case TOPO_CHEETAH: return TopoCheetah::name;
        case TOPO_SNAKE: return TopoSnake::name;
    }
    return "?";
}

static int lower_into(dartb_engine* e) {
    lower::Result res;
    std::string why = lower::lower_model(e->model, e->task, res);
    if (!why.empty()) return fail("model cannot be lowered to the planar kernels: " + why);
    int nlim = 0;
    for (int i = 0; i < res.m.nb; i++) nlim += res.m.limited[i] ? 1 : 0;
    int topo = pick_topo(res.signature, nlim);
    if (res.m.ns > LOOP_MAXS || res.m.nb > LOOP_MAXB) return fail("model too large for the planar kernels");
    static int forced_variant = -1;
    if (forced_variant < 0) { const char* ev = getenv("DARTB_VARIANT"); forced_variant = ev ? atoi(ev) : 0; }
    if (topo < 0 || e->variant_request == 1 || (e->variant_request < 0 && forced_variant == 1)) e->variant = 1;
    else e->variant = 0;
    e->topo = topo;
    e->md = res.m; e->td = res.t;
    lower::convert(res.m, e->mf);
    lower::convert(res.t, e->tf);
    e->nd = res.m.nb;
    e->max_contacts = res.max_contacts;
    e->n_orig_bodies = e->model.n_bodies;
    const char* plane = std::fabs(res.m.en[2]) > 0.5 ? "planar-xy" : (std::fabs(res.m.en[1]) > 0.5 ? "planar-zx" : "planar-yz");
    e->kernel_name = std::string(plane) + (e->variant == 1 ? std::string("/loop:generic") : std::string("/static:") + topo_name(topo)) +
                     (e->f64 ? "/f64" : "/f32");
    return 0;
}

template <typename R> struct Sel;
template <> struct Sel<float> {
    static const PModel<float>& m(const dartb_engine* e) { return e->mf; }
    static const PTask<float>& t(const dartb_engine* e) { return e->tf; }
};
template <> struct Sel<double> {
    static const PModel<double>& m(const dartb_engine* e) { return e->md; }
    static const PTask<double>& t(const dartb_engine* e) { return e->td; }
};

static int block_for(int n) {
    // One warp per SM cannot hide instruction-fetch latency of the unrolled stepper (ncu: stall_no_inst
    // dominant); several warps per SM share the instruction stream.  DARTB_BLOCK overrides (experiments).
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("DARTB_BLOCK"); forced = e ? atoi(e) : 0; }
    if (forced >= 32 && forced <= 256 && forced % 32 == 0) return forced;
    return n <= 148 * 32 * 4 ? 32 : (n <= 148 * 64 * 8 ? 64 : 128);
}

template <typename R>
static StepArgs<R> make_args(dartb_engine* e) {
    StepArgs<R> a;
    std::memset(&a, 0, sizeof a);
    a.n = e->n; a.q = (R*)e->q; a.dq = (R*)e->dq; a.episode = e->episode; a.elapsed = e->elapsed;
    a.truncated = e->truncated;
    a.lcp_mode = e->lcp_mode; a.pgs_iters = e->pgs_iters; a.max_episode_steps = e->max_episode_steps;
    a.seed = e->seed; a.world_offset = e->world_offset;
    a.sink.count = e->ccount; a.sink.body = e->cbody; a.sink.data = e->cdata; a.sink.maxc = e->max_contacts;
    return a;
}

#define DISPATCH_TOPO(e, R, CALL)                                   \
    switch ((e)->topo) {                                            \
        case TOPO_HOPPER: { using T = TopoHopper; CALL; } break;    \
        case TOPO_WALKER: { using T = TopoWalker; CALL; } break;    \
        case TOPO_CHEETAH: { using T = TopoCheetah; CALL; } break;  \
        case TOPO_SNAKE: { using T = TopoSnake; CALL; } break;      \
        default: return fail("bad topology id");                    \
    }

template <typename R>
static int launch_step(dartb_engine* e, const float* action, float* obs, float* reward, uint8_t* done, int auto_reset,
                       cudaStream_t st) {
    StepArgs<R> a = make_args<R>(e);
    a.action = action; a.obs = obs; a.reward = reward; a.done = done; a.auto_reset = auto_reset;
    const int bs = block_for(e->n), grid = (e->n + bs - 1) / bs;
    const PTask<R>& K = Sel<R>::t(e);
    const int stage = K.n_obs > K.n_act ? K.n_obs : K.n_act;
    const size_t shm = (size_t)(bs / 32) * 32 * stage * sizeof(float);
    if (e->variant == 1) k_env_step_loop<R><<<grid, bs, shm, st>>>(Sel<R>::m(e), K, a);
    else DISPATCH_TOPO(e, R, (k_env_step<T, R><<<grid, bs, shm, st>>>(Sel<R>::m(e), K, a)));
    e->launches++;
    CK(cudaGetLastError());
    return 0;
}
template <typename R>
static int launch_reset(dartb_engine* e, const uint8_t* mask, float* obs, cudaStream_t st) {
    StepArgs<R> a = make_args<R>(e);
    a.mask = mask; a.obs = obs;
    const int bs = block_for(e->n), grid = (e->n + bs - 1) / bs;
    const PTask<R>& K = Sel<R>::t(e);
    const size_t shm = (size_t)(bs / 32) * 32 * K.n_obs * sizeof(float);
    if (e->variant == 1) k_reset_loop<R><<<grid, bs, shm, st>>>(Sel<R>::m(e), K, a);
    else DISPATCH_TOPO(e, R, (k_reset<T, R><<<grid, bs, shm, st>>>(Sel<R>::m(e), K, a)));
    e->launches++;
    CK(cudaGetLastError());
    return 0;
}
template <typename R>
static int launch_substep(dartb_engine* e, const R* tau, const R* fext, cudaStream_t st) {
    ContactSink<R> sink;
    sink.count = e->ccount; sink.body = e->cbody; sink.data = e->cdata; sink.maxc = e->max_contacts;
    const int bs = block_for(e->n), grid = (e->n + bs - 1) / bs;
    if (e->variant == 1)
        k_substep_loop<R><<<grid, bs, 0, st>>>(Sel<R>::m(e), e->n, (R*)e->q, (R*)e->dq, tau, fext, e->lcp_mode, e->pgs_iters, sink);
    else
        DISPATCH_TOPO(e, R, (k_substep<T, R><<<grid, bs, 0, st>>>(Sel<R>::m(e), e->n, (R*)e->q, (R*)e->dq, tau, fext,
                                                                 e->lcp_mode, e->pgs_iters, sink)));
    e->launches++;
    CK(cudaGetLastError());
    return 0;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int create_impl(const dartb_model_t* model, const dartb_task_t* task, int32_t n, int32_t device, uint64_t seed,
                       int64_t world_offset, bool f64, dartb_handle_t* out) {
    if (!model || !task || !out) return fail("null argument");
    if (n <= 0) return fail("n_worlds must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("no CUDA device: the B200 engine has no CPU fallback");
    if (device < 0 || device >= ndev) return fail("bad device index");
    dartb_engine* e = new (std::nothrow) dartb_engine();
    if (!e) return fail("out of memory");
    e->device = device; e->n = n; e->seed = seed; e->world_offset = world_offset; e->f64 = f64;
    e->model = *model; e->task = *task;
    if (lower_into(e)) { delete e; return 1; }
    DeviceGuard g(device);
    const size_t rs = f64 ? 8 : 4;
    const size_t sc = (size_t)n * (size_t)std::max(e->nd, e->n_orig_bodies * 3);
    cudaError_t err = cudaSuccess;
    auto A = [&](void** p, size_t bytes) { if (err == cudaSuccess) { err = cudaMalloc(p, bytes); if (err == cudaSuccess) err = cudaMemset(*p, 0, bytes); } };
    A(&e->q, rs * n * e->nd); A(&e->dq, rs * n * e->nd); A(&e->scratch, rs * sc);
    A((void**)&e->episode, 4 * (size_t)n); A((void**)&e->elapsed, 4 * (size_t)n); A((void**)&e->truncated, (size_t)n);
    A((void**)&e->ccount, 4 * (size_t)n); A((void**)&e->cbody, 4 * (size_t)n * e->max_contacts);
    A((void**)&e->cdata, 4 * (size_t)n * e->max_contacts * 10);
    if (err != cudaSuccess) { dartb_destroy(e); return fail(std::string("cudaMalloc: ") + cudaGetErrorString(err)); }
    *out = e;
    // initial state = q_init / dq_init (the pydart World constructor resets the world)
    {
        const int nd = e->nd;
        if (f64) {
            std::vector<double> hq((size_t)n * nd), hv((size_t)n * nd);
            for (int d = 0; d < nd; d++) for (int w = 0; w < n; w++) { hq[(size_t)d * n + w] = e->md.qinit[d]; hv[(size_t)d * n + w] = e->md.dqinit[d]; }
            cudaMemcpy(e->q, hq.data(), hq.size() * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(e->dq, hv.data(), hv.size() * 8, cudaMemcpyHostToDevice);
        } else {
            std::vector<float> hq((size_t)n * nd), hv((size_t)n * nd);
            for (int d = 0; d < nd; d++) for (int w = 0; w < n; w++) { hq[(size_t)d * n + w] = e->mf.qinit[d]; hv[(size_t)d * n + w] = e->mf.dqinit[d]; }
            cudaMemcpy(e->q, hq.data(), hq.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(e->dq, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice);
        }
    }
    return 0;
}

template <typename S>
static int set_state_impl(dartb_handle_t e, const S* q, const S* dq, cudaStream_t st) {
    const int tot = e->n * e->nd, bs = 256, grid = (tot + bs - 1) / bs;
    if (e->f64) {
        if (q) { k_to_soa<S, double><<<grid, bs, 0, st>>>(e->n, e->nd, q, (double*)e->q); e->launches++; }
        if (dq) { k_to_soa<S, double><<<grid, bs, 0, st>>>(e->n, e->nd, dq, (double*)e->dq); e->launches++; }
    } else {
        if (q) { k_to_soa<S, float><<<grid, bs, 0, st>>>(e->n, e->nd, q, (float*)e->q); e->launches++; }
        if (dq) { k_to_soa<S, float><<<grid, bs, 0, st>>>(e->n, e->nd, dq, (float*)e->dq); e->launches++; }
    }
    CK(cudaGetLastError());
    return 0;
}
template <typename D>
static int get_state_impl(dartb_handle_t e, D* q, D* dq, cudaStream_t st) {
    const int tot = e->n * e->nd, bs = 256, grid = (tot + bs - 1) / bs;
    if (e->f64) {
        if (q) { k_from_soa<double, D><<<grid, bs, 0, st>>>(e->n, e->nd, (const double*)e->q, q); e->launches++; }
        if (dq) { k_from_soa<double, D><<<grid, bs, 0, st>>>(e->n, e->nd, (const double*)e->dq, dq); e->launches++; }
    } else {
        if (q) { k_from_soa<float, D><<<grid, bs, 0, st>>>(e->n, e->nd, (const float*)e->q, q); e->launches++; }
        if (dq) { k_from_soa<float, D><<<grid, bs, 0, st>>>(e->n, e->nd, (const float*)e->dq, dq); e->launches++; }
    }
    CK(cudaGetLastError());
    return 0;
}
extern "C" {

int dartb_create(const dartb_model_t* model, const dartb_task_t* task, int32_t n_worlds, int32_t device, uint64_t seed,
                 int64_t world_offset, dartb_handle_t* out) {
    return create_impl(model, task, n_worlds, device, seed, world_offset, false, out);
}
int dartb_create_f64(const dartb_model_t* model, const dartb_task_t* task, int32_t n_worlds, int32_t device,
                     uint64_t seed, int64_t world_offset, dartb_handle_t* out) {
    return create_impl(model, task, n_worlds, device, seed, world_offset, true, out);
}

int dartb_destroy(dartb_handle_t e) {
    if (!e) return 0;
    DeviceGuard g(e->device);
    cudaFree(e->q); cudaFree(e->dq); cudaFree(e->scratch); cudaFree(e->episode); cudaFree(e->elapsed);
    cudaFree(e->truncated); cudaFree(e->ccount); cudaFree(e->cbody); cudaFree(e->cdata);
    delete e;
    return 0;
}

int dartb_set_option(dartb_handle_t e, int32_t key, double value) {
    if (!e) return fail("null handle");
    switch (key) {
        case DARTB_OPT_LCP_MODE:
            if (value != 0 && value != 1) return fail("lcp mode must be 0 (exact) or 1 (PGS)");
            e->lcp_mode = (int)value; return 0;
        case DARTB_OPT_PGS_ITERS:
            if (value < 1 || value > 10000) return fail("bad PGS iteration count");
            e->pgs_iters = (int)value; return 0;
        case DARTB_OPT_FRICTION_ALL:
            for (int i = 0; i < e->model.n_bodies; i++) e->model.bodies[i].friction_coeff = value;
            return lower_into(e);
        case DARTB_OPT_KERNEL_VARIANT:
            if (value != 0 && value != 1) return fail("kernel variant must be 0 (unrolled) or 1 (loop)");
            e->variant_request = (int)value;
            return lower_into(e);
        case DARTB_OPT_MAX_EPISODE_STEPS:
            if (value < 0) return fail("bad max_episode_steps");
            e->max_episode_steps = (int)value; return 0;
    }
    return fail("unknown option key");
}

int dartb_reset(dartb_handle_t e, const uint8_t* d_mask, float* d_obs, void* stream) {
    if (!e) return fail("null handle");
    DeviceGuard g(e->device);
    return e->f64 ? launch_reset<double>(e, d_mask, d_obs, (cudaStream_t)stream)
                  : launch_reset<float>(e, d_mask, d_obs, (cudaStream_t)stream);
}

int dartb_set_state(dartb_handle_t e, const float* q, const float* dq, void* s) {
    if (!e) return fail("null handle");
    DeviceGuard g(e->device);
    return set_state_impl<float>(e, q, dq, (cudaStream_t)s);
}
int dartb_get_state(dartb_handle_t e, float* q, float* dq, void* s) {
    if (!e) return fail("null handle");
    DeviceGuard g(e->device);
    return get_state_impl<float>(e, q, dq, (cudaStream_t)s);
}
int dartb_set_state_f64(dartb_handle_t e, const double* q, const double* dq, void* s) {
    if (!e) return fail("null handle");
    DeviceGuard g(e->device);
    return set_state_impl<double>(e, q, dq, (cudaStream_t)s);
}
int dartb_get_state_f64(dartb_handle_t e, double* q, double* dq, void* s) {
    if (!e) return fail("null handle");
    DeviceGuard g(e->device);
    return get_state_impl<double>(e, q, dq, (cudaStream_t)s);
}

int dartb_step(dartb_handle_t e, const float* d_action, float* d_obs, float* d_reward, uint8_t* d_done,
               int32_t auto_reset, void* stream) {
    if (!e) return fail("null handle");
    if (!d_action || !d_obs || !d_reward || !d_done) return fail("null device pointer");
    DeviceGuard g(e->device);
    return e->f64 ? launch_step<double>(e, d_action, d_obs, d_reward, d_done, auto_reset, (cudaStream_t)stream)
                  : launch_step<float>(e, d_action, d_obs, d_reward, d_done, auto_reset, (cudaStream_t)stream);
}

int dartb_substep(dartb_handle_t e, const float* d_tau, const float* d_fext, void* stream) {
    if (!e) return fail("null handle");
    DeviceGuard g(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (!e->f64) return launch_substep<float>(e, d_tau, d_fext, st);
    // fp64 engine fed fp32 inputs: widen through the scratch buffer
    double* sc = (double*)e->scratch;
    const double* tau = nullptr; const double* fx = nullptr;
    if (d_tau && d_fext) return fail("fp64 engine: pass tau and fext through dartb_substep_f64");
    if (d_tau) { size_t k = (size_t)e->n * e->nd; k_convert<float, double><<<(unsigned)((k + 255) / 256), 256, 0, st>>>(k, d_tau, sc); tau = sc; e->launches++; }
    if (d_fext) { size_t k = (size_t)e->n * e->n_orig_bodies * 3; k_convert<float, double><<<(unsigned)((k + 255) / 256), 256, 0, st>>>(k, d_fext, sc); fx = sc; e->launches++; }
    return launch_substep<double>(e, tau, fx, st);
}
int dartb_substep_f64(dartb_handle_t e, const double* d_tau, const double* d_fext, void* stream) {
    if (!e) return fail("null handle");
    DeviceGuard g(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (e->f64) return launch_substep<double>(e, d_tau, d_fext, st);
    if (d_tau && d_fext) return fail("fp32 engine: pass tau and fext through dartb_substep");
    float* sc = (float*)e->scratch;
    const float* tau = nullptr; const float* fx = nullptr;
    if (d_tau) { size_t k = (size_t)e->n * e->nd; k_convert<double, float><<<(unsigned)((k + 255) / 256), 256, 0, st>>>(k, d_tau, sc); tau = sc; e->launches++; }
    if (d_fext) { size_t k = (size_t)e->n * e->n_orig_bodies * 3; k_convert<double, float><<<(unsigned)((k + 255) / 256), 256, 0, st>>>(k, d_fext, sc); fx = sc; e->launches++; }
    return launch_substep<float>(e, tau, fx, st);
}

int dartb_get_contacts(dartb_handle_t e, int32_t* d_count, int32_t* d_body, float* d_data, void* stream) {
    if (!e) return fail("null handle");
    DeviceGuard g(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (d_count) CK(cudaMemcpyAsync(d_count, e->ccount, 4 * (size_t)e->n, cudaMemcpyDeviceToDevice, st));
    if (d_body) CK(cudaMemcpyAsync(d_body, e->cbody, 4 * (size_t)e->n * e->max_contacts, cudaMemcpyDeviceToDevice, st));
    if (d_data) CK(cudaMemcpyAsync(d_data, e->cdata, 4 * (size_t)e->n * e->max_contacts * 10, cudaMemcpyDeviceToDevice, st));
    return 0;
}
int dartb_get_truncated(dartb_handle_t e, uint8_t* d_out, void* stream) {
    if (!e || !d_out) return fail("null argument");
    DeviceGuard g(e->device);
    CK(cudaMemcpyAsync(d_out, e->truncated, (size_t)e->n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}
int32_t dartb_max_contacts(dartb_handle_t e) { return e ? e->max_contacts : 0; }
int32_t dartb_num_worlds(dartb_handle_t e) { return e ? e->n : 0; }
int32_t dartb_num_dofs(dartb_handle_t e) { return e ? e->nd : 0; }
int32_t dartb_is_f64(dartb_handle_t e) { return e && e->f64 ? 1 : 0; }
int64_t dartb_launch_count(dartb_handle_t e) { return e ? e->launches : 0; }
const char* dartb_kernel_name(dartb_handle_t e) { return e ? e->kernel_name.c_str() : ""; }
const char* dartb_last_error(void) { return g_err.c_str(); }

int dartb_describe(const dartb_model_t* model, const dartb_task_t* task, char* buf, int32_t len) {
    if (!model || !task || !buf || len < 2) return fail("null argument");
    dartb_engine e;
    e.model = *model; e.task = *task;
    if (lower_into(&e)) return 1;
    std::snprintf(buf, (size_t)len, "%s nd=%d max_contacts=%d", e.kernel_name.c_str(), e.nd, e.max_contacts);
    return 0;
}
const char* dartb_version(void) { return "dartb 0.1 (sm_100a)"; }

}  // extern "C"
