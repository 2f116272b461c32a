// kernels.cuh — the __global__ kernels of the stepper (templated on topology / precision) and the
// launcher declarations that let each (topology, precision) pair live in its own translation unit
// (inst.cu is compiled once per pair, in parallel; see build.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dartb.h"
#include "planar_kernels.cuh"
#include "planar_loop.cuh"
#include "planar_coop.cuh"
#include "task_kinds.cuh"

// resident blocks (of 128 threads) per SM the per-thread kernels are compiled for: 1 = up to 255 registers
#ifndef DARTB_STEP_MIN_BLOCKS
#define DARTB_STEP_MIN_BLOCKS 1
#endif

// Experiment (DARTB_PREFETCH_PARAMS): touch one word of every 64-byte line of the kernel parameters at entry so that the
// constant-cache misses of the model overlap instead of being met one by one along K1 / K2 of the first DART step.
#ifndef DARTB_PREFETCH_PARAMS
#define DARTB_PREFETCH_PARAMS 0
#endif
template <typename S>
DEVI uint32_t touch_params(const S& s) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(&s);
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(S) / 4); i += 16) acc |= p[i];
    return acc;
}

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
// FLUID (snake_7link.py:35-47) is a template parameter so that the contact topologies do not carry a second,
// never-executed copy of the whole DART step (r1: 48.6 k SASS instructions per kernel, half of them that copy)
template <class T, typename R, bool FLUID>
__global__ void __launch_bounds__(128, DARTB_STEP_MIN_BLOCKS)
k_env_step(const __grid_constant__ PModel<R> M, const __grid_constant__ PTask<R> K, const __grid_constant__ StepArgs<R> a) {
    constexpr int NB = T::NB;
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wb = (blockIdx.x * (blockDim.x >> 5) + warp) * a.wpw;   // first world of this warp
    const int w = wb + lane;
    const int cnt = min(a.wpw, a.n - wb);             // worlds this warp owns (<= 0: idle warp)
    const bool active = lane < cnt;
    const int stage = K.n_obs > K.n_act ? K.n_obs : K.n_act;
    float* sw = smem + warp * 32 * stage;

    // the state loads are issued BEFORE the action staging below waits for its loads (host-facing step: a PCIe round
    // trip), so that the prologue pays one memory latency, not two
    R q[NB], dq[NB], tau[NB], zero[NB];
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        q[i] = active ? a.q[(size_t)i * a.n + w] : M.qinit[i];
        dq[i] = active ? a.dq[(size_t)i * a.n + w] : (R)0;
        zero[i] = 0;
    });
    // coalesced action load -> smem [lane][n_act]
    if (cnt > 0) for (int k = lane; k < cnt * K.n_act; k += 32) sw[k] = a.action[(size_t)wb * K.n_act + k];
    __syncwarp();

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
    // contact read-back is opt-in (DARTB_OPT_CONTACTS): 356 B per world, more than the rest of the step's HBM traffic
    const ContactSink<R>* sink = a.sink.count ? &a.sink : nullptr;
    uint64_t hint = active ? a.hint[w] : ~(uint64_t)0;
#pragma unroll 1
    for (int f = 0; f < K.frame_skip; f++) {
        const ContactSink<R>* sk = (active && f == K.frame_skip - 1) ? sink : nullptr;
        substep<T, R, false, FLUID>(M, q, dq, tau, zero, zero, zero, K.fluid_offset, K.fluid_coef, a.lcp_mode, a.pgs_iters, sk, w, hint);
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
        reset_state<T, R>(M, K, reset_seed(a, w), reset_world(a, w), ep, q, dq);
        a.episode[w] = ep + 1;
        hint = ~(uint64_t)0;
    }
    if (active) a.hint[w] = hint;
    // obs (of the reset state for auto-reset worlds: gym/vector/sync_vector_env.py:76-79)
    if (active) write_obs<T, R>(M, K, q, dq, sw + lane * K.n_obs);
    __syncwarp();
    if (cnt > 0) for (int k = lane; k < cnt * K.n_obs; k += 32) store_obs(a, (size_t)wb * K.n_obs + k, sw[k]);
    if (active) {
        static_for<0, NB>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            a.q[(size_t)i * a.n + w] = q[i];
            a.dq[(size_t)i * a.n + w] = dq[i];
        });
        if (a.reward64) { a.reward64[w] = (double)r; a.done[w] = done ? 1 : 0; }
        else {
            a.reward[w] = (float)r;
            a.done[w] = (uint8_t)((done ? 1 : 0) | (trunc ? 2 : 0));  // bit 0 done, bit 1 TimeLimit.truncated
        }
        if (a.truncated) a.truncated[w] = trunc ? 1 : 0;
    }
}

// ------------------------------------------------------------------------ env.step() kernel, group ("quad") form
// G = 2, 4 or 8 lanes per world (substep<..., G>): lane / G picks the world inside the warp, all lanes of a group carry
// the same state and the same task-layer arithmetic, lane % G == 0 writes.  Worlds per warp = 32 / G.
#ifndef DARTB_QUAD_MIN_BLOCKS
#define DARTB_QUAD_MIN_BLOCKS 1
#endif
template <class T, typename R, bool FLUID, int G>
__global__ void __launch_bounds__(128, DARTB_QUAD_MIN_BLOCKS)
k_env_step_quad(const __grid_constant__ PModel<R> M, const __grid_constant__ PTask<R> K, const __grid_constant__ StepArgs<R> a) {
    constexpr int NB = T::NB, WPW = 32 / G;
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gi = lane / G, l = lane % G;
    const int wb = (blockIdx.x * (blockDim.x >> 5) + warp) * WPW;   // first world of this warp
    const int w = wb + gi;
    const int cnt = min(WPW, a.n - wb);
    const bool active = gi < cnt;
    const int stage = K.n_obs > K.n_act ? K.n_obs : K.n_act;
    float* sw = smem + warp * WPW * stage;
#if DARTB_PREFETCH_PARAMS
    if ((touch_params(M) | touch_params(K)) == 0x7fc54321u && a.obs) a.obs[0] = 0;   // (never true: keeps the loads)
#endif
    // (state and counter loads first, then the action staging: one memory latency for the prologue, see k_env_step)
    R q[NB], dq[NB], tau[NB], zero[NB];
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        q[i] = active ? a.q[(size_t)i * a.n + w] : M.qinit[i];
        dq[i] = active ? a.dq[(size_t)i * a.n + w] : (R)0;
        zero[i] = 0;
    });
    // per-world counters: every lane of the group reads them before lane 0 writes them back at the end
    const int el_in = (active && a.max_episode_steps > 0) ? a.elapsed[w] : 0;
    const uint32_t ep_in = active ? a.episode[w] : 0u;
    uint64_t hint = active ? a.hint[w] : ~(uint64_t)0;
    if (cnt > 0) for (int k = lane; k < cnt * K.n_act; k += 32) sw[k] = a.action[(size_t)wb * K.n_act + k];
    __syncwarp();
    R a2 = 0;
    if (active) for (int j = 0; j < K.n_act; j++) { const R v = (R)sw[gi * K.n_act + j]; a2 += v * v; }
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        R t = 0;
        if (active && K.dof_act[i] >= 0) {
            R v = (R)sw[gi * K.n_act + K.dof_act[i]];
            v = v > K.dof_hi[i] ? K.dof_hi[i] : v;
            v = v < K.dof_lo[i] ? K.dof_lo[i] : v;
            t = v * K.dof_scale[i];
        }
        tau[i] = t;
    });
    __syncwarp();
    const R posbefore = q[0];
    const ContactSink<R>* sink = (a.sink.count && l == 0) ? &a.sink : nullptr;
#pragma unroll 1
    for (int f = 0; f < K.frame_skip; f++) {
        const ContactSink<R>* sk = (active && f == K.frame_skip - 1) ? sink : nullptr;
        substep<T, R, false, FLUID, G>(M, q, dq, tau, zero, zero, zero, K.fluid_offset, K.fluid_coef, a.lcp_mode, a.pgs_iters, sk, w, hint);
    }
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
    int el_out = el_in;
    if (active && a.max_episode_steps > 0) {
        const int el = el_in + 1;
        if (el >= a.max_episode_steps) { trunc = !done; done = true; }
        el_out = (done && a.auto_reset) ? 0 : el;
    }
    const bool do_reset = active && done && a.auto_reset;
    if (do_reset) {
        reset_state<T, R>(M, K, reset_seed(a, w), reset_world(a, w), ep_in, q, dq);
        hint = ~(uint64_t)0;
    }
    __syncwarp();
    const bool writer = active && l == 0;
    if (writer) {
        if (a.max_episode_steps > 0) a.elapsed[w] = el_out;
        if (do_reset) a.episode[w] = ep_in + 1;
        a.hint[w] = hint;
        write_obs<T, R>(M, K, q, dq, sw + gi * K.n_obs);
    }
    __syncwarp();
    if (cnt > 0) for (int k = lane; k < cnt * K.n_obs; k += 32) store_obs(a, (size_t)wb * K.n_obs + k, sw[k]);
    if (writer) {
        static_for<0, NB>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            a.q[(size_t)i * a.n + w] = q[i];
            a.dq[(size_t)i * a.n + w] = dq[i];
        });
        if (a.reward64) { a.reward64[w] = (double)r; a.done[w] = done ? 1 : 0; }
        else {
            a.reward[w] = (float)r;
            a.done[w] = (uint8_t)((done ? 1 : 0) | (trunc ? 2 : 0));  // bit 0 done, bit 1 TimeLimit.truncated
        }
        if (a.truncated) a.truncated[w] = trunc ? 1 : 0;
    }
}

// exactly `skel.set_forces(tau); world.step()`, quad form (no external forces)
template <class T, typename R, int G>
__global__ void __launch_bounds__(128, DARTB_QUAD_MIN_BLOCKS)
k_substep_quad(const __grid_constant__ PModel<R> M, int n, R* qs, R* dqs, const R* tau_in /*[n,nd]*/, int lcp_mode, int pgs_iters,
               const __grid_constant__ ContactSink<R> sink) {
    constexpr int NB = T::NB, WPW = 32 / G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gi = lane / G, l = lane % G;
    const int w = (blockIdx.x * (blockDim.x >> 5) + warp) * WPW + gi;
    const bool active = w < n;
    const int wr = active ? w : 0;
    R q[NB], dq[NB], tau[NB], zero[NB];
    static_for<0, NB>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        q[i] = active ? qs[(size_t)i * n + wr] : M.qinit[i];
        dq[i] = active ? dqs[(size_t)i * n + wr] : (R)0;
        tau[i] = (active && tau_in) ? tau_in[(size_t)wr * NB + i] : (R)0;
        zero[i] = 0;
    });
    uint64_t hint = ~(uint64_t)0;  // the literal World.step() drop-in is stateless
    __syncwarp();                  // (every lane has read its world's state before lane 0 of a group writes it)
    substep<T, R, false, false, G>(M, q, dq, tau, zero, zero, zero, (R)0, (R)0, lcp_mode, pgs_iters, (active && l == 0) ? &sink : nullptr, wr, hint);
    __syncwarp();
    if (active && l == 0)
        static_for<0, NB>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            qs[(size_t)i * n + w] = q[i];
            dqs[(size_t)i * n + w] = dq[i];
        });
}

// ------------------------------------------------------------------------ reset kernel
template <class T, typename R>
__global__ void __launch_bounds__(128, DARTB_STEP_MIN_BLOCKS)
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
        reset_state<T, R>(M, K, reset_seed(a, w), reset_world(a, w), ep, q, dq);
        a.episode[w] = ep + 1;
        a.elapsed[w] = 0;
        a.hint[w] = ~(uint64_t)0;
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
__global__ void __launch_bounds__(128, DARTB_STEP_MIN_BLOCKS)
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
        uint64_t hint = ~(uint64_t)0;  // the literal World.step() drop-in is stateless
        substep<T, R, true, false>(M, q, dq, tau, eft, efx, efy, (R)0, (R)0, lcp_mode, pgs_iters, &sink, w, hint);
    } else {
        uint64_t hint = ~(uint64_t)0;
        substep<T, R, false, false>(M, q, dq, tau, eft, efx, efy, (R)0, (R)0, lcp_mode, pgs_iters, &sink, w, hint);
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

// snake_7link.py:115-120 inside the kernel: at every reset_model of a world its bodynode masses become original + U(-r, r)
// and its friction coefficients original + U(-r, r), both clipped at 0 (the reference would hand DART the negative value),
// written into the world's column of the per-world table.  One draw per planar body (= bodynode: dartb.cu refuses the
// options for skeletons with welded bodynodes); a capsule's coefficient is min(its body's, ground = 1) as in lower.h.
// Draws come from the reset generator at indices past the state noise (2 PM_MAXB + i, 3 PM_MAXB + i).
template <typename R>
DEVI void redraw_dynamics(const PModel<R>& M, const StepArgs<R>& a, int w, uint64_t seed, int64_t gw, uint32_t ep) {
    if (!a.wpar || !(a.rand_mass > 0 || a.rand_mu > 0)) return;
    const int nb = M.nb, ns = M.ns;
    const size_t n = (size_t)a.n;
    if (a.rand_mass > 0)
        for (int i = 0; i < nb; i++) {
            const float m = __fadd_rn((float)M.mass[i], __fmul_rn(reset_uniform(seed, gw, ep, 2 * PM_MAXB + i), a.rand_mass));
            a.wpar[(size_t)i * n + w] = (R)(m > 0.0f ? m : 0.0f);
        }
    if (a.rand_mu > 0)
        for (int s = 0; s < ns; s++) {
            float mu = __fadd_rn((float)M.smu[s], __fmul_rn(reset_uniform(seed, gw, ep, 3 * PM_MAXB + M.sbody[s]), a.rand_mu));
            mu = mu > 0.0f ? mu : 0.0f;
            a.wpar[(size_t)(4 * nb + s) * n + w] = (R)(mu < 1.0f ? mu : 1.0f);
        }
}

template <typename R>
__global__ void __launch_bounds__(128, DARTB_STEP_MIN_BLOCKS)
k_env_step_loop(const __grid_constant__ PModel<R> M, const __grid_constant__ PTask<R> K, const __grid_constant__ StepArgs<R> a) {
    extern __shared__ float smem[];
    const int nb = M.nb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wb = (blockIdx.x * (blockDim.x >> 5) + warp) * a.wpw;
    const int w = wb + lane, cnt = min(a.wpw, a.n - wb);
    const bool active = lane < cnt;
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
    R tg[3] = {0, 0, 0};
    if (a.aux && active) { tg[0] = a.aux[w]; tg[1] = a.aux[(size_t)a.n + w]; tg[2] = a.aux[2 * (size_t)a.n + w]; }
    for (int f = 0; f < K.frame_skip; f++) {
        const ContactSink<R>* sk = (active && f == K.frame_skip - 1 && a.sink.count) ? &a.sink : nullptr;
        substep_loop<R>(M, q, dq, tau, false, tau, tau, tau, K.fluid_force != 0, K.fluid_offset, K.fluid_coef, a.lcp_mode,
                        a.pgs_iters, sk, w, (a.wpar && active) ? a.wpar + w : nullptr, (size_t)a.n);
    }
    if (K.kind != DARTB_TASK_LOCOMOTION) {
        // the contact-free envs: their own reward / done / reset / obs formulas, same TimeLimit and auto-reset plumbing
        R r; bool done, trunc = false;
        task_kind_eval<R>(M, K, q, dq, tg, a2, r, done);
        if (active && a.max_episode_steps > 0) {
            const int el = a.elapsed[w] + 1;
            if (el >= a.max_episode_steps) { trunc = !done; done = true; }
            a.elapsed[w] = (done && a.auto_reset) ? 0 : el;
        }
        if (active && done && a.auto_reset) {
            const uint32_t ep = a.episode[w];
            reset_state_kind<R>(M, K, reset_seed(a, w), reset_world(a, w), ep, q, dq, tg);
            redraw_dynamics<R>(M, a, w, reset_seed(a, w), reset_world(a, w), ep);
            a.episode[w] = ep + 1;
            if (a.aux) { a.aux[w] = tg[0]; a.aux[(size_t)a.n + w] = tg[1]; a.aux[2 * (size_t)a.n + w] = tg[2]; }
        }
        if (active) write_obs_kind<R>(M, K, q, dq, tg, sw + lane * K.n_obs);
        __syncwarp();
        if (cnt > 0) for (int k = lane; k < cnt * K.n_obs; k += 32) store_obs(a, (size_t)wb * K.n_obs + k, sw[k]);
        if (active) {
            for (int i = 0; i < nb; i++) { a.q[(size_t)i * a.n + w] = q[i]; a.dq[(size_t)i * a.n + w] = dq[i]; }
            if (a.reward64) { a.reward64[w] = (double)r; a.done[w] = done ? 1 : 0; }
            else { a.reward[w] = (float)r; a.done[w] = (uint8_t)((done ? 1 : 0) | (trunc ? 2 : 0)); }
            if (a.truncated) a.truncated[w] = trunc ? 1 : 0;
        }
        return;
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
        reset_state_loop<R>(M, K, reset_seed(a, w), reset_world(a, w), ep, q, dq);
        redraw_dynamics<R>(M, a, w, reset_seed(a, w), reset_world(a, w), ep);
        a.episode[w] = ep + 1;
    }
    if (active) write_obs_loop<R>(M, K, q, dq, sw + lane * K.n_obs);
    __syncwarp();
    if (cnt > 0) for (int k = lane; k < cnt * K.n_obs; k += 32) store_obs(a, (size_t)wb * K.n_obs + k, sw[k]);
    if (active) {
        for (int i = 0; i < nb; i++) { a.q[(size_t)i * a.n + w] = q[i]; a.dq[(size_t)i * a.n + w] = dq[i]; }
        if (a.reward64) { a.reward64[w] = (double)r; a.done[w] = done ? 1 : 0; }
        else {
            a.reward[w] = (float)r;
            a.done[w] = (uint8_t)((done ? 1 : 0) | (trunc ? 2 : 0));  // bit 0 done, bit 1 TimeLimit.truncated
        }
        if (a.truncated) a.truncated[w] = trunc ? 1 : 0;
    }
}

template <typename R>
__global__ void __launch_bounds__(128, DARTB_STEP_MIN_BLOCKS)
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
    R tg[3] = {0, 0, 0};
    if (a.aux && active) { tg[0] = a.aux[w]; tg[1] = a.aux[(size_t)a.n + w]; tg[2] = a.aux[2 * (size_t)a.n + w]; }
    if (doit) {
        const uint32_t ep = a.episode[w];
        if (K.kind == DARTB_TASK_LOCOMOTION) reset_state_loop<R>(M, K, reset_seed(a, w), reset_world(a, w), ep, q, dq);
        else {
            reset_state_kind<R>(M, K, reset_seed(a, w), reset_world(a, w), ep, q, dq, tg);
            if (a.aux) { a.aux[w] = tg[0]; a.aux[(size_t)a.n + w] = tg[1]; a.aux[2 * (size_t)a.n + w] = tg[2]; }
        }
        redraw_dynamics<R>(M, a, w, reset_seed(a, w), reset_world(a, w), ep);
        a.episode[w] = ep + 1;
        a.elapsed[w] = 0;
        a.hint[w] = ~(uint64_t)0;
        for (int i = 0; i < nb; i++) { a.q[(size_t)i * a.n + w] = q[i]; a.dq[(size_t)i * a.n + w] = dq[i]; }
        if (a.sink.count) a.sink.count[w] = 0;
    }
    if (a.obs) {
        if (active) { if (K.kind == DARTB_TASK_LOCOMOTION) write_obs_loop<R>(M, K, q, dq, sw + lane * K.n_obs); else write_obs_kind<R>(M, K, q, dq, tg, sw + lane * K.n_obs); }
        __syncwarp();
        if (cnt > 0) for (int k = lane; k < cnt * K.n_obs; k += 32) a.obs[(size_t)wb * K.n_obs + k] = sw[k];
    }
}

template <typename R>
__global__ void __launch_bounds__(128, DARTB_STEP_MIN_BLOCKS)
k_substep_loop(const __grid_constant__ PModel<R> M, int n, R* qs, R* dqs, const R* tau_in, const R* fext, int lcp_mode,
               int pgs_iters, const __grid_constant__ ContactSink<R> sink, const R* wpar) {
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
        substep_loop<R>(M, q, dq, tau, true, eft, efx, efy, false, (R)0, (R)0, lcp_mode, pgs_iters, &sink, w, wpar ? wpar + w : nullptr, (size_t)n);
    } else {
        substep_loop<R>(M, q, dq, tau, false, eft, efx, efy, false, (R)0, (R)0, lcp_mode, pgs_iters, &sink, w, wpar ? wpar + w : nullptr, (size_t)n);
    }
    for (int i = 0; i < nb; i++) { qs[(size_t)i * n + w] = q[i]; dqs[(size_t)i * n + w] = dq[i]; }
}


// ------------------------------------------------------------------------ launchers (one set per TU)
template <typename R>
struct Launchers {
    void (*step)(int grid, int bs, size_t shm, cudaStream_t st, const PModel<R>& M, const PTask<R>& K, const StepArgs<R>& a);
    void (*reset)(int grid, int bs, size_t shm, cudaStream_t st, const PModel<R>& M, const PTask<R>& K, const StepArgs<R>& a);
    void (*substep)(int grid, int bs, cudaStream_t st, const PModel<R>& M, int n, R* q, R* dq, const R* tau, const R* fext,
                    int lcp_mode, int pgs_iters, const ContactSink<R>& sink, const R* wpar);   // wpar: loop variant only
    // lane-cooperative kernels (planar_coop.cuh); null for the loop variant.  They size their own grid.
    void (*step_coop)(cudaStream_t st, const PModel<R>& M, const PTask<R>& K, const StepArgs<R>& a, const void* lane_table);
    void (*substep_coop)(cudaStream_t st, const PModel<R>& M, const void* lane_table, int n, R* q, R* dq, const R* tau, int lcp_mode,
                         int pgs_iters, const ContactSink<R>& sink);
    // quad form of the per-thread kernels (4 lanes per world); null for the loop variant.  They size their own grid.
    void (*step_quad)(cudaStream_t st, int lanes, const PModel<R>& M, const PTask<R>& K, const StepArgs<R>& a);   // lanes per world: 2, 4, 8
    void (*substep_quad)(cudaStream_t st, int lanes, const PModel<R>& M, int n, R* q, R* dq, const R* tau, int lcp_mode, int pgs_iters,
                         const ContactSink<R>& sink);
    void (*coop_table)(const PModel<R>& M, const PTask<R>& K, void* host_out);   // fills coop_table_bytes of per-lane constants
    size_t coop_table_bytes;
    int coop_lanes;   // lanes per world of the cooperative kernels (0: none)
};
#define DARTB_DECLARE_LAUNCHERS(SUFFIX, R) extern const Launchers<R> dartb_launchers_##SUFFIX;
DARTB_DECLARE_LAUNCHERS(hopper_f, float)
DARTB_DECLARE_LAUNCHERS(hopper_d, double)
DARTB_DECLARE_LAUNCHERS(walker_f, float)
DARTB_DECLARE_LAUNCHERS(walker_d, double)
DARTB_DECLARE_LAUNCHERS(cheetah_f, float)
DARTB_DECLARE_LAUNCHERS(cheetah_d, double)
DARTB_DECLARE_LAUNCHERS(snake_f, float)
DARTB_DECLARE_LAUNCHERS(snake_d, double)
DARTB_DECLARE_LAUNCHERS(loop_f, float)
DARTB_DECLARE_LAUNCHERS(loop_d, double)
