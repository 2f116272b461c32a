// inst.cu — one (topology, precision) instantiation of the kernels per translation unit.
// Compiled by build.py with -DINST_TOPO=<TopoX|LOOP> -DINST_REAL=<float|double> -DINST_SUFFIX=<name>.
#include <cstdlib>
#include <cstdint>
#ifndef DARTB_COOP_TMA_DEFAULT
#define DARTB_COOP_TMA_DEFAULT 0
#endif

#include "kernels.cuh"

#define CAT2_(a, b) a##b
#define CAT2(a, b) CAT2_(a, b)
typedef INST_REAL R_;

#ifdef INST_LOOP
static void l_step(int grid, int bs, size_t shm, cudaStream_t st, const PModel<R_>& M, const PTask<R_>& K, const StepArgs<R_>& a) {
    k_env_step_loop<R_><<<grid, bs, shm, st>>>(M, K, a);
}
static void l_reset(int grid, int bs, size_t shm, cudaStream_t st, const PModel<R_>& M, const PTask<R_>& K, const StepArgs<R_>& a) {
    k_reset_loop<R_><<<grid, bs, shm, st>>>(M, K, a);
}
static void l_substep(int grid, int bs, cudaStream_t st, const PModel<R_>& M, int n, R_* q, R_* dq, const R_* tau, const R_* fext,
                      int lcp_mode, int pgs_iters, const ContactSink<R_>& sink, const R_* wpar) {
    k_substep_loop<R_><<<grid, bs, 0, st>>>(M, n, q, dq, tau, fext, lcp_mode, pgs_iters, sink, wpar);
}
#else
typedef INST_TOPO T_;
static void l_step(int grid, int bs, size_t shm, cudaStream_t st, const PModel<R_>& M, const PTask<R_>& K, const StepArgs<R_>& a) {
    // the fluid-force instantiation exists for capsule-free topologies only (the snake: the one task that has it);
    // dartb.cu::lower_into sends fluid tasks on other topologies to the topology-generic loop kernel
    if constexpr (T_::NS == 0) {
        if (K.fluid_force) { k_env_step<T_, R_, true><<<grid, bs, shm, st>>>(M, K, a); return; }
    }
    k_env_step<T_, R_, false><<<grid, bs, shm, st>>>(M, K, a);
}
static void l_reset(int grid, int bs, size_t shm, cudaStream_t st, const PModel<R_>& M, const PTask<R_>& K, const StepArgs<R_>& a) {
    k_reset<T_, R_><<<grid, bs, shm, st>>>(M, K, a);
}
static void l_substep(int grid, int bs, cudaStream_t st, const PModel<R_>& M, int n, R_* q, R_* dq, const R_* tau, const R_* fext,
                      int lcp_mode, int pgs_iters, const ContactSink<R_>& sink, const R_*) {
    k_substep<T_, R_><<<grid, bs, 0, st>>>(M, n, q, dq, tau, fext, lcp_mode, pgs_iters, sink);
}
// lane-cooperative kernels: 4 warps per block, Coop<T>::WPW worlds per warp
static int coop_warps() {   // warps per block (1, 2 or 4; DARTB_COOP_WARPS overrides the default)
    static int v = 0;
    if (!v) { const char* e = getenv("DARTB_COOP_WARPS"); v = e ? atoi(e) : 4; if (v != 1 && v != 2 && v != 4) v = 4; }
    return v;
}
#define COOP_WARPS coop_warps()
static int coop_tma_mode() {   // DARTB_COOP_TMA: 0 plain loads, 1 TMA-staged state + lane table + obs store, 2 = actions too
    static int v = -1;
    if (v < 0) { const char* e = getenv("DARTB_COOP_TMA"); v = e ? atoi(e) : DARTB_COOP_TMA_DEFAULT; if (v < 0 || v > 2) v = 0; }
    return v;
}
static void l_step_coop(cudaStream_t st, const PModel<R_>& M, const PTask<R_>& K, const StepArgs<R_>& a_in, const void* tab) {
    const int per_block = COOP_WARPS * Coop<T_>::WPW, grid = (a_in.n + per_block - 1) / per_block;
    const CoopLane<T_, R_>* t = (const CoopLane<T_, R_>*)tab;
    StepArgs<R_> a = a_in;
    // TMA staging needs full tiles whose pieces are 16-byte aligned multiples of 16 bytes
    a.tma = 0;
    if (coop_tma_mode() > 0 && a.n_obs_peers == 0 && a.n % per_block == 0 && (per_block * sizeof(R_)) % 16 == 0 && (per_block * K.n_obs * sizeof(float)) % 16 == 0 &&
        ((uintptr_t)a.obs % 16) == 0 && ((size_t)a.n * sizeof(R_)) % 16 == 0) {
        a.tma = 1;
        if (coop_tma_mode() > 1 && (per_block * K.n_act * sizeof(float)) % 16 == 0 && ((uintptr_t)a.action % 16) == 0) a.tma = 2;
    }
    const size_t shm = a.tma ? coop_stage_offset<T_, R_>(COOP_WARPS, K.n_obs) + coop_stage_bytes<T_, R_>(COOP_WARPS, K.n_act)
                             : coop_shared_bytes<T_, R_>(COOP_WARPS, K.n_obs);
    // the fluid-force instantiation exists for capsule-free topologies only (the snake: the one task that has it);
    // dartb.cu::lower_into keeps fluid tasks on other topologies on the per-thread kernels
    if constexpr (T_::NS == 0) {
        if (K.fluid_force) { k_env_step_coop<T_, R_, true><<<grid, COOP_WARPS * 32, shm, st>>>(M, K, a, t); return; }
    }
    k_env_step_coop<T_, R_, false><<<grid, COOP_WARPS * 32, shm, st>>>(M, K, a, t);
}
static void l_substep_coop(cudaStream_t st, const PModel<R_>& M, const void* tab, int n, R_* q, R_* dq, const R_* tau, int lcp_mode,
                           int pgs_iters, const ContactSink<R_>& sink) {
    const int per_block = COOP_WARPS * Coop<T_>::WPW, grid = (n + per_block - 1) / per_block;
    k_substep_coop<T_, R_><<<grid, COOP_WARPS * 32, coop_shared_bytes<T_, R_>(COOP_WARPS, 0), st>>>(M, (const CoopLane<T_, R_>*)tab, n, q, dq, tau,
                                                                                                    lcp_mode, pgs_iters, sink);
}
// group ("quad") forms: 4 warps per block, 32 / lanes worlds per warp
template <int G>
static void launch_quad(cudaStream_t st, const PModel<R_>& M, const PTask<R_>& K, const StepArgs<R_>& a) {
    const int per_block = 4 * (32 / G), grid = (a.n + per_block - 1) / per_block;
    const int stage = K.n_obs > K.n_act ? K.n_obs : K.n_act;
    const size_t shm = (size_t)per_block * stage * sizeof(float);
    if constexpr (T_::NS == 0) {
        if (K.fluid_force) { k_env_step_quad<T_, R_, true, G><<<grid, 128, shm, st>>>(M, K, a); return; }
    }
    k_env_step_quad<T_, R_, false, G><<<grid, 128, shm, st>>>(M, K, a);
}
static void l_step_quad(cudaStream_t st, int lanes, const PModel<R_>& M, const PTask<R_>& K, const StepArgs<R_>& a) {
    (void)lanes;   // 2 and 8 lanes per world were measured and lost to 4 almost everywhere (profiles/r2_experiments.md section 6)
    launch_quad<4>(st, M, K, a);
}
static void l_substep_quad(cudaStream_t st, int lanes, const PModel<R_>& M, int n, R_* q, R_* dq, const R_* tau, int lcp_mode, int pgs_iters,
                           const ContactSink<R_>& sink) {
    (void)lanes;
    const int per_block = 4 * 8, grid = (n + per_block - 1) / per_block;
    k_substep_quad<T_, R_, 4><<<grid, 128, 0, st>>>(M, n, q, dq, tau, lcp_mode, pgs_iters, sink);
}
static void l_coop_table(const PModel<R_>& M, const PTask<R_>& K, void* out) { coop_build_table<T_, R_>(M, &K, (CoopLane<T_, R_>*)out); }
#endif
#ifdef INST_LOOP
extern const Launchers<R_> CAT2(dartb_launchers_, INST_SUFFIX) = {l_step, l_reset, l_substep, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
#else
extern const Launchers<R_> CAT2(dartb_launchers_, INST_SUFFIX) = {l_step, l_reset, l_substep, l_step_coop, l_substep_coop, l_step_quad, l_substep_quad, l_coop_table, sizeof(CoopLane<T_, R_>) * Coop<T_>::G, Coop<T_>::G};
#endif
