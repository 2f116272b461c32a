// inst.cu — one (topology, precision) instantiation of the kernels per translation unit.
// Compiled by build.py with -DINST_TOPO=<TopoX|LOOP> -DINST_REAL=<float|double> -DINST_SUFFIX=<name>.
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
                      int lcp_mode, int pgs_iters, const ContactSink<R_>& sink) {
    k_substep_loop<R_><<<grid, bs, 0, st>>>(M, n, q, dq, tau, fext, lcp_mode, pgs_iters, sink);
}
#else
typedef INST_TOPO T_;
static void l_step(int grid, int bs, size_t shm, cudaStream_t st, const PModel<R_>& M, const PTask<R_>& K, const StepArgs<R_>& a) {
    k_env_step<T_, R_><<<grid, bs, shm, st>>>(M, K, a);
}
static void l_reset(int grid, int bs, size_t shm, cudaStream_t st, const PModel<R_>& M, const PTask<R_>& K, const StepArgs<R_>& a) {
    k_reset<T_, R_><<<grid, bs, shm, st>>>(M, K, a);
}
static void l_substep(int grid, int bs, cudaStream_t st, const PModel<R_>& M, int n, R_* q, R_* dq, const R_* tau, const R_* fext,
                      int lcp_mode, int pgs_iters, const ContactSink<R_>& sink) {
    k_substep<T_, R_><<<grid, bs, 0, st>>>(M, n, q, dq, tau, fext, lcp_mode, pgs_iters, sink);
}
#endif
extern const Launchers<R_> CAT2(dartb_launchers_, INST_SUFFIX) = {l_step, l_reset, l_substep};
