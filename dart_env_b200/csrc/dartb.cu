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
#include "kernels.cuh"

// ------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));   \
    } while (0)

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

// ------------------------------------------------------------------------ developer overrides (environment)
// read once, through a C++11 function-local static (thread-safe initialisation): handles may be created from
// several host threads
struct EnvCfg {
    int variant, block, wpw, zerocopy; long coop_max, quad_max;
    EnvCfg() {
        auto geti = [](const char* k, long d) { const char* v = getenv(k); return v ? atol(v) : d; };
        variant = (int)geti("DARTB_VARIANT", -1); block = (int)geti("DARTB_BLOCK", 0); wpw = (int)geti("DARTB_WPW", 0);
        zerocopy = (int)geti("DARTB_ZEROCOPY", 1); coop_max = geti("DARTB_COOP_MAX_WORLDS", -1); quad_max = geti("DARTB_QUAD_MAX_WORLDS", -1);
    }
};
static const EnvCfg& envcfg() { static const EnvCfg c; return c; }

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
    uint64_t* seeds = nullptr;                // [n] per-world seeds (dartb_seed_worlds) or null
    float* obs_peer[DARTB_MAX_PEERS] = {};    // dartb_set_obs_peers: fused observation all-gather targets
    int n_obs_peers = 0; long long obs_peer_off = 0;
    void* aux = nullptr;                      // [3][n] of Real: per-world task state (reacher target), or null
    void* wpar = nullptr;                     // [4 nb + ns][n] of Real: per-world dynamics parameters (dartb_set_body_params), or null
    std::vector<double> body_mass, body_mu;   // [n][n_bodies] as given to dartb_set_body_params (empty: the model's)
    float rand_mass = 0, rand_mu = 0;         // DARTB_OPT_RANDOMIZE_*: per-reset redraw of the table's mass / friction rows
    std::string signature;                    // topology signature of the lowered model
    void* scratch = nullptr;                  // [n * nd | n * nbd*3] of Real: tau / fext precision conversion
    uint32_t* episode = nullptr; int32_t* elapsed = nullptr; uint8_t* truncated = nullptr;
    uint64_t* hint = nullptr;                 // LCP warm-start sets, see planar_kernels.cuh::substep
    int32_t* ccount = nullptr; int32_t* cbody = nullptr; float* cdata = nullptr;
    int lcp_mode = 0, pgs_iters = 30, max_episode_steps = 0;
    int variant_request = -1;                 // -1 auto (DARTB_VARIANT env, else by batch size), 0, 1, 2
    int variant = 0;                          // 0 = unrolled static topology (one world per thread), 1 = loop / generic
                                              // topology, 2 = lane-cooperative (8/16 lanes per world, planar_coop.cuh)
    int wpw_request = 0;                      // worlds per warp of k_env_step, 0 = auto (wpw_for)
    void* coop_tab = nullptr; size_t coop_tab_bytes = 0; bool coop_tab_dirty = true;   // per-lane constants of the cooperative kernels
    int64_t launches = 0;
    bool contacts = false;                    // DARTB_OPT_CONTACTS: the fused step records the last sub-step's contacts
    bool contacts_valid = false;              // the contact buffers describe the last stepping call
    // host-facing step (dartb_step_host): pinned staging + device mirrors, one stream
    float* h_stage = nullptr; float* d_stage = nullptr; float* h_stage_dev = nullptr; size_t stage_floats = 0;
    const void* zc_h[3] = {nullptr, nullptr, nullptr}; void* zc_d[3] = {nullptr, nullptr, nullptr};   // zero-copy output aliases
    struct HostRange { const char* h; size_t bytes; char* d; };
    std::vector<HostRange> host_ranges;       // dartb_register_host: page-locked ranges whose device alias is known
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
    const int forced_variant = envcfg().variant;
    const int want = e->variant_request >= 0 ? e->variant_request : forced_variant;
    // the compiled topologies carry the fluid force only where the reference has it (capsule-free: the snake); a fluid
    // task on a skeleton with capsules runs on the topology-generic loop kernel
    const bool coop_ok = true;
    // per-world dynamics parameters are read by the loop kernels only (the compiled topologies read one __grid_constant__ model)
    const bool per_world = !e->body_mass.empty() || !e->body_mu.empty() || e->rand_mass > 0 || e->rand_mu > 0;
    if (topo < 0 || res.m.any_coulomb || want == 1 || per_world || res.t.kind != DARTB_TASK_LOCOMOTION || (res.t.fluid_force && res.m.ns > 0)) e->variant = 1;
    else if (want == 2 && coop_ok) e->variant = 2;
    else if (want == 3) e->variant = 3;
    else if (!coop_ok) e->variant = 0;
    else if (want == 0) e->variant = 0;
    else {
        // auto, by batch size.  Small batches: the lane-cooperative form (8 / 16 lanes per world: ~2.5x the instructions per
        // world for ~5x less latency per DART step).  Mid-size batches: the quad form (4 lanes per world share the constraint
        // phase: 4x the warps of the per-thread form, a warp runs the union of 8 worlds' branches instead of 32).  Large
        // batches: one world per thread.  The quad form's PGS gathers A onto every lane, so PGS mode skips it.
        // Crossovers measured on B200 (gpurun_out/r2g_sweep.log, r2h_sweep.log, r2i_xover.log; us per env step):
        //   Hopper      2048: coop 32, quad 34     4096: quad 35.4, coop 38.8, static 47.7   8192: quad 41.9, static 55    16384: static 51.6, quad 75
        //   Walker2d    4096: quad 85, coop 94, static 111     8192: quad 113, static 130, coop 178      16384: static 149, quad 213
        //   HalfCheetah 4096: coop 168, quad 284, static 344   8192: coop 304, quad 372, static 420      16384: static 486, quad 509
        //   Snake7Link  2048: quad 27.5, coop 34    4096: quad 31.4, static 41.6, coop 60     8192: quad 36.3     32768: static 55.7, quad 102
        long lim_coop = envcfg().coop_max, lim_quad = envcfg().quad_max;
        //   (crossovers, r2i_xover.log)  Hopper 2048: coop 32.5, quad 34.2; 3072: quad 34.5, coop 36.7; 10240: static 50.3, quad 65.1
        //   Walker2d 2048: coop 54.5, quad 73.5; 3072: 80.1 / 79.6; 10240: static 130, quad 140   HalfCheetah 10240: coop 369, static 419;
        //   12288: 438 / 441   Snake7Link 1024: quad 27.2, coop 28.9; 12288: static 45.2, quad 53.9
        //   after the branch-free solvers (r2p_sweep.log; coop / quad / static):  Hopper 1536: 30.0 / 30.4; 2048: 30.9 / 30.7; 2560: 33.3 / 30.7;
        //   8192: - / 37.5 / 41.0; 10240: - / 59.4 / 41.6   Walker2d 2048: 51.9 / 63.0; 2560: 69.3 / 66.8; 3072: 76.3 / 69.4; 8192: - / 102 / 122;
        //   10240: - / 129 / 115   HalfCheetah 8192: 286 / 346 / 399; 10240: 349 / 371 / 394; 12288: 414 / 401 / 416; 16384: 549 / 464 / 458
        //   Snake7Link 512: 27.0 / 24.9; 768: 27.1 / 25.3; 8192: - / 35.0 / 41.5; 12288: - / 51.9 / 42.2
        //   (the quad form steps up past 8880 worlds = one wave of 148 x 60 worlds, the 16-lane cooperative form past 2368)
        if (lim_coop < 0) lim_coop = topo == TOPO_HOPPER ? 2048 : (topo == TOPO_WALKER ? 2368 : (topo == TOPO_CHEETAH ? 11264 : 448));
        if (lim_quad < 0) lim_quad = topo == TOPO_HOPPER ? 8880 : (topo == TOPO_WALKER ? 8880 : (topo == TOPO_CHEETAH ? 13312 : 8880));
        if (e->lcp_mode == 1) lim_quad = 0;
        e->variant = (e->n <= lim_coop) ? 2 : ((e->n <= lim_quad) ? 3 : 0);
    }
    e->topo = topo;
    e->signature = res.signature;
    e->coop_tab_dirty = true;
    e->md = res.m; e->td = res.t;
    lower::convert(res.m, e->mf);
    lower::convert(res.t, e->tf);
    e->nd = res.m.nb;
    e->max_contacts = res.max_contacts;
    e->n_orig_bodies = e->model.n_bodies;
    const char* plane = std::fabs(res.m.en[2]) > 0.5 ? "planar-xy" : (std::fabs(res.m.en[1]) > 0.5 ? "planar-zx" : "planar-yz");
    e->kernel_name = std::string(plane) + (e->variant == 1 ? std::string("/loop:generic") : std::string(e->variant == 2 ? "/coop:" : (e->variant == 3 ? "/quad:" : "/static:")) + topo_name(topo)) +
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
    const int forced = envcfg().block;
    if (forced >= 32 && forced <= 128 && forced % 32 == 0) return forced;
    return n <= 148 * 32 * 4 ? 32 : (n <= 148 * 64 * 8 ? 64 : 128);
}

// Worlds per warp of k_env_step.  The stepper is latency-bound per warp (ncu: IPC ~0.1 with one warp per SM)
// and a warp executes the UNION of its worlds' branches (contact sets, LCP size classes), so while the
// grid does not yet fill the 148 x 4 warp schedulers a batch is spread over MORE, narrower warps: fewer
// worlds per warp = a shorter union path, and the idle schedulers run them concurrently.
// DARTB_WPW / DARTB_OPT_WORLDS_PER_WARP override.
static int wpw_for(const dartb_engine* e) {
    const int forced = envcfg().wpw;
    int w = e->wpw_request > 0 ? e->wpw_request : forced;
    if (w >= 1 && w <= 32) return w;
    // one warp per scheduler: measured (gpurun_out/sweep_wpw.log) a second narrower warp per scheduler
    // costs more than the shorter union path saves (Walker2d 16384 worlds: 224 us at 16/warp vs 173 at 32)
    const int slots = 148 * 4;
    w = 32;
    while (w > 8 && (e->n + w / 2 - 1) / (w / 2) <= slots) w /= 2;
    return w;
}

template <typename R>
static StepArgs<R> make_args(dartb_engine* e) {
    StepArgs<R> a;
    std::memset(&a, 0, sizeof a);
    a.n = e->n; a.q = (R*)e->q; a.dq = (R*)e->dq; a.episode = e->episode; a.elapsed = e->elapsed;
    a.truncated = e->truncated;
    a.hint = e->hint;
    a.aux = (R*)e->aux;
    a.wpar = (R*)e->wpar;
    a.rand_mass = e->rand_mass; a.rand_mu = e->rand_mu;
    a.lcp_mode = e->lcp_mode; a.pgs_iters = e->pgs_iters; a.max_episode_steps = e->max_episode_steps;
    a.seed = e->seed; a.world_offset = e->world_offset; a.seeds = e->seeds;
    a.n_obs_peers = e->n_obs_peers; a.obs_peer_off = e->obs_peer_off;
    for (int p = 0; p < e->n_obs_peers; p++) a.obs_peer[p] = e->obs_peer[p];
    if (e->contacts) { a.sink.count = e->ccount; a.sink.body = e->cbody; a.sink.data = e->cdata; }
    a.sink.maxc = e->max_contacts;
    return a;
}

template <typename R> struct LTab;
template <> struct LTab<float> {
    static const Launchers<float>& get(const dartb_engine* e) {
        if (e->variant == 1) return dartb_launchers_loop_f;
        switch (e->topo) {
            case TOPO_HOPPER: return dartb_launchers_hopper_f;
            case TOPO_WALKER: return dartb_launchers_walker_f;
            case TOPO_CHEETAH: return dartb_launchers_cheetah_f;
            default: return dartb_launchers_snake_f;
        }
    }
};
template <> struct LTab<double> {
    static const Launchers<double>& get(const dartb_engine* e) {
        if (e->variant == 1) return dartb_launchers_loop_d;
        switch (e->topo) {
            case TOPO_HOPPER: return dartb_launchers_hopper_d;
            case TOPO_WALKER: return dartb_launchers_walker_d;
            case TOPO_CHEETAH: return dartb_launchers_cheetah_d;
            default: return dartb_launchers_snake_d;
        }
    }
};

// (re)upload the per-lane constant table of the cooperative kernels after the model was (re)lowered
template <typename R>
static int coop_table_sync(dartb_engine* e, cudaStream_t st) {
    if (!e->coop_tab_dirty) return 0;
    const Launchers<R>& L = LTab<R>::get(e);
    if (!L.coop_table) return fail("this topology has no cooperative kernel");
    if (e->coop_tab_bytes < L.coop_table_bytes) {
        if (e->coop_tab) cudaFree(e->coop_tab);
        e->coop_tab = nullptr; e->coop_tab_bytes = 0;
        CK(cudaMalloc(&e->coop_tab, L.coop_table_bytes));
        e->coop_tab_bytes = L.coop_table_bytes;
    }
    std::vector<unsigned char> host(L.coop_table_bytes);
    L.coop_table(Sel<R>::m(e), Sel<R>::t(e), host.data());
    CK(cudaStreamSynchronize(st));   // a previous launch may still be reading the old table
    CK(cudaMemcpy(e->coop_tab, host.data(), host.size(), cudaMemcpyHostToDevice));
    e->coop_tab_dirty = false;
    return 0;
}

template <typename R>
static int launch_step(dartb_engine* e, const float* action, float* obs, float* reward, uint8_t* done, int auto_reset,
                       cudaStream_t st, double* reward64 = nullptr, uint8_t* truncated = nullptr) {
    StepArgs<R> a = make_args<R>(e);
    a.action = action; a.obs = obs; a.reward = reward; a.done = done; a.auto_reset = auto_reset;
    a.reward64 = reward64;
    if (truncated) a.truncated = truncated;
    e->contacts_valid = e->contacts;
    if (e->variant == 2) {
        if (coop_table_sync<R>(e, st)) return 1;
        LTab<R>::get(e).step_coop(st, Sel<R>::m(e), Sel<R>::t(e), a, e->coop_tab);
        e->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    if (e->variant == 3) {
        LTab<R>::get(e).step_quad(st, 4, Sel<R>::m(e), Sel<R>::t(e), a);
        e->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    a.wpw = wpw_for(e);
    const int warps = (e->n + a.wpw - 1) / a.wpw;
    const int bs = block_for(warps * 32), grid = (warps + bs / 32 - 1) / (bs / 32);
    const PTask<R>& K = Sel<R>::t(e);
    const int stage = K.n_obs > K.n_act ? K.n_obs : K.n_act;
    const size_t shm = (size_t)(bs / 32) * 32 * stage * sizeof(float);
    LTab<R>::get(e).step(grid, bs, shm, st, Sel<R>::m(e), K, a);
    e->launches++;
    CK(cudaGetLastError());
    return 0;
}
template <typename R>
static int launch_reset(dartb_engine* e, const uint8_t* mask, float* obs, cudaStream_t st) {
    StepArgs<R> a = make_args<R>(e);
    a.mask = mask; a.obs = obs;
    a.sink.count = e->ccount;   // a reset world has no contacts (world.reset() clears collision_result)
    const int bs = block_for(e->n), grid = (e->n + bs - 1) / bs;
    const PTask<R>& K = Sel<R>::t(e);
    const size_t shm = (size_t)(bs / 32) * 32 * K.n_obs * sizeof(float);
    LTab<R>::get(e).reset(grid, bs, shm, st, Sel<R>::m(e), K, a);
    e->launches++;
    CK(cudaGetLastError());
    return 0;
}
template <typename R>
static int launch_substep(dartb_engine* e, const R* tau, const R* fext, cudaStream_t st) {
    ContactSink<R> sink;
    sink.count = e->ccount; sink.body = e->cbody; sink.data = e->cdata; sink.maxc = e->max_contacts;
    const int bs = block_for(e->n), grid = (e->n + bs - 1) / bs;
    e->contacts_valid = true;   // the literal World.step() always refreshes collision_result
    if (e->variant == 2 && !fext) {
        if (coop_table_sync<R>(e, st)) return 1;
        LTab<R>::get(e).substep_coop(st, Sel<R>::m(e), e->coop_tab, e->n, (R*)e->q, (R*)e->dq, tau, e->lcp_mode, e->pgs_iters, sink);
    }
    else if (e->variant == 3 && !fext)
        LTab<R>::get(e).substep_quad(st, 4, Sel<R>::m(e), e->n, (R*)e->q, (R*)e->dq, tau, e->lcp_mode, e->pgs_iters, sink);
    else
        LTab<R>::get(e).substep(grid, bs, st, Sel<R>::m(e), e->n, (R*)e->q, (R*)e->dq, tau, fext, e->lcp_mode, e->pgs_iters, sink, (const R*)e->wpar);
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
    const size_t sc = (size_t)n * (size_t)(e->nd + e->n_orig_bodies * 3);   // tau region, then fext region
    cudaError_t err = cudaSuccess;
    auto A = [&](void** p, size_t bytes) { if (err == cudaSuccess) { err = cudaMalloc(p, bytes); if (err == cudaSuccess) err = cudaMemset(*p, 0, bytes); } };
    A(&e->q, rs * n * e->nd); A(&e->dq, rs * n * e->nd); A(&e->scratch, rs * sc);
    A((void**)&e->episode, 4 * (size_t)n); A((void**)&e->elapsed, 4 * (size_t)n); A((void**)&e->truncated, (size_t)n);
    A((void**)&e->hint, 8 * (size_t)n);
    if (err == cudaSuccess) err = cudaMemset(e->hint, 0xFF, 8 * (size_t)n);
    if (task->kind == DARTB_TASK_REACHER2D) {
        A(&e->aux, rs * 3 * (size_t)n);
        // self.target = [0.1, 0.01, -0.1] until the first reset draws one (reacher2d.py:7)
        const double t0[3] = {0.1, 0.01, -0.1};
        std::vector<double> hd(3 * (size_t)n); std::vector<float> hf(3 * (size_t)n);
        for (int c = 0; c < 3; c++) for (int w = 0; w < n; w++) { hd[(size_t)c * n + w] = t0[c]; hf[(size_t)c * n + w] = (float)t0[c]; }
        if (err == cudaSuccess) err = cudaMemcpy(e->aux, f64 ? (const void*)hd.data() : (const void*)hf.data(), rs * 3 * (size_t)n, cudaMemcpyHostToDevice);
    }
    A((void**)&e->ccount, 4 * (size_t)n); A((void**)&e->cbody, 4 * (size_t)n * e->max_contacts);
    A((void**)&e->cdata, 4 * (size_t)n * e->max_contacts * 10);
    if (err != cudaSuccess) { dartb_destroy(e); return fail(std::string("cudaMalloc: ") + cudaGetErrorString(err)); }
    *out = e;
    // the cooperative kernels' lane table is uploaded now, so that no launch ever synchronises
    if (e->variant == 2 && (f64 ? coop_table_sync<double>(e, 0) : coop_table_sync<float>(e, 0))) { dartb_destroy(e); *out = nullptr; return 1; }
    // initial state = q_init / dq_init (the pydart World constructor resets the world)
    {
        const int nd = e->nd;
        if (f64) {
            std::vector<double> hq((size_t)n * nd), hv((size_t)n * nd);
            for (int d = 0; d < nd; d++) for (int w = 0; w < n; w++) { hq[(size_t)d * n + w] = e->md.qinit[d]; hv[(size_t)d * n + w] = e->md.dqinit[d]; }
            err = cudaMemcpy(e->q, hq.data(), hq.size() * 8, cudaMemcpyHostToDevice);
            if (err == cudaSuccess) err = cudaMemcpy(e->dq, hv.data(), hv.size() * 8, cudaMemcpyHostToDevice);
        } else {
            std::vector<float> hq((size_t)n * nd), hv((size_t)n * nd);
            for (int d = 0; d < nd; d++) for (int w = 0; w < n; w++) { hq[(size_t)d * n + w] = e->mf.qinit[d]; hv[(size_t)d * n + w] = e->mf.dqinit[d]; }
            err = cudaMemcpy(e->q, hq.data(), hq.size() * 4, cudaMemcpyHostToDevice);
            if (err == cudaSuccess) err = cudaMemcpy(e->dq, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice);
        }
        if (err != cudaSuccess) { dartb_destroy(e); *out = nullptr; return fail(std::string("cudaMemcpy (initial state): ") + cudaGetErrorString(err)); }
    }
    return 0;
}

template <typename S>
static int set_state_impl(dartb_handle_t e, const S* q, const S* dq, cudaStream_t st) {
    const int tot = e->n * e->nd, bs = 256, grid = (tot + bs - 1) / bs;
    CK(cudaMemsetAsync(e->hint, 0xFF, 8 * (size_t)e->n, st));  // a new state invalidates the LCP warm start
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
    cudaFree(e->q); cudaFree(e->dq); cudaFree(e->aux); cudaFree(e->wpar); cudaFree(e->seeds); cudaFree(e->scratch); cudaFree(e->episode); cudaFree(e->elapsed);
    cudaFree(e->truncated); cudaFree(e->hint); cudaFree(e->ccount); cudaFree(e->cbody); cudaFree(e->cdata);
    if (e->coop_tab) cudaFree(e->coop_tab);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    if (e->d_stage) cudaFree(e->d_stage);
    delete e;
    return 0;
}

// options that change the lowered model: lower again and refresh the lane table (the only synchronising path)
static int upload_body_params(dartb_engine* e);
static int relower(dartb_engine* e) {
    if (lower_into(e)) return 1;
    if (upload_body_params(e)) return 1;
    if (e->variant != 2) return 0;
    DeviceGuard g(e->device);
    return e->f64 ? coop_table_sync<double>(e, 0) : coop_table_sync<float>(e, 0);
}

// Per-world planar parameters from per-world DART bodynode masses / friction coefficients: every world's model is lowered
// again with its own values (the weld merge mixes the masses of a group into its mass, COM and izz), so the kernel's numbers
// are by construction what a single-world engine built from that model would use.
static int upload_body_params(dartb_engine* e) {
    const bool pm = !e->body_mass.empty(), pf = !e->body_mu.empty();
    DeviceGuard g(e->device);
    if (!pm && !pf && !(e->rand_mass > 0 || e->rand_mu > 0)) {
        if (e->wpar) { cudaDeviceSynchronize(); cudaFree(e->wpar); e->wpar = nullptr; }
        return 0;
    }
    std::vector<double> host;
    std::string why = lower::body_param_table(e->model, e->task, e->n, pm ? e->body_mass.data() : nullptr, pf ? e->body_mu.data() : nullptr,
                                              e->signature, e->md.nb, e->md.ns, host);
    if (!why.empty()) return fail("dartb_set_body_params: " + why);
    const size_t esz = e->f64 ? sizeof(double) : sizeof(float);
    cudaDeviceSynchronize();   // a stepping call in flight may still read the old table
    if (!e->wpar) CK(cudaMalloc(&e->wpar, host.size() * esz));
    if (e->f64) CK(cudaMemcpy(e->wpar, host.data(), host.size() * esz, cudaMemcpyHostToDevice));
    else {
        std::vector<float> hf(host.begin(), host.end());
        CK(cudaMemcpy(e->wpar, hf.data(), hf.size() * esz, cudaMemcpyHostToDevice));
    }
    return 0;
}

int dartb_set_body_params(dartb_handle_t e, const double* h_mass, const double* h_friction) {
    if (!e) return fail("null handle");
    const size_t k = (size_t)e->n * e->model.n_bodies;
    for (size_t i = 0; i < k; i++) {
        if (h_mass && !(h_mass[i] >= 0.0 && std::isfinite(h_mass[i]))) return fail("dartb_set_body_params: masses must be finite and >= 0");
        if (h_friction && !(h_friction[i] >= 0.0 && std::isfinite(h_friction[i]))) return fail("dartb_set_body_params: friction coefficients must be finite and >= 0");
    }
    std::vector<double> old_m, old_f;
    old_m.swap(e->body_mass); old_f.swap(e->body_mu);
    if (h_mass) e->body_mass.assign(h_mass, h_mass + k);
    if (h_friction) e->body_mu.assign(h_friction, h_friction + k);
    if (relower(e)) {   // leave the engine as it was
        e->body_mass.swap(old_m); e->body_mu.swap(old_f);
        std::string keep = g_err;
        relower(e);
        g_err = keep;
        return 1;
    }
    return 0;
}

int dartb_get_body_table(dartb_handle_t e, double* h_out, int32_t* n_rows) {
    if (!e) return fail("null handle");
    if (!e->wpar) return fail("dartb_get_body_table: no per-world parameters are set");
    if (n_rows) *n_rows = 4 * e->md.nb + e->md.ns;
    if (!h_out) return 0;
    DeviceGuard g(e->device);
    const size_t k = (size_t)(4 * e->md.nb + e->md.ns) * e->n;
    CK(cudaDeviceSynchronize());
    if (e->f64) { CK(cudaMemcpy(h_out, e->wpar, k * sizeof(double), cudaMemcpyDeviceToHost)); return 0; }
    std::vector<float> hf(k);
    CK(cudaMemcpy(hf.data(), e->wpar, k * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < k; i++) h_out[i] = hf[i];
    return 0;
}

int dartb_set_option(dartb_handle_t e, int32_t key, double value) {
    if (!e) return fail("null handle");
    switch (key) {
        case DARTB_OPT_LCP_MODE:
            if (value != 0 && value != 1) return fail("lcp mode must be 0 (exact) or 1 (PGS)");
            e->lcp_mode = (int)value;
            return e->variant_request < 0 ? relower(e) : 0;   // the automatic kernel choice depends on the solver
        case DARTB_OPT_PGS_ITERS:
            if (value < 1 || value > 10000) return fail("bad PGS iteration count");
            e->pgs_iters = (int)value; return 0;
        case DARTB_OPT_FRICTION_ALL:
            for (int i = 0; i < e->model.n_bodies; i++) e->model.bodies[i].friction_coeff = value;
            return relower(e);
        case DARTB_OPT_KERNEL_VARIANT:
            if (!(value == -1 || (value >= 0 && value <= 3 && value == (int)value)))
                return fail("kernel variant must be -1 (auto), 0 (one world per thread), 1 (loop), 2 (lane-cooperative) or 3 (quad: 4 lanes per world)");
            e->variant_request = (int)value;
            return relower(e);
        case DARTB_OPT_WORLDS_PER_WARP:
            if (value < 0 || value > 32) return fail("worlds per warp must be 0 (auto) or 1..32");
            e->wpw_request = (int)value; return 0;
        case DARTB_OPT_CONTACTS:
            if (value != 0 && value != 1) return fail("contacts option must be 0 or 1");
            e->contacts = value != 0; return 0;
        case DARTB_OPT_RANDOMIZE_MASS:
        case DARTB_OPT_RANDOMIZE_FRICTION: {
            if (!(value >= 0) || !std::isfinite(value)) return fail("randomisation half range must be finite and >= 0");
            if (value > 0) {
                // one draw per planar body stands for one bodynode: no weld merges (their mass / COM / izz mix on the host)
                bool plain = e->md.nbd == e->md.nb;
                for (int i = 0; plain && i < e->md.nbd; i++) plain = e->md.dgroup[i] == i;
                if (!plain) return fail("per-reset randomisation needs a skeleton without welded bodynodes; use dartb_set_body_params");
            }
            float& slot = key == DARTB_OPT_RANDOMIZE_MASS ? e->rand_mass : e->rand_mu;
            const float old = slot;
            slot = (float)value;
            if (relower(e)) { slot = old; std::string keep = g_err; relower(e); g_err = keep; return 1; }
            return 0;
        }
        case DARTB_OPT_MAX_EPISODE_STEPS:
            if (value < 0) return fail("bad max_episode_steps");
            e->max_episode_steps = (int)value; return 0;
    }
    return fail("unknown option key");
}

int dartb_seed(dartb_handle_t e, uint64_t seed) {
    if (!e) return fail("null handle");
    e->seed = seed;   // a kernel argument: takes effect with the next launch, no synchronisation
    if (e->seeds) {   // back to one key for the whole batch
        DeviceGuard g(e->device);
        CK(cudaDeviceSynchronize());
        cudaFree(e->seeds); e->seeds = nullptr;
    }
    return 0;
}
int dartb_seed_worlds(dartb_handle_t e, const uint64_t* h_seeds) {
    if (!e || !h_seeds) return fail("null argument");
    DeviceGuard g(e->device);
    if (!e->seeds) CK(cudaMalloc((void**)&e->seeds, 8 * (size_t)e->n));
    CK(cudaMemcpy(e->seeds, h_seeds, 8 * (size_t)e->n, cudaMemcpyHostToDevice));   // synchronous: seeding is not a hot path
    return 0;
}

int dartb_reset(dartb_handle_t e, const uint8_t* d_mask, float* d_obs, void* stream) {
    if (!e) return fail("null handle");
    if (e->task.n_obs == 0 && d_obs) return fail("physics-only handle has no observation layout");
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
    if (e->task.n_obs == 0) return fail("physics-only handle: no task layer configured (use dartb_substep)");
    DeviceGuard g(e->device);
    return e->f64 ? launch_step<double>(e, d_action, d_obs, d_reward, d_done, auto_reset, (cudaStream_t)stream)
                  : launch_step<float>(e, d_action, d_obs, d_reward, d_done, auto_reset, (cudaStream_t)stream);
}

// device-visible alias of a page-locked host pointer (UVA: normally the same value), or null if the
// memory is pageable.  Queries the driver: several microseconds per call (measured: five of them per step cost
// more than the step kernel), so the per-step paths look registered ranges up first (dartb_register_host).
static void* mapped_alias(const void* h) {
    cudaPointerAttributes pa;
    void* d = nullptr;
    if (cudaPointerGetAttributes(&pa, h) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
        cudaHostGetDevicePointer(&d, const_cast<void*>(h), 0) == cudaSuccess)
        return d;
    cudaGetLastError();  // a pageable pointer may leave cudaErrorInvalidValue behind on old drivers
    return nullptr;
}

static void* alias_of(const dartb_engine* e, const void* h) {
    const char* p = (const char*)h;
    for (const auto& r : e->host_ranges)
        if (p >= r.h && p < r.h + r.bytes) return r.d + (p - r.h);
    return mapped_alias(h);
}

// the pinned (host-mapped) staging block of the host-facing step and its device mirror
static int ensure_stage(dartb_engine* e, size_t need) {
    if (e->stage_floats >= need) return 0;
    if (e->h_stage) cudaFreeHost(e->h_stage);
    if (e->d_stage) cudaFree(e->d_stage);
    e->h_stage = nullptr; e->d_stage = nullptr; e->h_stage_dev = nullptr; e->stage_floats = 0;
    CK(cudaHostAlloc((void**)&e->h_stage, need * 4, cudaHostAllocMapped));
    CK(cudaMalloc((void**)&e->d_stage, need * 4));
    void* alias = nullptr;
    if (cudaHostGetDevicePointer(&alias, e->h_stage, 0) == cudaSuccess) e->h_stage_dev = (float*)alias;
    cudaGetLastError();
    e->stage_floats = need;
    return 0;
}

int dartb_register_host(dartb_handle_t e, const void* h_ptr, size_t bytes) {
    if (!e || !h_ptr || !bytes) return fail("null argument");
    DeviceGuard g(e->device);
    void* d = mapped_alias(h_ptr);
    if (!d) return fail("dartb_register_host: the range is not page-locked, device-mapped host memory");
    for (auto& r : e->host_ranges) if (r.h == (const char*)h_ptr) { r.bytes = bytes; r.d = (char*)d; return 0; }
    e->host_ranges.push_back({(const char*)h_ptr, bytes, (char*)d});
    return 0;
}
int dartb_unregister_host(dartb_handle_t e, const void* h_ptr) {
    if (!e) return fail("null handle");
    for (size_t i = 0; i < e->host_ranges.size(); i++)
        if (e->host_ranges[i].h == (const char*)h_ptr) { e->host_ranges.erase(e->host_ranges.begin() + i); return 0; }
    return fail("dartb_unregister_host: unknown range");
}

int dartb_step_host(dartb_handle_t e, const float* h_action, float* h_obs, float* h_reward, uint8_t* h_done,
                    int32_t auto_reset, void* stream) {
    if (!e) return fail("null handle");
    if (!h_action || !h_obs || !h_reward || !h_done) return fail("null host pointer");
    if (e->task.n_obs == 0) return fail("physics-only handle: no task layer configured (use dartb_substep)");
    DeviceGuard g(e->device);
    const int n = e->n, na = e->task.n_act, no = e->task.n_obs;
    // layout of the staging block (floats): [action n*na | obs n*no | reward n | done n bytes]
    const size_t fa = (size_t)n * na, fo = (size_t)n * no, fr = (size_t)n, fd = ((size_t)n + 3) / 4;
    const size_t need = fa + fo + fr + fd;
    if (ensure_stage(e, need)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (envcfg().zerocopy && e->h_stage_dev) {
        // Zero-copy: the step kernel reads the actions from, and writes obs / reward / done to, page-locked
        // host memory over PCIe itself (coalesced 128 B lines through its shared-memory staging), so the
        // whole host step is ONE launch + ONE sync: no memcpy nodes (each costs ~5 us of latency, more than
        // moving these ~250 KB does).  Pageable caller buffers go through the pinned staging block.
        const float* a_dev = (const float*)alias_of(e, h_action);
        if (!a_dev) { std::memcpy(e->h_stage, h_action, fa * 4); a_dev = e->h_stage_dev; }
        // (the caller's output buffers are normally the same page-locked arrays every step: look them up once)
        if (h_obs != e->zc_h[0] || h_reward != e->zc_h[1] || h_done != e->zc_h[2]) {
            e->zc_h[0] = h_obs; e->zc_h[1] = h_reward; e->zc_h[2] = h_done;
            e->zc_d[0] = alias_of(e, h_obs); e->zc_d[1] = alias_of(e, h_reward); e->zc_d[2] = alias_of(e, h_done);
        }
        float* o_dev = (float*)e->zc_d[0];
        float* r_dev = (float*)e->zc_d[1];
        uint8_t* d_dev = (uint8_t*)e->zc_d[2];
        const bool direct = o_dev && r_dev && d_dev;
        if (!direct) { o_dev = e->h_stage_dev + fa; r_dev = o_dev + fo; d_dev = (uint8_t*)(r_dev + fr); }
        int rc = e->f64 ? launch_step<double>(e, a_dev, o_dev, r_dev, d_dev, auto_reset, st)
                        : launch_step<float>(e, a_dev, o_dev, r_dev, d_dev, auto_reset, st);
        if (rc) return rc;
        CK(cudaStreamSynchronize(st));
        if (!direct) {
            std::memcpy(h_obs, e->h_stage + fa, fo * 4);
            std::memcpy(h_reward, e->h_stage + fa + fo, fr * 4);
            std::memcpy(h_done, e->h_stage + fa + fo + fr, (size_t)n);
        }
        return 0;
    }
    std::memcpy(e->h_stage, h_action, fa * 4);
    CK(cudaMemcpyAsync(e->d_stage, e->h_stage, fa * 4, cudaMemcpyHostToDevice, st));
    float* d_obs = e->d_stage + fa;
    float* d_rew = d_obs + fo;
    uint8_t* d_done = (uint8_t*)(d_rew + fr);
    int rc = e->f64 ? launch_step<double>(e, e->d_stage, d_obs, d_rew, d_done, auto_reset, st)
                    : launch_step<float>(e, e->d_stage, d_obs, d_rew, d_done, auto_reset, st);
    if (rc) return rc;
    // results: DMA straight into the caller's buffers when they are page-locked (the DartEnv wrapper
    // allocates its output arrays pinned), else one D2H into the pinned staging block + memcpy
    const bool direct = alias_of(e, h_obs) && alias_of(e, h_reward) && alias_of(e, h_done);
    if (direct) {
        CK(cudaMemcpyAsync(h_obs, d_obs, fo * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_reward, d_rew, fr * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_done, d_done, (size_t)n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return 0;
    }
    CK(cudaMemcpyAsync(e->h_stage + fa, d_obs, (fo + fr + fd) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::memcpy(h_obs, e->h_stage + fa, fo * 4);
    std::memcpy(h_reward, e->h_stage + fa + fo, fr * 4);
    std::memcpy(h_done, e->h_stage + fa + fo + fr, (size_t)n);
    return 0;
}

int dartb_step_host_gym(dartb_handle_t e, const float* h_action, float* obs_out, double* reward_out, uint8_t* done_out,
                        uint8_t* truncated_out, int32_t auto_reset, void* stream) {
    if (!e) return fail("null handle");
    if (!h_action || !obs_out || !reward_out || !done_out) return fail("null host pointer");
    if (e->task.n_obs == 0) return fail("physics-only handle: no task layer configured (use dartb_substep)");
    const int n = e->n, na = e->task.n_act, no = e->task.n_obs;
    const size_t fa = (size_t)n * na, fo = (size_t)n * no, fr = (size_t)n, fd = ((size_t)n + 3) / 4;
    DeviceGuard g(e->device);
    if (ensure_stage(e, fa + fo + fr + fd)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (envcfg().zerocopy && e->h_stage_dev) {
        // Page-locked outputs (the DartEnv wrapper hands out arrays of its pinned pool): the kernel itself writes the
        // reference's return types (float32 obs, float64 rewards, bool dones / truncated flags) over PCIe: ONE launch
        // and ONE sync per env.step(), no conversion pass, no memcpy.
        float* o_dev = (float*)alias_of(e, obs_out);
        double* r_dev = (double*)alias_of(e, reward_out);
        uint8_t* d_dev = (uint8_t*)alias_of(e, done_out);
        uint8_t* t_dev = truncated_out ? (uint8_t*)alias_of(e, truncated_out) : nullptr;
        if (o_dev && r_dev && d_dev && (t_dev || !truncated_out)) {
            const float* a_dev = (const float*)alias_of(e, h_action);
            if (!a_dev) { std::memcpy(e->h_stage, h_action, fa * 4); a_dev = e->h_stage_dev; }
            int rc = e->f64 ? launch_step<double>(e, a_dev, o_dev, nullptr, d_dev, auto_reset, st, r_dev, t_dev)
                            : launch_step<float>(e, a_dev, o_dev, nullptr, d_dev, auto_reset, st, r_dev, t_dev);
            if (rc) return rc;
            CK(cudaStreamSynchronize(st));
            return 0;
        }
    }
    // pageable outputs: the kernel writes into the page-locked staging block; the conversion to the reference's
    // return types (gym/vector/sync_vector_env.py:44-47: float64 rewards, bool dones) happens here, in one pass
    float* so = e->h_stage + fa;
    float* sr = so + fo;
    uint8_t* sd = (uint8_t*)(sr + fr);
    int rc = dartb_step_host(e, h_action, so, sr, sd, auto_reset, stream);
    if (rc) return rc;
    std::memcpy(obs_out, so, fo * 4);
    for (int i = 0; i < n; i++) reward_out[i] = (double)sr[i];
    for (int i = 0; i < n; i++) done_out[i] = sd[i] & 1;
    if (truncated_out) for (int i = 0; i < n; i++) truncated_out[i] = (sd[i] >> 1) & 1;
    return 0;
}

int dartb_substep(dartb_handle_t e, const float* d_tau, const float* d_fext, void* stream) {
    if (!e) return fail("null handle");
    DeviceGuard g(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (!e->f64) return launch_substep<float>(e, d_tau, d_fext, st);
    // fp64 engine fed fp32 inputs: widen through the scratch buffer
    double* sc = (double*)e->scratch;
    const double* tau = nullptr; const double* fx = nullptr;
    double* sf = sc + (size_t)e->n * e->nd;
    if (d_tau) { size_t k = (size_t)e->n * e->nd; k_convert<float, double><<<(unsigned)((k + 255) / 256), 256, 0, st>>>(k, d_tau, sc); tau = sc; e->launches++; }
    if (d_fext) { size_t k = (size_t)e->n * e->n_orig_bodies * 3; k_convert<float, double><<<(unsigned)((k + 255) / 256), 256, 0, st>>>(k, d_fext, sf); fx = sf; e->launches++; }
    return launch_substep<double>(e, tau, fx, st);
}
int dartb_substep_f64(dartb_handle_t e, const double* d_tau, const double* d_fext, void* stream) {
    if (!e) return fail("null handle");
    DeviceGuard g(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (e->f64) return launch_substep<double>(e, d_tau, d_fext, st);
    float* sc = (float*)e->scratch;
    float* sf = sc + (size_t)e->n * e->nd;
    const float* tau = nullptr; const float* fx = nullptr;
    if (d_tau) { size_t k = (size_t)e->n * e->nd; k_convert<double, float><<<(unsigned)((k + 255) / 256), 256, 0, st>>>(k, d_tau, sc); tau = sc; e->launches++; }
    if (d_fext) { size_t k = (size_t)e->n * e->n_orig_bodies * 3; k_convert<double, float><<<(unsigned)((k + 255) / 256), 256, 0, st>>>(k, d_fext, sf); fx = sf; e->launches++; }
    return launch_substep<float>(e, tau, fx, st);
}

int dartb_set_obs_peers(dartb_handle_t e, void* const* d_peers, int32_t n_peers, int64_t float_offset) {
    if (!e) return fail("null handle");
    if (n_peers < 0 || n_peers > DARTB_MAX_PEERS) return fail("dartb_set_obs_peers: 0..8 peers");
    if (n_peers > 0 && (!d_peers || float_offset < 0)) return fail("dartb_set_obs_peers: bad arguments");
    for (int p = 0; p < n_peers; p++) { if (!d_peers[p]) return fail("dartb_set_obs_peers: null peer pointer"); e->obs_peer[p] = (float*)d_peers[p]; }
    e->n_obs_peers = n_peers; e->obs_peer_off = float_offset;
    return 0;
}

int dartb_set_aux(dartb_handle_t e, const double* d_aux, void* stream) {
    if (!e || !d_aux) return fail("null argument");
    if (!e->aux) return fail("this task kind has no auxiliary per-world state");
    DeviceGuard g(e->device);
    const int tot = e->n * 3, bs = 256, grid = (tot + bs - 1) / bs;
    if (e->f64) k_to_soa<double, double><<<grid, bs, 0, (cudaStream_t)stream>>>(e->n, 3, d_aux, (double*)e->aux);
    else k_to_soa<double, float><<<grid, bs, 0, (cudaStream_t)stream>>>(e->n, 3, d_aux, (float*)e->aux);
    e->launches++;
    CK(cudaGetLastError());
    return 0;
}
int dartb_get_aux(dartb_handle_t e, double* d_aux, void* stream) {
    if (!e || !d_aux) return fail("null argument");
    if (!e->aux) return fail("this task kind has no auxiliary per-world state");
    DeviceGuard g(e->device);
    const int tot = e->n * 3, bs = 256, grid = (tot + bs - 1) / bs;
    if (e->f64) k_from_soa<double, double><<<grid, bs, 0, (cudaStream_t)stream>>>(e->n, 3, (const double*)e->aux, d_aux);
    else k_from_soa<float, double><<<grid, bs, 0, (cudaStream_t)stream>>>(e->n, 3, (const float*)e->aux, d_aux);
    e->launches++;
    CK(cudaGetLastError());
    return 0;
}

int dartb_get_contacts(dartb_handle_t e, int32_t* d_count, int32_t* d_body, float* d_data, void* stream) {
    if (!e) return fail("null handle");
    if (!e->contacts_valid)
        return fail("contact read-back of the fused step is off: dartb_set_option(h, DARTB_OPT_CONTACTS, 1) before stepping");
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
    e.n = 1 << 30;   // describe the topology-level lowering (the batch-size dependent choice of the cooperative form is per handle)
    if (lower_into(&e)) return 1;
    std::snprintf(buf, (size_t)len, "%s nd=%d max_contacts=%d", e.kernel_name.c_str(), e.nd, e.max_contacts);
    return 0;
}
const char* dartb_version(void) { return "dartb 0.1 (sm_100a)"; }

}  // extern "C"
