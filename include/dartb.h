/*
 * dartb.h — C-ABI of the B200-native batched DART stepper (libdartb.so).
 *
 * This is the drop-in boundary for the pydart2 calls made by the reference's
 * gym/envs/dart/dart_env.py.  Every entry point names the reference
 * call site it replaces.  Plain C types only: no torch, no C++ in the signatures.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; dartb_last_error()
 *     returns a thread-local message for the last failure.  No exception crosses the ABI.
 *   - `d_*` pointers are DEVICE pointers owned by the caller (e.g. torch tensors);
 *     `h_*` pointers are HOST pointers.  Nothing is retained past the call.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All
 *     work is enqueued on it; no hidden synchronisation except in the h_* calls.
 *   - boundary arrays are row-major [n_worlds, k] float32 (north_star: "fp32").
 *     Internally state is SoA [k][n_worlds].
 *   - a handle is bound to one device and used from one host thread at a time.
 *   - there is NO CPU fallback: without a CUDA device dartb_create fails.
 */
#ifndef DARTB_H
#define DARTB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DARTB_MAX_BODIES 24
#define DARTB_MAX_SHAPES 24
#define DARTB_MAX_PEERS  8
#define DARTB_MAX_GROUND 4
#define DARTB_MAX_ACT 16

enum { DARTB_JOINT_WELD = 0, DARTB_JOINT_REVOLUTE = 1, DARTB_JOINT_PRISMATIC = 2 };
enum { DARTB_SHAPE_BOX = 0, DARTB_SHAPE_CAPSULE = 1, DARTB_SHAPE_SPHERE = 2,
       DARTB_SHAPE_ELLIPSOID = 3, DARTB_SHAPE_CYLINDER = 4 };

/* One BodyNode + its parent Joint (DART creates them as a pair).  Filled by the host
 * model compiler (dart_env_b200/skel.py) which restates DART's SkelParser; replaces
 * pydart.World(dt, skel_path) at dart_env.py:54-55. */
typedef struct dartb_body {
    int32_t parent;            /* -1 = world */
    int32_t joint_type;        /* DARTB_JOINT_* */
    int32_t dof;               /* index into q, -1 for weld */
    int32_t limit_enforced;    /* dart_env.py:64-67 */
    double  T_parent_joint[12];/* 3x4 row-major [R|p]: joint frame in the parent body frame */
    double  T_child_joint[12]; /* joint frame in the child body frame */
    double  axis[3];           /* joint axis in the joint frame (unit) */
    double  q_lo, q_hi;
    double  damping, coulomb, spring_k, spring_rest;
    double  q_init, dq_init;
    double  mass;
    double  com[3];            /* local COM */
    double  inertia[9];        /* moment about the COM, body axes, row-major */
    double  friction_coeff;    /* BodyNode friction, default 1.0 */
} dartb_body_t;

typedef struct dartb_shape {
    int32_t body;              /* robot body index; -1 = world-fixed */
    int32_t type;              /* DARTB_SHAPE_* */
    double  size[3];           /* box xyz | capsule (radius,height,-) | sphere (r,-,-) */
    double  T[12];             /* 3x4 row-major, body-local (or world for body = -1) */
} dartb_shape_t;

typedef struct dartb_model {
    double  dt;
    double  gravity[3];
    int32_t n_bodies, n_dofs, n_shapes, n_ground;
    dartb_body_t  bodies[DARTB_MAX_BODIES];
    dartb_shape_t shapes[DARTB_MAX_SHAPES];
    dartb_shape_t ground[DARTB_MAX_GROUND];
} dartb_model_t;

/* Task layer: the arithmetic in gym/envs/dart/{hopper,walker2d,half_cheetah,snake_7link}.py
 * step()/advance()/_get_obs()/reset_model(), parameterised. */
enum { DARTB_OBS_Q1_DQ = 0,       /* [q[1:], dq]                 half_cheetah.py:79-85, snake_7link.py:89-99 */
       DARTB_OBS_HEIGHT_Q2_DQ = 1 /* [com_y(body), q[2:], clip(dq)] hopper.py:67-74, walker2d.py:67-74 */ };

typedef struct dartb_task {
    int32_t frame_skip;            /* dart_env.py:170 n_frames */
    int32_t n_act, n_obs;
    int32_t act_dof[DARTB_MAX_ACT];/* tau[act_dof[i]] = clamp(a[i], lo, hi) * scale[i] */
    double  act_scale[DARTB_MAX_ACT];
    double  act_lo[DARTB_MAX_ACT], act_hi[DARTB_MAX_ACT];
    int32_t obs_mode;              /* DARTB_OBS_* */
    double  dq_clip;               /* > 0: clip dq in obs to +-dq_clip (hopper.py:70), 0: none */
    int32_t height_body;           /* bodynodes[i].com()[1] (hopper.py:42); -1: unused */
    double  height_lo, height_hi;  /* done unless lo < height < hi; +-inf disables */
    double  ang_max;               /* done unless |q[2]| < ang_max */
    double  alive_bonus;
    double  ctrl_cost;             /* reward -= ctrl_cost * sum(a^2) with the RAW action */
    double  vel_weight;
    int32_t limit_pen_dof;         /* hopper.py:45-50 joint-limit penalty dof, -1: none */
    double  limit_pen_margin, limit_pen_weight;
    double  dev_cost;              /* snake_7link.py:80: reward -= dev_cost * |q[2]| */
    int32_t zero_reward_on_blowup; /* half_cheetah.py:55-59 */
    int32_t fluid_force;           /* snake_7link.py:35-50 per-substep fluid force */
    double  fluid_offset, fluid_coef; /* 0.05, 50.0 */
    double  reset_noise;           /* U(-noise, +noise) on q and dq (hopper.py:78-79) */
    double  state_bound;           /* 100: done if any |s[2:]| >= bound or non-finite */
    /* the contact-free envs' step()/_get_obs()/reset_model(), selected by `kind` (0 = the
     * locomotion layer parameterised above).  Actuators use act_dof / act_scale / act_lo / act_hi (+-inf = the
     * reference does not clamp). */
    int32_t kind;                  /* DARTB_TASK_* */
    double  reset_noise_dq;        /* noise on dq when it differs from reset_noise (< 0: same) */
    int32_t probe_body[2];         /* DOUBLE_PENDULUM: 'cart', 'weight' (to_world()[1], inverted_double_pendulum.py:28-31);
                                      REACHER2D: [0] = bodynodes[-1] (com(), reacher2d.py:31) */
    double  probe_local[2][3];     /* the body-local point of each probe (origin, or the local COM) */
} dartb_task_t;

enum { DARTB_TASK_LOCOMOTION = 0,      /* hopper / walker2d / half_cheetah / snake_7link */
       DARTB_TASK_CARTPOLE = 1,        /* cart_pole.py:12-36: obs [q, dq], reward 1, done unless finite and |q1| <= 0.2 */
       DARTB_TASK_SWINGUP = 2,         /* cartpole_swingup.py:14-52: reward 6 - |q1| - 0.01 a^2 - 0.01 |q0|; reset flips q1 by +-pi */
       DARTB_TASK_DOUBLE_PENDULUM = 3, /* inverted_double_pendulum.py:19-63: obs [q0, sin, cos, dq]; randn velocity noise */
       DARTB_TASK_REACHER2D = 4        /* reacher2d.py:17-64: per-world target (dartb_set_aux), reward -|tip - target| - a^2 */ };

/* Options for dartb_set_option */
enum { DARTB_OPT_LCP_MODE = 1,    /* 0 = exact (Dantzig-equivalent), 1 = PGS */
       DARTB_OPT_PGS_ITERS = 2,
       DARTB_OPT_FRICTION_ALL = 3,/* set_friction_coeff(mu) on every body (snake_7link.py:29-31) */
       DARTB_OPT_MAX_EPISODE_STEPS = 4,/* TimeLimit (gym/wrappers/time_limit.py:14-21); 0 = off */
       DARTB_OPT_KERNEL_VARIANT = 5, /* -1 = auto, 0 = unrolled per-topology kernel (one world per thread), 1 = loop / topology-generic kernel, 2 = lane-cooperative kernel (8/16 lanes per world), 3 = quad form of the per-thread kernel (4 lanes per world share the constraint phase); auto: lane-cooperative, quad, per-thread by growing batch size */
       DARTB_OPT_WORLDS_PER_WARP = 6,/* launch shape of dartb_step: worlds per warp, 0 = auto; results do not depend on it */
       DARTB_OPT_CONTACTS = 7,     /* 1 = dartb_step records world.collision_result.contacts of its last sub-step for
                                      dartb_get_contacts (walker2d.py:38-41); default 0: the record is 356 B per world
                                      per step, more than the rest of the step's HBM traffic.  dartb_substep (the
                                      literal World.step()) always records. */
       DARTB_OPT_RANDOMIZE_MASS = 8,    /* half range r >= 0: at EVERY reset of a world (dartb_reset and the auto-reset
                                      inside dartb_step) its bodynode masses become original + U(-r, r), clipped at 0 —
                                      snake_7link.py:115-118 (r = 1.5) inside the kernel; 0 = off */
       DARTB_OPT_RANDOMIZE_FRICTION = 9 /* the same for the friction coefficients, original + U(-r, r) clipped at 0,
                                      snake_7link.py:119-120 (r = 0.5).  Both write the per-world table of
                                      dartb_set_body_params (the batch runs on the loop kernels) and need a skeleton
                                      without welded bodynodes (one planar body per bodynode). */ };

typedef struct dartb_engine* dartb_handle_t;

/* Build an engine stepping n_worlds copies of `model` on CUDA device `device`.
 * `world_offset` is the global id of this handle's first world (RNG streams are keyed by
 * global world id so results do not depend on how worlds are sharded across GPUs).
 * Replaces pydart.init() + DartWorld(dt, path) (dart_env.py:19,54-59). */
int dartb_create(const dartb_model_t* model, const dartb_task_t* task, int32_t n_worlds,
                 int32_t device, uint64_t seed, int64_t world_offset, dartb_handle_t* out);
int dartb_destroy(dartb_handle_t h);

int dartb_set_option(dartb_handle_t h, int32_t key, double value);

/* env.seed(s) (dart_env.py:117-119): re-keys the reset-noise generator (Philox4x32-10 keyed by seed, global
 * world id and episode counter).  Like the reference it touches nothing else: physics state, episode counters
 * and options stay as they are. */
int dartb_seed(dartb_handle_t h, uint64_t seed);
/* VectorEnv.seed([s_0, ..., s_{n-1}]) (gym/vector/sync_vector_env.py:50-57): world i draws what a single env seeded
 * s_i draws.  h_seeds is a HOST array of n_worlds values; synchronous. */
int dartb_seed_worlds(dartb_handle_t h, const uint64_t* h_seeds);

/* world.reset() + reset_model(): q0 + U(+-noise), dq0 + U(+-noise), returns obs.
 * d_mask: uint8[n] (NULL = all worlds). d_obs may be NULL.  (dart_world.py:20-22, hopper.py:76-84) */
int dartb_reset(dartb_handle_t h, const uint8_t* d_mask, float* d_obs, void* stream);

/* set_positions/set_velocities and q/dq reads (dart_env.py:145-148, 211-215). [n, nd] fp32. */
int dartb_set_state(dartb_handle_t h, const float* d_q, const float* d_dq, void* stream);
int dartb_get_state(dartb_handle_t h, float* d_q, float* d_dq, void* stream);
/* fp64 variants used by the tight-tolerance parity tests (engine created with fp64 state). */
int dartb_set_state_f64(dartb_handle_t h, const double* d_q, const double* d_dq, void* stream);
int dartb_get_state_f64(dartb_handle_t h, double* d_q, double* d_dq, void* stream);

/* One env.step(): clamp/scale action, frame_skip x {set_forces; world.step()}, obs/reward/done,
 * optional auto-reset of done worlds (gym/vector/sync_vector_env.py:76-79 semantics: the returned
 * obs of a done world is its reset obs).  d_action [n,n_act], d_obs [n,n_obs], d_reward [n],
 * d_done uint8[n]: non-zero = done; bit 1 set = the episode was cut by the time limit only
 * (info['TimeLimit.truncated'], gym/wrappers/time_limit.py:18-20).
 * (hopper.py:24-65 and siblings; dart_env.py:158-175) */
int dartb_step(dartb_handle_t h, const float* d_action, float* d_obs, float* d_reward,
               uint8_t* d_done, int32_t auto_reset, void* stream);

/* The same env.step() for HOST buffers (numpy arrays of the reference-facing wrapper), synchronous:
 * when it returns, h_obs / h_reward / h_done hold the results.  h_* are ordinary host pointers.
 * Page-locked buffers are read / written by the step kernel itself over PCIe (zero-copy: one launch
 * and one sync, no memcpy nodes); pageable ones go through the library's pinned staging block.
 * DARTB_ZEROCOPY=0 selects the explicit H2D / D2H copy path instead.  This is the end-to-end call
 * a DartEnv user makes per step (bench.py "e2e"). */
int dartb_step_host(dartb_handle_t h, const float* h_action, float* h_obs, float* h_reward,
                    uint8_t* h_done, int32_t auto_reset, void* stream);

/* Tell the handle that [h_ptr, h_ptr + bytes) is page-locked, device-mapped host memory (cudaHostAlloc /
 * cudaHostRegister / torch pin_memory) that stays so until it is unregistered or the handle is destroyed.
 * dartb_step_host / dartb_step_host_gym then use pointers inside it zero-copy WITHOUT asking the driver on
 * every step (cudaPointerGetAttributes costs microseconds: several per step exceed the step kernel). */
int dartb_register_host(dartb_handle_t h, const void* h_ptr, size_t bytes);
int dartb_unregister_host(dartb_handle_t h, const void* h_ptr);

/* The same step with the reference's vectorised RETURN TYPES (gym/vector/sync_vector_env.py:44-47,73-84):
 * obs_out float32 [n, n_obs] (a fresh copy: VectorEnv(copy=True)), reward_out float64 [n], done_out bool [n]
 * (1 byte each), truncated_out bool [n] or NULL (TimeLimit.truncated, gym/wrappers/time_limit.py:14-21).
 * Page-locked output arrays are written by the step kernel itself in these types over PCIe (one launch, one
 * sync, no conversion pass); pageable ones are filled from the engine's page-locked staging block after the sync. */
int dartb_step_host_gym(dartb_handle_t h, const float* h_action, float* obs_out, double* reward_out,
                        uint8_t* done_out, uint8_t* truncated_out, int32_t auto_reset, void* stream);

/* Exactly `skel.set_forces(tau); world.step()` (dart_env.py:174-175): one DART time step
 * with generalized forces d_tau [n, nd]; optional external world-frame forces at body
 * origins d_fext [n, n_bodies, 3] (bn.add_ext_force, snake_7link.py:47), may be NULL. */
int dartb_substep(dartb_handle_t h, const float* d_tau, const float* d_fext, void* stream);
int dartb_substep_f64(dartb_handle_t h, const double* d_tau, const double* d_fext, void* stream);

/* Fused observation all-gather (the one optional collective, replacing the shared-memory observation buffer
 * of gym/vector/async_vector_env.py:90-94).  After this call dartb_step ALSO stores every observation row it writes to
 * d_obs into each of the n_peers buffers, at float offset `float_offset` + the row's offset in d_obs: with the peers'
 * gather buffers mapped into this device's address space (NVLink peer access / symmetric memory) and float_offset =
 * rank * n_worlds * n_obs, the all-gather happens inside the step kernel, tile by tile, instead of as a collective after
 * it.  The caller synchronises the ranks before reading (any barrier: the stores are complete when the step kernel is).
 * n_peers = 0 switches it off.  d_peers is a HOST array of device pointers. */
int dartb_set_obs_peers(dartb_handle_t h, void* const* d_peers, int32_t n_peers, int64_t float_offset);

/* Per-world auxiliary task state, [n, 3] fp64 (converted to the engine precision inside): the reacher's target
 * (`self.target`, reacher2d.py:7,57-63: world x, y, z).  Resets redraw it inside the kernel; these calls back the
 * attribute for callers and tests.  Fails for task kinds without auxiliary state. */
int dartb_set_aux(dartb_handle_t h, const double* d_aux, void* stream);
int dartb_get_aux(dartb_handle_t h, double* d_aux, void* stream);

/* Per-world dynamics parameters (the dynamics randomisation of snake_7link.py:20-25,115-120 —
 * bodynodes[i].set_mass(m), bodynodes[i].set_friction_coeff(mu) on one world of the batch).  h_mass / h_friction are HOST
 * arrays [n_worlds, model.n_bodies] in DART bodynode order (either may be NULL = the model's values; both NULL returns the
 * engine to the shared model).  set_mass keeps the body's moment of inertia (DART 6 Inertia::setMass); a contact's friction
 * is min(body, ground) as in the shared model.  Synchronises the device; while per-world parameters are set the batch runs
 * on the topology-generic loop kernels, the only ones that read them (dartb_kernel_name reports "loop:generic"). */
int dartb_set_body_params(dartb_handle_t h, const double* h_mass, const double* h_friction);
/* The table the kernels read, HOST array [4 nb + ns][n_worlds] fp64: rows mass, cx, cy, izz of planar body 0..nb-1, then
 * the friction coefficient of capsule 0..ns-1.  *n_rows (may be NULL) receives 4 nb + ns; h_out may be NULL to ask for
 * the row count only.  Fails when no per-world parameters are set.  Synchronises. */
int dartb_get_body_table(dartb_handle_t h, double* h_out, int32_t* n_rows);

/* world.collision_result.contacts of the LAST sub-step (walker2d.py:38-41).
 * d_count int32[n]; d_body int32[n, max_contacts] (robot body index per contact, -1 padded,
 * ordered by collision-shape index); d_data float[n, max_contacts, 10] =
 * point(3) normal(3) depth(1) force(3).  Any pointer may be NULL. */
int dartb_get_contacts(dartb_handle_t h, int32_t* d_count, int32_t* d_body, float* d_data,
                       void* stream);
int32_t dartb_max_contacts(dartb_handle_t h);
/* info['TimeLimit.truncated'] of the last dartb_step (time_limit.py:18-20): uint8[n] */
int dartb_get_truncated(dartb_handle_t h, uint8_t* d_out, void* stream);

/* Introspection used by the host wrapper and the tests */
int32_t dartb_num_worlds(dartb_handle_t h);
int32_t dartb_num_dofs(dartb_handle_t h);
int32_t dartb_is_f64(dartb_handle_t h);
/* number of CUDA kernels this handle has launched since creation (bench gpu_launches) */
int64_t dartb_launch_count(dartb_handle_t h);
/* Which kernel specialisation the model was lowered to, e.g. "planar-xy/static:hopper6" */
const char* dartb_kernel_name(dartb_handle_t h);

/* Engines created with this flag set keep fp64 state and run the fp64 instantiation of the
 * kernels (validation path: separates algorithmic from precision differences). */
int dartb_create_f64(const dartb_model_t* model, const dartb_task_t* task, int32_t n_worlds,
                     int32_t device, uint64_t seed, int64_t world_offset, dartb_handle_t* out);

/* Host-only: lower `model` and report the kernel specialisation it maps to (or fail with the
 * reason it is out of the planar kernels' scope).  Needs no GPU. */
int dartb_describe(const dartb_model_t* model, const dartb_task_t* task, char* buf, int32_t len);

const char* dartb_last_error(void);
const char* dartb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DARTB_H */
