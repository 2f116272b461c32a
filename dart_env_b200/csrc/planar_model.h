// planar_model.h — the lowered (planar) model the sm_100a kernels consume.
//
// Every skeleton in SURVEY.md §8's scope (hopper, walker2d, half_cheetah, snake_7link; next:
// cartpole, double pendulum, reacher2d) is PLANAR: all revolute axes share one direction n, all
// prismatic axes are orthogonal to it.  The host lowering (lower.h) proves that, merges weld
// joints into their parents, and re-expresses the DART model (include/dartb.h, restating
// DART's SkelParser output) in plane coordinates (e1, e2) with e1 x e2 = n:
//   - one dof per body; body frame origin = joint origin; orientation angle theta_i
//   - spatial quantities are 3-vectors [angular; lin_x; lin_y] in world axes at the body origin
//   - articulated inertias are symmetric 3x3 (6 scalars) instead of DART's 6x6
// so a whole world fits in registers.  Results are identical (up to rounding) to the 3-D
// formulation for planar skeletons; non-planar skeletons are rejected at dartb_create.
#pragma once
#include <stdint.h>

#define PM_MAXB 12   /* bodies after weld merge (= dofs) */
#define PM_MAXS 12   /* capsule shapes */
#define PM_MAXA 8    /* actuators */
#define PM_MAXD 24   /* DART bodynodes before the weld merge (= DARTB_MAX_BODIES) */

enum { PM_REV = 1, PM_PRI = 2 };

template <typename R>
struct PModel {
    int32_t nb, ns;
    R dt, gx, gy;
    // bodies (index = dof index)
    int32_t parent[PM_MAXB];
    int32_t jtype[PM_MAXB];
    R sgn[PM_MAXB];            // revolute: theta_i = theta_parent + sgn * q
    R ax[PM_MAXB], ay[PM_MAXB];// joint anchor in the parent's planar frame
    R ux[PM_MAXB], uy[PM_MAXB];// prismatic axis in the parent's planar frame
    R mass[PM_MAXB], cx[PM_MAXB], cy[PM_MAXB], izz[PM_MAXB];
    R ox[PM_MAXB], oy[PM_MAXB];// DART body origin in the planar body frame (add_ext_force point)
    R damping[PM_MAXB], kspring[PM_MAXB], rest[PM_MAXB];
    R coulomb[PM_MAXB];        // joint Coulomb friction (force); row bounds are +-coulomb*dt
    int32_t any_coulomb;
    R qlo[PM_MAXB], qhi[PM_MAXB];
    int32_t limited[PM_MAXB];
    R qinit[PM_MAXB], dqinit[PM_MAXB];
    R fnx[PM_MAXB], fny[PM_MAXB]; // body-local ez (snake fluid normal) in the planar body frame
    int32_t orig_body[PM_MAXB];// DART bodynode index of this planar body (first of a merged group)
    // DART bodynodes (before the weld merge): which planar body carries them, and where their
    // origin sits in that planar body frame (bn.add_ext_force application point)
    int32_t nbd;
    int32_t dgroup[PM_MAXD];
    R dox[PM_MAXD], doy[PM_MAXD];
    // capsule shapes
    int32_t sbody[PM_MAXS];    // planar body
    int32_t sorig[PM_MAXS];    // DART bodynode index the shape belongs to (contact read-back)
    R scx[PM_MAXS], scy[PM_MAXS];   // capsule centre, planar body frame
    R sdx[PM_MAXS], sdy[PM_MAXS];   // capsule axis (unit), planar body frame
    R shalf[PM_MAXS], srad[PM_MAXS];
    R smu[PM_MAXS];            // min(body friction, ground friction = 1)
    // one static box in plane coordinates (axis aligned)
    int32_t has_ground;
    R gcx, gcy, ghx, ghy;      // centre / half extents in (e1, e2)
    R gupx, gupy;              // box local +y in plane coordinates (deep-penetration fallback)
    R ghup;                    // half extent along that direction
    // plane basis in world coordinates, and the constant out-of-plane coordinate
    R e1[3], e2[3], en[3];
    R hz;
};

template <typename R>
struct PTask {
    int32_t frame_skip, n_act, n_obs;
    int32_t dof_act[PM_MAXB];  // actuator index driving dof i, -1: none   (tau[3:] = a * scale)
    R dof_scale[PM_MAXB], dof_lo[PM_MAXB], dof_hi[PM_MAXB];
    int32_t obs_mode;
    R dq_clip;
    int32_t height_body;       // planar body, -1 unused
    R hcx, hcy;                // COM of the DART height body in that planar body frame
    R wy1, wy2, wy0;           // world_y = wy1*X + wy2*Y + wy0
    R height_lo, height_hi, ang_max;
    R alive_bonus, ctrl_cost, vel_weight;
    int32_t limit_pen_dof;
    R limit_pen_margin, limit_pen_weight;
    R dev_cost;
    int32_t zero_reward_on_blowup, fluid_force;
    R fluid_offset, fluid_coef;
    R reset_noise, state_bound;
    R inv_dt_env;
    // contact-free task kinds (DARTB_TASK_*): probes = body-fixed points whose world position the task reads
    int32_t kind;
    R noise_dq;                // reset noise on dq
    int32_t probe_body[2];     // planar body
    R probe_x[2], probe_y[2];  // point in that planar body frame
    R probe_n[2];              // its (constant) out-of-plane coordinate
};
