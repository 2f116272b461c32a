// lower.h — host-side model lowering: dartb_model_t (DART / SkelParser semantics, 3-D)
//           -> PModel<double> / PTask<double> (planar, weld-merged; see planar_model.h).
//
// Native replacement for the part of DART's dynamics::Skeleton construction that the kernels
// need (reference call site: pydart.World(dt, skel_path), dart_env.py:54-55).  Everything is
// checked: a model the planar kernels cannot represent EXACTLY is rejected with a message
// (there is no silent approximation and no CPU fallback).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dartb.h"
#include "planar_model.h"

namespace lower {

struct V3 { double x, y, z; };
static inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline double norm(V3 a) { return std::sqrt(dot(a, a)); }

struct Xf { double R[9]; V3 p; };
static inline Xf xf_from12(const double t[12]) {
    Xf o;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o.R[3 * i + j] = t[4 * i + j];
    o.p = v3(t[3], t[7], t[11]);
    return o;
}
static inline V3 rot(const Xf& T, V3 v) {
    return v3(T.R[0] * v.x + T.R[1] * v.y + T.R[2] * v.z, T.R[3] * v.x + T.R[4] * v.y + T.R[5] * v.z,
              T.R[6] * v.x + T.R[7] * v.y + T.R[8] * v.z);
}
static inline V3 apply(const Xf& T, V3 v) { return rot(T, v) + T.p; }
static inline Xf mul(const Xf& a, const Xf& b) {
    Xf o;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            o.R[3 * i + j] = a.R[3 * i] * b.R[j] + a.R[3 * i + 1] * b.R[3 + j] + a.R[3 * i + 2] * b.R[6 + j];
    o.p = apply(a, b.p);
    return o;
}
static inline Xf inv(const Xf& a) {
    Xf o;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o.R[3 * i + j] = a.R[3 * j + i];
    V3 t = rot(o, a.p);
    o.p = v3(-t.x, -t.y, -t.z);
    return o;
}

struct Result {
    PModel<double> m;
    PTask<double> t;
    std::string signature;  // topology signature used to pick a static kernel instantiation
    int max_contacts;
};

// Returns "" on success, otherwise the reason the model is out of the planar kernels' scope.
static std::string lower_model(const dartb_model_t& dm, const dartb_task_t& dt_, Result& out) {
    char buf[256];
    const int nbd = dm.n_bodies;
    if (nbd < 1 || nbd > DARTB_MAX_BODIES) return "bad body count";
    PModel<double>& m = out.m;
    PTask<double>& t = out.t;
    std::memset(&m, 0, sizeof m);
    std::memset(&t, 0, sizeof t);

    // ---- 3-D forward kinematics at q = 0
    std::vector<Xf> Tw(nbd), Tcj(nbd);
    for (int i = 0; i < nbd; i++) {
        const dartb_body_t& b = dm.bodies[i];
        if (b.parent >= i) return "bodies are not topologically ordered";
        Xf Tpj = xf_from12(b.T_parent_joint);
        Tcj[i] = xf_from12(b.T_child_joint);
        Xf rel = mul(Tpj, inv(Tcj[i]));
        Tw[i] = b.parent < 0 ? rel : mul(Tw[b.parent], rel);
    }
    // ---- plane normal
    V3 n = v3(0, 0, 0);
    bool have_n = false;
    for (int i = 0; i < nbd && !have_n; i++)
        if (dm.bodies[i].joint_type == DARTB_JOINT_REVOLUTE) {
            const double* a = dm.bodies[i].axis;
            n = rot(Tw[i], rot(Tcj[i], v3(a[0], a[1], a[2])));
            have_n = true;
        }
    if (!have_n) {
        std::vector<V3> pa;
        for (int i = 0; i < nbd; i++)
            if (dm.bodies[i].joint_type == DARTB_JOINT_PRISMATIC) {
                const double* a = dm.bodies[i].axis;
                pa.push_back(rot(Tw[i], rot(Tcj[i], v3(a[0], a[1], a[2]))));
            }
        if (pa.empty()) return "model has no degrees of freedom";
        n = v3(0, 0, 1);
        for (size_t k = 1; k < pa.size(); k++) {
            V3 c = cross(pa[0], pa[k]);
            if (norm(c) > 1e-6) { n = (1.0 / norm(c)) * c; break; }
        }
        if (std::fabs(dot(n, pa[0])) > 1e-9) n = std::fabs(pa[0].z) < 0.9 ? cross(pa[0], v3(0, 0, 1)) : cross(pa[0], v3(1, 0, 0));
        n = (1.0 / norm(n)) * n;
    }
    V3 e1, e2, en;
    if (std::fabs(std::fabs(n.z) - 1) < 1e-9) { en = v3(0, 0, 1); e1 = v3(1, 0, 0); e2 = v3(0, 1, 0); }
    else if (std::fabs(std::fabs(n.y) - 1) < 1e-9) { en = v3(0, 1, 0); e1 = v3(0, 0, 1); e2 = v3(1, 0, 0); }
    else if (std::fabs(std::fabs(n.x) - 1) < 1e-9) { en = v3(1, 0, 0); e1 = v3(0, 1, 0); e2 = v3(0, 0, 1); }
    else return "plane normal is not a world axis (unsupported)";
    m.e1[0] = e1.x; m.e1[1] = e1.y; m.e1[2] = e1.z;
    m.e2[0] = e2.x; m.e2[1] = e2.y; m.e2[2] = e2.z;
    m.en[0] = en.x; m.en[1] = en.y; m.en[2] = en.z;

    // ---- weld merge: group[b] = planar body index
    std::vector<int> group(nbd, -1), root;  // root[g] = DART body owning the joint
    for (int i = 0; i < nbd; i++) {
        const dartb_body_t& b = dm.bodies[i];
        if (b.joint_type == DARTB_JOINT_WELD) {
            if (b.parent < 0) return "weld to world is not supported";
            group[i] = group[b.parent];
        } else if (b.joint_type == DARTB_JOINT_REVOLUTE || b.joint_type == DARTB_JOINT_PRISMATIC) {
            group[i] = (int)root.size();
            root.push_back(i);
            if (b.dof != group[i]) return "dof order does not follow body order";
        } else return "unsupported joint type";
    }
    const int nb = (int)root.size();
    if (nb > PM_MAXB) return "too many dofs for the planar kernels";
    if (nb != dm.n_dofs) return "dof count mismatch";
    m.nb = nb;
    m.dt = dm.dt;
    V3 g = v3(dm.gravity[0], dm.gravity[1], dm.gravity[2]);
    m.gx = dot(e1, g);
    m.gy = dot(e2, g);

    std::vector<V3> P0(nb);
    std::string sig;
    for (int gi = 0; gi < nb; gi++) {
        int r = root[gi];
        const dartb_body_t& b = dm.bodies[r];
        P0[gi] = apply(Tw[r], Tcj[r].p);
        int pg = b.parent < 0 ? -1 : group[b.parent];
        V3 Ppar = pg < 0 ? v3(0, 0, 0) : P0[pg];
        V3 d = P0[gi] - Ppar;
        m.parent[gi] = pg;
        m.ax[gi] = dot(e1, d);
        m.ay[gi] = dot(e2, d);
        V3 aw = rot(Tw[r], rot(Tcj[r], v3(b.axis[0], b.axis[1], b.axis[2])));
        if (b.joint_type == DARTB_JOINT_REVOLUTE) {
            double c = dot(aw, en);
            if (std::fabs(std::fabs(c) - 1) > 1e-9) {
                std::snprintf(buf, sizeof buf, "revolute joint %d is not parallel to the plane normal (non-planar skeleton)", r);
                return buf;
            }
            m.jtype[gi] = PM_REV;
            m.sgn[gi] = c > 0 ? 1.0 : -1.0;
        } else {
            if (std::fabs(dot(aw, en)) > 1e-9) {
                std::snprintf(buf, sizeof buf, "prismatic joint %d leaves the plane (non-planar skeleton)", r);
                return buf;
            }
            m.jtype[gi] = PM_PRI;
            m.ux[gi] = dot(e1, aw);
            m.uy[gi] = dot(e2, aw);
        }
        m.damping[gi] = b.damping;
        m.kspring[gi] = b.spring_k;
        m.rest[gi] = b.spring_rest;
        m.qlo[gi] = b.q_lo;
        m.qhi[gi] = b.q_hi;
        m.limited[gi] = b.limit_enforced;
        m.qinit[gi] = b.q_init;
        m.dqinit[gi] = b.dq_init;
        m.orig_body[gi] = r;
        m.coulomb[gi] = b.coulomb;
        if (b.coulomb != 0.0) m.any_coulomb = 1;
        V3 o = Tw[r].p - P0[gi];
        m.ox[gi] = dot(e1, o);
        m.oy[gi] = dot(e2, o);
        V3 ez = rot(Tw[r], v3(0, 0, 1));
        if (dt_.fluid_force) {
            // the fluid force mixes the BODY-frame COM velocity with the WORLD-frame normal (snake_7link.py:37-45); the
            // planar kernels restate that for body frames that coincide with the world axes at the zero pose
            for (int a = 0; a < 3; a++) for (int c = 0; c < 3; c++)
                if (std::fabs(Tw[r].R[3 * a + c] - (a == c ? 1.0 : 0.0)) > 1e-9)
                    return "fluid force needs body frames aligned with the world axes at the zero pose";
        }
        m.fnx[gi] = dot(e1, ez);
        m.fny[gi] = dot(e2, ez);
        std::snprintf(buf, sizeof buf, "%c%d,", b.joint_type == DARTB_JOINT_REVOLUTE ? 'R' : 'P', pg);
        sig += buf;
    }
    m.nbd = nbd;
    for (int i = 0; i < nbd; i++) {
        int gi = group[i];
        V3 o = Tw[i].p - P0[gi];
        m.dgroup[i] = gi;
        m.dox[i] = dot(e1, o);
        m.doy[i] = dot(e2, o);
    }
    // ---- mass properties per group
    for (int gi = 0; gi < nb; gi++) {
        double mass = 0;
        V3 mc = v3(0, 0, 0);
        for (int i = 0; i < nbd; i++)
            if (group[i] == gi) {
                const dartb_body_t& b = dm.bodies[i];
                mass += b.mass;
                mc = mc + b.mass * apply(Tw[i], v3(b.com[0], b.com[1], b.com[2]));
            }
        V3 C = mass > 0 ? (1.0 / mass) * mc : P0[gi];
        double izz = 0;
        for (int i = 0; i < nbd; i++)
            if (group[i] == gi) {
                const dartb_body_t& b = dm.bodies[i];
                V3 nl = rot(inv(Tw[i]), en);  // plane normal in body axes
                double In = 0;
                double nv[3] = {nl.x, nl.y, nl.z};
                for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) In += nv[r] * b.inertia[3 * r + c] * nv[c];
                V3 d = apply(Tw[i], v3(b.com[0], b.com[1], b.com[2])) - C;
                double dx = dot(e1, d), dy = dot(e2, d);
                izz += In + b.mass * (dx * dx + dy * dy);
            }
        m.mass[gi] = mass;
        m.izz[gi] = izz;
        V3 d = C - P0[gi];
        m.cx[gi] = dot(e1, d);
        m.cy[gi] = dot(e2, d);
    }
    // ---- static box
    double c_n = 0, h_n = 0;
    if (dm.n_ground > 1) return "more than one static collision shape is not supported";
    if (dm.n_ground == 1) {
        const dartb_shape_t& gs = dm.ground[0];
        if (gs.type != DARTB_SHAPE_BOX) return "static shape must be a box";
        Xf Tg = xf_from12(gs.T);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
            if (std::fabs(Tg.R[3 * i + j] - (i == j ? 1.0 : 0.0)) > 1e-9) return "static box must be axis aligned";
        V3 half = v3(0.5 * gs.size[0], 0.5 * gs.size[1], 0.5 * gs.size[2]);
        auto absdot = [](V3 a, V3 h) { return std::fabs(a.x) * h.x + std::fabs(a.y) * h.y + std::fabs(a.z) * h.z; };
        m.has_ground = 1;
        m.gcx = dot(e1, Tg.p); m.gcy = dot(e2, Tg.p);
        m.ghx = absdot(e1, half); m.ghy = absdot(e2, half);
        c_n = dot(en, Tg.p); h_n = absdot(en, half);
        V3 up = v3(0, 1, 0);
        m.gupx = dot(e1, up); m.gupy = dot(e2, up);
        m.ghup = half.y;
    }
    // ---- capsule shapes
    int ns = 0;
    bool hz_set = false;
    for (int si = 0; si < dm.n_shapes && m.has_ground; si++) {
        const dartb_shape_t& s = dm.shapes[si];
        if (s.body < 0 || s.body >= nbd) return "shape with a bad body index";
        Xf Ts = mul(Tw[s.body], xf_from12(s.T));
        double hzs = dot(en, Ts.p);
        double reach;
        if (s.type == DARTB_SHAPE_CAPSULE) reach = s.size[0];
        else if (s.type == DARTB_SHAPE_SPHERE) reach = s.size[0];
        else reach = norm(v3(s.size[0], s.size[1], s.size[2]));
        double gap_n = std::fabs(hzs - c_n) - h_n;
        if (gap_n > reach + 1e-12) continue;  // can never touch the static box: all motion is in-plane
        if (s.type != DARTB_SHAPE_CAPSULE) return "only capsule robot shapes collide with the ground (SURVEY 8f.3)";
        if (gap_n > 0) return "capsule straddles the static box out of plane (unsupported)";
        V3 ax = rot(Ts, v3(0, 0, 1));
        if (std::fabs(dot(ax, en)) > 1e-6) return "capsule axis leaves the plane (unsupported)";
        if (ns >= PM_MAXS) return "too many collision shapes";
        int gi = group[s.body];
        V3 d = Ts.p - P0[gi];
        double dx = dot(e1, ax), dy = dot(e2, ax), nn = std::sqrt(dx * dx + dy * dy);
        m.sbody[ns] = gi;
        m.sorig[ns] = s.body;
        m.scx[ns] = dot(e1, d); m.scy[ns] = dot(e2, d);
        m.sdx[ns] = dx / nn; m.sdy[ns] = dy / nn;
        m.shalf[ns] = 0.5 * s.size[1];
        m.srad[ns] = s.size[0];
        double mu = dm.bodies[s.body].friction_coeff;
        m.smu[ns] = mu < 1.0 ? mu : 1.0;
        if (!hz_set) { m.hz = hzs; hz_set = true; }
        std::snprintf(buf, sizeof buf, "S%d,", gi);
        sig += buf;
        ns++;
    }
    m.ns = ns;
    out.max_contacts = ns > 0 ? ns : 1;
    out.signature = sig;

    // ---- task (n_obs == 0: physics-only handle, dartb_substep / state access only)
    if (dt_.n_obs == 0 && dt_.n_act == 0) {
        for (int i = 0; i < PM_MAXB; i++) t.dof_act[i] = -1;
        t.frame_skip = dt_.frame_skip > 0 ? dt_.frame_skip : 1;
        t.height_body = -1; t.limit_pen_dof = -1;
        t.inv_dt_env = 1.0 / (dm.dt * t.frame_skip);
        return "";
    }
    if (dt_.n_act > PM_MAXA) return "too many actuators";
    t.frame_skip = dt_.frame_skip; t.n_act = dt_.n_act; t.n_obs = dt_.n_obs;
    for (int i = 0; i < PM_MAXB; i++) t.dof_act[i] = -1;
    for (int i = 0; i < dt_.n_act; i++) {
        int d = dt_.act_dof[i];
        if (d < 0 || d >= nb) return "actuator dof out of range";
        if (t.dof_act[d] >= 0) return "two actuators on one dof";
        t.dof_act[d] = i;
        t.dof_scale[d] = dt_.act_scale[i]; t.dof_lo[d] = dt_.act_lo[i]; t.dof_hi[d] = dt_.act_hi[i];
    }
    t.obs_mode = dt_.obs_mode;
    t.kind = dt_.kind;
    t.noise_dq = dt_.reset_noise_dq >= 0 ? dt_.reset_noise_dq : dt_.reset_noise;
    int expect = (dt_.obs_mode == DARTB_OBS_HEIGHT_Q2_DQ) ? (1 + (nb - 2) + nb) : ((nb - 1) + nb);
    int nprobe = 0;
    switch (dt_.kind) {
        case DARTB_TASK_LOCOMOTION: break;
        case DARTB_TASK_CARTPOLE: case DARTB_TASK_SWINGUP: expect = 2 * nb; break;
        case DARTB_TASK_DOUBLE_PENDULUM: expect = 1 + 2 * (nb - 1) + nb; nprobe = 2; break;
        case DARTB_TASK_REACHER2D: expect = 2 * nb + 2 + nb + 3; nprobe = 1; break;
        default: return "unknown task kind";
    }
    if (dt_.n_obs != expect) return "n_obs does not match the task's observation layout";
    if (dt_.kind != DARTB_TASK_LOCOMOTION && (nb < 2 || dt_.fluid_force)) return "contact-free task kinds need >= 2 dofs and no fluid force";
    for (int k = 0; k < 2; k++) {
        t.probe_body[k] = 0; t.probe_x[k] = 0; t.probe_y[k] = 0; t.probe_n[k] = 0;
        if (k >= nprobe) continue;
        int pb = dt_.probe_body[k];
        if (pb < 0 || pb >= nbd) return "probe body out of range";
        V3 C = apply(Tw[pb], v3(dt_.probe_local[k][0], dt_.probe_local[k][1], dt_.probe_local[k][2]));
        V3 d = C - P0[group[pb]];
        t.probe_body[k] = group[pb];
        t.probe_x[k] = dot(e1, d); t.probe_y[k] = dot(e2, d); t.probe_n[k] = dot(en, C);
    }
    t.dq_clip = dt_.dq_clip;
    t.height_body = -1;
    if (dt_.height_body >= 0) {
        if (dt_.height_body >= nbd) return "height_body out of range";
        int hb = dt_.height_body, gi = group[hb];
        const dartb_body_t& b = dm.bodies[hb];
        V3 C = apply(Tw[hb], v3(b.com[0], b.com[1], b.com[2]));
        V3 d = C - P0[gi];
        t.height_body = gi;
        t.hcx = dot(e1, d); t.hcy = dot(e2, d);
        t.wy1 = e1.y; t.wy2 = e2.y; t.wy0 = en.y * dot(en, C);
    }
    t.height_lo = dt_.height_lo; t.height_hi = dt_.height_hi; t.ang_max = dt_.ang_max;
    t.alive_bonus = dt_.alive_bonus; t.ctrl_cost = dt_.ctrl_cost; t.vel_weight = dt_.vel_weight;
    t.limit_pen_dof = dt_.limit_pen_dof;
    if (t.limit_pen_dof >= nb) return "limit_pen_dof out of range";
    t.limit_pen_margin = dt_.limit_pen_margin; t.limit_pen_weight = dt_.limit_pen_weight;
    t.dev_cost = dt_.dev_cost;
    t.zero_reward_on_blowup = dt_.zero_reward_on_blowup;
    t.fluid_force = dt_.fluid_force;
    if (t.fluid_force && nb != nbd) return "fluid force with welded bodies is not supported";
    t.fluid_offset = dt_.fluid_offset; t.fluid_coef = dt_.fluid_coef;
    t.reset_noise = dt_.reset_noise; t.state_bound = dt_.state_bound;
    t.inv_dt_env = 1.0 / (dm.dt * dt_.frame_skip);
    if (dt_.kind == DARTB_TASK_LOCOMOTION && nb < 3) return "task layer needs at least 3 dofs (q[0], q[2] are read)";
    return "";
}

template <typename R>
static void convert(const PModel<double>& a, PModel<R>& b) {
    b.nb = a.nb; b.ns = a.ns; b.dt = (R)a.dt; b.gx = (R)a.gx; b.gy = (R)a.gy;
    for (int i = 0; i < PM_MAXB; i++) {
        b.parent[i] = a.parent[i]; b.jtype[i] = a.jtype[i]; b.sgn[i] = (R)a.sgn[i];
        b.ax[i] = (R)a.ax[i]; b.ay[i] = (R)a.ay[i]; b.ux[i] = (R)a.ux[i]; b.uy[i] = (R)a.uy[i];
        b.mass[i] = (R)a.mass[i]; b.cx[i] = (R)a.cx[i]; b.cy[i] = (R)a.cy[i]; b.izz[i] = (R)a.izz[i];
        b.ox[i] = (R)a.ox[i]; b.oy[i] = (R)a.oy[i];
        b.damping[i] = (R)a.damping[i]; b.kspring[i] = (R)a.kspring[i]; b.rest[i] = (R)a.rest[i];
        b.coulomb[i] = (R)a.coulomb[i];
        b.qlo[i] = (R)a.qlo[i]; b.qhi[i] = (R)a.qhi[i]; b.limited[i] = a.limited[i];
        b.qinit[i] = (R)a.qinit[i]; b.dqinit[i] = (R)a.dqinit[i];
        b.fnx[i] = (R)a.fnx[i]; b.fny[i] = (R)a.fny[i]; b.orig_body[i] = a.orig_body[i];
    }
    b.nbd = a.nbd;
    for (int i = 0; i < PM_MAXD; i++) { b.dgroup[i] = a.dgroup[i]; b.dox[i] = (R)a.dox[i]; b.doy[i] = (R)a.doy[i]; }
    for (int i = 0; i < PM_MAXS; i++) {
        b.sbody[i] = a.sbody[i]; b.sorig[i] = a.sorig[i];
        b.scx[i] = (R)a.scx[i]; b.scy[i] = (R)a.scy[i]; b.sdx[i] = (R)a.sdx[i]; b.sdy[i] = (R)a.sdy[i];
        b.shalf[i] = (R)a.shalf[i]; b.srad[i] = (R)a.srad[i]; b.smu[i] = (R)a.smu[i];
    }
    b.any_coulomb = a.any_coulomb;
    b.has_ground = a.has_ground;
    b.gcx = (R)a.gcx; b.gcy = (R)a.gcy; b.ghx = (R)a.ghx; b.ghy = (R)a.ghy;
    b.gupx = (R)a.gupx; b.gupy = (R)a.gupy; b.ghup = (R)a.ghup;
    for (int k = 0; k < 3; k++) { b.e1[k] = (R)a.e1[k]; b.e2[k] = (R)a.e2[k]; b.en[k] = (R)a.en[k]; }
    b.hz = (R)a.hz;
}
template <typename R>
static void convert(const PTask<double>& a, PTask<R>& b) {
    b.frame_skip = a.frame_skip; b.n_act = a.n_act; b.n_obs = a.n_obs;
    for (int i = 0; i < PM_MAXB; i++) {
        b.dof_act[i] = a.dof_act[i]; b.dof_scale[i] = (R)a.dof_scale[i];
        b.dof_lo[i] = (R)a.dof_lo[i]; b.dof_hi[i] = (R)a.dof_hi[i];
    }
    b.obs_mode = a.obs_mode; b.dq_clip = (R)a.dq_clip; b.height_body = a.height_body;
    b.hcx = (R)a.hcx; b.hcy = (R)a.hcy; b.wy1 = (R)a.wy1; b.wy2 = (R)a.wy2; b.wy0 = (R)a.wy0;
    b.height_lo = (R)a.height_lo; b.height_hi = (R)a.height_hi; b.ang_max = (R)a.ang_max;
    b.alive_bonus = (R)a.alive_bonus; b.ctrl_cost = (R)a.ctrl_cost; b.vel_weight = (R)a.vel_weight;
    b.limit_pen_dof = a.limit_pen_dof; b.limit_pen_margin = (R)a.limit_pen_margin;
    b.limit_pen_weight = (R)a.limit_pen_weight; b.dev_cost = (R)a.dev_cost;
    b.zero_reward_on_blowup = a.zero_reward_on_blowup; b.fluid_force = a.fluid_force;
    b.fluid_offset = (R)a.fluid_offset; b.fluid_coef = (R)a.fluid_coef;
    b.reset_noise = (R)a.reset_noise; b.state_bound = (R)a.state_bound; b.inv_dt_env = (R)a.inv_dt_env;
    b.kind = a.kind; b.noise_dq = (R)a.noise_dq;
    for (int k = 0; k < 2; k++) {
        b.probe_body[k] = a.probe_body[k]; b.probe_x[k] = (R)a.probe_x[k]; b.probe_y[k] = (R)a.probe_y[k]; b.probe_n[k] = (R)a.probe_n[k];
    }
}

// Per-world dynamics parameters (dartb_set_body_params): lowers every world's model with its own bodynode masses / friction
// coefficients ([n][n_bodies], either may be null) and tabulates what the loop kernels read per world:
// out[(k nb + i) n + w], k = 0..3: mass, cx, cy, izz of planar body i;  out[(4 nb + s) n + w]: friction of capsule s.
// Returns "" or why a world cannot be lowered to the topology `signature` of the shared model.
static std::string body_param_table(const dartb_model_t& base, const dartb_task_t& task, int n, const double* mass, const double* mu,
                                    const std::string& signature, int nb, int ns, std::vector<double>& out) {
    const int nbd = base.n_bodies;
    out.assign((size_t)(4 * nb + ns) * n, 0.0);
    dartb_model_t dm = base;
    Result res;
    for (int w = 0; w < n; w++) {
        for (int i = 0; i < nbd; i++) {
            if (mass) dm.bodies[i].mass = mass[(size_t)w * nbd + i];
            if (mu) dm.bodies[i].friction_coeff = mu[(size_t)w * nbd + i];
        }
        std::string why = lower_model(dm, task, res);
        if (!why.empty()) return "world " + std::to_string(w) + " cannot be lowered: " + why;
        if (res.signature != signature || res.m.nb != nb || res.m.ns != ns)
            return "world " + std::to_string(w) + " lowers to a different topology";
        for (int i = 0; i < nb; i++) {
            out[(size_t)i * n + w] = res.m.mass[i];
            out[(size_t)(nb + i) * n + w] = res.m.cx[i];
            out[(size_t)(2 * nb + i) * n + w] = res.m.cy[i];
            out[(size_t)(3 * nb + i) * n + w] = res.m.izz[i];
        }
        for (int k = 0; k < ns; k++) out[(size_t)(4 * nb + k) * n + w] = res.m.smu[k];
    }
    return "";
}

}  // namespace lower
