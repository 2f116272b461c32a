/*
 * dart_oracle.c — CPU restatement (fp64, scalar C) of the DART 6 / pydart2 time step that
 * sits under gym/envs/dart/dart_env.py:170-175 of the reference.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load it.  The product path
 * (dart_env_b200/ -> libdartb.so) never links, imports or calls anything in oracle/.
 *
 * PARITY UNPINNED: the arithmetic restated here lives in third-party code that is NOT in
 * /root/reference and is not installed in this image:
 *     pydart2 (github.com/sehoonha/pydart2, unpinned; reference tox.ini:52-55)
 *     DART 6.x (libdart6-all-dev, unpinned; reference test.dockerfile:27-29)
 *     ODE (narrow phase dCollideCapsuleBox + Dantzig dSolveLCP, bundled with DART)
 * and the reference holds no golden vectors for it (gym/envs/tests/rollout.json = {}).
 * The restatement follows the published algorithms as summarised in SURVEY.md Appendix B and
 * is anchored on the reference's own call sites:
 *     World.step()                 dart_env.py:175, snake_7link.py:50
 *     Skeleton.set_forces(tau)     dart_env.py:174
 *     set_positions/velocities     dart_env.py:147-148
 *     q / dq                       dart_env.py:213-214
 *     bodynodes[i].com()           hopper.py:42
 *     bn.com_spatial_velocity(), to_world(), add_ext_force()   snake_7link.py:37-47
 *     collision_result.contacts[*].force                        walker2d.py:38-41
 *     set_position_limit_enforced  dart_env.py:64-67
 * The task layer (obs / reward / done / reset) restates hopper.py:24-84, walker2d.py:22-82,
 * half_cheetah.py:27-101, snake_7link.py:35-122 and IS pinned: tests/golden/ npz files are produced
 * by running those reference classes unmodified on top of this physics (oracle/pydart2_shim).
 *
 * Step semantics (DART World::step, Appendix B.3):
 *   1. forward dynamics by the Articulated Body Algorithm, gravity as a body force, joint
 *      damping/spring folded implicitly into the projected articulated inertia     (B.4)
 *   2. dq += dt * ddq
 *   3. collide robot shapes against world-fixed boxes (ODE capsule-box: ONE contact) (B.5)
 *   4. contact rows [n, t1, t2] + joint-limit rows, A = J M^-1 J^T (+CFM), b         (B.6)
 *   5. boxed LCP: Dantzig pivoting with ODE's friction handling (bounds fixed from the
 *      frictionless normal impulses), or PGS with a fixed sweep count                (B.7)
 *   6. dq += M^-1 J^T x ; q += dt * dq ; clear tau and external forces               (B.8)
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/dartb.h"

#define NB DARTB_MAX_BODIES
#define MAXC DARTB_MAX_SHAPES /* at most one contact per robot shape / ground shape pair kept */
#define MAXROWS (3 * MAXC + 2 * NB)

/* DART constants (ContactConstraint.cpp / JointLimitConstraint.cpp) */
#define CONTACT_ERP 0.01
#define CONTACT_MAX_ERV 1e-3
#define CONTACT_CFM 1e-5
#define CONTACT_EPS 1e-6
#define FRICTION_THRESHOLD 1e-3
#define LIMIT_CFM 1e-9
#define INERT_DIAG 1e-14 /* rows whose A_ii is below this never move anything: x = 0 */

/* ------------------------------------------------------------------ small linear algebra */
typedef struct { double R[9]; double p[3]; } xf_t; /* rigid transform */

static void v3cross(const double a[3], const double b[3], double o[3]) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static double v3dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void m3v(const double R[9], const double v[3], double o[3]) {
    double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void m3tv(const double R[9], const double v[3], double o[3]) {
    double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void m3m(const double A[9], const double B[9], double O[9]) {
    double T[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(O, T, sizeof T);
}
static void xf_from12(const double t[12], xf_t* o) {
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) o->R[3 * i + j] = t[4 * i + j];
        o->p[i] = t[4 * i + 3];
    }
}
static void xf_mul(const xf_t* a, const xf_t* b, xf_t* o) {
    xf_t t;
    m3m(a->R, b->R, t.R);
    m3v(a->R, b->p, t.p);
    for (int i = 0; i < 3; i++) t.p[i] += a->p[i];
    *o = t;
}
static void xf_inv(const xf_t* a, xf_t* o) {
    xf_t t;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t.R[3 * i + j] = a->R[3 * j + i];
    m3v(t.R, a->p, t.p);
    for (int i = 0; i < 3; i++) t.p[i] = -t.p[i];
    *o = t;
}
static void xf_point(const xf_t* a, const double p[3], double o[3]) {
    double t[3];
    m3v(a->R, p, t);
    for (int i = 0; i < 3; i++) o[i] = t[i] + a->p[i];
}
/* rotation by angle about unit axis (Rodrigues; DART math::expAngular) */
static void rot_axis(const double a[3], double th, double R[9]) {
    double c = cos(th), s = sin(th), v = 1.0 - c;
    R[0] = c + a[0] * a[0] * v;        R[1] = a[0] * a[1] * v - a[2] * s; R[2] = a[0] * a[2] * v + a[1] * s;
    R[3] = a[1] * a[0] * v + a[2] * s; R[4] = c + a[1] * a[1] * v;        R[5] = a[1] * a[2] * v - a[0] * s;
    R[6] = a[2] * a[0] * v - a[1] * s; R[7] = a[2] * a[1] * v + a[0] * s; R[8] = c + a[2] * a[2] * v;
}

/* spatial vectors: [angular(3); linear(3)], body frame (DART convention) */
/* motion transform parent->child, T = pose of child in parent */
static void sx_motion(const xf_t* T, const double V[6], double O[6]) {
    double t[3], w[3];
    v3cross(V, T->p, t); /* w x p */
    for (int i = 0; i < 3; i++) t[i] += V[3 + i];
    m3tv(T->R, V, w);
    m3tv(T->R, t, t);
    O[0] = w[0]; O[1] = w[1]; O[2] = w[2]; O[3] = t[0]; O[4] = t[1]; O[5] = t[2];
}
/* force transform child->parent */
static void sx_force_T(const xf_t* T, const double F[6], double O[6]) {
    double n[3], f[3], c[3];
    m3v(T->R, F, n);
    m3v(T->R, F + 3, f);
    v3cross(T->p, f, c);
    O[0] = n[0] + c[0]; O[1] = n[1] + c[1]; O[2] = n[2] + c[2]; O[3] = f[0]; O[4] = f[1]; O[5] = f[2];
}
static void sx_cross_motion(const double V[6], const double M[6], double O[6]) {
    double a[3], b[3], c[3];
    v3cross(V, M, a);
    v3cross(V, M + 3, b);
    v3cross(V + 3, M, c);
    O[0] = a[0]; O[1] = a[1]; O[2] = a[2]; O[3] = b[0] + c[0]; O[4] = b[1] + c[1]; O[5] = b[2] + c[2];
}
static void sx_cross_force(const double V[6], const double F[6], double O[6]) {
    double a[3], b[3], c[3];
    v3cross(V, F, a);
    v3cross(V + 3, F + 3, b);
    v3cross(V, F + 3, c);
    O[0] = a[0] + b[0]; O[1] = a[1] + b[1]; O[2] = a[2] + b[2]; O[3] = c[0]; O[4] = c[1]; O[5] = c[2];
}
static void m6v(const double M[36], const double v[6], double o[6]) {
    double t[6];
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int j = 0; j < 6; j++) s += M[6 * i + j] * v[j];
        t[i] = s;
    }
    memcpy(o, t, sizeof t);
}
/* 6x6 motion-transform matrix X (parent->child) for T = pose of child in parent */
static void sx_matrix(const xf_t* T, double X[36]) {
    double Rt[9], px[9], B[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rt[3 * i + j] = T->R[3 * j + i];
    px[0] = 0; px[1] = -T->p[2]; px[2] = T->p[1];
    px[3] = T->p[2]; px[4] = 0; px[5] = -T->p[0];
    px[6] = -T->p[1]; px[7] = T->p[0]; px[8] = 0;
    m3m(Rt, px, B); /* R^T [p]x ; lower-left = -R^T [p]x */
    memset(X, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            X[6 * i + j] = Rt[3 * i + j];
            X[6 * (i + 3) + (j + 3)] = Rt[3 * i + j];
            X[6 * (i + 3) + j] = -B[3 * i + j];
        }
}
/* O += X^T M X */
static void m6_congruence_add(const double X[36], const double M[36], double O[36]) {
    double T[36];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += M[6 * i + k] * X[6 * k + j];
            T[6 * i + j] = s;
        }
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += X[6 * k + i] * T[6 * k + j];
            O[6 * i + j] += s;
        }
}

/* ------------------------------------------------------------------ world */
typedef struct {
    int body;          /* robot body index */
    int shape;         /* robot shape index */
    double point[3], normal[3], depth;
    double force[3];   /* impulse/dt after the solve (pydart2 contact.force) */
    int nrows;         /* 3 with friction, 1 without */
    double mu;
} contact_t;

typedef struct orc_world {
    dartb_model_t m;
    int nb, nd;
    int dof_body[NB];
    double q[NB], dq[NB], tau[NB], ddq[NB];
    double fext[NB][6];          /* body-frame spatial external force accumulators */
    /* per-body constants */
    xf_t Tpj[NB], Tcj_inv[NB];
    double S[NB][6];             /* joint motion subspace in the child body frame */
    double I6[NB][36];           /* spatial inertia, body frame */
    /* per-step kinematics */
    xf_t Trel[NB], Tw[NB];
    double V[NB][6], eta[NB][6];
    /* articulated quantities: [0] = implicit (forward dynamics), [1] = plain (impulses) */
    double IA[2][NB][36], U[2][NB][6], Dinv[2][NB];
    /* contacts of the last step */
    int ncontacts;
    contact_t contacts[MAXC];
    /* last LCP (for inspection) */
    int nrows;
    double lcpA[MAXROWS * MAXROWS], lcpx[MAXROWS], lcpb[MAXROWS], lcplo[MAXROWS], lcphi[MAXROWS];
    int lcpfindex[MAXROWS];
    int limit_active[NB];        /* per dof: 0 none, -1 lower, +1 upper (last step) */
    double shape_gap[MAXC];      /* per robot shape: min over statics of (distance - radius) */
    double shape_tilt[MAXC];     /* per robot shape: |height difference of the capsule ends| */
    /* options */
    int lcp_mode;                /* 0 dantzig, 1 pgs */
    int pgs_iters;
    double time;
    long frame;
    int lcp_fail;
} orc_world_t;

static void build_constants(orc_world_t* w) {
    const dartb_model_t* m = &w->m;
    w->nb = m->n_bodies;
    w->nd = m->n_dofs;
    for (int i = 0; i < w->nb; i++) {
        const dartb_body_t* b = &m->bodies[i];
        xf_t Tcj;
        xf_from12(b->T_parent_joint, &w->Tpj[i]);
        xf_from12(b->T_child_joint, &Tcj);
        xf_inv(&Tcj, &w->Tcj_inv[i]);
        double a[3];
        m3v(Tcj.R, b->axis, a);
        memset(w->S[i], 0, sizeof w->S[i]);
        if (b->joint_type == DARTB_JOINT_REVOLUTE) {
            double pv[3];
            v3cross(Tcj.p, a, pv);
            for (int k = 0; k < 3; k++) { w->S[i][k] = a[k]; w->S[i][3 + k] = pv[k]; }
        } else if (b->joint_type == DARTB_JOINT_PRISMATIC) {
            for (int k = 0; k < 3; k++) w->S[i][3 + k] = a[k];
        }
        if (b->dof >= 0) w->dof_body[b->dof] = i;
        /* spatial inertia about the body origin */
        double* I = w->I6[i];
        memset(I, 0, 36 * sizeof(double));
        const double* c = b->com;
        double mass = b->mass, cc = v3dot(c, c);
        for (int r = 0; r < 3; r++)
            for (int s = 0; s < 3; s++)
                I[6 * r + s] = b->inertia[3 * r + s] + mass * ((r == s ? cc : 0.0) - c[r] * c[s]);
        double cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
        for (int r = 0; r < 3; r++)
            for (int s = 0; s < 3; s++) {
                I[6 * r + (3 + s)] = mass * cx[3 * r + s];
                I[6 * (3 + r) + s] = -mass * cx[3 * r + s];
            }
        for (int r = 0; r < 3; r++) I[6 * (3 + r) + (3 + r)] = mass;
    }
}

orc_world_t* orc_create(const dartb_model_t* model) {
    orc_world_t* w = (orc_world_t*)calloc(1, sizeof(orc_world_t));
    w->m = *model;
    build_constants(w);
    w->pgs_iters = 30;
    for (int d = 0; d < w->nd; d++) {
        w->q[d] = model->bodies[w->dof_body[d]].q_init;
        w->dq[d] = model->bodies[w->dof_body[d]].dq_init;
    }
    return w;
}
void orc_destroy(orc_world_t* w) { free(w); }
int orc_num_dofs(const orc_world_t* w) { return w->nd; }
int orc_num_bodies(const orc_world_t* w) { return w->nb; }
void orc_set_option(orc_world_t* w, int key, double val) {
    if (key == DARTB_OPT_LCP_MODE) w->lcp_mode = (int)val;
    else if (key == DARTB_OPT_PGS_ITERS) w->pgs_iters = (int)val;
    else if (key == DARTB_OPT_FRICTION_ALL)
        for (int i = 0; i < w->nb; i++) w->m.bodies[i].friction_coeff = val;
}
void orc_set_mass(orc_world_t* w, int body, double mass) {
    /* bn.set_mass (snake_7link.py:117) -> DART 6 BodyNode::setMass -> Inertia::setMass: the mass changes, the moment of
     * inertia about the COM (the <moment_of_inertia> of the .skel, or the shape-derived one computed at load) stays, and
     * the spatial tensor is rebuilt from both (DART is a pydart2 dependency absent from the reference tree: published
     * behaviour restated) */
    w->m.bodies[body].mass = mass;
    build_constants(w);
}
void orc_set_friction(orc_world_t* w, int body, double mu) { w->m.bodies[body].friction_coeff = mu; }
void orc_set_state(orc_world_t* w, const double* q, const double* dq) {
    if (q) memcpy(w->q, q, w->nd * sizeof(double));
    if (dq) memcpy(w->dq, dq, w->nd * sizeof(double));
}
void orc_get_state(const orc_world_t* w, double* q, double* dq) {
    if (q) memcpy(q, w->q, w->nd * sizeof(double));
    if (dq) memcpy(dq, w->dq, w->nd * sizeof(double));
}
void orc_set_forces(orc_world_t* w, const double* tau) { memcpy(w->tau, tau, w->nd * sizeof(double)); }

/* ------------------------------------------------------------------ kinematics (B.4 pass 1) */
static void forward_kinematics(orc_world_t* w) {
    for (int i = 0; i < w->nb; i++) {
        const dartb_body_t* b = &w->m.bodies[i];
        xf_t J;
        memset(&J, 0, sizeof J);
        J.R[0] = J.R[4] = J.R[8] = 1.0;
        double qd = 0.0;
        if (b->dof >= 0) qd = w->dq[b->dof];
        if (b->joint_type == DARTB_JOINT_REVOLUTE) rot_axis(b->axis, w->q[b->dof], J.R);
        else if (b->joint_type == DARTB_JOINT_PRISMATIC)
            for (int k = 0; k < 3; k++) J.p[k] = b->axis[k] * w->q[b->dof];
        xf_t t;
        xf_mul(&w->Tpj[i], &J, &t);
        xf_mul(&t, &w->Tcj_inv[i], &w->Trel[i]);
        if (b->parent >= 0) xf_mul(&w->Tw[b->parent], &w->Trel[i], &w->Tw[i]);
        else w->Tw[i] = w->Trel[i];
        double Vp[6] = {0, 0, 0, 0, 0, 0}, Sdq[6];
        if (b->parent >= 0) sx_motion(&w->Trel[i], w->V[b->parent], Vp);
        for (int k = 0; k < 6; k++) { Sdq[k] = w->S[i][k] * qd; w->V[i][k] = Vp[k] + Sdq[k]; }
        sx_cross_motion(w->V[i], Sdq, w->eta[i]); /* partial acceleration, dS/dt = 0 */
    }
}

/* articulated inertias; which = 0 implicit (D += dt*d + dt^2*k), 1 plain */
static void articulated_inertia(orc_world_t* w, int which) {
    double dt = w->m.dt;
    for (int i = 0; i < w->nb; i++) memcpy(w->IA[which][i], w->I6[i], 36 * sizeof(double));
    for (int i = w->nb - 1; i >= 0; i--) {
        const dartb_body_t* b = &w->m.bodies[i];
        double* IA = w->IA[which][i];
        double Pi[36];
        memcpy(Pi, IA, sizeof Pi);
        if (b->dof >= 0) {
            double* U = w->U[which][i];
            m6v(IA, w->S[i], U);
            double D = 0;
            for (int k = 0; k < 6; k++) D += w->S[i][k] * U[k];
            if (which == 0) D += dt * b->damping + dt * dt * b->spring_k;
            w->Dinv[which][i] = 1.0 / D;
            for (int r = 0; r < 6; r++)
                for (int s = 0; s < 6; s++) Pi[6 * r + s] -= U[r] * U[s] * w->Dinv[which][i];
        }
        if (b->parent >= 0) {
            double X[36];
            sx_matrix(&w->Trel[i], X);
            m6_congruence_add(X, Pi, w->IA[which][b->parent]);
        }
    }
}

/* ddq = M^-1 rhs using the factors `which` (pure joint-space solve, no bias terms) */
static void minv_mul(const orc_world_t* w, int which, const double* rhs, double* out) {
    double pA[NB][6], u[NB], a[NB][6];
    memset(pA, 0, sizeof pA);
    for (int i = w->nb - 1; i >= 0; i--) {
        const dartb_body_t* b = &w->m.bodies[i];
        double pa[6];
        memcpy(pa, pA[i], sizeof pa);
        if (b->dof >= 0) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += w->S[i][k] * pA[i][k];
            u[i] = rhs[b->dof] - s;
            for (int k = 0; k < 6; k++) pa[k] += w->U[which][i][k] * u[i] * w->Dinv[which][i];
        }
        if (b->parent >= 0) {
            double f[6];
            sx_force_T(&w->Trel[i], pa, f);
            for (int k = 0; k < 6; k++) pA[b->parent][k] += f[k];
        }
    }
    for (int i = 0; i < w->nb; i++) {
        const dartb_body_t* b = &w->m.bodies[i];
        double ap[6] = {0, 0, 0, 0, 0, 0};
        if (b->parent >= 0) sx_motion(&w->Trel[i], a[b->parent], ap);
        memcpy(a[i], ap, sizeof ap);
        if (b->dof >= 0) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += w->U[which][i][k] * ap[k];
            double dd = w->Dinv[which][i] * (u[i] - s);
            out[b->dof] = dd;
            for (int k = 0; k < 6; k++) a[i][k] += w->S[i][k] * dd;
        }
    }
}

/* full forward dynamics (B.4): fills w->ddq. Assumes forward_kinematics done. */
static void forward_dynamics(orc_world_t* w) {
    double dt = w->m.dt;
    articulated_inertia(w, 0);
    double pA[NB][6], u[NB], a[NB][6];
    for (int i = 0; i < w->nb; i++) {
        /* bias force: V x* (I V) - f_gravity - f_ext */
        double IV[6], g_b[3], fg[6], gv[6] = {0, 0, 0, 0, 0, 0};
        m6v(w->I6[i], w->V[i], IV);
        sx_cross_force(w->V[i], IV, pA[i]);
        m3tv(w->Tw[i].R, w->m.gravity, g_b);
        gv[3] = g_b[0]; gv[4] = g_b[1]; gv[5] = g_b[2];
        m6v(w->I6[i], gv, fg); /* I * [0; R^T g] */
        for (int k = 0; k < 6; k++) pA[i][k] -= fg[k] + w->fext[i][k];
    }
    for (int i = w->nb - 1; i >= 0; i--) {
        const dartb_body_t* b = &w->m.bodies[i];
        double pa[6], t6[6];
        m6v(w->IA[0][i], w->eta[i], t6);
        if (b->dof >= 0) {
            int d = b->dof;
            double s = 0;
            for (int k = 0; k < 6; k++) s += w->S[i][k] * (pA[i][k] + t6[k]);
            double spring = -b->spring_k * (w->q[d] - b->spring_rest + dt * w->dq[d]);
            double damp = -b->damping * w->dq[d];
            u[i] = w->tau[d] + spring + damp - s;
            for (int k = 0; k < 6; k++) pa[k] = pA[i][k] + t6[k] + w->U[0][i][k] * u[i] * w->Dinv[0][i];
        } else {
            for (int k = 0; k < 6; k++) pa[k] = pA[i][k] + t6[k];
        }
        if (b->parent >= 0) {
            double f[6];
            sx_force_T(&w->Trel[i], pa, f);
            for (int k = 0; k < 6; k++) pA[b->parent][k] += f[k];
        }
    }
    for (int i = 0; i < w->nb; i++) {
        const dartb_body_t* b = &w->m.bodies[i];
        double ap[6] = {0, 0, 0, 0, 0, 0};
        if (b->parent >= 0) sx_motion(&w->Trel[i], a[b->parent], ap);
        if (b->dof >= 0) {
            /* u already contains -S^T IA eta; ddq = Dinv (u - U^T ap) */
            double s = 0;
            for (int k = 0; k < 6; k++) s += w->U[0][i][k] * ap[k];
            double dd = w->Dinv[0][i] * (u[i] - s);
            w->ddq[b->dof] = dd;
            for (int k = 0; k < 6; k++) a[i][k] = ap[k] + w->eta[i][k] + w->S[i][k] * dd;
        } else {
            for (int k = 0; k < 6; k++) a[i][k] = ap[k] + w->eta[i][k];
        }
    }
}

/* ------------------------------------------------------------------ collision (B.5) */
/* Closest points between segment p1->p2 and a box (centre c, rotation R, full sides).
 * Restates ODE's dClosestLineBoxPoints: exact minimiser of the convex piecewise-quadratic
 * distance along the segment; ties (segment parallel to the nearest face) resolve to t = 0. */
static void closest_segment_box(const double p1[3], const double p2[3], const double c[3],
                                const double R[9], const double side[3], double lret[3], double bret[3]) {
    double tmp[3], s[3], v[3], sign[3], v2[3], h[3], tanchor[3], d12[3];
    int region[3];
    for (int i = 0; i < 3; i++) tmp[i] = p1[i] - c[i];
    m3tv(R, tmp, s);
    for (int i = 0; i < 3; i++) d12[i] = p2[i] - p1[i];
    m3tv(R, d12, v);
    for (int i = 0; i < 3; i++) {
        if (v[i] < 0) { s[i] = -s[i]; v[i] = -v[i]; sign[i] = -1; } else sign[i] = 1;
        v2[i] = v[i] * v[i];
        h[i] = 0.5 * side[i];
    }
    const double tanchor_eps = 1e-19;
    for (int i = 0; i < 3; i++) {
        if (v[i] > tanchor_eps) {
            if (s[i] < -h[i]) { region[i] = -1; tanchor[i] = (-h[i] - s[i]) / v[i]; }
            else { region[i] = (s[i] > h[i]); tanchor[i] = (h[i] - s[i]) / v[i]; }
        } else { region[i] = 0; tanchor[i] = 2; }
    }
    double t = 0, dd2dt = 0;
    for (int i = 0; i < 3; i++) dd2dt -= (region[i] ? v2[i] : 0) * tanchor[i];
    if (dd2dt < 0) {
        int done = 0;
        do {
            double next_t = 1;
            for (int i = 0; i < 3; i++)
                if (tanchor[i] > t && tanchor[i] < 1 && tanchor[i] < next_t) next_t = tanchor[i];
            double next_dd2dt = 0;
            for (int i = 0; i < 3; i++) next_dd2dt += (region[i] ? v2[i] : 0) * (next_t - tanchor[i]);
            if (next_dd2dt >= 0) {
                double mm = (next_dd2dt - dd2dt) / (next_t - t);
                t -= dd2dt / mm;
                done = 1;
                break;
            }
            for (int i = 0; i < 3; i++)
                if (tanchor[i] == next_t) { tanchor[i] = (h[i] - s[i]) / v[i]; region[i]++; }
            t = next_t;
            dd2dt = next_dd2dt;
        } while (t < 1);
        if (!done) t = 1;
    }
    for (int i = 0; i < 3; i++) lret[i] = p1[i] + t * d12[i];
    for (int i = 0; i < 3; i++) {
        tmp[i] = sign[i] * (s[i] + t * v[i]);
        if (tmp[i] < -h[i]) tmp[i] = -h[i];
        else if (tmp[i] > h[i]) tmp[i] = h[i];
    }
    m3v(R, tmp, s);
    for (int i = 0; i < 3; i++) bret[i] = s[i] + c[i];
}

/* sphere (centre pl, radius r) against point pb: ODE dCollideSpheres with r2 = 0 */
static int sphere_point_contact(const double pl[3], double r, const double pb[3], double pos[3],
                                double normal[3], double* depth) {
    double dv[3] = {pl[0] - pb[0], pl[1] - pb[1], pl[2] - pb[2]};
    double d = sqrt(v3dot(dv, dv));
    if (d < 1e-15) return -1; /* ODE dCollideCapsuleBox: mindist (double build) -> box-box fallback */
    if (d > r) return 0;
    for (int i = 0; i < 3; i++) normal[i] = dv[i] / d;
    double k = 0.5 * (-r - d);
    for (int i = 0; i < 3; i++) pos[i] = pl[i] + normal[i] * k;
    *depth = r - d;
    return 1;
}

static void collide(orc_world_t* w) {
    w->ncontacts = 0;
    for (int si = 0; si < w->m.n_shapes; si++) {
        const dartb_shape_t* sh = &w->m.shapes[si];
        xf_t Tl, Ts;
        xf_from12(sh->T, &Tl);
        xf_mul(&w->Tw[sh->body], &Tl, &Ts);
        w->shape_gap[si] = INFINITY;
        w->shape_tilt[si] = INFINITY;
        for (int gi = 0; gi < w->m.n_ground; gi++) {
            const dartb_shape_t* g = &w->m.ground[gi];
            if (g->type != DARTB_SHAPE_BOX) continue; /* only box statics collide (in scope) */
            xf_t Tg;
            xf_from12(g->T, &Tg);
            double pl[3], pb[3], radius;
            if (sh->type == DARTB_SHAPE_CAPSULE) {
                /* ODE dCollideCapsuleBox: p1 = c + (h/2) z, p2 = c - (h/2) z */
                double ax[3] = {Ts.R[2], Ts.R[5], Ts.R[8]}, p1[3], p2[3], hl = 0.5 * sh->size[1];
                for (int k = 0; k < 3; k++) { p1[k] = Ts.p[k] + hl * ax[k]; p2[k] = Ts.p[k] - hl * ax[k]; }
                closest_segment_box(p1, p2, Tg.p, Tg.R, g->size, pl, pb);
                radius = sh->size[0];
                double up[3] = {Tg.R[1], Tg.R[4], Tg.R[7]}, dd[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
                double tilt = fabs(v3dot(up, dd));
                if (tilt < w->shape_tilt[si]) w->shape_tilt[si] = tilt;
            } else if (sh->type == DARTB_SHAPE_SPHERE) {
                closest_segment_box(Ts.p, Ts.p, Tg.p, Tg.R, g->size, pl, pb);
                radius = sh->size[0];
            } else {
                continue; /* box / ellipsoid robot shapes: out of scope */
            }
            {
                double dv[3] = {pl[0] - pb[0], pl[1] - pb[1], pl[2] - pb[2]};
                double gap = sqrt(v3dot(dv, dv)) - radius;
                if (gap < w->shape_gap[si]) w->shape_gap[si] = gap;
            }
            if (w->ncontacts >= MAXC) continue;
            contact_t* c = &w->contacts[w->ncontacts];
            int r = sphere_point_contact(pl, radius, pb, c->point, c->normal, &c->depth);
            if (r == -1) {
                /* capsule axis inside the box (ODE falls back to box-box): push out through the
                 * nearest face of the box along its local +y. Practically unreachable here. */
                double up[3] = {Tg.R[1], Tg.R[4], Tg.R[7]}, rel[3];
                for (int k = 0; k < 3; k++) rel[k] = pl[k] - Tg.p[k];
                double hgt = v3dot(rel, up);
                for (int k = 0; k < 3; k++) { c->normal[k] = up[k]; c->point[k] = pl[k]; }
                c->depth = radius + (0.5 * g->size[1] - hgt);
                r = 1;
            }
            if (r != 1) continue;
            if (v3dot(c->normal, c->normal) < CONTACT_EPS * CONTACT_EPS) continue;
            c->body = sh->body;
            c->shape = si;
            c->force[0] = c->force[1] = c->force[2] = 0;
            w->ncontacts++;
        }
    }
}

/* ------------------------------------------------------------------ Jacobian rows */
/* d(velocity of world point P fixed on `body`)/d(dq) . dir */
static void point_jacobian_row(const orc_world_t* w, int body, const double P[3], const double dir[3],
                               double* row) {
    for (int d = 0; d < w->nd; d++) row[d] = 0;
    for (int i = body; i >= 0; i = w->m.bodies[i].parent) {
        const dartb_body_t* b = &w->m.bodies[i];
        if (b->dof < 0) continue;
        /* joint axis / origin in world: S is expressed in body i's frame */
        double aw[3], lw[3];
        m3v(w->Tw[i].R, w->S[i], aw);
        m3v(w->Tw[i].R, w->S[i] + 3, lw);
        /* velocity at P = lw + aw x (P - o_i) */
        double r[3] = {P[0] - w->Tw[i].p[0], P[1] - w->Tw[i].p[1], P[2] - w->Tw[i].p[2]}, cr[3];
        v3cross(aw, r, cr);
        row[b->dof] = (lw[0] + cr[0]) * dir[0] + (lw[1] + cr[1]) * dir[1] + (lw[2] + cr[2]) * dir[2];
    }
}

/* ------------------------------------------------------------------ LCP (B.7) */
static int chol_solve(int n, const double* A, const double* rhs, double* x) {
    /* dense Cholesky A = L L^T (A symmetric PD, n small) */
    double L[MAXROWS * MAXROWS];
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = A[i * n + j];
            for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
            if (i == j) {
                if (!(s > 0)) return -1;
                L[i * n + i] = sqrt(s);
            } else L[i * n + j] = s / L[j * n + j];
        }
    double y[MAXROWS];
    for (int i = 0; i < n; i++) {
        double s = rhs[i];
        for (int k = 0; k < i; k++) s -= L[i * n + k] * y[k];
        y[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = y[i];
        for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
        x[i] = s / L[i * n + i];
    }
    return 0;
}

/* Boxed LCP  A x = b + w,  lo <= x <= hi, complementarity; friction rows (findex >= 0) have
 * hi = mu and get bounds +-mu*x[findex] fixed when the first of them is reached (ODE dSolveLCP).
 * Restates the Dantzig driving-index loop of ODE lcp.cpp; the incremental LDL^T of A_CC is
 * replaced by a fresh dense solve (same pivots, same answer up to rounding).
 * lo/hi are modified in place like ODE does.  Returns 0 ok, 1 = gave up (s <= 0 / singular). */
int orc_solve_lcp_dantzig(int n, const double* A, double* x, const double* b, double* w,
                          double* lo, double* hi, const int* findex) {
    int order[MAXROWS], inC[MAXROWS], inN[MAXROWS], state[MAXROWS], fail = 0;
    int no = 0;
    for (int i = 0; i < n; i++) if (findex[i] < 0) order[no++] = i;
    for (int i = 0; i < n; i++) if (findex[i] >= 0) order[no++] = i;
    for (int i = 0; i < n; i++) { x[i] = 0; w[i] = 0; inC[i] = inN[i] = state[i] = 0; }
    int hit_friction = 0;
    double dx[MAXROWS], dw[MAXROWS];
    for (int oi = 0; oi < n && !fail; oi++) {
        int i = order[oi];
        if (!hit_friction && findex[i] >= 0) {
            for (int ok = oi; ok < n; ok++) {
                int k = order[ok];
                double wfk = x[findex[k]];
                if (wfk == 0) { hi[k] = 0; lo[k] = 0; }
                else { hi[k] = fabs(hi[k] * wfk); lo[k] = -hi[k]; }
            }
            hit_friction = 1;
        }
        if (!(A[i * n + i] > INERT_DIAG)) { /* inert row (zero Jacobian): contributes nothing */
            x[i] = 0; w[i] = 0; inN[i] = 1; state[i] = 0; lo[i] = 0; hi[i] = 0;
            continue;
        }
        double wi = -b[i];
        for (int j = 0; j < n; j++) if (inC[j] || inN[j]) wi += A[i * n + j] * x[j];
        w[i] = wi;
        if (lo[i] == 0 && w[i] >= 0) { inN[i] = 1; state[i] = 0; continue; }
        if (hi[i] == 0 && w[i] <= 0) { inN[i] = 1; state[i] = 1; continue; }
        if (w[i] == 0) { inC[i] = 1; continue; }
        int placed = 0;
        for (int guard = 0; guard < 10 * n + 50; guard++) {
            int dir = (w[i] <= 0) ? 1 : -1;
            double dirf = dir;
            /* dx_C = -dir * A_CC^-1 A_Ci */
            int Cidx[MAXROWS], nC = 0;
            for (int j = 0; j < n; j++) if (inC[j]) Cidx[nC++] = j;
            for (int j = 0; j < n; j++) dx[j] = 0;
            if (nC > 0) {
                double Acc[MAXROWS * MAXROWS], rhs[MAXROWS], sol[MAXROWS];
                for (int r = 0; r < nC; r++) {
                    for (int c = 0; c < nC; c++) Acc[r * nC + c] = A[Cidx[r] * n + Cidx[c]];
                    rhs[r] = -dirf * A[Cidx[r] * n + i];
                }
                if (chol_solve(nC, Acc, rhs, sol)) { fail = 1; break; }
                for (int r = 0; r < nC; r++) dx[Cidx[r]] = sol[r];
            }
            /* dw = A dx (+ A_:i dir) on N and i */
            for (int j = 0; j < n; j++) {
                if (!(inN[j] || j == i)) continue;
                double s = A[j * n + i] * dirf;
                for (int r = 0; r < nC; r++) s += A[j * n + Cidx[r]] * dx[Cidx[r]];
                dw[j] = s;
            }
            int cmd = 1, si = 0;
            double s = -w[i] / dw[i];
            if (dir > 0) {
                if (hi[i] < INFINITY) { double s2 = (hi[i] - x[i]) * dirf; if (s2 < s) { s = s2; cmd = 3; } }
            } else {
                if (lo[i] > -INFINITY) { double s2 = (lo[i] - x[i]) * dirf; if (s2 < s) { s = s2; cmd = 2; } }
            }
            for (int k = 0; k < n; k++) {
                if (!inN[k]) continue;
                if ((!state[k] && dw[k] < 0) || (state[k] && dw[k] > 0)) {
                    if (lo[k] == 0 && hi[k] == 0) continue;
                    double s2 = -w[k] / dw[k];
                    if (s2 < s) { s = s2; cmd = 4; si = k; }
                }
            }
            for (int r = 0; r < nC; r++) {
                int k = Cidx[r];
                if (dx[k] < 0 && lo[k] > -INFINITY) {
                    double s2 = (lo[k] - x[k]) / dx[k];
                    if (s2 < s) { s = s2; cmd = 5; si = k; }
                }
                if (dx[k] > 0 && hi[k] < INFINITY) {
                    double s2 = (hi[k] - x[k]) / dx[k];
                    if (s2 < s) { s = s2; cmd = 6; si = k; }
                }
            }
            if (!(s > 0.0)) {
                /* ODE: "LCP internal error, s <= 0": zero the remaining rows and stop */
                if (s < 0 || s != s) {
                    for (int ok = oi; ok < n; ok++) { x[order[ok]] = 0; w[order[ok]] = 0; }
                    fail = 1;
                    break;
                }
            }
            for (int r = 0; r < nC; r++) x[Cidx[r]] += s * dx[Cidx[r]];
            x[i] += s * dirf;
            for (int k = 0; k < n; k++) if (inN[k]) w[k] += s * dw[k];
            w[i] += s * dw[i];
            switch (cmd) {
                case 1: w[i] = 0; inC[i] = 1; break;
                case 2: x[i] = lo[i]; state[i] = 0; inN[i] = 1; break;
                case 3: x[i] = hi[i]; state[i] = 1; inN[i] = 1; break;
                case 4: w[si] = 0; inN[si] = 0; inC[si] = 1; break;
                case 5: x[si] = lo[si]; state[si] = 0; inC[si] = 0; inN[si] = 1; break;
                case 6: x[si] = hi[si]; state[si] = 1; inC[si] = 0; inN[si] = 1; break;
            }
            if (cmd <= 3) { placed = 1; break; }
        }
        if (!placed && !fail) { fail = 1; inN[i] = 1; }
    }
    return fail;
}

/* Projected Gauss-Seidel, DART PGSLCPSolver shape: natural order, rows with A_ii < 1e-9 get
 * x = 0, friction bounds +-mu*x[findex] from the CURRENT normal impulse, x0 = 0, exactly
 * `iters` sweeps (no early exit so the GPU does identical work). */
int orc_solve_lcp_pgs(int n, const double* A, double* x, const double* b, const double* lo,
                      const double* hi, const int* findex, int iters) {
    for (int i = 0; i < n; i++) x[i] = 0;
    for (int it = 0; it < iters; it++)
        for (int i = 0; i < n; i++) {
            double aii = A[i * n + i];
            if (aii < 1e-9) { x[i] = 0; continue; }
            double s = b[i];
            for (int j = 0; j < n; j++) if (j != i) s -= A[i * n + j] * x[j];
            s /= aii;
            double l = lo[i], h = hi[i];
            if (findex[i] >= 0) { h = hi[i] * x[findex[i]]; l = -h; }
            if (s > h) s = h;
            if (s < l) s = l;
            x[i] = s;
        }
    return 0;
}

/* ------------------------------------------------------------------ constraint solve (B.6-B.8) */
static void tangent_basis(const double n[3], double t1[3], double t2[3]) {
    /* DART getTangentBasisMatrixODE: t1 = normalize(z x n) (fallback x x n), t2 = n x t1 */
    double z[3] = {0, 0, 1}, xx[3] = {1, 0, 0};
    v3cross(z, n, t1);
    double nn = sqrt(v3dot(t1, t1));
    if (nn < CONTACT_EPS) { v3cross(xx, n, t1); nn = sqrt(v3dot(t1, t1)); }
    for (int k = 0; k < 3; k++) t1[k] /= nn;
    v3cross(n, t1, t2);
}

static void solve_constraints(orc_world_t* w) {
    int nd = w->nd;
    static __thread double J[MAXROWS][NB], MinvJt[MAXROWS][NB];
    double* A = w->lcpA; double* b = w->lcpb; double* lo = w->lcplo; double* hi = w->lcphi; int* findex = w->lcpfindex;
    double dirs[MAXROWS][3];
    int row_contact[MAXROWS];
    double cfm[MAXROWS];
    int n = 0;
    double inv_dt = 1.0 / w->m.dt;
    for (int ci = 0; ci < w->ncontacts; ci++) {
        contact_t* c = &w->contacts[ci];
        double mu = w->m.bodies[c->body].friction_coeff; /* min(mu_robot, mu_ground = 1) */
        if (mu > 1.0) mu = 1.0;
        c->mu = mu;
        c->nrows = (mu > FRICTION_THRESHOLD) ? 3 : 1;
        double t1[3], t2[3];
        tangent_basis(c->normal, t1, t2);
        const double* dd[3] = {c->normal, t1, t2};
        int base = n;
        for (int r = 0; r < c->nrows; r++) {
            for (int k = 0; k < 3; k++) dirs[n][k] = dd[r][k];
            point_jacobian_row(w, c->body, c->point, dirs[n], J[n]);
            double v = 0;
            for (int d = 0; d < nd; d++) v += J[n][d] * w->dq[d];
            b[n] = -v;
            if (r == 0) {
                double bounce = c->depth; /* error allowance 0 */
                if (bounce < 0) bounce = 0;
                else { bounce *= inv_dt * CONTACT_ERP; if (bounce > CONTACT_MAX_ERV) bounce = CONTACT_MAX_ERV; }
                b[n] += bounce;
                lo[n] = 0; hi[n] = INFINITY; findex[n] = -1;
            } else { lo[n] = -mu; hi[n] = mu; findex[n] = base; }
            cfm[n] = CONTACT_CFM;
            row_contact[n] = ci;
            n++;
        }
    }
    for (int d = 0; d < nd; d++) {
        const dartb_body_t* bd = &w->m.bodies[w->dof_body[d]];
        w->limit_active[d] = 0;
        if (!bd->limit_enforced) continue;
        /* JointLimitConstraint::update — q is the position BEFORE this step's integration */
        double viol = w->q[d] - bd->q_lo;
        int act = 0;
        if (viol <= 0.0) act = -1;
        else { viol = w->q[d] - bd->q_hi; if (viol >= 0.0) act = 1; }
        if (!act) continue;
        for (int k = 0; k < nd; k++) J[n][k] = 0;
        J[n][d] = 1.0;
        b[n] = -w->dq[d];
        if (act < 0) { lo[n] = 0; hi[n] = INFINITY; } else { lo[n] = -INFINITY; hi[n] = 0; }
        findex[n] = -1;
        cfm[n] = LIMIT_CFM;
        row_contact[n] = -1;
        w->limit_active[d] = act;
        n++;
    }
    /* JointCoulombFrictionConstraint: one row per dof with Coulomb friction and non-zero velocity,
     * b = -dq, impulse bounds +-friction*dt ("Coulomb friction is force not impulse"), CFM 1e-9.
     * DART activates them after the joint-limit constraints (ConstraintSolver::updateConstraints). */
    for (int d = 0; d < nd; d++) {
        const dartb_body_t* bd = &w->m.bodies[w->dof_body[d]];
        if (bd->coulomb == 0.0 || w->dq[d] == 0.0 || n >= MAXROWS) continue;
        for (int k = 0; k < nd; k++) J[n][k] = 0;
        J[n][d] = 1.0;
        b[n] = -w->dq[d];
        hi[n] = bd->coulomb * w->m.dt; lo[n] = -hi[n];
        findex[n] = -1;
        cfm[n] = LIMIT_CFM;
        row_contact[n] = -1;
        n++;
    }
    w->nrows = n;
    if (n == 0) return;
    articulated_inertia(w, 1);
    for (int r = 0; r < n; r++) minv_mul(w, 1, J[r], MinvJt[r]);
    for (int r = 0; r < n; r++)
        for (int s = 0; s < n; s++) {
            double v = 0;
            for (int d = 0; d < nd; d++) v += J[s][d] * MinvJt[r][d];
            A[r * n + s] = v;
        }
    for (int r = 0; r < n; r++) A[r * n + r] *= (1.0 + cfm[r]);
    double wv[MAXROWS];
    if (w->lcp_mode == 1) orc_solve_lcp_pgs(n, A, w->lcpx, b, lo, hi, findex, w->pgs_iters);
    else w->lcp_fail |= orc_solve_lcp_dantzig(n, A, w->lcpx, b, wv, lo, hi, findex);
    /* apply impulses */
    for (int r = 0; r < n; r++) {
        double xr = w->lcpx[r];
        if (xr == 0) continue;
        for (int d = 0; d < nd; d++) w->dq[d] += MinvJt[r][d] * xr;
        if (row_contact[r] >= 0)
            for (int k = 0; k < 3; k++) w->contacts[row_contact[r]].force[k] += dirs[r][k] * xr * inv_dt;
    }
}

/* ------------------------------------------------------------------ World::step */
void orc_step(orc_world_t* w) {
    double dt = w->m.dt;
    forward_kinematics(w);
    forward_dynamics(w);
    for (int d = 0; d < w->nd; d++) w->dq[d] += dt * w->ddq[d];
    /* body velocities are only needed through J*dq below; positions unchanged */
    collide(w);
    solve_constraints(w);
    for (int d = 0; d < w->nd; d++) w->q[d] += dt * w->dq[d];
    memset(w->tau, 0, sizeof w->tau);
    memset(w->fext, 0, sizeof w->fext);
    w->time += dt;
    w->frame++;
}

void orc_reset(orc_world_t* w) {
    for (int d = 0; d < w->nd; d++) {
        w->q[d] = w->m.bodies[w->dof_body[d]].q_init;
        w->dq[d] = w->m.bodies[w->dof_body[d]].dq_init;
    }
    memset(w->tau, 0, sizeof w->tau);
    memset(w->fext, 0, sizeof w->fext);
    w->ncontacts = 0;
    w->time = 0;
    w->frame = 0;
}

/* ------------------------------------------------------------------ pydart2-style queries */
void orc_update_kinematics(orc_world_t* w) { forward_kinematics(w); }
void orc_body_transform(orc_world_t* w, int body, double out[12]) {
    forward_kinematics(w);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) out[4 * i + j] = w->Tw[body].R[3 * i + j];
        out[4 * i + 3] = w->Tw[body].p[i];
    }
}
void orc_body_com(orc_world_t* w, int body, double out[3]) { /* bodynode.com() world */
    forward_kinematics(w);
    xf_point(&w->Tw[body], w->m.bodies[body].com, out);
}
void orc_body_com_spatial_velocity(orc_world_t* w, int body, double out[6]) {
    /* [omega; v_com] in BODY coordinates: pydart2's no-argument com_spatial_velocity() is DART's
     * BodyNode::getCOMSpatialVelocity() = getSpatialVelocity(getLocalCOM()), "expressed in coordinates of this
     * Frame and relative to the World Frame" [RECALLED: DART 6 BodyNode.hpp / Frame.cpp; not in /root/reference] */
    forward_kinematics(w);
    const double* V = w->V[body];
    double t[3];
    v3cross(V, w->m.bodies[body].com, t);
    for (int k = 0; k < 3; k++) { out[k] = V[k]; out[3 + k] = V[3 + k] + t[k]; }
}
void orc_add_ext_force(orc_world_t* w, int body, const double f[3]) {
    /* world-frame force applied at the body origin (offset 0): pure linear part in body frame */
    forward_kinematics(w);
    double fb[3];
    m3tv(w->Tw[body].R, f, fb);
    for (int k = 0; k < 3; k++) w->fext[body][3 + k] += fb[k];
}
int orc_num_contacts(const orc_world_t* w) { return w->ncontacts; }
void orc_get_contact(const orc_world_t* w, int i, int* body, double* point, double* normal, double* depth,
                     double* force) {
    const contact_t* c = &w->contacts[i];
    *body = c->body;
    *depth = c->depth;
    for (int k = 0; k < 3; k++) { point[k] = c->point[k]; normal[k] = c->normal[k]; force[k] = c->force[k]; }
}
double orc_shape_gap(const orc_world_t* w, int si) { return w->shape_gap[si]; }
double orc_shape_tilt(const orc_world_t* w, int si) { return w->shape_tilt[si]; }
int orc_limit_active(const orc_world_t* w, int dof) { return w->limit_active[dof]; }
int orc_lcp_rows(const orc_world_t* w) { return w->nrows; }
void orc_get_lcp(const orc_world_t* w, double* A, double* x, double* b, double* lo, double* hi, int* findex) {
    int n = w->nrows;
    memcpy(A, w->lcpA, n * n * sizeof(double));
    memcpy(x, w->lcpx, n * sizeof(double));
    memcpy(b, w->lcpb, n * sizeof(double));
    memcpy(lo, w->lcplo, n * sizeof(double));
    memcpy(hi, w->lcphi, n * sizeof(double));
    memcpy(findex, w->lcpfindex, n * sizeof(int));
}
int orc_lcp_failed(const orc_world_t* w) { return w->lcp_fail; }

/* ------------------------------------------------------------------ independent cross-checks */
/* Mass matrix by sum_b J_b^T I_b J_b (body Jacobians in body frames): no shared code with ABA. */
void orc_mass_matrix(orc_world_t* w, double* M) {
    int nd = w->nd;
    forward_kinematics(w);
    memset(M, 0, nd * nd * sizeof(double));
    for (int i = 0; i < w->nb; i++) {
        double Jb[NB][6];
        int used[NB];
        memset(used, 0, sizeof used);
        xf_t Tinv;
        xf_inv(&w->Tw[i], &Tinv);
        for (int j = i; j >= 0; j = w->m.bodies[j].parent) {
            int d = w->m.bodies[j].dof;
            if (d < 0) continue;
            /* S_j from frame j to frame i: pose of i in j = Tw[j]^-1 Tw[i] */
            xf_t Tj_inv, Tij;
            xf_inv(&w->Tw[j], &Tj_inv);
            xf_mul(&Tj_inv, &w->Tw[i], &Tij);
            sx_motion(&Tij, w->S[j], Jb[d]);
            used[d] = 1;
        }
        for (int a = 0; a < nd; a++) {
            if (!used[a]) continue;
            double IJ[6];
            m6v(w->I6[i], Jb[a], IJ);
            for (int c = 0; c < nd; c++) {
                if (!used[c]) continue;
                double s = 0;
                for (int k = 0; k < 6; k++) s += Jb[c][k] * IJ[k];
                M[c * nd + a] += s;
            }
        }
    }
}
/* ABA forward dynamics with the current tau / fext; returns ddq (no state change) */
void orc_forward_dynamics(orc_world_t* w, double* ddq) {
    forward_kinematics(w);
    forward_dynamics(w);
    memcpy(ddq, w->ddq, w->nd * sizeof(double));
}
/* kinetic + potential energy (springs excluded) */
double orc_energy(orc_world_t* w) {
    forward_kinematics(w);
    double E = 0;
    for (int i = 0; i < w->nb; i++) {
        double IV[6], c[3];
        m6v(w->I6[i], w->V[i], IV);
        double ke = 0;
        for (int k = 0; k < 6; k++) ke += 0.5 * w->V[i][k] * IV[k];
        xf_point(&w->Tw[i], w->m.bodies[i].com, c);
        E += ke - w->m.bodies[i].mass * v3dot(w->m.gravity, c);
    }
    return E;
}

/* ------------------------------------------------------------------ counter-based RNG */
/* Philox4x32-10; identical constants on the GPU so reset noise is bit-identical. */
static void philox4x32(uint32_t ctr[4], const uint32_t key_in[2]) {
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0], p1 = (uint64_t)0xCD9E8D57u * ctr[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ k1, n3 = (uint32_t)p0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
/* i-th uniform in [-1, 1) for (seed, world, episode); fp32 arithmetic, no FMA contraction */
float orc_reset_uniform(uint64_t seed, int64_t world, uint32_t episode, int i) {
    uint32_t ctr[4] = {(uint32_t)world, (uint32_t)((uint64_t)world >> 32), episode, (uint32_t)(i >> 2)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    philox4x32(ctr, key);
    uint32_t bits = ctr[i & 3];
    volatile float u = (float)(bits >> 8) * (1.0f / 16777216.0f); /* [0,1) exact */
    volatile float r = u * 2.0f;
    r = r - 1.0f;
    return r;
}

/* ------------------------------------------------------------------ task layer */
typedef struct orc_env {
    orc_world_t* w;
    dartb_task_t t;
    uint64_t seed;
    int64_t world_id;
    uint32_t episode;
} orc_env_t;

orc_env_t* orc_env_create(const dartb_model_t* m, const dartb_task_t* t, uint64_t seed, int64_t world_id) {
    orc_env_t* e = (orc_env_t*)calloc(1, sizeof(orc_env_t));
    e->w = orc_create(m);
    e->t = *t;
    e->seed = seed;
    e->world_id = world_id;
    return e;
}
void orc_env_destroy(orc_env_t* e) { orc_destroy(e->w); free(e); }
orc_world_t* orc_env_world(orc_env_t* e) { return e->w; }

void orc_env_obs(orc_env_t* e, double* obs) {
    orc_world_t* w = e->w;
    const dartb_task_t* t = &e->t;
    int nd = w->nd, k = 0;
    if (t->obs_mode == DARTB_OBS_HEIGHT_Q2_DQ) {
        double c[3];
        orc_body_com(w, t->height_body, c);
        obs[k++] = c[1];
        for (int d = 2; d < nd; d++) obs[k++] = w->q[d];
    } else {
        for (int d = 1; d < nd; d++) obs[k++] = w->q[d];
    }
    for (int d = 0; d < nd; d++) {
        double v = w->dq[d];
        if (t->dq_clip > 0) { if (v > t->dq_clip) v = t->dq_clip; if (v < -t->dq_clip) v = -t->dq_clip; }
        obs[k++] = v;
    }
}

/* reset_model(): world.reset(); q0 + U(+-noise), dq0 + U(+-noise); returns obs */
void orc_env_reset(orc_env_t* e, double* obs) {
    orc_world_t* w = e->w;
    orc_reset(w);
    int nd = w->nd;
    float noise = (float)e->t.reset_noise;
    for (int d = 0; d < nd; d++) {
        volatile float a = orc_reset_uniform(e->seed, e->world_id, e->episode, d) * noise;
        volatile float b = orc_reset_uniform(e->seed, e->world_id, e->episode, nd + d) * noise;
        volatile float q0 = (float)w->m.bodies[w->dof_body[d]].q_init, v0 = (float)w->m.bodies[w->dof_body[d]].dq_init;
        volatile float qq = q0 + a, vv = v0 + b;
        w->q[d] = (double)qq;
        w->dq[d] = (double)vv;
    }
    e->episode++;
    if (obs) orc_env_obs(e, obs);
}

/* snake_7link.py:35-47 fluid force, applied before each sub-step */
static void fluid_forces(orc_env_t* e) {
    orc_world_t* w = e->w;
    forward_kinematics(w);
    for (int i = 0; i < w->nb; i++) {
        double sv[6], nrm[3] = {w->Tw[i].R[2], w->Tw[i].R[5], w->Tw[i].R[8]}; /* R * ez */
        /* bn.com_spatial_velocity(): BODY-frame components (see orc_body_com_spatial_velocity); the reference
         * combines them with the WORLD-frame norm_dir component by component (snake_7link.py:37-45) */
        const double* V = w->V[i];
        double tt[3];
        v3cross(V, w->m.bodies[i].com, tt);
        for (int k = 0; k < 3; k++) { sv[k] = V[k]; sv[3 + k] = V[3 + k] + tt[k]; }
        double cr[3], vp[3], vn[3], f[3] = {0, 0, 0};
        v3cross(sv, nrm, cr);
        for (int k = 0; k < 3; k++) { vp[k] = sv[3 + k] + cr[k] * e->t.fluid_offset; vn[k] = sv[3 + k] - cr[k] * e->t.fluid_offset; }
        double dp = v3dot(vp, nrm), dn = v3dot(vn, nrm);
        if (dp > 0.0) for (int k = 0; k < 3; k++) f[k] = -e->t.fluid_coef * dp * nrm[k];
        if (dn < 0.0) for (int k = 0; k < 3; k++) f[k] = -e->t.fluid_coef * dn * nrm[k];
        double fb[3];
        m3tv(w->Tw[i].R, f, fb);
        for (int k = 0; k < 3; k++) w->fext[i][3 + k] += fb[k];
    }
}

void orc_env_step(orc_env_t* e, const double* action, double* obs, double* reward, int* done) {
    orc_world_t* w = e->w;
    const dartb_task_t* t = &e->t;
    int nd = w->nd;
    double tau[NB];
    memset(tau, 0, sizeof tau);
    double a2 = 0;
    for (int i = 0; i < t->n_act; i++) {
        double a = action[i];
        a2 += a * a; /* RAW action in the control cost (hopper.py:55) */
        if (a > t->act_hi[i]) a = t->act_hi[i];
        if (a < t->act_lo[i]) a = t->act_lo[i];
        tau[t->act_dof[i]] = a * t->act_scale[i];
    }
    double posbefore = w->q[0];
    for (int f = 0; f < t->frame_skip; f++) {
        if (t->fluid_force) fluid_forces(e);
        orc_set_forces(w, tau);
        orc_step(w);
    }
    double posafter = w->q[0], ang = w->q[2];
    double dt_env = w->m.dt * t->frame_skip;
    double height = 0;
    if (t->height_body >= 0) { double c[3]; orc_body_com(w, t->height_body, c); height = c[1]; }
    double r = (posafter - posbefore) / dt_env * t->vel_weight;
    r += t->alive_bonus;
    r -= t->ctrl_cost * a2;
    if (t->limit_pen_dof >= 0) {
        int j = t->limit_pen_dof;
        const dartb_body_t* bd = &w->m.bodies[w->dof_body[j]];
        double pen = 0;
        if ((bd->q_lo - w->q[j]) > -t->limit_pen_margin) pen += 1.5;
        if ((bd->q_hi - w->q[j]) < t->limit_pen_margin) pen += 1.5;
        r -= t->limit_pen_weight * pen;
    }
    r -= t->dev_cost * fabs(ang);
    int finite = 1, bounded = 1;
    for (int d = 0; d < nd; d++) {
        if (!isfinite(w->q[d]) || !isfinite(w->dq[d])) finite = 0;
        if (d >= 2 && !(fabs(w->q[d]) < t->state_bound)) bounded = 0;
        if (!(fabs(w->dq[d]) < t->state_bound)) bounded = 0;
    }
    int ok = finite && bounded;
    if (t->zero_reward_on_blowup && !ok) r = 0;
    if (t->height_body >= 0) ok = ok && (height > t->height_lo) && (height < t->height_hi);
    ok = ok && (fabs(ang) < t->ang_max);
    *reward = r;
    *done = !ok;
    orc_env_obs(e, obs);
}

/* ------------------------------------------------------------------ threaded CPU baseline driver */
typedef struct {
    const dartb_model_t* m; const dartb_task_t* t;
    int first, count, steps; uint64_t seed; long done_steps; double checksum;
} job_t;

static uint32_t lcg(uint32_t* s) { *s = *s * 1664525u + 1013904223u; return *s; }

static void* run_job(void* arg) {
    job_t* j = (job_t*)arg;
    double obs[64], rew, act[DARTB_MAX_ACT];
    int done;
    for (int wi = 0; wi < j->count; wi++) {
        orc_env_t* e = orc_env_create(j->m, j->t, j->seed, j->first + wi);
        uint32_t s = (uint32_t)(j->seed * 2654435761u + (uint32_t)(j->first + wi));
        orc_env_reset(e, obs);
        for (int st = 0; st < j->steps; st++) {
            for (int a = 0; a < j->t->n_act; a++) act[a] = (double)(lcg(&s) >> 8) / 8388608.0 - 1.0;
            orc_env_step(e, act, obs, &rew, &done);
            j->checksum += rew;
            if (done) orc_env_reset(e, obs); /* auto-reset (sync_vector_env.py:76-79) */
            j->done_steps++;
        }
        orc_env_destroy(e);
    }
    return NULL;
}

/* Steps n_worlds envs for n_steps each with random actions on n_threads threads.
 * Returns env-steps/second (wall clock); *checksum gets the reward sum (keeps work live). */
double orc_bench(const dartb_model_t* m, const dartb_task_t* t, int n_worlds, int n_steps, int n_threads,
                 uint64_t seed, double* checksum) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    job_t jobs[256];
    int per = (n_worlds + n_threads - 1) / n_threads, first = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int used = 0;
    for (int i = 0; i < n_threads && first < n_worlds; i++) {
        int cnt = per; if (first + cnt > n_worlds) cnt = n_worlds - first;
        jobs[i] = (job_t){m, t, first, cnt, n_steps, seed, 0, 0.0};
        pthread_create(&th[i], NULL, run_job, &jobs[i]);
        first += cnt; used++;
    }
    long total = 0; double cs = 0;
    for (int i = 0; i < used; i++) { pthread_join(th[i], NULL); total += jobs[i].done_steps; cs += jobs[i].checksum; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double sec = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    if (checksum) *checksum = cs;
    return (double)total / sec;
}
