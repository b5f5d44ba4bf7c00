/* ORACLE (test infrastructure, not product code) - float64 CPU restatement of the articulated
 * rigid-body control step of the B200 environment kernel (vid2player3d_b200/csrc).
 *
 * PARITY UNPINNED: in the reference this arithmetic lives inside NVIDIA Isaac Gym Preview 4
 * (closed-source PhysX 5 binary, un-vendored, un-installable here).  The reference
 * only fixes the INPUTS of the solver - the call sites `gym.simulate` (embodied_pose/env/tasks/
 * base_task.py:450-454), the assets (data/assets/mjcf/smpl_mesh_humanoid_amass_v1.xml), the sim
 * block (cfg/amass_im.yaml:37-52, utils/config.py:198-231), the actor properties
 * (humanoid_smpl_im.py:273-276,356-389) and the actuation semantics (:125-157,391-396).  There is
 * no golden trajectory in the reference.  This file therefore restates OUR model:
 *
 *   - 23 spherical joints (exp-map coordinates, child-frame relative angular velocity) + free root,
 *   - Featherstone articulated-body algorithm written with classical accelerations in world-aligned
 *     axes with each body's reference point at its own joint origin,
 *   - implicit ("stable") PD drive + armature + joint-limit springs on the joint-space diagonal,
 *   - compliant ground contact of every convex-hull vertex, implicit in the velocity (the
 *     spring/damper/viscous-friction terms enter the body's articulated inertia as rank-1 updates),
 *   - semi-implicit Euler on SO(3), h = sim_dt / substeps.
 *
 * It is deliberately written differently from the kernel (serial per env, full 6x6 matrices,
 * generic matrix products) so that agreement is evidence, not tautology.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may call
 * into this library.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/b200env.h"

typedef double real;

static void q_mul(const real* a, const real* b, real* o) {
  real x1 = a[0], y1 = a[1], z1 = a[2], w1 = a[3], x2 = b[0], y2 = b[1], z2 = b[2], w2 = b[3];
  o[0] = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2;
  o[1] = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2;
  o[2] = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2;
  o[3] = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2;
}
static void q_norm(real* q) {
  real n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) q[k] /= n;
}
static void q_to_mat(const real* q, real R[3][3]) {
  real x = q[0], y = q[1], z = q[2], w = q[3];
  R[0][0] = 1 - 2 * (y * y + z * z); R[0][1] = 2 * (x * y - z * w); R[0][2] = 2 * (x * z + y * w);
  R[1][0] = 2 * (x * y + z * w); R[1][1] = 1 - 2 * (x * x + z * z); R[1][2] = 2 * (y * z - x * w);
  R[2][0] = 2 * (x * z - y * w); R[2][1] = 2 * (y * z + x * w); R[2][2] = 1 - 2 * (x * x + y * y);
}
/* exp: rotation vector -> unit quaternion */
static void q_exp(const real* v, real* q) {
  real a2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  real a = sqrt(a2), s;
  if (a < 1e-6) s = 0.5 - a2 / 48.0; else s = sin(0.5 * a) / a;
  q[0] = s * v[0]; q[1] = s * v[1]; q[2] = s * v[2]; q[3] = cos(0.5 * a);
}
/* log: unit quaternion -> rotation vector with angle in [0, pi] */
static void q_log(const real* qin, real* v) {
  real q[4] = {qin[0], qin[1], qin[2], qin[3]};
  if (q[3] < 0) for (int k = 0; k < 4; k++) q[k] = -q[k];
  real s2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
  real s = sqrt(s2), f;
  if (s < 1e-6) f = 2.0 + s2 / 3.0; else f = 2.0 * atan2(s, q[3]) / s;
  v[0] = f * q[0]; v[1] = f * q[1]; v[2] = f * q[2];
}
static void cross(const real* a, const real* b, real* o) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static void matvec3(real M[3][3], const real* v, real* o) {
  real t[3];
  for (int i = 0; i < 3; i++) t[i] = M[i][0] * v[0] + M[i][1] * v[1] + M[i][2] * v[2];
  o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}
static void matTvec3(real M[3][3], const real* v, real* o) {
  real t[3];
  for (int i = 0; i < 3; i++) t[i] = M[0][i] * v[0] + M[1][i] * v[1] + M[2][i] * v[2];
  o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}
static void skew(const real* r, real X[3][3]) {
  X[0][0] = 0; X[0][1] = -r[2]; X[0][2] = r[1];
  X[1][0] = r[2]; X[1][1] = 0; X[1][2] = -r[0];
  X[2][0] = -r[1]; X[2][1] = r[0]; X[2][2] = 0;
}
static int inv3(real M[3][3], real O[3][3]) {
  real c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1];
  real c01 = M[1][2] * M[2][0] - M[1][0] * M[2][2];
  real c02 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
  real det = M[0][0] * c00 + M[0][1] * c01 + M[0][2] * c02;
  if (!(fabs(det) > 0)) return -1;
  real id = 1.0 / det;
  O[0][0] = c00 * id; O[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) * id; O[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * id;
  O[1][0] = c01 * id; O[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) * id; O[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * id;
  O[2][0] = c02 * id; O[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) * id; O[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * id;
  return 0;
}
/* solve the 6x6 SPD system M x = b by Cholesky */
static int solve6(real M[6][6], const real* b, real* x) {
  real L[6][6];
  memset(L, 0, sizeof(L));
  for (int i = 0; i < 6; i++)
    for (int j = 0; j <= i; j++) {
      real s = M[i][j];
      for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
      if (i == j) { if (!(s > 0)) return -1; L[i][i] = sqrt(s); }
      else L[i][j] = s / L[j][j];
    }
  real y[6];
  for (int i = 0; i < 6; i++) { real s = b[i]; for (int k = 0; k < i; k++) s -= L[i][k] * y[k]; y[i] = s / L[i][i]; }
  for (int i = 5; i >= 0; i--) { real s = y[i]; for (int k = i + 1; k < 6; k++) s -= L[k][i] * x[k]; x[i] = s / L[i][i]; }
  return 0;
}

typedef struct {
  real Q[4], R[3][3], p[3], w[3], v[3];      /* world pose / velocity of the body origin */
  real r[3];                                 /* p - p_parent */
  real zeta[6];                              /* velocity-product acceleration */
  real IA[6][6], bA[6];
  real Dinv[3][3], u[3];
  real A[6];                                 /* (alpha, a) */
  real qj[4], wt[3];                         /* joint quaternion (child in parent), joint velocity (child frame) */
} body_t;

static void rank1(real IA[6][6], const real* J, real k) {
  for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) IA[a][b] += k * J[a] * J[b];
}

/* one substep of length h for one env; ext = wrench (force, torque) on body 0 or NULL */
struct ball_s;
static int substep(const b200_model_t* m, const float* verts, const b200_cfg_t* cfg, real h, body_t* B,
                   const real* pd_tar, const real* ext, real* contact_f, const real* racket_react /* F[3], X[3] or NULL */) {
  int nb = m->nb;
  real g[3] = {0, 0, cfg->gravity_z};
  /* 1. kinematics */
  q_to_mat(B[0].Q, B[0].R);
  for (int i = 1; i < nb; i++) {
    body_t* b = &B[i]; body_t* P = &B[m->parent[i]];
    real off[3] = {m->offset[i][0], m->offset[i][1], m->offset[i][2]};
    matvec3(P->R, off, b->r);
    for (int k = 0; k < 3; k++) b->p[k] = P->p[k] + b->r[k];
    real wxr[3]; cross(P->w, b->r, wxr);
    for (int k = 0; k < 3; k++) b->v[k] = P->v[k] + wxr[k];
    if (m->fixed[i]) {
      memcpy(b->Q, P->Q, sizeof(b->Q)); memcpy(b->R, P->R, sizeof(b->R)); memcpy(b->w, P->w, sizeof(b->w));
      continue;
    }
    q_mul(P->Q, b->qj, b->Q); q_norm(b->Q); q_to_mat(b->Q, b->R);
    real wj[3]; matvec3(b->R, b->wt, wj);
    for (int k = 0; k < 3; k++) b->w[k] = P->w[k] + wj[k];
    cross(P->w, wj, &b->zeta[0]);
    cross(P->w, wxr, &b->zeta[3]);
  }
  /* 2. rigid-body inertia, bias, external + contact + joint forces */
  for (int i = 0; i < nb; i++) {
    if (m->fixed[i]) { if (contact_f) contact_f[3 * i] = contact_f[3 * i + 1] = contact_f[3 * i + 2] = 0; continue; }
    body_t* b = &B[i];
    real ms = m->mass[i];
    real cl[3] = {m->com[i][0], m->com[i][1], m->com[i][2]}, c[3];
    matvec3(b->R, cl, c);
    real Ib[3][3] = {{m->inertia[i][0], m->inertia[i][3], m->inertia[i][4]},
                     {m->inertia[i][3], m->inertia[i][1], m->inertia[i][5]},
                     {m->inertia[i][4], m->inertia[i][5], m->inertia[i][2]}};
    real Io[3][3];
    for (int a = 0; a < 3; a++) for (int d = 0; d < 3; d++) {
      real s = 0;
      for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) s += b->R[a][k] * Ib[k][l] * b->R[d][l];
      Io[a][d] = s + ms * ((a == d ? (c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) : 0.0) - c[a] * c[d]);
    }
    real Cx[3][3]; skew(c, Cx);
    memset(b->IA, 0, sizeof(b->IA));
    for (int a = 0; a < 3; a++) for (int d = 0; d < 3; d++) {
      b->IA[a][d] = Io[a][d];
      b->IA[a][3 + d] = ms * Cx[a][d];
      b->IA[3 + a][d] = ms * Cx[d][a];
      b->IA[3 + a][3 + d] = (a == d) ? ms : 0.0;
    }
    real Iw[3], wxIw[3], wxc[3], wxwxc[3], cxg[3];
    matvec3(Io, b->w, Iw); cross(b->w, Iw, wxIw);
    cross(b->w, c, wxc); cross(b->w, wxc, wxwxc);
    cross(c, g, cxg);
    for (int k = 0; k < 3; k++) { b->bA[k] = wxIw[k] - ms * cxg[k]; b->bA[3 + k] = ms * wxwxc[k] - ms * g[k]; }
    if (racket_react && cfg->racket_body >= 0 && i == m->parent[cfg->racket_body]) { /* last racket impact's reaction on the wrist */
      real dx[3] = {racket_react[3] - b->p[0], racket_react[4] - b->p[1], racket_react[5] - b->p[2]}, t[3];
      cross(dx, racket_react, t);
      for (int k = 0; k < 3; k++) { b->bA[k] -= t[k]; b->bA[3 + k] -= racket_react[k]; }
    }
    if (i == 0 && ext) { /* force acts at the COM of body 0 (apply_rigid_body_force_tensors, ENV_SPACE) */
      real cxF[3]; cross(c, ext, cxF);
      for (int k = 0; k < 3; k++) { b->bA[k] -= ext[3 + k] + cxF[k]; b->bA[3 + k] -= ext[k]; }
    }
    /* ground contact of hull vertices (plane z = 0) */
    real fsum[3] = {0, 0, 0};
    if (m->nverts[i] > 0 && b->p[2] - m->radius[i] < 0) {
      for (int k = 0; k < m->nverts[i]; k++) {
        const float* vl = verts + ((size_t)i * m->vmax + k) * 3;
        real vb[3] = {vl[0], vl[1], vl[2]}, rho[3];
        matvec3(b->R, vb, rho);
        real pen = -(b->p[2] + rho[2]);
        if (!(pen > 0)) continue;
        real wxrho[3], uu[3]; cross(b->w, rho, wxrho);
        for (int a = 0; a < 3; a++) uu[a] = b->v[a] + wxrho[a];
        real fn0 = cfg->contact_kn * pen - cfg->contact_cn * uu[2];
        if (!(fn0 > 0)) continue;
        real n[3] = {0, 0, 1}, tx[3] = {1, 0, 0}, ty[3] = {0, 1, 0};
        real Jn[6], Jx[6], Jy[6];
        cross(rho, n, Jn); cross(rho, tx, Jx); cross(rho, ty, Jy);
        for (int a = 0; a < 3; a++) { Jn[3 + a] = n[a]; Jx[3 + a] = tx[a]; Jy[3 + a] = ty[a]; }
        real kn_imp = h * cfg->contact_cn + h * h * cfg->contact_kn;
        rank1(b->IA, Jn, kn_imp);
        real ut = sqrt(uu[0] * uu[0] + uu[1] * uu[1]);
        real ct = cfg->friction_mu * fn0 / fmax(ut, (real)cfg->friction_vs);
        rank1(b->IA, Jx, h * ct); rank1(b->IA, Jy, h * ct);
        for (int a = 0; a < 6; a++) b->bA[a] -= Jn[a] * fn0 - ct * (Jx[a] * uu[0] + Jy[a] * uu[1]);
        fsum[0] += -ct * uu[0]; fsum[1] += -ct * uu[1]; fsum[2] += fn0;
      }
    }
    if (contact_f) { contact_f[3 * i] = fsum[0]; contact_f[3 * i + 1] = fsum[1]; contact_f[3 * i + 2] = fsum[2]; }
    /* joint drive: implicit PD + armature + limit springs (child frame, exp-map chart) */
    if (i > 0) {
      int d0 = m->dof_of_body[i];
      real q[3]; q_log(b->qj, q);
      real tau[3], e[3];
      for (int k = 0; k < 3; k++) {
        real kp = m->kp[d0 + k], kd = m->kd[d0 + k];
        e[k] = m->armature[d0 + k] + h * kd + h * h * kp;
        tau[k] = kp * (pd_tar[d0 + k] - q[k] - h * b->wt[k]) - kd * b->wt[k];
        real lo = m->lim_lo[d0 + k], hi = m->lim_hi[d0 + k];
        if (q[k] < lo) { tau[k] += cfg->limit_k * (lo - q[k] - h * b->wt[k]) - cfg->limit_c * b->wt[k]; e[k] += h * cfg->limit_c + h * h * cfg->limit_k; }
        else if (q[k] > hi) { tau[k] += cfg->limit_k * (hi - q[k] - h * b->wt[k]) - cfg->limit_c * b->wt[k]; e[k] += h * cfg->limit_c + h * h * cfg->limit_k; }
      }
      matvec3(b->R, tau, b->u); /* u holds tau_w for now */
      for (int a = 0; a < 3; a++) for (int d = 0; d < 3; d++) {
        real s = 0;
        for (int k = 0; k < 3; k++) s += b->R[a][k] * e[k] * b->R[d][k];
        b->Dinv[a][d] = s; /* Dinv holds E_w for now */
      }
    }
  }
  /* 3. backward pass */
  for (int i = nb - 1; i >= 1; i--) {
    if (m->fixed[i]) continue;
    body_t* b = &B[i]; body_t* P = &B[m->parent[i]];
    real D[3][3];
    for (int a = 0; a < 3; a++) for (int d = 0; d < 3; d++) D[a][d] = b->IA[a][d] + b->Dinv[a][d];
    if (inv3(D, b->Dinv)) return -1;
    for (int a = 0; a < 3; a++) b->u[a] -= b->bA[a];
    real UD[6][3];
    for (int a = 0; a < 6; a++) for (int d = 0; d < 3; d++) {
      real s = 0; for (int k = 0; k < 3; k++) s += b->IA[a][k] * b->Dinv[k][d]; UD[a][d] = s;
    }
    real Ia[6][6], ba[6];
    for (int a = 0; a < 6; a++) for (int d = 0; d < 6; d++) {
      real s = 0; for (int k = 0; k < 3; k++) s += UD[a][k] * b->IA[d][k];
      Ia[a][d] = b->IA[a][d] - s;
    }
    for (int a = 0; a < 6; a++) {
      real s = b->bA[a];
      for (int k = 0; k < 6; k++) s += Ia[a][k] * b->zeta[k];
      for (int k = 0; k < 3; k++) s += UD[a][k] * b->u[k];
      ba[a] = s;
    }
    /* Phi = [[1,0],[-[r]x,1]]  (A_child = Phi A_parent + ...);  parent += Phi^T Ia Phi, Phi^T ba */
    real Phi[6][6]; memset(Phi, 0, sizeof(Phi));
    real X[3][3]; skew(b->r, X);
    for (int a = 0; a < 6; a++) Phi[a][a] = 1;
    for (int a = 0; a < 3; a++) for (int d = 0; d < 3; d++) Phi[3 + a][d] = -X[a][d];
    real T[6][6];
    for (int a = 0; a < 6; a++) for (int d = 0; d < 6; d++) { real s = 0; for (int k = 0; k < 6; k++) s += Ia[a][k] * Phi[k][d]; T[a][d] = s; }
    for (int a = 0; a < 6; a++) for (int d = 0; d < 6; d++) { real s = 0; for (int k = 0; k < 6; k++) s += Phi[k][a] * T[k][d]; P->IA[a][d] += s; }
    for (int a = 0; a < 6; a++) { real s = 0; for (int k = 0; k < 6; k++) s += Phi[k][a] * ba[k]; P->bA[a] += s; }
  }
  /* 4. root */
  real nb0[6]; for (int a = 0; a < 6; a++) nb0[a] = -B[0].bA[a];
  if (solve6(B[0].IA, nb0, B[0].A)) return -2;
  /* 5. forward pass */
  for (int i = 1; i < nb; i++) {
    if (m->fixed[i]) continue;
    body_t* b = &B[i]; body_t* P = &B[m->parent[i]];
    real Ap[6], axr[3]; cross(&P->A[0], b->r, axr);
    for (int k = 0; k < 3; k++) { Ap[k] = P->A[k] + b->zeta[k]; Ap[3 + k] = P->A[3 + k] + axr[k] + b->zeta[3 + k]; }
    real t[3];
    for (int a = 0; a < 3; a++) { real s = b->u[a]; for (int k = 0; k < 6; k++) s -= b->IA[k][a] * Ap[k]; t[a] = s; }
    real gam[3]; matvec3(b->Dinv, t, gam);
    for (int k = 0; k < 3; k++) { b->A[k] = Ap[k] + gam[k]; b->A[3 + k] = Ap[3 + k]; }
    real wd[3]; matTvec3(b->R, gam, wd);
    /* 6a. integrate joint velocity, damping, clamp */
    real damp = 1.0 - h * cfg->ang_damping;
    for (int k = 0; k < 3; k++) b->wt[k] = (b->wt[k] + h * wd[k]) * damp;
    real nn = sqrt(b->wt[0] * b->wt[0] + b->wt[1] * b->wt[1] + b->wt[2] * b->wt[2]);
    if (nn > cfg->max_ang_vel) for (int k = 0; k < 3; k++) b->wt[k] *= cfg->max_ang_vel / nn;
  }
  /* 6b. root velocity + all positions */
  {
    body_t* b = &B[0];
    real damp = 1.0 - h * cfg->ang_damping;
    for (int k = 0; k < 3; k++) { b->w[k] = (b->w[k] + h * b->A[k]) * damp; b->v[k] += h * b->A[3 + k]; }
    real nn = sqrt(b->w[0] * b->w[0] + b->w[1] * b->w[1] + b->w[2] * b->w[2]);
    if (nn > cfg->max_ang_vel) for (int k = 0; k < 3; k++) b->w[k] *= cfg->max_ang_vel / nn;
    for (int k = 0; k < 3; k++) b->p[k] += h * b->v[k];
    real hv[3] = {h * b->w[0], h * b->w[1], h * b->w[2]}, dq[4], qn[4];
    q_exp(hv, dq); q_mul(dq, b->Q, qn); q_norm(qn); memcpy(b->Q, qn, sizeof(qn));
  }
  for (int i = 1; i < nb; i++) {
    if (m->fixed[i]) continue;
    body_t* b = &B[i];
    real hv[3] = {h * b->wt[0], h * b->wt[1], h * b->wt[2]}, dq[4], qn[4];
    q_exp(hv, dq); q_mul(b->qj, dq, qn); q_norm(qn); memcpy(b->qj, qn, sizeof(qn));
  }
  return 0;
}

/* forward kinematics only: fills Q,R,p,w,v of every body from the current state */
static void fk(const b200_model_t* m, body_t* B) {
  q_to_mat(B[0].Q, B[0].R);
  for (int i = 1; i < m->nb; i++) {
    body_t* b = &B[i]; body_t* P = &B[m->parent[i]];
    real off[3] = {m->offset[i][0], m->offset[i][1], m->offset[i][2]};
    matvec3(P->R, off, b->r);
    real wxr[3]; cross(P->w, b->r, wxr);
    for (int k = 0; k < 3; k++) { b->p[k] = P->p[k] + b->r[k]; b->v[k] = P->v[k] + wxr[k]; }
    if (m->fixed[i]) { memcpy(b->Q, P->Q, sizeof(b->Q)); memcpy(b->R, P->R, sizeof(b->R)); memcpy(b->w, P->w, sizeof(b->w)); continue; }
    q_mul(P->Q, b->qj, b->Q); q_norm(b->Q); q_to_mat(b->Q, b->R);
    real wj[3]; matvec3(b->R, b->wt, wj);
    for (int k = 0; k < 3; k++) b->w[k] = P->w[k] + wj[k];
  }
}


/* ------------------------------------------------------------------------------------------------
 * Tennis ball (vid2player): our model, float64 restatement of csrc/b200env.cu::ball_substep.
 * Reference inputs honoured: tennis_ball.urdf (r, m, I), aerodynamic force of apply_external_force_to_ball
 * (vid2player/env/tasks/humanoid_smpl_im_mvae.py:711-739, refreshed once per sim step :752-756), material
 * restitution / friction (:414-416,436-438; PhysX "average" combine), racket head cylinder (federer.xml:190).
 * Parity vs PhysX is unpinned like the rest of the physics. */
typedef struct {
  real p[3], v[3], w[3], fa[3], rF[3], rX[3];
  int hits, has_bounce, bounce_now;
  real bpos[3];
} ball_t;

static void q_rot(const real* q, const real* v, real* o) {
  real R[3][3]; q_to_mat(q, R); matvec3(R, v, o);
}
static void q_rot_inv(const real* q, const real* v, real* o) {
  real R[3][3]; q_to_mat(q, R); matTvec3(R, v, o);
}
static void ball_aero_ref(const real* vel, const real* angvel, real spin_scale, real* f) {
  const real KF = 0.0019462794807519486, CD = 0.55, PI2 = 6.283185307179586;
  real vs = sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
  if (vs == 0) vs += 1;
  real vn[3] = {vel[0] / vs, vel[1] / vs, vel[2] / vs}, down[3] = {0, 0, -1}, vt[3], c[3];
  cross(vn, down, vt);
  real vspin = sqrt(angvel[0] * angvel[0] + angvel[1] * angvel[1] + angvel[2] * angvel[2]) / PI2;
  real cl = 1.0 / (2.0 + fabs(vs / (vspin * spin_scale + 1e-6)));
  if (vspin > 0) cl = -cl;
  cross(vt, vn, c);
  for (int k = 0; k < 3; k++) f[k] = -KF * CD * vs * vel[k] - KF * cl * vs * vs * c[k];
}
static void ball_impulse_ref(const b200_cfg_t* cfg, ball_t* B, const real* n, const real* vo, real e, real mu, real* J) {
  real m = cfg->ball_mass, I = cfg->ball_inertia, R = cfg->ball_radius;
  real rn[3] = {-R * n[0], -R * n[1], -R * n[2]}, wxr[3], u[3];
  cross(B->w, rn, wxr);
  for (int k = 0; k < 3; k++) u[k] = B->v[k] + wxr[k] - vo[k];
  real un = u[0] * n[0] + u[1] * n[1] + u[2] * n[2];
  J[0] = J[1] = J[2] = 0;
  if (!(un < 0)) return;
  real jn = ((-un > cfg->bounce_threshold_velocity) ? (1.0 + e) : 1.0) * (-un) * m;
  real ut[3] = {u[0] - un * n[0], u[1] - un * n[1], u[2] - un * n[2]};
  real utn = sqrt(ut[0] * ut[0] + ut[1] * ut[1] + ut[2] * ut[2]), jt = 0;
  if (utn > 1e-9) {
    real stick = m * utn / (1.0 + m * R * R / I);
    jt = mu * jn < stick ? mu * jn : stick;
    for (int k = 0; k < 3; k++) ut[k] /= utn;
  }
  real rxJ[3];
  for (int k = 0; k < 3; k++) J[k] = jn * n[k] - jt * ut[k];
  cross(rn, J, rxJ);
  for (int k = 0; k < 3; k++) { B->v[k] += J[k] / m; B->w[k] += rxJ[k] / I; }
}
static void ball_substep_ref(const b200_cfg_t* cfg, real h, ball_t* B, const body_t* racket) {
  real m = cfg->ball_mass, R = cfg->ball_radius;
  for (int k = 0; k < 3; k++) B->v[k] += h * B->fa[k] / m;
  B->v[2] += h * cfg->gravity_z;
  for (int k = 0; k < 3; k++) B->rF[k] = 0;
  real thit = -1, nl = 0;
  if (racket) {
    real d[3], wxd[3], vrel[3], d0[3], vr[3];
    for (int k = 0; k < 3; k++) d[k] = B->p[k] - racket->p[k];
    cross(racket->w, d, wxd);
    for (int k = 0; k < 3; k++) vrel[k] = B->v[k] - racket->v[k] - wxd[k];
    real hq[4] = {cfg->racket_head_quat[0], cfg->racket_head_quat[1], cfg->racket_head_quat[2], cfg->racket_head_quat[3]}, hQ[4];
    q_mul(racket->Q, hq, hQ);   /* head frame in the world: the string-bed normal is its +y */
    q_rot_inv(hQ, d, d0); q_rot_inv(hQ, vrel, vr);
    for (int k = 0; k < 3; k++) d0[k] -= cfg->racket_head_center[k];
    real H = cfg->racket_head_halfthick + R, Rad = cfg->racket_head_radius + R;
    if (fabs(d0[1]) < H) {
      if (d0[0] * d0[0] + d0[2] * d0[2] < Rad * Rad) { thit = 0; nl = d0[1] >= 0 ? 1 : -1; }
    } else {
      real t = -1;
      if (d0[1] >= H && vr[1] < 0) t = (d0[1] - H) / (-vr[1]);
      else if (d0[1] <= -H && vr[1] > 0) t = (-H - d0[1]) / vr[1];
      if (t >= 0 && t <= h) {
        real hx = d0[0] + t * vr[0], hz = d0[2] + t * vr[2];
        if (hx * hx + hz * hz < Rad * Rad) { thit = t; nl = d0[1] >= 0 ? 1 : -1; }
      }
    }
    if (thit >= 0) {
      real ny[3] = {0, nl, 0}, n[3], J[3], xc[3], dx[3], wxx[3], vo[3], vb[3];
      q_rot(hQ, ny, n);
      for (int k = 0; k < 3; k++) { xc[k] = B->p[k] + thit * B->v[k] - R * n[k]; dx[k] = xc[k] - racket->p[k]; vb[k] = B->v[k]; }
      cross(racket->w, dx, wxx);
      for (int k = 0; k < 3; k++) vo[k] = racket->v[k] + wxx[k];
      ball_impulse_ref(cfg, B, n, vo, cfg->ball_e_racket, cfg->ball_mu_racket, J);
      if (J[0] != 0 || J[1] != 0 || J[2] != 0) {
        for (int k = 0; k < 3; k++) { B->rF[k] = -J[k] / h; B->rX[k] = xc[k]; B->p[k] += thit * vb[k] + (h - thit) * B->v[k]; }
        B->hits++;
      } else thit = -1;
    }
  }
  if (thit < 0) for (int k = 0; k < 3; k++) B->p[k] += h * B->v[k];
  if (B->p[2] < R && B->v[2] < 0) {
    real n[3] = {0, 0, 1}, vo[3] = {0, 0, 0}, J[3];
    ball_impulse_ref(cfg, B, n, vo, cfg->ball_e_ground, cfg->ball_mu_ground, J);
    B->p[2] = R;
  }
}

/* ball-body contact (b200_cfg_t::ball_body_contact): per-body radius of the spheres that stand for the hull vertices - the same float
 * arithmetic as csrc/dyn_common.cuh::hull_vertex_radius */
static void hull_vertex_radius_ref(const b200_model_t* model, const float* verts, float* vrho) {
  for (int b = 0; b < B200_MAX_BODIES; b++) vrho[b] = 0.0f;
  for (int b = 0; b < model->nb; b++) {
    const int nv = model->nverts[b];
    if (nv < 2) continue;
    const float* v = verts + (size_t)b * model->vmax * 3;
    double sum = 0.0;
    for (int i = 0; i < nv; i++) {
      double best = 1e30;
      for (int j = 0; j < nv; j++) {
        if (j == i) continue;
        const double dx = (double)v[i * 3] - v[j * 3], dy = (double)v[i * 3 + 1] - v[j * 3 + 1], dz = (double)v[i * 3 + 2] - v[j * 3 + 2];
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 > 1e-12 && d2 < best) best = d2;
      }
      sum += sqrt(best);
    }
    double r = 0.5 * sum / nv;
    r = r < 0.005 ? 0.005 : (r > 0.05 ? 0.05 : r);
    vrho[b] = (float)r;
  }
}

/* ---- exact sphere / convex-hull test (round 2).  The hull faces come with the compiled model (model_compiler.hull_faces: triangles with
 * outward winding + their plane equations); phys_ref_set_hull_faces installs them (test infrastructure: a process-wide pointer).  A body
 * without faces falls back to the round-1 stand-in (spheres on the hull vertices).
 * Query for a ball centre c (body frame) and radius R: (1) s = max over the face planes of n.c - d: s > R separates exactly; (2) s <= 0: the
 * centre is inside the hull: depth R - s along the least-penetrated face normal; (3) otherwise the closest point of the hull surface =
 * the closest point over all triangles (Ericson, Real-Time Collision Detection 5.1.5): depth R - distance, normal from it to c. */
static const float* g_planes = 0;     /* [nb][tmax][4] */
static const unsigned char* g_tris = 0; /* [nb][tmax][4] (3 vertex indices + pad) */
static const int32_t* g_ntris = 0;    /* [B200_MAX_BODIES] */
static int g_tmax = 0;
int phys_ref_set_hull_faces(const float* planes, const unsigned char* tris, const int32_t* ntris, int tmax) {
  g_planes = planes; g_tris = tris; g_ntris = ntris; g_tmax = tmax;
  return 0;
}
static void closest_on_triangle(const real* p, const real* a, const real* b, const real* c, real* q) {
  real ab[3], ac[3], ap[3];
  for (int k = 0; k < 3; k++) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
  real d1 = ab[0] * ap[0] + ab[1] * ap[1] + ab[2] * ap[2], d2 = ac[0] * ap[0] + ac[1] * ap[1] + ac[2] * ap[2];
  if (d1 <= 0 && d2 <= 0) { for (int k = 0; k < 3; k++) q[k] = a[k]; return; }
  real bp[3];
  for (int k = 0; k < 3; k++) bp[k] = p[k] - b[k];
  real d3 = ab[0] * bp[0] + ab[1] * bp[1] + ab[2] * bp[2], d4 = ac[0] * bp[0] + ac[1] * bp[1] + ac[2] * bp[2];
  if (d3 >= 0 && d4 <= d3) { for (int k = 0; k < 3; k++) q[k] = b[k]; return; }
  real vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { real v = d1 / (d1 - d3); for (int k = 0; k < 3; k++) q[k] = a[k] + v * ab[k]; return; }
  real cp[3];
  for (int k = 0; k < 3; k++) cp[k] = p[k] - c[k];
  real d5 = ab[0] * cp[0] + ab[1] * cp[1] + ab[2] * cp[2], d6 = ac[0] * cp[0] + ac[1] * cp[1] + ac[2] * cp[2];
  if (d6 >= 0 && d5 <= d6) { for (int k = 0; k < 3; k++) q[k] = c[k]; return; }
  real vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { real w = d2 / (d2 - d6); for (int k = 0; k < 3; k++) q[k] = a[k] + w * ac[k]; return; }
  real va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    real w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    for (int k = 0; k < 3; k++) q[k] = b[k] + w * (c[k] - b[k]);
    return;
  }
  real den = 1 / (va + vb + vc), v = vb * den, w = vc * den;
  for (int k = 0; k < 3; k++) q[k] = a[k] + ab[k] * v + ac[k] * w;
}
/* returns 1 and (pen, outward unit normal nl in the body frame) when the sphere (c, R) touches the hull of body b */
static int hull_sphere_ref(const b200_model_t* m, const float* verts, int b, const real* c, real R, real* pen, real* nl) {
  const int nt = g_ntris[b];
  const float* pl = g_planes + (size_t)b * g_tmax * 4;
  const unsigned char* tr = g_tris + (size_t)b * g_tmax * 4;
  const float* vb = verts + (size_t)b * m->vmax * 3;
  real smax = -1e30;
  int imax = 0;
  for (int t = 0; t < nt; t++) {
    real sd = (real)pl[t * 4] * c[0] + (real)pl[t * 4 + 1] * c[1] + (real)pl[t * 4 + 2] * c[2] - (real)pl[t * 4 + 3];
    if (sd > R) return 0;                          /* separating plane */
    if (sd > smax) { smax = sd; imax = t; }
  }
  if (smax <= 0) {
    *pen = R - smax;
    for (int k = 0; k < 3; k++) nl[k] = pl[imax * 4 + k];
    return 1;
  }
  real best = 1e30, qb[3] = {0, 0, 0};
  for (int t = 0; t < nt; t++) {
    /* the closest point of a convex body to an outside point lies on a face the point sees: back faces skipped */
    if (!((real)pl[t * 4] * c[0] + (real)pl[t * 4 + 1] * c[1] + (real)pl[t * 4 + 2] * c[2] - (real)pl[t * 4 + 3] > 0)) continue;
    real a[3], bb[3], cc[3], q[3];
    for (int k = 0; k < 3; k++) { a[k] = vb[tr[t * 4] * 3 + k]; bb[k] = vb[tr[t * 4 + 1] * 3 + k]; cc[k] = vb[tr[t * 4 + 2] * 3 + k]; }
    closest_on_triangle(c, a, bb, cc, q);
    real e2 = (c[0] - q[0]) * (c[0] - q[0]) + (c[1] - q[1]) * (c[1] - q[1]) + (c[2] - q[2]) * (c[2] - q[2]);
    if (e2 < best) { best = e2; qb[0] = q[0]; qb[1] = q[1]; qb[2] = q[2]; }
  }
  real dist = sqrt(best);
  if (!(dist < R) || dist <= 1e-9) return 0;
  *pen = R - dist;
  for (int k = 0; k < 3; k++) nl[k] = (c[k] - qb[k]) / dist;
  return 1;
}
/* test hook: the query alone (body-frame centre) */
int phys_ref_hull_sphere(const b200_model_t* m, const float* verts, int b, const double* c, double R, double* pen, double* nl) {
  real p = 0, n[3] = {0, 0, 0}, cc[3] = {c[0], c[1], c[2]};
  int hit = (g_planes && g_ntris[b] > 0) ? hull_sphere_ref(m, verts, b, cc, R, &p, n) : -1;
  *pen = p; nl[0] = n[0]; nl[1] = n[1]; nl[2] = n[2];
  return hit;
}

/* float64 restatement of csrc/packed.cuh::pk_ball_contacts_extra: the ball against the bodies (hull vertices as spheres) and the racket
 * handle (capsule), poses of the start of the substep, deepest contact only, kinematic obstacle */
static void ball_contacts_extra_ref(const b200_model_t* m, const float* verts, const b200_cfg_t* cfg, const float* vrho, const body_t* B,
                                    ball_t* ball) {
  real R = cfg->ball_radius;
  {
    real dx = ball->p[0] - B[0].p[0], dy = ball->p[1] - B[0].p[1], dz = ball->p[2] - B[0].p[2];
    if (dx * dx + dy * dy + dz * dz > 2.25) return;
  }
  real best = 0, bn[3] = {0, 0, 1}, bvo[3] = {0, 0, 0}, be = cfg->ball_e_body, bmu = cfg->ball_mu_body;
  for (int b = 0; b < m->nb; b++) {
    const int handle = b == cfg->racket_body && cfg->racket_handle[6] > 0;
    const int nv = m->nverts[b];
    if (nv == 0 && !handle) continue;
    const body_t* bd = &B[b];
    real d[3] = {ball->p[0] - bd->p[0], ball->p[1] - bd->p[1], ball->p[2] - bd->p[2]};
    real d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    real reach = handle ? (real)0.6 : (real)m->radius[b] + ((g_planes && g_ntris[b] > 0) ? (real)0 : (real)vrho[b]) + R;
    if (d2 > reach * reach) continue;
    real dl[3], pen = 0, nl[3] = {0, 0, 1};
    q_rot_inv(bd->Q, d, dl);
    if (!handle && g_planes && g_ntris[b] > 0) {     /* exact: the body's convex hull */
      if (!hull_sphere_ref(m, verts, b, dl, R, &pen, nl)) continue;
    } else {
      real el[3] = {0, 0, 0}, dist = 0, rad = 0;
      if (handle) {
        const float* hd = cfg->racket_handle;
        real a[3] = {(real)hd[3] - hd[0], (real)hd[4] - hd[1], (real)hd[5] - hd[2]};
        real q0[3] = {dl[0] - hd[0], dl[1] - hd[1], dl[2] - hd[2]};
        real t = (q0[0] * a[0] + q0[1] * a[1] + q0[2] * a[2]) / (a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        t = t < 0 ? 0 : (t > 1 ? 1 : t);
        for (int k = 0; k < 3; k++) el[k] = q0[k] - t * a[k];
        dist = sqrt(el[0] * el[0] + el[1] * el[1] + el[2] * el[2]);
        rad = hd[6];
      } else {                                         /* no faces installed: spheres on the hull vertices */
        const float* vb = verts + (size_t)b * m->vmax * 3;
        real m2 = 1e30;
        for (int k = 0; k < nv; k++) {
          real ex = dl[0] - vb[k * 3], ey = dl[1] - vb[k * 3 + 1], ez = dl[2] - vb[k * 3 + 2];
          real e2 = ex * ex + ey * ey + ez * ez;
          if (e2 < m2) { m2 = e2; el[0] = ex; el[1] = ey; el[2] = ez; }
        }
        dist = sqrt(m2);
        rad = vrho[b];
      }
      pen = R + rad - dist;
      if (!(dist > 1e-9)) continue;
      nl[0] = el[0] / dist; nl[1] = el[1] / dist; nl[2] = el[2] / dist;
    }
    if (pen > best) {
      best = pen;
      q_rot(bd->Q, nl, bn);
      real x[3] = {d[0] - R * bn[0], d[1] - R * bn[1], d[2] - R * bn[2]}, wxx[3];
      cross(bd->w, x, wxx);
      for (int k = 0; k < 3; k++) bvo[k] = bd->v[k] + wxx[k];
      be = handle ? cfg->ball_e_racket : cfg->ball_e_body;
      bmu = handle ? cfg->ball_mu_racket : cfg->ball_mu_body;
    }
  }
  if (best > 0) {
    real J[3];
    ball_impulse_ref(cfg, ball, bn, bvo, be, bmu, J);
    for (int k = 0; k < 3; k++) ball->p[k] += best * bn[k];
  }
}

/* One control step (control_freq_inv sim steps x substeps) for n envs.  Same I/O contract as
 * b200env_physics_only(prec=1) in include/b200env.h.  Returns 0, or -(env+1) on a failed solve. */
int phys_ref_control_step_ball(const b200_model_t* m, const float* verts, const b200_cfg_t* cfg, int n, double* root,
                               double* dof_pos, double* dof_vel, const double* pd_tar, const double* ext_wrench,
                               double* rb_out, double* contact_out, double* ballio, int32_t* hits) {
  int nb = m->nb, nd = m->nd, fail = 0;
  real h = (real)cfg->sim_dt / cfg->substeps;
  const int with_ball = cfg->has_ball && ballio != 0;
  float vrho[B200_MAX_BODIES];
  hull_vertex_radius_ref(m, verts, vrho);
#pragma omp parallel for schedule(static)
  for (int e = 0; e < n; e++) {
    body_t B[B200_MAX_BODIES];
    memset(B, 0, sizeof(B));
    real* rs = root + (size_t)e * 13;
    memcpy(B[0].p, rs, 3 * sizeof(real)); memcpy(B[0].Q, rs + 3, 4 * sizeof(real)); q_norm(B[0].Q);
    memcpy(B[0].v, rs + 7, 3 * sizeof(real)); memcpy(B[0].w, rs + 10, 3 * sizeof(real));
    for (int i = 1; i < nb; i++) {
      if (m->fixed[i]) continue;
      int d0 = m->dof_of_body[i];
      q_exp(dof_pos + (size_t)e * nd + d0, B[i].qj);
      memcpy(B[i].wt, dof_vel + (size_t)e * nd + d0, 3 * sizeof(real));
    }
    ball_t ball;
    memset(&ball, 0, sizeof(ball));
    if (with_ball) {
      real* bs = ballio + (size_t)e * 13;
      memcpy(ball.p, bs, 3 * sizeof(real)); memcpy(ball.v, bs + 7, 3 * sizeof(real)); memcpy(ball.w, bs + 10, 3 * sizeof(real));
    }
    real cf[3 * B200_MAX_BODIES];
    int err = 0;
    for (int s = 0; s < cfg->control_freq_inv && !err; s++) {
      if (with_ball) {
        ball_aero_ref(ball.v, ball.w, cfg->spin_scale, ball.fa);
        real thr = cfg->substeps > 2 ? cfg->ball_radius * 6 : cfg->ball_radius * 4;
        if (!ball.has_bounce && ball.p[2] <= thr) { ball.has_bounce = 1; ball.bounce_now = 1; memcpy(ball.bpos, ball.p, sizeof(ball.bpos)); }
      }
      for (int k = 0; k < cfg->substeps && !err; k++) {
        real react[6] = {ball.rF[0], ball.rF[1], ball.rF[2], ball.rX[0], ball.rX[1], ball.rX[2]};
        const body_t root0 = B[0];   /* substep() integrates the root in place; the other bodies keep their start-of-substep pose */
        err = substep(m, verts, cfg, h, B, pd_tar + (size_t)e * nd, (s == 0 && ext_wrench) ? ext_wrench + (size_t)e * 6 : 0, cf,
                      with_ball ? react : 0);
        if (with_ball && !err) {
          ball_substep_ref(cfg, h, &ball, cfg->racket_body >= 0 ? &B[cfg->racket_body] : 0);
          if (cfg->ball_body_contact) {   /* every body at its start-of-substep pose, like the kernel's records at that point */
            const body_t root1 = B[0];
            B[0] = root0;
            ball_contacts_extra_ref(m, verts, cfg, vrho, B, &ball);
            B[0] = root1;
          }
        }
      }
    }
    if (err) {
#pragma omp critical
      if (!fail) fail = -(e + 1);
      continue;
    }
    fk(m, B);
    memcpy(rs, B[0].p, 3 * sizeof(real)); memcpy(rs + 3, B[0].Q, 4 * sizeof(real));
    memcpy(rs + 7, B[0].v, 3 * sizeof(real)); memcpy(rs + 10, B[0].w, 3 * sizeof(real));
    for (int i = 1; i < nb; i++) {
      if (m->fixed[i]) continue;
      int d0 = m->dof_of_body[i];
      q_log(B[i].qj, dof_pos + (size_t)e * nd + d0);
      memcpy(dof_vel + (size_t)e * nd + d0, B[i].wt, 3 * sizeof(real));
    }
    for (int i = 0; i < nb; i++) {
      real* o = rb_out + ((size_t)e * nb + i) * 13;
      memcpy(o, B[i].p, 3 * sizeof(real)); memcpy(o + 3, B[i].Q, 4 * sizeof(real));
      memcpy(o + 7, B[i].v, 3 * sizeof(real)); memcpy(o + 10, B[i].w, 3 * sizeof(real));
      if (contact_out) memcpy(contact_out + ((size_t)e * nb + i) * 3, cf + 3 * i, 3 * sizeof(real));
    }
    if (with_ball) {
      real* bs = ballio + (size_t)e * 13;
      memcpy(bs, ball.p, 3 * sizeof(real)); memcpy(bs + 7, ball.v, 3 * sizeof(real)); memcpy(bs + 10, ball.w, 3 * sizeof(real));
      if (hits) hits[e] = ball.hits;
    }
  }
  return fail;
}

int phys_ref_control_step(const b200_model_t* m, const float* verts, const b200_cfg_t* cfg, int n, double* root,
                          double* dof_pos, double* dof_vel, const double* pd_tar, const double* ext_wrench,
                          double* rb_out, double* contact_out) {
  return phys_ref_control_step_ball(m, verts, cfg, n, root, dof_pos, dof_vel, pd_tar, ext_wrench, rb_out, contact_out, 0, 0);
}

/* diagnostics for invariant tests: total mass, COM, linear momentum, angular momentum about the
 * COM, kinetic energy and gravitational potential of one env state.  out[0..13]. */
/* number of OpenMP threads of the env loop (bench.py picks the count that is fastest on the box: on a 64-core / 128-thread host,
 * 1024 envs per step run 3.5x faster on 32 threads than on 128); returns the previous setting, 1 without OpenMP */
#ifdef _OPENMP
#include <omp.h>
int phys_ref_set_threads(int n) { int old = omp_get_max_threads(); if (n > 0) omp_set_num_threads(n); return old; }
#else
int phys_ref_set_threads(int n) { (void)n; return 1; }
#endif

int phys_ref_diagnostics(const b200_model_t* m, const b200_cfg_t* cfg, const double* root, const double* dof_pos,
                         const double* dof_vel, double* out) {
  body_t B[B200_MAX_BODIES];
  memset(B, 0, sizeof(B));
  memcpy(B[0].p, root, 3 * sizeof(real)); memcpy(B[0].Q, root + 3, 4 * sizeof(real)); q_norm(B[0].Q);
  memcpy(B[0].v, root + 7, 3 * sizeof(real)); memcpy(B[0].w, root + 10, 3 * sizeof(real));
  for (int i = 1; i < m->nb; i++) {
    if (m->fixed[i]) continue;
    q_exp(dof_pos + m->dof_of_body[i], B[i].qj);
    memcpy(B[i].wt, dof_vel + m->dof_of_body[i], 3 * sizeof(real));
  }
  fk(m, B);
  real M = 0, com[3] = {0, 0, 0}, P[3] = {0, 0, 0}, L0[3] = {0, 0, 0}, ke = 0, pe = 0;
  for (int i = 0; i < m->nb; i++) {
    if (m->fixed[i]) continue;
    body_t* b = &B[i];
    real ms = m->mass[i], cl[3] = {m->com[i][0], m->com[i][1], m->com[i][2]}, c[3], x[3], wxc[3], vc[3];
    matvec3(b->R, cl, c); cross(b->w, c, wxc);
    for (int k = 0; k < 3; k++) { x[k] = b->p[k] + c[k]; vc[k] = b->v[k] + wxc[k]; }
    real Ib[3][3] = {{m->inertia[i][0], m->inertia[i][3], m->inertia[i][4]},
                     {m->inertia[i][3], m->inertia[i][1], m->inertia[i][5]},
                     {m->inertia[i][4], m->inertia[i][5], m->inertia[i][2]}};
    real wl[3], Iwl[3], Iw[3], xv[3];
    matTvec3(b->R, b->w, wl); matvec3(Ib, wl, Iwl); matvec3(b->R, Iwl, Iw); cross(x, vc, xv);
    M += ms;
    for (int k = 0; k < 3; k++) { com[k] += ms * x[k]; P[k] += ms * vc[k]; L0[k] += Iw[k] + ms * xv[k]; }
    ke += 0.5 * ms * (vc[0] * vc[0] + vc[1] * vc[1] + vc[2] * vc[2]) + 0.5 * (wl[0] * Iwl[0] + wl[1] * Iwl[1] + wl[2] * Iwl[2]);
    pe += -ms * cfg->gravity_z * x[2];
  }
  for (int k = 0; k < 3; k++) com[k] /= M;
  real cP[3]; cross(com, P, cP);
  out[0] = M;
  for (int k = 0; k < 3; k++) { out[1 + k] = com[k]; out[4 + k] = P[k]; out[7 + k] = L0[k] - cP[k]; }
  out[10] = ke; out[11] = pe;
  return 0;
}
