// b200env.cu - B200 (sm_100a) kernels + C ABI of the vectorised humanoid environment step.
//
// One warp per env, one lane per rigid body.  Per-asset constants (tree tables, inertias, PD
// gains, convex-hull vertices) are staged once per CTA into shared memory with a TMA bulk copy
// (cp.async.bulk + mbarrier).  The three articulated-body passes run level by level over the
// kinematic tree with warp shuffles between parent and child lanes; everything between the
// state load and the obs/reward/reset store stays in registers.  No tensor cores: there is no
// dense contraction on this path (BASELINE.json north_star).
//
// Reference behaviour being replaced (paths relative to /root/reference/embodied_pose):
//   BaseTask.step                      env/tasks/base_task.py:147-165
//   pre_physics_step / PD target clamp env/tasks/humanoid_smpl_im.py:125-157, 391-396
//   gym.simulate x controlFreqInv      env/tasks/base_task.py:450-454   (physics: DESIGN.md 3)
//   post_physics_step                  env/tasks/humanoid_smpl_im.py:398-418
//   MotionLib.get_motion_state         utils/motion_lib.py:164-266, 427-436, 460-488
//   obs / reward / reset               env/tasks/humanoid_smpl_im.py:653-668, 773-850, 918-987
//   quaternion helpers                 utils/torch_utils.py:70-243
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/b200env.h"
#include "dyn_common.cuh"

#ifndef WARPS_PER_CTA
#define WARPS_PER_CTA 8  // A/B on B200 (profiles/r1_notes.md): 8 warps x 2 CTAs/SM, barrier per substep
#endif
#ifndef PT_STEP_SYNC
#define PT_STEP_SYNC 1   // step_kernel_tmem: CTA barrier at every substep (keeps the 14 warps on the same instruction-cache lines)
#endif
#ifndef STEP_SYNC
#define STEP_SYNC 0  // (packed kernel A/B: 434 us without any barrier, 416 with this one, 391 with POST_SYNC only)
// legacy note: 1: CTA barrier at every substep boundary keeps the warps of a CTA on the same code (I-cache sharing)
#endif
#define SCRATCH_FLOATS 320
#ifndef POST_SYNC
#define POST_SYNC 1
#endif
#ifndef PK_WARP_TICKETS
#define PK_WARP_TICKETS 0   // packed kernel: 1 = every warp pulls its own groups of EPW envs (no CTA barrier in the loop)
#endif
#ifndef STEP_MIN_CTAS
#define STEP_MIN_CTAS 2
#endif


struct b200env {
  int device;
  int num_envs;
  b200_model_t model;
  b200_cfg_t cfg;
  b200_buffers_t bufs;
  b200_motion_lib_t ml;
  bool bound, has_ml;
  void* d_blob;
  size_t blob_bytes;
  float* d_face_planes;                // b200env_set_hull_faces
  unsigned char* d_face_tris;
  b200_cfg_t* d_cfg;
  unsigned long long* d_ticket;
  unsigned long long ticket_base;
  float* d_ext;                        // [num_envs rows of the bound tensors, 6] residual root wrench between pre_kernel and the physics launch
  int ext_rows;
  int sort;                            // 1: envs handed out by contact load (B200ENV_SORT=1; split form, 4-envs-per-warp kernel)
  int32_t *d_bin, *d_perm;             // [num_envs]
  uint32_t *d_pos, *d_cnt;             // [num_envs], [SORT_BINS]
  int packed3;                         // 1: the physics launch is step_kernel_packed3 (B200ENV_KERNEL=packed3)
  int tmem;                            // 1: the physics launch is step_kernel_tmem (one wave, private fields in tensor memory; B200ENV_KERNEL=tmem|packed)
  int split;                           // 1: three launches (pre / physics / post), 0: one fused launch.  env B200ENV_SPLIT=0|1
  int env_first = 0, env_stride = 1;   // b200env_set_env_slice: local env i = row env_first + env_stride * i of the bound tensors
  int step_grid;
  int packed;        // 1: 4-envs-per-warp kernels (packed.cuh); 0: lane-per-body kernels.  env B200ENV_KERNEL=lane|packed
  int packed_ok;     // tree fits the packed layout (<= 8 bodies per depth, <= 25 bodies)
  int64_t launches;
  int timing;                          // b200env_set_kernel_timing: event pairs around the physics launch
  std::vector<cudaEvent_t>* tev;       // recorded (start, stop) pairs
};

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}
#define CUDA_OK(x)                                                   \
  do {                                                               \
    cudaError_t _e = (x);                                            \
    if (_e != cudaSuccess) return fail(-10, "CUDA error: %s", cudaGetErrorString(_e)); \
  } while (0)

template <typename T>
__device__ __forceinline__ void substep(const DevBlob& B, const float* __restrict__ verts, const PhysCfg<T>& c,
                                        const LaneConst& lc, int lane, Lane<T>& L, const T* pdtar, bool ext_on,
                                        const T* extF, const T* extT, T* cf, Ball<T>& ball, bool lsync = false) {
  const b200_model_t& M = B.m;
  T r[3] = {0, 0, 0}, zeta[6] = {0, 0, 0, 0, 0, 0};
  fk_pass<T, true>(M, lc, lane, L, r, zeta, lsync);

  // articulated inertia  [[A, Bm], [Bm^T, C]]  (A, C symmetric) and bias (bn, bf)
  T A[6] = {0, 0, 0, 0, 0, 0}, Bm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, C[6] = {0, 0, 0, 0, 0, 0};
  T bn[3] = {0, 0, 0}, bf[3] = {0, 0, 0};
  T E[6] = {0, 0, 0, 0, 0, 0}, u[3] = {0, 0, 0};
  T R[9];
  qmat(L.Q, R);
  cf[0] = cf[1] = cf[2] = T(0);
  T rF[3] = {0, 0, 0}, rX[3] = {0, 0, 0};
  if (c.has_ball && c.racket_body >= 0) {  // (warp-uniform branch) fetch the ball lane's last racket reaction
#pragma unroll
    for (int k = 0; k < 3; k++) { rF[k] = shfl(ball.rF[k], BALL_LANE); rX[k] = shfl(ball.rX[k], BALL_LANE); }
  }
  if (lc.dyn) {
    const T ms = T(M.mass[lane]);
    T cl[3] = {T(M.com[lane][0]), T(M.com[lane][1]), T(M.com[lane][2])}, cw[3];
    mv3(R, cl, cw);
    // Io = R Ib R^T + m (|c|^2 1 - c c^T)
    {
      T Ib[6];
#pragma unroll
      for (int k = 0; k < 6; k++) Ib[k] = T(M.inertia[lane][k]);
      T F[9], RF[9];
      sym_full(Ib, F);
      // RF = R * F
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) RF[i * 3 + j] = R[i * 3] * F[j] + R[i * 3 + 1] * F[3 + j] + R[i * 3 + 2] * F[6 + j];
      T c2 = cw[0] * cw[0] + cw[1] * cw[1] + cw[2] * cw[2];
      A[0] = RF[0] * R[0] + RF[1] * R[1] + RF[2] * R[2] + ms * (c2 - cw[0] * cw[0]);
      A[1] = RF[3] * R[3] + RF[4] * R[4] + RF[5] * R[5] + ms * (c2 - cw[1] * cw[1]);
      A[2] = RF[6] * R[6] + RF[7] * R[7] + RF[8] * R[8] + ms * (c2 - cw[2] * cw[2]);
      A[3] = RF[0] * R[3] + RF[1] * R[4] + RF[2] * R[5] - ms * cw[0] * cw[1];
      A[4] = RF[0] * R[6] + RF[1] * R[7] + RF[2] * R[8] - ms * cw[0] * cw[2];
      A[5] = RF[3] * R[6] + RF[4] * R[7] + RF[5] * R[8] - ms * cw[1] * cw[2];
    }
    // Bm = m [c]x
    Bm[1] = -ms * cw[2]; Bm[2] = ms * cw[1];
    Bm[3] = ms * cw[2]; Bm[5] = -ms * cw[0];
    Bm[6] = -ms * cw[1]; Bm[7] = ms * cw[0];
    C[0] = C[1] = C[2] = ms;
    // bias: [w x Io w ; m w x (w x c)] - gravity wrench
    {
      T Iw[3], t1[3], t2[3];
      sym_mv(A, L.w, Iw);
      cross3(L.w, Iw, bn);
      cross3(L.w, cw, t1);
      cross3(L.w, t1, t2);
      // c x (m g) with g = (0,0,gz):  (c_y gz, -c_x gz, 0) * m
      bn[0] -= ms * cw[1] * c.gz;
      bn[1] += ms * cw[0] * c.gz;
      bf[0] = ms * t2[0]; bf[1] = ms * t2[1]; bf[2] = ms * t2[2] - ms * c.gz;
    }
    if (c.has_ball && c.racket_body >= 0) {  // reaction of the previous substep's racket impact, applied to the wrist link
      if (lane == M.parent[c.racket_body]) {
        T dx[3] = {rX[0] - L.p[0], rX[1] - L.p[1], rX[2] - L.p[2]}, t[3];
        cross3(dx, rF, t);
#pragma unroll
        for (int k = 0; k < 3; k++) { bn[k] -= t[k]; bf[k] -= rF[k]; }
      }
    }
    if (lane == 0 && ext_on) {  // residual wrench: force at the COM + torque, world frame
      T cxF[3];
      cross3(cw, extF, cxF);
#pragma unroll
      for (int k = 0; k < 3; k++) { bn[k] -= extT[k] + cxF[k]; bf[k] -= extF[k]; }
    }
    // ground contact of the hull vertices, implicit in the velocity
    const int nv = M.nverts[lane];
    if (nv > 0 && L.p[2] - T(M.radius[lane]) < T(0))
      contact_hull<T>(verts + (size_t)lane * M.vmax * 3, M.vmax, nv, c, R, L.p, L.v, L.w, A, Bm, C, bn, bf, cf);
    // joint drive (child frame, exp-map chart): implicit PD + armature + limit springs
    if (lane > 0) {
      T q[3], tau[3], e[3];
      qlog(L.qj, q);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        T kp = T(M.kp[lc.dof0 + k]), kd = T(M.kd[lc.dof0 + k]);
        e[k] = T(M.armature[lc.dof0 + k]) + c.h * kd + c.h * c.h * kp;
        tau[k] = kp * (pdtar[k] - q[k] - c.h * L.wt[k]) - kd * L.wt[k];
        T lo = T(M.lim_lo[lc.dof0 + k]), hi = T(M.lim_hi[lc.dof0 + k]);
        if (q[k] < lo) { tau[k] += c.limk * (lo - q[k] - c.h * L.wt[k]) - c.limc * L.wt[k]; e[k] += c.h * c.limc + c.h * c.h * c.limk; }
        else if (q[k] > hi) { tau[k] += c.limk * (hi - q[k] - c.h * L.wt[k]) - c.limc * L.wt[k]; e[k] += c.h * c.limc + c.h * c.h * c.limk; }
      }
      mv3(R, tau, u);  // tau_w
      // E_w = R diag(e) R^T
      E[0] = R[0] * R[0] * e[0] + R[1] * R[1] * e[1] + R[2] * R[2] * e[2];
      E[1] = R[3] * R[3] * e[0] + R[4] * R[4] * e[1] + R[5] * R[5] * e[2];
      E[2] = R[6] * R[6] * e[0] + R[7] * R[7] * e[1] + R[8] * R[8] * e[2];
      E[3] = R[0] * R[3] * e[0] + R[1] * R[4] * e[1] + R[2] * R[5] * e[2];
      E[4] = R[0] * R[6] * e[0] + R[1] * R[7] * e[1] + R[2] * R[8] * e[2];
      E[5] = R[3] * R[6] * e[0] + R[4] * R[7] * e[1] + R[5] * R[8] * e[2];
    }
  }

  // ---- backward pass: leaves -> root, one tree level at a time
  T Dinv[6] = {0, 0, 0, 0, 0, 0};
  for (int d = M.max_depth; d >= 1; d--) {
    if (LEVEL_SYNC && lsync) __syncthreads();
    T out[27];
#pragma unroll
    for (int k = 0; k < 27; k++) out[k] = T(0);
    if (lc.dyn && lc.depth == d) {
      T D[6];
#pragma unroll
      for (int k = 0; k < 6; k++) D[k] = A[k] + E[k];
      sym_inv(D, Dinv);
      u[0] -= bn[0]; u[1] -= bn[1]; u[2] -= bn[2];
      T Af[9], Df[9];
      sym_full(A, Af);
      sym_full(Dinv, Df);
      // G = Dinv * A ; K = Dinv * Bm   (3x3 full)
      T G[9], K[9];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          G[i * 3 + j] = Df[i * 3] * Af[j] + Df[i * 3 + 1] * Af[3 + j] + Df[i * 3 + 2] * Af[6 + j];
          K[i * 3 + j] = Df[i * 3] * Bm[j] + Df[i * 3 + 1] * Bm[3 + j] + Df[i * 3 + 2] * Bm[6 + j];
        }
      // Ia blocks: aA = A - A G (sym), aB = Bm - A K, aC = C - Bm^T K (sym)
      T aA[6], aB[9], aC[6];
      aA[0] = A[0] - (Af[0] * G[0] + Af[1] * G[3] + Af[2] * G[6]);
      aA[1] = A[1] - (Af[3] * G[1] + Af[4] * G[4] + Af[5] * G[7]);
      aA[2] = A[2] - (Af[6] * G[2] + Af[7] * G[5] + Af[8] * G[8]);
      aA[3] = A[3] - (Af[0] * G[1] + Af[1] * G[4] + Af[2] * G[7]);
      aA[4] = A[4] - (Af[0] * G[2] + Af[1] * G[5] + Af[2] * G[8]);
      aA[5] = A[5] - (Af[3] * G[2] + Af[4] * G[5] + Af[5] * G[8]);
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) aB[i * 3 + j] = Bm[i * 3 + j] - (Af[i * 3] * K[j] + Af[i * 3 + 1] * K[3 + j] + Af[i * 3 + 2] * K[6 + j]);
      aC[0] = C[0] - (Bm[0] * K[0] + Bm[3] * K[3] + Bm[6] * K[6]);
      aC[1] = C[1] - (Bm[1] * K[1] + Bm[4] * K[4] + Bm[7] * K[7]);
      aC[2] = C[2] - (Bm[2] * K[2] + Bm[5] * K[5] + Bm[8] * K[8]);
      aC[3] = C[3] - (Bm[0] * K[1] + Bm[3] * K[4] + Bm[6] * K[7]);
      aC[4] = C[4] - (Bm[0] * K[2] + Bm[3] * K[5] + Bm[6] * K[8]);
      aC[5] = C[5] - (Bm[1] * K[2] + Bm[4] * K[5] + Bm[7] * K[8]);
      // ba = b + Ia zeta + Ucol Dinv u
      T s[3], an[3], af[3], t1[3], t2[3];
      sym_mv(Dinv, u, s);
      sym_mv(aA, zeta, t1);
      mv3(aB, zeta + 3, t2);
      T As[3];
      sym_mv(A, s, As);
#pragma unroll
      for (int k = 0; k < 3; k++) an[k] = bn[k] + t1[k] + t2[k] + As[k];
      mtv3(aB, zeta, t1);
      sym_mv(aC, zeta + 3, t2);
      T Bts[3];
      mtv3(Bm, s, Bts);
#pragma unroll
      for (int k = 0; k < 3; k++) af[k] = bf[k] + t1[k] + t2[k] + Bts[k];
      // shift to the parent's origin (X = [r]x):  B' = aB + X aC ; A' = aA + X B'^T + (X aB^T)^T
      T Cf[9];
      sym_full(aC, Cf);
      T Bp[9];
#pragma unroll
      for (int j = 0; j < 3; j++) {  // column j of X*Cf = r x (column j of Cf)
        T col[3] = {Cf[j], Cf[3 + j], Cf[6 + j]}, x[3];
        cross3(r, col, x);
        Bp[j] = aB[j] + x[0]; Bp[3 + j] = aB[3 + j] + x[1]; Bp[6 + j] = aB[6 + j] + x[2];
      }
      T P1[9], P2[9];  // P1 = X Bp^T (column j = r x row j of Bp), P2 = X aB^T
#pragma unroll
      for (int j = 0; j < 3; j++) {
        T x[3];
        cross3(r, Bp + 3 * j, x);
        P1[j] = x[0]; P1[3 + j] = x[1]; P1[6 + j] = x[2];
        cross3(r, aB + 3 * j, x);
        P2[j] = x[0]; P2[3 + j] = x[1]; P2[6 + j] = x[2];
      }
      out[0] = aA[0] + P1[0] + P2[0];
      out[1] = aA[1] + P1[4] + P2[4];
      out[2] = aA[2] + P1[8] + P2[8];
      out[3] = aA[3] + P1[1] + P2[3];
      out[4] = aA[4] + P1[2] + P2[6];
      out[5] = aA[5] + P1[5] + P2[7];
#pragma unroll
      for (int k = 0; k < 9; k++) out[6 + k] = Bp[k];
#pragma unroll
      for (int k = 0; k < 6; k++) out[15 + k] = aC[k];
      T rxf[3];
      cross3(r, af, rxf);
#pragma unroll
      for (int k = 0; k < 3; k++) { out[21 + k] = an[k] + rxf[k]; out[24 + k] = af[k]; }
    }
    const int rounds = B.t.maxch[d - 1];
    for (int cix = 0; cix < rounds; cix++) {
      int src = (lc.active && lc.depth == d - 1) ? B.t.child[lane][cix] : -1;
      const bool has = src >= 0;
      src = has ? src : lane;
#pragma unroll
      for (int k = 0; k < 27; k++) {
        T val = shfl(out[k], src);
        if (has) {
          if (k < 6) A[k] += val;
          else if (k < 15) Bm[k - 6] += val;
          else if (k < 21) C[k - 15] += val;
          else if (k < 24) bn[k - 21] += val;
          else bf[k - 24] += val;
        }
      }
    }
  }

  // ---- root: solve [[A,Bm],[Bm^T,C]] [alpha; a] = -[bn; bf] by block elimination
  T acc[6] = {0, 0, 0, 0, 0, 0};  // (alpha, a) of this body
  if (lane == 0) {
    T Ci[6];
    sym_inv(C, Ci);
    T Cif[9];
    sym_full(Ci, Cif);
    T BC[9];  // Bm * Ci
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) BC[i * 3 + j] = Bm[i * 3] * Cif[j] + Bm[i * 3 + 1] * Cif[3 + j] + Bm[i * 3 + 2] * Cif[6 + j];
    T S[6];  // A - Bm Ci Bm^T
    S[0] = A[0] - (BC[0] * Bm[0] + BC[1] * Bm[1] + BC[2] * Bm[2]);
    S[1] = A[1] - (BC[3] * Bm[3] + BC[4] * Bm[4] + BC[5] * Bm[5]);
    S[2] = A[2] - (BC[6] * Bm[6] + BC[7] * Bm[7] + BC[8] * Bm[8]);
    S[3] = A[3] - (BC[0] * Bm[3] + BC[1] * Bm[4] + BC[2] * Bm[5]);
    S[4] = A[4] - (BC[0] * Bm[6] + BC[1] * Bm[7] + BC[2] * Bm[8]);
    S[5] = A[5] - (BC[3] * Bm[6] + BC[4] * Bm[7] + BC[5] * Bm[8]);
    T Si[6];
    sym_inv(S, Si);
    T rhs[3], t[3];
    mv3(BC, bf, t);
#pragma unroll
    for (int k = 0; k < 3; k++) rhs[k] = -bn[k] + t[k];
    sym_mv(Si, rhs, acc);
    mtv3(Bm, acc, t);
#pragma unroll
    for (int k = 0; k < 3; k++) t[k] = -bf[k] - t[k];
    sym_mv(Ci, t, acc + 3);
  }

  // ---- forward pass: root -> leaves
  for (int d = 1; d <= M.max_depth; d++) {
    if (LEVEL_SYNC && lsync) __syncthreads();
    T pa[6];
#pragma unroll
    for (int k = 0; k < 6; k++) pa[k] = shfl(acc[k], lc.par);
    if (lc.dyn && lc.depth == d) {
      T axr[3], Ap[6];
      cross3(pa, r, axr);
#pragma unroll
      for (int k = 0; k < 3; k++) { Ap[k] = pa[k] + zeta[k]; Ap[3 + k] = pa[3 + k] + axr[k] + zeta[3 + k]; }
      // t = u - Ucol^T A' = u - A Ap_ang - Bm Ap_lin
      T t1[3], t2[3], t[3], gam[3], wd[3];
      sym_mv(A, Ap, t1);
      mv3(Bm, Ap + 3, t2);
#pragma unroll
      for (int k = 0; k < 3; k++) t[k] = u[k] - t1[k] - t2[k];
      sym_mv(Dinv, t, gam);
#pragma unroll
      for (int k = 0; k < 3; k++) { acc[k] = Ap[k] + gam[k]; acc[3 + k] = Ap[3 + k]; }
      mtv3(R, gam, wd);
#pragma unroll
      for (int k = 0; k < 3; k++) L.wt[k] = (L.wt[k] + c.h * wd[k]) * c.damp;
      T n2 = L.wt[0] * L.wt[0] + L.wt[1] * L.wt[1] + L.wt[2] * L.wt[2];
      if (n2 > c.wmax * c.wmax) { T sc = c.wmax * rsqrt_(n2); L.wt[0] *= sc; L.wt[1] *= sc; L.wt[2] *= sc; }
    }
  }

  // ---- integrate
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) { L.w[k] = (L.w[k] + c.h * acc[k]) * c.damp; L.v[k] += c.h * acc[3 + k]; }
    T n2 = L.w[0] * L.w[0] + L.w[1] * L.w[1] + L.w[2] * L.w[2];
    if (n2 > c.wmax * c.wmax) { T sc = c.wmax * rsqrt_(n2); L.w[0] *= sc; L.w[1] *= sc; L.w[2] *= sc; }
    T hv[3] = {c.h * L.w[0], c.h * L.w[1], c.h * L.w[2]}, dq[4], qn[4];
#pragma unroll
    for (int k = 0; k < 3; k++) L.p[k] += c.h * L.v[k];
    qexp_small(hv, dq);
    qmul(dq, L.Q, qn);
    qnormalize(qn);
#pragma unroll
    for (int k = 0; k < 4; k++) L.Q[k] = qn[k];
  } else if (lc.dyn) {
    T hv[3] = {c.h * L.wt[0], c.h * L.wt[1], c.h * L.wt[2]}, dq[4], qn[4];
    qexp_small(hv, dq);
    qmul(L.qj, dq, qn);
    qnormalize(qn);
#pragma unroll
    for (int k = 0; k < 4; k++) L.qj[k] = qn[k];
  }
  // ---- ball (lane BALL_LANE); the welded racket lane still holds its start-of-substep pose and velocity
  if (c.has_ball) {
    T rQ[4] = {0, 0, 0, 1}, rp[3] = {0, 0, 0}, rv[3] = {0, 0, 0}, rw[3] = {0, 0, 0};
    const bool has_racket = c.racket_body >= 0;
    if (has_racket) {
#pragma unroll
      for (int k = 0; k < 4; k++) rQ[k] = shfl(L.Q[k], c.racket_body);
#pragma unroll
      for (int k = 0; k < 3; k++) { rp[k] = shfl(L.p[k], c.racket_body); rv[k] = shfl(L.v[k], c.racket_body); rw[k] = shfl(L.w[k], c.racket_body); }
    }
    if (lane == BALL_LANE) ball_substep<T>(c, ball, has_racket, rQ, rp, rv, rw);
  }
}

// control_freq_inv sim steps x substeps; external wrench only during the first sim step
template <typename T>
__device__ __forceinline__ void control_step(const DevBlob& B, const float* verts, const PhysCfg<T>& c, const LaneConst& lc,
                                             int lane, Lane<T>& L, const T* pdtar, const T* extF, const T* extT, T* cf,
                                             Ball<T>& ball, bool cta_sync = false) {
  for (int s = 0; s < c.cfi; s++) {
    if (c.has_ball && lane == BALL_LANE) {  // once per sim step, like apply_external_force_to_ball (:752-756)
      ball_aero<T>(ball.v, ball.w, c.spin_scale, ball.fa);
      const T thr = c.substeps > 2 ? c.bR * T(6) : c.bR * T(4);
      if (!ball.has_bounce && ball.p[2] <= thr) {
        ball.has_bounce = 1; ball.bounce_now = 1;
        ball.bpos[0] = ball.p[0]; ball.bpos[1] = ball.p[1]; ball.bpos[2] = ball.p[2];
      }
    }
    for (int k = 0; k < c.substeps; k++) {
      if (cta_sync) __syncthreads();
      substep<T>(B, verts, c, lc, lane, L, pdtar, s == 0, extF, extT, cf, ball, cta_sync);
    }
  }
  T dummy[3], dz[6];
  fk_pass<T, false>(B.m, lc, lane, L, dummy, dz);
}

#include "packed.cuh"
#include "packed_t.cuh"
// A/B variants that lost their measurements live outside the product build: -DB200ENV_WITH_PACKED3=1 compiles the
// 2-envs-per-warp / 3-lanes-per-body kernel of tools/variants/packed3.cuh back in (B200ENV_KERNEL=packed3; 16 % slower, profiles/r1j, r2a)
#ifndef B200ENV_WITH_PACKED3
#define B200ENV_WITH_PACKED3 0
#endif
#if B200ENV_WITH_PACKED3
#include "../../tools/variants/packed3.cuh"
#endif

// ------------------------------------------------------------------------------------------
// TMA bulk load of the constant block into shared memory (one elected thread issues it)
__device__ __forceinline__ void load_blob(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
  const uint32_t mb = (uint32_t)__cvta_generic_to_shared(mbar);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(gsrc), "r"(bytes), "r"(mb)
                 : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mb)
        : "memory");
  }
}


// ------------------------------------------------------------------------------------------
// reference float32 helpers (utils/torch_utils.py) - op order kept close to the reference
__device__ __forceinline__ void ref_quat_rotate(const float* q, const float* v, float* o) {  // my_quat_rotate :70-79
  float qw = q[3];
  float a = 2.0f * qw * qw - 1.0f;
  float cx = q[1] * v[2] - q[2] * v[1], cy = q[2] * v[0] - q[0] * v[2], cz = q[0] * v[1] - q[1] * v[0];
  float d = q[0] * v[0] + q[1] * v[1] + q[2] * v[2];
  o[0] = v[0] * a + cx * qw * 2.0f + q[0] * d * 2.0f;
  o[1] = v[1] * a + cy * qw * 2.0f + q[1] * d * 2.0f;
  o[2] = v[2] * a + cz * qw * 2.0f + q[2] * d * 2.0f;
}
__device__ __forceinline__ void ref_tan_norm(const float* q, float* o) {  // quat_to_tan_norm :122-134
  const float ex[3] = {1.f, 0.f, 0.f}, ez[3] = {0.f, 0.f, 1.f};
  ref_quat_rotate(q, ex, o);
  ref_quat_rotate(q, ez, o + 3);
}
__device__ __forceinline__ float ref_normalize_angle(float x) { return atan2f(sinf(x), cosf(x)); }
__device__ __forceinline__ void ref_quat_to_angle_axis(const float* q, float& angle, float* axis) {  // :82-102
  // 1 - w*w cancels for small angles: keep torch's unfused rounding (no fma contraction)
  float sin_theta = sqrtf(__fsub_rn(1.0f, __fmul_rn(q[3], q[3])));
  angle = ref_normalize_angle(2.0f * acosf(q[3]));
  bool mask = fabsf(sin_theta) > 1e-5f;
  if (mask) { axis[0] = q[0] / sin_theta; axis[1] = q[1] / sin_theta; axis[2] = q[2] / sin_theta; }
  else { angle = 0.f; axis[0] = 0.f; axis[1] = 0.f; axis[2] = 1.f; }
}
__device__ __forceinline__ void ref_exp_map_to_quat(const float* e, float* q) {  // :143-166
  float angle = sqrtf(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
  float ax[3] = {e[0] / angle, e[1] / angle, e[2] / angle};
  angle = ref_normalize_angle(angle);
  if (!(fabsf(angle) > 1e-5f)) { angle = 0.f; ax[0] = 0.f; ax[1] = 0.f; ax[2] = 1.f; }
  float n = fmaxf(sqrtf(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]), 1e-9f);  // quat_from_angle_axis: normalize(axis)
  float s = sinf(angle * 0.5f), cw = cosf(angle * 0.5f);
  q[0] = ax[0] / n * s; q[1] = ax[1] / n * s; q[2] = ax[2] / n * s; q[3] = cw;
  float qn = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-9f);  // quat_unit
  q[0] /= qn; q[1] /= qn; q[2] /= qn; q[3] /= qn;
}
__device__ __forceinline__ void ref_slerp(const float* q0, const float* q1in, float t, float* o) {  // :168-190
  float c = q0[0] * q1in[0] + q0[1] * q1in[1] + q0[2] * q1in[2] + q0[3] * q1in[3];
  float sg = c < 0.f ? -1.f : 1.f;
  float q1[4] = {sg * q1in[0], sg * q1in[1], sg * q1in[2], sg * q1in[3]};
  c = fabsf(c);
  float half = acosf(c);
  float sh = sqrtf(__fsub_rn(1.0f, __fmul_rn(c, c)));  // unfused like torch (cancellation near c = 1)
  float ra = sinf((1.f - t) * half) / sh, rb = sinf(t * half) / sh;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float v = ra * q0[k] + rb * q1[k];
    if (fabsf(sh) < 0.001f) v = 0.5f * q0[k] + 0.5f * q1[k];
    if (fabsf(c) >= 1.f) v = q0[k];
    o[k] = v;
  }
}
__device__ __forceinline__ void ref_remove_base_rot(const float* q, float* o) {  // humanoid_smpl_im.py:766-770
  const float b[4] = {-0.5f, -0.5f, -0.5f, 0.5f};
  qmul(q, b, o);
}
__device__ __forceinline__ float ref_calc_heading(const float* q) {  // :192-203
  const float ex[3] = {1.f, 0.f, 0.f};
  float d[3];
  ref_quat_rotate(q, ex, d);
  return atan2f(d[1], d[0]);
}
__device__ __forceinline__ void ref_heading_quat(float heading, float* q) {  // quat_from_angle_axis(heading, z)
  float s = sinf(heading * 0.5f), c = cosf(heading * 0.5f);
  float n = fmaxf(sqrtf(s * s + c * c), 1e-9f);
  q[0] = 0.f; q[1] = 0.f; q[2] = s / n; q[3] = c / n;
}

// MoCap sampling for one (motion id, time): lane b returns its body's blended pose.
struct MotionSample {
  float rb_pos[3], rb_rot[4], dof[3];
  float root_vel[3], root_ang_vel[3];  // every lane holds the same values
  int64_t f0;
};
__device__ __forceinline__ MotionSample sample_motion(const b200_motion_lib_t& ml, const b200_model_t& M, int64_t mid,
                                                      float time, int lane, float ground_tol) {
  MotionSample s;
  const float len = ml.motion_lengths[mid];
  const int64_t nf = ml.num_frames[mid];
  const float dt = ml.motion_dt[mid];
  // _calc_frame_blend (motion_lib.py:427-436), float32 op by op (no fma contraction)
  float phase = __fdiv_rn(time, len);
  phase = fminf(fmaxf(phase, 0.0f), 1.0f);
  int64_t i0 = (int64_t)__fmul_rn(phase, (float)(nf - 1));
  int64_t i1 = (i0 + 1 < nf - 1) ? i0 + 1 : nf - 1;
  float blend = __fdiv_rn(__fsub_rn(time, __fmul_rn((float)i0, dt)), dt);
  const int64_t f0 = i0 + ml.length_starts[mid], f1 = i1 + ml.length_starts[mid];
  s.f0 = f0;
  const int nbl = ml.num_lib_bodies;
  const float minvh = ml.min_verts_h[mid] - ground_tol;
  const float omb = __fsub_rn(1.0f, blend);
  if (lane < nbl) {
    const float* g0 = ml.gts + (f0 * nbl + lane) * 3;
    const float* g1 = ml.gts + (f1 * nbl + lane) * 3;
#pragma unroll
    for (int k = 0; k < 3; k++) s.rb_pos[k] = __fadd_rn(__fmul_rn(omb, g0[k]), __fmul_rn(blend, g1[k]));
    s.rb_pos[2] -= minvh;
    const float4 r0 = *reinterpret_cast<const float4*>(ml.grs + (f0 * nbl + lane) * 4);
    const float4 r1 = *reinterpret_cast<const float4*>(ml.grs + (f1 * nbl + lane) * 4);
    float a[4] = {r0.x, r0.y, r0.z, r0.w}, b[4] = {r1.x, r1.y, r1.z, r1.w};
    ref_slerp(a, b, blend, s.rb_rot);
    const float4 l0 = *reinterpret_cast<const float4*>(ml.lrs + (f0 * nbl + lane) * 4);
    const float4 l1 = *reinterpret_cast<const float4*>(ml.lrs + (f1 * nbl + lane) * 4);
    float la[4] = {l0.x, l0.y, l0.z, l0.w}, lb[4] = {l1.x, l1.y, l1.z, l1.w}, lq[4];
    ref_slerp(la, lb, blend, lq);
    float ang, ax[3];
    ref_quat_to_angle_axis(lq, ang, ax);  // quat_to_exp_map (:113-120)
    s.dof[0] = ang * ax[0]; s.dof[1] = ang * ax[1]; s.dof[2] = ang * ax[2];
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++) { s.rb_pos[k] = 0.f; s.dof[k] = 0.f; }
    s.rb_rot[0] = s.rb_rot[1] = s.rb_rot[2] = 0.f; s.rb_rot[3] = 1.f;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) { s.root_vel[k] = ml.grvs[f0 * 3 + k]; s.root_ang_vel[k] = ml.gravs[f0 * 3 + k]; }
  return s;
}

// write one MoCap sample as the "target" rows of env e
__device__ __forceinline__ void store_targets(const b200_buffers_t& bf, const b200_cfg_t& cfg, const b200_motion_lib_t& ml,
                                              const b200_model_t& M, const MotionSample& s, int64_t e, int lane, int dof0) {
  const int nbl = ml.num_lib_bodies, nd = M.nd;
  if (lane < nbl) {
#pragma unroll
    for (int k = 0; k < 3; k++) bf.t_rb_pos[(e * nbl + lane) * 3 + k] = s.rb_pos[k];
    *reinterpret_cast<float4*>(bf.t_rb_rot + (e * nbl + lane) * 4) = make_float4(s.rb_rot[0], s.rb_rot[1], s.rb_rot[2], s.rb_rot[3]);
    if (dof0 >= 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) bf.t_dof_pos[e * nd + dof0 + k] = s.dof[k];
    }
    for (int k = 0; k < cfg.num_key; k++)
      if (cfg.key_body[k] == lane) {
#pragma unroll
        for (int j = 0; j < 3; j++) bf.t_key_pos[(e * cfg.num_key + k) * 3 + j] = s.rb_pos[j];
      }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      bf.t_root_pos[e * 3 + k] = s.rb_pos[k];
      bf.t_root_vel[e * 3 + k] = s.root_vel[k];
      bf.t_root_ang_vel[e * 3 + k] = s.root_ang_vel[k];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) bf.t_root_rot[e * 4 + k] = s.rb_rot[k];
  }
  {
    float dv[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { const int k = lane + 32 * i; dv[i] = k < nd ? ml.dvs[s.f0 * nd + k] : 0.f; }
#pragma unroll
    for (int i = 0; i < 3; i++) { const int k = lane + 32 * i; if (k < nd) bf.t_dof_vel[e * nd + k] = dv[i]; }
  }
}

// raw-state observation row (humanoid_smpl_im.py:653-668, obs_names :198)
__device__ __forceinline__ void store_obs_raw(float* obs, int nb, int nd, int shape_dim, int lane, bool is_body, int dof0,
                                              const float* pos, const float* rot, const float* vel, const float* angvel,
                                              const float* dq, const float* dqd, const float* motion_bodies) {
  int o = 0;
  if (is_body) {
#pragma unroll
    for (int k = 0; k < 3; k++) obs[o + lane * 3 + k] = pos[k];
  }
  o += nb * 3;
  if (is_body) {
#pragma unroll
    for (int k = 0; k < 4; k++) obs[o + lane * 4 + k] = rot[k];
  }
  o += nb * 4;
  if (is_body && dof0 >= 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) { obs[o + dof0 + k] = dq[k]; obs[o + nd + dof0 + k] = dqd[k]; }
  }
  o += 2 * nd;
  if (is_body) {
#pragma unroll
    for (int k = 0; k < 3; k++) { obs[o + lane * 3 + k] = vel[k]; obs[o + nb * 3 + lane * 3 + k] = angvel[k]; }
  }
  o += nb * 6;
  if (lane < shape_dim) obs[o + lane] = motion_bodies[lane];
}

// ---- per-env prologue (lane = body): state rows -> registers, PD targets, residual wrench, previous targets, ball
__device__ __forceinline__ void step_prologue(const b200_buffers_t& bf, const b200_motion_lib_t& ml, const b200_cfg_t& cfg,
                                              const b200_model_t& M, const LaneConst& lc, int lane, float* scr,
                                              const float* __restrict__ actions, int64_t e, Lane<float>& L, float* pdtar, float* extF,
                                              float* extT, Ball<float>& ball) {
  const int nd = M.nd, na = bf.num_actions;   // row width of the action tensors: nd, or nd + 6 with the residual root wrench
  __syncwarp();  // scratch reuse when called back to back
  // ---- load state rows (coalesced) into the warp's scratch
  const float* rs = bf.root_states + e * bf.actors_per_env * 13;
  const float* ds = bf.dof_state + e * nd * 2;
  const float* ac = actions + e * na;
  const bool was_reset = bf.reset_buf[e] == 1;
  // all loads of the env's rows are issued before the first dependent store (memory-level parallelism: one latency, not ten)
  {
    const int nbl = ml.num_lib_bodies;
    float r0 = 0.f, d[6], a[3], c0[3], c1[3], c2[3], c3[4];
    if (lane < 13) r0 = rs[lane];
#pragma unroll
    for (int i = 0; i < 6; i++) { const int k = lane + 32 * i; d[i] = k < nd * 2 ? ds[k] : 0.f; }
#pragma unroll
    for (int i = 0; i < 3; i++) { const int k = lane + 32 * i; a[i] = k < na ? ac[k] : 0.f; }
#pragma unroll
    for (int i = 0; i < 3; i++) { const int k = lane + 32 * i; c0[i] = k < nd ? bf.t_dof_pos[e * nd + k] : 0.f; c1[i] = k < nd ? bf.t_dof_vel[e * nd + k] : 0.f; }
#pragma unroll
    for (int i = 0; i < 3; i++) { const int k = lane + 32 * i; c2[i] = k < nbl * 3 ? bf.t_rb_pos[e * nbl * 3 + k] : 0.f; }
#pragma unroll
    for (int i = 0; i < 4; i++) { const int k = lane + 32 * i; c3[i] = k < nbl * 4 ? bf.t_rb_rot[e * nbl * 4 + k] : 0.f; }
    if (lane < 13) scr[lane] = r0;
#pragma unroll
    for (int i = 0; i < 6; i++) { const int k = lane + 32 * i; if (k < nd * 2) scr[16 + k] = d[i]; }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const int k = lane + 32 * i;
      if (k < na) {
        const float av = (was_reset && cfg.task_mode == 0) ? 0.0f : a[i];  // actions[self.reset_buf == 1] = 0   (:126, embodied_pose only)
        scr[16 + 2 * nd + k] = av;
        bf.actions_used[e * na + k] = av;
      }
    }
    // previous targets <- current targets (:626-636); the reward in the epilogue uses them (:677-680)
#pragma unroll
    for (int i = 0; i < 3; i++) { const int k = lane + 32 * i; if (k < nd) { bf.p_dof_pos[e * nd + k] = c0[i]; bf.p_dof_vel[e * nd + k] = c1[i]; } }
#pragma unroll
    for (int i = 0; i < 3; i++) { const int k = lane + 32 * i; if (k < nbl * 3) bf.p_rb_pos[e * nbl * 3 + k] = c2[i]; }
#pragma unroll
    for (int i = 0; i < 4; i++) { const int k = lane + 32 * i; if (k < nbl * 4) bf.p_rb_rot[e * nbl * 4 + k] = c3[i]; }
  }
  __syncwarp();

#pragma unroll
  for (int k = 0; k < 4; k++) { L.Q[k] = 0.f; L.qj[k] = 0.f; }
  L.Q[3] = 1.f; L.qj[3] = 1.f;
#pragma unroll
  for (int k = 0; k < 3; k++) { L.p[k] = 0.f; L.w[k] = 0.f; L.v[k] = 0.f; L.wt[k] = 0.f; }
  pdtar[0] = pdtar[1] = pdtar[2] = 0.f;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) { L.p[k] = scr[k]; L.v[k] = scr[7 + k]; L.w[k] = scr[10 + k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) L.Q[k] = scr[3 + k];
    qnormalize(L.Q);
  }
  if (lc.dyn && lane > 0) {
    float q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      q[k] = scr[16 + (lc.dof0 + k) * 2];
      L.wt[k] = scr[16 + (lc.dof0 + k) * 2 + 1];
      float a = scr[16 + 2 * nd + lc.dof0 + k];
      // _action_to_pd_targets: clamp(action, q -+ lim) (humanoid_smpl_im.py:391-396) or, for the vid2player player env,
      // clamp(target_dof + action, q -+ lim) (humanoid_smpl_im_mvae.py:693-709, no_scale_action / pd_target_base target_pos)
      if (cfg.pd_mode == 1) a += bf.t_dof_pos[e * nd + lc.dof0 + k];
      pdtar[k] = fmaxf(fminf(a, q[k] + cfg.pd_tar_lim), q[k] - cfg.pd_tar_lim);
      bf.pd_targets[e * nd + lc.dof0 + k] = pdtar[k];
    }
    qexp(q, L.qj);
  }
  // residual root wrench rotated by the heading of the de-based root rotation (:141-154)
  extF[0] = extF[1] = extF[2] = 0.f; extT[0] = extT[1] = extT[2] = 0.f;
  if (lane == 0 && cfg.res_force_scale > 0.f) {
    const float* rq = bf.rigid_body_state + e * bf.bodies_per_env * 13 + 3;  // self._rigid_body_rot[:, 0]
    float q0[4] = {rq[0], rq[1], rq[2], rq[3]}, qb[4], hq[4];
    ref_remove_base_rot(q0, qb);
    ref_heading_quat(ref_calc_heading(qb), hq);
    float f[3], t[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      f[k] = scr[16 + 2 * nd + nd + k] * cfg.res_force_scale;
      t[k] = scr[16 + 2 * nd + nd + 3 + k] * cfg.res_torque_scale;
    }
    ref_quat_rotate(hq, f, extF);
    ref_quat_rotate(hq, t, extT);
  }
  __syncwarp();  // the reward below reads p_* rows written by other lanes of this warp

  ball_clear(ball);
  float* ball_row = bf.root_states + (e * bf.actors_per_env + 1) * 13;  // actor 1 = ball (only dereferenced if has_ball)
  if (cfg.has_ball && lane == BALL_LANE) {
#pragma unroll
    for (int k = 0; k < 3; k++) { ball.p[k] = ball_row[k]; ball.v[k] = ball_row[7 + k]; ball.w[k] = ball_row[10 + k]; }
    ball.has_bounce = bf.has_bounce[e];
  }
}

// ball rows + flags of env e (called by the lane that carries the ball)
__device__ __forceinline__ void ball_writeback(const b200_buffers_t& bf, int64_t e, const Ball<float>& ball) {
  float* ball_row = bf.root_states + (e * bf.actors_per_env + 1) * 13;
  float* brb = bf.rigid_body_state + (e * bf.bodies_per_env + bf.bodies_per_env - 1) * 13;  // last rigid-body row = ball
#pragma unroll
  for (int k = 0; k < 3; k++) {
    ball_row[k] = ball.p[k]; ball_row[7 + k] = ball.v[k]; ball_row[10 + k] = ball.w[k];
    brb[k] = ball.p[k]; brb[7 + k] = ball.v[k]; brb[10 + k] = ball.w[k];
  }
  bf.has_bounce_now[e] = ball.bounce_now;   // cleared at the start of the step (:688), set by the aero pass (:730-734)
  if (ball.bounce_now) {
    bf.has_bounce[e] = 1;
#pragma unroll
    for (int k = 0; k < 3; k++) bf.bounce_pos[e * 3 + k] = ball.bpos[k];
  }
  bf.racket_hit_now[e] = ball.hits > 0;
}
__device__ __forceinline__ void ball_load(const b200_buffers_t& bf, int64_t e, Ball<float>& ball) {
  const float* ball_row = bf.root_states + (e * bf.actors_per_env + 1) * 13;
#pragma unroll
  for (int k = 0; k < 3; k++) { ball.p[k] = ball_row[k]; ball.v[k] = ball_row[7 + k]; ball.w[k] = ball_row[10 + k]; }
  ball.has_bounce = bf.has_bounce[e];
}

// ---- per-env epilogue, part 1 (lane = body): write back the simulation state (what gym.refresh_* exposes); dq = log(qj)
__device__ __forceinline__ void epilogue_writeback(const b200_buffers_t& bf, const b200_cfg_t& cfg, const b200_model_t& M,
                                                   const LaneConst& lc, int lane, int64_t e, const Lane<float>& L, const float* cf,
                                                   float* dq) {
  const int nd = M.nd;
  dq[0] = dq[1] = dq[2] = 0.f;
  // ---- write back the simulation state (what gym.refresh_* exposes)
  if (lane == 0) {
    float* wrs = bf.root_states + e * bf.actors_per_env * 13;
#pragma unroll
    for (int k = 0; k < 3; k++) { wrs[k] = L.p[k]; wrs[7 + k] = L.v[k]; wrs[10 + k] = L.w[k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) wrs[3 + k] = L.Q[k];
  }
  if (lc.dyn && lane > 0) {
    qlog(L.qj, dq);
    float* wds = bf.dof_state + e * nd * 2;
#pragma unroll
    for (int k = 0; k < 3; k++) { wds[(lc.dof0 + k) * 2] = dq[k]; wds[(lc.dof0 + k) * 2 + 1] = L.wt[k]; }
  }
  if (lc.active) {
    float* rb = bf.rigid_body_state + (e * bf.bodies_per_env + lane) * 13;
#pragma unroll
    for (int k = 0; k < 3; k++) { rb[k] = L.p[k]; rb[7 + k] = L.v[k]; rb[10 + k] = L.w[k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) rb[3 + k] = L.Q[k];
    float* cfo = bf.contact_forces + (e * bf.bodies_per_env + lane) * 3;
    if (cf) { cfo[0] = cf[0]; cfo[1] = cf[1]; cfo[2] = cf[2]; }
    else if (!lc.dyn) { cfo[0] = 0.f; cfo[1] = 0.f; cfo[2] = 0.f; }   // step_kernel_tmem: the owners of the dynamic bodies wrote theirs in the last substep
  }

  if (cfg.task_mode == 1 && lane == 0) bf.progress_buf[e] += 1;  // vid2player player env: post_physics_step (:785-797) only advances
}                                                                 // progress; obs / targets come from post_mvae_step, rewards / resets from the controller

// ---- per-env epilogue, part 2 (lane = body; embodied_pose task): MoCap target, obs, reward, reset.  Needs p, Q, v, w, wt of the
// lane's body and dq - either still in registers (fused kernel) or read back from the state rows (post_kernel).
__device__ __forceinline__ void epilogue_post(const b200_buffers_t& bf, const b200_motion_lib_t& ml, const b200_cfg_t& cfg,
                                              const b200_model_t& M, const LaneConst& lc, int lane, int64_t e, const Lane<float>& L,
                                              const float* dq) {
  const int nd = M.nd;
  const bool was_reset = bf.reset_buf[e] == 1;  // still the pre-step value: reset_buf is only written at the end of this function
  // ---- post-physics (:398-418)
  const int64_t progress = bf.progress_buf[e] + 1;
  const float ref_t = __fadd_rn(bf.ref_motion_times[e], __fmul_rn((float)cfg.control_freq_inv, cfg.sim_dt));
  const float step_dt = __fmul_rn((float)cfg.control_freq_inv, cfg.sim_dt);
  const int64_t mid = bf.motion_ids[e];
  if (lane == 0) { bf.progress_buf[e] = progress; bf.ref_motion_times[e] = ref_t; }
  if (ABL != 4) {
  const MotionSample ms = sample_motion(ml, M, mid, __fadd_rn(ref_t, step_dt), lane, cfg.ground_tolerance);
  store_targets(bf, cfg, ml, M, ms, e, lane, lc.dof0);
  }

  // observation
  const int nbl = ml.num_lib_bodies;
  const bool is_body = lane < nbl;
  if (ABL != 5)
  store_obs_raw(bf.obs_buf + e * bf.num_obs, nbl, nd, cfg.shape_dim, lane, is_body, lc.dof0, L.p, L.Q, L.v, L.w, dq, L.wt,
                bf.motion_bodies + e * cfg.shape_dim);

  // reward against the PREVIOUS targets (compute_humanoid_reward :918-953)
  float s_dof = 0.f, s_vel = 0.f, s_pos = 0.f, s_rot = 0.f;
  if (is_body && ABL != 3) {
    if (lc.dof0 >= 0) {
      float tq[3], tv[3];
#pragma unroll
      for (int k = 0; k < 3; k++) { tq[k] = bf.p_dof_pos[e * nd + lc.dof0 + k]; tv[k] = bf.p_dof_vel[e * nd + lc.dof0 + k]; }
      float qa[4], qb[4], oa[6], ob[6];
      ref_exp_map_to_quat(dq, qa);
      ref_exp_map_to_quat(tq, qb);
      ref_tan_norm(qa, oa);
      ref_tan_norm(qb, ob);
#pragma unroll
      for (int k = 0; k < 6; k++) { float d = oa[k] - ob[k]; s_dof += d * d; }
#pragma unroll
      for (int k = 0; k < 3; k++) { float d = tv[k] - L.wt[k]; s_vel += d * d; }
    }
    const float wgt = cfg.body_pos_weight[lane];
    float tr[4], cj[4], dqr[4];
#pragma unroll
    for (int k = 0; k < 3; k++) { float d = (bf.p_rb_pos[(e * nbl + lane) * 3 + k] - L.p[k]) * wgt; s_pos += d * d; }
#pragma unroll
    for (int k = 0; k < 4; k++) tr[k] = bf.p_rb_rot[(e * nbl + lane) * 4 + k];
    cj[0] = -L.Q[0]; cj[1] = -L.Q[1]; cj[2] = -L.Q[2]; cj[3] = L.Q[3];
    qmul(tr, cj, dqr);
    float ang, ax[3];
    ref_quat_to_angle_axis(dqr, ang, ax);
    s_rot = ang * ang;
  }
  s_dof = warp_sum(s_dof); s_vel = warp_sum(s_vel); s_pos = warp_sum(s_pos); s_rot = warp_sum(s_rot);
  // reset (compute_humanoid_reset :956-987 + caller :724-739)
  bool fall = false;
  if (is_body && cfg.enable_early_termination) fall = (L.p[2] < cfg.termination_height[lane]) && !cfg.contact_body[lane];
  const bool any_fall = __any_sync(FULL, fall);
  if (lane == 0) {
    const int njoint = nd / 3;
    float r_dof = expf(-cfg.k_dof * (s_dof / (float)(njoint * 6)));
    float r_vel = expf(-cfg.k_vel * (s_vel / (float)nd));
    float r_pos = expf(-cfg.k_pos * (s_pos / 3.0f / (float)nbl));
    float r_rot = expf(-cfg.k_rot * (s_rot / (float)nbl));
    float rew = cfg.w_dof * r_dof + cfg.w_vel * r_vel + cfg.w_pos * r_pos + cfg.w_rot * r_rot;
    if (was_reset) { rew = 0.f; r_dof = r_vel = r_pos = r_rot = 0.f; }  // :688-691
    bf.rew_buf[e] = rew;
    float* sr = bf.sub_rewards + e * 4;
    sr[0] = r_dof; sr[1] = r_vel; sr[2] = r_pos; sr[3] = r_rot;
    int64_t terminated = (cfg.enable_early_termination && any_fall && progress > 1) ? 1 : 0;
    const bool cond = (progress >= (int64_t)cfg.max_episode_length - 1) || (ref_t >= ml.motion_lengths[mid]);
    int64_t reset = cond ? 1 : terminated;
    if (was_reset) { reset = 1; terminated = bf.terminate_buf[e]; }
    bf.reset_buf[e] = reset;
    bf.terminate_buf[e] = terminated;
  }
}

// ---- per-env epilogue of the fused kernels
__device__ __forceinline__ void step_epilogue(const b200_buffers_t& bf, const b200_motion_lib_t& ml, const b200_cfg_t& cfg,
                                              const b200_model_t& M, const LaneConst& lc, int lane, int64_t e, const Lane<float>& L,
                                              const float* cf, const Ball<float>& ball, bool ball_in_lane31 = true) {
  if (cfg.has_ball && lane == BALL_LANE && ball_in_lane31) ball_writeback(bf, e, ball);
  float dq[3];
  epilogue_writeback(bf, cfg, M, lc, lane, e, L, cf, dq);
  if (cfg.task_mode == 1 || ABL == 2) return;
  epilogue_post(bf, ml, cfg, M, lc, lane, e, L, dq);
}

// ------------------------------------------------------------------------------------------
// fused env step:  pre-physics -> substeps -> MoCap target -> obs -> reward -> reset
// Persistent: the grid is sized to the resident CTA slots; each warp pulls env indices from a global ticket
// counter (monotonic across launches: ticket - base = env index) so the tail is balanced per env, not per CTA.
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, STEP_MIN_CTAS)
step_kernel(const DevBlob* __restrict__ gblob, uint32_t blob_bytes, const b200_cfg_t* __restrict__ gcfg, b200_buffers_t bf,
            b200_motion_lib_t ml, const float* __restrict__ actions, int num_envs, unsigned long long* __restrict__ ticket,
            int env_first, int env_stride) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar;
  load_blob(smem, gblob, blob_bytes, &mbar);
  const DevBlob& B = *reinterpret_cast<const DevBlob*>(smem);
  const b200_model_t& M = B.m;
  const float* verts = reinterpret_cast<const float*>(smem + sizeof(DevBlob));
  float* scratch_all = reinterpret_cast<float*>(smem + ((blob_bytes + 15) & ~15u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* scr = scratch_all + warp * SCRATCH_FLOATS;
  const b200_cfg_t& cfg = *gcfg;
  const LaneConst lc = lane_const(M, lane);
  const int nd = M.nd;
  const PhysCfg<float> pc = make_phys_cfg<float>(cfg);

 __shared__ unsigned long long s_tk;
 for (;;) {
  // one ticket per CTA = a batch of WARPS_PER_CTA consecutive envs (CTA-uniform control flow -> barriers are legal)
  __syncthreads();
  if (threadIdx.x == 0) s_tk = atomicAdd(ticket, (unsigned long long)WARPS_PER_CTA);
  __syncthreads();
  const int64_t e0 = (int64_t)s_tk;
  if (e0 >= num_envs) break;
  const bool full_batch = e0 + WARPS_PER_CTA <= num_envs;
  if (!full_batch && e0 + warp >= num_envs) continue;  // ragged last batch: no CTA barriers below (full_batch is CTA-uniform)
  const int64_t e = env_first + (int64_t)env_stride * (e0 + warp);  // row of this env in the bound tensors (env slice)

  Lane<float> L;
  float pdtar[3], extF[3], extT[3], cf[3];
  Ball<float> ball;
  step_prologue(bf, ml, cfg, M, lc, lane, scr, actions, e, L, pdtar, extF, extT, ball);
  control_step<float>(B, verts, pc, lc, lane, L, pdtar, extF, extT, cf, ball, STEP_SYNC && full_batch);
#if POST_SYNC
  if (full_batch) __syncthreads();  // re-align the CTA's warps: the once-per-step epilogue is straight-line code, fetched once per CTA
#endif
  step_epilogue(bf, ml, cfg, M, lc, lane, e, L, cf, ball);
 }  // ticket loop
}


// ------------------------------------------------------------------------------------------
// three-launch form of the env step (b200env_step with h->split): pre_kernel -> step_kernel_packed<true> -> post_kernel.
// state of env row e as the physics wants it, from the rows the pre_kernel left behind (lane = body)
__device__ __forceinline__ void load_state_rows(const b200_buffers_t& bf, const b200_model_t& M, const LaneConst& lc, int lane, int64_t e,
                                                const float* __restrict__ ext_wrench, Lane<float>& L, float* pdtar, float* extF, float* extT) {
  const int nd = M.nd;
#pragma unroll
  for (int k = 0; k < 4; k++) { L.Q[k] = 0.f; L.qj[k] = 0.f; }
  L.Q[3] = 1.f; L.qj[3] = 1.f;
#pragma unroll
  for (int k = 0; k < 3; k++) { L.p[k] = 0.f; L.w[k] = 0.f; L.v[k] = 0.f; L.wt[k] = 0.f; pdtar[k] = 0.f; extF[k] = 0.f; extT[k] = 0.f; }
  if (lane == 0) {
    const float* rs = bf.root_states + e * bf.actors_per_env * 13;
#pragma unroll
    for (int k = 0; k < 3; k++) { L.p[k] = rs[k]; L.v[k] = rs[7 + k]; L.w[k] = rs[10 + k]; extF[k] = ext_wrench[e * 6 + k]; extT[k] = ext_wrench[e * 6 + 3 + k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) L.Q[k] = rs[3 + k];
    qnormalize(L.Q);
  }
  if (lc.dyn && lane > 0) {
    const float* ds = bf.dof_state + (e * nd + lc.dof0) * 2;
    float q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { q[k] = ds[2 * k]; L.wt[k] = ds[2 * k + 1]; pdtar[k] = bf.pd_targets[e * nd + lc.dof0 + k]; }
    qexp(q, L.qj);
  }
}

// launch 1: pre_physics_step of every env (zeroed actions of reset envs, PD targets, residual root wrench, previous <- current
// targets).  Warp per env, lane per body, 8 warps per CTA, no big shared-memory footprint -> full occupancy.
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
pre_kernel(const DevBlob* __restrict__ gblob, const b200_cfg_t* __restrict__ gcfg, b200_buffers_t bf, b200_motion_lib_t ml,
           const float* __restrict__ actions, int num_envs, int env_first, int env_stride, float* __restrict__ ext_wrench,
           unsigned long long* __restrict__ ticket) {
  __shared__ float s_scr[WARPS_PER_CTA * SCRATCH_FLOATS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) *ticket = 0ull;   // hand-out counter of the physics launch that follows (one memset node less per step)
  const int64_t i = (int64_t)blockIdx.x * WARPS_PER_CTA + warp;
  if (i >= num_envs) return;
  const int64_t e = env_first + (int64_t)env_stride * i;
  const b200_model_t& M = gblob->m;
  const LaneConst lc = lane_const(M, lane);
  Lane<float> L;
  float pdtar[3], extF[3], extT[3];
  Ball<float> dummy;
  step_prologue(bf, ml, *gcfg, M, lc, lane, s_scr + warp * SCRATCH_FLOATS, actions, e, L, pdtar, extF, extT, dummy);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) { ext_wrench[e * 6 + k] = extF[k]; ext_wrench[e * 6 + 3 + k] = extT[k]; }
  }
}

// launch 3: post_physics_step of every env (MoCap target, obs, reward, reset) from the state rows launch 2 wrote
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
post_kernel(const DevBlob* __restrict__ gblob, const b200_cfg_t* __restrict__ gcfg, b200_buffers_t bf, b200_motion_lib_t ml, int num_envs,
            int env_first, int env_stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * WARPS_PER_CTA + warp;
  if (i >= num_envs) return;
  const int64_t e = env_first + (int64_t)env_stride * i;
  const b200_model_t& M = gblob->m;
  const LaneConst lc = lane_const(M, lane);
  const int nd = M.nd;
  Lane<float> L;
#pragma unroll
  for (int k = 0; k < 4; k++) { L.Q[k] = 0.f; L.qj[k] = 0.f; }
  L.Q[3] = 1.f; L.qj[3] = 1.f;
  float dq[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 3; k++) { L.p[k] = 0.f; L.w[k] = 0.f; L.v[k] = 0.f; L.wt[k] = 0.f; }
  if (lc.active) {
    const float* rb = bf.rigid_body_state + (e * bf.bodies_per_env + lane) * 13;
#pragma unroll
    for (int k = 0; k < 3; k++) { L.p[k] = rb[k]; L.v[k] = rb[7 + k]; L.w[k] = rb[10 + k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) L.Q[k] = rb[3 + k];
  }
  if (lc.dyn && lane > 0) {
    const float* ds = bf.dof_state + (e * nd + lc.dof0) * 2;
#pragma unroll
    for (int k = 0; k < 3; k++) { dq[k] = ds[2 * k]; L.wt[k] = ds[2 * k + 1]; }
  }
  epilogue_post(bf, ml, *gcfg, M, lc, lane, e, L, dq);
}

// ------------------------------------------------------------------------------------------
// packed variant of the fused step: 4 envs per warp in the physics (packed.cuh), lane-per-body prologue / epilogue per env
#ifndef PK_WARPS
#define PK_WARPS 7          // 7 warps x 4 envs x 7.2 KB records + 20 KB constants = 227 KB of shared memory: one CTA per SM
#endif
#define PK_SCRATCH 232
// PK_SHADOW=1 (experiment, tools/shadow.sh): launch 2 x PK_WARPS warps; warps PK_WARPS.. are "shadows" that run the instruction stream
// of warp - PK_WARPS on the same records with every store suppressed: what would 14 resident warps cost (issue slots, shared-memory
// bandwidth, 128 registers)?  Results of the real warps are unchanged.
#ifndef PK_SHADOW
#define PK_SHADOW 0
#endif
#define PK_LAUNCH_WARPS (PK_WARPS * (PK_SHADOW ? 2 : 1))
#ifndef PK_BOUND_WARPS
#define PK_BOUND_WARPS PK_LAUNCH_WARPS    // A/B: a larger bound = the register budget of that many warps at the product's warp count
#endif

// SPLIT = false: the whole env step in this launch.  SPLIT = true: the middle launch of pre_kernel -> this -> post_kernel; here only
// the state rows / PD targets / residual wrench are read and the state rows written (the once-per-step task logic runs in the two
// high-occupancy kernels, where its memory latency is covered by 48+ warps per SM instead of the 7 this kernel can hold).
template <bool SPLIT>
__global__ void __launch_bounds__(PK_BOUND_WARPS * 32, 1)
step_kernel_packed(const DevBlob* __restrict__ gblob, uint32_t blob_bytes, const b200_cfg_t* __restrict__ gcfg, b200_buffers_t bf,
                   b200_motion_lib_t ml, const float* __restrict__ actions, int num_envs, unsigned long long* __restrict__ ticket,
                   int env_first, int env_stride, const float* __restrict__ ext_wrench, const int32_t* __restrict__ perm) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar;
  load_blob(smem, gblob, blob_bytes, &mbar);
  const DevBlob& B = *reinterpret_cast<const DevBlob*>(smem);
  const b200_model_t& M = B.m;
  const float* verts = reinterpret_cast<const float*>(smem + sizeof(DevBlob));
  float* scratch_all = reinterpret_cast<float*>(smem + ((blob_bytes + 15) & ~15u));
  const int warp = (threadIdx.x >> 5) % PK_WARPS, lane = threadIdx.x & 31;
  const bool shadow = PK_SHADOW && (int)threadIdx.x >= PK_WARPS * 32;
  float* scr = scratch_all + warp * PK_SCRATCH;   // prologue / epilogue scratch of the fused form; the split form has none
  float* wrec = scratch_all + (SPLIT ? 0 : PK_WARPS * PK_SCRATCH) + (size_t)warp * EPW * ENV_STRIDE;
  __shared__ b200_cfg_t s_cfg;  // constants in shared memory: no global (long-scoreboard) reloads inside the substep loop
  for (int k = threadIdx.x; k < (int)(sizeof(b200_cfg_t) / 4); k += blockDim.x) reinterpret_cast<uint32_t*>(&s_cfg)[k] = reinterpret_cast<const uint32_t*>(gcfg)[k];
  __syncthreads();
  const b200_cfg_t& cfg = s_cfg;
  LaneConst lc = lane_const(M, lane);
  if (lc.active) lc.rix = B.t.rix[lane];
  const PhysCfg<float> pc = make_phys_cfg<float>(cfg);
  const int g = lane >> 3, s = lane & 7;
  // hand-out order of the envs: identity, or (B200ENV_SORT=1) the permutation sort_perm_kernel built from the contact load, so that
  // the envs of a warp / of a CTA batch have similar ground-contact work (tools/contact_imbalance.py)
  auto row = [&](int64_t i) -> int64_t { return env_first + (int64_t)env_stride * (perm ? (int64_t)perm[i] : i); };
  // The CTA's warps form PK_GROUPS independent groups, each pulling its own batches and meeting at its own named barrier:
  // while one group is in the (memory-latency-bound) prologue / epilogue the other is in the (issue-bound) physics.
#ifndef PK_GROUPS
#define PK_GROUPS 1   // 2 measured 1.5 % slower (343 vs 338 us): mixing the phases costs more I-cache than it hides latency
#endif
  constexpr int G0 = PK_GROUPS == 1 ? PK_WARPS : (PK_WARPS + 1) / 2;       // warps in group 0
  const int grp = warp < G0 ? 0 : 1;
  const int gw0 = grp == 0 ? 0 : G0, gwn = grp == 0 ? G0 : PK_WARPS - G0;     // first warp / number of warps of my group
  const int BATCH = gwn * EPW;
  const int gthreads = gwn * 32 * (PK_SHADOW ? 2 : 1);
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(gthreads) : "memory"); };

  __shared__ unsigned long long s_tk[2];
  for (;;) {
#if PK_WARP_TICKETS
    // one ticket per WARP = EPW consecutive envs: no CTA barrier anywhere in the loop, a warp whose envs were cheap (no
    // ground contact) simply comes back for the next group earlier
    unsigned long long tk = 0;
    if (lane == 0) tk = atomicAdd(ticket, (unsigned long long)EPW);
    tk = __shfl_sync(FULL, tk, 0);
    const int64_t eb = (int64_t)tk;
    if (eb >= num_envs) break;
    const bool full_batch = false;
    (void)s_tk; (void)BATCH; (void)group_sync;
#else
    group_sync();
    if ((int)threadIdx.x == gw0 * 32) s_tk[grp] = atomicAdd(ticket, (unsigned long long)BATCH);
    group_sync();
    const int64_t e0 = (int64_t)s_tk[grp];
    if (e0 >= num_envs) break;
    const bool full_batch = e0 + BATCH <= num_envs;
    const int64_t eb = e0 + (int64_t)(warp - gw0) * EPW;
    if (!full_batch && eb >= num_envs) continue;
#endif

    if (SPLIT) {
      // all loads of the warp's EPW envs are issued before the first use (one memory latency instead of EPW)
      float rq[EPW][3], rw[EPW][3], rp[EPW][3], r1[EPW];   // r1: lane j < 13 holds root_states[e][j], 13 <= j < 19 the residual wrench
      const int nd = M.nd;
#pragma unroll
      for (int k = 0; k < EPW; k++) {
        const int64_t ek = eb + k < num_envs ? eb + k : (int64_t)num_envs - 1;   // ragged tail: re-read the last env, never stored
        const int64_t e = row(ek);
        if (lc.dyn && lane > 0) {
          const float* ds = bf.dof_state + (e * nd + lc.dof0) * 2;
#pragma unroll
          for (int j = 0; j < 3; j++) { rq[k][j] = ds[2 * j]; rw[k][j] = ds[2 * j + 1]; rp[k][j] = bf.pd_targets[e * nd + lc.dof0 + j]; }
        }
        r1[k] = lane < 13 ? bf.root_states[e * bf.actors_per_env * 13 + lane] : (lane < 19 ? ext_wrench[e * 6 + lane - 13] : 0.f);
      }
#pragma unroll
      for (int k = 0; k < EPW; k++) {
        if (eb + k >= num_envs) break;
        Lane<float> L;
        float pdtar[3] = {0.f, 0.f, 0.f}, extF[3] = {0.f, 0.f, 0.f}, extT[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; j++) { L.Q[j] = 0.f; L.qj[j] = 0.f; }
        L.Q[3] = 1.f; L.qj[3] = 1.f;
#pragma unroll
        for (int j = 0; j < 3; j++) { L.p[j] = 0.f; L.w[j] = 0.f; L.v[j] = 0.f; L.wt[j] = 0.f; }
        float rr[19];
#pragma unroll
        for (int j = 0; j < 19; j++) rr[j] = __shfl_sync(FULL, r1[k], j);
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 3; j++) { L.p[j] = rr[j]; L.v[j] = rr[7 + j]; L.w[j] = rr[10 + j]; extF[j] = rr[13 + j]; extT[j] = rr[16 + j]; }
#pragma unroll
          for (int j = 0; j < 4; j++) L.Q[j] = rr[3 + j];
          qnormalize(L.Q);
        }
        if (lc.dyn && lane > 0) {
#pragma unroll
          for (int j = 0; j < 3; j++) { L.wt[j] = rw[k][j]; pdtar[j] = rp[k][j]; }
          qexp(rq[k], L.qj);
        }
        pk_store_state<float>(wrec + k * ENV_STRIDE, lc, lane, L, pdtar, extF, extT);
      }
    } else {
    for (int k = 0; k < EPW; k++) {
      if (eb + k >= num_envs) break;
      const int64_t e = env_first + (int64_t)env_stride * (eb + k);  // row in the bound tensors (env slice)
      Lane<float> L;
      float pdtar[3], extF[3], extT[3];
      Ball<float> dummy;
      step_prologue(bf, ml, cfg, M, lc, lane, scr, actions, e, L, pdtar, extF, extT, dummy);
      pk_store_state<float>(wrec + k * ENV_STRIDE, lc, lane, L, pdtar, extF, extT);
    }
    }
    __syncwarp();
    const bool valid = eb + g < num_envs;
    Ball<float> ball;
    ball_clear(ball);
    const int64_t erow_g = SPLIT ? row(valid ? eb + g : eb) : env_first + (int64_t)env_stride * (eb + g);
    if (cfg.has_ball && valid && s == BALL_SLOT) ball_load(bf, erow_g, ball);
    if (ABL != 1) control_step_packed<float>(B, verts, pc, wrec, lane, valid, ball, STEP_SYNC && full_batch);
#if POST_SYNC
    // fused form only: re-aligns the warps before the long straight-line epilogue (I-cache).  The split form's epilogue is a
    // state write-back; without the barrier it measured 295.3 vs 298.3 us per 8192-env step (profiles/r1h_ab.log).
    if (!SPLIT && full_batch) group_sync();
#endif
    if (cfg.has_ball && valid && s == BALL_SLOT && !shadow) ball_writeback(bf, erow_g, ball);
    for (int k = 0; k < EPW; k++) {
      if (eb + k >= num_envs || shadow) break;
      const int64_t e = SPLIT ? row(eb + k) : env_first + (int64_t)env_stride * (eb + k);
      Lane<float> L;
      float cf[3];
      pk_load_state<float>(wrec + k * ENV_STRIDE, lc, lane, L, cf);
      if (SPLIT) {
        float dq[3];
        epilogue_writeback(bf, cfg, M, lc, lane, e, L, cf, dq);
      } else {
        step_epilogue(bf, ml, cfg, M, lc, lane, e, L, cf, ball, false);
      }
    }
    __syncwarp();
  }
}


// ------------------------------------------------------------------------------------------
// One-wave form of the middle launch (pre_kernel -> this -> post_kernel): csrc/packed_t.cuh.  14 warps x 4 envs = 56 envs per CTA
// (3.5 KB of shared memory per env + the constant block = 220 KB), the lane-private fields of the bodies in tensor memory: warp w owns
// the 32 TMEM lanes of quadrant w % 4 and the 128 columns starting at 128 * (w / 4).  8192 envs = 147 CTAs: ONE round on 148 SMs.
#ifndef PT_WARPS
#define PT_WARPS 14
#endif
__device__ __forceinline__ uint32_t pt_tmem_alloc(uint32_t* holder, int warp) {
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(holder)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *holder;
}
__device__ __forceinline__ void pt_tmem_free(uint32_t base, int warp) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}

__global__ void __launch_bounds__(PT_WARPS * 32, 1)
step_kernel_tmem(const DevBlob* __restrict__ gblob, uint32_t blob_bytes, const b200_cfg_t* __restrict__ gcfg, b200_buffers_t bf, int num_envs,
                 unsigned long long* __restrict__ ticket, int env_first, int env_stride, const float* __restrict__ ext_wrench) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t s_tmem;
  load_blob(smem, gblob, blob_bytes, &mbar);
  const DevBlob& B = *reinterpret_cast<const DevBlob*>(smem);
  const b200_model_t& M = B.m;
  const float* verts = reinterpret_cast<const float*>(smem + sizeof(DevBlob));
  const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // the shuffle tells the compiler that warp is warp-uniform: the
                                                                                          // window addresses of the private store stay in uniform registers
  float* wrec = reinterpret_cast<float*>(smem + ((blob_bytes + 15) & ~15u)) + (size_t)warp * EPW * PT_ENV_STRIDE;
  __shared__ b200_cfg_t s_cfg;
  for (int k = threadIdx.x; k < (int)(sizeof(b200_cfg_t) / 4); k += blockDim.x) reinterpret_cast<uint32_t*>(&s_cfg)[k] = reinterpret_cast<const uint32_t*>(gcfg)[k];
  const uint32_t tbase = pt_tmem_alloc(&s_tmem, warp);   // has a __syncthreads: s_cfg is complete behind it
  const PrivTmem ps{tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * PT_WARP_COLS};
  const b200_cfg_t& cfg = s_cfg;
  LaneConst lc = lane_const(M, lane);
  if (lc.active) lc.rix = B.t.rix[lane];
  __shared__ PhysCfg<float> s_pc;   // derived constants in shared memory, not in ~50 registers per thread (128 registers at 14 warps)
  if (threadIdx.x == 0) s_pc = make_phys_cfg<float>(cfg);
  __syncthreads();
  const PhysCfg<float>& pc = s_pc;
  const int g = lane >> 3, s = lane & 7;
  auto row = [&](int64_t i) -> int64_t { return env_first + (int64_t)env_stride * i; };
  constexpr int BATCH = PT_WARPS * EPW;
  __shared__ unsigned long long s_tk;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_tk = atomicAdd(ticket, (unsigned long long)BATCH);
    __syncthreads();
    const int64_t e0 = (int64_t)s_tk;
    if (e0 >= num_envs) break;
    const bool full_batch = e0 + BATCH <= num_envs;
    const int64_t eb = e0 + (int64_t)warp * EPW;
    if (!full_batch && eb >= num_envs) continue;
    {
      // all loads of the warp's EPW envs are issued before the first use (one memory latency instead of EPW)
      float rq[EPW][3], rw[EPW][3], rp[EPW][3], r1[EPW];   // r1: lane j < 13 holds root_states[e][j], 13 <= j < 19 the residual wrench
      const int nd = M.nd;
#pragma unroll
      for (int k = 0; k < EPW; k++) {
        const int64_t ek = eb + k < num_envs ? eb + k : (int64_t)num_envs - 1;   // ragged tail: re-read the last env, never stored
        const int64_t e = row(ek);
        if (lc.dyn && lane > 0) {
          const float* ds = bf.dof_state + (e * nd + lc.dof0) * 2;
#pragma unroll
          for (int j = 0; j < 3; j++) { rq[k][j] = ds[2 * j]; rw[k][j] = ds[2 * j + 1]; rp[k][j] = bf.pd_targets[e * nd + lc.dof0 + j]; }
        }
        r1[k] = lane < 13 ? bf.root_states[e * bf.actors_per_env * 13 + lane] : (lane < 19 ? ext_wrench[e * 6 + lane - 13] : 0.f);
      }
#pragma unroll
      for (int k = 0; k < EPW; k++) {
        if (eb + k >= num_envs) break;
        Lane<float> L;
        float pdtar[3] = {0.f, 0.f, 0.f}, extF[3] = {0.f, 0.f, 0.f}, extT[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; j++) { L.Q[j] = 0.f; L.qj[j] = 0.f; }
        L.Q[3] = 1.f; L.qj[3] = 1.f;
#pragma unroll
        for (int j = 0; j < 3; j++) { L.p[j] = 0.f; L.w[j] = 0.f; L.v[j] = 0.f; L.wt[j] = 0.f; }
        float rr[19];
#pragma unroll
        for (int j = 0; j < 19; j++) rr[j] = __shfl_sync(FULL, r1[k], j);
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 3; j++) { L.p[j] = rr[j]; L.v[j] = rr[7 + j]; L.w[j] = rr[10 + j]; extF[j] = rr[13 + j]; extT[j] = rr[16 + j]; }
#pragma unroll
          for (int j = 0; j < 4; j++) L.Q[j] = rr[3 + j];
          qnormalize(L.Q);
        }
        if (lc.dyn && lane > 0) {
#pragma unroll
          for (int j = 0; j < 3; j++) { L.wt[j] = rw[k][j]; pdtar[j] = rp[k][j]; }
          qexp(rq[k], L.qj);
        }
        pt_stage_in<float>(wrec + k * PT_ENV_STRIDE, lc, lane, L, pdtar, extF, extT);
      }
    }
    __syncwarp();
    const bool valid = eb + g < num_envs;
    pt_adopt<float>(B, wrec, lane, valid, ps);
    Ball<float> ball;
    ball_clear(ball);
    const int64_t erow_g = row(valid ? eb + g : eb);
    if (cfg.has_ball && valid && s == BALL_SLOT) ball_load(bf, erow_g, ball);
    float* cf_env = bf.contact_forces + erow_g * bf.bodies_per_env * 3;
    control_step_t<float>(B, verts, pc, wrec, lane, valid, ball, ps, cf_env, PT_STEP_SYNC && full_batch);
    if (cfg.has_ball && valid && s == BALL_SLOT) ball_writeback(bf, erow_g, ball);
    pt_publish<float>(B, wrec, lane, valid, ps);
    __syncwarp();
    for (int k = 0; k < EPW; k++) {
      if (eb + k >= num_envs) break;
      const int64_t e = row(eb + k);
      Lane<float> L;
      float dq[3];
      pt_load_state<float>(wrec + k * PT_ENV_STRIDE, lc, lane, L);
      epilogue_writeback(bf, cfg, M, lc, lane, e, L, nullptr, dq);
    }
    __syncwarp();
  }
  pt_tmem_free(tbase, warp);
}

// the test entry of the same device code (b200env_physics_only with the handle in tmem mode): float only - the private store of the
// double instantiation does not fit the tensor memory; float64 parity of this form is checked on the CPU lane emulator (tests/emu)
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
physics_kernel_tmem(const DevBlob* __restrict__ gblob, uint32_t blob_bytes, const b200_cfg_t* __restrict__ gcfg, int n, int n_steps,
                    float* root, float* dof_pos, float* dof_vel, const float* pd_tar, const float* ext, float* rb_out, float* contact_out,
                    float* ballio, int32_t* hits) {
  typedef float T;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t s_tmem;
  load_blob(smem, gblob, blob_bytes, &mbar);
  const DevBlob& B = *reinterpret_cast<const DevBlob*>(smem);
  const b200_model_t& M = B.m;
  const float* verts = reinterpret_cast<const float*>(smem + sizeof(DevBlob));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  T* wrec = reinterpret_cast<T*>(smem + ((blob_bytes + 15) & ~15u)) + (size_t)warp * EPW * PT_ENV_STRIDE;
  const uint32_t tbase = pt_tmem_alloc(&s_tmem, warp);
  const PrivTmem ps{tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * PT_WARP_COLS};
  const int64_t eb = ((int64_t)blockIdx.x * WARPS + warp) * EPW;
  if (eb < n) {
    LaneConst lc = lane_const(M, lane);
    if (lc.active) lc.rix = B.t.rix[lane];
    const int nb = M.nb, nd = M.nd;
    PhysCfg<T> pc = make_phys_cfg<T>(*gcfg);
    const bool with_ball = pc.has_ball && ballio != nullptr;
    pc.has_ball = with_ball;
    const int g = lane >> 3, s = lane & 7;
    for (int k = 0; k < EPW; k++) {
      const int64_t e = eb + k;
      if (e >= n) break;
      Lane<T> L;
      for (int j = 0; j < 4; j++) { L.Q[j] = 0; L.qj[j] = 0; }
      L.Q[3] = 1; L.qj[3] = 1;
      for (int j = 0; j < 3; j++) { L.p[j] = 0; L.w[j] = 0; L.v[j] = 0; L.wt[j] = 0; }
      T pdt[3] = {0, 0, 0}, eF[3] = {0, 0, 0}, eT[3] = {0, 0, 0};
      if (lane == 0) {
        const T* rs = root + e * 13;
        for (int j = 0; j < 3; j++) { L.p[j] = rs[j]; L.v[j] = rs[7 + j]; L.w[j] = rs[10 + j]; }
        for (int j = 0; j < 4; j++) L.Q[j] = rs[3 + j];
        qnormalize(L.Q);
        if (ext) for (int j = 0; j < 3; j++) { eF[j] = ext[e * 6 + j]; eT[j] = ext[e * 6 + 3 + j]; }
      }
      if (lc.dyn && lane > 0) {
        T q[3];
        for (int j = 0; j < 3; j++) { q[j] = dof_pos[e * nd + lc.dof0 + j]; L.wt[j] = dof_vel[e * nd + lc.dof0 + j]; pdt[j] = pd_tar[e * nd + lc.dof0 + j]; }
        qexp(q, L.qj);
      }
      pt_stage_in<T>(wrec + k * PT_ENV_STRIDE, lc, lane, L, pdt, eF, eT);
    }
    __syncwarp();
    const bool valid = eb + g < n;
    pt_adopt<T>(B, wrec, lane, valid, ps);
    Ball<T> ball;
    ball_clear(ball);
    if (with_ball && valid && s == BALL_SLOT) {
      const int64_t e = eb + g;
      for (int j = 0; j < 3; j++) { ball.p[j] = ballio[e * 13 + j]; ball.v[j] = ballio[e * 13 + 7 + j]; ball.w[j] = ballio[e * 13 + 10 + j]; }
    }
    T* cf_env = (valid && contact_out) ? contact_out + (eb + g) * nb * 3 : nullptr;
    for (int st_ = 0; st_ < n_steps; st_++) {
      control_step_t<T>(B, verts, pc, wrec, lane, valid, ball, ps, cf_env, false);
      if (st_ + 1 < n_steps) pt_requantize<T>(B, lane, valid, ps);   // the state crosses control steps as exp-map coordinates
    }
    pt_publish<T>(B, wrec, lane, valid, ps);
    __syncwarp();
    for (int k = 0; k < EPW; k++) {
      const int64_t e = eb + k;
      if (e >= n) break;
      Lane<T> L;
      pt_load_state<T>(wrec + k * PT_ENV_STRIDE, lc, lane, L);
      if (lane == 0) {
        T* rs = root + e * 13;
        for (int j = 0; j < 3; j++) { rs[j] = L.p[j]; rs[7 + j] = L.v[j]; rs[10 + j] = L.w[j]; }
        for (int j = 0; j < 4; j++) rs[3 + j] = L.Q[j];
      }
      if (lc.dyn && lane > 0) {
        T q[3];
        qlog(L.qj, q);
        for (int j = 0; j < 3; j++) { dof_pos[e * nd + lc.dof0 + j] = q[j]; dof_vel[e * nd + lc.dof0 + j] = L.wt[j]; }
      }
      if (lc.active) {
        T* rb = rb_out + (e * nb + lane) * 13;
        for (int j = 0; j < 3; j++) { rb[j] = L.p[j]; rb[7 + j] = L.v[j]; rb[10 + j] = L.w[j]; }
        for (int j = 0; j < 4; j++) rb[3 + j] = L.Q[j];
        if (contact_out && !lc.dyn) for (int j = 0; j < 3; j++) contact_out[(e * nb + lane) * 3 + j] = T(0);
      }
    }
    if (with_ball && valid && s == BALL_SLOT) {
      const int64_t e = eb + g;
      for (int j = 0; j < 3; j++) { ballio[e * 13 + j] = ball.p[j]; ballio[e * 13 + 7 + j] = ball.v[j]; ballio[e * 13 + 10 + j] = ball.w[j]; }
      if (hits) hits[e] = ball.hits;
    }
  }
  pt_tmem_free(tbase, warp);
}

// ------------------------------------------------------------------------------------------
// hand-out order by contact load (B200ENV_SORT=1, split form).  The per-vertex loop of contact_hull runs on the body's lane; a warp
// executes, per body round, as many iterations as its slowest lane, and a CTA batch waits for its slowest warp: with fallen humanoids
// handed out in env order the slowest warp of a batch has ~75 % more vertex iterations than the average one, sorted by load ~10 %
// (tools/contact_imbalance.py).  sort_count_kernel: serial vertex iterations of every env from the rigid-body rows (the state the
// step starts from) -> 16 bins, position inside the bin by atomicAdd (the order inside a bin does not matter: envs are independent,
// every env's result is bit-identical whatever the order).  sort_perm_kernel: heaviest bin first.
#define SORT_BINS 16
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
sort_count_kernel(const DevBlob* __restrict__ gblob, b200_buffers_t bf, int num_envs, int env_first, int env_stride, int32_t* __restrict__ bin,
                  uint32_t* __restrict__ pos, uint32_t* __restrict__ cnt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * WARPS_PER_CTA + warp;
  if (i >= num_envs) return;
  const int64_t e = env_first + (int64_t)env_stride * i;
  const b200_model_t& M = gblob->m;
  const float* verts = reinterpret_cast<const float*>(gblob + 1);   // SoA [nb][3][vmax] behind the header
  int c = 0;
  if (lane < M.nb && M.nverts[lane] > 0 && !M.fixed[lane]) {
    const float* rb = bf.rigid_body_state + (e * bf.bodies_per_env + lane) * 13;
    const float pz = rb[2];
    if (pz - M.radius[lane] < 0.f) {
      const float x = rb[3], y = rb[4], z = rb[5], w = rb[6];
      const float r6 = 2.f * (x * z - y * w), r7 = 2.f * (y * z + x * w), r8 = 1.f - 2.f * (x * x + y * y);   // z row of the rotation
      const float* vb = verts + (size_t)lane * M.vmax * 3;
      for (int k = 0; k < M.nverts[lane]; k++) c += (pz + r6 * vb[k] + r7 * vb[M.vmax + k] + r8 * vb[2 * M.vmax + k]) < 0.f;
    }
  }
  // lanes 8r .. 8r+7 = the bodies of round r of the packed body pass: max per round, then the sum of the rounds
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) c = max(c, __shfl_xor_sync(FULL, c, o));
  const int serial = __shfl_sync(FULL, c, 0) + __shfl_sync(FULL, c, 8) + __shfl_sync(FULL, c, 16) + __shfl_sync(FULL, c, 24);
  if (lane == 0) {
    const int b = min(serial / 3, SORT_BINS - 1);
    bin[i] = b;
    pos[i] = atomicAdd(&cnt[b], 1u);
  }
}
__global__ void sort_perm_kernel(int num_envs, const int32_t* __restrict__ bin, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ cnt,
                                 int32_t* __restrict__ perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_envs) return;
  uint32_t off = 0;
  for (int b = SORT_BINS - 1; b > bin[i]; b--) off += cnt[b];   // heaviest bin first
  perm[off + pos[i]] = i;
}

// ------------------------------------------------------------------------------------------
#if B200ENV_WITH_PACKED3
// packed3 form of the middle launch (pre_kernel -> this -> post_kernel): 2 envs per warp, 14 warps per CTA (csrc/packed3.cuh).
// Same shared-memory footprint and the same 28 envs per CTA batch as step_kernel_packed<split>; twice the warps.
#ifndef PK3_WARPS
#define PK3_WARPS 14
#endif
__global__ void __launch_bounds__(PK3_WARPS * 32, 1)
step_kernel_packed3(const DevBlob* __restrict__ gblob, uint32_t blob_bytes, const b200_cfg_t* __restrict__ gcfg, b200_buffers_t bf,
                    int num_envs, unsigned long long* __restrict__ ticket, int env_first, int env_stride,
                    const float* __restrict__ ext_wrench) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar;
  load_blob(smem, gblob, blob_bytes, &mbar);
  const DevBlob& B = *reinterpret_cast<const DevBlob*>(smem);
  const b200_model_t& M = B.m;
  const float* verts = reinterpret_cast<const float*>(smem + sizeof(DevBlob));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wrec = reinterpret_cast<float*>(smem + ((blob_bytes + 15) & ~15u)) + (size_t)warp * EPW3 * ENV_STRIDE;
  __shared__ b200_cfg_t s_cfg;
  for (int k = threadIdx.x; k < (int)(sizeof(b200_cfg_t) / 4); k += blockDim.x) reinterpret_cast<uint32_t*>(&s_cfg)[k] = reinterpret_cast<const uint32_t*>(gcfg)[k];
  __syncthreads();
  const b200_cfg_t& cfg = s_cfg;
  LaneConst lc = lane_const(M, lane);
  if (lc.active) lc.rix = B.t.rix[lane];
  const PhysCfg<float> pc = make_phys_cfg<float>(cfg);
  const int g = lane >> 4, u = lane & 15;
  constexpr int BATCH = PK3_WARPS * EPW3;
  __shared__ unsigned long long s_tk;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_tk = atomicAdd(ticket, (unsigned long long)BATCH);
    __syncthreads();
    const int64_t e0 = (int64_t)s_tk;
    if (e0 >= num_envs) break;
    const int64_t eb = e0 + (int64_t)warp * EPW3;
    if (eb >= num_envs) continue;
    {
      // all loads of the warp's envs are issued before the first use
      float rq[EPW3][3], rw[EPW3][3], rp[EPW3][3], r1[EPW3];
      const int nd = M.nd;
#pragma unroll
      for (int k = 0; k < EPW3; k++) {
        const int64_t ek = eb + k < num_envs ? eb + k : (int64_t)num_envs - 1;
        const int64_t e = env_first + (int64_t)env_stride * ek;
        if (lc.dyn && lane > 0) {
          const float* ds = bf.dof_state + (e * nd + lc.dof0) * 2;
#pragma unroll
          for (int j = 0; j < 3; j++) { rq[k][j] = ds[2 * j]; rw[k][j] = ds[2 * j + 1]; rp[k][j] = bf.pd_targets[e * nd + lc.dof0 + j]; }
        }
        r1[k] = lane < 13 ? bf.root_states[e * bf.actors_per_env * 13 + lane] : (lane < 19 ? ext_wrench[e * 6 + lane - 13] : 0.f);
      }
#pragma unroll
      for (int k = 0; k < EPW3; k++) {
        if (eb + k >= num_envs) break;
        Lane<float> L;
        float pdtar[3] = {0.f, 0.f, 0.f}, extF[3] = {0.f, 0.f, 0.f}, extT[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; j++) { L.Q[j] = 0.f; L.qj[j] = 0.f; }
        L.Q[3] = 1.f; L.qj[3] = 1.f;
#pragma unroll
        for (int j = 0; j < 3; j++) { L.p[j] = 0.f; L.w[j] = 0.f; L.v[j] = 0.f; L.wt[j] = 0.f; }
        float rr[19];
#pragma unroll
        for (int j = 0; j < 19; j++) rr[j] = __shfl_sync(FULL, r1[k], j);
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 3; j++) { L.p[j] = rr[j]; L.v[j] = rr[7 + j]; L.w[j] = rr[10 + j]; extF[j] = rr[13 + j]; extT[j] = rr[16 + j]; }
#pragma unroll
          for (int j = 0; j < 4; j++) L.Q[j] = rr[3 + j];
          qnormalize(L.Q);
        }
        if (lc.dyn && lane > 0) {
#pragma unroll
          for (int j = 0; j < 3; j++) { L.wt[j] = rw[k][j]; pdtar[j] = rp[k][j]; }
          qexp(rq[k], L.qj);
        }
        pk_store_state<float>(wrec + k * ENV_STRIDE, lc, lane, L, pdtar, extF, extT);
      }
    }
    __syncwarp();
    const bool valid = eb + g < num_envs;
    Ball<float> ball;
    ball_clear(ball);
    const int64_t erow_g = env_first + (int64_t)env_stride * (valid ? eb + g : eb);
    if (cfg.has_ball && valid && u == BALL_SLOT3) ball_load(bf, erow_g, ball);
    control_step_packed3<float>(B, verts, pc, wrec, lane, valid, ball, false);
    if (cfg.has_ball && valid && u == BALL_SLOT3) ball_writeback(bf, erow_g, ball);
    for (int k = 0; k < EPW3; k++) {
      if (eb + k >= num_envs) break;
      const int64_t e = env_first + (int64_t)env_stride * (eb + k);
      Lane<float> L;
      float cf[3], dq[3];
      pk_load_state<float>(wrec + k * ENV_STRIDE, lc, lane, L, cf);
      epilogue_writeback(bf, cfg, M, lc, lane, e, L, cf, dq);
    }
    __syncwarp();
  }
}
#endif  // B200ENV_WITH_PACKED3


template <typename T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
physics_kernel_packed(const DevBlob* __restrict__ gblob, uint32_t blob_bytes, const b200_cfg_t* __restrict__ gcfg, int n, int n_steps,
                      T* root, T* dof_pos, T* dof_vel, const T* pd_tar, const T* ext, T* rb_out, T* contact_out, T* ballio, int32_t* hits) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar;
  load_blob(smem, gblob, blob_bytes, &mbar);
  const DevBlob& B = *reinterpret_cast<const DevBlob*>(smem);
  const b200_model_t& M = B.m;
  const float* verts = reinterpret_cast<const float*>(smem + sizeof(DevBlob));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  T* wrec = reinterpret_cast<T*>(smem + ((blob_bytes + 15) & ~15u)) + (size_t)warp * EPW * ENV_STRIDE;
  const int64_t eb = ((int64_t)blockIdx.x * WARPS + warp) * EPW;
  if (eb >= n) return;
  LaneConst lc = lane_const(M, lane);
  if (lc.active) lc.rix = B.t.rix[lane];
  const int nb = M.nb, nd = M.nd;
  PhysCfg<T> pc = make_phys_cfg<T>(*gcfg);
  const bool with_ball = pc.has_ball && ballio != nullptr;
  pc.has_ball = with_ball;
  const int g = lane >> 3, s = lane & 7;
  for (int k = 0; k < EPW; k++) {
    const int64_t e = eb + k;
    if (e >= n) break;
    Lane<T> L;
    for (int j = 0; j < 4; j++) { L.Q[j] = 0; L.qj[j] = 0; }
    L.Q[3] = 1; L.qj[3] = 1;
    for (int j = 0; j < 3; j++) { L.p[j] = 0; L.w[j] = 0; L.v[j] = 0; L.wt[j] = 0; }
    T pdt[3] = {0, 0, 0}, eF[3] = {0, 0, 0}, eT[3] = {0, 0, 0};
    if (lane == 0) {
      const T* rs = root + e * 13;
      for (int j = 0; j < 3; j++) { L.p[j] = rs[j]; L.v[j] = rs[7 + j]; L.w[j] = rs[10 + j]; }
      for (int j = 0; j < 4; j++) L.Q[j] = rs[3 + j];
      qnormalize(L.Q);
      if (ext) for (int j = 0; j < 3; j++) { eF[j] = ext[e * 6 + j]; eT[j] = ext[e * 6 + 3 + j]; }
    }
    if (lc.dyn && lane > 0) {
      T q[3];
      for (int j = 0; j < 3; j++) { q[j] = dof_pos[e * nd + lc.dof0 + j]; L.wt[j] = dof_vel[e * nd + lc.dof0 + j]; pdt[j] = pd_tar[e * nd + lc.dof0 + j]; }
      qexp(q, L.qj);
    }
    pk_store_state<T>(wrec + k * ENV_STRIDE, lc, lane, L, pdt, eF, eT);
  }
  __syncwarp();
  const bool valid = eb + g < n;
  Ball<T> ball;
  ball_clear(ball);
  if (with_ball && valid && s == BALL_SLOT) {
    const int64_t e = eb + g;
    for (int j = 0; j < 3; j++) { ball.p[j] = ballio[e * 13 + j]; ball.v[j] = ballio[e * 13 + 7 + j]; ball.w[j] = ballio[e * 13 + 10 + j]; }
  }
  for (int st_ = 0; st_ < n_steps; st_++) {
    control_step_packed<T>(B, verts, pc, wrec, lane, valid, ball, false);
    if (st_ + 1 < n_steps) {  // the state crosses control steps as exp-map coordinates
      for (int k = 0; k < EPW; k++) {
        if (eb + k >= n) break;
        if (lc.dyn && lane > 0) {
          T* rec = wrec + k * ENV_STRIDE + lc.rix * REC;
          T qj[4], q[3];
          ldr<R_QJ, 4>(rec, qj);
          qlog(qj, q);
          qexp(q, qj);
          str<R_QJ, 4>(rec, qj);
        }
        if (lane == 0) {  // the racket reaction does not carry over a control step in the lane kernel either (ball.rF is per substep)
        }
      }
      __syncwarp();
    }
  }
  for (int k = 0; k < EPW; k++) {
    const int64_t e = eb + k;
    if (e >= n) break;
    Lane<T> L;
    T cf[3];
    pk_load_state<T>(wrec + k * ENV_STRIDE, lc, lane, L, cf);
    if (lane == 0) {
      T* rs = root + e * 13;
      for (int j = 0; j < 3; j++) { rs[j] = L.p[j]; rs[7 + j] = L.v[j]; rs[10 + j] = L.w[j]; }
      for (int j = 0; j < 4; j++) rs[3 + j] = L.Q[j];
    }
    if (lc.dyn && lane > 0) {
      T q[3];
      qlog(L.qj, q);
      for (int j = 0; j < 3; j++) { dof_pos[e * nd + lc.dof0 + j] = q[j]; dof_vel[e * nd + lc.dof0 + j] = L.wt[j]; }
    }
    if (lc.active) {
      T* rb = rb_out + (e * nb + lane) * 13;
      for (int j = 0; j < 3; j++) { rb[j] = L.p[j]; rb[7 + j] = L.v[j]; rb[10 + j] = L.w[j]; }
      for (int j = 0; j < 4; j++) rb[3 + j] = L.Q[j];
      if (contact_out) for (int j = 0; j < 3; j++) contact_out[(e * nb + lane) * 3 + j] = cf[j];
    }
  }
  if (with_ball && valid && s == BALL_SLOT) {
    const int64_t e = eb + g;
    for (int j = 0; j < 3; j++) { ballio[e * 13 + j] = ball.p[j]; ballio[e * 13 + 7 + j] = ball.v[j]; ballio[e * 13 + 10 + j] = ball.w[j]; }
    if (hits) hits[e] = ball.hits;
  }
}

// ------------------------------------------------------------------------------------------
// reset: ref-state init from the MoCap buffer (_reset_ref_state_init :489-528, _set_env_state :741-755)
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
reset_kernel(const DevBlob* __restrict__ gblob, const b200_cfg_t* __restrict__ gcfg, b200_buffers_t bf, b200_motion_lib_t ml,
             const int64_t* __restrict__ env_ids, const float* __restrict__ times, int n) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * WARPS_PER_CTA + warp;
  if (i >= n) return;
  const b200_model_t& M = gblob->m;
  const b200_cfg_t& cfg = *gcfg;
  const int64_t e = env_ids[i];
  const float t = times[i];
  const int64_t mid = bf.motion_ids[e];
  const int nd = M.nd, nbl = ml.num_lib_bodies;
  const int dof0 = lane < M.nb ? M.dof_of_body[lane] : -1;
  const MotionSample s = sample_motion(ml, M, mid, t, lane, cfg.ground_tolerance);
  const bool is_body = lane < nbl;
  float zero3[3] = {0.f, 0.f, 0.f};
  if (lane == 0) {
    float* rs = bf.root_states + e * bf.actors_per_env * 13;
#pragma unroll
    for (int k = 0; k < 3; k++) { rs[k] = s.rb_pos[k]; rs[7 + k] = s.root_vel[k]; rs[10 + k] = s.root_ang_vel[k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) rs[3 + k] = s.rb_rot[k];
    bf.progress_buf[e] = 0; bf.reset_buf[e] = 0; bf.terminate_buf[e] = 0;  // humanoid_smpl.py:170-172
    bf.ref_motion_times[e] = t;
  }
  float dv[3] = {0.f, 0.f, 0.f};
  if (is_body) {
    float* rb = bf.rigid_body_state + (e * bf.bodies_per_env + lane) * 13;  // :746-749 (+ _refresh_sim_tensors :457-463)
#pragma unroll
    for (int k = 0; k < 3; k++) { rb[k] = s.rb_pos[k]; rb[7 + k] = 0.f; rb[10 + k] = 0.f; }
#pragma unroll
    for (int k = 0; k < 4; k++) rb[3 + k] = s.rb_rot[k];
    if (dof0 >= 0) {
      float* ds = bf.dof_state + e * nd * 2;
#pragma unroll
      for (int k = 0; k < 3; k++) { dv[k] = ml.dvs[s.f0 * nd + dof0 + k]; ds[(dof0 + k) * 2] = s.dof[k]; ds[(dof0 + k) * 2 + 1] = dv[k]; }
    }
  } else if (lane < M.nb) {  // welded extra bodies (Racket): placed by the first step's FK; park at the parent pose
    float* rb = bf.rigid_body_state + (e * bf.bodies_per_env + lane) * 13;
    const int par = M.parent[lane];
    const float* prow = bf.rigid_body_state + (e * bf.bodies_per_env + par) * 13;
    (void)prow;
#pragma unroll
    for (int k = 0; k < 13; k++) rb[k] = (k == 6) ? 1.f : 0.f;
  }
  // targets = MoCap state one control step ahead (:525 -> _set_target_motion_state :594-624)
  const float step_dt = __fmul_rn((float)cfg.control_freq_inv, cfg.sim_dt);
  const MotionSample tg = sample_motion(ml, M, mid, __fadd_rn(t, step_dt), lane, cfg.ground_tolerance);
  store_targets(bf, cfg, ml, M, tg, e, lane, dof0);
  // _compute_observations(env_ids) (humanoid_smpl.py:158)
  store_obs_raw(bf.obs_buf + e * bf.num_obs, nbl, nd, cfg.shape_dim, lane, is_body, dof0, s.rb_pos, s.rb_rot, zero3, zero3, s.dof, dv,
                bf.motion_bodies + e * cfg.shape_dim);
}

// MotionLib.get_motion_state for arbitrary (id, time) pairs
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
motion_state_kernel(const DevBlob* __restrict__ gblob, const b200_cfg_t* __restrict__ gcfg, b200_motion_lib_t ml,
                    const int64_t* __restrict__ ids, const float* __restrict__ times, int n, float* root_pos, float* root_rot,
                    float* dof_pos, float* root_vel, float* root_ang_vel, float* dof_vel, float* key_pos, float* rb_pos,
                    float* rb_rot) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * WARPS_PER_CTA + warp;
  if (i >= n) return;
  const b200_model_t& M = gblob->m;
  const b200_cfg_t& cfg = *gcfg;
  const int nd = M.nd, nbl = ml.num_lib_bodies;
  const int dof0 = lane < M.nb ? M.dof_of_body[lane] : -1;
  const MotionSample s = sample_motion(ml, M, ids[i], times[i], lane, cfg.ground_tolerance);
  if (lane < nbl) {
    if (rb_pos) for (int k = 0; k < 3; k++) rb_pos[(i * nbl + lane) * 3 + k] = s.rb_pos[k];
    if (rb_rot) for (int k = 0; k < 4; k++) rb_rot[(i * nbl + lane) * 4 + k] = s.rb_rot[k];
    if (dof_pos && dof0 >= 0) for (int k = 0; k < 3; k++) dof_pos[i * nd + dof0 + k] = s.dof[k];
    if (key_pos)
      for (int k = 0; k < cfg.num_key; k++)
        if (cfg.key_body[k] == lane) for (int j = 0; j < 3; j++) key_pos[(i * cfg.num_key + k) * 3 + j] = s.rb_pos[j];
  }
  if (lane == 0) {
    for (int k = 0; k < 3; k++) {
      if (root_pos) root_pos[i * 3 + k] = s.rb_pos[k];
      if (root_vel) root_vel[i * 3 + k] = s.root_vel[k];
      if (root_ang_vel) root_ang_vel[i * 3 + k] = s.root_ang_vel[k];
    }
    if (root_rot) for (int k = 0; k < 4; k++) root_rot[i * 4 + k] = s.rb_rot[k];
  }
  if (dof_vel) for (int k = lane; k < nd; k += 32) dof_vel[i * nd + k] = ml.dvs[s.f0 * nd + k];
}

// _init_context (humanoid_smpl_im.py:530-563): the P-frame MoCap window of each listed env written straight into
// context_feat[row, j, :] = rb_pos | rb_rot | dof_pos | rb_pos | dof_pos and context_mask[row, j]; warp per (env, frame).
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
motion_context_kernel(const DevBlob* __restrict__ gblob, const b200_cfg_t* __restrict__ gcfg, b200_motion_lib_t ml,
                      const int64_t* __restrict__ env_ids, const int64_t* __restrict__ ids, const float* __restrict__ times, int n, int P,
                      int first, float dt, float two_dt, float* __restrict__ feat, uint8_t* __restrict__ mask) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * WARPS_PER_CTA + warp;
  if (w >= (int64_t)n * P) return;
  const int64_t i = w / P;
  const int j = (int)(w - i * P);
  const b200_model_t& M = gblob->m;
  const int nd = M.nd, nbl = ml.num_lib_bodies;
  const int dof0 = lane < M.nb ? M.dof_of_body[lane] : -1;
  const int64_t mid = ids[i];
  // torch: (motion_times + dt).unsqueeze(-1) + dt * arange(first, first + P), all float32, op by op
  const float t = __fadd_rn(__fadd_rn(times[i], dt), __fmul_rn((float)(first + j), dt));
  const MotionSample s = sample_motion(ml, M, mid, t, lane, gcfg->ground_tolerance);
  const int64_t row = env_ids ? env_ids[i] : i;
  const int W = nbl * 3 + nbl * 4 + nd + nbl * 3 + nd;
  float* o = feat + (row * P + j) * W;
  if (lane < nbl) {
#pragma unroll
    for (int k = 0; k < 3; k++) { o[lane * 3 + k] = s.rb_pos[k]; o[nbl * 7 + nd + lane * 3 + k] = s.rb_pos[k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) o[nbl * 3 + lane * 4 + k] = s.rb_rot[k];
    if (dof0 >= 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) { o[nbl * 7 + dof0 + k] = s.dof[k]; o[nbl * 10 + nd + dof0 + k] = s.dof[k]; }
    }
  }
  if (lane == 0) mask[row * P + j] = t <= __fadd_rn(ml.motion_lengths[mid], two_dt) ? 1 : 0;
}

// compute_humanoid_observations_imitation (humanoid_smpl_im.py:773-850): warp per env, lane per body
// element strides of the state arrays of obs_imitation_kernel: contiguous [n, nb, 3] / [n, nb, 4] / [n, nd] copies (legacy entry) or the
// Isaac-layout rows themselves (rigid_body_state [n, bodies_per_env, 13], dof_state [n, nd, 2]) - no gather copies
#define OBS_ROW_MAX 768   // widest observation row: 734 (24 bodies, 69 dof, shape 11)
struct ObsStrides { int p_row, p_elem, q_row, q_elem, d_row, d_elem; };
struct ObsBf16 { __nv_bfloat16* out; int ld; const float* mean; const float* rstd; float clamp; };
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
obs_imitation_kernel(int n, int nb, int nd, int shape_dim, const float* __restrict__ body_pos, const float* __restrict__ body_rot,
                     const float* __restrict__ target_pos, const float* __restrict__ target_rot, const float* __restrict__ dof_pos,
                     const float* __restrict__ dof_vel, const float* __restrict__ target_dof_pos, const float* __restrict__ body_vel,
                     const float* __restrict__ body_ang_vel, const float* __restrict__ motion_bodies, int local_root_obs,
                     int root_height_obs, float* __restrict__ obs, int jpos, ObsStrides st, ObsBf16 ob) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * WARPS_PER_CTA + warp;
  if (e >= n) return;
  // jpos = compute_humanoid_observations_imitation_jpos (:853-915): no rotation / heading / dof targets
  const int W = jpos ? 1 + (nb - 1) * 3 + nb * 6 + nb * 3 + nb * 3 + nd + 1 + 2 + nb * 3 + shape_dim
                     : 1 + (nb - 1) * 3 + nb * 6 + nb * 3 + nb * 3 + nd + 1 + 6 + 2 + 2 + nd + nb * 3 + nb * 6 + shape_dim;
  // every value goes through put() into the warp's row in shared memory; the row leaves at the end as coalesced 8-byte stores
  // (the per-lane pieces - 3 / 6 values per body - are 12 / 24 bytes apart: written straight to global memory they made the launch
  // ~30 us for 36 MB) - the float row the reference returns and, when asked for, the bf16 operand row of the policy's first layer =
  // clamp((x - mean) * rstd, -clamp, clamp) (RunningMeanStd + the +-5 clamp of im_player.py:187-190) from the same values
  __shared__ __align__(16) float s_obs[WARPS_PER_CTA][OBS_ROW_MAX];
  if (W > OBS_ROW_MAX) return;            // cannot happen with the 24-body SMPL humanoid (W = 734 / 513); the host entry points check nb
  float* srow = s_obs[warp];
  auto put = [&](int idx, float val) { srow[idx] = val; };
  const bool act = lane < nb;
  float p[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, v[3] = {0, 0, 0}, w[3] = {0, 0, 0}, tp[3] = {0, 0, 0}, tq[4] = {0, 0, 0, 1};
  if (act) {
    const int64_t i3 = (e * nb + lane) * 3, i4 = (e * nb + lane) * 4;
    const int64_t s3 = e * st.p_row + (int64_t)lane * st.p_elem, s4 = e * st.q_row + (int64_t)lane * st.q_elem;   // state arrays: own strides
#pragma unroll
    for (int k = 0; k < 3; k++) { p[k] = body_pos[s3 + k]; v[k] = body_vel[s3 + k]; w[k] = body_ang_vel[s3 + k]; tp[k] = target_pos[i3 + k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) { q[k] = body_rot[s4 + k]; tq[k] = target_rot[i4 + k]; }
  }
  float rp[3], rq[4], trp[3], trq[4];
#pragma unroll
  for (int k = 0; k < 3; k++) { rp[k] = shfl(p[k], 0); trp[k] = shfl(tp[k], 0); }
#pragma unroll
  for (int k = 0; k < 4; k++) { rq[k] = shfl(q[k], 0); trq[k] = shfl(tq[k], 0); }
  float root_rot[4], hq[4], t_root_rot[4];
  ref_remove_base_rot(rq, root_rot);
  const float heading = ref_calc_heading(root_rot);
  ref_heading_quat(-heading, hq);
  ref_remove_base_rot(trq, t_root_rot);
  const float t_heading = ref_calc_heading(t_root_rot);
  int off = 0;
  if (lane == 0) put(0, root_height_obs ? rp[2] : 0.f);
  off = 1;
  if (act) {
    float d[3] = {p[0] - rp[0], p[1] - rp[1], p[2] - rp[2]}, l[3];
    ref_quat_rotate(hq, d, l);
    if (lane > 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) put(off + (lane - 1) * 3 + k, l[k]);
    }
  }
  off += (nb - 1) * 3;
  if (act) {
    float lq[4], tn[6];
    qmul(hq, q, lq);
    ref_tan_norm(lq, tn);
    if (lane == 0 && local_root_obs) ref_tan_norm(root_rot, tn);  // quirk kept (:806-809)
#pragma unroll
    for (int k = 0; k < 6; k++) put(off + lane * 6 + k, tn[k]);
  }
  off += nb * 6;
  if (act) {
    float l[3];
    ref_quat_rotate(hq, v, l);
#pragma unroll
    for (int k = 0; k < 3; k++) put(off + lane * 3 + k, l[k]);
    ref_quat_rotate(hq, w, l);
#pragma unroll
    for (int k = 0; k < 3; k++) put(off + nb * 3 + lane * 3 + k, l[k]);
  }
  off += nb * 6;
  for (int k = lane; k < nd; k += 32) put(off + k, dof_vel[e * st.d_row + (int64_t)k * st.d_elem]);
  off += nd;
  if (jpos) {
    if (lane == 0) {
      put(off, rp[2] - trp[2]);
      float d[3] = {trp[0] - rp[0], trp[1] - rp[1], trp[2] - rp[2]}, l[3];
      ref_quat_rotate(hq, d, l);
      put(off + 1, l[0]); put(off + 2, l[1]);
    }
    off += 3;
  } else {
    if (lane == 0) {
      put(off, rp[2] - trp[2]);
      float cj[4] = {-root_rot[0], -root_rot[1], -root_rot[2], root_rot[3]}, rel[4], tn[6];
      qmul(t_root_rot, cj, rel);
      ref_tan_norm(rel, tn);
#pragma unroll
      for (int k = 0; k < 6; k++) put(off + 1 + k, tn[k]);
      float d[3] = {trp[0] - rp[0], trp[1] - rp[1], trp[2] - rp[2]}, l[3];
      ref_quat_rotate(hq, d, l);
      put(off + 7, l[0]); put(off + 8, l[1]);
      const float dh = t_heading - heading;
      put(off + 9, cosf(dh)); put(off + 10, sinf(dh));
    }
    off += 11;
    for (int k = lane; k < nd; k += 32) put(off + k, target_dof_pos[e * nd + k] - dof_pos[e * st.d_row + (int64_t)k * st.d_elem]);
    off += nd;
  }
  if (act) {
    float d[3] = {tp[0] - p[0], tp[1] - p[1], tp[2] - p[2]}, l[3];
    ref_quat_rotate(hq, d, l);
#pragma unroll
    for (int k = 0; k < 3; k++) put(off + lane * 3 + k, l[k]);
  }
  off += nb * 3;
  if (!jpos) {
    if (act) {
      float cj[4] = {-q[0], -q[1], -q[2], q[3]}, rel[4], tn[6];
      qmul(cj, tq, rel);
      ref_tan_norm(rel, tn);
#pragma unroll
      for (int k = 0; k < 6; k++) put(off + lane * 6 + k, tn[k]);
    }
    off += nb * 6;
  }
  if (lane < shape_dim) put(off + lane, motion_bodies[e * shape_dim + lane]);
  __syncwarp();
  float* orow = obs + e * W;
  if ((((uintptr_t)orow) & 7) == 0 && (W & 1) == 0) {
    for (int k = lane; k < W / 2; k += 32) reinterpret_cast<float2*>(orow)[k] = reinterpret_cast<const float2*>(srow)[k];
  } else {
    for (int k = lane; k < W; k += 32) orow[k] = srow[k];
  }
  if (ob.out) {
    __nv_bfloat16* brow = ob.out + e * (int64_t)ob.ld;
    for (int k = 2 * lane; k < W; k += 64) {
      float x0 = srow[k], x1 = k + 1 < W ? srow[k + 1] : 0.f;
      if (ob.mean) {
        x0 = (x0 - __ldg(ob.mean + k)) * __ldg(ob.rstd + k);
        if (k + 1 < W) x1 = (x1 - __ldg(ob.mean + k + 1)) * __ldg(ob.rstd + k + 1);
      }
      x0 = fminf(fmaxf(x0, -ob.clamp), ob.clamp);
      x1 = fminf(fmaxf(x1, -ob.clamp), ob.clamp);
      if (k + 1 < W) *reinterpret_cast<__nv_bfloat162*>(brow + k) = __floats2bfloat162_rn(x0, x1);   // ld is a multiple of 64: 4-byte aligned
      else brow[k] = __float2bfloat16_rn(x0);
    }
  }
}

// physics-only entry used by the parity tests (float or double)
template <typename T>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
physics_kernel(const DevBlob* __restrict__ gblob, uint32_t blob_bytes, const b200_cfg_t* __restrict__ gcfg, int n, int n_steps,
               T* root, T* dof_pos, T* dof_vel, const T* pd_tar, const T* ext, T* rb_out, T* contact_out, T* ballio, int32_t* hits) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar;
  load_blob(smem, gblob, blob_bytes, &mbar);
  const DevBlob& B = *reinterpret_cast<const DevBlob*>(smem);
  const b200_model_t& M = B.m;
  const float* verts = reinterpret_cast<const float*>(smem + sizeof(DevBlob));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * WARPS_PER_CTA + warp;
  if (e >= n) return;
  const LaneConst lc = lane_const(M, lane);
  const int nb = M.nb, nd = M.nd;
  const PhysCfg<T> pc = make_phys_cfg<T>(*gcfg);
  Lane<T> L;
  for (int k = 0; k < 4; k++) { L.Q[k] = 0; L.qj[k] = 0; }
  L.Q[3] = 1; L.qj[3] = 1;
  for (int k = 0; k < 3; k++) { L.p[k] = 0; L.w[k] = 0; L.v[k] = 0; L.wt[k] = 0; }
  T pdt[3] = {0, 0, 0}, eF[3] = {0, 0, 0}, eT[3] = {0, 0, 0}, cf[3] = {0, 0, 0};
  if (lane == 0) {
    const T* rs = root + e * 13;
    for (int k = 0; k < 3; k++) { L.p[k] = rs[k]; L.v[k] = rs[7 + k]; L.w[k] = rs[10 + k]; }
    for (int k = 0; k < 4; k++) L.Q[k] = rs[3 + k];
    qnormalize(L.Q);
    if (ext) for (int k = 0; k < 3; k++) { eF[k] = ext[e * 6 + k]; eT[k] = ext[e * 6 + 3 + k]; }
  }
  if (lc.dyn && lane > 0) {
    T q[3];
    for (int k = 0; k < 3; k++) { q[k] = dof_pos[e * nd + lc.dof0 + k]; L.wt[k] = dof_vel[e * nd + lc.dof0 + k]; pdt[k] = pd_tar[e * nd + lc.dof0 + k]; }
    qexp(q, L.qj);
  }
  Ball<T> ball;
  ball_clear(ball);
  const bool with_ball = pc.has_ball && ballio != nullptr;
  PhysCfg<T> pcb = pc;
  pcb.has_ball = with_ball;
  if (with_ball && lane == BALL_LANE) {
    for (int k = 0; k < 3; k++) { ball.p[k] = ballio[e * 13 + k]; ball.v[k] = ballio[e * 13 + 7 + k]; ball.w[k] = ballio[e * 13 + 10 + k]; }
  }
  for (int s = 0; s < n_steps; s++) {
    control_step<T>(B, verts, pcb, lc, lane, L, pdt, eF, eT, cf, ball);
    if (s + 1 < n_steps && lc.dyn && lane > 0) {  // the state crosses control steps as exp-map coordinates
      T q[3];
      qlog(L.qj, q);
      qexp(q, L.qj);
    }
  }
  if (lane == 0) {
    T* rs = root + e * 13;
    for (int k = 0; k < 3; k++) { rs[k] = L.p[k]; rs[7 + k] = L.v[k]; rs[10 + k] = L.w[k]; }
    for (int k = 0; k < 4; k++) rs[3 + k] = L.Q[k];
  }
  if (lc.dyn && lane > 0) {
    T q[3];
    qlog(L.qj, q);
    for (int k = 0; k < 3; k++) { dof_pos[e * nd + lc.dof0 + k] = q[k]; dof_vel[e * nd + lc.dof0 + k] = L.wt[k]; }
  }
  if (lc.active) {
    T* rb = rb_out + (e * nb + lane) * 13;
    for (int k = 0; k < 3; k++) { rb[k] = L.p[k]; rb[7 + k] = L.v[k]; rb[10 + k] = L.w[k]; }
    for (int k = 0; k < 4; k++) rb[3 + k] = L.Q[k];
    if (contact_out) for (int k = 0; k < 3; k++) contact_out[(e * nb + lane) * 3 + k] = cf[k];
  }
  if (with_ball && lane == BALL_LANE) {
    for (int k = 0; k < 3; k++) { ballio[e * 13 + k] = ball.p[k]; ballio[e * 13 + 7 + k] = ball.v[k]; ballio[e * 13 + 10 + k] = ball.w[k]; }
    if (hits) hits[e] = ball.hits;
  }
}

// ------------------------------------------------------------------------------------------
// C ABI
extern "C" {

int b200env_abi_version(void) { return B200_ABI_VERSION; }
const char* b200env_last_error(void) { return g_err; }

int b200env_create(const b200_model_t* model, const float* verts, const b200_cfg_t* cfg, int32_t num_envs, int32_t device,
                   b200env_handle* out) {
  if (!model || !verts || !cfg || !out) return fail(-1, "b200env_create: null argument%s");
  if (model->nb < 1 || model->nb > B200_MAX_BODIES || model->nd > B200_MAX_DOF || model->nd % 3)
    return fail(-2, "b200env_create: model dimensions out of range%s");
  if (model->max_depth >= MAX_LEVELS) return fail(-2, "b200env_create: kinematic tree too deep%s");
  if (model->vmax % 4) return fail(-2, "b200env_create: vmax must be a multiple of 4%s");
  for (int b = 0; b < model->nb; b++)
    if (model->nverts[b] < 0 || model->nverts[b] > 64 || model->nverts[b] > model->vmax)
      return fail(-2, "b200env_create: at most 64 hull vertices per body (contact_hull keeps a 64-bit penetration mask)%s");
  if (num_envs < 1) return fail(-2, "b200env_create: num_envs must be positive%s");
  if (cfg->has_ball && model->nb > BALL_LANE) return fail(-2, "b200env_create: has_ball needs nb <= 31 (lane 31 integrates the ball)%s");
  if (cfg->has_ball && cfg->racket_body >= model->nb) return fail(-2, "b200env_create: racket_body out of range%s");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(-3, "b200env_create: no CUDA device - this library has no CPU fallback%s");
  CUDA_OK(cudaSetDevice(device));
  b200env* h = new b200env();
  memset(h, 0, sizeof(*h));
  h->env_stride = 1;
  h->device = device;
  h->num_envs = num_envs;
  h->model = *model;
  h->cfg = *cfg;
  // tree tables
  DevBlob hb;
  int slots_ok = 1;
  const int trc = build_dev_blob(model, hb, &slots_ok);
  if (trc == -1) { delete h; return fail(-2, "b200env_create: bodies must be in topological order%s"); }
  if (trc == -2) { delete h; return fail(-2, "b200env_create: too many children per body%s"); }
  h->packed_ok = model->nb <= B200_MAX_BODIES_PK && slots_ok;
  hull_vertex_radius(model, verts, hb.t.vrho);
  hull_bounding_spheres(model, verts, hb.t.bs);
  const char* kv = getenv("B200ENV_KERNEL");
  h->packed = h->packed_ok && !(kv && strcmp(kv, "lane") == 0);
  const char* sv = getenv("B200ENV_SPLIT");
  h->split = h->packed && !(sv && strcmp(sv, "0") == 0);
  // packed3 (2 envs per warp, 3 lanes per body in the backward pass; csrc/packed3.cuh): needs the split form and at most 5 bodies
  // per tree depth.  B200ENV_KERNEL=packed3 selects it (A/B against the 4-envs-per-warp kernel, profiles/).
#if B200ENV_WITH_PACKED3
  int wide = 0;
  for (int d = 0; d < MAX_LEVELS; d++) if (hb.t.lvl_all[d][SLOTS3] >= 0) wide = 1;
  h->packed3 = h->split && !wide && kv && strcmp(kv, "packed3") == 0;
#else
  h->packed3 = 0;
#endif
  const char* so = getenv("B200ENV_SORT");
  h->sort = h->split && !h->packed3 && so && strcmp(so, "1") == 0;
  // one-wave form (csrc/packed_t.cuh): needs the split form and a tree that fits 3 column blocks x 8 owner slots
  // It is the default where it applies (round 2: -8 % per config-2 step, -6 % per config-3 step, profiles/r2t_tmem.md);
  // B200ENV_KERNEL=packed keeps the two-round kernel for A/B.
  h->tmem = h->split && !h->packed3 && !h->sort && hb.t.pt_ok && hb.t.pt_nmbox <= PT_MBOX_MAX && !(kv && strcmp(kv, "packed") == 0);
  if (cfg->has_ball && cfg->ball_body_contact && !h->packed) {
    delete h;
    return fail(-2, "b200env_create: ball_body_contact needs the packed kernels (the lane-per-body kernel does not have it)%s");
  }
  const size_t vbytes = (size_t)model->nb * model->vmax * 3 * sizeof(float);
  h->blob_bytes = sizeof(DevBlob) + ((vbytes + 15) & ~(size_t)15);
  CUDA_OK(cudaMalloc(&h->d_blob, h->blob_bytes));
  CUDA_OK(cudaMemset(h->d_blob, 0, h->blob_bytes));
  CUDA_OK(cudaMemcpy(h->d_blob, &hb, sizeof(hb), cudaMemcpyHostToDevice));
  {  // AoS [nb][vmax][3] (ABI) -> SoA [nb][3][vmax] (what contact_hull reads with 128-bit loads)
    std::vector<float> soa((size_t)model->nb * model->vmax * 3, 0.0f);
    verts_to_soa(model, verts, soa.data());
    CUDA_OK(cudaMemcpy((char*)h->d_blob + sizeof(DevBlob), soa.data(), vbytes, cudaMemcpyHostToDevice));
  }
  CUDA_OK(cudaMalloc(&h->d_cfg, sizeof(b200_cfg_t)));
  CUDA_OK(cudaMemcpy(h->d_cfg, cfg, sizeof(b200_cfg_t), cudaMemcpyHostToDevice));
  if (h->sort) {
    CUDA_OK(cudaMalloc(&h->d_bin, sizeof(int32_t) * num_envs));
    CUDA_OK(cudaMalloc(&h->d_perm, sizeof(int32_t) * num_envs));
    CUDA_OK(cudaMalloc(&h->d_pos, sizeof(uint32_t) * num_envs));
    CUDA_OK(cudaMalloc(&h->d_cnt, sizeof(uint32_t) * SORT_BINS));
  }
  CUDA_OK(cudaMalloc(&h->d_ticket, sizeof(unsigned long long)));
  CUDA_OK(cudaMemset(h->d_ticket, 0, sizeof(unsigned long long)));
  *out = h;
  return 0;
}

int b200env_destroy(b200env_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaFree(h->d_blob);
  cudaFree(h->d_face_planes); cudaFree(h->d_face_tris);
  cudaFree(h->d_cfg);
  cudaFree(h->d_ticket);
  cudaFree(h->d_ext);
  cudaFree(h->d_bin); cudaFree(h->d_perm); cudaFree(h->d_pos); cudaFree(h->d_cnt);
  if (h->tev) { for (cudaEvent_t e : *h->tev) cudaEventDestroy(e); delete h->tev; }
  delete h;
  return 0;
}

int b200env_set_hull_faces(b200env_handle h, const float* planes, const uint8_t* tris, const int32_t* ntris, int32_t tmax) {
  if (!h || !planes || !tris || !ntris) return fail(-1, "b200env_set_hull_faces: null argument%s");
  if (tmax < 1 || tmax > 255) return fail(-2, "b200env_set_hull_faces: tmax must be in 1..255%s");
  const int nb = h->model.nb;
  for (int b = 0; b < nb; b++) {
    if (ntris[b] < 0 || ntris[b] > tmax) return fail(-2, "b200env_set_hull_faces: ntris out of range%s");
    for (int t = 0; t < ntris[b]; t++)
      for (int k = 0; k < 3; k++)
        if (tris[((size_t)b * tmax + t) * 4 + k] >= h->model.nverts[b]) return fail(-2, "b200env_set_hull_faces: vertex index out of range%s");
  }
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaDeviceSynchronize());
  cudaFree(h->d_face_planes); cudaFree(h->d_face_tris);
  h->d_face_planes = nullptr; h->d_face_tris = nullptr;
  const size_t nf = (size_t)nb * tmax;
  CUDA_OK(cudaMalloc(&h->d_face_planes, nf * 4 * sizeof(float)));
  CUDA_OK(cudaMalloc(&h->d_face_tris, nf * 4));
  CUDA_OK(cudaMemcpy(h->d_face_planes, planes, nf * 4 * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(h->d_face_tris, tris, nf * 4, cudaMemcpyHostToDevice));
  // patch the face fields of the tree block in place (the rest of the blob stays what b200env_create uploaded)
  DevTree t;
  memset(&t, 0, sizeof(t));
  t.face_planes = h->d_face_planes;
  t.face_tris = h->d_face_tris;
  t.face_tmax = tmax;
  for (int b = 0; b < nb; b++) t.ntris[b] = ntris[b];
  const size_t o0 = offsetof(DevTree, face_planes), o1 = offsetof(DevTree, pt_blk);   // the face fields only: the tables behind them stay
  CUDA_OK(cudaMemcpy((char*)h->d_blob + offsetof(DevBlob, t) + o0, (const char*)&t + o0, o1 - o0, cudaMemcpyHostToDevice));
  return 0;
}

int b200env_bind(b200env_handle h, const b200_buffers_t* bufs) {
  if (!h || !bufs) return fail(-1, "b200env_bind: null argument%s");
  const void* req[] = {bufs->root_states, bufs->dof_state, bufs->rigid_body_state, bufs->contact_forces, bufs->obs_buf, bufs->rew_buf,
                       bufs->sub_rewards, bufs->reset_buf, bufs->progress_buf, bufs->terminate_buf, bufs->motion_ids,
                       bufs->ref_motion_times, bufs->motion_bodies, bufs->t_root_pos, bufs->t_root_rot, bufs->t_dof_pos,
                       bufs->t_root_vel, bufs->t_root_ang_vel, bufs->t_dof_vel, bufs->t_key_pos, bufs->t_rb_pos, bufs->t_rb_rot,
                       bufs->p_dof_pos, bufs->p_dof_vel, bufs->p_rb_pos, bufs->p_rb_rot, bufs->pd_targets, bufs->actions_used};
  for (size_t i = 0; i < sizeof(req) / sizeof(req[0]); i++)
    if (!req[i]) return fail(-1, "b200env_bind: a required buffer pointer is null%s");
  if (bufs->bodies_per_env < h->model.nb || bufs->actors_per_env < 1) return fail(-2, "b200env_bind: bad bodies/actors per env%s");
  if (h->cfg.has_ball) {
    if (bufs->actors_per_env < 2 || bufs->bodies_per_env < h->model.nb + 1) return fail(-2, "b200env_bind: has_ball needs 2 actors and nb+1 rigid-body rows per env%s");
    if (!bufs->has_bounce || !bufs->has_bounce_now || !bufs->bounce_pos || !bufs->racket_hit_now) return fail(-1, "b200env_bind: ball flag buffers are null%s");
  }
  if (((uintptr_t)bufs->t_rb_rot | (uintptr_t)bufs->p_rb_rot) & 15) return fail(-2, "b200env_bind: quaternion rows must be 16-byte aligned%s");
  if (bufs->num_actions != h->model.nd && bufs->num_actions != h->model.nd + 6)
    return fail(-2, "b200env_bind: num_actions must be nd (no residual wrench) or nd + 6%s");
  if (h->cfg.res_force_scale > 0.f && bufs->num_actions != h->model.nd + 6)
    return fail(-2, "b200env_bind: res_force_scale > 0 needs num_actions = nd + 6 (force + torque columns)%s");
  h->bufs = *bufs;
  h->bound = true;
  return 0;
}

int b200env_set_motion_lib(b200env_handle h, const b200_motion_lib_t* ml) {
  if (!h || !ml) return fail(-1, "b200env_set_motion_lib: null argument%s");
  if (!ml->gts || !ml->grs || !ml->lrs || !ml->grvs || !ml->gravs || !ml->dvs || !ml->motion_lengths || !ml->num_frames ||
      !ml->motion_dt || !ml->length_starts || !ml->min_verts_h)
    return fail(-1, "b200env_set_motion_lib: null array%s");
  if (ml->num_lib_bodies < 1 || ml->num_lib_bodies > h->model.nb) return fail(-2, "b200env_set_motion_lib: body count mismatch%s");
  if (((uintptr_t)ml->grs | (uintptr_t)ml->lrs) & 15) return fail(-2, "b200env_set_motion_lib: grs/lrs must be 16-byte aligned%s");
  h->ml = *ml;
  h->has_ml = true;
  return 0;
}

static int need_ctas(const b200env* h) { return (h->num_envs + WARPS_PER_CTA - 1) / WARPS_PER_CTA; }
static size_t step_smem(const b200env* h) { return ((h->blob_bytes + 15) & ~(size_t)15) + WARPS_PER_CTA * SCRATCH_FLOATS * sizeof(float); }

int b200env_step(b200env_handle h, const float* actions, void* stream) {
  if (!h || !actions) return fail(-1, "b200env_step: null argument%s");
  if (!h->bound || (!h->has_ml && h->cfg.task_mode == 0)) return fail(-4, "b200env_step: bind buffers and a motion lib first%s");
  cudaSetDevice(h->device);
  if (h->packed) {
    const size_t psmem = h->tmem ? ((h->blob_bytes + 15) & ~(size_t)15) + (size_t)PT_WARPS * EPW * PT_ENV_STRIDE * sizeof(float)
                                 : ((h->blob_bytes + 15) & ~(size_t)15) + (h->split ? 0 : PK_WARPS * PK_SCRATCH * sizeof(float)) +
                                       (size_t)PK_WARPS * EPW * ENV_STRIDE * sizeof(float);
    const int batch = (h->tmem ? PT_WARPS : PK_WARPS) * EPW;
    const int need = (h->num_envs + batch - 1) / batch;
    if (h->step_grid == 0) {
      if (!h->tmem) {
        CUDA_OK(cudaFuncSetAttribute(step_kernel_packed<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        CUDA_OK(cudaFuncSetAttribute(step_kernel_packed<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
      }
      int per_sm = 0, sms = 0;
      if (h->tmem) {
        CUDA_OK(cudaFuncSetAttribute(step_kernel_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_kernel_tmem, PT_WARPS * 32, psmem));
        if (per_sm > 1) per_sm = 1;   // the CTA takes the SM's whole tensor memory
      } else
#if B200ENV_WITH_PACKED3
      if (h->packed3) {
        CUDA_OK(cudaFuncSetAttribute(step_kernel_packed3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_kernel_packed3, PK3_WARPS * 32, psmem));
      } else
#endif
      if (h->split) { CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_kernel_packed<true>, PK_LAUNCH_WARPS * 32, psmem)); }
      else { CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_kernel_packed<false>, PK_LAUNCH_WARPS * 32, psmem)); }
      CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
      h->step_grid = per_sm * sms < need ? per_sm * sms : need;
      if (h->step_grid < 1) return fail(-5, "b200env_step: step_kernel_packed does not fit on this device%s");
    }
    // the ticket counter is re-zeroed on the stream before every launch: nothing in the launch depends on host-side
    // history, so a step can be captured into a CUDA graph and replayed
    if (!h->split) CUDA_OK(cudaMemsetAsync(h->d_ticket, 0, sizeof(unsigned long long), (cudaStream_t)stream));   // split form: pre_kernel zeroes it
    if (h->split) {
      const int rows = h->env_first + h->env_stride * (h->num_envs - 1) + 1;
      if (rows > h->ext_rows) {  // first step (or a wider env slice): never inside a graph capture - the bound task warms up first
        cudaFree(h->d_ext);
        h->d_ext = nullptr;
        CUDA_OK(cudaMalloc(&h->d_ext, (size_t)rows * 6 * sizeof(float)));
        h->ext_rows = rows;
      }
      const int io_grid = need_ctas(h);
      if (h->sort) {
        CUDA_OK(cudaMemsetAsync(h->d_cnt, 0, sizeof(uint32_t) * SORT_BINS, (cudaStream_t)stream));
        sort_count_kernel<<<io_grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>((const DevBlob*)h->d_blob, h->bufs, h->num_envs, h->env_first,
                                                                                    h->env_stride, h->d_bin, h->d_pos, h->d_cnt);
        sort_perm_kernel<<<(h->num_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->num_envs, h->d_bin, h->d_pos, h->d_cnt, h->d_perm);
        h->launches += 2;
      }
      pre_kernel<<<io_grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>((const DevBlob*)h->d_blob, h->d_cfg, h->bufs, h->ml, actions,
                                                                           h->num_envs, h->env_first, h->env_stride, h->d_ext, h->d_ticket);
      cudaEvent_t tv0 = nullptr, tv1 = nullptr;
      if (h->timing && h->tev && h->tev->size() < 8192) {
        if (cudaEventCreate(&tv0) == cudaSuccess && cudaEventCreate(&tv1) == cudaSuccess) cudaEventRecord(tv0, (cudaStream_t)stream);
        else tv0 = tv1 = nullptr;
      }
      if (h->tmem)
        step_kernel_tmem<<<h->step_grid, PT_WARPS * 32, psmem, (cudaStream_t)stream>>>(
            (const DevBlob*)h->d_blob, (uint32_t)h->blob_bytes, h->d_cfg, h->bufs, h->num_envs, h->d_ticket, h->env_first, h->env_stride, h->d_ext);
      else
#if B200ENV_WITH_PACKED3
      if (h->packed3)
        step_kernel_packed3<<<h->step_grid, PK3_WARPS * 32, psmem, (cudaStream_t)stream>>>(
            (const DevBlob*)h->d_blob, (uint32_t)h->blob_bytes, h->d_cfg, h->bufs, h->num_envs, h->d_ticket, h->env_first, h->env_stride, h->d_ext);
      else
#endif
        step_kernel_packed<true><<<h->step_grid, PK_LAUNCH_WARPS * 32, psmem, (cudaStream_t)stream>>>(
            (const DevBlob*)h->d_blob, (uint32_t)h->blob_bytes, h->d_cfg, h->bufs, h->ml, actions, h->num_envs, h->d_ticket, h->env_first,
            h->env_stride, h->d_ext, h->sort ? h->d_perm : nullptr);
      if (tv0) { cudaEventRecord(tv1, (cudaStream_t)stream); h->tev->push_back(tv0); h->tev->push_back(tv1); }
      if (h->cfg.task_mode == 0)
        post_kernel<<<io_grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>((const DevBlob*)h->d_blob, h->d_cfg, h->bufs, h->ml, h->num_envs,
                                                                              h->env_first, h->env_stride);
      CUDA_OK(cudaGetLastError());
      h->launches += h->cfg.task_mode == 0 ? 3 : 2;
      return 0;
    }
    step_kernel_packed<false><<<h->step_grid, PK_LAUNCH_WARPS * 32, psmem, (cudaStream_t)stream>>>(
        (const DevBlob*)h->d_blob, (uint32_t)h->blob_bytes, h->d_cfg, h->bufs, h->ml, actions, h->num_envs, h->d_ticket, h->env_first,
        h->env_stride, nullptr, nullptr);
    CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
  }
  const size_t smem = step_smem(h);
  if (h->step_grid == 0) {
    CUDA_OK(cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, sms = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_kernel, WARPS_PER_CTA * 32, smem));
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    const int need = need_ctas(h);
    h->step_grid = per_sm * sms < need ? per_sm * sms : need;
    if (h->step_grid < 1) return fail(-5, "b200env_step: step_kernel does not fit on this device%s");
  }
  CUDA_OK(cudaMemsetAsync(h->d_ticket, 0, sizeof(unsigned long long), (cudaStream_t)stream));
  step_kernel<<<h->step_grid, WARPS_PER_CTA * 32, smem, (cudaStream_t)stream>>>((const DevBlob*)h->d_blob, (uint32_t)h->blob_bytes,
                                                                                 h->d_cfg, h->bufs, h->ml, actions, h->num_envs,
                                                                                 h->d_ticket, h->env_first, h->env_stride);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  return 0;
}

int b200env_reset(b200env_handle h, const int64_t* env_ids, const float* motion_times, int32_t n, void* stream) {
  if (!h) return fail(-1, "b200env_reset: null handle%s");
  if (n == 0) return 0;
  if (!env_ids || !motion_times || n < 0) return fail(-1, "b200env_reset: bad arguments%s");
  if (!h->bound || !h->has_ml) return fail(-4, "b200env_reset: bind buffers and a motion lib first%s");
  cudaSetDevice(h->device);
  const int grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  reset_kernel<<<grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>((const DevBlob*)h->d_blob, h->d_cfg, h->bufs, h->ml, env_ids,
                                                                      motion_times, n);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  return 0;
}

int b200env_motion_state(b200env_handle h, const int64_t* motion_ids, const float* motion_times, int32_t n, float* root_pos,
                         float* root_rot, float* dof_pos, float* root_vel, float* root_ang_vel, float* dof_vel, float* key_pos,
                         float* rb_pos, float* rb_rot, void* stream) {
  if (!h) return fail(-1, "b200env_motion_state: null handle%s");
  if (n == 0) return 0;
  if (!motion_ids || !motion_times || n < 0) return fail(-1, "b200env_motion_state: bad arguments%s");
  if (!h->has_ml) return fail(-4, "b200env_motion_state: set a motion lib first%s");
  cudaSetDevice(h->device);
  const int grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  motion_state_kernel<<<grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>((const DevBlob*)h->d_blob, h->d_cfg, h->ml, motion_ids,
                                                                             motion_times, n, root_pos, root_rot, dof_pos, root_vel,
                                                                             root_ang_vel, dof_vel, key_pos, rb_pos, rb_rot);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  return 0;
}

int b200env_motion_context(b200env_handle h, const int64_t* env_ids, const int64_t* motion_ids, const float* motion_times, int32_t n,
                           int32_t num_frames, int32_t first_frame, float dt, float* context_feat, uint8_t* context_mask, void* stream) {
  if (!h) return fail(-1, "b200env_motion_context: null handle%s");
  if (n == 0) return 0;
  if (!motion_ids || !motion_times || !context_feat || !context_mask || n < 0 || num_frames < 1)
    return fail(-1, "b200env_motion_context: bad arguments%s");
  if (!h->has_ml) return fail(-4, "b200env_motion_context: set a motion lib first%s");
  cudaSetDevice(h->device);
  const int64_t warps = (int64_t)n * num_frames;
  const int grid = (int)((warps + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
  motion_context_kernel<<<grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>((const DevBlob*)h->d_blob, h->d_cfg, h->ml, env_ids, motion_ids,
                                                                               motion_times, n, num_frames, first_frame, dt,
                                                                               (float)(2.0 * (double)dt), context_feat, context_mask);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  return 0;
}

int b200env_obs_imitation(b200env_handle h, int32_t n, const float* body_pos, const float* body_rot, const float* target_pos,
                          const float* target_rot, const float* dof_pos, const float* dof_vel, const float* target_dof_pos,
                          const float* body_vel, const float* body_ang_vel, const float* motion_bodies, int32_t local_root_obs,
                          int32_t root_height_obs, float* obs, void* stream) {
  if (!h) return fail(-1, "b200env_obs_imitation: null handle%s");
  if (n == 0) return 0;
  if (!body_pos || !body_rot || !target_pos || !target_rot || !dof_pos || !dof_vel || !target_dof_pos || !body_vel || !body_ang_vel ||
      !motion_bodies || !obs || n < 0)
    return fail(-1, "b200env_obs_imitation: bad arguments%s");
  cudaSetDevice(h->device);
  const int nbl = h->has_ml ? h->ml.num_lib_bodies : (h->model.fixed[h->model.nb - 1] ? h->model.nb - 1 : h->model.nb);
  const int grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  obs_imitation_kernel<<<grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>(n, nbl, h->model.nd, h->cfg.shape_dim, body_pos, body_rot,
                                                                              target_pos, target_rot, dof_pos, dof_vel, target_dof_pos,
                                                                              body_vel, body_ang_vel, motion_bodies, local_root_obs & 1,
                                                                              root_height_obs, obs, (local_root_obs >> 1) & 1,
                                                                              ObsStrides{nbl * 3, 3, nbl * 4, 4, h->model.nd, 1}, ObsBf16{nullptr, 0, nullptr, nullptr, 0.f});
  CUDA_OK(cudaGetLastError());
  h->launches++;
  return 0;
}

int b200env_obs_imitation_rows(b200env_handle h, int32_t n, const float* rigid_body_state, int32_t bodies_per_env, const float* dof_state,
                               const float* target_pos, const float* target_rot, const float* target_dof_pos, const float* motion_bodies,
                               int32_t local_root_obs, int32_t root_height_obs, float* obs, void* obs_bf16, int32_t ld_bf16, const float* mean,
                               const float* rstd, float clamp, void* stream) {
  if (!h || !rigid_body_state || !dof_state || !target_pos || !target_rot || !target_dof_pos || !motion_bodies || !obs)
    return fail(-1, "b200env_obs_imitation_rows: null argument%s");
  if (n <= 0) return fail(-2, "b200env_obs_imitation_rows: n must be positive%s");
  const int nbl = h->has_ml ? h->ml.num_lib_bodies : (h->model.fixed[h->model.nb - 1] ? h->model.nb - 1 : h->model.nb);
  if (bodies_per_env < nbl) return fail(-2, "b200env_obs_imitation_rows: bodies_per_env smaller than the humanoid%s");
  if ((mean == nullptr) != (rstd == nullptr)) return fail(-2, "b200env_obs_imitation_rows: mean and rstd go together%s");
  if (obs_bf16 && !(clamp > 0.f)) return fail(-2, "b200env_obs_imitation_rows: the bf16 row needs a positive clamp%s");
  cudaSetDevice(h->device);
  const int grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  const int nd = h->model.nd;
  obs_imitation_kernel<<<grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>(
      n, nbl, nd, h->cfg.shape_dim, rigid_body_state, rigid_body_state + 3, target_pos, target_rot, dof_state, dof_state + 1, target_dof_pos,
      rigid_body_state + 7, rigid_body_state + 10, motion_bodies, local_root_obs & 1, root_height_obs, obs, (local_root_obs >> 1) & 1,
      ObsStrides{bodies_per_env * 13, 13, bodies_per_env * 13, 13, nd * 2, 2},
      ObsBf16{reinterpret_cast<__nv_bfloat16*>(obs_bf16), ld_bf16, mean, rstd, clamp});
  CUDA_OK(cudaGetLastError());
  h->launches++;
  return 0;
}

int b200env_physics_only(b200env_handle h, int32_t prec, int32_t n, int32_t n_steps, void* root, void* dof_pos, void* dof_vel,
                         const void* pd_tar, const void* ext_wrench, void* rb_out, void* contact_out, void* ball, int32_t* ball_hits,
                         void* stream) {
  if (!h || !root || !dof_pos || !dof_vel || !pd_tar || !rb_out) return fail(-1, "b200env_physics_only: null argument%s");
  if (n <= 0 || n_steps <= 0) return fail(-2, "b200env_physics_only: n and n_steps must be positive%s");
  cudaSetDevice(h->device);
  if (h->packed) {
    const size_t bsm = (h->blob_bytes + 15) & ~(size_t)15;
    if (prec == 0 && h->tmem) {
      constexpr int W = 6;   // more than 4 warps: windows in two column ranges of the tensor memory
      const size_t sm = bsm + (size_t)W * EPW * PT_ENV_STRIDE * sizeof(float);
      CUDA_OK(cudaFuncSetAttribute(physics_kernel_tmem<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      physics_kernel_tmem<W><<<(n + W * EPW - 1) / (W * EPW), W * 32, sm, (cudaStream_t)stream>>>(
          (const DevBlob*)h->d_blob, (uint32_t)h->blob_bytes, h->d_cfg, n, n_steps, (float*)root, (float*)dof_pos, (float*)dof_vel,
          (const float*)pd_tar, (const float*)ext_wrench, (float*)rb_out, (float*)contact_out, (float*)ball, ball_hits);
    } else if (prec == 0) {
      constexpr int W = 4;
      const size_t sm = bsm + (size_t)W * EPW * ENV_STRIDE * sizeof(float);
      CUDA_OK(cudaFuncSetAttribute(physics_kernel_packed<float, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      physics_kernel_packed<float, W><<<(n + W * EPW - 1) / (W * EPW), W * 32, sm, (cudaStream_t)stream>>>(
          (const DevBlob*)h->d_blob, (uint32_t)h->blob_bytes, h->d_cfg, n, n_steps, (float*)root, (float*)dof_pos, (float*)dof_vel,
          (const float*)pd_tar, (const float*)ext_wrench, (float*)rb_out, (float*)contact_out, (float*)ball, ball_hits);
    } else if (prec == 1) {
      constexpr int W = 2;
      const size_t sm = bsm + (size_t)W * EPW * ENV_STRIDE * sizeof(double);
      CUDA_OK(cudaFuncSetAttribute(physics_kernel_packed<double, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      physics_kernel_packed<double, W><<<(n + W * EPW - 1) / (W * EPW), W * 32, sm, (cudaStream_t)stream>>>(
          (const DevBlob*)h->d_blob, (uint32_t)h->blob_bytes, h->d_cfg, n, n_steps, (double*)root, (double*)dof_pos, (double*)dof_vel,
          (const double*)pd_tar, (const double*)ext_wrench, (double*)rb_out, (double*)contact_out, (double*)ball, ball_hits);
    } else {
      return fail(-2, "b200env_physics_only: prec must be 0 (float) or 1 (double)%s");
    }
    CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
  }
  const int grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  const size_t smem = (h->blob_bytes + 15) & ~(size_t)15;
  if (prec == 0) {
    CUDA_OK(cudaFuncSetAttribute(physics_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    physics_kernel<float><<<grid, WARPS_PER_CTA * 32, smem, (cudaStream_t)stream>>>(
        (const DevBlob*)h->d_blob, (uint32_t)h->blob_bytes, h->d_cfg, n, n_steps, (float*)root, (float*)dof_pos, (float*)dof_vel,
        (const float*)pd_tar, (const float*)ext_wrench, (float*)rb_out, (float*)contact_out, (float*)ball, ball_hits);
  } else if (prec == 1) {
    CUDA_OK(cudaFuncSetAttribute(physics_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    physics_kernel<double><<<grid, WARPS_PER_CTA * 32, smem, (cudaStream_t)stream>>>(
        (const DevBlob*)h->d_blob, (uint32_t)h->blob_bytes, h->d_cfg, n, n_steps, (double*)root, (double*)dof_pos, (double*)dof_vel,
        (const double*)pd_tar, (const double*)ext_wrench, (double*)rb_out, (double*)contact_out, (double*)ball, ball_hits);
  } else {
    return fail(-2, "b200env_physics_only: prec must be 0 (float) or 1 (double)%s");
  }
  CUDA_OK(cudaGetLastError());
  h->launches++;
  return 0;
}

#if PT_PROF
// debug builds only (tools/pt_prof.sh): the per-warp phase cycle counts of the last step_kernel_tmem launch
extern "C" int b200env_debug_prof(unsigned long long* out, int n) {
  return cudaMemcpyFromSymbol(out, g_pt_prof, sizeof(unsigned long long) * (size_t)n) == cudaSuccess ? 0 : -1;
}
#endif
int64_t b200env_launch_count(b200env_handle h) { return h ? h->launches : 0; }
int32_t b200env_kernel_form(b200env_handle h) { return !h ? -1 : (h->tmem ? 3 : (h->packed3 ? 2 : (h->packed ? 1 : 0))); }

int b200env_set_kernel_timing(b200env_handle h, int32_t on) {
  if (!h) return fail(-1, "b200env_set_kernel_timing: null handle%s");
  if (!h->tev) h->tev = new std::vector<cudaEvent_t>();
  for (cudaEvent_t e : *h->tev) cudaEventDestroy(e);
  h->tev->clear();
  h->timing = on ? 1 : 0;
  return 0;
}

int b200env_kernel_ms(b200env_handle h, double* mean_ms, int32_t* count) {
  if (!h || !mean_ms || !count) return fail(-1, "b200env_kernel_ms: null argument%s");
  *mean_ms = 0.0;
  *count = 0;
  if (!h->tev || h->tev->empty()) return 0;
  cudaSetDevice(h->device);
  double sum = 0.0;
  int n = 0;
  for (size_t i = 0; i + 1 < h->tev->size(); i += 2) {
    float ms = 0.f;
    if (cudaEventSynchronize((*h->tev)[i + 1]) == cudaSuccess && cudaEventElapsedTime(&ms, (*h->tev)[i], (*h->tev)[i + 1]) == cudaSuccess) { sum += ms; n++; }
  }
  for (cudaEvent_t e : *h->tev) cudaEventDestroy(e);
  h->tev->clear();
  cudaGetLastError();
  if (n) *mean_ms = sum / n;
  *count = n;
  return 0;
}

int b200env_set_env_slice(b200env_handle h, int32_t env_first, int32_t env_stride) {
  if (!h) return fail(-1, "b200env_set_env_slice: null handle%s");
  if (env_first < 0 || env_stride < 1) return fail(-2, "b200env_set_env_slice: env_first >= 0 and env_stride >= 1 required%s");
  h->env_first = env_first;
  h->env_stride = env_stride;
  return 0;
}

}  // extern "C"

#include "ballgen.cuh"
