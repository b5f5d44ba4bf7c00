// b200env_v2p.cu - kernels + C ABI for the vid2player rows of the hot path:
// SMPL-FK targets, ball aerodynamics / reset, state-from-sim, and the high-level env's
// bounce/estimator bookkeeping + reward + observation + reset FSM fused into one launch.
// Reference (paths relative to /root/reference/vid2player):
//   _smpl_to_sim / _forward_kinematics        env/tasks/humanoid_smpl_im_mvae.py:897-946, utils/hybrik.py:597-652
//   rotation_matrix_to_quaternion/_angle_axis utils/konia_transform.py:348-438, 558-655
//   apply_external_force_to_ball              env/tasks/humanoid_smpl_im_mvae.py:711-739 (+ utils/tennis_ball.py:15-37)
//   _reset_balls + offline pool               env/tasks/humanoid_smpl_im_mvae.py:503-524, utils/tennis_ball.py:422-456
//   _update_state_from_sim                    env/tasks/humanoid_smpl_im_mvae.py:799-860
//   controller post_physics_step              env/tasks/physics_mvae_controller.py:271-314,333-360,368-436,481-602
//   TennisBallOutEstimator.estimate           utils/tennis_ball_out_estimator.py:126-205
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200env_v2p.h"

#define FULL 0xffffffffu
#define V2P_WARPS 4

static thread_local char g_verr[256] = "";
static int vfail(int code, const char* msg) {
  snprintf(g_verr, sizeof(g_verr), "%s", msg);
  return code;
}
#define V_CUDA_OK()                                                        \
  do {                                                                     \
    cudaError_t _e = cudaGetLastError();                                   \
    if (_e != cudaSuccess) return vfail(-10, cudaGetErrorString(_e));      \
  } while (0)

// ------------------------------------------------------------------------------------------ konia conversions
__device__ __forceinline__ float safe_div(float num, float den) {  // safe_zero_division, eps 1e-6
  if (fabsf(den) < 1e-6f) den += 1e-6f;
  return num / den;
}
// rotation matrix (row-major) -> quaternion wxyz  (konia_transform.py:348-438)
__device__ __forceinline__ void rotmat_to_quat_wxyz(const float* m, float* q) {
  const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
  const float trace = m00 + m11 + m22;
  if (trace > 0.0f) {
    float sq = sqrtf(fmaxf(trace + 1.0f, 1e-6f)) * 2.0f;
    q[0] = 0.25f * sq; q[1] = safe_div(m21 - m12, sq); q[2] = safe_div(m02 - m20, sq); q[3] = safe_div(m10 - m01, sq);
  } else if ((m00 > m11) && (m00 > m22)) {
    float sq = sqrtf(fmaxf(1.0f + m00 - m11 - m22, 1e-6f)) * 2.0f;
    q[0] = safe_div(m21 - m12, sq); q[1] = 0.25f * sq; q[2] = safe_div(m01 + m10, sq); q[3] = safe_div(m02 + m20, sq);
  } else if (m11 > m22) {
    float sq = sqrtf(fmaxf(1.0f + m11 - m00 - m22, 1e-6f)) * 2.0f;
    q[0] = safe_div(m02 - m20, sq); q[1] = safe_div(m01 + m10, sq); q[2] = 0.25f * sq; q[3] = safe_div(m12 + m21, sq);
  } else {
    float sq = sqrtf(fmaxf(1.0f + m22 - m00 - m11, 1e-6f)) * 2.0f;
    q[0] = safe_div(m10 - m01, sq); q[1] = safe_div(m02 + m20, sq); q[2] = safe_div(m12 + m21, sq); q[3] = 0.25f * sq;
  }
}
__device__ __forceinline__ float safe_atan2(float y, float x) {  // torch_safe_atan2 (:41-49)
  if (fabsf(y) < 1e-6f && fabsf(x) < 1e-6f) y += 1e-6f;
  return atan2f(y, x);
}
// quaternion wxyz -> rotation vector (konia_transform.py:558-628)
__device__ __forceinline__ void quat_wxyz_to_angle_axis(const float* q, float* aa) {
  const float c = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  const float s2 = q1 * q1 + q2 * q2 + q3 * q3;
  const float s = sqrtf(fmaxf(s2, 1e-6f));
  const float two_theta = 2.0f * (c < 0.0f ? safe_atan2(-s, -c) : safe_atan2(s, c));
  const float k = s2 > 0.0f ? safe_div(two_theta, s) : 2.0f;
  aa[0] = q1 * k; aa[1] = q2 * k; aa[2] = q3 * k;
}
// quaternion (given in the converter's WXYZ slot order) -> rotation matrix (konia_transform.py:474-555)
__device__ __forceinline__ void quat_wxyz_to_rotmat(const float* qi, float* m) {
  const float n = fmaxf(sqrtf(qi[0] * qi[0] + qi[1] * qi[1] + qi[2] * qi[2] + qi[3] * qi[3]), 1e-12f);
  const float w = qi[0] / n, x = qi[1] / n, y = qi[2] / n, z = qi[3] / n;
  const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  m[0] = 1.0f - (tyy + tzz); m[1] = txy - twz; m[2] = txz + twy;
  m[3] = txy + twz; m[4] = 1.0f - (txx + tzz); m[5] = tyz - twx;
  m[6] = txz - twy; m[7] = tyz + twx; m[8] = 1.0f - (txx + tyy);
}
__device__ __forceinline__ void qmul_xyzw(const float* a, const float* b, float* o) {
  const float x1 = a[0], y1 = a[1], z1 = a[2], w1 = a[3], x2 = b[0], y2 = b[1], z2 = b[2], w2 = b[3];
  o[0] = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2;
  o[1] = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2;
  o[2] = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2;
  o[3] = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2;
}

// ------------------------------------------------------------------------------------------ a13: SMPL FK targets
__global__ void __launch_bounds__(V2P_WARPS * 32)
smpl_to_sim_kernel(int n, const float* __restrict__ root_pos, const float* __restrict__ rotmat, const float* __restrict__ rest_all,
                   int num_rest, const int32_t* __restrict__ parents, const int32_t* __restrict__ smpl_2_mujoco, float dt,
                   const float* prev_root_pos, const float* __restrict__ prev_rb_rot, float* root_rot, float* dof_pos,
                   float* root_vel, float* root_ang_vel, float* dof_vel, float* rb_pos, float* rb_rot, float* prev_root_pos_update,
                   float* target_root_pos_out, const uint8_t* __restrict__ only_mask) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * V2P_WARPS + warp;
  if (e >= n) return;
  if (only_mask && !only_mask[e]) return;      // mask-driven reset: FK of the envs being reset only
  const float* rest = rest_all + (e % num_rest) * 72;   // per-shape rest joints: env e has shape e % num_rest (dual: players alternate)
  const bool act = lane < 24;
  const int j = act ? lane : 0;
  const int par = parents[j] < 0 ? 0 : parents[j];
  int mj = 0;  // mujoco index of SMPL joint j (inverse of smpl_2_mujoco)
  for (int k = 0; k < 24; k++)
    if (smpl_2_mujoco[k] == j) mj = k;
  float Rl[9], G[9], P[3], rel[3];
#pragma unroll
  for (int k = 0; k < 9; k++) Rl[k] = rotmat[(e * 24 + j) * 9 + k];
#pragma unroll
  for (int k = 0; k < 3; k++) rel[k] = rest[j * 3 + k] - (j > 0 ? rest[par * 3 + k] : 0.0f);
  // local angle-axis -> dof targets (root excluded), mujoco order
  if (act && mj > 0) {
    float q[4], aa[3];
    rotmat_to_quat_wxyz(Rl, q);
    quat_wxyz_to_angle_axis(q, aa);
#pragma unroll
    for (int k = 0; k < 3; k++) dof_pos[e * 69 + (mj - 1) * 3 + k] = aa[k];
  }
  // FK chain (hybrik.py:597-652): G_j = G_parent R_j ; P_j = G_parent rel_j + P_parent
  bool done = (lane == 0);
#pragma unroll
  for (int k = 0; k < 9; k++) G[k] = Rl[k];
#pragma unroll
  for (int k = 0; k < 3; k++) P[k] = rel[k];
  for (int it = 0; it < 9; it++) {
    float pG[9], pP[3];
#pragma unroll
    for (int k = 0; k < 9; k++) pG[k] = __shfl_sync(FULL, G[k], par);
#pragma unroll
    for (int k = 0; k < 3; k++) pP[k] = __shfl_sync(FULL, P[k], par);
    const bool pdone = __shfl_sync(FULL, (int)done, par) != 0;
    if (act && !done && pdone) {
#pragma unroll
      for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 3; c++) G[r * 3 + c] = pG[r * 3] * Rl[c] + pG[r * 3 + 1] * Rl[3 + c] + pG[r * 3 + 2] * Rl[6 + c];
        P[r] = pG[r * 3] * rel[0] + pG[r * 3 + 1] * rel[1] + pG[r * 3 + 2] * rel[2] + pP[r];
      }
      done = true;
    }
  }
  float qw[4], q[4];
  rotmat_to_quat_wxyz(G, qw);
  q[0] = qw[1]; q[1] = qw[2]; q[2] = qw[3]; q[3] = qw[0];
  if (act) {
#pragma unroll
    for (int k = 0; k < 3; k++) rb_pos[(e * 24 + mj) * 3 + k] = P[k] - (rest[k] - root_pos[e * 3 + k]);
#pragma unroll
    for (int k = 0; k < 4; k++) rb_rot[(e * 24 + mj) * 4 + k] = q[k];
    if (mj == 0) {
#pragma unroll
      for (int k = 0; k < 4; k++) root_rot[e * 4 + k] = q[k];
    }
    if (prev_root_pos && prev_rb_rot) {
      // diff = quat_normalize(quat_mul(conj(prev), cur)) ; axis*angle/dt   (:909-919)
      const float* pq = prev_rb_rot + (e * 24 + mj) * 4;
      float cj[4] = {-pq[0], -pq[1], -pq[2], pq[3]}, d[4];
      qmul_xyzw(cj, q, d);
      const float sg = d[3] < 0.0f ? -1.0f : 1.0f;  // quat_pos
#pragma unroll
      for (int k = 0; k < 4; k++) d[k] *= sg;
      const float nn = fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3]), 1e-9f);
#pragma unroll
      for (int k = 0; k < 4; k++) d[k] /= nn;
      // quat_to_angle_axis (utils/torch_utils.py:82-102)
      float st = sqrtf(__fsub_rn(1.0f, __fmul_rn(d[3], d[3])));
      float ang = 2.0f * acosf(d[3]);
      ang = atan2f(sinf(ang), cosf(ang));
      float ax[3];
      if (fabsf(st) > 1e-5f) { ax[0] = d[0] / st; ax[1] = d[1] / st; ax[2] = d[2] / st; }
      else { ang = 0.0f; ax[0] = 0.0f; ax[1] = 0.0f; ax[2] = 1.0f; }
      float dv[3] = {ax[0] * ang / dt, ax[1] * ang / dt, ax[2] * ang / dt};
      if (mj == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
          root_ang_vel[e * 3 + k] = dv[k] / dt;  // reference quirk: divided by dt twice (:917-919)
          root_vel[e * 3 + k] = (root_pos[e * 3 + k] - prev_root_pos[e * 3 + k]) / dt;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 3; k++) dof_vel[e * 69 + (mj - 1) * 3 + k] = dv[k];
      }
    } else {
      if (mj == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { root_ang_vel[e * 3 + k] = 0.0f; root_vel[e * 3 + k] = 0.0f; }
      } else {
#pragma unroll
        for (int k = 0; k < 3; k++) dof_vel[e * 69 + (mj - 1) * 3 + k] = 0.0f;
      }
    }
    // _save_prev_target_motion_state (:741-750) / `self._target_root_pos = root_pos` folded in: the lane that just read the previous
    // root position stores the current one (prev_root_pos_update may alias prev_root_pos)
    if (mj == 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float r = root_pos[e * 3 + k];
        if (prev_root_pos_update) prev_root_pos_update[e * 3 + k] = r;
        if (target_root_pos_out) target_root_pos_out[e * 3 + k] = r;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ a13b: head look-at correction
// angle_axis_to_rotation_matrix (konia_transform.py:250-334)
__device__ __forceinline__ void aa_to_rotmat(const float* aa, float* m) {
  const float theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > 1e-6f) {
    const float theta = sqrtf(fmaxf(theta2, 1e-6f));
    const float wx = aa[0] / (theta + 1e-6f), wy = aa[1] / (theta + 1e-6f), wz = aa[2] / (theta + 1e-6f);
    const float c = cosf(theta), s = sinf(theta), k = 1.0f - c;
    m[0] = c + wx * wx * k; m[1] = wx * wy * k - wz * s; m[2] = wy * s + wx * wz * k;
    m[3] = wz * s + wx * wy * k; m[4] = c + wy * wy * k; m[5] = -wx * s + wy * wz * k;
    m[6] = -wy * s + wx * wz * k; m[7] = wx * s + wy * wz * k; m[8] = c + wz * wz * k;
  } else {
    m[0] = 1.0f; m[1] = -aa[2]; m[2] = aa[1]; m[3] = aa[2]; m[4] = 1.0f; m[5] = -aa[0]; m[6] = -aa[1]; m[7] = aa[0]; m[8] = 1.0f;
  }
}
__global__ void fix_head_kernel(int n, const float* __restrict__ rb_pos, const float* __restrict__ rb_rot, int head_body,
                                const float* __restrict__ ball_pos, const float* __restrict__ root_pos, float* joint_rotmat) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float* hq = rb_rot + (e * 24 + head_body) * 4;
  const float qw[4] = {hq[3], hq[0], hq[1], hq[2]};
  float m[9];
  quat_wxyz_to_rotmat(qw, m);
  float lx = m[2], ly = m[5];   // head_rotmat @ (0,0,1), xy part
  float ln = fmaxf(sqrtf(lx * lx + ly * ly), 1e-12f);
  lx /= ln; ly /= ln;
  float bx = ball_pos[e * 3] - rb_pos[(e * 24 + head_body) * 3], by = ball_pos[e * 3 + 1] - rb_pos[(e * 24 + head_body) * 3 + 1];
  float bn = fmaxf(sqrtf(bx * bx + by * by), 1e-12f);
  bx /= bn; by /= bn;
  float d = atan2f(by, bx) - atan2f(ly, lx);
  if (d > 3.14159265358979323846f) d -= 6.283185307179586f;
  if (d < -3.14159265358979323846f) d += 6.283185307179586f;
  const bool miss = (ball_pos[e * 3 + 1] < root_pos[e * 3 + 1] - 0.5f) || (fabsf(ball_pos[e * 3]) > 4.0f);
  if (miss) d = 0.0f;
  const int joints[2] = {15, 12};  // SMPLPose.Head, SMPLPose.Neck
#pragma unroll
  for (int j = 0; j < 2; j++) {
    float* R = joint_rotmat + (e * 24 + joints[j]) * 9;
    float Rl[9], q[4], aa[3], o[9];
#pragma unroll
    for (int k = 0; k < 9; k++) Rl[k] = R[k];
    rotmat_to_quat_wxyz(Rl, q);
    quat_wxyz_to_angle_axis(q, aa);
    aa[1] += d / 2.0f;
    aa_to_rotmat(aa, o);
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = o[k];
  }
}

// ------------------------------------------------------------------------------------------ a10: ball aerodynamics
#define BALL_R 0.032f
#define BALL_KF 0.0019462794807519486f  // rho*pi*R^2/2, rho = 1.21 (tennis_ball.py:17-22)
#define BALL_CD 0.55f
__device__ __forceinline__ void ball_aero_force(const float* vel, const float* angvel, float spin_scale, float* force) {
  float vs = sqrtf(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
  if (vs == 0.0f) vs += 1.0f;
  const float vn[3] = {vel[0] / vs, vel[1] / vs, vel[2] / vs};
  // vel_tan = vn x (0,0,-1)
  const float vt[3] = {-vn[1], vn[0], 0.0f};
  const float vspin = sqrtf(angvel[0] * angvel[0] + angvel[1] * angvel[1] + angvel[2] * angvel[2]) / 6.283185307179586f;
  float cl = 1.0f / (2.0f + fabsf(vs / (vspin * spin_scale + 1e-6f)));
  cl = cl * (vspin > 0.0f ? -1.0f : 1.0f);
  // vel_tan x vn
  const float cx = vt[1] * vn[2] - vt[2] * vn[1], cy = vt[2] * vn[0] - vt[0] * vn[2], cz = vt[0] * vn[1] - vt[1] * vn[0];
  const float kd = -BALL_KF * BALL_CD * vs, kl = -BALL_KF * cl * vs * vs;
  force[0] = kd * vel[0] + kl * cx;
  force[1] = kd * vel[1] + kl * cy;
  force[2] = kd * vel[2] + kl * cz;
}
__global__ void ball_aero_kernel(int n, const float* __restrict__ ball_states, int stride, uint8_t* has_bounce, uint8_t* has_bounce_now,
                                 float* bounce_pos, float* force, int substeps, float spin_scale) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float* b = ball_states + e * stride;
  float f[3];
  ball_aero_force(b + 7, b + 10, spin_scale, f);
  force[e * 3] = f[0]; force[e * 3 + 1] = f[1]; force[e * 3 + 2] = f[2];
  const float thr = substeps > 2 ? BALL_R * 6.0f : BALL_R * 4.0f;
  const bool now = !has_bounce[e] && (b[2] <= thr);
  if (now) {
    has_bounce_now[e] = 1; has_bounce[e] = 1;
    bounce_pos[e * 3] = b[0]; bounce_pos[e * 3 + 1] = b[1]; bounce_pos[e * 3 + 2] = b[2];
  }
}

// ------------------------------------------------------------------------------------------ a17: ball reset from the pool
__global__ void ball_reset_kernel(int n, const int64_t* __restrict__ env_ids, const int64_t* __restrict__ pool_index,
                                  const float* __restrict__ pool, float* ball_states, int stride, float* ball_pos, float* ball_vel,
                                  uint8_t* has_bounce, float* bounce_pos, uint8_t* has_contact, float* traj) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int64_t e = env_ids[i];
  const float* row = pool + pool_index[i] * 307;
  for (int k = threadIdx.x; k < 300; k += blockDim.x) traj[e * 300 + k] = row[7 + k];
  if (threadIdx.x == 0) {
    float* b = ball_states + e * stride;
    const float v[3] = {row[3], row[4], row[5]};
    // ang vel = vspin * 2pi * normalize(v x (0,0,-1))   (:508-509)
    float c[3] = {-v[1], v[0], 0.0f};
    const float nn = fmaxf(sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]), 1e-12f);
    const float s = row[6] * 3.141592653589793f * 2.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) { b[k] = row[k]; b[7 + k] = v[k]; b[10 + k] = s * (c[k] / nn); ball_pos[e * 3 + k] = row[k]; ball_vel[e * 3 + k] = v[k]; bounce_pos[e * 3 + k] = 0.0f; }
    has_bounce[e] = 0; has_contact[e] = 0;
  }
}

// ------------------------------------------------------------------------------------------ a11 / a12: state from sim
__global__ void update_state_kernel(b200v2p_state_t s) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= s.n) return;
  if (s.only_mask && !s.only_mask[e]) return;     // reset path: the reference refreshes the state views when humanoids were reset (:186-187)
  const float* rb = s.rigid_body_state + e * s.bodies_per_env * 13;
  const float* ball = s.ball_states + e * s.ball_stride;
  // velocity-jump contact detector (substeps > 2): uses the ball velocity of the previous control step (:800-808)
  const bool now = !s.has_contact[e] && (ball[8] > 0.0f) && ((ball[8] - s.ball_vel[e * 3 + 1]) > 10.0f);
  s.has_contact_now[e] = now ? 1 : 0;
  if (now) s.has_contact[e] = 1;
  const bool second = s.dual && (e & 1);
  const int wrist_body = second ? s.wrist_body2 : s.wrist_body, racket_body = second ? s.racket_body2 : s.racket_body;
  const float* grip = second ? s.grip_normal2 : s.grip_normal;
  const float* wq = rb + wrist_body * 13 + 3;
  const float qw[4] = {wq[3], wq[0], wq[1], wq[2]};
  float m[9];
  quat_wxyz_to_rotmat(qw, m);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    s.root_pos[e * 3 + k] = rb[k];
    s.root_vel[e * 3 + k] = s.root_states[e * s.root_stride + 7 + k];
    s.racket_pos[e * 3 + k] = rb[racket_body * 13 + k];
    s.racket_vel[e * 3 + k] = rb[racket_body * 13 + 7 + k];
    s.racket_normal[e * 3 + k] = m[k * 3] * grip[0] + m[k * 3 + 1] * grip[1] + m[k * 3 + 2] * grip[2];
    s.ball_pos[e * 3 + k] = ball[k];
    s.ball_vel[e * 3 + k] = ball[7 + k];
  }
  s.ball_vspin[e] = sqrtf(ball[10] * ball[10] + ball[11] * ball[11] + ball[12] * ball[12]) / 6.283185307179586f;
}

// ------------------------------------------------------------------------------------------ a14-a16: controller post step
__device__ __forceinline__ bool in_court(float x, float y) { return x > -4.11f && x < 4.11f && y > 0.0f && y < 11.89f; }
__device__ __forceinline__ float est_index(float v, const float* r) {  // clamp + round (:126-162)
  v = fminf(fmaxf(v, r[0]), r[1] - r[2]);
  return rintf((v - r[0]) / r[2]);
}
__global__ void __launch_bounds__(V2P_WARPS * 32) controller_post_kernel(b200v2p_ctrl_t c) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * V2P_WARPS + warp;
  if (e >= c.n) return;
  const bool touched = c.touch_mask && c.touch_mask[e];
  if (c.obs_only && c.touch_mask && !touched && !c.reset_reaction[e] && !c.reset_recovery[e]) return;   // row still current
  const float* rb = c.rigid_body_state + e * c.bodies_per_env * 13;
  if (c.advance && !c.obs_only) {
    // tail of physics_step (:364-366) + head of post_physics_step (:441-444) folded in: the future-trajectory window moves on by one
    // frame (roll(-1), last frame zeroed), the reaction timer and the episode progress count one step.  In place inside the warp.
    float* tr = c.ball_traj + e * 300;
    float tmp[10];
#pragma unroll
    for (int i = 0; i < 10; i++) {
      const int k = lane + 32 * i;
      tmp[i] = k < 297 ? tr[k + 3] : 0.0f;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 10; i++) {
      const int k = lane + 32 * i;
      if (k < 300) tr[k] = tmp[i];
    }
    if (lane == 0) { c.tar_time[e] += 1; c.progress_buf[e] += 1; }
    __syncwarp();
  }
  const float rp[3] = {c.root_pos[e * 3], c.root_pos[e * 3 + 1], c.root_pos[e * 3 + 2]};
  const float bp[3] = {c.ball_pos[e * 3], c.ball_pos[e * 3 + 1], c.ball_pos[e * 3 + 2]};
  const float kp[3] = {c.racket_pos[e * 3], c.racket_pos[e * 3 + 1], c.racket_pos[e * 3 + 2]};
  const bool has_contact = c.has_contact[e] != 0;
  // ---- _update_state (:271-314): true bounce in court, estimated bounce of the outgoing ball
  if (lane == 0 && !c.obs_only) {
    if (c.tar_action[e] == 0 && c.has_bounce_now[e]) c.bounce_in[e] = in_court(c.bounce_pos[e * 3], c.bounce_pos[e * 3 + 1]) ? 1 : 0;
    if (c.has_contact_now[e] && c.est_x) {
      const float* b = c.ball_states + e * c.ball_stride;
      const float *VX = c.est_params, *VY = c.est_params + 3, *VS = c.est_params + 6, *TX = c.est_params + 9, *TY = c.est_params + 12;
      bool valid = (b[8] > VX[0]) && (b[9] > VY[0]) && (b[9] < VY[1]) && (b[2] < TY[1]);
      const float x_net = b[0] + b[7] * fabsf(b[1] / b[8]);
      valid = valid && (x_net > -4.0f) && (x_net < 4.0f);
      if (valid) {
        const float vel_x = sqrtf(b[7] * b[7] + b[8] * b[8]);
        const float vspin = sqrtf(b[10] * b[10] + b[11] * b[11] + b[12] * b[12]) / 6.283185307179586f;
        const float d1 = (VY[1] - VY[0]) / VY[2], d2 = (VS[1] - VS[0]) / VS[2];
        const int64_t ti = (int64_t)(est_index(vel_x, VX) * d1 * d2 + est_index(b[9], VY) * d2 + est_index(vspin, VS));
        const float* tx = c.est_x + ti * c.est_nx;
        const float* ty = c.est_y + ti * c.est_ny * 2;
        const int hi = (int)est_index(b[2], TY);
        float bx = b[0] + ty[hi * 2] * b[7] / vel_x, by = b[1] + ty[hi * 2] * b[8] / vel_x, bt = ty[hi * 2 + 1];
        const float net_dist = -b[1] / b[8] * vel_x;
        const int ni = (int)est_index(net_dist, TX);
        if (tx[ni] + b[2] < 1.07f) { bx = 0.0f; by = 0.0f; bt = 0.0f; }  // into the net (NET_HEIGHT)
        float mx = tx[0];
        for (int k = 1; k < c.est_nx; k++) mx = fmaxf(mx, tx[k]);
        c.est_bounce_pos[e * 3] = bx; c.est_bounce_pos[e * 3 + 1] = by;
        c.est_bounce_time[e] = bt;
        c.est_max_height[e] = b[2] + mx;
        c.est_bounce_in[e] = in_court(bx, by) ? 1 : 0;
      }
    }
  }
  __syncwarp();
  // ---- reward (:368-406, jit :493-602)
  if (lane == 0 && !c.obs_only) {
    const float phase = c.phase[e];
    float d2 = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) d2 += (bp[k] - kp[k]) * (bp[k] - kp[k]);
    float rew = 0.0f, s0 = 0.0f, s1 = 0.0f;
    const float tp[3] = {c.target_bounce_pos[e * 3], c.target_bounce_pos[e * 3 + 1], c.target_bounce_pos[e * 3 + 2]};
    if (c.reward_type == 0) {  // reach
      const float cp = c.swing_type[e] == -1 ? 3.0f : 3.14159265358979323846f;
      s0 = (c.tar_action[e] == 1 ? 1.0f : 0.0f) * expf(-c.scale_pos * d2) * expf(-c.scale_phase * (phase - cp) * (phase - cp));
      rew = s0 * c.w_pos;
    } else {
      const int64_t st = c.reward_type == 1 ? c.swing_type[e] : c.swing_type_cycle[e];
      const float cp = st >= 2 ? 3.0f : 3.14159265358979323846f;
      s0 = (has_contact ? 0.0f : 1.0f) * expf(-c.scale_pos * d2) * expf(-c.scale_phase * (phase - cp) * (phase - cp)) + (has_contact ? 1.0f : 0.0f);
      if (c.reward_type == 1) {  // return
        float err = 0.0f;
        const float* src = c.has_bounce[e] ? (c.bounce_pos + e * 3) : bp;
#pragma unroll
        for (int k = 0; k < 3; k++) err += (src[k] - tp[k]) * (src[k] - tp[k]);
        s1 = (has_contact ? 1.0f : 0.0f) * fminf(fmaxf((400.0f - err) / 400.0f, 0.0f), 1.0f);
      } else {  // return_w_estimate
        float err = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; k++) err += (c.est_bounce_pos[e * 3 + k] - tp[k]) * (c.est_bounce_pos[e * 3 + k] - tp[k]);
        s1 = (c.est_bounce_in[e] ? 1.0f : 0.0f) * expf(-c.scale_bounce_pos * err) * expf(-c.scale_bounce_time * c.est_bounce_time[e]);
      }
      rew = c.w_pos * s0 + c.w_ball_pos * s1;
    }
    c.rew_buf[e] = rew;
    c.sub_rewards[e * 2] = s0; c.sub_rewards[e * 2 + 1] = s1;
  }
  // ---- observation (:316-360): actor 225 | ball trajectory 3*L relative to the racket | target - root xy
  float* o = c.obs_buf + e * c.num_obs;
  bool nan = false;
  if (lane < 3) { o[lane] = rp[lane]; o[3 + lane] = c.root_vel[e * 3 + lane]; o[222 + lane] = c.racket_normal[e * 3 + lane]; nan |= isnan(o[lane]) || isnan(o[3 + lane]) || isnan(o[222 + lane]); }
  if (lane < 24) {
    const float* b1 = rb + (lane + 1) * 13;  // bodies 1..24 (23 humanoid + Racket) relative to the root
#pragma unroll
    for (int k = 0; k < 3; k++) { float v = b1[k] - rp[k]; o[6 + lane * 3 + k] = v; nan |= isnan(v); }
    const float* q = rb + lane * 13 + 3;  // xyzw fed to the WXYZ converter as is (reference quirk, :339)
    float m[9];
    quat_wxyz_to_rotmat(q, m);
    float* r6 = o + 78 + lane * 6;  // rotmat_to_rot6d: cat(mat[...,0], mat[...,1]) = first and second COLUMN
    r6[0] = m[0]; r6[1] = m[3]; r6[2] = m[6]; r6[3] = m[1]; r6[4] = m[4]; r6[5] = m[7];
#pragma unroll
    for (int k = 0; k < 6; k++) nan |= isnan(r6[k]);
  }
  const float* rk = rb + c.racket_body * 13;
  const int L3 = c.obs_traj_len * 3;
  if (c.ball_obs) {   // _compute_task_obs :345-346: _ball_obs <- roll(-1) with the current ball position appended (kept in every mode)
    float* hst = c.ball_obs + e * L3;
    const bool touch = !c.obs_only || c.reset_reaction[e] || (!c.dual && c.reset_recovery[e]);   // obs_only: the reference refreshes
    if (touch) {                                                                                // the reset ids only (:200-201; dual :64-66)
      float tmp[10];   // L3 <= 300 -> at most 10 values per lane: read everything, then write (in-place shift inside the warp)
#pragma unroll
      for (int i = 0; i < 10; i++) {
        const int k = lane + 32 * i;
        tmp[i] = k < L3 - 3 ? hst[k + 3] : (k < L3 ? c.ball_pos[e * 3 + k - (L3 - 3)] : 0.0f);
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 10; i++) {
        const int k = lane + 32 * i;
        if (k < L3) hst[k] = tmp[i];
      }
      __syncwarp();
    }
  }
  const float* src = c.use_history ? c.ball_obs + e * L3 : c.ball_traj + e * 300;
  for (int k = lane; k < L3; k += 32) { float v = src[k] - rk[k % 3]; o[225 + k] = v; nan |= isnan(v); }
  if (c.use_target && lane < 2) { float v = c.target_bounce_pos[e * 3 + lane] - rp[lane]; o[225 + c.obs_traj_len * 3 + lane] = v; nan |= isnan(v); }
  const bool has_nan = __any_sync(FULL, nan);
  // ---- dual reset FSM (physics_mvae_controller_dual.py:92-120): the opponent's flags are recomputed from its inputs (no exchange)
  if (lane == 0 && !c.obs_only && c.dual) {
    const int64_t p = e ^ 1;
    bool rec[2], term[2];
    for (int w = 0; w < 2; w++) {
      const int64_t i = w ? p : e;
      const bool ta1 = c.tar_action[i] == 1, ta0 = c.tar_action[i] == 0, hc = c.has_contact[i] != 0, hb = c.has_bounce[i] != 0;
      // bounce_in AFTER this step's _update_state: recomputed when it is being rewritten (by this or the partner's warp)
      const bool bin = (ta0 && c.has_bounce_now[i]) ? in_court(c.bounce_pos[i * 3], c.bounce_pos[i * 3 + 1]) : (c.bounce_in[i] != 0);
      const bool miss_ball = c.ball_pos[i * 3 + 1] < c.root_pos[i * 3 + 1] - 1.0f;
      const bool twice = hb && c.ball_pos[i * 3 + 2] < 0.05f;
      rec[w] = ta1 && (hc || miss_ball || twice);
      term[w] = (rec[w] && !hc) || (ta0 && hb && !bin);
    }
    c.distance[e] += sqrtf(c.root_vel[e * 3] * c.root_vel[e * 3] + c.root_vel[e * 3 + 1] * c.root_vel[e * 3 + 1]);  // _compute_stats
    bool reaction = rec[1], recovery = rec[0];     // "recovery also marks the reaction of their opponent"
    if (term[0] || term[1]) { c.reset_buf[e] = 1; reaction = false; recovery = false; }
    c.reset_reaction[e] = reaction ? 1 : 0;
    c.reset_recovery[e] = recovery ? 1 : 0;
  }
  // ---- reset FSM (:408-436)
  if (lane == 0 && !c.obs_only && !c.dual) {
    const bool out = rp[0] < c.court_min[0] || rp[1] < c.court_min[1] || rp[0] > c.court_max[0] || rp[1] > c.court_max[1];
    bool terminate = out || has_nan;
    int64_t reset = (c.progress_buf[e] >= (int64_t)c.max_episode_length - 1) ? 1 : (terminate ? 1 : 0);
    bool reaction = c.tar_time[e] == c.tar_time_total[e];
    const bool behind = bp[1] < rp[1] - 1.0f;
    bool recovery = (c.tar_action[e] == 1) && (has_contact || behind);
    c.distance[e] += sqrtf(c.root_vel[e * 3] * c.root_vel[e * 3] + c.root_vel[e * 3 + 1] * c.root_vel[e * 3 + 1]);  // _compute_stats
    if (c.early_termination) {
      terminate = terminate || (recovery && !has_contact) || behind;
      if (c.reward_type == 2) terminate = terminate || (has_contact && !c.est_bounce_in[e]);
    }
    if (terminate) { reset = 1; recovery = false; }
    reaction = reaction || (reset != 0);
    c.terminate_buf[e] = terminate ? 1 : 0;
    c.reset_buf[e] = reset;
    c.reset_reaction[e] = reaction ? 1 : 0;
    c.reset_recovery[e] = recovery ? 1 : 0;
  }
  if (lane == 0 && c.obs_only && touched) { c.reset_reaction[e] = 0; c.reset_recovery[e] = 0; }   // tail of _reset_envs (:176-177) for the reset humanoids
}

// ------------------------------------------------------------------------------------------ dual mode: incoming-ball table
struct InParams { float lo[4], hi_m_step[4], step[4], d1, d2, d3; };   // HEIGHT, VEL_X, VEL_Y, VSPIN
__device__ __forceinline__ float in_round(float v, const InParams& P, int k) {   // clamp, (v - lo) / step, round half-even (:27-45)
  v = fminf(fmaxf(v, P.lo[k]), P.hi_m_step[k]);
  return rintf(__fdiv_rn(__fsub_rn(v, P.lo[k]), P.step[k]));
}
__global__ void __launch_bounds__(V2P_WARPS * 32) ball_in_estimate_kernel(int n, const int64_t* __restrict__ contact_ids,
                                                                          const float* __restrict__ ball_states, int stride,
                                                                          const float* __restrict__ table, int64_t rows, InParams P,
                                                                          float* __restrict__ traj, float* __restrict__ s_in,
                                                                          float* __restrict__ s_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * V2P_WARPS + warp;
  if (i >= n) return;
  const float* b = ball_states + contact_ids[i] * stride;
  const float vel_x = sqrtf(__fadd_rn(__fmul_rn(b[7], b[7]), __fmul_rn(b[8], b[8])));
  const float dx = __fdiv_rn(b[7], vel_x), dy = __fdiv_rn(b[8], vel_x);
  const float vspin = __fdiv_rn(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(b[10], b[10]), __fmul_rn(b[11], b[11])), __fmul_rn(b[12], b[12]))),
                                6.283185307179586f);
  const float rh = in_round(b[2], P, 0), rx = in_round(vel_x, P, 1), ry = in_round(b[9], P, 2), rs = in_round(vspin, P, 3);
  // float32 index sum, truncated like .long() (:38-41)
  const float fi = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(__fmul_rn(rh, P.d1), P.d2), P.d3), __fmul_rn(__fmul_rn(rx, P.d2), P.d3)),
                                       __fmul_rn(ry, P.d3)), rs);
  int64_t row = (int64_t)fi;
  row = row < 0 ? 0 : (row >= rows ? rows - 1 : row);   // the reference would raise on an out-of-range row; clamp instead of faulting
  const float* t = table + row * 100;
  for (int k = lane; k < 50; k += 32) {   // traj_trans (:63-65): distance along the hit direction -> xy, mirrored for the receiver
    const float d = t[k * 2], z = t[k * 2 + 1];
    float* o = traj + (i * 50 + k) * 3;
    o[0] = -__fadd_rn(__fmul_rn(d, dx), b[0]);
    o[1] = -__fadd_rn(__fmul_rn(d, dy), b[1]);
    o[2] = z;
  }
  if (lane == 0) {
    const float height = __fadd_rn(__fmul_rn(rh, P.step[0]), P.lo[0]), vx = __fadd_rn(__fmul_rn(rx, P.step[1]), P.lo[1]);
    const float vy = __fadd_rn(__fmul_rn(ry, P.step[2]), P.lo[2]), vs = __fadd_rn(__fmul_rn(rs, P.step[3]), P.lo[3]);
    float* si = s_in + i * 13;
    float* so = s_out + i * 13;
    const float vin[2] = {__fmul_rn(-vx, dx), __fmul_rn(-vx, dy)};
    const float sp = __fmul_rn(__fmul_rn(vs, 3.141592653589793f), 2.0f);
    // omega = vspin * 2 pi * normalize(v x (0,0,-1)) = (-v_y, v_x, 0) / |.|
    const float nn = fmaxf(sqrtf(__fadd_rn(__fmul_rn(vin[1], vin[1]), __fmul_rn(vin[0], vin[0]))), 1e-12f);
    const float wi[3] = {__fmul_rn(sp, __fdiv_rn(-vin[1], nn)), __fmul_rn(sp, __fdiv_rn(vin[0], nn)), 0.0f};
    si[0] = -b[0]; si[1] = -b[1]; si[2] = height;
    so[0] = b[0];  so[1] = b[1];  so[2] = height;
#pragma unroll
    for (int k = 3; k < 7; k++) { si[k] = b[k]; so[k] = b[k]; }
    si[7] = vin[0]; si[8] = vin[1]; si[9] = vy;
    so[7] = -vin[0]; so[8] = -vin[1]; so[9] = vy;
#pragma unroll
    for (int k = 0; k < 3; k++) { si[10 + k] = wi[k]; so[10 + k] = k < 2 ? -wi[k] : 0.0f; }
  }
}

// ------------------------------------------------------------------------------------------ mask-driven task reset
__global__ void __launch_bounds__(V2P_WARPS * 32) task_reset_kernel(b200v2p_treset_t r) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * V2P_WARPS + warp;
  if (e >= r.n) return;
  const bool rea = r.reset_reaction[e] != 0, rec = r.reset_recovery[e] != 0;
  if (!rea && !rec) return;
  if (rea) {  // _reset_balls (:503-524) with the sample_random branch of the offline pool (tennis_ball.py:436-444)
    int64_t idx = r.pool_rand[e];
    if (r.ball_pos[e * 3 + 1] > 0.0f) {
      idx = (int64_t)((r.ball_pos[e * 3] + 4.0f) / 8.0f * (float)r.pool_size) + r.side_rand[e];
      idx = idx < 0 ? 0 : (idx > r.pool_size - 1 ? r.pool_size - 1 : idx);
    }
    const float* row = r.pool + idx * 307;
    __syncwarp();
    if (r.ball_obs) {   // use_history_ball_obs: _ball_traj is left alone (:187-188), _ball_obs <- the launch position repeated (:213-214)
      for (int k = lane; k < r.obs_traj_len * 3; k += 32) r.ball_obs[e * r.obs_traj_len * 3 + k] = row[k % 3];
    } else {
      for (int k = lane; k < 300; k += 32) r.ball_traj[e * 300 + k] = row[7 + k];
    }
    if (lane == 0) {
      float* b = r.ball_states + e * r.ball_stride;
      float* rb = r.rigid_body_state + (e * r.bodies_per_env + r.bodies_per_env - 1) * 13;
      const float v[3] = {row[3], row[4], row[5]};
      const float c[3] = {-v[1], v[0], 0.0f};
      const float nn = fmaxf(sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]), 1e-12f);
      const float sp = row[6] * 3.141592653589793f * 2.0f;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float w = sp * (c[k] / nn);
        b[k] = row[k]; b[7 + k] = v[k]; b[10 + k] = w;
        rb[k] = row[k]; rb[7 + k] = v[k]; rb[10 + k] = w;
        r.ball_pos[e * 3 + k] = row[k]; r.ball_vel[e * 3 + k] = v[k]; r.bounce_pos[e * 3 + k] = 0.0f;
      }
      r.has_bounce[e] = 0; r.has_contact[e] = 0;
    }
  }
  if (lane != 0) return;
  if (rec) {  // _reset_recovery_tasks (:242-245)
    r.tar_action[e] = 0;
    r.has_bounce[e] = 0;
    r.bounce_pos[e * 3] = 0.0f; r.bounce_pos[e * 3 + 1] = 0.0f; r.bounce_pos[e * 3 + 2] = 0.0f;
  }
  if (rea) {  // _reset_reaction_tasks (:203-240)
    r.tar_time[e] = 0;
    r.tar_action[e] = 1;
    r.num_reset_reaction[e] += 1;
    r.bounce_in[e] = 0;
    r.est_bounce_pos[e * 3] = 0.0f; r.est_bounce_pos[e * 3 + 1] = 0.0f; r.est_bounce_pos[e * 3 + 2] = 0.0f;
    r.est_bounce_time[e] = 0.0f;
    r.est_bounce_in[e] = 0;
    r.est_max_height[e] = 0.0f;
    r.swing_type_cycle[e] = -1;
    r.tar_time_total[e] = (int64_t)r.reaction_nframes + r.frame_rand[e];
    if (r.target_mode == 1) {        // 'continuous': one target shared by the envs reset in this call (:226-229)
#pragma unroll
      for (int k = 0; k < 3; k++) r.target_bounce_pos[e * 3 + k] = r.target_seed[k] * (r.target_max[k] - r.target_min[k]) + r.target_min[k];
    } else if (r.target_mode == 2) { // left / middle / right (:230-237)
      const float sd = r.target_seed[e];
      r.target_bounce_pos[e * 3] = sd < 0.33f ? -3.0f : (sd > 0.67f ? 3.0f : 0.0f);
      r.target_bounce_pos[e * 3 + 1] = 10.0f; r.target_bounce_pos[e * 3 + 2] = 0.0f;
    }
  }
}

// ------------------------------------------------------------------------------------------ humanoid reset from the FK pose
__global__ void __launch_bounds__(V2P_WARPS * 32) actor_reset_kernel(b200v2p_areset_t r) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * V2P_WARPS + warp;
  if (i >= r.n) return;
  const int64_t e = r.env_ids[i];
  if (r.mask && !r.mask[e]) return;
  const int nd = r.num_dof;
  float* rb = r.rigid_body_state + e * r.bodies_per_env * 13;
  if (lane < 24) {
    float* row = rb + lane * 13;
    const float* p = r.src_rb_pos + (e * 24 + lane) * 3;
    const float* q = r.src_rb_rot + (e * 24 + lane) * 4;
#pragma unroll
    for (int k = 0; k < 3; k++) { row[k] = p[k]; row[7 + k] = 0.0f; row[10 + k] = 0.0f; }
#pragma unroll
    for (int k = 0; k < 4; k++) { row[3 + k] = q[k]; r.prev_target_rb_rot[(e * 24 + lane) * 4 + k] = q[k]; }
  }
  if (lane == 24 && r.racket_body >= 0) {  // welded racket row = parent pose + rotated offset (first obs is consistent)
    const bool second = r.dual && (e & 1);
    const int rpar = second ? r.racket_parent2 : r.racket_parent;
    const float* pp = r.src_rb_pos + (e * 24 + rpar) * 3;
    const float* q = r.src_rb_rot + (e * 24 + rpar) * 4;
    const float* ro = second ? r.racket_offset2 : r.racket_offset;
    const float o[3] = {ro[0], ro[1], ro[2]};
    float t[3] = {2.0f * (q[1] * o[2] - q[2] * o[1]), 2.0f * (q[2] * o[0] - q[0] * o[2]), 2.0f * (q[0] * o[1] - q[1] * o[0])};
    float u[3] = {q[1] * t[2] - q[2] * t[1], q[2] * t[0] - q[0] * t[2], q[0] * t[1] - q[1] * t[0]};
    float* row = rb + r.racket_body * 13;
#pragma unroll
    for (int k = 0; k < 3; k++) { row[k] = pp[k] + o[k] + q[3] * t[k] + u[k]; row[7 + k] = 0.0f; row[10 + k] = 0.0f; }
#pragma unroll
    for (int k = 0; k < 4; k++) row[3 + k] = q[k];
  }
  for (int k = lane; k < nd; k += 32) {
    const float v = r.src_dof_pos[e * nd + k];
    r.dof_state[(e * nd + k) * 2] = v; r.dof_state[(e * nd + k) * 2 + 1] = 0.0f;
    r.pd_target_dof_pos[e * nd + k] = v;
  }
  if (lane == 0) {
    float* rs = r.root_states + e * r.root_stride;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float v = r.src_root_pos[e * 3 + k];
      rs[k] = v; rs[7 + k] = 0.0f; rs[10 + k] = 0.0f;
      r.prev_target_root_pos[e * 3 + k] = v; r.root_pos[e * 3 + k] = v; r.root_vel[e * 3 + k] = 0.0f; r.target_root_pos[e * 3 + k] = v;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) rs[3 + k] = r.src_root_rot[e * 4 + k];
    r.progress_buf[e] = 0; r.reset_buf[e] = 0; r.terminate_buf[e] = 0;
  }
}

// ------------------------------------------------------------------------------------------ controller bookkeeping of a humanoid reset
// PhysicsMVAEController._reset_envs (:176-181) for the envs whose mask is set: the controller's own progress / reset / terminate
// counters and statistics, one launch instead of seven masked fills
__global__ void ctrl_reset_kernel(int n, const uint8_t* __restrict__ mask, int64_t* progress, int64_t* reset_buf, int64_t* terminate,
                                  int64_t* num_reset_reaction, float* distance, int64_t* num_reset) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || !mask[e]) return;
  progress[e] = 0; reset_buf[e] = 0; terminate[e] = 0; num_reset_reaction[e] = 0; distance[e] = 0.0f; num_reset[e] += 1;
}

// ------------------------------------------------------------------------------------------ high-level pre_physics_step
// PhysicsMVAEController.pre_physics_step (:247-262) before the motion generator runs: the latent part of the action scaled by
// vae_action_scale, replaced by clamp(N(0,1), -5, 5) for the envs in recovery (random_walk_in_recovery), the residual-dof part scaled
// by residual_dof_scale.  The normal deviates come from a counter-based generator (splitmix64 of seed, step counter, element index +
// Box-Muller) so that the step needs no separate RNG launch; the step counter lives on the device and is advanced by the last
// block to finish, which keeps the launch replayable inside a CUDA graph.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ float counter_normal(uint64_t seed, uint64_t step, uint64_t idx) {
  const uint64_t h = splitmix64(splitmix64(seed ^ (step * 0xD1B54A32D192ED03ull)) + idx);
  const float u1 = ((uint32_t)(h >> 40) + 1u) * (1.0f / 16777217.0f);     // (0, 1)
  const float u2 = (uint32_t)((h >> 8) & 0xFFFFFFu) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
}
__device__ __forceinline__ void last_block_advance(int64_t* counter, unsigned int* done, int64_t inc) {
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) { counter[0] += inc; *done = 0u; __threadfence(); }
}
__global__ void pre_step_kernel(b200v2p_prestep_t p) {
  const int W = p.num_latent + p.num_res_dof;
  const int64_t step = p.step_counter[0];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int64_t)p.n * W) {
    const int64_t e = i / W;
    const int k = (int)(i - e * W);
    const float a = p.actions[e * p.num_actions + k];
    if (k < p.num_latent) {
      float v = a * p.vae_action_scale;
      if (p.random_walk_in_recovery && p.tar_action[e] == 0) v = fminf(fmaxf(counter_normal(p.seed, (uint64_t)step, (uint64_t)i), -5.0f), 5.0f);
      p.mvae_actions[e * p.num_latent + k] = v;
    } else {
      p.res_dof_actions[e * p.num_res_dof + (k - p.num_latent)] = a * p.residual_dof_scale;
    }
  }
  last_block_advance(p.step_counter, p.done_counter, 1);
}

// ------------------------------------------------------------------------------------------ resident kinematic target stream
// StreamMotionPlayer (tasks/physics_mvae_controller.py; SURVEY.md 8d "synthetic kinematic target stream"): env e reads frame
// (t + off[e]) % K of the ring kept in HBM into the live buffers the FK kernel and the controller read.  One launch instead of six
// index_selects and the index arithmetic; `advance` = 1 moves the stream clock on by one frame first (step), 0 re-reads (reset).
__global__ void __launch_bounds__(V2P_WARPS * 32) stream_gather_kernel(b200v2p_stream_t s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * V2P_WARPS + warp;
  const int64_t t = s.clock[0] + s.advance;
  if (e < s.n && (!s.reseed_mask || s.reseed_mask[e])) {
    if (s.reseed_mask) {          // reset of the masked envs: a fresh offset into the stream (counter-based draw), others untouched
      const uint64_t h = splitmix64(splitmix64(s.seed ^ ((uint64_t)t * 0xD1B54A32D192ED03ull)) + (uint64_t)e);
      if (lane == 0) s.offset[e] = (int64_t)(h % (uint64_t)s.frames);
      __syncwarp();
    }
    int64_t f = (t + s.offset[e]) % s.frames;
    if (f < 0) f += s.frames;
    const int64_t row = f * s.n + e;
    const float4* src = reinterpret_cast<const float4*>(s.ring_rotmat + row * 216);
    float4* dst = reinterpret_cast<float4*>(s.rotmat + e * 216);
    for (int k = lane; k < 54; k += 32) dst[k] = src[k];
    if (lane < 3) { s.root_pos[e * 3 + lane] = s.ring_root_pos[row * 3 + lane]; s.racket_pos[e * 3 + lane] = s.ring_racket_pos[row * 3 + lane]; }
    if (lane == 3) s.phase[e] = s.ring_phase[row];
    if (lane == 4) s.swing_type[e] = s.ring_swing_type[row];
    if (lane == 5) s.swing_type_cycle[e] = s.ring_swing_type_cycle[row];
  }
  if (s.advance) last_block_advance(s.clock, s.done_counter, s.advance);
}

// ------------------------------------------------------------------------------------------ C ABI
extern "C" {

const char* b200v2p_last_error(void) { return g_verr; }

int b200v2p_smpl_to_sim(int32_t n, const float* root_pos, const float* joint_rotmat, const float* rest, int32_t num_rest,
                        const int32_t* parents, const int32_t* smpl_2_mujoco, float dt, const float* prev_root_pos, const float* prev_rb_rot, float* root_rot,
                        float* dof_pos, float* root_vel, float* root_ang_vel, float* dof_vel, float* rb_pos, float* rb_rot, float* prev_root_pos_update,
                        float* target_root_pos_out, const uint8_t* only_mask, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || !root_pos || !joint_rotmat || !rest || !parents || !smpl_2_mujoco || !root_rot || !dof_pos || !root_vel || !root_ang_vel ||
      !dof_vel || !rb_pos || !rb_rot)
    return vfail(-1, "b200v2p_smpl_to_sim: bad arguments");
  if (!(dt > 0)) return vfail(-2, "b200v2p_smpl_to_sim: dt must be positive");
  if (num_rest < 1) return vfail(-2, "b200v2p_smpl_to_sim: num_rest must be >= 1");
  smpl_to_sim_kernel<<<(n + V2P_WARPS - 1) / V2P_WARPS, V2P_WARPS * 32, 0, (cudaStream_t)stream>>>(
      n, root_pos, joint_rotmat, rest, num_rest, parents, smpl_2_mujoco, dt, prev_root_pos, prev_rb_rot, root_rot, dof_pos, root_vel, root_ang_vel,
      dof_vel, rb_pos, rb_rot, prev_root_pos_update, target_root_pos_out, only_mask);
  V_CUDA_OK();
  return 0;
}

int b200v2p_fix_head(int32_t n, const float* rb_pos, const float* rb_rot, int32_t head_body, const float* ball_pos, const float* root_pos,
                     float* joint_rotmat, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || !rb_pos || !rb_rot || !ball_pos || !root_pos || !joint_rotmat || head_body < 0 || head_body > 23)
    return vfail(-1, "b200v2p_fix_head: bad arguments");
  fix_head_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, rb_pos, rb_rot, head_body, ball_pos, root_pos, joint_rotmat);
  V_CUDA_OK();
  return 0;
}

int b200v2p_ball_aero(int32_t n, const float* ball_states, int32_t stride, uint8_t* has_bounce, uint8_t* has_bounce_now, float* bounce_pos,
                      float* force, int32_t substeps, float spin_scale, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || !ball_states || !has_bounce || !has_bounce_now || !bounce_pos || !force || stride < 13) return vfail(-1, "b200v2p_ball_aero: bad arguments");
  ball_aero_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, ball_states, stride, has_bounce, has_bounce_now, bounce_pos, force,
                                                                     substeps, spin_scale);
  V_CUDA_OK();
  return 0;
}

int b200v2p_ball_reset(int32_t n, const int64_t* env_ids, const int64_t* pool_index, const float* pool, float* ball_states, int32_t stride,
                       float* ball_pos, float* ball_vel, uint8_t* has_bounce, float* bounce_pos, uint8_t* has_contact, float* traj,
                       void* stream) {
  if (n == 0) return 0;
  if (n < 0 || !env_ids || !pool_index || !pool || !ball_states || !ball_pos || !ball_vel || !has_bounce || !bounce_pos || !has_contact || !traj)
    return vfail(-1, "b200v2p_ball_reset: bad arguments");
  ball_reset_kernel<<<n, 64, 0, (cudaStream_t)stream>>>(n, env_ids, pool_index, pool, ball_states, stride, ball_pos, ball_vel, has_bounce,
                                                        bounce_pos, has_contact, traj);
  V_CUDA_OK();
  return 0;
}

int b200v2p_ball_in_estimate(int32_t n, const int64_t* contact_ids, const float* ball_states, int32_t stride, const float* table,
                             int64_t table_rows, const double* params, float* traj, float* states_in, float* states_out, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || !contact_ids || !ball_states || !table || table_rows < 1 || !params || !traj || !states_in || !states_out)
    return vfail(-1, "b200v2p_ball_in_estimate: bad arguments");
  InParams P;
  double dim[4];
  for (int k = 0; k < 4; k++) {   // python-double arithmetic of the reference, then the float32 cast torch applies to scalars
    const double lo = params[k * 3], hi = params[k * 3 + 1], st = params[k * 3 + 2];
    P.lo[k] = (float)lo; P.hi_m_step[k] = (float)(hi - st); P.step[k] = (float)st;
    dim[k] = (hi - lo) / st;
  }
  P.d1 = (float)dim[1]; P.d2 = (float)dim[2]; P.d3 = (float)dim[3];
  ball_in_estimate_kernel<<<(n + V2P_WARPS - 1) / V2P_WARPS, V2P_WARPS * 32, 0, (cudaStream_t)stream>>>(
      n, contact_ids, ball_states, stride, table, table_rows, P, traj, states_in, states_out);
  V_CUDA_OK();
  return 0;
}

int b200v2p_update_state(const b200v2p_state_t* s, void* stream) {
  if (!s) return vfail(-1, "b200v2p_update_state: null");
  if (s->n == 0) return 0;
  update_state_kernel<<<(s->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*s);
  V_CUDA_OK();
  return 0;
}

int b200v2p_actor_reset(const b200v2p_areset_t* r, void* stream) {
  if (!r) return vfail(-1, "b200v2p_actor_reset: null");
  if (r->n == 0) return 0;
  if (r->n < 0 || !r->env_ids || !r->src_root_pos || !r->root_states) return vfail(-1, "b200v2p_actor_reset: bad arguments");
  actor_reset_kernel<<<(r->n + V2P_WARPS - 1) / V2P_WARPS, V2P_WARPS * 32, 0, (cudaStream_t)stream>>>(*r);
  V_CUDA_OK();
  return 0;
}

int b200v2p_task_reset(const b200v2p_treset_t* r, void* stream) {
  if (!r) return vfail(-1, "b200v2p_task_reset: null");
  if (r->n == 0) return 0;
  if (r->pool_size < 1 || !r->pool || !r->reset_reaction || !r->reset_recovery) return vfail(-1, "b200v2p_task_reset: bad arguments");
  task_reset_kernel<<<(r->n + V2P_WARPS - 1) / V2P_WARPS, V2P_WARPS * 32, 0, (cudaStream_t)stream>>>(*r);
  V_CUDA_OK();
  return 0;
}

int b200v2p_ctrl_reset(int32_t n, const uint8_t* mask, int64_t* progress_buf, int64_t* reset_buf, int64_t* terminate_buf, int64_t* num_reset_reaction,
                       float* distance, int64_t* num_reset, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || !mask || !progress_buf || !reset_buf || !terminate_buf || !num_reset_reaction || !distance || !num_reset)
    return vfail(-1, "b200v2p_ctrl_reset: bad arguments");
  ctrl_reset_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, mask, progress_buf, reset_buf, terminate_buf, num_reset_reaction, distance, num_reset);
  V_CUDA_OK();
  return 0;
}

int b200v2p_pre_step(const b200v2p_prestep_t* p, void* stream) {
  if (!p) return vfail(-1, "b200v2p_pre_step: null");
  if (p->n == 0) return 0;
  if (p->n < 0 || !p->actions || !p->mvae_actions || !p->step_counter || !p->done_counter || p->num_latent < 1 || p->num_res_dof < 0 ||
      p->num_actions < p->num_latent + p->num_res_dof || (p->num_res_dof > 0 && !p->res_dof_actions) || (p->random_walk_in_recovery && !p->tar_action))
    return vfail(-1, "b200v2p_pre_step: bad arguments");
  const int64_t total = (int64_t)p->n * (p->num_latent + p->num_res_dof);
  pre_step_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*p);
  V_CUDA_OK();
  return 0;
}

int b200v2p_stream_gather(const b200v2p_stream_t* s, void* stream) {
  if (!s) return vfail(-1, "b200v2p_stream_gather: null");
  if (s->n == 0) return 0;
  if (s->n < 0 || s->frames < 1 || !s->clock || !s->done_counter || !s->offset || !s->ring_rotmat || !s->rotmat || !s->ring_root_pos || !s->root_pos ||
      !s->ring_racket_pos || !s->racket_pos || !s->ring_phase || !s->phase || !s->ring_swing_type || !s->swing_type || !s->ring_swing_type_cycle ||
      !s->swing_type_cycle || s->advance < 0)
    return vfail(-1, "b200v2p_stream_gather: bad arguments");
  if (((uintptr_t)s->ring_rotmat | (uintptr_t)s->rotmat) & 15) return vfail(-2, "b200v2p_stream_gather: rotation matrices must be 16-byte aligned");
  stream_gather_kernel<<<(s->n + V2P_WARPS - 1) / V2P_WARPS, V2P_WARPS * 32, 0, (cudaStream_t)stream>>>(*s);
  V_CUDA_OK();
  return 0;
}

int b200v2p_controller_post(const b200v2p_ctrl_t* c, void* stream) {
  if (!c) return vfail(-1, "b200v2p_controller_post: null");
  if (c->n == 0) return 0;
  if (c->obs_traj_len < 0 || c->obs_traj_len > 100 || c->num_obs < 225 + 3 * c->obs_traj_len + (c->use_target ? 2 : 0))
    return vfail(-2, "b200v2p_controller_post: observation width / trajectory length mismatch");
  if (c->reward_type < 0 || c->reward_type > 2) return vfail(-2, "b200v2p_controller_post: reward_type must be 0 (reach), 1 (return) or 2 (return_w_estimate)");
  if (c->use_history && !c->ball_obs) return vfail(-2, "b200v2p_controller_post: use_history needs the ball_obs buffer");
  controller_post_kernel<<<(c->n + V2P_WARPS - 1) / V2P_WARPS, V2P_WARPS * 32, 0, (cudaStream_t)stream>>>(*c);
  V_CUDA_OK();
  return 0;
}

}  // extern "C"
