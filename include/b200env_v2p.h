/* b200env_v2p.h - C ABI of the vid2player rows of the rollout hot path.
 * Stateless entry points: plain device pointers + sizes, asynchronous on `stream`, no allocation, no host sync.
 * Bool tensors are 1 byte per element (torch.bool storage).  Paths below are relative to /root/reference/vid2player. */
#ifndef B200ENV_V2P_H
#define B200ENV_V2P_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* b200v2p_last_error(void);

/* replaces HumanoidSMPLIMMVAE._smpl_to_sim + _forward_kinematics (env/tasks/humanoid_smpl_im_mvae.py:897-946, utils/hybrik.py:597-652):
 * joint_rotmat [n,24,3,3] and rest [num_rest,24,3] in SMPL joint order (env e uses shape e % num_rest: 1 = one player,
 * 2 = dual mode's alternating players, n = a shape per env); outputs in MuJoCo body order (smpl_2_mujoco);
 * prev_* NULL -> zero velocities.  prev_root_pos_update / target_root_pos_out (either may be NULL): after the finite difference the
 * root position of this call is stored there - `_save_prev_target_motion_state` (:741-750) and `self._target_root_pos = root_pos`
 * (:655) folded into the launch; prev_root_pos_update may be the same buffer as prev_root_pos.  only_mask (may be NULL): bool [n], the
 * envs whose entry is 0 are skipped (mask-driven reset: FK of the envs being reset only). */
int b200v2p_smpl_to_sim(int32_t n, const float* root_pos, const float* joint_rotmat, const float* rest, int32_t num_rest,
                        const int32_t* parents, const int32_t* smpl_2_mujoco, float dt, const float* prev_root_pos, const float* prev_rb_rot, float* root_rot,
                        float* dof_pos, float* root_vel, float* root_ang_vel, float* dof_vel, float* rb_pos, float* rb_rot, float* prev_root_pos_update,
                        float* target_root_pos_out, const uint8_t* only_mask, void* stream);

/* replaces the fix_head_orientation block of _set_target_motion_state (env/tasks/humanoid_smpl_im_mvae.py:605-634):
 * rb_pos/rb_rot [n,24,*] = FK of the uncorrected pose (MuJoCo order, head_body = 13); joint_rotmat [n,24,3,3] (SMPL order) is
 * corrected in place at SMPL joints Head (15) and Neck (12). */
int b200v2p_fix_head(int32_t n, const float* rb_pos, const float* rb_rot, int32_t head_body, const float* ball_pos, const float* root_pos,
                     float* joint_rotmat, void* stream);

/* replaces apply_external_force_to_ball (env/tasks/humanoid_smpl_im_mvae.py:711-739): drag + Magnus lift, bounce flag */
int b200v2p_ball_aero(int32_t n, const float* ball_states, int32_t stride, uint8_t* has_bounce, uint8_t* has_bounce_now, float* bounce_pos,
                      float* force, int32_t substeps, float spin_scale, void* stream);

/* replaces _reset_balls + TennisBallGeneratorOffline.generate (:503-524, utils/tennis_ball.py:435-456): pool [P,307] */
int b200v2p_ball_reset(int32_t n, const int64_t* env_ids, const int64_t* pool_index, const float* pool, float* ball_states, int32_t stride,
                       float* ball_pos, float* ball_vel, uint8_t* has_bounce, float* bounce_pos, uint8_t* has_contact, float* traj,
                       void* stream);

/* replaces _update_state_from_sim (:799-860), substeps > 2 contact detector */
typedef struct b200v2p_state {
  int32_t n, bodies_per_env, ball_stride, root_stride, racket_body, wrist_body;
  float grip_normal[3];
  /* dual_mode 'different' (:839-842, :79-84): odd envs use the second player's grip / racket bodies when dual != 0 */
  int32_t dual, racket_body2, wrist_body2;
  float grip_normal2[3];
  const float* rigid_body_state; /* [n, bodies_per_env, 13] */
  const float* root_states;      /* humanoid root row per env, stride root_stride floats */
  const float* ball_states;      /* ball root row per env, stride ball_stride floats */
  uint8_t *has_contact, *has_contact_now;
  float *root_pos, *root_vel, *racket_pos, *racket_vel, *racket_normal, *ball_pos, *ball_vel, *ball_vspin;
  const uint8_t* only_mask;      /* NULL, or bool [n]: refresh the envs whose entry is set only (reset path) */
} b200v2p_state_t;
int b200v2p_update_state(const b200v2p_state_t* s, void* stream);

/* replaces PhysicsMVAEController.post_physics_step's device work (env/tasks/physics_mvae_controller.py:441-452):
 * _update_state (:271-314, incl. TennisBallOutEstimator.estimate utils/tennis_ball_out_estimator.py:164-205),
 * _compute_reward (:368-406, jit :493-602), _compute_observations (:316-360), _compute_reset (:408-436). */
typedef struct b200v2p_ctrl {
  int32_t n, bodies_per_env, ball_stride, racket_body, num_obs, obs_traj_len, use_target, reward_type, early_termination,
      max_episode_length, est_nx, est_ny;
  int32_t obs_only; /* 1: refresh obs_buf only (the _compute_observations call of _reset_envs :200-201), touch nothing else */
  int32_t dual;     /* 1: reset FSM of PhysicsMVAEControllerDual._compute_reset (env/tasks/physics_mvae_controller_dual.py:92-120):
                       envs (2k, 2k+1) are opponents; reset_buf is only ever SET, terminate_buf is not written */
  int32_t use_history; /* cfg use_history_ball_obs (:348-351): the task observation is the HISTORY of ball positions (ball_obs)
                          instead of the future trajectory window (ball_traj) */
  int32_t advance;     /* 1 (ignored with obs_only): first do the tail of physics_step and the head of post_physics_step (:364-366,
                          :441-444) - ball_traj <- roll(-1) with a zero last frame, tar_time += 1, progress_buf += 1 - then the rest */
  float scale_pos, scale_phase, scale_bounce_pos, scale_bounce_time, w_pos, w_ball_pos;
  float court_min[2], court_max[2];
  float est_params[15]; /* VEL_X, VEL_Y, VSPIN, TRAJ_X, TRAJ_Y ranges (lo, hi, step) */
  const float* rigid_body_state;
  const float* ball_states;
  const float *root_pos, *root_vel, *racket_pos, *racket_normal, *ball_pos;
  const uint8_t *has_contact, *has_contact_now, *has_bounce, *has_bounce_now;
  const float* bounce_pos;
  float* ball_traj; /* [n,100,3]; written only with advance = 1 */
  const float* target_bounce_pos;
  const float* phase;
  const int64_t *swing_type, *swing_type_cycle, *tar_action;
  int64_t* tar_time;            /* written only with advance = 1 */
  const int64_t* tar_time_total;
  int64_t* progress_buf;        /* written only with advance = 1 */
  const float *est_x, *est_y; /* estimator tables [rows, est_nx], [rows, est_ny, 2]; NULL = no estimator (dual mode) */
  uint8_t *bounce_in, *est_bounce_in, *reset_reaction, *reset_recovery;
  float *est_bounce_pos, *est_bounce_time, *est_max_height, *distance;
  float *obs_buf, *rew_buf, *sub_rewards; /* sub_rewards [n,2] */
  int64_t *reset_buf, *terminate_buf;
  float* ball_obs; /* [n, obs_traj_len, 3] _ball_obs (:65): rolled by one and appended with ball_pos at every observation (:345-346);
                      in obs_only mode only for the envs whose reset masks are set (the reference refreshes those ids only).
                      NULL = not kept (then use_history must be 0) */
  const uint8_t* touch_mask; /* NULL, or bool [n] with obs_only: the humanoids just reset.  Rows of envs that are neither in it nor flagged
                                reset_reaction / reset_recovery are still current and skipped; the task flags of the envs in it are cleared
                                at the end (tail of _reset_envs, :176-177) */
} b200v2p_ctrl_t;
int b200v2p_controller_post(const b200v2p_ctrl_t* c, void* stream);

/* replaces the action handling of PhysicsMVAEController.pre_physics_step (env/tasks/physics_mvae_controller.py:247-262) that precedes
 * the motion generator: mvae_actions = actions[:, :num_latent] * vae_action_scale, replaced by clamp(N(0,1), -5, 5) where
 * tar_action == 0 when random_walk_in_recovery; res_dof_actions = actions[:, num_latent : num_latent + num_res_dof] *
 * residual_dof_scale.  The normal deviates are counter-based (seed, *step_counter, element), *step_counter is advanced by the launch
 * (done_counter: one zero-initialised uint32 of scratch), so that a captured step replays with fresh numbers. */
typedef struct b200v2p_prestep {
  int32_t n, num_actions, num_latent, num_res_dof, random_walk_in_recovery;
  float vae_action_scale, residual_dof_scale;
  uint64_t seed;
  const float* actions;        /* [n, num_actions] */
  const int64_t* tar_action;   /* [n] */
  int64_t* step_counter;       /* [1] */
  uint32_t* done_counter;      /* [1] scratch, zero between launches */
  float* mvae_actions;         /* [n, num_latent] */
  float* res_dof_actions;      /* [n, num_res_dof] or NULL */
} b200v2p_prestep_t;
int b200v2p_pre_step(const b200v2p_prestep_t* p, void* stream);

/* the controller's own bookkeeping of a humanoid reset (env/tasks/physics_mvae_controller.py:176-181, 203-210) for the envs whose mask
 * is set: progress / reset / terminate / num_reset_reaction <- 0, distance <- 0, num_reset += 1 */
int b200v2p_ctrl_reset(int32_t n, const uint8_t* mask, int64_t* progress_buf, int64_t* reset_buf, int64_t* terminate_buf, int64_t* num_reset_reaction,
                       float* distance, int64_t* num_reset, void* stream);

/* The resident kinematic target stream that stands in for the unreleased MVAE motion generator: `frames`
 * frames of every field for all n envs are kept in HBM ([frames * n, ...], frame-major); env e reads frame (clock + advance +
 * offset[e]) % frames into the live buffers MVAEPlayer exposes (players/mvae_player.py: _joint_rotmat [n,24,3,3], _root_pos,
 * _racket_pos [n,3], _phase_pred [n], _swing_type, _swing_type_cycle [n] int64).  advance = 1: a step (the clock moves on), 0: re-read
 * after a reset changed offsets. */
typedef struct b200v2p_stream {
  int32_t n, frames, advance, pad_;
  int64_t* clock;              /* [1] */
  uint32_t* done_counter;      /* [1] scratch, zero between launches */
  int64_t* offset;             /* [n] */
  const uint8_t* reseed_mask;  /* NULL, or bool [n] (reset): the envs whose entry is set draw a new offset (counter-based: seed, clock, env)
                                  and re-read their frame; the others are left alone */
  uint64_t seed;
  const float* ring_rotmat;    float* rotmat;
  const float* ring_root_pos;  float* root_pos;
  const float* ring_racket_pos; float* racket_pos;
  const float* ring_phase;     float* phase;
  const int64_t* ring_swing_type; int64_t* swing_type;
  const int64_t* ring_swing_type_cycle; int64_t* swing_type_cycle;
} b200v2p_stream_t;
int b200v2p_stream_gather(const b200v2p_stream_t* s, void* stream);

/* replaces TennisBallInEstimator.estimate (utils/tennis_ball_in_estimator.py:22-81), the table lookup of
 * HumanoidSMPLIMMVAEDual._reset_balls (env/tasks/humanoid_smpl_im_mvae_dual.py:63-72): for query i the ball row
 * ball_states + contact_ids[i]*stride (the ball the opponent just hit) is snapped to the table grid; outputs the incoming
 * trajectory in the receiver's frame traj[n,50,3] and the snapped in / out ball states [n,13].  The caller scatters them
 * (states[ball_ids] = in, then states[contact_ids] = out).  table [rows,50,2]; params = HEIGHT, VEL_X, VEL_Y, VSPIN (lo,hi,step). */
int b200v2p_ball_in_estimate(int32_t n, const int64_t* contact_ids, const float* ball_states, int32_t stride, const float* table,
                             int64_t table_rows, const double* params, float* traj, float* states_in, float* states_out, void* stream);

/* replaces the per-step task part of PhysicsMVAEController._reset_envs when no humanoid needs a reset
 * (env/tasks/physics_mvae_controller.py:173-201: _reset_balls for the reaction envs :503-524, _reset_recovery_tasks :242-245,
 * _reset_reaction_tasks :203-240), driven by the device-side masks instead of nonzero() id lists: no host sync.
 * Random draws come in as device tensors (torch RNG): pool_rand[n] in [0,P), side_rand[n] in [-1000,1000), frame_rand[n] in [-5,5),
 * target_seed: 3 floats (continuous mode) or n floats (discrete mode). */
typedef struct b200v2p_treset {
  int32_t n, pool_size, ball_stride, bodies_per_env, reaction_nframes, target_mode; /* 0 none, 1 continuous, 2 discrete */
  float target_min[3], target_max[3];
  const uint8_t *reset_reaction, *reset_recovery;
  const int64_t *pool_rand, *side_rand, *frame_rand;
  const float* target_seed;
  const float* pool;          /* [P,307] */
  float* ball_states;         /* first ball root row, stride ball_stride */
  float* rigid_body_state;    /* ball = last body row of each env */
  float *ball_pos, *ball_vel, *bounce_pos, *ball_traj, *est_bounce_pos, *est_bounce_time, *est_max_height, *target_bounce_pos;
  uint8_t *has_bounce, *has_contact, *bounce_in, *est_bounce_in;
  int64_t *tar_time, *tar_time_total, *tar_action, *num_reset_reaction, *swing_type_cycle;
  float* ball_obs;       /* use_history_ball_obs (:213-214): rows of the reaction envs <- the new ball position repeated; NULL = off */
  int32_t obs_traj_len, pad_;
} b200v2p_treset_t;
int b200v2p_task_reset(const b200v2p_treset_t* r, void* stream);

/* replaces HumanoidSMPLIMMVAE._reset_actors + _set_env_state + _reset_env_tensors for an id list
 * (env/tasks/humanoid_smpl_im_mvae.py:463-501, 562-581): the FK of the motion generator's pose (b200v2p_smpl_to_sim over all
 * envs, prev = NULL) is scattered into the simulation state of the listed envs; velocities zero. */
typedef struct b200v2p_areset {
  int32_t n, num_dof, bodies_per_env, root_stride, racket_body, racket_parent;
  float racket_offset[3];
  int32_t dual;             /* 1: odd envs use racket_offset2 / racket_parent2 (second player's asset) */
  float racket_offset2[3];
  int32_t racket_parent2, pad_;
  const int64_t* env_ids;
  const float *src_root_pos, *src_root_rot, *src_dof_pos, *src_rb_pos, *src_rb_rot; /* [N,...] FK results */
  float *root_states, *dof_state, *rigid_body_state;
  float *prev_target_root_pos, *prev_target_rb_rot, *root_pos, *root_vel, *pd_target_dof_pos, *target_root_pos;
  int64_t *progress_buf, *reset_buf, *terminate_buf;
  const uint8_t* mask;      /* optional [N]: row env_ids[i] is reset only when mask[env_ids[i]] != 0 (mask-driven form: env_ids =
                               0..N-1, n = N; no id list has to be built on the host, the launch can live in a CUDA graph) */
} b200v2p_areset_t;
int b200v2p_actor_reset(const b200v2p_areset_t* r, void* stream);

#ifdef __cplusplus
}
#endif
#endif
