/*
 * pulse_b200.h -- C ABI of the B200-native PULSE hot path (libpulse_b200.so).
 *
 * Boundary rules:
 *   - plain C types only: device pointers, element strides, sizes, a cudaStream_t passed as void*;
 *   - every buffer is owned by the caller (torch tensors or Isaac Gym gymtorch views); nothing is
 *     allocated on the device inside the library;
 *   - no hidden synchronisation: kernels are enqueued on the caller's stream and the call returns;
 *   - every entry point returns 0 on success or a negative pulse_status; the message is available
 *     from pulse_last_error() (thread-local).  No exceptions cross the boundary;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     PULSE_ERR_CUDA;
 *   - ONE device per process (the reference's model: one Isaac Gym sim per process, run_hydra.py:117-131): launch attributes,
 *     the SM count and the persistent-grid sizes are cached per process on first use, so a process must not drive two
 *     different devices through this library; calls are made from one host thread per process, on the caller's stream.
 *
 * Each entry point cites the reference interface (file:line under the PULSE tree) it replaces.
 */
#ifndef PULSE_B200_H_
#define PULSE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PULSE_ABI_VERSION 2

enum pulse_status {
  PULSE_OK = 0,
  PULSE_ERR_ARG = -1,    /* null pointer / bad size / bad stride / misaligned table */
  PULSE_ERR_CUDA = -2,   /* CUDA runtime error (launch, no device, ...) */
  PULSE_ERR_UNSUPPORTED = -3
};

#define PULSE_NUM_BODIES 24        /* SMPL humanoid rigid bodies (smpl_humanoid.xml)        */
#define PULSE_NUM_DOF 69           /* 23 joints x 3                                           */
#define PULSE_BODY_STATE_W 13      /* pos3 quat4(xyzw) linvel3 angvel3, humanoid.py:215-222  */
#define PULSE_SELF_OBS 358         /* humanoid.py:1675-1731                                   */
#define PULSE_TASK_OBS_V6 576      /* humanoid_im.py:1328-1378                                */
#define PULSE_IM_OBS (PULSE_SELF_OBS + PULSE_TASK_OBS_V6)
#define PULSE_AMP_OBS 196          /* humanoid_amp.py:924-969 with dof_subset                 */
#define PULSE_FRAME_REC 312        /* packed per-frame record: pos72 | rot96 | vel72 | angvel72 */
#define PULSE_AUX_REC 240          /* packed per-frame record: lrs96 | dvs69 | aa72 | pad3    */

int pulse_abi_version(void);
const char* pulse_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py "gpu_launches"). */
int64_t pulse_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * MotionLib tables.  Replaces the device-resident buffers MotionLibBase.load_motions builds
 * (phc/utils/motion_lib_base.py:287-316) with two packed per-frame records so that one query
 * gathers one contiguous 1248-byte row per frame instead of rows of six separate tables.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pulse_motionlib pulse_motionlib_t;

typedef struct {
  /* reference tables, device pointers, contiguous fp32 / int64 (motion_lib_base.py:297-315) */
  const float* gts;   /* [F,24,3] */
  const float* grs;   /* [F,24,4] */
  const float* lrs;   /* [F,24,4] */
  const float* gvs;   /* [F,24,3] */
  const float* gavs;  /* [F,24,3] */
  const float* dvs;   /* [F,23,3] */
  const float* motion_aa; /* [F,72] (may be NULL -> zeros) */
  const float* lengths;         /* [M] _motion_lengths */
  const float* dt;              /* [M] _motion_dt */
  const int64_t* num_frames;    /* [M] _motion_num_frames */
  const int64_t* length_starts; /* [M] length_starts */
  int64_t total_frames;         /* F */
  int64_t num_motions;          /* M */
  /* caller-allocated packed outputs, 16-byte aligned, filled by pulse_motionlib_create */
  float* frame_rec;   /* [F, PULSE_FRAME_REC] */
  float* aux_rec;     /* [F, PULSE_AUX_REC]   */
} pulse_motionlib_desc_t;

/* Packs the tables (one kernel on `stream`) and returns a host-side handle that keeps the pointers.
 * The per-motion arrays and the packed records must outlive the handle. */
int pulse_motionlib_create(const pulse_motionlib_desc_t* desc, void* stream, pulse_motionlib_t** out);
int pulse_motionlib_destroy(pulse_motionlib_t* lib);

/* MotionLibBase.get_motion_state(motion_ids, motion_times, offset)  motion_lib_base.py:434-517
 * (+ _calc_frame_blend :546-556, _local_rotation_to_dof_smpl :561-564).  Any output may be NULL. */
typedef struct {
  const int64_t* motion_ids;   /* [n] */
  const float* motion_times;   /* [n] */
  const float* offset;         /* [n,3] or NULL */
  float* root_pos;      /* [n,3]   */
  float* root_rot;      /* [n,4]   */
  float* dof_pos;       /* [n,69]  */
  float* root_vel;      /* [n,3]   */
  float* root_ang_vel;  /* [n,3]   */
  float* dof_vel;       /* [n,69]  */
  float* motion_aa;     /* [n,72]  */
  float* rg_pos;        /* [n,24,3] */
  float* rb_rot;        /* [n,24,4] */
  float* body_vel;      /* [n,24,3] */
  float* body_ang_vel;  /* [n,24,3] */
  int64_t* frame_idx0;  /* [n] (diagnostic: _calc_frame_blend) */
  int64_t* frame_idx1;  /* [n] */
  float* blend;         /* [n] */
} pulse_motion_query_t;
int pulse_motion_state(const pulse_motionlib_t* lib, const pulse_motion_query_t* q, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused HumanoidIm post-physics step: reward(t) -> reset(t) -> observation(t+dt) in ONE kernel.
 * Replaces, for the default task configuration (obs_v 6, self_obs_v 1, full-body reward, 24 tracked
 * bodies, upright start, local_root_obs, root_height_obs):
 *   HumanoidIm._compute_reward        phc/env/tasks/humanoid_im.py:853-919  (compute_imitation_reward :1543-1574)
 *   HumanoidIm._compute_reset         humanoid_im.py:1119-1192              (compute_humanoid_im_reset :1600-1628)
 *   HumanoidIm._compute_observations  humanoid_im.py:677-706, _compute_task_obs :708-851
 *                                     (compute_imitation_observations_v6 :1328-1378)
 *   Humanoid._compute_humanoid_obs    phc/env/tasks/humanoid.py:1137-1213   (compute_humanoid_observations_smpl_max :1675-1731)
 *   and the two MotionLib queries behind _get_state_from_motionlib_cache (humanoid_im.py:950-964).
 * `progress_buf` is the value AFTER the reference's `progress_buf += 1` (humanoid.py:1317).
 * ---------------------------------------------------------------------------------------------- */
#define PULSE_STEP_REWARD 1u
#define PULSE_STEP_RESET 2u
#define PULSE_STEP_OBS 4u
#define PULSE_STEP_ALL 7u
#define PULSE_STEP_ADVANCE 8u   /* ABI 2: the kernel itself performs `progress_buf += 1` (humanoid.py:1317) through progress_rw before using it */

typedef struct {
  /* simulator state (Isaac Gym views, read-only) */
  const float* body_state;   /* rigid body state; env e, body j at body_state + e*body_env_stride + j*13 */
  int64_t body_env_stride;   /* floats between envs = bodies_per_env*13 (humanoid.py:215-222) */
  const float* dof_vel;      /* dof velocity; element k of env e at dof_vel + e*dof_env_stride + k*dof_elem_stride */
  int64_t dof_env_stride;    /* Isaac Gym dof-state view: dofs_per_env*2, elem stride 2 (humanoid.py:207-210) */
  int64_t dof_elem_stride;
  const float* dof_force;    /* [N,69] dof_force_tensor (humanoid.py:189-190); NULL disables the power term */
  int64_t dof_force_stride;
  /* optional env subset: warp i processes env env_ids[i] (reset path, humanoid_im.py:677-681); NULL = envs 0..n-1 */
  const int64_t* env_ids;
  /* task buffers (read-only) */
  const int64_t* progress_buf;      /* [N] */
  const int64_t* motion_ids;        /* [N] _sampled_motion_ids */
  const float* motion_start_times;  /* [N] */
  const float* motion_start_offset; /* [N] _motion_start_times_offset */
  const float* global_offset;       /* [N,3] */
  const int32_t* cycle_counter;     /* [N] or NULL */
  const int64_t* reset_buf_in;      /* unused by the reference's formula (kept for signature parity); may be NULL */
  const float* termination_distances; /* [24] _termination_distances (humanoid_im.py:1166-1186) */
  uint32_t reset_body_mask;         /* bit j set = body j in reset_bodies (env_im.yaml:38) */
  uint32_t flags;                   /* PULSE_STEP_* */
  float dt;                         /* control dt, fp32(2/60) */
  float k_pos, k_rot, k_vel, k_ang_vel, w_pos, w_rot, w_vel, w_ang_vel; /* reward_specs humanoid_im.py:55 */
  float power_coefficient;          /* humanoid_im.py:91; used when dof_force != NULL */
  int32_t cycle_motion;             /* 0: pass_time = t >= motion_len; 1: progress >= max_episode_length-1 */
  int64_t max_episode_length;
  int32_t enable_early_termination;
  int32_t use_mean_reset;           /* flags.im_eval && !strict_eval (humanoid_im.py:1606) */
  /* outputs (any may be NULL when its stage is disabled) */
  float* obs_buf;        /* [N, obs_stride], first 934 floats written */
  int64_t obs_stride;
  float* self_obs_buf;   /* [N,358] optional copy (humanoid_im.py:683) */
  float* rew_buf;        /* [N] */
  float* reward_raw;     /* [N, raw_stride]: pos, rot, vel, ang_vel (, power) */
  int64_t raw_stride;
  int64_t* reset_buf;    /* [N] */
  int64_t* terminate_buf;/* [N] */
  uint8_t* pass_time;    /* [N] optional: t >= motion_len mask (needed by the cycle_motion host path) */
  float* ref_body_pos;   /* [N,24,3] optional (humanoid_im.py:835-848) */
  float* ref_body_vel;   /* [N,24,3] optional */
  float* ref_body_rot;   /* [N,24,4] optional */
  float* ref_dof_pos;    /* [N,69]   optional (costs 2 extra 960-byte gathers per env) */
  /* ---- ABI 2 ---- */
  const int32_t* env_count;        /* optional DEVICE-side length of env_ids (the compacted list pulse_reset_ref_state writes): only the
                                      first min(num_envs, *env_count) entries are processed -- no host read of the count is ever needed */
  const int32_t* recovery_counter; /* [N] or NULL.  HumanoidImGetup._compute_reset (humanoid_im_getup.py:203-210): for envs with
                                      recovery_counter > 0 the reset / terminate outputs are forced to 0 and progress_buf is decremented
                                      (through progress_rw) BEFORE the observation time is formed */
  int64_t* progress_rw;            /* writable alias of progress_buf; required with recovery_counter */
  float* fdones_out;               /* [N] optional float copy of reset_buf (the experience buffer's `dones`, amp_agent.py:383) */
} pulse_im_step_args_t;
/* num_envs = number of envs processed (= len(env_ids) when env_ids is given). */
int pulse_im_step(const pulse_motionlib_t* lib, const pulse_im_step_args_t* args, int64_t num_envs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * AMP observation + history shift.
 *   HumanoidAMP._update_hist_amp_obs      phc/env/tasks/humanoid_amp.py:622-630
 *   HumanoidAMP._compute_amp_observations humanoid_amp.py:632-667 (build_amp_observations_smpl :924-969,
 *                                         dof_to_obs_smpl humanoid.py:1436-1446), has_dof_subset = True.
 * amp_obs_buf is [N, num_steps, 196], index 0 = newest; in place: buf[:,1:] <- buf[:,:-1]; buf[:,0] <- new.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* body_state; int64_t body_env_stride;
  const float* dof_pos; const float* dof_vel; int64_t dof_env_stride; int64_t dof_elem_stride;
  float* amp_obs_buf;   /* [N, num_steps, 196] */
  int32_t num_steps;    /* numAMPObsSteps (env_im.yaml:31) */
  int32_t shift_history;/* 1: shift then write slot 0;  0: write slot 0 only */
} pulse_amp_obs_args_t;
int pulse_amp_obs(const pulse_amp_obs_args_t* args, int64_t num_envs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Rollout glue of AMPAgent.play_steps (phc/learning/amp_agent.py:341-439), written straight into the experience-buffer slices.
 *   pulse_policy_post   get_action_values' sampling (common_agent.py:262-288; ModelA2CContinuousLogStd [rl_games]): action = mu +
 *                       exp(logstd) * eps, neglogp, de-normalised value (running_mean_std.py:84-87), optional PD targets
 *                       (Humanoid._action_to_pd_targets, humanoid.py:1392-1394).  eps: injected, or Philox4x32-10(seed,
 *                       row, *rng_offset + rng_step) drawn in the kernel (a device-side offset keeps CUDA-graph replays fresh).
 *   pulse_value_post    next_values = unnormalise(critic(next obs)) * (1 - terminated)   (amp_agent.py:396-398)
 *   pulse_amp_obs_row   AMP observation row of this step = [current 196 | first (steps-1)*196 floats of the previous row], written
 *                       into its experience slice (humanoid_amp.py:622-667 + amp_agent.py:385)
 *   pulse_bump_counter  *counter += by (one thread): advances the device-side RNG offset once per iteration
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* mu; int64_t ld_mu;            /* [rows, A] actor head output */
  const float* logstd;                       /* [A] */
  const float* eps; int64_t ld_eps;          /* [rows, A] injected standard-normal draws, or NULL -> Philox */
  uint64_t seed; const uint64_t* rng_offset; uint64_t rng_step;
  int32_t num_actions; int32_t reserved;
  float* actions; int64_t ld_actions;        /* out */
  float* neglogp; int64_t ld_neglogp;        /* out, element stride */
  float* mus_out; int64_t ld_mus;            /* optional copy of mu (NULL when the head GEMM already wrote the experience slice) */
  const float* value; int64_t ld_value;      /* [rows] normalised critic output (optional) */
  const double* value_mean; const double* value_var; float value_eps; int32_t reserved2;   /* RunningMeanStd of the value (NULL = identity) */
  float* values_out; int64_t ld_values;      /* optional */
  const float* pd_offset; const float* pd_scale; float* pd_targets; int64_t ld_pd;   /* optional */
} pulse_policy_post_args_t;
int pulse_policy_post(const pulse_policy_post_args_t* args, int64_t rows, void* stream);
int pulse_value_post(const float* value, int64_t ld_value, const double* mean, const double* var, float eps, const int64_t* terminate,
                     float* out, int64_t ld_out, int64_t rows, void* stream);
typedef struct {
  const float* body_state; int64_t body_env_stride;
  const float* dof_pos; const float* dof_vel; int64_t dof_env_stride; int64_t dof_elem_stride;
  const float* prev; int64_t ld_prev;        /* previous step's row of every env (floats between envs) */
  float* out; int64_t ld_out;                /* this step's row */
  int32_t num_steps; int32_t reserved;
  int32_t* fresh;                            /* [N] optional flags set by pulse_reset_ref_state; cleared here */
  const float* fresh_rows;                   /* [N, num_steps, 196] the back-filled rows of reset envs */
} pulse_amp_row_args_t;
int pulse_amp_obs_row(const pulse_amp_row_args_t* args, int64_t num_envs, void* stream);
int pulse_bump_counter(uint64_t* counter, uint64_t by, void* stream);
/* Timing events that stay readable when the launches around them are captured into a CUDA graph (cudaEventRecordExternal): bench.py
 * times the fused step kernel live inside a whole-rollout graph with these. */
int pulse_event_create(void** event);
int pulse_event_destroy(void* event);
int pulse_event_record(void* event, void* stream);
int pulse_event_elapsed_ms(void* start, void* stop, float* ms);

/* ------------------------------------------------------------------------------------------------
 * Per-step env reset, fused and free of host synchronisation.  Replaces, for the envs whose
 * reset_buf is set (mask mode: the `done_indices` of AMPAgent.play_steps, phc/learning/amp_agent.py:352 -> env_reset ->
 * VecTaskPythonWrapper.reset -> Humanoid.reset, humanoid.py:526-541) or for an explicit id list:
 *   HumanoidIm._reset_ref_state_init     phc/env/tasks/humanoid_im.py:921-948   (start offset / global offset / cycle counter <- 0)
 *   HumanoidAMP._reset_ref_state_init    humanoid_amp.py:468-488, _sample_ref_state humanoid_im.py:966-989
 *   MotionLibBase.sample_time_interval   phc/utils/motion_lib_base.py:411-420   (uniform draw injected or Philox)
 *   HumanoidAMP._set_env_state           humanoid_amp.py:565-597  (root 13, dof pos / vel 69, rigid bodies 24 x 13, written in place
 *                                        into the Isaac Gym views; the rigid-body write is the reference's own post-refresh hack :604-614)
 *   Humanoid._reset_env_tensors          humanoid.py:589-609      (progress / reset / terminate <- 0, contact forces <- 0, and the
 *                                        int32 actor-id list for gym.set_*_tensor_indexed, built on the device)
 *   HumanoidAMP._init_amp_obs            humanoid_amp.py:519-563  (current AMP observation + the num_steps-1 history frames at
 *                                        t - k*dt from the reference motion, no offset)
 * Two launches (ordered compaction; one warp per (env, history step)).  The observation of the reset envs follows with
 * pulse_im_step(flags = PULSE_STEP_OBS, env_ids = env_list, env_count = count).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t* reset_buf;            /* [N] mask mode: envs with reset_buf != 0 are reset; cleared for them afterwards */
  const int64_t* env_ids_in;     /* list mode: explicit env ids [num_ids] (Humanoid.reset(env_ids)); NULL = mask mode */
  int64_t num_ids;
  const float* phase;            /* [N] uniform [0,1) draw per ENV (tests inject it) or NULL: Philox4x32-10(seed, env, offset) */
  uint64_t seed, offset;
  const int64_t* motion_ids;     /* [N] _sampled_motion_ids (each env keeps its clip, humanoid_im.py:966-989) */
  float* motion_start_times;     /* [N] <- sampled start time */
  float* motion_start_offset;    /* [N] <- 0 */
  float* global_offset;          /* [N,3] <- 0 */
  int32_t* cycle_counter;        /* [N] <- 0 (may be NULL) */
  int64_t* progress_buf;         /* [N] <- 0 */
  int64_t* terminate_buf;        /* [N] <- 0 (may be NULL) */
  float* root_states; int64_t root_env_stride;           /* _humanoid_root_states: env e at root_states + e*root_env_stride, 13 floats */
  float* dof_pos; float* dof_vel; int64_t dof_env_stride; int64_t dof_elem_stride;   /* Isaac Gym dof-state views (elem stride 2) */
  float* rigid_body_state; int64_t body_env_stride;      /* [N, bodies_per_env, 13]; may be NULL */
  float* contact_forces; int64_t contact_env_stride; int32_t contact_bodies;   /* [N, bodies_per_env, 3] <- 0; may be NULL */
  int32_t num_amp_steps;         /* numAMPObsSteps; 0 with amp_obs_buf NULL */
  float* amp_obs_buf;            /* [N, num_amp_steps, 196]; may be NULL */
  float dt;                      /* control dt */
  int32_t reserved;
  const int32_t* actor_ids;      /* [N] _humanoid_actor_ids (humanoid.py:590) or NULL */
  int64_t* env_list;             /* [N] out: the reset env ids, ascending (what `nonzero` returns) */
  int32_t* actor_list;           /* [N] out: actor ids of those envs (argument of gym.set_*_tensor_indexed); may be NULL */
  int32_t* count;                /* [1] out, device side: number of reset envs */
  int32_t* amp_fresh;            /* [N] optional: set to 1 for every reset env -- tells pulse_amp_obs_row to take that env's history from
                                    amp_obs_buf (the back-filled rows) at the next step */
  const uint64_t* offset_dev;    /* optional device-side counter added to `offset` (fresh draws on every CUDA-graph replay) */
} pulse_reset_args_t;
int pulse_reset_ref_state(const pulse_motionlib_t* lib, const pulse_reset_args_t* args, int64_t num_envs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GAE / returns.  CommonAgent.discount_values  phc/learning/common_agent.py:493-505, mb_returns =
 * mb_advs + mb_values (amp_agent.py:427), and the first half of _calc_advs (:589-599): sums for the
 * advantage mean / unbiased std.  Inputs are [T,N] time-major as in the rl_games ExperienceBuffer;
 * outputs are written ENV-MAJOR [N,T] (swap_and_flatten01 layout) ready for minibatch slicing.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* rewards;      /* [T,N] */
  const float* values;       /* [T,N] */
  const float* next_values;  /* [T,N] already multiplied by (1-terminated), amp_agent.py:396-398 */
  const float* fdones;       /* [T,N] 0/1 */
  float gamma, tau;
  float* advantages;         /* [N,T] env-major */
  float* returns;            /* [N,T] env-major */
  double* adv_sum;           /* [2]: sum(adv), sum(adv^2) accumulated with atomics; caller zeroes; may be NULL */
} pulse_gae_args_t;
int pulse_gae(const pulse_gae_args_t* args, int32_t horizon, int64_t num_envs, void* stream);
/* advantages <- (adv - mean) / (std + 1e-8) with unbiased std from adv_sum (common_agent.py:596-597) */
int pulse_normalize_advantages(float* advantages, const double* adv_sum, int64_t count, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense layers on the tensor cores: D[M,N] = epilogue(alpha * A[M,K] . B[N,K]^T), bf16 operands (both
 * K-major: row-major with the reduction dimension contiguous), fp32 accumulation in TMEM (tcgen05).
 * Replaces the nn.Linear + activation stacks of the policy / value / discriminator / VAE networks
 * (phc/learning/network_builder.py:105-124, amp_network_builder.py:58-249, amp_network_z_builder.py:341-467)
 * and their autograd backward: forward (A=X, B=W), dgrad (A=dY, B=W^T), wgrad (A=dY^T, B=X^T).
 * ---------------------------------------------------------------------------------------------- */
typedef uint16_t pulse_bf16_t;  /* raw bfloat16 bits */
#define PULSE_ACT_NONE 0
#define PULSE_ACT_RELU 1
#define PULSE_ACT_SILU 2

typedef struct {
  const float* bias;         /* [N] added before the activation, or NULL */
  int32_t act;               /* PULSE_ACT_* applied to (alpha*acc + bias) */
  int32_t gate_mode;         /* PULSE_ACT_RELU / PULSE_ACT_SILU: multiply by act'(gate) (backward through the activation) */
  const pulse_bf16_t* gate;  /* [M, ldg] saved tensor: ReLU -> the layer OUTPUT, SiLU -> the PRE-activation; NULL = off */
  int64_t ldg;
  float alpha;
  pulse_bf16_t* out;         /* [M, ldo] bf16 row-major, or NULL */
  int64_t ldo;
  pulse_bf16_t* out_t;       /* [N, ldot] bf16 TRANSPOSED copy (operand of the next wgrad), or NULL */
  int64_t ldot;
  float* out_f32;            /* [M, ldf] fp32 (heads, weight-gradient slabs), or NULL */
  int64_t ldf;
  int64_t split_stride;      /* floats between split-K slabs of out_f32 */
  pulse_bf16_t* preact;      /* [M, ldp] bf16 pre-activation (saved for SiLU backward), or NULL */
  int64_t ldp;
  float* colsum;             /* [N] += column sums of the final values (bias gradient of the layer whose dY this GEMM writes), or NULL */
  int32_t accumulate;        /* 1: out_f32 += result with fp32 atomics (weight gradients; caller zeroes), 0: overwrite */
  int32_t reserved;
  double* sumsq;             /* *sumsq += sum of squares of the final values over the valid [M,N] region (fp64 atomics), or NULL */
  /* ---- ABI 2: ReLU masks as bit words.  Forward: bit i of relu_mask[(col/32) * ld_rmask + row] = (pre-activation of column
   * 32*(col/32)+i of `row`) > 0, written next to the bf16 activations (2 MB instead of the 32 MB the backward pass used to re-read
   * for a 16384 x 1024 layer).  ReLU-dgrad: gate_mask in the same layout replaces `gate`: one coalesced 4-byte load per
   * (row, 32-column chunk).  Chunk-major ([ceil(N/32), ld] words, ld >= M) so that consecutive rows are consecutive words. */
  uint32_t* relu_mask; int64_t ld_rmask;
  const uint32_t* gate_mask; int64_t ld_gmask;
} pulse_gemm_epilogue_t;

#define PULSE_GEMM_A_MN 1u   /* A is given as [K, M] row-major (the reduction dimension is the ROW index) */
#define PULSE_GEMM_B_MN 2u   /* B is given as [K, N] row-major */

/* lda / ldb in elements, multiples of 8, >= K; A and B 16-byte aligned.  split_k > 1: fp32 slabs only
 * (slab z at out_f32 + z*split_stride); pulse_gemm_num_splits gives the number of slabs actually written. */
int pulse_gemm_bf16_nt(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n, int64_t k,
                       const pulse_gemm_epilogue_t* ep, int32_t split_k, void* stream);
/* General form: D[M,N] = epilogue(sum_k A(m,k) B(n,k)).  flags select, per operand, K-major storage (A[M,K] / B[N,K],
 * the NT case above) or MN-major storage (A[K,M] / B[K,N] row-major), so activations, output gradients and weights
 * are consumed exactly as they sit in memory:  dgrad dX = dY . W  -> A = dY (K-major), B = W [N_out,K_in] as MN-major;
 * wgrad dW = dY^T X -> A = dY [batch,N] MN-major, B = X [batch,K] MN-major.  No transposed copies anywhere. */
int pulse_gemm_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n, int64_t k,
                    const pulse_gemm_epilogue_t* ep, int32_t split_k, uint32_t flags, void* stream);
int pulse_gemm_num_splits(int64_t k, int32_t split_k);

/* Several GEMMs of the same kind in ONE persistent launch (work items of all problems concatenated): forward groups
 * (flags 0), ReLU-dgrad groups (PULSE_GEMM_B_MN) or weight-gradient groups (PULSE_GEMM_A_MN | PULSE_GEMM_B_MN, fp32 atomic
 * accumulation), at most 4 problems.  EXPERIMENTAL in round 1 (compiled, not yet validated on a device). */
typedef struct {
  const void* a; int64_t lda;
  const void* b; int64_t ldb;
  int64_t m, n, k;
  pulse_gemm_epilogue_t ep;
  int32_t split_k, reserved;
} pulse_gemm_problem_t;
int pulse_gemm_bf16_grouped(const pulse_gemm_problem_t* problems, int32_t count, uint32_t flags, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Element-wise / reduction kernels around the GEMMs.
 * ---------------------------------------------------------------------------------------------- */
/* RunningMeanStd.forward, eval path (phc/utils/running_mean_std.py:69-95): y = clamp((x-mean)*rstd, -5, 5),
 * written as bf16 [rows, ld_out] (columns >= cols zero-filled up to ld_out) and optionally transposed
 * bf16 [ld_out, ld_t] (operand of the first layer's wgrad).  mean / rstd: fp32 [cols] (rstd = 1/sqrt(var+eps),
 * prepared by the caller from the fp64 statistics); NULL mean = plain cast.  pad_one: value of the FIRST pad column (index cols) when
 * ld_out > cols -- 1.0 makes it the "ones" column of a bias-augmented GEMM operand (the layer's bias then sits in column `cols` of its
 * weight matrix: bias add and bias gradient ride the tensor cores), 0.0 = plain zero fill. */
int pulse_normalize_to_bf16(const float* x, int64_t ldx, int64_t rows, int64_t cols, const float* mean, const float* rstd,
                            pulse_bf16_t* out, int64_t ld_out, pulse_bf16_t* out_t, int64_t ld_t, float pad_one, void* stream);

/* Batch moments for RunningMeanStd's training-mode update (:96-107): per-column sum and sum of squares of
 * fp32 x [rows, cols] accumulated in fp64 into sums[2*cols] (caller zeroes). */
int pulse_column_moments(const float* x, int64_t ldx, int64_t rows, int64_t cols, double* sums, void* stream);

/* pulse_normalize_to_bf16 + pulse_column_moments in ONE pass over x: RunningMeanStd.forward in training mode
 * (phc/utils/running_mean_std.py:91-107) normalises with the statistics from before the batch and merges the
 * batch afterwards, so both read the same rows.  out [rows, ld_out] (padding columns zeroed), sums[2*cols] += . */
int pulse_normalize_moments(const float* x, int64_t ldx, int64_t rows, int64_t cols, const float* mean, const float* rstd,
                            pulse_bf16_t* out, int64_t ld_out, double* sums, float pad_one, void* stream);

/* RunningMeanStd._update_mean_var_count_from_moments (:54-66) on the device: merges the batch sums of
 * pulse_column_moments (n rows) into the fp64 running mean / var / count and refreshes the fp32 mean / rstd
 * vectors pulse_normalize_to_bf16 reads, then zeroes `sums` for the next batch.  One launch, no host round trip. */
int pulse_rms_merge(double* sums, int64_t n, int32_t size, double* mean, double* var, double* count, float eps,
                    float* mean_f32, float* rstd_f32, void* stream);

/* Single-output head (the critic's `value` Linear, network_builder.py:171; the discriminator's `_disc_logits`,
 * amp_network_builder.py:245-249): out[m] = h[m,:] . w + bias.  h bf16 [rows, k] (row stride ldh), w bf16 [k]. */
int pulse_head1_forward(const pulse_bf16_t* h, int64_t ldh, int64_t rows, int32_t k, const pulse_bf16_t* w, const float* bias, float* out,
                        int64_t ldo, void* stream);

/* Backward of that head through the ReLU below it, ONE pass over h:  dh[m,j] = dv[m] w[j] (h[m,j] > 0) (bf16, may be
 * NULL);  dw[j] += sum_m dv[m] h[m,j];  db += sum_m dv[m];  dbias_prev[j] += sum_m dh[m,j] (bias gradient of the layer
 * that produced h; may be NULL).  k <= 2048, multiple of 8. */
int pulse_head1_backward(const pulse_bf16_t* h, int64_t ldh, int64_t rows, int32_t k, const pulse_bf16_t* dv, int64_t ld_dv,
                         const pulse_bf16_t* w, pulse_bf16_t* dh, int64_t ld_dh, float* dw, float* db, float* dbias_prev, void* stream);

/* Gaussian policy head (rl_games ModelA2CContinuousLogStd, fixed sigma: im.yaml:21-25):
 * actions = mu + exp(logstd)*eps;  neglogp = 0.5*sum(((a-mu)/sigma)^2) + 0.5*A*log(2*pi) + sum(logstd). */
int pulse_gaussian_sample(const float* mu, int64_t ld_mu, const float* eps, const float* logstd, int64_t rows, int32_t num_actions,
                          float* actions, float* neglogp, void* stream);

/* PPO actor / critic / bound losses and their gradients w.r.t. the network outputs, one pass
 * (common_agent.py:512-520, :564-587; amp_agent.py:691-710; torch_ext.policy_kl):
 *   a = max(-A r, -A clip(r, 1-e, 1+e)), r = exp(old_neglogp - neglogp);  c = (ret - v)^2;
 *   b = sum(clamp_min(mu-1,0)^2 + clamp_max(mu+1,0)^2);  loss = mean(a) + critic_coef*mean(c) + bounds_coef*mean(b).
 * Outputs: dmu bf16 [rows, ld_dmu] (+ transposed [A_pad, ld_t]), dvalue bf16 [rows, ld_dv] (+ transposed),
 * stats[0..5] fp64 accumulators: sum a, sum c, sum b, sum kl, clipped count, sum neglogp (caller zeroes). */
typedef struct {
  const float* mu; int64_t ld_mu;       /* [rows, A] network output */
  const float* value; int64_t ld_value; /* [rows, 1] */
  const float* actions;                 /* [rows, A] contiguous */
  const float* old_neglogp;             /* [rows] */
  const float* advantages;              /* [rows] */
  const float* returns;                 /* [rows] (already value-normalised) */
  const float* old_mu;                  /* [rows, A] contiguous, for the KL statistic; may be NULL */
  const float* logstd;                  /* [A] */
  int32_t num_actions;
  float e_clip, critic_coef, bounds_coef;
  pulse_bf16_t* dmu; int64_t ld_dmu; pulse_bf16_t* dmu_t; int64_t ld_dmu_t;
  pulse_bf16_t* dvalue; int64_t ld_dv; pulse_bf16_t* dvalue_t; int64_t ld_dv_t;
  double* stats;
} pulse_ppo_loss_args_t;
int pulse_ppo_loss(const pulse_ppo_loss_args_t* args, int64_t rows, void* stream);

/* AMP discriminator loss pieces (AMPAgent._disc_loss, phc/learning/amp_agent.py:895-952):
 *  - prediction loss 0.5*(BCE(agent U replay, 0) + BCE(demo, 1)) and its gradient w.r.t. the logits
 *    (rows [0,n_agent) agent/replay, rows [n_agent, n_agent+n_demo) demo), scaled by `scale` (= disc_coef);
 *    stats[0..3] += sum softplus(l) agent, sum softplus(-l) demo, #agent l<0, #demo l>0 (accuracies, :954-959);
 *  - pulse_relu_mask_scale: out = (h > 0) * w, the first factor of the ANALYTIC input gradient of the ReLU
 *    discriminator used for the gradient penalty (:910-929) instead of autograd's double backward;
 *  - pulse_axpy: y += a*x (logit regulariser :905-908 and weight decay :932-937 gradients, 2*coef*w). */
int pulse_disc_loss(const float* logits, int64_t ld, int64_t n_agent, int64_t n_demo, float scale, pulse_bf16_t* dlogit, int64_t ld_d,
                    double* stats, void* stream);
int pulse_relu_mask_scale(const pulse_bf16_t* h, int64_t ldh, int64_t rows, int64_t cols, const float* w, pulse_bf16_t* out, int64_t ldo,
                          void* stream);
int pulse_axpy(float a, const float* x, float* y, int64_t count, void* stream);
/* Weight decay / logit regulariser gradients + the sums of squares of the logged terms for up to four weight blocks in one launch
 * (AMPAgent._disc_loss, amp_agent.py:905-908, :932-937): g[r,c] += coef * w[r,c] for c < cols of a [rows, ld] matrix; *sumsq (and
 * *sumsq2) += sum w^2 in fp64.  g / sumsq / sumsq2 may be NULL. */
typedef struct {
  const float* w; float* g; int64_t rows, cols, ld; float coef; int32_t reserved; double* sumsq; double* sumsq2;
} pulse_weight_block_t;
typedef struct { pulse_weight_block_t block[4]; int32_t count; int32_t reserved; } pulse_weight_reg_t;
int pulse_weight_reg(const pulse_weight_reg_t* desc, void* stream);

/* out[c] (+)= sum over rows of bf16 x[rows, ldx] (bias gradients). */
int pulse_column_sum_bf16(const pulse_bf16_t* x, int64_t ldx, int64_t rows, int64_t cols, float* out, void* stream);
/* dst[i] = sum_s slabs[s*slab_stride + i]  (split-K weight-gradient slabs -> flat gradient buffer) */
int pulse_reduce_slabs(const float* slabs, int64_t slab_stride, int32_t num_slabs, int64_t count, float* dst, void* stream);
/* sumsq[0] += sum(x^2) in fp64 (global gradient norm; caller zeroes) */
int pulse_sum_squares(const float* x, int64_t count, double* sumsq, void* stream);
/* clip_grad_norm_(max_norm) + Adam step over one flat parameter buffer (amp_agent.py:725-750; torch.optim.Adam
 * defaults beta 0.9/0.999): scale = min(1, max_norm/(sqrt(sumsq)+1e-6)) read on the device, no host sync. */
/* `step` is a DEVICE counter (int32[1]) incremented by this call, so the launch sequence is CUDA-graph replayable.
 * params_bf16 (optional, same flat layout): bf16 copy of the updated parameters = the GEMM operands. */
#define PULSE_ADAM_ZERO_GRADS 1u      /* the kernel zeroes every gradient it has consumed (the next minibatch accumulates from zero) */
#define PULSE_ADAM_SELF_CONTAINED 2u  /* no bump / memset launches: this launch is step *step + 1; its last block stores the new step and
                                         re-zeroes *grad_sumsq (block_counter: one zero-initialised uint32 owned by the optimizer) */
int pulse_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t count, double* grad_sumsq,
                    float max_norm, float lr, float beta1, float beta2, float eps, int32_t* step, pulse_bf16_t* params_bf16, uint32_t flags,
                    uint32_t* block_counter, void* stream);
/* Multi-GPU optimizer step over NVLink peer memory (csrc/peer_adam.cu): the reference's Horovod gradient averaging
 * (hvd.DistributedOptimizer, amp_agent.py:735-742) + clip_grad_norm_ + torch.optim.Adam (amp_agent.py:725-750) as ONE kernel per rank --
 * reduce-scatter of the flat gradient buffers by peer loads (or multimem.ld_reduce), exchange of the slice norms, Adam on the rank's
 * slice (sharded moments), push of the new fp32 masters + bf16 operands into every rank's buffers (peer stores or multimem.st), clearing
 * of the gradients.  Replaces pulse_sum_squares + pulse_adam_step + the NCCL all-reduce when the ranks' buffers are peer-mapped.
 *   grads / params / params_bf16 / signals [p]: rank p's buffer as mapped into THIS process ([rank] = the local one).  A signal block
 *   is PULSE_PEER_SIGNAL_BYTES of zero-initialised memory: uint32 flags[3][PULSE_PEER_MAX] then double norms[PULSE_PEER_MAX].
 *   mc_*: multicast aliases of the same buffers (NVLS), or NULL.  exp_avg / exp_avg_sq: local, full size; only this rank's slice
 *   [rank * ceil(count/4/world) * 4, ...) is read or written.  step: device Adam step counter (incremented).  epoch: device uint32[1]
 *   call counter, zero-initialised, same value on every rank.  cta_partials: double[PULSE_PEER_MAX_GRID]; grid_bar: uint64[1] zero-
 *   initialised; grid: CTAs (0 = one per SM) -- must not change between calls that share grid_bar.
 * Every rank must make the call (it waits for its peers, bounded by timeout_ms, then the launch fails); count is a multiple of 4. */
#define PULSE_PEER_MAX 8
#define PULSE_PEER_MAX_GRID 256
#define PULSE_PEER_SIGNAL_BYTES (3 * PULSE_PEER_MAX * 4 + PULSE_PEER_MAX * 8)
typedef struct {
  int32_t rank, world;
  float* grads[PULSE_PEER_MAX];
  float* params[PULSE_PEER_MAX];
  pulse_bf16_t* params_bf16[PULSE_PEER_MAX];
  uint32_t* signals[PULSE_PEER_MAX];
  const float* mc_grads; float* mc_params; pulse_bf16_t* mc_params_bf16;
  float* exp_avg; float* exp_avg_sq;
  int64_t count;
  float max_norm, lr, beta1, beta2, eps;
  int32_t grid;
  uint32_t timeout_ms;                     /* bound of every wait on a peer (0 = 30 min: ranks drift apart around rank-0-only work) */
  uint32_t reserved;
  int32_t* step;
  uint32_t* epoch;
  double* cta_partials;
  unsigned long long* grid_bar;
} pulse_peer_adam_args_t;
int pulse_peer_reduce_adam(const pulse_peer_adam_args_t* args, void* stream);

/* refresh the bf16 operand copies of one weight matrix W fp32 [n, k] (contiguous):
 *   w_bf16 [n, ld_k] (K-major, forward / wgrad-free) and wt_bf16 [k, ld_n] (transposed, dgrad operand); pads zeroed. */
int pulse_refresh_weight_bf16(const float* w, int64_t n, int64_t k, pulse_bf16_t* w_bf16, int64_t ld_k, pulse_bf16_t* wt_bf16,
                              int64_t ld_n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PULSE VAE distillation, Z-task decode (K20), reach task (K21), PD targets (K22).
 * The dense layers run on pulse_gemm_bf16; these are the row-wise pieces between them.
 * ---------------------------------------------------------------------------------------------- */
/* y = (x - mean) * rstd, optionally clamped to [-clamp, clamp] (clamp <= 0: no clamp -- HumanoidZ.compute_z_actions feeds the
 * prior the UNCLAMPED normalised self observation, humanoid_z.py:87 vs :147), written as bf16 into out[rows, 0:cols];
 * columns [cols, zero_to) of each out row are zero-filled, nothing beyond is touched (out may be a column window of a wider
 * operand buffer).  mean/rstd NULL = plain cast. */
int pulse_normalize_cols(const float* x, int64_t ldx, int64_t rows, int64_t cols, const float* mean, const float* rstd, float clamp,
                         pulse_bf16_t* out, int64_t ld_out, int64_t zero_to, void* stream);
/* dst1[r, 0:cols] = dst2[r, 0:cols] = src[r, 0:cols] (bf16; dst2 may be NULL): the normalised self-observation columns of
 * the encoder input feed the prior MLP and the decoder input window (amp_network_z_builder.py:229, :443-445). */
int pulse_copy_cols_bf16(const pulse_bf16_t* src, int64_t ld_src, int64_t rows, int64_t cols, pulse_bf16_t* dst1, int64_t ld1,
                         pulse_bf16_t* dst2, int64_t ld2, void* stream);

/* Latent sample (form_embedding / reparameterize, amp_network_z_builder.py:79-121, :243-246).  head fp32 [rows, >= 2*latent]:
 * columns [0, latent) = mu, [latent, 2*latent) = raw log-variance (clamped to [clamp_lo, clamp_hi] when clamp != 0).
 *   mode PULSE_Z_SAMPLE: z = mu + exp(0.5*logvar)*noise;  PULSE_Z_MEAN: z = mu (flags.test, :94-95);
 *   PULSE_Z_RESIDUAL:    z = mu + noise  (HumanoidZ.compute_z_actions: prior_mu + action_z, humanoid_z.py:104-107).
 * z is written as bf16 into z_bf16[rows, 0:latent] (the decoder input window) and/or fp32 z_f32[rows, latent]. */
#define PULSE_Z_SAMPLE 0
#define PULSE_Z_MEAN 1
#define PULSE_Z_RESIDUAL 2
int pulse_vae_reparam(const float* head, int64_t ld_head, const float* noise, int64_t ld_noise, int64_t rows, int32_t latent,
                      int32_t mode, int32_t clamp, float clamp_lo, float clamp_hi, pulse_bf16_t* z_bf16, int64_t ld_z, float* z_f32,
                      int64_t ld_zf, void* stream);

/* kin_action_loss = mean_rows ||pred - gt||_2 (amp_agent.py:782): stats[0] += sum of row norms (fp64; caller zeroes);
 * dpred bf16 [rows, ld_d] = (pred - gt) / (||pred - gt|| * rows) (0 where the norm is 0, as torch.norm's backward), columns
 * [num_actions, zero_to) zero-filled. */
int pulse_vae_action_loss(const float* pred, int64_t ld_pred, const float* gt, int64_t ld_gt, int64_t rows, int32_t num_actions,
                          pulse_bf16_t* dpred, int64_t ld_d, int64_t zero_to, double* stats, void* stream);

/* Latent-space terms of AMPAgent._optimize_kin (amp_agent.py:784-816) and their gradients w.r.t. the encoder / prior heads,
 * one warp per row (latent <= 32):
 *   KLD  = mean_rows kl_multi(q || p)                                   (loss_functions.py:3-11)      * kld_coef
 *   AR1  = mean over (rows/horizon)*(horizon-1) pairs of ||mu[t+1] - phi mu[t]||, pairs masked where the progress counter is
 *          not consecutive or either step has progress <= 2 (:792-808)                                   * ar1_coef
 *   REGU = 0.001*(mean pm^2 + mean qm^2 + mean pv^2 + mean qv^2) (:810-814)                              * regu_coef
 * plus the reparameterisation path of dz = dLoss/dz (from the decoder's input gradient):  dmu += dz,
 * dlogvar += dz * 0.5*exp(0.5*logvar)*noise; clamp gates (gradient passes where lo <= raw <= hi).
 * Rows are env-major [rows/horizon, horizon] (amp_datasets.py:54-79).  progress NULL or ar1_coef == 0: no AR(1) term.
 * stats (fp64, caller zeroes): [0] sum KL rows, [1] sum AR1 pair norms, [2] sum pm^2, [3] sum qm^2, [4] sum pv^2, [5] sum qv^2. */
typedef struct {
  const float* enc_head; int64_t ld_enc;       /* [rows, >= 2*latent]: mu | raw logvar of the posterior */
  const float* prior_head; int64_t ld_prior;   /* [rows, >= 2*latent]: mu | raw logvar of the prior */
  const float* noise; int64_t ld_noise;        /* [rows, latent] */
  const float* dz; int64_t ld_dz;              /* [rows, latent] fp32, or NULL */
  const int64_t* progress;                     /* [rows] progress_buf recorded with the sample, or NULL */
  int32_t latent, horizon, clamp, reserved;
  float clamp_lo, clamp_hi, kld_coef, ar1_coef, regu_coef, phi;
  pulse_bf16_t* d_enc_head; int64_t ld_de;     /* [rows, 2*latent] gradient w.r.t. enc_head */
  pulse_bf16_t* d_prior_head; int64_t ld_dp;   /* [rows, 2*latent] gradient w.r.t. prior_head */
  double* stats;
} pulse_vae_latent_args_t;
int pulse_vae_latent_loss(const pulse_vae_latent_args_t* args, int64_t rows, void* stream);

/* Distillation teacher output (HumanoidImDistill.step, humanoid_im_distill.py:193-198): out[r, :] = sum_k act(w[r, k]) *
 * acts[k][r, :], acts = num_prim column outputs fp32 at acts + k*prim_stride, w = raw composer head [rows, num_prim];
 * act = PULSE_ACT_SILU for the composer rebuilt by load_mcp_mlp (network_loader.py:37-39). */
int pulse_pnn_compose(const float* acts, int64_t prim_stride, int64_t ld_a, const float* w, int64_t ld_w, int32_t act, int64_t rows,
                      int32_t num_actions, int32_t num_prim, float* out, int64_t ld_out, void* stream);

/* Humanoid._action_to_pd_targets + the freeze_hand / freeze_toe zeroing of pre_physics_step (humanoid.py:1222-1247,1392-1394):
 * out = freeze[d] ? 0 : offset[d] + scale[d]*action[r, d].  freeze: uint8 [dofs] or NULL. */
int pulse_pd_targets(const float* action, int64_t ld_a, const float* offset, const float* scale, const uint8_t* freeze, int64_t rows,
                     int32_t dofs, float* out, int64_t ld_out, void* stream);

/* HumanoidReach._update_task / _reset_task (humanoid_reach.py:126-147) with the uniform draws supplied by the caller:
 * where progress >= tar_change_steps: tar_pos = (dist_max*(2u-1), dist_max*(2v-1), h_min + (h_max-h_min)*w),
 * tar_change_steps = progress + steps.  rand01 fp32 [n, 3], steps int64 [n]. */
int pulse_reach_update_task(const int64_t* progress, int64_t* tar_change_steps, float* tar_pos, const float* rand01, const int64_t* steps,
                            float dist_max, float h_min, float h_max, int64_t num_envs, void* stream);

/* HumanoidReach post-physics step (humanoid_reach.py:149-166, :224-250; humanoid.py:1573-1608, :1675-1731), one warp per env:
 * reward = exp(-4 ||tar - reach_body||^2); reset / terminate = compute_humanoid_reset (contact force > 0.1 on a non-contact body
 * AND a non-contact body below its termination height, progress > 1; or progress >= max_episode_length - 1);
 * obs[env] = [self observation 358 | heading-frame target offset 3]. */
typedef struct {
  const float* body_state; int64_t body_env_stride;       /* [N, >=24, 13] pos quat(xyzw) linvel angvel */
  const float* contact_forces; int64_t contact_env_stride; /* [N, >=24, 3] or NULL (no early termination) */
  const float* termination_heights;                        /* [24] */
  const float* tar_pos;                                    /* [N, 3] */
  const int64_t* progress_buf;                             /* [N] */
  uint32_t contact_body_mask;                              /* bit j: body j may touch the ground (contact_bodies) */
  int32_t reach_body_id;
  int32_t enable_early_termination, reserved;
  int64_t max_episode_length;
  float* obs_buf; int64_t obs_stride;                      /* [N, 361] */
  float* rew_buf;                                          /* [N] */
  int64_t* reset_buf; int64_t* terminate_buf;              /* [N] */
} pulse_reach_step_args_t;
#define PULSE_REACH_OBS 361
int pulse_reach_step(const pulse_reach_step_args_t* args, int64_t num_envs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Post-physics step of the downstream latent-space tasks HumanoidSpeedZ / HumanoidStrikeZ, siblings of pulse_reach_step:
 * self observation (humanoid.py:1675-1731) + task observation + reward + reset in one launch.
 *   PULSE_ZTASK_SPEED   compute_speed_observations / compute_speed_reward (phc/env/tasks/humanoid_speed.py:310-343), power term
 *                       (:215-222; dof_force NULL = off), compute_humanoid_reset (humanoid.py:1573-1608).  obs 358 + 3.
 *   PULSE_ZTASK_STRIKE  compute_strike_observations / compute_strike_reward (humanoid_strike.py:270-328), the strike variant of
 *                       compute_humanoid_reset (:330-375).  obs 358 + 15.
 * prev_root_pos [N,3] = root position before the physics step (pre_physics_step, humanoid_speed.py:73-76).  Not covered: the
 * power_usage_reward terms (:224-240, humanoid_strike.py:186-198), the input-noise suffix (humanoid_speed.py:194-195).
 * ---------------------------------------------------------------------------------------------- */
#define PULSE_ZTASK_SPEED 1
#define PULSE_ZTASK_STRIKE 2
#define PULSE_SPEED_OBS 361
#define PULSE_STRIKE_OBS 373
typedef struct {
  int32_t kind, enable_early_termination;
  const float* body_state; int64_t body_env_stride;
  const float* contact_forces; int64_t contact_env_stride;      /* [N, B, 3] view or NULL */
  const float* termination_heights;                              /* [24] */
  uint32_t contact_body_mask;                                    /* bodies allowed to touch the ground (_contact_body_ids) */
  uint32_t strike_body_mask;                                     /* strike: bodies allowed to hit the target (_strike_body_ids) */
  const int64_t* progress_buf; int64_t max_episode_length;
  const float* prev_root_pos; float dt; float power_coefficient;
  const float* tar_speed;                                        /* speed: [N] */
  const float* target_states; int64_t target_env_stride;         /* strike: [N, 13] view of the target actor's root state */
  const float* tar_contact_forces; int64_t tar_contact_env_stride;   /* strike: [N, 3] view */
  const float* dof_force; int64_t dof_force_stride;              /* speed power term: [N, 69] */
  const float* dof_vel; int64_t dof_env_stride, dof_elem_stride;
  float* obs_buf; int64_t obs_stride; float* rew_buf; float* reward_raw; int64_t raw_stride;
  int64_t* reset_buf; int64_t* terminate_buf;
} pulse_ztask_step_args_t;
int pulse_ztask_step(const pulse_ztask_step_args_t* args, int64_t num_envs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Task observation for every observation version / tracked-body subset / number of future samples: replaces the
 * dispatch of HumanoidIm._compute_task_obs (phc/env/tasks/humanoid_im.py:757-833) over compute_imitation_observations (:1222-1258,
 * obs_v 1), _v2 (:1261-1301), _v3 (:1304-1326), _v6 (:1328-1378, obs_v 4 / 6), _v7 (:1381-1413), _v8 (:1415-1479, time_steps 1) and
 * _v9 (:1482-1540) on the `_track_bodies_id` rows.  The reference states are the outputs of pulse_motion_state for the N * time_steps
 * sample times in repeat_interleave order (row env * time_steps + t; fut_tracks, humanoid_im.py:723-729), full 24-body arrays.
 * Not covered: zero_out_far / occlusion rewrites (:763-784), the one-hot suffix of obs_v 5, the multi-sample branches of v2 / v8.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* body_state; int64_t body_env_stride;   /* [N, B>=24, 13] rigid-body state view */
  const int32_t* track_ids;                            /* [num_track] device array of body indices < 24; [0] = 0 for versions 2 and 9 */
  int32_t num_track, time_steps, version, upright;     /* upright = _has_upright_start */
  const float* ref_pos; const float* ref_rot; const float* ref_vel; const float* ref_ang_vel;   /* [N * time_steps, 24, 3 | 4] */
  const float* dof_pos; int64_t dof_env_stride, dof_elem_stride;   /* version 2: simulator dof positions (view strides) */
  const float* ref_dof_pos;                            /* version 2: [N, 69] */
  float* obs; int64_t obs_stride;                      /* [N, >= pulse_task_obs_size] */
  int64_t num_envs;
} pulse_task_obs_args_t;
int pulse_task_obs_size(int32_t version, int32_t num_track, int32_t time_steps);   /* floats per env, -1 for an unknown version */
int pulse_im_task_obs(const pulse_task_obs_args_t* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluation metrics on the device.  One call per evaluation step replaces the per-step bookkeeping of
 * IMAmpAgent._post_step_eval (phc/learning/im_amp.py:244-363: termination state :249-251, the curr_max stopping rule :252-268, :275,
 * the per-sequence `[:(num_steps - 1)]` frame slices :283-287) and the per-frame metrics compute_metrics_lite derives from the
 * frames the reference copies to the host every step (humanoid_im.py:664-673; smpl_sim [3P]): global / root-relative /
 * Procrustes-aligned MPJPE, velocity and acceleration errors -- accumulated into per-env fp64 sums (metres) and frame counts.
 *   ctrl int32[8], zeroed at the start of a chunk: [0] steps taken, [1] chunk finished (later calls are no-ops), [2..4] scratch.
 *   terminate_state int32[N], hist float[N,2,24,3], sums double[N,5] (mpjpe_g, mpjpe_l, mpjpe_pa, vel, accel), counts int32[N,3]
 *   (frames behind sums 0-2, 3, 4): zeroed at the start of a chunk.  bound: envs [0, bound) hold distinct clips (wrapped last chunk,
 *   im_amp.py:254-262), otherwise num_envs.  max_steps_all = max(num_steps).  mpjpe_out float[N] (optional): extras['mpjpe'].
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* body_pos; int64_t pos_env_stride, pos_body_stride;     /* body j of env e at body_pos + e*env_stride + j*body_stride */
  const float* body_pos_gt; int64_t gt_env_stride, gt_body_stride;    /* motion_res['rg_pos'] */
  const int64_t* terminate;    /* [N] terminate_buf */
  const int32_t* num_steps;    /* [N] get_motion_num_steps() (motion_lib_base.py:428-432) */
  int32_t num_envs, bound, max_steps_all, reserved;
  int32_t* ctrl; int32_t* terminate_state; float* hist; double* sums; int32_t* counts; float* mpjpe_out;
} pulse_eval_args_t;
int pulse_eval_step(const pulse_eval_args_t* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MotionLib loader on the device (parity green against the reference's tables, not yet timed).  Per clip: optional heading rotation, local rotations, forward kinematics, gaussian-filtered linear /
 * angular velocities and dof velocities (motion_lib_smpl.py:101-174, poselib skeleton3d.py:389-462, :1100-1118,
 * motion_lib_base.py:47-70) from the on-disk clip arrays concatenated over clips; fills the six fp32 tables
 * pulse_motionlib_create packs.  All pointers are device pointers.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const double* pose_quat_global;  /* [F, 24, 4] xyzw, as stored by convert_amass_isaac.py */
  const double* root_trans;        /* [F, 3] root_trans_offset */
  const int32_t* frame_clip;       /* [F] clip index of every frame */
  const int64_t* clip_start;       /* [M + 1] first frame of every clip, clip_start[M] = F */
  const float* fps;                /* [M] */
  const double* headings;          /* [M] heading angle drawn per clip (motion_lib_smpl.py:134-135), or NULL (im_eval / test) */
  const int32_t* parents;          /* [24] skeleton parent indices (-1 = root) */
  const float* local_translation;  /* [24, 3] skeleton offsets */
  int64_t total_frames, num_clips;
  float* gts; float* grs; float* lrs; float* gvs; float* gavs; float* dvs;   /* outputs, shapes as in pulse_motionlib_desc_t */
  float* tmp_vel; float* tmp_ang;  /* workspaces [F, 24, 3] */
} pulse_loader_args_t;
int pulse_motionlib_load_clips(const pulse_loader_args_t* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PULSE_B200_H_ */
