// Rollout glue of AMPAgent.play_steps (phc/learning/amp_agent.py:341-439) as three row-wise kernels that write STRAIGHT into the
// experience-buffer slices (pointer + stride), so the per-step ATen copies / elementwise chains of round 1 disappear:
//   policy_post_kernel  ModelA2CContinuousLogStd sampling [rl_games]: a = mu + sigma * eps with eps drawn in-kernel (Philox4x32-10 or
//                       injected), neglogp, value de-normalisation (RunningMeanStd.forward(unnorm=True), running_mean_std.py:84-87),
//                       PD targets (Humanoid._action_to_pd_targets, humanoid.py:1392-1394) -- get_action_values + the experience
//                       buffer updates of amp_agent.py:361-378 + pre_physics_step's target computation in ONE launch;
//   value_post_kernel   next_values = unnormalise(critic(next obs)) * (1 - terminated)  (amp_agent.py:396-398);
//   amp_row_kernel      this step's AMP observation row [cur | previous row's first (steps-1)*196 floats] written directly into the
//                       experience slice (humanoid_amp.py:622-667 + amp_agent.py:385): 7 056 B read + 7 840 B written per env instead
//                       of the in-place shift (7 056 + 7 840) followed by a 7 840 + 7 840 B copy.  Envs reset since the last step take
//                       their previous row from the rows `pulse_reset_ref_state` back-filled (`fresh` flags, cleared here).
#include "philox.cuh"
#include "pulse_common.cuh"
#include "quat_math.cuh"

namespace pulse {
namespace {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// RunningMeanStd.forward(unnorm=True): clamp(y, -5, 5) * sqrt(var.float() + eps) + mean.float()
__device__ __forceinline__ float value_unnorm(float y, const double* mean, const double* var, float eps) {
  if (mean == nullptr) return y;
  const float sd = sqrtf(__fadd_rn(static_cast<float>(var[0]), eps));
  return __fadd_rn(__fmul_rn(fminf(fmaxf(y, -5.0f), 5.0f), sd), static_cast<float>(mean[0]));
}

__global__ void __launch_bounds__(128) policy_post_kernel(const pulse_policy_post_args_t a, long long rows) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int A = a.num_actions;
  const unsigned long long off = a.rng_offset != nullptr ? *a.rng_offset + a.rng_step : a.rng_step;
  float acc = 0.0f, ls = 0.0f;
  // lanes take PAIRS of actions (2*lane + 64*i): one Philox call yields four words = two Box-Muller pairs
  for (int i = 0; 2 * lane + 64 * i < A; ++i) {
    const int k0 = 2 * lane + 64 * i;
    float e0, e1;
    if (a.eps != nullptr) {
      e0 = a.eps[row * a.ld_eps + k0];
      e1 = k0 + 1 < A ? a.eps[row * a.ld_eps + k0 + 1] : 0.0f;
    } else {
      const Philox4 r = philox4x32_10(a.seed, static_cast<unsigned long long>(row) * 64ull + static_cast<unsigned long long>(lane + 32 * i), off);
      box_muller(r.x, r.y, e0, e1);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = k0 + h;
      if (k >= A) break;
      const float l = a.logstd[k];
      const float sg = expf(l);
      const float m = a.mu[row * a.ld_mu + k];
      const float act = m + sg * (h == 0 ? e0 : e1);
      a.actions[row * a.ld_actions + k] = act;
      if (a.mus_out != nullptr) a.mus_out[row * a.ld_mus + k] = m;
      if (a.pd_targets != nullptr) {
        const float tgt = __fadd_rn(a.pd_offset[k], __fmul_rn(a.pd_scale[k], act));     // humanoid.py:1392-1394
        a.pd_targets[row * a.ld_pd + k] = tgt;
      }
      const float z = (act - m) / sg;
      acc += z * z;
      ls += l;
    }
  }
  acc = wsum(acc);
  ls = wsum(ls);
  if (lane == 0) {
    a.neglogp[row * a.ld_neglogp] = 0.5f * acc + 0.5f * 1.8378770664093453f * A + ls;   // log(2*pi)
    if (a.values_out != nullptr) a.values_out[row * a.ld_values] = value_unnorm(a.value[row * a.ld_value], a.value_mean, a.value_var, a.value_eps);
  }
}

__global__ void __launch_bounds__(256) value_post_kernel(const float* __restrict__ value, long long ld_value, const double* mean, const double* var,
                                                         float eps, const long long* __restrict__ terminate, float* __restrict__ out,
                                                         long long ld_out, long long rows) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float v = value_unnorm(value[r * ld_value], mean, var, eps);
  if (terminate != nullptr) v = __fmul_rn(v, __fsub_rn(1.0f, static_cast<float>(terminate[r])));   // next_vals *= (1.0 - terminated)
  out[r * ld_out] = v;
}

__constant__ int a_kept_joint[19] = {0, 1, 2, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 19, 20, 21};
__constant__ int a_key_body[4] = {7, 3, 22, 17};
constexpr int kAmp = PULSE_AMP_OBS;

__global__ void __launch_bounds__(128) amp_row_kernel(const pulse_amp_row_args_t a, long long n) {
  const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= n) return;
  float* out = a.out + e * a.ld_out;
  const int hist = (a.num_steps - 1) * kAmp;
  // ---- history part: out[196 + k] <- prev[k], k < hist (prev and out never alias: different experience slices) ------------------
  const bool fresh = a.fresh != nullptr && a.fresh[e] != 0;
  const float* prev = fresh ? a.fresh_rows + e * (long long)a.num_steps * kAmp : a.prev + e * a.ld_prev;
  {
    // (steps-1)*49 16-byte units per env (441 for 10 steps): all loads of the warp are issued before the first store (16 in flight per lane)
    const float4* src = reinterpret_cast<const float4*>(prev);
    float4* dst = reinterpret_cast<float4*>(out + kAmp);
    const int nvec = hist / 4;
    float4 regs[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) regs[i] = __ldcs(src + c);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) __stcs(dst + c, regs[i]);
    }
    for (int c = lane + 512; c < nvec; c += 32) dst[c] = src[c];   // more than 11 history steps
  }
  if (fresh && lane == 0) a.fresh[e] = 0;
  // ---- current observation (same arithmetic as amp_obs_kernel) ---------------------------------------------------------------
  const float* bs = a.body_state + e * a.body_env_stride;
  const Vec3 p0 = {bs[0], bs[1], bs[2]};
  const Quat q0 = {bs[3], bs[4], bs[5], bs[6]};
  float hs, hc;
  heading_half(q0, hs, hc);
  const Quat h_inv = {0.0f, 0.0f, -hs, hc};
  const Yaw yr = make_yaw(h_inv);
  float* o = out;
  if (lane == 0) {
    o[0] = p0.z;
    float six[6];
    qsix(qmul(h_inv, q0), six);
#pragma unroll
    for (int i = 0; i < 6; ++i) o[1 + i] = six[i];
    const Vec3 lv = yaw_rot(yr, {bs[7], bs[8], bs[9]});
    const Vec3 lw = yaw_rot(yr, {bs[10], bs[11], bs[12]});
    o[7] = lv.x; o[8] = lv.y; o[9] = lv.z;
    o[10] = lw.x; o[11] = lw.y; o[12] = lw.z;
  }
  const float* dp = a.dof_pos + e * a.dof_env_stride;
  const float* dv = a.dof_vel + e * a.dof_env_stride;
  if (lane < 19) {
    const int jt = a_kept_joint[lane];
    const Vec3 em = {dp[(3 * jt + 0) * a.dof_elem_stride], dp[(3 * jt + 1) * a.dof_elem_stride], dp[(3 * jt + 2) * a.dof_elem_stride]};
    float six[6];
    qsix(exp_map_quat(em), six);
#pragma unroll
    for (int i = 0; i < 6; ++i) o[13 + 6 * lane + i] = six[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) o[127 + 3 * lane + i] = dv[(3 * jt + i) * a.dof_elem_stride];
  } else if (lane < 23) {
    const int kb = a_key_body[lane - 19];
    const float* bk = bs + kb * PULSE_BODY_STATE_W;
    const Vec3 lp = yaw_rot(yr, Vec3{bk[0], bk[1], bk[2]} - p0);
    o[184 + 3 * (lane - 19) + 0] = lp.x;
    o[184 + 3 * (lane - 19) + 1] = lp.y;
    o[184 + 3 * (lane - 19) + 2] = lp.z;
  }
}

__global__ void bump_counter_kernel(unsigned long long* c, unsigned long long by) { *c += by; }

}  // namespace
}  // namespace pulse

extern "C" int pulse_policy_post(const pulse_policy_post_args_t* args, int64_t rows, void* stream) {
  using namespace pulse;
  PULSE_REQUIRE(args != nullptr, "pulse_policy_post: null args");
  PULSE_REQUIRE(rows >= 0, "pulse_policy_post: negative rows");
  if (rows == 0) return PULSE_OK;
  const pulse_policy_post_args_t& a = *args;
  PULSE_REQUIRE(a.mu && a.logstd && a.actions && a.neglogp, "pulse_policy_post: null mu / logstd / actions / neglogp");
  PULSE_REQUIRE(a.num_actions >= 1 && a.num_actions <= 128, "pulse_policy_post: num_actions %d outside [1,128]", a.num_actions);
  PULSE_REQUIRE(a.ld_mu >= a.num_actions && a.ld_actions >= a.num_actions && a.ld_neglogp >= 1, "pulse_policy_post: leading dimensions too small");
  PULSE_REQUIRE(a.eps == nullptr || a.ld_eps >= a.num_actions, "pulse_policy_post: ld_eps too small");
  PULSE_REQUIRE(a.mus_out == nullptr || a.ld_mus >= a.num_actions, "pulse_policy_post: ld_mus too small");
  PULSE_REQUIRE(a.values_out == nullptr || (a.value != nullptr && a.ld_value >= 1 && a.ld_values >= 1), "pulse_policy_post: values_out needs value");
  PULSE_REQUIRE((a.value_mean == nullptr) == (a.value_var == nullptr), "pulse_policy_post: value_mean and value_var go together");
  PULSE_REQUIRE(a.pd_targets == nullptr || (a.pd_offset && a.pd_scale && a.ld_pd >= a.num_actions), "pulse_policy_post: pd_targets needs offset / scale");
  policy_post_kernel<<<static_cast<unsigned>((rows * 32 + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(a, (long long)rows);
  PULSE_LAUNCH_OK("policy_post_kernel");
  return PULSE_OK;
}

extern "C" int pulse_value_post(const float* value, int64_t ld_value, const double* mean, const double* var, float eps, const int64_t* terminate,
                                float* out, int64_t ld_out, int64_t rows, void* stream) {
  using namespace pulse;
  PULSE_REQUIRE(value != nullptr && out != nullptr, "pulse_value_post: null value / out");
  PULSE_REQUIRE(rows >= 0 && ld_value >= 1 && ld_out >= 1, "pulse_value_post: bad sizes");
  PULSE_REQUIRE((mean == nullptr) == (var == nullptr), "pulse_value_post: mean and var go together");
  if (rows == 0) return PULSE_OK;
  value_post_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      value, ld_value, mean, var, eps, reinterpret_cast<const long long*>(terminate), out, ld_out, (long long)rows);
  PULSE_LAUNCH_OK("value_post_kernel");
  return PULSE_OK;
}

extern "C" int pulse_amp_obs_row(const pulse_amp_row_args_t* args, int64_t num_envs, void* stream) {
  using namespace pulse;
  PULSE_REQUIRE(args != nullptr, "pulse_amp_obs_row: null args");
  PULSE_REQUIRE(num_envs >= 0, "pulse_amp_obs_row: negative num_envs");
  if (num_envs == 0) return PULSE_OK;
  const pulse_amp_row_args_t& a = *args;
  PULSE_REQUIRE(a.body_state && a.dof_pos && a.dof_vel && a.prev && a.out, "pulse_amp_obs_row: null buffer");
  PULSE_REQUIRE(a.num_steps >= 1 && a.num_steps <= 16, "pulse_amp_obs_row: num_steps %d outside [1,16]", a.num_steps);
  PULSE_REQUIRE(a.ld_prev >= (a.num_steps - 1) * PULSE_AMP_OBS && a.ld_out >= a.num_steps * PULSE_AMP_OBS, "pulse_amp_obs_row: row strides too small");
  PULSE_REQUIRE(aligned16(a.prev) && aligned16(a.out) && (a.ld_prev % 4) == 0 && (a.ld_out % 4) == 0, "pulse_amp_obs_row: rows must be 16-byte aligned");
  PULSE_REQUIRE(a.prev != a.out, "pulse_amp_obs_row: prev and out must be different experience slices (use pulse_amp_obs for the in-place shift)");
  PULSE_REQUIRE((a.fresh == nullptr) == (a.fresh_rows == nullptr), "pulse_amp_obs_row: fresh flags and fresh_rows go together");
  PULSE_REQUIRE(a.fresh_rows == nullptr || aligned16(a.fresh_rows), "pulse_amp_obs_row: fresh_rows not 16-byte aligned");
  PULSE_REQUIRE(a.body_env_stride >= PULSE_NUM_BODIES * PULSE_BODY_STATE_W, "pulse_amp_obs_row: body_env_stride too small");
  const long long threads = num_envs * 32;
  amp_row_kernel<<<static_cast<unsigned>((threads + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(a, (long long)num_envs);
  PULSE_LAUNCH_OK("amp_row_kernel");
  return PULSE_OK;
}

extern "C" int pulse_bump_counter(uint64_t* counter, uint64_t by, void* stream) {
  using namespace pulse;
  PULSE_REQUIRE(counter != nullptr, "pulse_bump_counter: null counter");
  bump_counter_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long*>(counter), by);
  PULSE_LAUNCH_OK("bump_counter_kernel");
  return PULSE_OK;
}

// ---- timing events that survive CUDA-graph capture ---------------------------------------------------------------------------
// torch.cuda.Event.record() inside a capture becomes an internal dependency node that cannot be queried; bench.py has to time the
// fused step kernel LIVE inside the timed region even when the whole rollout is one graph, so these wrap cudaEventRecordWithFlags
// (cudaEventRecordExternal): captured as an event-record NODE, the event is re-recorded by every replay and elapsed times can be read.
extern "C" int pulse_event_create(void** event) {
  using namespace pulse;
  PULSE_REQUIRE(event != nullptr, "pulse_event_create: null out pointer");
  cudaEvent_t e;
  PULSE_CUDA_OK(cudaEventCreate(&e));
  *event = e;
  return PULSE_OK;
}
extern "C" int pulse_event_destroy(void* event) {
  using namespace pulse;
  if (event != nullptr) PULSE_CUDA_OK(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
  return PULSE_OK;
}
extern "C" int pulse_event_record(void* event, void* stream) {
  using namespace pulse;
  PULSE_REQUIRE(event != nullptr, "pulse_event_record: null event");
  // the external flag is only legal while the stream is being captured; outside a capture this is an ordinary record
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  PULSE_CUDA_OK(cudaStreamIsCapturing(static_cast<cudaStream_t>(stream), &st));
  PULSE_CUDA_OK(cudaEventRecordWithFlags(static_cast<cudaEvent_t>(event), static_cast<cudaStream_t>(stream),
                                         st == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault));
  return PULSE_OK;
}
extern "C" int pulse_event_elapsed_ms(void* start, void* stop, float* ms) {
  using namespace pulse;
  PULSE_REQUIRE(start != nullptr && stop != nullptr && ms != nullptr, "pulse_event_elapsed_ms: null argument");
  PULSE_CUDA_OK(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
  return PULSE_OK;
}
