// Fused HumanoidIm post-physics step kernel: reward(t) -> reset(t) -> observation(t+dt).
//
// One warp owns one env; lane j (< 24) owns rigid body j.  Per env the warp
//   1. reads the per-env task scalars and the per-motion constants, computes the two motion times
//      and the frame indices with the reference's exact fp32 operation order;
//   2. stages up to four packed 1248-byte frame records (deduplicated when the reward and the
//      observation query share a frame) plus the env's 24x13 rigid-body state into shared memory
//      with cp.async (16-byte, L1-bypassing), and the dof force / velocity rows for the power term;
//   3. blends the reference pose per lane (lerp / slerp), reduces the four reward errors and the
//      termination test across the warp with shuffles / ballot;
//   4. builds the 934-float observation row in shared memory and streams it out with 16-byte
//      stores (row start may be only 8-byte aligned: 934*4 = 3736; the staging buffer is phase
//      shifted so that global 16-byte boundaries coincide with shared ones).
//
// HBM-bound by design: algorithmic traffic is 9 396 B per env-step and nothing is
// re-read.  References: humanoid_im.py:853-919, :1119-1192, :677-851, :1328-1378, :1543-1628;
// humanoid.py:1675-1731; motion_lib_base.py:434-517, :546-556.
#include "pulse_common.cuh"
#include "quat_math.cuh"

namespace pulse {
namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kFrame = PULSE_FRAME_REC;  // 312 floats
constexpr int kNB = PULSE_NUM_BODIES;
constexpr int kObs = PULSE_IM_OBS;  // 934
constexpr int kObsPad = 944;        // 934 + 3 phase slack, rounded up to a multiple of 4

struct __align__(16) WarpStage {
  float frames[4][kFrame];  // 4 x 1248 B
  float body[kFrame + 8];   // rigid-body state rows of this env (+ phase slack)
  float obs[kObsPad];       // observation row staging
};
static_assert(sizeof(WarpStage) % 16 == 0, "stage must keep 16-byte alignment");

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

__device__ __forceinline__ Vec3 ld3(const float* p) { return {p[0], p[1], p[2]}; }
__device__ __forceinline__ Quat ld4(const float* p) {
  float4 v = *reinterpret_cast<const float4*>(p);
  return {v.x, v.y, v.z, v.w};
}

// Reference pose of body j blended between two staged frame records.
struct RefPose {
  Vec3 p, v, w;
  Quat q;
};
__device__ __forceinline__ RefPose blend_pose(const float* f0, const float* f1, int j, float b, Vec3 goff) {
  RefPose r;
  // position feeds the reset mask: reproduce ((1-b)*p0 + b*p1) + off without contraction
  r.p.x = __fadd_rn(lerp_rn(f0[3 * j + 0], f1[3 * j + 0], b), goff.x);
  r.p.y = __fadd_rn(lerp_rn(f0[3 * j + 1], f1[3 * j + 1], b), goff.y);
  r.p.z = __fadd_rn(lerp_rn(f0[3 * j + 2], f1[3 * j + 2], b), goff.z);
  r.q = slerp(ld4(f0 + 72 + 4 * j), ld4(f1 + 72 + 4 * j), b);
  float a = 1.0f - b;
  const float* v0 = f0 + 168 + 3 * j;
  const float* v1 = f1 + 168 + 3 * j;
  r.v = {a * v0[0] + b * v1[0], a * v0[1] + b * v1[1], a * v0[2] + b * v1[2]};
  const float* w0 = f0 + 240 + 3 * j;
  const float* w1 = f1 + 240 + 3 * j;
  r.w = {a * w0[0] + b * w1[0], a * w0[1] + b * w1[1], a * w0[2] + b * w1[2]};
  return r;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32) im_step_kernel(const pulse_motionlib_desc_t lib,
                                                                    const pulse_im_step_args_t a, long long num_envs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long widx = (long long)blockIdx.x * kWarpsPerCta + warp;
  if (widx >= num_envs) return;
  const long long e = a.env_ids != nullptr ? a.env_ids[widx] : widx;
  WarpStage& st = reinterpret_cast<WarpStage*>(smem_raw)[warp];

  const bool do_rew = a.flags & PULSE_STEP_REWARD;
  const bool do_reset = a.flags & PULSE_STEP_RESET;
  const bool do_obs = a.flags & PULSE_STEP_OBS;

  // ---- 1. per-env scalars, motion constants, frame indices (uniform across the warp) ----------
  const long long prog = a.progress_buf[e];
  const long long mid = a.motion_ids[e];
  const float t_start = a.motion_start_times[e];
  const float t_off = a.motion_start_offset[e];
  const Vec3 goff = {a.global_offset[3 * e + 0], a.global_offset[3 * e + 1], a.global_offset[3 * e + 2]};
  const float mlen = lib.lengths[mid];
  const float mdt = lib.dt[mid];
  const long long nf = lib.num_frames[mid];
  const long long row0 = lib.length_starts[mid];

  const float t_rew = motion_time_rn(prog, a.dt, t_start, t_off);
  const float t_obs = motion_time_rn(prog + 1, a.dt, t_start, t_off);
  long long i0r, i1r, i0o, i1o;
  float b_rew, b_obs;
  frame_blend_rn(t_rew, mlen, nf, mdt, i0r, i1r, b_rew);
  frame_blend_rn(t_obs, mlen, nf, mdt, i0o, i1o, b_obs);
  const bool need_t = do_rew || do_reset;
  long long rows[4] = {row0 + i0r, row0 + i1r, row0 + i0o, row0 + i1o};
  int slot[4];
  bool fetch[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    bool wanted = (k < 2) ? need_t : do_obs;
    slot[k] = k;
    fetch[k] = wanted;
#pragma unroll
    for (int m = 0; m < k; ++m) {
      if (fetch[k] && fetch[m] && slot[m] == m && rows[m] == rows[k]) {
        slot[k] = m;
        fetch[k] = false;
      }
    }
  }

  // ---- 2. stage frame records + body state -----------------------------------------------------
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (fetch[k]) {
      const float* src = lib.frame_rec + rows[k] * kFrame;
      for (int c = lane; c < kFrame / 4; c += 32) cp_async16(&st.frames[k][4 * c], src + 4 * c);
    }
  }
  const float* bsrc = a.body_state + e * a.body_env_stride;
  const int bphase = static_cast<int>((reinterpret_cast<uintptr_t>(bsrc) >> 2) & 3);
  float* body = st.body + bphase;  // 16-byte boundaries of the source line up with shared memory
  {
    const int head = (4 - bphase) & 3;  // scalars before the first aligned chunk
    const int nvec = (kFrame - head) / 4;
    const int tail0 = head + 4 * nvec;
    if (lane < head) cp_async4(body + lane, bsrc + lane);
    for (int c = lane; c < nvec; c += 32) cp_async16(body + head + 4 * c, bsrc + head + 4 * c);
    if (lane < kFrame - tail0) cp_async4(body + tail0 + lane, bsrc + tail0 + lane);
  }
  // power term operands straight to registers while the copies are in flight
  float pw = 0.0f;
  const bool do_power = do_rew && a.dof_force != nullptr;
  if (do_power) {
    const float* fr = a.dof_force + e * a.dof_force_stride;
    const float* dv = a.dof_vel + e * a.dof_env_stride;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int d = lane + 32 * k;
      if (d < PULSE_NUM_DOF) pw += fabsf(fr[d] * dv[d * a.dof_elem_stride]);
    }
  }
  float term_j = 0.0f;
  if (do_reset && lane < kNB) term_j = a.termination_distances[lane];
  const int cyc = (do_reset && a.cycle_counter != nullptr) ? a.cycle_counter[e] : 0;
  cp_async_wait_all();
  __syncwarp();

  // ---- 3. per-body state, heading ----------------------------------------------------------------
  const int j = lane < kNB ? lane : kNB - 1;
  const bool active = lane < kNB;
  const float* bj = body + j * PULSE_BODY_STATE_W;
  const Vec3 p = {bj[0], bj[1], bj[2]};
  const Quat q = {bj[3], bj[4], bj[5], bj[6]};
  const Vec3 v = {bj[7], bj[8], bj[9]};
  const Vec3 w = {bj[10], bj[11], bj[12]};
  const Vec3 p_root = {body[0], body[1], body[2]};
  const Quat q_root = {body[3], body[4], body[5], body[6]};

  // ---- reward + reset at t -----------------------------------------------------------------------
  if (need_t) {
    const RefPose r = blend_pose(st.frames[slot[0]], st.frames[slot[1]], j, b_rew, goff);
    const bool pass_time = a.cycle_motion ? (prog >= a.max_episode_length - 1) : (t_rew >= mlen);
    if (do_rew) {
      float e_pos = active ? sq3(r.p - p) : 0.0f;
      float e_vel = active ? sq3(r.v - v) : 0.0f;
      float e_ang = active ? sq3(r.w - w) : 0.0f;
      float th = quat_angle(qmul(r.q, qconj(q)));
      float e_rot = active ? th * th : 0.0f;
      e_pos = warp_sum(e_pos) * (1.0f / (3.0f * kNB));
      e_vel = warp_sum(e_vel) * (1.0f / (3.0f * kNB));
      e_ang = warp_sum(e_ang) * (1.0f / (3.0f * kNB));
      e_rot = warp_sum(e_rot) * (1.0f / kNB);
      const float r_pos = expf(-a.k_pos * e_pos);
      const float r_rot = expf(-a.k_rot * e_rot);
      const float r_vel = expf(-a.k_vel * e_vel);
      const float r_ang = expf(-a.k_ang_vel * e_ang);
      float rew = a.w_pos * r_pos + a.w_rot * r_rot + a.w_vel * r_vel + a.w_ang_vel * r_ang;
      float p_rew = 0.0f;
      if (do_power) {
        pw = warp_sum(pw);
        p_rew = (prog <= 3) ? 0.0f : -a.power_coefficient * pw;
        rew += p_rew;
      }
      if (lane == 0) a.rew_buf[e] = rew;
      if (a.reward_raw != nullptr) {
        float val = lane == 0 ? r_pos : lane == 1 ? r_rot : lane == 2 ? r_vel : lane == 3 ? r_ang : p_rew;
        if (lane < 4 || (lane == 4 && do_power)) a.reward_raw[e * a.raw_stride + lane] = val;
      }
    }
    if (do_reset) {
      const bool in_mask = active && ((a.reset_body_mask >> j) & 1u);
      const float dist = norm3_rn(__fsub_rn(p.x, r.p.x), __fsub_rn(p.y, r.p.y), __fsub_rn(p.z, r.p.z));
      bool fallen;
      if (a.use_mean_reset) {
        // mean over the reset bodies vs the first reset body's distance (humanoid_im.py:1606)
        const unsigned m = a.reset_body_mask & 0xffffffu;
        const float mean = warp_sum(in_mask ? dist : 0.0f) / static_cast<float>(__popc(m));
        const float d0 = __shfl_sync(kFull, term_j, __ffs(m) - 1);
        fallen = mean > d0;
      } else {
        fallen = __ballot_sync(kFull, in_mask && (dist > term_j)) != 0u;
      }
      fallen = fallen && (prog > 1) && a.enable_early_termination;
      long long terminated = fallen ? 1 : 0;
      long long reset = pass_time ? 1 : terminated;
      if (!pass_time && cyc > 0) {  // recovering envs: humanoid_im.py:1188-1190
        reset = 0;
        terminated = 0;
      }
      if (lane == 0) {
        a.reset_buf[e] = reset;
        a.terminate_buf[e] = terminated;
      }
    }
    if (lane == 0 && a.pass_time != nullptr) a.pass_time[e] = (t_rew >= mlen) ? 1 : 0;
  }

  // ---- observation at t + dt ---------------------------------------------------------------------
  if (do_obs) {
    const RefPose r = blend_pose(st.frames[slot[2]], st.frames[slot[3]], j, b_obs, goff);
    const float hd = heading_angle(q_root);
    const Quat h_inv = yaw_quat(-hd);
    const Quat h_fwd = yaw_quat(hd);
    const Yaw yr = make_yaw(h_inv);

    float* orow = a.obs_buf + e * a.obs_stride;
    const int ophase = static_cast<int>((reinterpret_cast<uintptr_t>(orow) >> 2) & 3);
    float* o = st.obs + ophase;
    if (active) {
      // self observation (humanoid.py:1675-1731)
      if (j == 0) {
        o[0] = p_root.z;
      } else {
        Vec3 lp = yaw_rot(yr, p - p_root);
        o[1 + 3 * (j - 1) + 0] = lp.x;
        o[1 + 3 * (j - 1) + 1] = lp.y;
        o[1 + 3 * (j - 1) + 2] = lp.z;
      }
      qsix(qmul(h_inv, q), o + 70 + 6 * j);
      Vec3 lv = yaw_rot(yr, v);
      o[214 + 3 * j + 0] = lv.x;
      o[214 + 3 * j + 1] = lv.y;
      o[214 + 3 * j + 2] = lv.z;
      Vec3 lw = yaw_rot(yr, w);
      o[286 + 3 * j + 0] = lw.x;
      o[286 + 3 * j + 1] = lw.y;
      o[286 + 3 * j + 2] = lw.z;
      // task observation v6 (humanoid_im.py:1328-1378), block-major
      float* t = o + PULSE_SELF_OBS;
      Vec3 dp = yaw_rot(yr, r.p - p);
      t[3 * j + 0] = dp.x;
      t[3 * j + 1] = dp.y;
      t[3 * j + 2] = dp.z;
      qsix(qmul(qmul(h_inv, qmul(r.q, qconj(q))), h_fwd), t + 72 + 6 * j);
      Vec3 dv = yaw_rot(yr, r.v - v);
      t[216 + 3 * j + 0] = dv.x;
      t[216 + 3 * j + 1] = dv.y;
      t[216 + 3 * j + 2] = dv.z;
      Vec3 dw = yaw_rot(yr, r.w - w);
      t[288 + 3 * j + 0] = dw.x;
      t[288 + 3 * j + 1] = dw.y;
      t[288 + 3 * j + 2] = dw.z;
      Vec3 rp = yaw_rot(yr, r.p - p_root);
      t[360 + 3 * j + 0] = rp.x;
      t[360 + 3 * j + 1] = rp.y;
      t[360 + 3 * j + 2] = rp.z;
      qsix(qmul(h_inv, r.q), t + 432 + 6 * j);
      // reference-pose side buffers (humanoid_im.py:835-848)
      if (a.ref_body_pos != nullptr) {
        float* d = a.ref_body_pos + e * (kNB * 3) + 3 * j;
        d[0] = r.p.x; d[1] = r.p.y; d[2] = r.p.z;
      }
      if (a.ref_body_vel != nullptr) {
        float* d = a.ref_body_vel + e * (kNB * 3) + 3 * j;
        d[0] = r.v.x; d[1] = r.v.y; d[2] = r.v.z;
      }
      if (a.ref_body_rot != nullptr) {
        float* d = a.ref_body_rot + e * (kNB * 4) + 4 * j;
        d[0] = r.q.x; d[1] = r.q.y; d[2] = r.q.z; d[3] = r.q.w;
      }
    }
    if (a.ref_dof_pos != nullptr && lane >= 1 && lane < kNB) {
      // dof_pos = exp_map(slerp(lrs[f0, j], lrs[f1, j], blend)), joints 1..23 (motion_lib_base.py:489-490)
      const float* x0 = lib.aux_rec + rows[2] * PULSE_AUX_REC + 4 * lane;
      const float* x1 = lib.aux_rec + rows[3] * PULSE_AUX_REC + 4 * lane;
      Vec3 em = quat_exp_map(slerp(ld4(x0), ld4(x1), b_obs));
      float* d = a.ref_dof_pos + e * PULSE_NUM_DOF + 3 * (lane - 1);
      d[0] = em.x; d[1] = em.y; d[2] = em.z;
    }
    __syncwarp();
    // stream the row out: 16-byte stores on aligned slots, scalars on the ragged ends
    const int nslot = (ophase + kObs + 3) / 4;
    for (int s = lane; s < nslot; s += 32) {
      const int k0 = 4 * s - ophase;  // obs index of the slot's first float
      if (k0 >= 0 && k0 + 3 < kObs) {
        *reinterpret_cast<float4*>(orow + k0) = *reinterpret_cast<const float4*>(st.obs + 4 * s);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int k = k0 + i;
          if (k >= 0 && k < kObs) orow[k] = st.obs[4 * s + i];
        }
      }
    }
    if (a.self_obs_buf != nullptr) {
      float* srow = a.self_obs_buf + e * PULSE_SELF_OBS;
      for (int k = lane; k < PULSE_SELF_OBS; k += 32) srow[k] = o[k];
    }
  }
}

}  // namespace
}  // namespace pulse

extern "C" int pulse_im_step(const pulse_motionlib_t* lib, const pulse_im_step_args_t* args, int64_t num_envs,
                             void* stream) {
  using namespace pulse;
  PULSE_REQUIRE(lib != nullptr && args != nullptr, "pulse_im_step: null lib/args");
  PULSE_REQUIRE(num_envs >= 0, "pulse_im_step: negative num_envs");
  if (num_envs == 0) return PULSE_OK;
  const pulse_im_step_args_t& a = *args;
  PULSE_REQUIRE((a.flags & PULSE_STEP_ALL) != 0 && (a.flags & ~PULSE_STEP_ALL) == 0, "pulse_im_step: bad flags 0x%x", a.flags);
  PULSE_REQUIRE(a.body_state && a.progress_buf && a.motion_ids && a.motion_start_times && a.motion_start_offset &&
                    a.global_offset, "pulse_im_step: null state/task buffer");
  PULSE_REQUIRE(a.body_env_stride >= PULSE_NUM_BODIES * PULSE_BODY_STATE_W, "pulse_im_step: body_env_stride %lld < 312",
                (long long)a.body_env_stride);
  PULSE_REQUIRE((reinterpret_cast<uintptr_t>(a.body_state) & 3u) == 0, "pulse_im_step: body_state not 4-byte aligned");
  if (a.flags & PULSE_STEP_REWARD) {
    PULSE_REQUIRE(a.rew_buf != nullptr, "pulse_im_step: rew_buf is null");
    if (a.dof_force) {
      PULSE_REQUIRE(a.dof_vel != nullptr, "pulse_im_step: dof_force given without dof_vel");
      PULSE_REQUIRE(!a.reward_raw || a.raw_stride >= 5, "pulse_im_step: raw_stride must be >= 5 with the power term");
    } else {
      PULSE_REQUIRE(!a.reward_raw || a.raw_stride >= 4, "pulse_im_step: raw_stride must be >= 4");
    }
  }
  if (a.flags & PULSE_STEP_RESET) {
    PULSE_REQUIRE(a.reset_buf && a.terminate_buf && a.termination_distances, "pulse_im_step: null reset buffer");
    PULSE_REQUIRE((a.reset_body_mask & 0xffffffu) != 0, "pulse_im_step: empty reset_body_mask");
  }
  if (a.flags & PULSE_STEP_OBS) {
    PULSE_REQUIRE(a.obs_buf != nullptr && a.obs_stride >= PULSE_IM_OBS, "pulse_im_step: obs_buf null or obs_stride < 934");
    PULSE_REQUIRE((reinterpret_cast<uintptr_t>(a.obs_buf) & 3u) == 0, "pulse_im_step: obs_buf misaligned");
    PULSE_REQUIRE(!a.ref_dof_pos || lib->d.aux_rec, "pulse_im_step: ref_dof_pos needs the aux records");
  }
  static bool attr_set = false;
  const size_t smem = sizeof(WarpStage) * kWarpsPerCta;
  if (!attr_set) {
    PULSE_CUDA_OK(cudaFuncSetAttribute(im_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const unsigned grid = static_cast<unsigned>((num_envs + kWarpsPerCta - 1) / kWarpsPerCta);
  im_step_kernel<<<grid, kWarpsPerCta * 32, smem, static_cast<cudaStream_t>(stream)>>>(lib->d, a, (long long)num_envs);
  PULSE_LAUNCH_OK("im_step_kernel");
  return PULSE_OK;
}
