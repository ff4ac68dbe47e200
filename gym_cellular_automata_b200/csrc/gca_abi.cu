// gca_abi.cu -- the extern "C" surface of libgca.so (see include/gca.h).  Host code only:
// argument checks, constant derivation (A0) and kernel dispatch by grid shape.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "gca_common.cuh"

namespace gca {
cudaError_t launch_pack(const gca_params&, const gca_state&, const float*, const float*, const int32_t*,
                        const int32_t*, const int32_t*, uint8_t*, int32_t*, cudaStream_t);
cudaError_t launch_unpack(const gca_params&, const gca_state&, float*, float*, int32_t*, cudaStream_t);
cudaError_t launch_move_modify(const gca_params&, const gca_state&, const int32_t*, cudaStream_t);
cudaError_t launch_reward_done(const gca_params&, const gca_state&, float*, uint8_t*, int32_t*, cudaStream_t);
cudaError_t launch_conditional_reset(const gca_params&, const gca_state&, const gca_state&, const float*, float*,
                                     uint8_t*, cudaStream_t);
cudaError_t launch_render(const gca_params&, int, const uint8_t*, const uint64_t*, const int32_t*, const uint8_t*,
                          const int32_t*, int, const uint8_t*, int, int, uint32_t*, void*, cudaStream_t);
cudaError_t launch_threefry_bits(const uint32_t*, long long, int, uint32_t*, cudaStream_t);
cudaError_t launch_threefry_split_part(const uint32_t*, int, uint32_t*, cudaStream_t);
cudaError_t launch_balance_order(int, const uint32_t*, int32_t*, cudaStream_t);
cudaError_t launch_generate_hidden(int, int, int, unsigned long long, int, int32_t*, int32_t*, float*, double*, float*,
                                   float*, cudaStream_t);
cudaError_t launch_episode_stats(int, const gca_episode_stats&, const float*, const uint8_t*, const uint8_t*,
                                 const uint8_t*, const int32_t*, cudaStream_t);
cudaError_t launch_windy_step(int, int, int, unsigned long long*, unsigned long long*, int32_t*, double*, const int32_t*,
                              const double*, const double*, int, double, double, double, double*, uint8_t*, int32_t*,
                              int32_t*, cudaStream_t);
cudaError_t launch_windy_pack(int, int, int, const uint8_t*, unsigned long long*, unsigned long long*, cudaStream_t);
cudaError_t launch_windy_unpack(int, int, int, const unsigned long long*, const unsigned long long*, uint8_t*,
                                cudaStream_t);
cudaError_t launch_tiled_env_step(const gca_params&, const gca_state&, const int32_t*, const gca_step_out&,
                                  const gca_inject&, uint32_t, uint8_t*, uint32_t*, int32_t*, uint8_t*, int, const gca_state*,
                                  const float*, cudaStream_t);
cudaError_t launch_auto_reset(const gca_params&, const gca_state&, const gca_state&, const float*, float*,
                              const uint8_t*, cudaStream_t);
bool bb_supported(const gca_params&);
cudaError_t launch_bb_env_step(const gca_params&, const gca_state&, const int32_t*, const gca_step_out&, const gca_inject&,
                               uint32_t, cudaStream_t);
}  // namespace gca

static thread_local char g_err[256] = "";

static int fail(int code, const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
static int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return GCA_OK;
  std::snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return GCA_ERR_CUDA;
}
static bool is64(const gca_params* p) { return p->H == 64 && p->W == 64; }

extern "C" {

#ifndef GCA_BUILD_ID
#define GCA_BUILD_ID "unknown"
#endif
int gca_version(void) { return GCA_VERSION; }
const char* gca_build_id(void) { return GCA_BUILD_ID; }
const char* gca_last_error(void) { return g_err; }

int gca_params_init(gca_params* p, int32_t nrows, int32_t ncols, int32_t K, double speed_move, double speed_act,
                    double t_any, double t_move, double t_shoot, double p_tree, double p_wind_change,
                    int32_t rng_mode, const float* winds72) {
  if (!p) return fail(GCA_ERR_ARG, "gca_params_init: null params");
  if (nrows < 8 || ncols < 8) return fail(GCA_ERR_UNSUPPORTED, "gca_params_init: grid smaller than 8x8");
  if (K < 1 || K > GCA_MAX_K) return fail(GCA_ERR_ARG, "gca_params_init: K out of range [1, 8]");
  if (rng_mode != GCA_RNG_LEGACY && rng_mode != GCA_RNG_PARTITIONABLE)
    return fail(GCA_ERR_ARG, "gca_params_init: bad rng_mode");
  std::memset(p, 0, sizeof(*p));
  p->H = nrows;
  p->W = ncols;
  p->K = K;
  p->rng_mode = rng_mode;
  // ca_alexandridis_jax.py:57-65 (grid_size = nrows, advanced_bulldozer.py:276-282)
  const int size = nrows;
  const int spread = size + size / 2;
  const double age_min = spread * 1.5, age_max = spread * 1.75;
  const int R = (int)std::ceil(std::log2((double)size)) - 2;
  if (R < 1 || R > GCA_MAX_R) return fail(GCA_ERR_UNSUPPORTED, "gca_params_init: burn radius out of range [1, 10]");
  p->R = R;
  p->dous_border = (float)(0.0007 * age_max * 0.50);
  p->dous_inner = (float)(0.006 * age_max * 0.50);
  // build_burn_kernel (:108-151): 0.065 in total, ring i takes 60 % of what is left
  double remaining = 0.065;
  for (int i = 0; i < R; ++i) {
    const int cells = (2 * i + 3) * (2 * i + 3) - (2 * i + 1) * (2 * i + 1) + (i == 0 ? 1 : 0);
    double lw;
    if (i == R - 1) lw = remaining / cells;
    else { lw = remaining * 0.60 / cells; remaining = remaining * 0.40; }
    p->ring_w[i + 1] = (float)lw;
  }
  p->ring_w[0] = p->ring_w[1];
  // jax.random.randint(key, shape, fire_age_min, fire_age_max): float bounds truncate
  const int lo = (int)age_min, hi = (int)age_max;
  uint32_t span = (uint32_t)(hi - lo);
  if (hi <= lo) span = 1;
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  p->age_lo = lo;
  p->age_span = span;
  p->age_mult = mult;
  if (lo <= K) return fail(GCA_ERR_UNSUPPORTED, "gca_params_init: fire_age_min must exceed K");
  p->day_length = 400;  // advanced_bulldozer.py:733
  p->p_tree = (float)p_tree;
  p->p_wind_change = (float)p_wind_change;
  p->t_any = (float)t_any;
  // advanced_bulldozer.py:238-246: scale = (nrows + ncols) // 2
  const int scale = (nrows + ncols) / 2;
  const double tm = t_move < 0 ? (1.0 / (speed_move * scale)) - t_any : t_move;
  const double ts = t_shoot < 0 ? (1.0 / (speed_act * scale)) - tm : t_shoot;
  for (int i = 0; i < 9; ++i) p->t_move[i] = (float)tm;  // every move costs the same (:746-754)
  for (int i = 0; i < 2; ++i) p->t_shoot[i] = (float)ts;
  // ca_alexandridis_jax.py:170-173
  const float pv[6] = {-999.f, (float)-0.1, (float)0.2, (float)0.5, (float)0.8, (float)1.2};
  const float pd[6] = {-999.f, (float)-0.2, (float)0.2, (float)0.5, (float)0.8, (float)1.2};
  for (int i = 0; i < 6; ++i) {
    p->onep_veg[i] = 1.0f + pv[i];
    p->onep_den[i] = 1.0f + pd[i];
  }
  if (winds72) {
    std::memcpy(p->winds, winds72, sizeof(float) * 72);
  } else {
    // init_utils.py:203-245
    static const int th[8][9] = {
        {45, 0, 45, 90, 0, 90, 135, 180, 135}, {90, 45, 0, 135, 0, 45, 180, 135, 90},
        {135, 90, 45, 180, 0, 0, 135, 90, 45}, {180, 135, 90, 135, 0, 45, 90, 45, 0},
        {135, 180, 135, 90, 0, 90, 45, 0, 45}, {90, 135, 180, 45, 0, 135, 0, 45, 90},
        {45, 90, 135, 0, 0, 180, 45, 90, 135}, {0, 45, 90, 45, 0, 135, 90, 135, 180}};
    for (int k = 0; k < 8; ++k)
      for (int i = 0; i < 9; ++i) {
        const double t = th[k][i] * (3.14159265358979323846 / 180.0);
        const double ft = std::exp(10 * 0.131 * (std::cos(t) - 1));
        p->winds[k * 9 + i] = i == 4 ? 0.0f : (float)(std::exp(0.045 * 10) * ft);
      }
  }
  return GCA_OK;
}

static int check_state(const gca_params* p, const gca_state* s, const char* who) {
  if (!p || !s) return fail(GCA_ERR_ARG, who);
  if (s->N <= 0) return fail(GCA_ERR_ARG, "N must be positive");
  if (!s->cell || !s->death || !s->doused || !s->tick || !s->key || !s->wind_index)
    return fail(GCA_ERR_ARG, "state: null cell/death/doused/tick/key/wind_index");
  return GCA_OK;
}

static int env_step_impl(const gca_params* p, const gca_state* s, const int32_t* actions, const gca_step_out* out,
                         const gca_inject* inj, const gca_state* snapshot, const float* snapshot_reward, uint32_t flags,
                         void* stream, bool completion_word);

int gca_env_step(const gca_params* p, const gca_state* s, const int32_t* actions, const gca_step_out* out,
                 const gca_inject* inj, const gca_state* snapshot, const float* snapshot_reward, uint32_t flags,
                 void* stream) {
  // device-resident callers order their work by the stream: no completion word, whatever the struct holds
  return env_step_impl(p, s, actions, out, inj, snapshot, snapshot_reward, flags, stream, false);
}

static int env_step_impl(const gca_params* p, const gca_state* s, const int32_t* actions, const gca_step_out* out,
                         const gca_inject* inj, const gca_state* snapshot, const float* snapshot_reward, uint32_t flags,
                         void* stream, bool completion_word) {
  int rc = check_state(p, s, "gca_env_step: null params/state");
  if (rc) return rc;
  if (!(flags & GCA_FLAG_CA_ONLY)) {
    if (!actions || !s->position || !s->time || !s->time_step || !s->is_night)
      return fail(GCA_ERR_ARG, "gca_env_step: null actions/position/time/time_step/is_night");
  }
  if (!(flags & GCA_FLAG_NO_HIDDEN) && !s->hidden) return fail(GCA_ERR_ARG, "gca_env_step: hidden is null");
  if ((flags & GCA_FLAG_AUTO_RESET) && (!snapshot || !snapshot_reward))
    return fail(GCA_ERR_ARG, "gca_env_step: auto-reset needs snapshot and snapshot_reward");
  gca_step_out o;
  std::memset(&o, 0, sizeof(o));
  if (out) o = *out;
  if (!completion_word) o.host_done = nullptr;
  gca_inject j;
  std::memset(&j, 0, sizeof(j));
  if (inj) j = *inj;
  gca_state sn;
  std::memset(&sn, 0, sizeof(sn));
  if (snapshot) sn = *snapshot;
  gca_state st = *s;
  if (flags & GCA_FLAG_NO_HIDDEN) { st.hidden = nullptr; st.pslope = nullptr; }
  if (flags & GCA_FLAG_RENDER) {
    if (!is64(p)) return fail(GCA_ERR_UNSUPPORTED, "gca_env_step: GCA_FLAG_RENDER is a feature of the 64x64 kernel (use gca_render_rgb)");
    if (!o.rgb || (flags & GCA_FLAG_CA_ONLY)) return fail(GCA_ERR_ARG, "gca_env_step: GCA_FLAG_RENDER needs out->rgb and a full env step");
  }
  if (is64(p)) {
    if (!s->row_min || !s->bb) return fail(GCA_ERR_ARG, "gca_env_step: 64x64 path needs row_min and bb");
    if ((flags & GCA_FLAG_AUTO_RESET) && !sn.bb) return fail(GCA_ERR_ARG, "gca_env_step: the snapshot needs bb");
    return check_cuda(gca::launch_env_step64(*p, st, actions, o, j, sn, snapshot_reward, flags, (cudaStream_t)stream),
                      "env_step64");
  }
  if (gca::bb_supported(*p) && !(flags & GCA_FLAG_GENERIC_TILES)) {
    // grids of whole 64-bit words up to 256x256: one launch per env step, the grid as bit-boards in shared memory
    rc = check_cuda(gca::launch_bb_env_step(*p, st, actions, o, j, flags, (cudaStream_t)stream), "env_step_bb");
  } else {
    // any other grid: generic tiled kernels (a tile list per env step), one per CA sub-step + 3, replayed as one CUDA
    // graph -- the fused conditional_reset included
    if (((long long)p->H * p->W) & 1) return fail(GCA_ERR_UNSUPPORTED, "gca_env_step: H*W must be even");
    if (!s->scratch_cell || !s->scratch_u32)
      return fail(GCA_ERR_ARG, "gca_env_step: grids other than 64x64 need scratch_cell and scratch_u32");
    const bool reset = (flags & GCA_FLAG_AUTO_RESET) != 0;
    if (reset && !o.terminated) return fail(GCA_ERR_ARG, "gca_env_step: auto-reset on the tiled path needs out->terminated");
    return check_cuda(gca::launch_tiled_env_step(*p, st, actions, o, j, flags, s->scratch_cell, s->scratch_u32,
                                                 reinterpret_cast<int32_t*>(s->scratch_u32) + 12 * GCA_MAX_K * (size_t)s->N,
                                                 reinterpret_cast<uint8_t*>(s->scratch_u32 + (12 * GCA_MAX_K + 2) * (size_t)s->N),
                                                 (flags & GCA_FLAG_NO_TMA) ? 0 : 1, reset ? &sn : nullptr, snapshot_reward,
                                                 (cudaStream_t)stream),
                      "env_step_tiled");
  }
  if (rc) return rc;
  if (flags & GCA_FLAG_AUTO_RESET) {
    if (!o.terminated) return fail(GCA_ERR_ARG, "gca_env_step: auto-reset on the tiled path needs out->terminated");
    return check_cuda(gca::launch_auto_reset(*p, st, sn, snapshot_reward, o.reward, o.terminated, (cudaStream_t)stream),
                      "auto_reset");
  }
  return GCA_OK;
}

// Device-visible alias of [host, host + bytes) when that range is pinned host memory mapped into the device's
// address space (cudaHostAlloc / cudaHostRegister under unified addressing); false for pageable memory.
static bool mapped_device_pointer(const void* host, size_t bytes, void** dev) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return false;
  cudaPointerAttributes b;  // the last byte must belong to a pinned range as well
  if (cudaPointerGetAttributes(&b, static_cast<const char*>(host) + bytes - 1) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (b.type != cudaMemoryTypeHost || b.devicePointer == nullptr) return false;
  *dev = a.devicePointer;
  return true;
}

int gca_env_step_host(const gca_params* p, const gca_state* s, const int32_t* host_actions, int32_t* dev_actions,
                      const gca_step_out* out, const gca_state* snapshot, const float* snapshot_reward, uint32_t flags,
                      float* host_reward, uint8_t* host_terminated, void* stream) {
  if (!p || !s || !host_actions || !dev_actions || !out || !out->reward || !out->terminated || !host_reward ||
      !host_terminated)
    return fail(GCA_ERR_ARG, "gca_env_step_host: null argument (out->reward and out->terminated are required)");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t N = (size_t)s->N;
  // transport per direction: the kernel itself over the bus (pinned, device-visible buffers; 64x64 kernel), else
  // the copy engine
  void *da = nullptr, *dr = nullptr, *dt = nullptr, *dw = nullptr;
  bool in_mapped, out_mapped, word_mapped = false;
  if ((flags & GCA_FLAG_HOST_MAPPED) && is64(p)) {
    // the caller vouches for cudaHostAlloc'ed buffers: under unified addressing their device address is the host address
    da = const_cast<int32_t*>(host_actions); dr = host_reward; dt = host_terminated; dw = out->host_done;
    in_mapped = !(flags & GCA_FLAG_HOST_COPY_IN);
    out_mapped = !(flags & GCA_FLAG_HOST_COPY_OUT);
    word_mapped = out->host_done != nullptr;
  } else {
    in_mapped = !(flags & GCA_FLAG_HOST_COPY_IN) && is64(p) &&
                mapped_device_pointer(host_actions, N * 3 * sizeof(int32_t), &da);
    out_mapped = !(flags & GCA_FLAG_HOST_COPY_OUT) && is64(p) &&
                 mapped_device_pointer(host_reward, N * sizeof(float), &dr) &&
                 mapped_device_pointer(host_terminated, N, &dt);
    if (out_mapped && out->host_done && out->done_counter)
      word_mapped = mapped_device_pointer(out->host_done, sizeof(uint32_t), &dw);
  }
  int rc;
  if (!in_mapped) {
    rc = check_cuda(cudaMemcpyAsync(dev_actions, host_actions, N * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, st),
                    "env_step_host: actions H2D");
    if (rc) return rc;
  }
  gca_step_out o = *out;
  bool use_word = false;
  o.host_done = nullptr;
  if (out_mapped) {
    o.host_reward = static_cast<float*>(dr);
    o.host_terminated = static_cast<uint8_t*>(dt);
    if (word_mapped && out->done_counter) {
      o.host_done = static_cast<uint32_t*>(dw);  // the kernel needs the device-visible alias of the word
      use_word = true;
    }
  }
  rc = env_step_impl(p, s, in_mapped ? static_cast<const int32_t*>(da) : dev_actions, &o, nullptr, snapshot,
                     snapshot_reward, flags, stream, use_word);
  if (rc) return rc;
  if (!out_mapped) {
    const bool adjacent = out->terminated == reinterpret_cast<const uint8_t*>(out->reward + N) &&
                          host_terminated == reinterpret_cast<const uint8_t*>(host_reward + N);
    if (adjacent) {  // [N] f32 + [N] u8 laid out back to back on both sides: one copy
      rc = check_cuda(cudaMemcpyAsync(host_reward, out->reward, N * 5, cudaMemcpyDeviceToHost, st),
                      "env_step_host: reward + terminated D2H");
      if (rc) return rc;
    } else {
      rc = check_cuda(cudaMemcpyAsync(host_reward, out->reward, N * sizeof(float), cudaMemcpyDeviceToHost, st),
                      "env_step_host: reward D2H");
      if (rc) return rc;
      rc = check_cuda(cudaMemcpyAsync(host_terminated, out->terminated, N, cudaMemcpyDeviceToHost, st),
                      "env_step_host: terminated D2H");
      if (rc) return rc;
    }
  }
  if (flags & GCA_FLAG_HOST_ASYNC) return GCA_OK;  // the caller waits: gca_host_wait
  return gca_host_wait(use_word ? out->host_done : nullptr, out->done_token, 30.0, stream);
}

int gca_host_wait(const uint32_t* host_done, uint32_t token, double timeout_s, void* stream) {
  if (!host_done) return check_cuda(cudaStreamSynchronize((cudaStream_t)stream), "host_wait: synchronize");
  const volatile uint32_t* w = host_done;
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spins = 0;; ++spins) {
    if (*w == token) {
      std::atomic_thread_fence(std::memory_order_acquire);
      return GCA_OK;
    }
    if ((spins & 0x3FFFu) == 0x3FFFu) {
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (dt > timeout_s) {
        const cudaError_t e = cudaStreamQuery((cudaStream_t)stream);
        if (e != cudaSuccess && e != cudaErrorNotReady) return check_cuda(e, "host_wait");
        return fail(GCA_ERR_CUDA, "gca_host_wait: the completion word did not arrive in time");
      }
    }
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
}

int gca_alexandridis_step(const gca_params* p, const gca_state* s, const gca_step_out* out, const gca_inject* inj,
                          uint32_t flags, void* stream) {
  return gca_env_step(p, s, nullptr, out, inj, nullptr, nullptr, (flags | GCA_FLAG_CA_ONLY) & ~GCA_FLAG_AUTO_RESET,
                      stream);
}

int gca_move_modify(const gca_params* p, const gca_state* s, const int32_t* actions, void* stream) {
  if (!p || !s || !actions || !s->position || !s->doused) return fail(GCA_ERR_ARG, "gca_move_modify: null argument");
  return check_cuda(gca::launch_move_modify(*p, *s, actions, (cudaStream_t)stream), "move_modify");
}

int gca_reward_done(const gca_params* p, const gca_state* s, float* reward, uint8_t* terminated, int32_t* counts,
                    void* stream) {
  if (!p || !s || !s->cell) return fail(GCA_ERR_ARG, "gca_reward_done: null argument");
  return check_cuda(gca::launch_reward_done(*p, *s, reward, terminated, counts, (cudaStream_t)stream), "reward_done");
}

int gca_conditional_reset(const gca_params* p, const gca_state* s, const gca_state* snapshot,
                          const float* snapshot_reward, float* reward, uint8_t* terminated, void* stream) {
  int rc = check_state(p, s, "gca_conditional_reset: null params/state");
  if (rc) return rc;
  if (!snapshot || !snapshot_reward || !terminated) return fail(GCA_ERR_ARG, "gca_conditional_reset: null argument");
  return check_cuda(
      gca::launch_conditional_reset(*p, *s, *snapshot, snapshot_reward, reward, terminated, (cudaStream_t)stream),
      "conditional_reset");
}

int gca_render_rgb(const gca_params* p, int32_t N, const uint8_t* cell, const uint64_t* doused,
                   const int32_t* position, const uint8_t* night, const int32_t* ext_action,
                   const uint8_t* env_mask, int32_t enable_extensions, int32_t rgb_u8, uint32_t* scratch,
                   void* rgb_out, void* stream) {
  if (!p || N <= 0 || !cell || !doused || !position || !night || !rgb_out)
    return fail(GCA_ERR_ARG, "gca_render_rgb: null argument");
  if (enable_extensions && !scratch) return fail(GCA_ERR_ARG, "gca_render_rgb: extensions need a [N] u32 scratch");
  return check_cuda(gca::launch_render(*p, N, cell, doused, position, night, ext_action, 1, env_mask,
                                       enable_extensions, rgb_u8, scratch, rgb_out, (cudaStream_t)stream),
                    "render_rgb");
}

int gca_render_rgb_actions(const gca_params* p, int32_t N, const uint8_t* cell, const uint64_t* doused,
                           const int32_t* position, const uint8_t* night, const int32_t* actions,
                           const uint8_t* env_mask, int32_t enable_extensions, int32_t rgb_u8, uint32_t* scratch,
                           void* rgb_out, void* stream) {
  if (!p || N <= 0 || !cell || !doused || !position || !night || !rgb_out)
    return fail(GCA_ERR_ARG, "gca_render_rgb_actions: null argument");
  if (enable_extensions && !scratch)
    return fail(GCA_ERR_ARG, "gca_render_rgb_actions: extensions need a [N] u32 scratch");
  return check_cuda(gca::launch_render(*p, N, cell, doused, position, night, actions ? actions + 2 : nullptr, 3, env_mask,
                                       enable_extensions, rgb_u8, scratch, rgb_out, (cudaStream_t)stream),
                    "render_rgb_actions");
}

int gca_pack_state(const gca_params* p, const gca_state* s, const float* true_grid, const float* fire_age,
                   const int32_t* dousing_count, const int32_t* vegetation, const int32_t* density,
                   uint8_t* hidden_out, int32_t* err_flag, void* stream) {
  int rc = check_state(p, s, "gca_pack_state: null params/state");
  if (rc) return rc;
  if (!true_grid || !fire_age || !dousing_count) return fail(GCA_ERR_ARG, "gca_pack_state: null input array");
  if (hidden_out && (!vegetation || !density)) return fail(GCA_ERR_ARG, "gca_pack_state: null vegetation/density");
  return check_cuda(gca::launch_pack(*p, *s, true_grid, fire_age, dousing_count, vegetation, density, hidden_out,
                                     err_flag, (cudaStream_t)stream),
                    "pack_state");
}

int gca_unpack_state(const gca_params* p, const gca_state* s, float* true_grid, float* fire_age,
                     int32_t* dousing_count, void* stream) {
  int rc = check_state(p, s, "gca_unpack_state: null params/state");
  if (rc) return rc;
  return check_cuda(gca::launch_unpack(*p, *s, true_grid, fire_age, dousing_count, (cudaStream_t)stream),
                    "unpack_state");
}

int gca_generate_hidden(int32_t N, int32_t H, int32_t W, uint64_t seed, int32_t env_offset, int32_t* vegetation,
                        int32_t* density, float* altitude, double* altitude_f64, float* slope9, float* pslope9,
                        void* stream) {
  if (N <= 0 || !vegetation || !density || !altitude || !altitude_f64 || !pslope9)
    return fail(GCA_ERR_ARG, "gca_generate_hidden: null argument");
  if (H < 16 || W < 16 || (long long)H * W > 0x7FFFFFFFll || env_offset < 0)
    return fail(GCA_ERR_UNSUPPORTED, "gca_generate_hidden: grids from 16 x 16 up to 2^31 cells, env_offset >= 0");
  return check_cuda(gca::launch_generate_hidden(N, H, W, (unsigned long long)seed, env_offset, vegetation, density, altitude,
                                                altitude_f64, slope9, pslope9, (cudaStream_t)stream), "generate_hidden");
}

int gca_episode_stats_update(int32_t N, const gca_episode_stats* st, const float* step_reward,
                             const uint8_t* terminated, const uint8_t* truncated, const uint8_t* obs_night,
                             const int32_t* actions, void* stream) {
  if (N <= 0 || !st || !step_reward || !terminated || !actions)
    return fail(GCA_ERR_ARG, "gca_episode_stats_update: null argument");
  const void* const* fields = reinterpret_cast<const void* const*>(st);
  for (size_t i = 0; i < sizeof(gca_episode_stats) / sizeof(void*); ++i)
    if (!fields[i]) return fail(GCA_ERR_ARG, "gca_episode_stats_update: every gca_episode_stats pointer must be set");
  return check_cuda(gca::launch_episode_stats(N, *st, step_reward, terminated, truncated, obs_night, actions,
                                              (cudaStream_t)stream), "episode_stats");
}

int gca_balance_order(int32_t N, const uint32_t* work, int32_t* order, void* stream) {
  if (N <= 0 || !work || !order) return fail(GCA_ERR_ARG, "gca_balance_order: bad argument");
  return check_cuda(gca::launch_balance_order(N, work, order, (cudaStream_t)stream), "balance_order");
}

int gca_windy_env_step(int32_t N, int32_t H, int32_t W, uint64_t* tree_bb, uint64_t* fire_bb, int32_t* position,
                       double* time, const int32_t* actions, const double* wind9, const double* rolls,
                       int32_t rmax, double t_move, double t_shoot, double t_any, double* reward,
                       uint8_t* terminated, int32_t* counts, int32_t* repeats_out, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || !tree_bb || !fire_bb || !position || !time || !actions || !wind9)
    return fail(GCA_ERR_ARG, "gca_windy_env_step: null/bad argument");
  if (rmax < 0 || (rmax > 0 && !rolls)) return fail(GCA_ERR_ARG, "gca_windy_env_step: rolls missing");
  if ((size_t)3 * H * ((W + 63) / 64) * 8 > 200 * 1024)
    return fail(GCA_ERR_UNSUPPORTED, "gca_windy_env_step: grid does not fit shared memory (max ~512x512)");
  return check_cuda(gca::launch_windy_step(N, H, W, (unsigned long long*)tree_bb, (unsigned long long*)fire_bb, position,
                                           time, actions, wind9, rolls, rmax, t_move, t_shoot, t_any, reward,
                                           terminated, counts, repeats_out, (cudaStream_t)stream),
                    "windy_env_step");
}

int gca_windy_pack(int32_t N, int32_t H, int32_t W, const uint8_t* cell, uint64_t* tree_bb, uint64_t* fire_bb,
                   void* stream) {
  if (N <= 0 || !cell || !tree_bb || !fire_bb) return fail(GCA_ERR_ARG, "gca_windy_pack: bad argument");
  return check_cuda(gca::launch_windy_pack(N, H, W, cell, (unsigned long long*)tree_bb, (unsigned long long*)fire_bb,
                                           (cudaStream_t)stream), "windy_pack");
}

int gca_windy_unpack(int32_t N, int32_t H, int32_t W, const uint64_t* tree_bb, const uint64_t* fire_bb,
                     uint8_t* cell, void* stream) {
  if (N <= 0 || !cell || !tree_bb || !fire_bb) return fail(GCA_ERR_ARG, "gca_windy_unpack: bad argument");
  return check_cuda(gca::launch_windy_unpack(N, H, W, (const unsigned long long*)tree_bb,
                                             (const unsigned long long*)fire_bb, cell, (cudaStream_t)stream),
                    "windy_unpack");
}

int gca_threefry_bits(const uint32_t* key2_dev, int64_t n, int32_t rng_mode, uint32_t* out_dev, void* stream) {
  if (!key2_dev || !out_dev || n <= 0) return fail(GCA_ERR_ARG, "gca_threefry_bits: bad argument");
  return check_cuda(gca::launch_threefry_bits(key2_dev, (long long)n, rng_mode, out_dev, (cudaStream_t)stream),
                    "threefry_bits");
}

int gca_threefry_split(const uint32_t* key2_dev, int32_t num, int32_t rng_mode, uint32_t* out_dev, void* stream) {
  if (!key2_dev || !out_dev || num <= 0) return fail(GCA_ERR_ARG, "gca_threefry_split: bad argument");
  if (rng_mode == GCA_RNG_LEGACY)  // split(key, num) = bits(key, 2 num).reshape(num, 2)
    return check_cuda(gca::launch_threefry_bits(key2_dev, 2ll * num, rng_mode, out_dev, (cudaStream_t)stream),
                      "threefry_split");
  return check_cuda(gca::launch_threefry_split_part(key2_dev, num, out_dev, (cudaStream_t)stream), "threefry_split");
}

}  // extern "C"
