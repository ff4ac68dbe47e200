// gca_aux.cu -- boundary kernels around the fused step: reference-layout <-> packed state,
// stand-alone move/douse, reward/done, conditional_reset, RGB observation, PRNG test hooks.
// Reference lines are cited per kernel (paths relative to
// /root/reference/gym_cellular_automata/).
#include <cstdlib>
#include <cstring>

#include "gca_common.cuh"

namespace gca {

// ---------------------------------------------------------------------------------------------
// pack / unpack: the float32/int32 context pytree of _initial_context_distribution
// (forest_fire/bulldozer/advanced_bulldozer.py:690-743) <-> packed state.
// One warp per grid row segment of 64 columns.
// ---------------------------------------------------------------------------------------------
__global__ void pack_state_kernel(gca_params P, gca_state S, const float* __restrict__ grid,
                                  const float* __restrict__ fire_age, const int32_t* __restrict__ dousing,
                                  const int32_t* __restrict__ veg, const int32_t* __restrict__ den,
                                  uint8_t* __restrict__ hidden_out, int32_t* err_flag) {
  const int H = P.H, W = P.W, WW = (W + 63) >> 6;
  const int lane = threadIdx.x & 31;
  const long long seg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nseg = (long long)S.N * H * WW;
  if (seg >= nseg) return;
  const int ws = (int)(seg % WW);
  const long long er = seg / WW;  // e * H + r
  const int e = (int)(er / H);
  const uint32_t tick = S.tick[e];
  uint32_t rowmin = 0xFFFFFFFFu;
  unsigned long long dmask = 0, tmask = 0, fmask = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = ws * 64 + h * 32 + lane;
    bool dbit = false;
    int ccode = 0;
    if (c < W) {
      const size_t i = (size_t)er * W + c;
      const float g = grid[i];
      const int code = g == 2.0f ? 2 : (g == 1.0f ? 1 : 0);
      if (!(g == 0.0f || g == 1.0f || g == 2.0f) && err_flag) atomicOr(err_flag, 1);
      S.cell[i] = (uint8_t)code;
      ccode = code;
      const float a = fire_age[i];
      const int ai = (int)a;
      uint32_t field;
      if (code == 2) {
        if (!(a >= 1.0f && a <= 32767.0f && (float)ai == a) && err_flag) atomicOr(err_flag, 2);
        const uint32_t dabs = tick + (uint32_t)max(ai, 1) - 1u;  // burns out at this tick
        field = dabs & 0xFFFFu;
        rowmin = min(rowmin, dabs);
      } else {
        if (!(a >= 0.0f && a <= 65535.0f && (float)ai == a) && err_flag) atomicOr(err_flag, 4);
        field = (uint32_t)min(max(ai, 0), 65535);
      }
      S.death[i] = (uint16_t)field;
      if (hidden_out != nullptr) {
        const int v = veg[i], d = den[i];
        if ((v < 0 || v > 7 || d < 0 || d > 7) && err_flag) atomicOr(err_flag, 8);
        hidden_out[i] = (uint8_t)((v & 7) | ((d & 7) << 3));
      }
      const int dc = dousing[i];
      if (!(dc == 0 || dc == 1) && err_flag) atomicOr(err_flag, 16);
      dbit = dc != 0;
    }
    dmask |= (unsigned long long)__ballot_sync(GCA_FULL, dbit) << (32 * h);
    tmask |= (unsigned long long)__ballot_sync(GCA_FULL, ccode == 1) << (32 * h);
    fmask |= (unsigned long long)__ballot_sync(GCA_FULL, ccode == 2) << (32 * h);
  }
  if (lane == 0) S.doused[seg] = dmask;
  if (lane == 0 && S.bb != nullptr) { S.bb[2 * seg] = tmask; S.bb[2 * seg + 1] = fmask; }
  // row_min is only defined (and only used) for single-segment rows, i.e. W <= 64
  rowmin = __reduce_min_sync(GCA_FULL, rowmin);
  if (lane == 0 && S.row_min != nullptr && WW == 1) S.row_min[er] = rowmin;
}

__global__ void unpack_state_kernel(gca_params P, gca_state S, float* __restrict__ grid,
                                    float* __restrict__ fire_age, int32_t* __restrict__ dousing) {
  const int H = P.H, W = P.W, WW = (W + 63) >> 6;
  const size_t n = (size_t)S.N * H * W;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % W);
  const size_t er = i / W;
  const int e = (int)(er / H);
  const int code = S.cell[i];
  if (grid) grid[i] = (float)code;
  if (fire_age) {
    const uint32_t f = S.death[i];
    fire_age[i] = code == 2 ? (float)(((f - S.tick[e]) & 0xFFFFu) + 1u) : (float)f;
  }
  if (dousing) dousing[i] = (int32_t)((S.doused[er * WW + (c >> 6)] >> (c & 63)) & 1ull);
}

// ---------------------------------------------------------------------------------------------
// MoveJax / ModifyJax (forest_fire/operators/move_modify_jax.py:39-62,102-114)
// ---------------------------------------------------------------------------------------------
__global__ void move_modify_kernel(gca_params P, gca_state S, const int32_t* __restrict__ actions) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S.N) return;
  const int WW = (P.W + 63) >> 6;
  int row = S.position[2 * e], col = S.position[2 * e + 1];
  move_position(actions[3 * e], P.H, P.W, row, col);
  S.position[2 * e] = row;
  S.position[2 * e + 1] = col;
  if (actions[3 * e + 1] == 1) S.doused[((size_t)e * P.H + row) * WW + (col >> 6)] |= 1ull << (col & 63);
}

// ---------------------------------------------------------------------------------------------
// count_cells + _award + _is_done (advanced_bulldozer.py:597-633,941-953): one CTA per env
// ---------------------------------------------------------------------------------------------
__global__ void reward_done_kernel(gca_params P, gca_state S, float* reward, uint8_t* terminated, int32_t* counts) {
  const int e = blockIdx.x;
  const size_t n = (size_t)P.H * P.W;
  const uint8_t* g = S.cell + (size_t)e * n;
  int t = 0, f = 0;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = g[i];
    t += c == 1;
    f += c == 2;
  }
  __shared__ int st, sf;
  if (threadIdx.x == 0) { st = 0; sf = 0; }
  __syncthreads();
  t = __reduce_add_sync(GCA_FULL, t);
  f = __reduce_add_sync(GCA_FULL, f);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&st, t); atomicAdd(&sf, f); }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (reward) reward[e] = award(st, sf);
    if (terminated) terminated[e] = sf == 0;
    if (counts) { counts[2 * e] = st; counts[2 * e + 1] = sf; }
  }
}

// ---------------------------------------------------------------------------------------------
// conditional_reset (advanced_bulldozer.py:422-518): one CTA per env; envs that are not
// terminated exit immediately.
// ---------------------------------------------------------------------------------------------
__global__ void conditional_reset_kernel(gca_params P, gca_state S, gca_state SNAP, const float* __restrict__ snap_reward,
                                         float* reward, uint8_t* terminated, int clear_flag) {
  const int e = blockIdx.x;
  if (!terminated[e]) return;
  const int H = P.H, W = P.W, WW = (W + 63) >> 6;
  const size_t n = (size_t)H * W, base = (size_t)e * n;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
    S.cell[base + i] = SNAP.cell[base + i];
    S.death[base + i] = SNAP.death[base + i];
  }
  for (size_t i = threadIdx.x; i < (size_t)H * WW; i += blockDim.x)
    S.doused[(size_t)e * H * WW + i] = SNAP.doused[(size_t)e * H * WW + i];
  if (S.row_min != nullptr && SNAP.row_min != nullptr && WW == 1)
    for (int i = threadIdx.x; i < H; i += blockDim.x) S.row_min[(size_t)e * H + i] = SNAP.row_min[(size_t)e * H + i];
  if (S.bb != nullptr && SNAP.bb != nullptr)
    for (size_t i = threadIdx.x; i < (size_t)H * WW * 2; i += blockDim.x)
      S.bb[(size_t)e * H * WW * 2 + i] = SNAP.bb[(size_t)e * H * WW * 2 + i];
  __syncthreads();
  if (threadIdx.x == 0) {
    S.key[2 * e] = SNAP.key[2 * e];
    S.key[2 * e + 1] = SNAP.key[2 * e + 1];
    S.wind_index[e] = SNAP.wind_index[e];
    S.position[2 * e] = SNAP.position[2 * e];
    S.position[2 * e + 1] = SNAP.position[2 * e + 1];
    S.time[e] = SNAP.time[e];
    S.tick[e] = SNAP.tick[e];
    if (S.steps_elapsed) S.steps_elapsed[e] = 0.0f;
    if (S.reward_accumulated) S.reward_accumulated[e] = 0.0f;
    if (reward) reward[e] = snap_reward[e];
    if (clear_flag) terminated[e] = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// Observation (advanced_bulldozer.py:988-1101; extension_utils.py:99-195).
// pass 1 (only with extensions): per-env flags  bit0 any(grid > 0), bit1 any(grid[0,:] > 0),
//                                               bit2 any(blur > 0), bit3 any(blur[0,:] > 0)
// pass 2: one thread per cell.
// blur = round(3 * sum_{3x3, edge padded}(g/3)/9) = (S >= 5) + (S >= 14) with S the integer sum:
// the mean S/9 is never within 0.05 of a rounding boundary, so float32 rounding cannot matter.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int blur_at(const uint8_t* g, int H, int W, int r, int c) {
  int s = 0;
#pragma unroll
  for (int di = -1; di <= 1; ++di) {
    const int rr = min(max(r + di, 0), H - 1);
#pragma unroll
    for (int dj = -1; dj <= 1; ++dj) {
      const int cc = min(max(c + dj, 0), W - 1);
      s += g[(size_t)rr * W + cc];
    }
  }
  return (s >= 5) + (s >= 14);
}

__global__ void render_flags_kernel(gca_params P, const uint8_t* __restrict__ cell, uint32_t* flags) {
  const int e = blockIdx.y;
  const int H = P.H, W = P.W;
  const uint32_t n = (uint32_t)H * (uint32_t)W;  // <= 4096 x 4096: 32-bit index arithmetic
  const uint8_t* g = cell + (size_t)e * n;
  uint32_t fl = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = (int)(i / (uint32_t)W), c = (int)(i - (uint32_t)r * (uint32_t)W);
    if (g[i] > 0) fl |= 1u | (r == 0 ? 2u : 0u);
    if (blur_at(g, H, W, r, c) > 0) fl |= 4u | (r == 0 ? 8u : 0u);
  }
  fl = __reduce_or_sync(GCA_FULL, fl);
  if ((threadIdx.x & 31) == 0 && fl) atomicOr(&flags[e], fl);
}

// extension logic of build_observation_on_extensions for one cell (see render_rgb_kernel)
__device__ __forceinline__ int render_display(int raw, int blur, int ext, uint32_t fl) {
  if (ext == 1) return (fl & 1u) ? ((fl & 2u) ? raw : 0) : blur;
  if (ext == 2) return (fl & 4u) ? ((fl & 8u) ? 0 : blur) : blur;
  return blur;
}

// Fast path (W % 4 == 0, H W % 16 == 0): a CTA renders 1024 consecutive cells of one env, each thread 4 cells
// of a row (one 32-bit load of the grid, one 64-bit load of the dousing board), the pixels are staged in shared
// memory and leave as fully coalesced 128-bit stores -- the observation is write-bound (3 B or 12 B per cell).
constexpr int RENDER_CELLS = 1024;
template <bool U8>
__global__ void __launch_bounds__(256) render_rgb4_kernel(gca_params P, uint32_t chunks, const uint8_t* __restrict__ cell,
                                                          const unsigned long long* __restrict__ doused,
                                                          const int32_t* __restrict__ position,
                                                          const uint8_t* __restrict__ night,
                                                          const int32_t* __restrict__ ext_action, int ext_stride,
                                                          const uint8_t* __restrict__ env_mask, int enable_ext,
                                                          const uint32_t* __restrict__ flags, void* out) {
  extern __shared__ uint4 stage[];  // U8: 3 x 256 32-bit words; float: 3 x 256 float4
  const uint32_t e = blockIdx.x / chunks, chunk = blockIdx.x - e * chunks;
  if (env_mask != nullptr && env_mask[e] == 0) return;
  const int H = P.H, W = P.W, WW = (W + 63) >> 6;
  const uint32_t n = (uint32_t)H * (uint32_t)W;
  const uint32_t c0 = chunk * RENDER_CELLS;
  const uint32_t ci = c0 + threadIdx.x * 4;
  if (ci < n) {
    const int r = (int)(ci / (uint32_t)W), c = (int)(ci - (uint32_t)r * (uint32_t)W);
    const uint8_t* g = cell + (size_t)e * n;
    const uint32_t four = *reinterpret_cast<const uint32_t*>(g + ci);
    const bool ng = night[e] != 0;
    const int pr = position[2 * e], pc = position[2 * e + 1];
    const unsigned long long dw = doused[((size_t)e * H + r) * WW + (c >> 6)] >> (c & 63);
    const int ext = (enable_ext && ext_action) ? ext_action[(size_t)e * ext_stride] : 0;
    const uint32_t fl = enable_ext ? flags[e] : 0u;
    float px[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int disp = (int)((four >> (8 * k)) & 255u);
      if (enable_ext) disp = render_display(disp, blur_at(g, H, W, r, c + k), ext, fl);
      render_pixel(disp, ng, (dw >> k) & 1ull, pr == r && pc == c + k, px[3 * k], px[3 * k + 1], px[3 * k + 2]);
    }
    if (U8) {
      uint32_t* s32 = reinterpret_cast<uint32_t*>(stage) + threadIdx.x * 3;
      uint32_t b[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) b[k] = (uint32_t)(uint8_t)px[k];
#pragma unroll
      for (int j = 0; j < 3; ++j)
        s32[j] = b[4 * j] | (b[4 * j + 1] << 8) | (b[4 * j + 2] << 16) | (b[4 * j + 3] << 24);
    } else {
      float4* s4 = reinterpret_cast<float4*>(stage) + threadIdx.x * 3;
#pragma unroll
      for (int j = 0; j < 3; ++j) s4[j] = make_float4(px[4 * j], px[4 * j + 1], px[4 * j + 2], px[4 * j + 3]);
    }
  }
  __syncthreads();
  const uint32_t cells = min((uint32_t)RENDER_CELLS, n - c0);
  const size_t first = ((size_t)e * n + c0) * 3;  // first output element of the CTA
  if (U8) {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(out) + first);
    for (uint32_t j = threadIdx.x; j < cells * 3 / 16; j += 256) o[j] = stage[j];
  } else {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<float*>(out) + first);
    for (uint32_t j = threadIdx.x; j < cells * 3 / 4; j += 256) o[j] = stage[j];
  }
}

// ---- 16 cells per thread (W % 16 == 0): byte-parallel (SWAR) blur, colours from a per-CTA table --------------
// 16 cells of row r from column c (c % 16 == 0): byte k of word j = cell c + 4 j + k
__device__ __forceinline__ uint4 load16(const uint8_t* g, int W, int r, int c) {
  return *reinterpret_cast<const uint4*>(g + (size_t)r * W + c);
}
// blur of those 16 cells, one byte each: S = 3x3 edge-padded sum (<= 18, no carries between bytes),
// blur = (S >= 5) + (S >= 14); `any` receives a non-zero value iff some blur byte is non-zero
__device__ __forceinline__ uint4 blur16(const uint8_t* g, int H, int W, int r, int c, uint32_t& any) {
  const int rm = max(r - 1, 0), rp = min(r + 1, H - 1);
  const uint4 a = load16(g, W, rm, c), b = load16(g, W, r, c), d = load16(g, W, rp, c);
  uint32_t V[4] = {a.x + b.x + d.x, a.y + b.y + d.y, a.z + b.z + d.z, a.w + b.w + d.w};
  const int cl = max(c - 1, 0), cr = min(c + 16, W - 1);
  const uint32_t vl = (uint32_t)g[(size_t)rm * W + cl] + g[(size_t)r * W + cl] + g[(size_t)rp * W + cl];
  const uint32_t vr = (uint32_t)g[(size_t)rm * W + cr] + g[(size_t)r * W + cr] + g[(size_t)rp * W + cr];
  uint32_t out[4];
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t prev = __funnelshift_l(j ? V[j - 1] : (vl << 24), V[j], 8);   // byte k: V of the cell to the left
    const uint32_t next = __funnelshift_r(V[j], j < 3 ? V[j + 1] : vr, 8);       // byte k: V of the cell to the right
    const uint32_t S = V[j] + prev + next;
    const uint32_t t1 = ((S + 0x7B7B7B7Bu) >> 7) & 0x01010101u;  // S >= 5
    const uint32_t t2 = ((S + 0x72727272u) >> 7) & 0x01010101u;  // S >= 14
    out[j] = t1 + t2;
    acc |= t1;
  }
  any = acc;
  return make_uint4(out[0], out[1], out[2], out[3]);
}

__global__ void __launch_bounds__(256) render_flags16_kernel(gca_params P, const uint8_t* __restrict__ cell,
                                                             uint32_t* flags) {
  const int e = blockIdx.y;
  const int H = P.H, W = P.W;
  const uint32_t n16 = (uint32_t)H * (uint32_t)W / 16u, w16 = (uint32_t)W / 16u;
  const uint8_t* g = cell + (size_t)e * H * W;
  uint32_t fl = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) {
    const int r = (int)(i / w16), c = (int)(i - (uint32_t)r * w16) * 16;
    const uint4 raw = load16(g, W, r, c);
    if (raw.x | raw.y | raw.z | raw.w) fl |= 1u | (r == 0 ? 2u : 0u);
    uint32_t any;
    blur16(g, H, W, r, c, any);
    if (any) fl |= 4u | (r == 0 ? 8u : 0u);
  }
  fl = __reduce_or_sync(GCA_FULL, fl);
  if ((threadIdx.x & 31) == 0 && fl) atomicOr(&flags[e], fl);
}

constexpr int RENDER16_THREADS = 128;
constexpr int RENDER16_CELLS = RENDER16_THREADS * 16;
constexpr int RENDER16_F32_STRIDE = 13;  // float4 per thread in the staging buffer (12 + 1 pad: conflict-free stores)
template <bool U8>
__global__ void __launch_bounds__(RENDER16_THREADS) render_rgb16_kernel(
    gca_params P, uint32_t chunks, const uint8_t* __restrict__ cell, const unsigned long long* __restrict__ doused,
    const int32_t* __restrict__ position, const uint8_t* __restrict__ night, const int32_t* __restrict__ ext_action, int ext_stride,
    const uint8_t* __restrict__ env_mask, int enable_ext, const uint32_t* __restrict__ flags, void* out) {
  extern __shared__ uint4 stage[];
  __shared__ float lut_f[8][3];   // [doused * 4 + min(display value, 3)]
  __shared__ uint32_t lut_u[8];   // the same colours as 0x00BBGGRR
  const uint32_t e = blockIdx.x / chunks, chunk = blockIdx.x - e * chunks;
  if (env_mask != nullptr && env_mask[e] == 0) return;
  const int H = P.H, W = P.W, WW = (W + 63) >> 6;
  const uint32_t n = (uint32_t)H * (uint32_t)W;
  const bool ng = night[e] != 0;
  if (threadIdx.x < 8) {
    float cr, cg, cb;
    render_pixel((int)(threadIdx.x & 3u), ng, threadIdx.x >= 4, false, cr, cg, cb);
    lut_f[threadIdx.x][0] = cr; lut_f[threadIdx.x][1] = cg; lut_f[threadIdx.x][2] = cb;
    lut_u[threadIdx.x] = (uint32_t)(uint8_t)cr | ((uint32_t)(uint8_t)cg << 8) | ((uint32_t)(uint8_t)cb << 16);
  }
  __syncthreads();
  const uint32_t c0 = chunk * RENDER16_CELLS;
  const uint32_t ci = c0 + threadIdx.x * 16;
  if (ci < n) {
    const int r = (int)(ci / (uint32_t)W), c = (int)(ci - (uint32_t)r * (uint32_t)W);
    const uint8_t* g = cell + (size_t)e * n;
    uint4 disp = load16(g, W, r, c);
    if (enable_ext) {
      // channel 0 = blurred grid; extension channel 0 = raw grid (action bit 0), channel 1 = blurred grid (bit 1);
      // the display channel index is the first row holding a positive extension value, clamped to 1
      const int ext = ext_action ? ext_action[(size_t)e * ext_stride] : 0;
      const uint32_t fl = flags[e];
      int src;  // 0: raw grid, 1: zeros, 2: blurred grid -- the same for every cell of the env
      if (ext == 1) src = (fl & 1u) ? ((fl & 2u) ? 0 : 1) : 2;
      else if (ext == 2) src = (fl & 4u) ? ((fl & 8u) ? 1 : 2) : 2;
      else src = 2;
      if (src == 2) { uint32_t any; disp = blur16(g, H, W, r, c, any); }
      else if (src == 1) disp = make_uint4(0u, 0u, 0u, 0u);
    }
    const uint32_t dw = (uint32_t)(doused[((size_t)e * H + r) * WW + (c >> 6)] >> (c & 63)) & 0xFFFFu;
    const uint32_t dv[4] = {disp.x, disp.y, disp.z, disp.w};
    if (U8) {
      uint32_t w[12];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          p[k] = lut_u[(((dw >> (4 * q + k)) & 1u) << 2) | min((dv[q] >> (8 * k)) & 255u, 3u)];
        w[3 * q] = p[0] | (p[1] << 24);
        w[3 * q + 1] = (p[1] >> 8) | (p[2] << 16);
        w[3 * q + 2] = (p[2] >> 16) | (p[3] << 8);
      }
      uint4* s = stage + threadIdx.x * 3;
      s[0] = make_uint4(w[0], w[1], w[2], w[3]);
      s[1] = make_uint4(w[4], w[5], w[6], w[7]);
      s[2] = make_uint4(w[8], w[9], w[10], w[11]);
    } else {
      float* s = reinterpret_cast<float*>(stage + threadIdx.x * RENDER16_F32_STRIDE);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float f[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float* l = lut_f[(((dw >> (4 * q + k)) & 1u) << 2) | min((dv[q] >> (8 * k)) & 255u, 3u)];
          f[3 * k] = l[0]; f[3 * k + 1] = l[1]; f[3 * k + 2] = l[2];
        }
#pragma unroll
        for (int j = 0; j < 3; ++j)
          reinterpret_cast<float4*>(s)[3 * q + j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
      }
    }
    // the bulldozer's pixel is black
    const int pr = position[2 * e], pk = position[2 * e + 1] - c;
    if (pr == r && pk >= 0 && pk < 16) {
      if (U8) {
        uint8_t* s = reinterpret_cast<uint8_t*>(stage + threadIdx.x * 3) + 3 * pk;
        s[0] = 0; s[1] = 0; s[2] = 0;
      } else {
        float* s = reinterpret_cast<float*>(stage + threadIdx.x * RENDER16_F32_STRIDE) + 3 * pk;
        s[0] = 0.f; s[1] = 0.f; s[2] = 0.f;
      }
    }
  }
  __syncthreads();
  const uint32_t cells = min((uint32_t)RENDER16_CELLS, n - c0);
  const size_t first = ((size_t)e * n + c0) * 3;  // first output element of the CTA
  if (U8) {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(out) + first);
    for (uint32_t j = threadIdx.x; j < cells * 3 / 16; j += RENDER16_THREADS) o[j] = stage[j];
  } else {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<float*>(out) + first);
    for (uint32_t j = threadIdx.x; j < cells * 3 / 4; j += RENDER16_THREADS)
      o[j] = stage[(j / 12u) * RENDER16_F32_STRIDE + j % 12u];
  }
}

template <bool U8>
__global__ void render_rgb_kernel(gca_params P, int N, const uint8_t* __restrict__ cell,
                                  const unsigned long long* __restrict__ doused, const int32_t* __restrict__ position,
                                  const uint8_t* __restrict__ night, const int32_t* __restrict__ ext_action, int ext_stride,
                                  const uint8_t* __restrict__ env_mask, int enable_ext,
                                  const uint32_t* __restrict__ flags, void* out) {
  const int H = P.H, W = P.W, WW = (W + 63) >> 6;
  const size_t n = (size_t)H * W;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * n) return;
  const int e = (int)(i / n);
  if (env_mask != nullptr && env_mask[e] == 0) return;
  const size_t ci = i % n;
  const int r = (int)(ci / W), c = (int)(ci % W);
  const uint8_t* g = cell + (size_t)e * n;
  int disp = g[ci];
  if (enable_ext) {
    // channel 0 = blurred grid; extension channel 0 = raw grid (action bit 0), channel 1 = blurred
    // grid (bit 1); the display channel index is the FIRST ROW holding a positive extension
    // value, clamped to 1 (advanced_bulldozer.py:1028-1032).
    const int ext = ext_action ? ext_action[(size_t)e * ext_stride] : 0;
    const uint32_t fl = flags[e];
    const int blur = blur_at(g, H, W, r, c);
    if (ext == 1) {        // bits (1,0): ext0 = raw grid, ext1 = 0
      if (fl & 1u) disp = (fl & 2u) ? disp : 0;
      else disp = blur;
    } else if (ext == 2) { // bits (0,1): ext0 = 0, ext1 = blurred grid
      if (fl & 4u) disp = (fl & 8u) ? 0 : blur;
      else disp = blur;
    } else {
      disp = blur;
    }
  }
  const bool ng = night[e] != 0;
  float cr, cg, cb;
  if (!ng) {
    if (disp == 1) { cr = 169.f; cg = 196.f; cb = 153.f; }       // #A9C499
    else if (disp == 2) { cr = 230.f; cg = 129.f; cb = 129.f; }  // #E68181
    else { cr = 221.f; cg = 209.f; cb = 211.f; }                 // #DDD1D3
  } else {
    if (disp == 1) { cr = 47.f; cg = 79.f; cb = 79.f; }          // #2F4F4F
    else if (disp == 2) { cr = 139.f; cg = 0.f; cb = 0.f; }      // #8B0000
    else { cr = 105.f; cg = 105.f; cb = 105.f; }                 // #696969
  }
  const bool ds = (doused[((size_t)e * H + r) * WW + (c >> 6)] >> (c & 63)) & 1ull;
  if (ds) {  // rgb * (1 - 0.75) + tint * 0.75, float32, tint blue by day / orange by night
    const float tr = ng ? 255.f : 0.f, tg = ng ? 165.f : 0.f, tb = ng ? 0.f : 200.f;
    cr = __fadd_rn(__fmul_rn(cr, 0.25f), __fmul_rn(tr, 0.75f));
    cg = __fadd_rn(__fmul_rn(cg, 0.25f), __fmul_rn(tg, 0.75f));
    cb = __fadd_rn(__fmul_rn(cb, 0.25f), __fmul_rn(tb, 0.75f));
  }
  if (position[2 * e] == r && position[2 * e + 1] == c) { cr = 0.f; cg = 0.f; cb = 0.f; }
  if (U8) {
    uint8_t* o = reinterpret_cast<uint8_t*>(out) + i * 3;
    o[0] = (uint8_t)cr; o[1] = (uint8_t)cg; o[2] = (uint8_t)cb;
  } else {
    float* o = reinterpret_cast<float*>(out) + i * 3;
    o[0] = cr; o[1] = cg; o[2] = cb;
  }
}

// ---------------------------------------------------------------------------------------------
// Load balancing of the 64x64 step kernel.  One CTA orders a chunk of up to 8192 envs by decreasing
// work with a 1024-bucket counting sort in shared memory (the order inside a bucket is irrelevant:
// only the step kernel's speed depends on it, never its results) and deals the sorted envs to the
// warp slots of the step kernel's CTAs (see the dealing rule below; wpc = envs per CTA there).
// ---------------------------------------------------------------------------------------------
constexpr int BAL_CHUNK = 8192;
constexpr int BAL_WAVE = 148;
constexpr int BAL_BUCKETS = 1024;
#ifndef GCA_BALANCE_SKEW_DEFAULT
#define GCA_BALANCE_SKEW_DEFAULT (GCA_S64_GROUPS == 1 ? GCA_S64_WARPS : 0)  /* measured on B200: 88.2 us (0), 85.1 (8), 83.5 (11), 83.3 (14) per step */
#endif
__global__ void __launch_bounds__(1024) balance_order_kernel(int N, int wpc, int groups, int skew, const uint32_t* __restrict__ work, int32_t* order) {
  extern __shared__ int keys[];            // [BAL_CHUNK] env (chunk-local) by rank
  __shared__ int hist[BAL_BUCKETS];        // bucket counts, then start offsets (descending buckets)
  __shared__ uint32_t s_max;
  const int base = blockIdx.x * BAL_CHUNK;
  const int n = min(BAL_CHUNK, N - base);
  const int tid = threadIdx.x;
  hist[tid] = 0;
  if (tid == 0) s_max = 0u;
  __syncthreads();
  uint32_t mx = 0;
  for (int i = tid; i < n; i += blockDim.x) mx = max(mx, work[base + i]);
  mx = __reduce_max_sync(0xFFFFFFFFu, mx);
  if ((tid & 31) == 0) atomicMax(&s_max, mx);
  __syncthreads();
  const int shift = max(0, 32 - __clz(s_max | 1u) - 10);  // work >> shift < 1024
  for (int i = tid; i < n; i += blockDim.x) atomicAdd(&hist[BAL_BUCKETS - 1 - (int)(work[base + i] >> shift)], 1);
  __syncthreads();
  // exclusive scan of the 1024 counts (bucket 0 = heaviest)
  {
    const int c = hist[tid];
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if ((tid & 31) >= d) incl += o;
    }
    __shared__ int wsum[32];
    if ((tid & 31) == 31) wsum[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
      int v = wsum[tid];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (tid >= d) v += o;
      }
      wsum[tid] = v;
    }
    __syncthreads();
    hist[tid] = incl - c + ((tid >> 5) ? wsum[(tid >> 5) - 1] : 0);
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    const int rank = atomicAdd(&hist[BAL_BUCKETS - 1 - (int)(work[base + i] >> shift)], 1);
    keys[rank] = i;
  }
  __syncthreads();
  // Pooled kernel (E warps = E envs per CTA, the heavy phases shared by the CTA): a CTA costs about the
  // SUM of its envs' work, so the sorted envs are dealt like cards, every other round in reverse (snake).
  // With up to two CTAs per SM (C <= 2 G full CTAs, G = 148 SMs) the bins of the deal are the SMs: the block
  // scheduler puts CTA b and CTA b + G on the same SM, and the warp scheduler favours the one launched
  // first (measured: 160 k against 205 k cycles for equal work), so the first CTA of a pair takes the
  // heaviest `skew` rounds of the bin and then every other one (skew = 0: equal sums; default: the heavier
  // half), the second the rest.
  // A trailing partial CTA keeps the lightest ranks.
  const int full = n / wpc;  // CTAs with all warps in range
  const int G = BAL_WAVE;
  const int P = full - G;    // CTA pairs (b, b + G), b < P; CTAs P..G-1 have an SM to themselves
  for (int s = threadIdx.x; s < n; s += blockDim.x) {
    const int b = s / wpc, w = s % wpc;  // CTA slot inside the chunk, warp
    int rank = s;
    if (b < full) {
      if (wpc == 1) {
        const int wave = b / BAL_WAVE, pos = b % BAL_WAVE;
        const int in_wave = min(BAL_WAVE, full - wave * BAL_WAVE);
        const int q = (wave & 1) ? (in_wave - 1 - pos) : pos;
        rank = wave * BAL_WAVE + q;
      } else if (groups == 2) {
        // one CTA per SM holding two lock-step groups: the bin is the CTA, its two groups split the rounds like
        // the two CTAs of an SM above (skew = 0: equal sums)
        const int ge = wpc / 2, g = w / ge, wg = w % ge;
        const int round = g == 0 ? (wg < skew ? wg : skew + 2 * (wg - skew)) : (wg < ge - skew ? skew + 1 + 2 * wg : ge + wg);
        rank = round * full + ((round & 1) ? (full - 1 - b) : b);
      } else if (P <= 0 || P > G) {
        rank = w * full + ((w & 1) ? (full - 1 - b) : b);  // one CTA per SM, or more than two waves
      } else {
        int bin, round;
        if (b < P) { bin = b; round = w < skew ? w : skew + 2 * (w - skew); }
        else if (b >= G) { bin = b - G; round = w < wpc - skew ? skew + 1 + 2 * w : wpc + w; }
        else { bin = b; round = w; }
        if (round < wpc) rank = round * G + ((round & 1) ? (G - 1 - bin) : bin);
        else rank = wpc * G + (round - wpc) * P + ((round & 1) ? (P - 1 - bin) : bin);
      }
    }
    order[base + s] = base + keys[rank];
  }
}

// ---------------------------------------------------------------------------------------------
// Episode statistics of the rollout loop (agents/jax_ppo.py:504-655), one CTA for all envs: thread t
// owns the contiguous env range [t*C, (t+1)*C), so a block scan of the per-thread finished counts
// gives every finished env its rank in env order -- the order the reference's serial scan visits.
// Only the last 10 ranks survive in the ring, and they land on distinct slots.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) episode_stats_kernel(int N, const __grid_constant__ gca_episode_stats E,
                                                             const float* __restrict__ step_reward,
                                                             const uint8_t* __restrict__ terminated,
                                                             const uint8_t* __restrict__ truncated,
                                                             const uint8_t* __restrict__ obs_night,
                                                             const int32_t* __restrict__ actions) {
  __shared__ int s_cnt[1024];
  __shared__ int s_term[32];
  const int tid = threadIdx.x;
  const int C = (N + 1023) / 1024;
  const int lo = min(N, tid * C), hi = min(N, lo + C);
  int nfin = 0, nterm = 0;
  for (int i = lo; i < hi; ++i) {
    const float new_ret = __fadd_rn(E.episode_returns[i], step_reward[i]);
    const int new_len = E.episode_lengths[i] + 1;
    const int night = obs_night ? (int)obs_night[i] : 0;
    const int ext = actions[3 * i + 2];
    E.current_day_correct[i] += (1 - night) * (ext == 2 ? 1 : 0);
    E.current_night_correct[i] += night * (ext == 1 ? 1 : 0);
    E.current_day_steps[i] += 1 - night;
    E.current_night_steps[i] += night;
    const int term = terminated[i], trunc = truncated ? truncated[i] : 0;
    const bool fin = (term + trunc) != 0;
    nterm += term;
    nfin += fin ? 1 : 0;
    // (new) * (1 - terminated) * (1 - truncated)
    E.episode_returns[i] = __fmul_rn(__fmul_rn(new_ret, (float)(1 - term)), (float)(1 - trunc));
    E.episode_lengths[i] = new_len * (1 - term) * (1 - trunc);
    if (fin) {
      E.returned_episode_returns[i] = new_ret;
      E.returned_episode_lengths[i] = new_len;
    }
  }
  s_cnt[tid] = nfin;
  const int wsum = __reduce_add_sync(0xFFFFFFFFu, nterm);
  if ((tid & 31) == 0) s_term[tid >> 5] = wsum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {  // inclusive Hillis-Steele scan
    const int v = tid >= d ? s_cnt[tid - d] : 0;
    __syncthreads();
    s_cnt[tid] += v;
    __syncthreads();
  }
  const int F = s_cnt[1023];
  const int idx0 = E.recent_idx[0];
  int rank = s_cnt[tid] - nfin;
  __syncthreads();
  for (int i = lo; i < hi; ++i) {
    const int term = terminated[i], trunc = truncated ? truncated[i] : 0;
    if ((term + trunc) == 0) continue;
    if (rank >= F - GCA_RECENT) {
      const int slot = (idx0 + rank) % GCA_RECENT;
      E.recent_returns[slot] = E.returned_episode_returns[i];
      E.recent_lengths[slot] = E.returned_episode_lengths[i];
      E.recent_day_correct[slot] = E.current_day_correct[i];
      E.recent_night_correct[slot] = E.current_night_correct[i];
      E.recent_day_steps[slot] = E.current_day_steps[i];
      E.recent_night_steps[slot] = E.current_night_steps[i];
    }
    ++rank;
  }
  if (tid == 0) {
    int t = 0;
    for (int w = 0; w < 32; ++w) t += s_term[w];
    E.amount_finished[0] += t;
    E.recent_idx[0] = (idx0 + F) % GCA_RECENT;
  }
}

// ---------------------------------------------------------------------------------------------
// PRNG test hooks
// ---------------------------------------------------------------------------------------------
__global__ void threefry_bits_kernel(const uint32_t* __restrict__ key, long long n, int mode, uint32_t* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const TfKey k = tf_key(key[0], key[1]);
  if (n == 1) { out[0] = bits_scalar(k, mode); return; }
  if (mode == GCA_RNG_LEGACY) {
    const long long m = n + (n & 1), half = m / 2;
    const bool first = i < half;
    uint32_t c0 = (uint32_t)(first ? i : i - half);
    uint32_t c1 = (uint32_t)(first ? i + half : i);
    if ((n & 1) && (first ? i + half : i) == m - 1) c1 = 0u;  // padded counter
    uint32_t o0, o1;
    threefry2x32(k, c0, c1, o0, o1);
    out[i] = first ? o0 : o1;
  } else {
    uint32_t o0, o1;
    threefry2x32(k, (uint32_t)((unsigned long long)i >> 32), (uint32_t)i, o0, o1);
    out[i] = o0 ^ o1;
  }
}

// partitionable split(key, num)[i] = both words of block (0, i)
__global__ void threefry_split_part_kernel(const uint32_t* __restrict__ key, int num, uint32_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num) return;
  uint32_t o0, o1;
  threefry2x32(tf_key(key[0], key[1]), 0u, (uint32_t)i, o0, o1);
  out[2 * i] = o0;
  out[2 * i + 1] = o1;
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_pack(const gca_params& p, const gca_state& s, const float* grid, const float* fire_age,
                        const int32_t* dousing, const int32_t* veg, const int32_t* den, uint8_t* hidden_out,
                        int32_t* err_flag, cudaStream_t st) {
  const int WW = (p.W + 63) >> 6;
  const long long nseg = (long long)s.N * p.H * WW;
  const int wpb = 8;
  pack_state_kernel<<<(unsigned)((nseg + wpb - 1) / wpb), wpb * 32, 0, st>>>(p, s, grid, fire_age, dousing, veg, den,
                                                                           hidden_out, err_flag);
  return cudaGetLastError();
}
cudaError_t launch_unpack(const gca_params& p, const gca_state& s, float* grid, float* fire_age, int32_t* dousing,
                          cudaStream_t st) {
  const size_t n = (size_t)s.N * p.H * p.W;
  unpack_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, s, grid, fire_age, dousing);
  return cudaGetLastError();
}
cudaError_t launch_move_modify(const gca_params& p, const gca_state& s, const int32_t* actions, cudaStream_t st) {
  move_modify_kernel<<<(s.N + 127) / 128, 128, 0, st>>>(p, s, actions);
  return cudaGetLastError();
}
cudaError_t launch_reward_done(const gca_params& p, const gca_state& s, float* reward, uint8_t* terminated,
                               int32_t* counts, cudaStream_t st) {
  reward_done_kernel<<<s.N, 256, 0, st>>>(p, s, reward, terminated, counts);
  return cudaGetLastError();
}
cudaError_t launch_conditional_reset(const gca_params& p, const gca_state& s, const gca_state& snap,
                                     const float* snap_reward, float* reward, uint8_t* terminated, cudaStream_t st) {
  conditional_reset_kernel<<<s.N, 256, 0, st>>>(p, s, snap, snap_reward, reward, terminated, 1);
  return cudaGetLastError();
}
// fused auto-reset of the tiled path: same restore, but `terminated` keeps the transition's done flag
cudaError_t launch_auto_reset(const gca_params& p, const gca_state& s, const gca_state& snap, const float* snap_reward,
                              float* reward, const uint8_t* terminated, cudaStream_t st) {
  conditional_reset_kernel<<<s.N, 256, 0, st>>>(p, s, snap, snap_reward, reward, const_cast<uint8_t*>(terminated), 0);
  return cudaGetLastError();
}
cudaError_t launch_render(const gca_params& p, int N, const uint8_t* cell, const uint64_t* doused,
                          const int32_t* position, const uint8_t* night, const int32_t* ext_action, int ext_stride,
                          const uint8_t* env_mask, int enable_ext, int rgb_u8, uint32_t* flags_scratch, void* out,
                          cudaStream_t st) {
  const size_t n = (size_t)N * p.H * p.W;
  if (enable_ext) {
    cudaError_t err = cudaMemsetAsync(flags_scratch, 0, sizeof(uint32_t) * N, st);
    if (err != cudaSuccess) return err;
    const size_t per = (size_t)p.H * p.W;
    // blockIdx.y = env: at most 65535 envs per launch
    for (int e0 = 0; e0 < N; e0 += 65535) {
      const unsigned ne = (unsigned)min(65535, N - e0);
      if (p.W % 16 == 0) {
        dim3 grid((unsigned)min((size_t)64, (per / 16 + 255) / 256), ne);
        render_flags16_kernel<<<grid, 256, 0, st>>>(p, cell + (size_t)e0 * per, flags_scratch + e0);
      } else {
        dim3 grid((unsigned)min((size_t)64, (per + 255) / 256), ne);
        render_flags_kernel<<<grid, 256, 0, st>>>(p, cell + (size_t)e0 * per, flags_scratch + e0);
      }
    }
  }
  const size_t per_env = (size_t)p.H * p.W;
  if (p.W % 16 == 0) {
    const uint32_t chunks = (uint32_t)((per_env + RENDER16_CELLS - 1) / RENDER16_CELLS);
    const size_t ctas = (size_t)N * chunks;
    if (ctas <= 0x7FFFFFFFull) {
      if (rgb_u8)
        render_rgb16_kernel<true><<<(unsigned)ctas, RENDER16_THREADS, RENDER16_THREADS * 48, st>>>(
            p, chunks, cell, (const unsigned long long*)doused, position, night, ext_action, ext_stride, env_mask, enable_ext,
            flags_scratch, out);
      else
        render_rgb16_kernel<false><<<(unsigned)ctas, RENDER16_THREADS, RENDER16_THREADS * RENDER16_F32_STRIDE * 16, st>>>(
            p, chunks, cell, (const unsigned long long*)doused, position, night, ext_action, ext_stride, env_mask, enable_ext,
            flags_scratch, out);
      return cudaGetLastError();
    }
  }
  if (p.W % 4 == 0 && per_env % 16 == 0) {
    const uint32_t chunks = (uint32_t)((per_env + RENDER_CELLS - 1) / RENDER_CELLS);
    const size_t ctas = (size_t)N * chunks;
    if (ctas <= 0x7FFFFFFFull) {
      if (rgb_u8)
        render_rgb4_kernel<true><<<(unsigned)ctas, 256, 256 * 12, st>>>(p, chunks, cell, (const unsigned long long*)doused,
                                                                     position, night, ext_action, ext_stride, env_mask, enable_ext,
                                                                     flags_scratch, out);
      else
        render_rgb4_kernel<false><<<(unsigned)ctas, 256, 256 * 48, st>>>(p, chunks, cell,
                                                                      (const unsigned long long*)doused, position, night,
                                                                      ext_action, ext_stride, env_mask, enable_ext, flags_scratch, out);
      return cudaGetLastError();
    }
  }
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (rgb_u8)
    render_rgb_kernel<true><<<blocks, 256, 0, st>>>(p, N, cell, (const unsigned long long*)doused, position, night,
                                                    ext_action, ext_stride, env_mask, enable_ext, flags_scratch, out);
  else
    render_rgb_kernel<false><<<blocks, 256, 0, st>>>(p, N, cell, (const unsigned long long*)doused, position, night,
                                                     ext_action, ext_stride, env_mask, enable_ext, flags_scratch, out);
  return cudaGetLastError();
}
cudaError_t launch_episode_stats(int N, const gca_episode_stats& e, const float* step_reward, const uint8_t* terminated,
                                 const uint8_t* truncated, const uint8_t* obs_night, const int32_t* actions,
                                 cudaStream_t st) {
  episode_stats_kernel<<<1, 1024, 0, st>>>(N, e, step_reward, terminated, truncated, obs_night, actions);
  return cudaGetLastError();
}
cudaError_t launch_balance_order(int N, const uint32_t* work, int32_t* order, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(balance_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BAL_CHUNK * 4);
    attr_set = true;
  }
  const int chunks = (N + BAL_CHUNK - 1) / BAL_CHUNK;
  // GCA_BALANCE_SKEW (0 .. envs per CTA) overrides how much heavier the first CTA of an SM is made
  static const int skew = [] {
    const char* v = getenv("GCA_BALANCE_SKEW");
    const int k = v ? atoi(v) : GCA_BALANCE_SKEW_DEFAULT;
    const int cap = GCA_S64_WARPS / GCA_S64_GROUPS;
    return k < 0 ? 0 : (k > cap ? cap : k);
  }();
  balance_order_kernel<<<chunks, 1024, BAL_CHUNK * 4, st>>>(N, GCA_S64_WARPS, GCA_S64_GROUPS, skew, work, order);
  return cudaGetLastError();
}
cudaError_t launch_threefry_split_part(const uint32_t* key, int num, uint32_t* out, cudaStream_t st) {
  threefry_split_part_kernel<<<(num + 255) / 256, 256, 0, st>>>(key, num, out);
  return cudaGetLastError();
}
cudaError_t launch_threefry_bits(const uint32_t* key, long long n, int mode, uint32_t* out, cudaStream_t st) {
  threefry_bits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(key, n, mode, out);
  return cudaGetLastError();
}

}  // namespace gca
