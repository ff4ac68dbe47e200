// gca_hidden.cu -- device-side generation of the static inputs of the advanced bulldozer env: vegetation and
// density (random rectangles over a random background), altitude (noise + cosine hills + ramps), the slope to
// the 8 neighbours and its factor exp(0.078 slope).
//
// Same layer models as /root/reference/gym_cellular_automata/forest_fire/bulldozer/utils/init_utils.py:10-73
// (patches), :76-116 (altitude), :166-200 (get_slope) and ca_alexandridis_jax.py:199-200 (slope factor).  The
// reference draws from Python's unseeded generators in per-cell loops (minutes at 65536 envs or 4096^2); here
// every random number is one counter-based threefry2x32 block addressed by (stream, env, index), so an env's
// layers depend only on (seed, global env index) -- not on the batch size or on how the envs are sharded over
// GPUs -- and oracle/hidden_device.py restates the generator in NumPy for the tests.
#include "gca_common.cuh"

namespace gca {
namespace {

enum { HS_VEG_RECT = 1, HS_VEG_FILL = 2, HS_DEN_RECT = 3, HS_DEN_FILL = 4, HS_ALT_NOISE = 5, HS_ALT_HILL = 6, HS_ALT_RAMP = 7 };

// word pair of (stream, env, idx): threefry2x32 with key (seed_lo ^ stream, seed_hi), counter (env, idx)
__device__ __forceinline__ void hrand(unsigned long long seed, uint32_t stream, uint32_t env, uint32_t idx, uint32_t& a,
                                      uint32_t& b) {
  threefry2x32(tf_key((uint32_t)seed ^ stream, (uint32_t)(seed >> 32)), env, idx, a, b);
}
// uniform double in [0, 1) from 53 bits
__device__ __forceinline__ double hunif(uint32_t a, uint32_t b) {
  return (double)((((unsigned long long)a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
}

// vegetation / density: 4..7 rectangles of type 1..5 (later ones overwrite earlier ones), the uncovered cells 1..3
__global__ void __launch_bounds__(256) hidden_patches_kernel(int H, int W, unsigned long long seed, int env_offset,
                                                             uint32_t rect_stream, uint32_t fill_stream, int32_t* out) {
  __shared__ int s_n, s_r0[8], s_r1[8], s_c0[8], s_c1[8], s_kind[8];
  const int e = blockIdx.y;
  const uint32_t ge = (uint32_t)(env_offset + e);
  const int tid = threadIdx.x;
  if (tid < 7) {
    uint32_t a, b, c, d, k, z;
    hrand(seed, rect_stream, ge, 3u * tid, a, b);
    hrand(seed, rect_stream, ge, 3u * tid + 1u, c, d);
    hrand(seed, rect_stream, ge, 3u * tid + 2u, k, z);
    const int cr = (int)(a % (uint32_t)H), cc = (int)(b % (uint32_t)W);
    const int ph = 3 + (int)(c % (uint32_t)(H / 2 - 3)), pw = 3 + (int)(d % (uint32_t)(W / 2 - 3));
    s_r0[tid] = max(0, cr - ph / 2); s_r1[tid] = min(H, cr + ph / 2);
    s_c0[tid] = max(0, cc - pw / 2); s_c1[tid] = min(W, cc + pw / 2);
    s_kind[tid] = 1 + (int)(k % 5u);
  }
  if (tid == 7) {
    uint32_t a, b;
    hrand(seed, rect_stream, ge, 100u, a, b);
    s_n = 4 + (int)(a % 4u);
  }
  __syncthreads();
  const int cell = blockIdx.x * blockDim.x + tid;
  if (cell >= H * W) return;
  const int r = cell / W, c = cell % W;
  int kind = 0;
  for (int k = 0; k < s_n; ++k)
    if (r >= s_r0[k] && r < s_r1[k] && c >= s_c0[k] && c < s_c1[k]) kind = s_kind[k];
  if (kind == 0) {
    uint32_t a, b;
    hrand(seed, fill_stream, ge, (uint32_t)cell, a, b);
    kind = 1 + (int)(a % 3u);
  }
  out[(size_t)e * H * W + cell] = kind;
}

// altitude: U(0,5) noise + 6..9 cosine hills + 4..7 linear ramps, / 10 (float64, like the reference's NumPy)
__global__ void __launch_bounds__(256) hidden_altitude_kernel(int H, int W, unsigned long long seed, int env_offset,
                                                              double* alt64, float* alt32) {
  __shared__ int s_nh, s_nr, s_hr[9], s_hc[9], s_rad[9], s_sr[8], s_sc[8], s_w[8], s_h[8];
  __shared__ double s_hh[9], s_diff[8];
  const int e = blockIdx.y;
  const uint32_t ge = (uint32_t)(env_offset + e);
  const int tid = threadIdx.x;
  if (tid < 9) {
    uint32_t a, b, c, d, x, y;
    hrand(seed, HS_ALT_HILL, ge, 3u * tid, a, b);
    hrand(seed, HS_ALT_HILL, ge, 3u * tid + 1u, c, d);
    hrand(seed, HS_ALT_HILL, ge, 3u * tid + 2u, x, y);
    s_hr[tid] = (int)(a % (uint32_t)H);
    s_hc[tid] = (int)(b % (uint32_t)W);
    s_rad[tid] = 2 + (int)(c % (uint32_t)(min(H, W) / 4 - 2));
    s_hh[tid] = 2.0 + 4.0 * hunif(x, y);
  } else if (tid >= 16 && tid < 23) {
    const int k = tid - 16;
    uint32_t a, b, c, d, x, y;
    hrand(seed, HS_ALT_RAMP, ge, 3u * k, a, b);
    hrand(seed, HS_ALT_RAMP, ge, 3u * k + 1u, c, d);
    hrand(seed, HS_ALT_RAMP, ge, 3u * k + 2u, x, y);
    s_sr[k] = (int)(a % (uint32_t)(H - 4));
    s_sc[k] = (int)(b % (uint32_t)(W - 4));
    s_w[k] = 3 + (int)(c % (uint32_t)(W / 4 - 3));
    s_h[k] = 3 + (int)(d % (uint32_t)(H / 4 - 3));
    s_diff[k] = 1.0 + 3.0 * hunif(x, y);
  } else if (tid == 32) {
    uint32_t a, b;
    hrand(seed, HS_ALT_HILL, ge, 100u, a, b);
    s_nh = 6 + (int)(a % 4u);
    hrand(seed, HS_ALT_RAMP, ge, 100u, a, b);
    s_nr = 4 + (int)(a % 4u);
  }
  __syncthreads();
  const int cell = blockIdx.x * blockDim.x + tid;
  if (cell >= H * W) return;
  const int r = cell / W, c = cell % W;
  uint32_t a, b;
  hrand(seed, HS_ALT_NOISE, ge, (uint32_t)cell, a, b);
  double alt = 5.0 * hunif(a, b);
  for (int k = 0; k < s_nh; ++k) {
    const double dr = (double)(r - s_hr[k]), dc = (double)(c - s_hc[k]);
    const double dist = sqrt(dr * dr + dc * dc);
    if (dist < (double)s_rad[k]) alt += s_hh[k] * cos(dist / (double)s_rad[k] * 3.141592653589793 / 2.0);
  }
  for (int k = 0; k < s_nr; ++k) {
    const int r1 = min(s_sr[k] + s_h[k], H), c1 = min(s_sc[k] + s_w[k], W);
    if (r >= s_sr[k] && r < r1 && c >= s_sc[k] && c < c1) alt += s_diff[k] * ((double)(r - s_sr[k]) / (double)s_h[k]);
  }
  alt /= 10.0;
  const size_t i = (size_t)e * H * W + cell;
  alt64[i] = alt;
  alt32[i] = (float)alt;
}

// slope (degrees) to the 8 neighbours and its factor: interior cells only, border cells are flat
__global__ void __launch_bounds__(256) hidden_slope_kernel(int H, int W, const double* __restrict__ alt64,
                                                           float* __restrict__ slope9, float* __restrict__ pslope9) {
  const int e = blockIdx.y;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= H * W) return;
  const int r = cell / W, c = cell % W;
  const double* a = alt64 + (size_t)e * H * W;
  const size_t o = ((size_t)e * H * W + cell) * 9;
  const bool interior = r > 0 && r < H - 1 && c > 0 && c < W - 1;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float s = 0.0f;
      if (interior && !(i == 1 && j == 1)) {
        double d = a[r * W + c] - a[(r - 1 + i) * W + (c - 1 + j)];
        if (i != 1 && j != 1) d = d / 1.414;
        s = (float)(atan(d) * (180.0 / 3.141592653589793));
      }
      if (slope9) slope9[o + i * 3 + j] = s;
      // exp(f32(0.078) * slope) with the float32 product of the reference, the exponential in double
      pslope9[o + i * 3 + j] = (float)exp((double)__fmul_rn(0.078f, s));
    }
}

}  // namespace

cudaError_t launch_generate_hidden(int N, int H, int W, unsigned long long seed, int env_offset, int32_t* veg, int32_t* den,
                                   float* alt32, double* alt64, float* slope9, float* pslope9, cudaStream_t st) {
  const dim3 grid((H * W + 255) / 256, N);
  hidden_patches_kernel<<<grid, 256, 0, st>>>(H, W, seed, env_offset, HS_DEN_RECT, HS_DEN_FILL, den);
  hidden_patches_kernel<<<grid, 256, 0, st>>>(H, W, seed, env_offset, HS_VEG_RECT, HS_VEG_FILL, veg);
  hidden_altitude_kernel<<<grid, 256, 0, st>>>(H, W, seed, env_offset, alt64, alt32);
  hidden_slope_kernel<<<grid, 256, 0, st>>>(H, W, alt64, slope9, pslope9);
  return cudaGetLastError();
}

}  // namespace gca
