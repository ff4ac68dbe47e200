// gca_windy.cu -- the v3 rule set (registered ForestFireBulldozer256x256-v3): WindyForestFire CA +
// Move / Modify (cut a tree) / RepeatCA clock + reward / done, batched, one CTA per env.
//
// Reference (paths relative to /root/reference/gym_cellular_automata/forest_fire/):
//   operators/ca_windy.py:41-139   one 3x3 uniform roll per CA update; a direction whose wind entry is
//                                  <= its roll fails; true convolution with the 8/2048 kernel; breaks
//                                  6144 / 6344 / 51200  ==>  fire -> empty, tree -> fire iff a burning
//                                  neighbour sits in a direction that did not fail, empty stays empty
//   operators/repeat_ca.py:32-45   clock in float64, int(repeats) CA updates per step
//   operators/move_modify.py:39-94 clamped 9-way move; shoot replaces a tree by empty
//   bulldozer/bulldozer.py:196-203 reward -(f / (t + f)), done = no fire
// Not a translation: the grid lives as two bit-boards (tree, fire); a CA update is 8 shifted ORs per
// 64-cell word in shared memory instead of an int64 convolution.  The reference draws its rolls
// from a freshly built, unseeded gymnasium Box, so rolls are an INPUT here (rule parity).
#include "gca_common.cuh"

namespace gca {

typedef unsigned long long u64;

// neighbour word of row r, word w shifted so that bit c holds column c + dc of that row
__device__ __forceinline__ u64 shifted_word(const u64* rows, int H, int WW, int r, int w, int dc) {
  if (r < 0 || r >= H) return 0ull;
  const u64* row = rows + (size_t)r * WW;
  const u64 x = row[w];
  if (dc == 0) return x;
  if (dc > 0) return (x >> 1) | (w + 1 < WW ? row[w + 1] << 63 : 0ull);
  return (x << 1) | (w > 0 ? row[w - 1] >> 63 : 0ull);
}

__global__ void __launch_bounds__(256)
windy_env_step_kernel(int N, int H, int W, u64* __restrict__ tree_g, u64* __restrict__ fire_g,
                      int32_t* __restrict__ position, double* __restrict__ time,
                      const int32_t* __restrict__ actions, const double* __restrict__ wind,
                      const double* __restrict__ rolls, int rmax, double t_move, double t_shoot, double t_any,
                      double* __restrict__ reward, uint8_t* __restrict__ terminated, int32_t* __restrict__ counts,
                      int32_t* __restrict__ repeats_out) {
  extern __shared__ u64 smem[];
  const int e = blockIdx.x;
  const int WW = (W + 63) >> 6, nw = H * WW;
  u64* tree = smem;
  u64* fa = smem + nw;
  u64* fb = smem + 2 * nw;
  __shared__ int s_rep, s_t, s_f;
  __shared__ uint32_t s_ok;
  const int tid = threadIdx.x;
  u64* tg = tree_g + (size_t)e * nw;
  u64* fg = fire_g + (size_t)e * nw;
  for (int i = tid; i < nw; i += blockDim.x) { tree[i] = tg[i]; fa[i] = fg[i]; }
  const int a0 = actions[2 * e], a1 = actions[2 * e + 1];
  if (tid == 0) {
    // RepeatCA.update: accu += time_action + time_state; accu, repeats = modf(accu)   (float64)
    const double ta = (a0 == 4 ? 0.0 : t_move) + (a1 == 0 ? 0.0 : t_shoot);
    double acc = time[e] + (ta + t_any);
    const double rep = trunc(acc);
    acc -= rep;
    time[e] = acc;
    s_rep = (int)rep;
    s_t = 0; s_f = 0;
  }
  __syncthreads();
  const int rep = min(s_rep, rmax);
  // last word of a row: columns >= W must stay clear
  const u64 tail = (W & 63) ? ((1ull << (W & 63)) - 1ull) : ~0ull;
  for (int k = 0; k < rep; ++k) {
    if (tid == 0) {
      uint32_t ok = 0;
      for (int q = 0; q < 9; ++q)
        if (q != 4 && !(wind[q] <= rolls[((size_t)e * rmax + k) * 9 + q])) ok |= 1u << q;  // not failed
      s_ok = ok;
    }
    __syncthreads();
    const uint32_t ok = s_ok;
    for (int i = tid; i < nw; i += blockDim.x) {
      const int r = i / WW, w = i % WW;
      u64 src = 0ull;
      // kernel element (ki, kj) multiplies the neighbour at (r + 1 - ki, c + 1 - kj) (true convolution)
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        if (q == 4) continue;
        if (ok & (1u << q)) src |= shifted_word(fa, H, WW, r + 1 - q / 3, w, 1 - q % 3);
      }
      u64 nf = tree[i] & src;
      if (w == WW - 1) nf &= tail;
      fb[i] = nf;          // tree with a burning neighbour in a live direction -> fire; old fire -> empty
      tree[i] &= ~nf;
    }
    __syncthreads();
    u64* t = fa; fa = fb; fb = t;
  }
  // Move + Modify
  if (tid == 0) {
    int row = position[2 * e], col = position[2 * e + 1];
    move_position(a0, H, W, row, col);
    position[2 * e] = row;
    position[2 * e + 1] = col;
    if (a1 != 0) tree[(size_t)row * WW + (col >> 6)] &= ~(1ull << (col & 63));  // cut: tree -> empty
    if (repeats_out) repeats_out[e] = s_rep;
  }
  __syncthreads();
  int nt = 0, nf = 0;
  for (int i = tid; i < nw; i += blockDim.x) {
    const u64 t = tree[i], f = fa[i];
    tg[i] = t;
    fg[i] = f;
    nt += __popcll(t);
    nf += __popcll(f);
  }
  nt = __reduce_add_sync(GCA_FULL, nt);
  nf = __reduce_add_sync(GCA_FULL, nf);
  if ((tid & 31) == 0) { atomicAdd(&s_t, nt); atomicAdd(&s_f, nf); }
  __syncthreads();
  if (tid == 0) {
    const int t = s_t, f = s_f;
    if (reward) reward[e] = (t + f) > 0 ? -((double)f / (double)(t + f)) : __longlong_as_double(0x7FF8000000000000ll);
    if (terminated) terminated[e] = f == 0;
    if (counts) { counts[2 * e] = t; counts[2 * e + 1] = f; }
  }
}

// u8 codes (0 empty, 1 tree, 2 fire) <-> the two bit-boards; one warp per 64-column word
__global__ void windy_pack_kernel(int N, int H, int W, const uint8_t* __restrict__ cell, u64* tree, u64* fire) {
  const int WW = (W + 63) >> 6;
  const long long word = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (word >= (long long)N * H * WW) return;
  const int lane = threadIdx.x & 31;
  const long long er = word / WW;
  const int w = (int)(word % WW);
  u64 t = 0, f = 0;
  for (int h = 0; h < 2; ++h) {
    const int c = w * 64 + h * 32 + lane;
    const int v = c < W ? cell[er * W + c] : 0;
    t |= (u64)__ballot_sync(GCA_FULL, v == 1) << (32 * h);
    f |= (u64)__ballot_sync(GCA_FULL, v == 2) << (32 * h);
  }
  if (lane == 0) { tree[word] = t; fire[word] = f; }
}
__global__ void windy_unpack_kernel(int N, int H, int W, const u64* __restrict__ tree, const u64* __restrict__ fire,
                                    uint8_t* cell) {
  const int WW = (W + 63) >> 6;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * H * W) return;
  const int c = (int)(i % W);
  const size_t er = i / W;
  const u64 t = tree[er * WW + (c >> 6)], f = fire[er * WW + (c >> 6)];
  cell[i] = ((f >> (c & 63)) & 1ull) ? 2 : (((t >> (c & 63)) & 1ull) ? 1 : 0);
}

cudaError_t launch_windy_step(int N, int H, int W, u64* tree, u64* fire, int32_t* position, double* time,
                              const int32_t* actions, const double* wind, const double* rolls, int rmax,
                              double t_move, double t_shoot, double t_any, double* reward, uint8_t* terminated,
                              int32_t* counts, int32_t* repeats_out, cudaStream_t st) {
  const int WW = (W + 63) >> 6;
  const size_t smem = (size_t)3 * H * WW * sizeof(u64);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t err = cudaFuncSetAttribute(windy_env_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    configured = smem;
  }
  windy_env_step_kernel<<<N, 256, smem, st>>>(N, H, W, tree, fire, position, time, actions, wind, rolls, rmax, t_move,
                                              t_shoot, t_any, reward, terminated, counts, repeats_out);
  return cudaGetLastError();
}
cudaError_t launch_windy_pack(int N, int H, int W, const uint8_t* cell, u64* tree, u64* fire, cudaStream_t st) {
  const long long words = (long long)N * H * ((W + 63) >> 6);
  windy_pack_kernel<<<(unsigned)((words + 7) / 8), 256, 0, st>>>(N, H, W, cell, tree, fire);
  return cudaGetLastError();
}
cudaError_t launch_windy_unpack(int N, int H, int W, const u64* tree, const u64* fire, uint8_t* cell, cudaStream_t st) {
  const size_t n = (size_t)N * H * W;
  windy_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(N, H, W, tree, fire, cell);
  return cudaGetLastError();
}

}  // namespace gca
