// gca_tiled.cu -- environment step for grids of any size (256x256 x 1024 envs, one 4096x4096
// grid, ...): 2-D tiles with a halo of R cells, ONE CA sub-step per launch.
//
// Same rule, same lazy counter-based draws and the same enclosure / exact-fallback logic as the
// 64x64 kernel (gca_step64.cu); what differs is the decomposition:
//   * tile of 32 x 64 cells per CTA; tile + halo (R <= 10) of the u8 grid is staged into shared
//     memory -- by TMA (cp.async.bulk.tensor, 3-D map (W, H, N); out-of-bounds coordinates are
//     zero-filled, which IS the reference's jnp.pad(constant_values=0) boundary,
//     ca_alexandridis_jax.py:26) when W % 16 == 0, by plain bounds-checked loads otherwise;
//   * front cells of the tile are compacted into a shared list (ballot + one atomic per warp) and
//     processed balanced over the CTA: window sum -> enclosure -> per burning direction one
//     threefry block addressed by the GLOBAL linear index ((r W + c) 9 + d), so results do not
//     depend on the tiling;
//   * only tiles that can change are worked on: a dense, vectorised pass (tile_flags_kernel, 1 byte
//     per cell read) marks the tiles that hold fire, and a tile is ACTIVE when its 3x3 tile
//     neighbourhood holds any (R <= 10 < the tile's extent, so nothing further away can ignite it);
//     the CTAs of the other tiles leave at once -- for a small fire on a 4096^2 grid that is all
//     but a handful.  Active tiles write their new cells to the scratch grid (neighbouring tiles
//     still read the old ones) and tile_apply_kernel copies them back, so S.cell is always the
//     current grid; burn-out ticks / ages are updated in place (own cell only);
//   * per-env scalars (key chain, wind walk, clock, move, douse, reward, done) live in two tiny
//     kernels around the K sub-step launches.
// Reference lines as in gca_step64.cu.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>
#include <cstring>

#include "gca_common.cuh"

namespace gca {

constexpr int T_TH = 32, T_TW = 64, T_THREADS = 256;
constexpr int T_MAXR = GCA_MAX_R;
constexpr int T_HC = 16;                              // column halo: TMA needs the box start 16-byte aligned,
                                                      // so the left halo is always 16 columns (>= R)
constexpr int T_PITCH_MAX = T_TW + 2 * T_HC;          // 96
constexpr int T_ROWS_MAX = T_TH + 2 * T_MAXR;         // 52
#define T_LO 0.9998779296875f  /* 1 - 2^-13: (2R+1)^2 <= 441 terms -> |err| <= 441 u |sum| */
#define T_HI 1.0001220703125f  /* 1 + 2^-13 */

// sched[e][*] written by tiled_sched_kernel for the current sub-step
enum { SC_BURN0 = 0, SC_BURN1, SC_GROW0, SC_GROW1, SC_AK10, SC_AK11, SC_AK20, SC_AK21, SC_WIND, SC_N = 12 };

// ---------------------------------------------------------------------------------------------
// per-env key schedule of ONE PartiallyObservableForestFireJax.update (thread per env)
// ---------------------------------------------------------------------------------------------
__global__ void tiled_sched_kernel(gca_params P, gca_state S, gca_inject J, int substep, uint32_t* sched) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S.N) return;
  const int mode = P.rng_mode;
  uint32_t k0 = S.key[2 * e], k1 = S.key[2 * e + 1];
  uint32_t K1[2], S1[2], Ka[2], Sb[2], Kb[2], Sg[2], Kc[2], Sa[2], K2[2], Sw[2], K3[2], Si[2], a1[2], a2[2], w1[2],
      w2[2];
  split_thread(k0, k1, mode, K1[0], K1[1], S1[0], S1[1]);
  split_thread(S1[0], S1[1], mode, Ka[0], Ka[1], Sb[0], Sb[1]);
  split_thread(Ka[0], Ka[1], mode, Kb[0], Kb[1], Sg[0], Sg[1]);
  split_thread(Kb[0], Kb[1], mode, Kc[0], Kc[1], Sa[0], Sa[1]);
  split_thread(Sa[0], Sa[1], mode, a1[0], a1[1], a2[0], a2[1]);
  split_thread(K1[0], K1[1], mode, K2[0], K2[1], Sw[0], Sw[1]);
  split_thread(K2[0], K2[1], mode, K3[0], K3[1], Si[0], Si[1]);
  split_thread(Si[0], Si[1], mode, w1[0], w1[1], w2[0], w2[1]);
  float u = bits_to_uniform(bits_scalar(tf_key(Sw[0], Sw[1]), mode));
  int step = randint_from_bits(bits_scalar(tf_key(w1[0], w1[1]), mode), bits_scalar(tf_key(w2[0], w2[1]), mode), 1,
                               7u, 4u);
  if (J.u_wind) u = J.u_wind[(size_t)substep * S.N + e];
  if (J.wind_step) step = J.wind_step[(size_t)substep * S.N + e];
  uint32_t* sc = sched + (size_t)e * SC_N;
  sc[SC_BURN0] = Sb[0]; sc[SC_BURN1] = Sb[1]; sc[SC_GROW0] = Sg[0]; sc[SC_GROW1] = Sg[1];
  sc[SC_AK10] = a1[0]; sc[SC_AK11] = a1[1]; sc[SC_AK20] = a2[0]; sc[SC_AK21] = a2[1];
  const int w = S.wind_index[e];
  sc[SC_WIND] = (uint32_t)w;  // wind used by THIS sub-step
  S.wind_index[e] = (u < P.p_wind_change) ? (w + step) % 8 : w;
  S.key[2 * e] = K3[0];
  S.key[2 * e + 1] = K3[1];
}

// ---------------------------------------------------------------------------------------------
// tile activity: flags[e][ty][tx] = the tile holds a burning cell; with want_counts the tree / fire
// cells of the whole grid are counted on the way (the step's reward / done need them once)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tile_flags_kernel(int H, int W, const uint8_t* __restrict__ cell,
                                                         uint8_t* __restrict__ tile_flags, int32_t* __restrict__ counts,
                                                         int want_counts, int all_active) {
  const int e = blockIdx.z, r0 = blockIdx.y * T_TH, c0 = blockIdx.x * T_TW;
  const int tid = threadIdx.x;
  const int r = r0 + (tid >> 2), c = c0 + (tid & 3) * 16;  // 16 cells per thread
  const uint8_t* src = cell + ((size_t)e * H + r) * W + c;
  uint32_t fire = 0;
  int nt = 0, nf = 0;
  if (r < H && c < W) {
    if ((W & 15) == 0) {
      const uint4 v = *reinterpret_cast<const uint4*>(src);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // cell codes 0 / 1 / 2: bit 0 = tree, bit 1 = fire
        fire |= w[k] & 0x02020202u;
        nt += __popc(w[k] & 0x01010101u);
        nf += __popc(w[k] & 0x02020202u);
      }
    } else {
      for (int k = 0; k < 16 && c + k < W; ++k) {
        const int v = src[k];
        fire |= v == 2;
        nt += v == 1;
        nf += v == 2;
      }
    }
  }
  const int any = __syncthreads_or(fire != 0u) | all_active;
  if (tid == 0) tile_flags[((size_t)e * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = (uint8_t)(any != 0);
  if (want_counts) {
    nt = __reduce_add_sync(GCA_FULL, nt);
    nf = __reduce_add_sync(GCA_FULL, nf);
    if ((tid & 31) == 0) {
      if (nt) atomicAdd(&counts[2 * e], nt);
      if (nf) atomicAdd(&counts[2 * e + 1], nf);
    }
  }
}

// does the 3x3 tile neighbourhood of this CTA's tile hold fire?  (uniform over the CTA)
__device__ __forceinline__ bool tile_active(const uint8_t* __restrict__ tile_flags) {
  const int tx = blockIdx.x, ty = blockIdx.y, TX = gridDim.x, TY = gridDim.y;
  const uint8_t* f = tile_flags + (size_t)blockIdx.z * TY * TX;
  bool a = false;
  if (threadIdx.x < 9) {
    const int y = ty + (int)threadIdx.x / 3 - 1, x = tx + (int)threadIdx.x % 3 - 1;
    if (y >= 0 && y < TY && x >= 0 && x < TX) a = f[y * TX + x] != 0;
  }
  return __syncthreads_or(a) != 0;
}

// copies the new cells of the active tiles from the scratch grid back into the grid
__global__ void __launch_bounds__(128) tile_apply_kernel(int H, int W, const uint8_t* __restrict__ tile_flags,
                                                         const uint8_t* __restrict__ scratch, uint8_t* __restrict__ cell) {
  if (!tile_active(tile_flags)) return;
  const int e = blockIdx.z, r0 = blockIdx.y * T_TH, c0 = blockIdx.x * T_TW;
  const int tid = threadIdx.x;
  const int r = r0 + (tid >> 2), c = c0 + (tid & 3) * 16;
  if (r >= H || c >= W) return;
  const size_t off = ((size_t)e * H + r) * W + c;
  if ((W & 15) == 0) {
    *reinterpret_cast<uint4*>(cell + off) = *reinterpret_cast<const uint4*>(scratch + off);
  } else {
    for (int k = 0; k < 16 && c + k < W; ++k) cell[off + k] = scratch[off + k];
  }
}

// ---------------------------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------------------------
struct TileSmem {
  alignas(128) uint8_t tile[T_ROWS_MAX * T_PITCH_MAX];  // cells incl. halo, pitch = params
  float wmat[(2 * T_MAXR + 1) * (2 * T_MAXR + 1)];      // burn kernel, row-major
  uint16_t list[T_TH * T_TW];                            // front cells of the tile: (lr << 6) | lc
  uint8_t ignite[T_TH * T_TW];                           // 1 = ignites this sub-step
  alignas(8) unsigned long long mbar;
  int nfront;
  int cnt_tree, cnt_fire;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 5x5 dousing window of (r, c) from the global bit-board: bit (5 i + j) <-> (r-2+i, c-2+j)
__device__ __forceinline__ uint32_t dous_window_g(const unsigned long long* __restrict__ db, int H, int W, int WW,
                                                  int r, int c) {
  uint32_t v = 0;
  for (int i = 0; i < 5; ++i) {
    const int rr = r - 2 + i;
    if (rr < 0 || rr >= H) continue;
    const unsigned long long* row = db + (size_t)rr * WW;
    for (int j = 0; j < 5; ++j) {
      const int cc = c - 2 + j;
      if (cc < 0 || cc >= W) continue;
      v |= (uint32_t)((row[cc >> 6] >> (cc & 63)) & 1ull) << (5 * i + j);
    }
  }
  return v;
}

template <bool USE_TMA>
__global__ void __launch_bounds__(T_THREADS)
ca_tiled_kernel(const __grid_constant__ gca_params P, const __grid_constant__ gca_state S,
                const __grid_constant__ gca_inject J, const __grid_constant__ CUtensorMap tmap,
                const uint8_t* __restrict__ cell_in, uint8_t* __restrict__ cell_out,
                const uint32_t* __restrict__ sched, int32_t* __restrict__ counts, unsigned long long* stats,
                const uint8_t* __restrict__ tile_flags, int substep, int pitch) {
  __shared__ TileSmem sm;
  if (!tile_active(tile_flags)) return;  // no fire within reach: nothing in this tile can change
  const int H = P.H, W = P.W, R = P.R, mode = P.rng_mode;
  const int WW = (W + 63) >> 6;
  const int e = blockIdx.z;
  const int r0 = blockIdx.y * T_TH, c0 = blockIdx.x * T_TW;
  const int tid = threadIdx.x, lane = tid & 31;
  const int rows = T_TH + 2 * R;
  const size_t env_off = (size_t)e * H * W;
  const int win = 2 * R + 1;

  if (tid == 0) { sm.nfront = 0; sm.cnt_tree = 0; sm.cnt_fire = 0; }
  // burn kernel weights: ring k = max(|di|, |dj|); centre shares ring 1's weight
  for (int i = tid; i < win * win; i += T_THREADS) {
    const int di = abs(i / win - R), dj = abs(i % win - R);
    sm.wmat[i] = P.ring_w[max(di, dj)];
  }
  // ---- stage tile + halo -------------------------------------------------------------------------
  if (USE_TMA) {
    if (tid == 0) {
      const uint32_t bar = smem_u32(&sm.mbar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const uint32_t bytes = (uint32_t)(rows * pitch);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      // box (pitch, rows, 1) at (c0 - 16, r0 - R, e); negative / beyond-edge coordinates are zero-filled.
      // The innermost start coordinate must keep the global address 16-byte aligned.
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"(smem_u32(sm.tile)), "l"(&tmap), "r"(c0 - T_HC), "r"(r0 - R), "r"(e), "r"(bar)
          : "memory");
    }
    __syncthreads();
    {
      const uint32_t bar = smem_u32(&sm.mbar);
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(0u)
            : "memory");
      }
    }
  } else {
    for (int i = tid; i < rows * pitch; i += T_THREADS) {
      const int lr = i / pitch, lc = i % pitch;
      const int gr = r0 - R + lr, gc = c0 - T_HC + lc;
      uint8_t v = 0;
      if (gr >= 0 && gr < H && gc >= 0 && gc < W) v = cell_in[env_off + (size_t)gr * W + gc];
      sm.tile[i] = v;
    }
    __syncthreads();
  }

  const uint32_t* sc = sched + (size_t)e * SC_N;
  const uint32_t tick = S.tick[e] + (uint32_t)substep;  // S.tick advances by K in the epilogue
  const uint32_t half_cell = (uint32_t)(((size_t)H * W) >> 1);
  const uint32_t half_burn = (uint32_t)((9ull * H * W) >> 1);

  // ---- find the tile's front cells ---------------------------------------------------------------
  const int lc = tid & 63, rg = tid >> 6;
  for (int k = 0; k < T_TH / 4; ++k) {
    const int lr = rg + 4 * k;
    const int gr = r0 + lr, gc = c0 + lc;
    const uint8_t* ctr = sm.tile + (lr + R) * pitch + (lc + T_HC);
    bool front = false;
    if (gr < H && gc < W && ctr[0] == 1) {
      front = ctr[-pitch - 1] == 2 || ctr[-pitch] == 2 || ctr[-pitch + 1] == 2 || ctr[-1] == 2 || ctr[1] == 2 ||
              ctr[pitch - 1] == 2 || ctr[pitch] == 2 || ctr[pitch + 1] == 2;
    }
    sm.ignite[lr * T_TW + lc] = 0;
    const uint32_t bal = __ballot_sync(GCA_FULL, front);
    if (bal) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&sm.nfront, __popc(bal));
      base = __shfl_sync(GCA_FULL, base, 0);
      if (front) sm.list[base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)((lr << 6) | lc);
    }
  }
  __syncthreads();

  // ---- balanced pass over the front cells --------------------------------------------------------
  const int nfront = sm.nfront;
  const TfKey kburn = tf_key(sc[SC_BURN0], sc[SC_BURN1]);
  const float* wind = P.winds + 9 * (int)sc[SC_WIND];
  uint32_t n_draws = 0, n_thresh = 0;
  for (int i = tid; i < nfront; i += T_THREADS) {
    const int lr = sm.list[i] >> 6, lcc = sm.list[i] & 63;
    const int gr = r0 + lr, gc = c0 + lcc;
    const size_t gcell = (size_t)gr * W + gc;
    const uint8_t* ctr = sm.tile + (lr + R) * pitch + (lcc + T_HC);
    // heat: any summation order is inside the enclosure
    float Hf = 0.0f;
    for (int di = 0; di < win; ++di) {
      const uint8_t* rowp = ctr + (di - R) * pitch - R;
      const float* wrow = sm.wmat + di * win;
      for (int dj = 0; dj < win; ++dj) Hf += (rowp[dj] == 2) ? wrow[dj] : 0.0f;
    }
    float Dlo = 0.0f, Dhi = 0.0f;
    uint32_t dwin = dous_window_g(reinterpret_cast<const unsigned long long*>(S.doused) + (size_t)e * H * WW, H, W,
                                  WW, gr, gc);
    if (dwin) {
      const int ni = __popc(dwin & ((0x0Eu << 5) | (0x0Eu << 10) | (0x0Eu << 15)));
      const int nb = __popc(dwin) - ni;
      const float Df = fmaf((float)nb, P.dous_border, (float)ni * P.dous_inner);
      Dlo = __fmul_rn(Df, T_LO);
      Dhi = __fmul_rn(Df, T_HI);
    }
    int hid = 3 | (3 << 3);
    if (S.hidden != nullptr) hid = S.hidden[env_off + gcell];
    const float a = P.onep_veg[clip15(hid & 7)], b = P.onep_den[clip15((hid >> 3) & 7)];
    const float blo = __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(Hf, T_LO), Dhi), a), b);
    const float bhi = __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(Hf, T_HI), Dlo), a), b);
    if (!(bhi > 0.0f)) continue;
    bool ig = false;
    float base_exact = 0.0f;
    bool have_exact = false;
    for (int d = 0; d < 9 && !ig; ++d) {
      if (d == 4) continue;
      if (ctr[(d / 3 - 1) * pitch + (d % 3 - 1)] != 2) continue;
      float u;
      if (J.u_burn) u = J.u_burn[(((size_t)substep * S.N + e) * H * W + gcell) * 9 + d];
      else u = bits_to_uniform(bits_at(kburn, (uint32_t)(gcell * 9 + d), half_burn, mode));
      ++n_draws;
      const float w = wind[d];
      const float s = S.pslope ? S.pslope[(env_off + gcell) * 8 + dir_slot(d)] : 1.0f;
      const float plo = __fmul_rn(__fmul_rn(blo, w), s), phi = __fmul_rn(__fmul_rn(bhi, w), s);
      if (u < plo) { ig = true; break; }
      if (u < phi) {
        // threshold cell: reference-order (row-major, float32, from +0) sums
        if (!have_exact) {
          float heat = 0.0f;
          for (int di = 0; di < win; ++di)
            for (int dj = 0; dj < win; ++dj)
              if (ctr[(di - R) * pitch + (dj - R)] == 2) heat = __fadd_rn(heat, sm.wmat[di * win + dj]);
          float dous = 0.0f;
          for (int q = 0; q < 25; ++q)
            if ((dwin >> q) & 1u) {
              const int qi = q / 5, qj = q % 5;
              const bool inner = qi >= 1 && qi <= 3 && qj >= 1 && qj <= 3;
              dous = __fadd_rn(dous, inner ? P.dous_inner : P.dous_border);
            }
          base_exact = __fmul_rn(__fmul_rn(__fsub_rn(heat, dous), a), b);
          have_exact = true;
          ++n_thresh;
        }
        if (u < __fmul_rn(__fmul_rn(base_exact, w), s)) ig = true;
      }
    }
    if (ig) sm.ignite[lr * T_TW + lcc] = 1;
  }
  __syncthreads();

  // ---- write the new grid, burn-out ticks, counts -------------------------------------------------
  const TfKey ka1 = tf_key(sc[SC_AK10], sc[SC_AK11]), ka2 = tf_key(sc[SC_AK20], sc[SC_AK21]);
  const TfKey kg = tf_key(sc[SC_GROW0], sc[SC_GROW1]);
  int nt = 0, nfire = 0;
  uint32_t n_ign = 0, n_ext = 0;
  for (int k = 0; k < T_TH / 4; ++k) {
    const int lr = rg + 4 * k;
    const int gr = r0 + lr, gc = c0 + lc;
    if (gr >= H || gc >= W) continue;
    const size_t gcell = (size_t)gr * W + gc;
    const int old = sm.tile[(lr + R) * pitch + (lc + T_HC)];
    int nw = old;
    if (old == 1 && sm.ignite[lr * T_TW + lc]) {
      nw = 2;
      int age;
      if (J.age_new) age = J.age_new[((size_t)substep * S.N + e) * H * W + gcell];
      else age = randint_from_bits(bits_at(ka1, (uint32_t)gcell, half_cell, mode),
                                   bits_at(ka2, (uint32_t)gcell, half_cell, mode), P.age_lo, P.age_span, P.age_mult);
      S.death[env_off + gcell] = (uint16_t)(tick + (uint32_t)age);  // burns out at tick + age
      ++n_ign;
    } else if (old == 0) {
      if (P.p_tree > 0.0f) {
        float u;
        if (J.u_grow) u = J.u_grow[((size_t)substep * S.N + e) * H * W + gcell];
        else u = bits_to_uniform(bits_at(kg, (uint32_t)gcell, half_cell, mode));
        if (u < P.p_tree) nw = 1;
      }
    } else if (old == 2) {
      if (((S.death[env_off + gcell] - tick) & 0xFFFFu) == 0u) {  // fire_age <= 1 -> empty, age ends at 0
        nw = 0;
        S.death[env_off + gcell] = 0;
        ++n_ext;
      }
    }
    cell_out[env_off + gcell] = (uint8_t)nw;
    nt += (nw == 1) - (old == 1);     // change of the env's tree / fire counts (tile_flags_kernel counted
    nfire += (nw == 2) - (old == 2);  // the grid as it was before this sub-step)
  }
  nt = __reduce_add_sync(GCA_FULL, nt);
  nfire = __reduce_add_sync(GCA_FULL, nfire);
  if (lane == 0) {
    if (nt) atomicAdd(&sm.cnt_tree, nt);
    if (nfire) atomicAdd(&sm.cnt_fire, nfire);
  }
  if (stats != nullptr) {
    const uint32_t a = __reduce_add_sync(GCA_FULL, n_draws), b = __reduce_add_sync(GCA_FULL, n_ign);
    const uint32_t c = __reduce_add_sync(GCA_FULL, n_ext), d = __reduce_add_sync(GCA_FULL, n_thresh);
    if (lane == 0) {
      if (a) atomicAdd(&stats[1], (unsigned long long)a);
      if (b) atomicAdd(&stats[2], (unsigned long long)b);
      if (c) atomicAdd(&stats[3], (unsigned long long)c);
      if (d) atomicAdd(&stats[4], (unsigned long long)d);
    }
  }
  __syncthreads();
  if (tid == 0) {
    if (sm.cnt_tree) atomicAdd(&counts[2 * e], sm.cnt_tree);
    if (sm.cnt_fire) atomicAdd(&counts[2 * e + 1], sm.cnt_fire);
    if (stats != nullptr && nfront) atomicAdd(&stats[0], (unsigned long long)nfront);
  }
}

// ---------------------------------------------------------------------------------------------
// per-env epilogue of the env step (thread per env): tick, clock, move, douse, day/night, reward,
// done, info counters (MDP.update tail + stateless_step, advanced_bulldozer.py:1112-1127,378-391)
// ---------------------------------------------------------------------------------------------
__global__ void tiled_finish_kernel(gca_params P, gca_state S, const int32_t* __restrict__ actions, gca_step_out O,
                                    const int32_t* __restrict__ counts, uint32_t flags) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S.N) return;
  const int WW = (P.W + 63) >> 6;
  S.tick[e] += (uint32_t)P.K;
  const int t = counts[2 * e], f = counts[2 * e + 1];
  const float rew = award(t, f);
  const bool done = f == 0;
  if (!(flags & GCA_FLAG_CA_ONLY)) {
    const int a0 = actions[3 * e], a1 = actions[3 * e + 1];
    const int a0c = min(max(a0, 0), 8), a1c = min(max(a1, 0), 1);
    const float tt = __fadd_rn(__fadd_rn(P.t_move[a0c], P.t_shoot[a1c]), P.t_any);
    const float ntm = __fadd_rn(S.time[e], tt);
    S.time[e] = __fsub_rn(ntm, truncf(ntm));
    int row = S.position[2 * e], col = S.position[2 * e + 1];
    move_position(a0, P.H, P.W, row, col);
    S.position[2 * e] = row;
    S.position[2 * e + 1] = col;
    if (a1 == 1) S.doused[((size_t)e * P.H + row) * WW + (col >> 6)] |= 1ull << (col & 63);
    const int ts = S.time_step[e] + 1;
    S.time_step[e] = ts;
    int night = S.is_night[e];
    if (O.obs_night) O.obs_night[e] = (uint8_t)night;
    if (ts % P.day_length == 0) night = 1 - night;
    S.is_night[e] = night;
    if (S.steps_elapsed) S.steps_elapsed[e] = __fadd_rn(S.steps_elapsed[e], 1.0f);
    if (S.reward_accumulated) S.reward_accumulated[e] = __fadd_rn(S.reward_accumulated[e], rew);
  }
  if (O.step_reward) O.step_reward[e] = rew;
  if (O.reward) O.reward[e] = rew;
  if (O.terminated) O.terminated[e] = done ? 1 : 0;
  if (O.counts) { O.counts[2 * e] = t; O.counts[2 * e + 1] = f; }
  if (O.stats != nullptr) atomicAdd(&O.stats[5], 1ull);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

static bool make_tmap(CUtensorMap* m, const uint8_t* base, int N, int H, int W, int pitch, int rows) {
  auto enc = get_encode();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (W & 15)) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)W * (cuuint64_t)H};
  const cuuint32_t box[3] = {(cuuint32_t)pitch, (cuuint32_t)rows, 1u};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Runs K sub-steps and the epilogue.  Per sub-step: key schedule, tile activity flags (+ cell counts in
// the last one), the CA kernel on the active tiles (S.cell -> scratch_cell), the copy back.
// scratch_cell: [N][H][W] u8; scratch_sched: [N][12] u32; scratch_counts: [N][2] i32;
// tile_flags: [N][ceil(H/32)][ceil(W/64)] u8.
static cudaError_t enqueue_tiled_env_step(const gca_params& p, const gca_state& s, const int32_t* actions,
                                          const gca_step_out& out, const gca_inject& inj, uint32_t flags,
                                          uint8_t* scratch_cell, uint32_t* scratch_sched, int32_t* scratch_counts,
                                          uint8_t* tile_flags, int use_tma, cudaStream_t st) {
  const int N = s.N, H = p.H, W = p.W, R = p.R;
  const int pitch = T_PITCH_MAX;
  const int rows = T_TH + 2 * R;
  const int TX = (W + T_TW - 1) / T_TW, TY = (H + T_TH - 1) / T_TH;
  cudaError_t err;
  // regrowth (p_tree > 0) can change any empty cell: every tile is active then
  const int all_active = p.p_tree > 0.0f ? 1 : 0;
  for (int j = 0; j < p.K; ++j) {
    const int last = j == p.K - 1;
    tiled_sched_kernel<<<(N + 127) / 128, 128, 0, st>>>(p, s, inj, j, scratch_sched);
    if (last && (err = cudaMemsetAsync(scratch_counts, 0, sizeof(int32_t) * 2 * N, st)) != cudaSuccess) return err;
    // gridDim.z holds the env: at most 65535 envs per launch
    for (int e0 = 0; e0 < N; e0 += 65535) {
      const int ne = min(65535, N - e0);
      dim3 grid(TX, TY, ne);
      gca_state sv = s;  // the slice's view of the per-env arrays the tile kernels index by blockIdx.z
      sv.N = ne;
      const size_t cells = (size_t)e0 * H * W;
      sv.cell = s.cell + cells;
      sv.death = s.death + cells;
      sv.hidden = s.hidden ? s.hidden + cells : nullptr;
      sv.pslope = s.pslope ? s.pslope + cells * 8 : nullptr;
      sv.doused = s.doused + (size_t)e0 * H * ((W + 63) >> 6);
      sv.tick = s.tick + e0;
      gca_inject jv = inj;  // injected fields are indexed [substep][N][...]: only whole-batch launches may use them
      if (e0 > 0 || ne < N) {
        if (inj.u_burn || inj.u_grow || inj.age_new) return cudaErrorInvalidValue;
      }
      CUtensorMap tm;
      memset(&tm, 0, sizeof(tm));
      const bool tma = use_tma && make_tmap(&tm, sv.cell, ne, H, W, pitch, rows);
      uint8_t* tf = tile_flags + (size_t)e0 * TX * TY;
      tile_flags_kernel<<<grid, 128, 0, st>>>(H, W, sv.cell, tf, scratch_counts + 2 * (size_t)e0, last, all_active);
      if (tma)
        ca_tiled_kernel<true><<<grid, T_THREADS, 0, st>>>(p, sv, jv, tm, sv.cell, scratch_cell + cells,
                                                          scratch_sched + (size_t)e0 * SC_N, scratch_counts + 2 * (size_t)e0,
                                                          out.stats, tf, j, pitch);
      else
        ca_tiled_kernel<false><<<grid, T_THREADS, 0, st>>>(p, sv, jv, tm, sv.cell, scratch_cell + cells,
                                                           scratch_sched + (size_t)e0 * SC_N, scratch_counts + 2 * (size_t)e0,
                                                           out.stats, tf, j, pitch);
      tile_apply_kernel<<<grid, 128, 0, st>>>(H, W, tf, scratch_cell + cells, sv.cell);
    }
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
  }
  tiled_finish_kernel<<<(N + 127) / 128, 128, 0, st>>>(p, s, actions, out, scratch_counts, flags);
  return cudaGetLastError();
}

// The 4 K + 1 launches of one env step as ONE CUDA graph launch: at 4096x4096 with a small fire nearly every tile
// exits at once and the step is bound by launch latency (85 us for K = 1).  The graph is captured once per set of
// buffers (thread-local cache of one entry; a capture stream of its own: the caller's may be the legacy stream, which
// cannot capture) and replayed with only the action pointer of the last kernel patched.  GCA_TILED_GRAPH=0 disables it.
namespace {
struct TiledGraphKey {
  gca_params p; gca_state s; gca_step_out out; uint32_t flags; const void *a, *b, *c, *d; int use_tma;
};
struct TiledGraphCache {
  bool valid = false, broken = false;
  TiledGraphKey key;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraphNode_t finish = nullptr;
  cudaStream_t cap = nullptr;
};
thread_local TiledGraphCache g_tiled_graph;
}  // namespace

cudaError_t launch_tiled_env_step(const gca_params& p, const gca_state& s, const int32_t* actions,
                                  const gca_step_out& out, const gca_inject& inj, uint32_t flags, uint8_t* scratch_cell,
                                  uint32_t* scratch_sched, int32_t* scratch_counts, uint8_t* tile_flags, int use_tma,
                                  cudaStream_t st) {
  static const bool enabled = [] { const char* v = getenv("GCA_TILED_GRAPH"); return !(v && v[0] == '0'); }();
  TiledGraphCache& G = g_tiled_graph;
  const bool injected = inj.u_burn || inj.u_grow || inj.age_new || inj.u_wind || inj.wind_step;
  if (!enabled || G.broken || injected)
    return enqueue_tiled_env_step(p, s, actions, out, inj, flags, scratch_cell, scratch_sched, scratch_counts, tile_flags,
                                  use_tma, st);
  TiledGraphKey key;
  memset(&key, 0, sizeof(key));
  key.p = p; key.s = s; key.out = out; key.flags = flags;
  key.out.done_token = 0; key.out.rgb = nullptr; key.out.rgb_u8 = 0;  // per-step values no tiled kernel reads
  key.out.host_done = nullptr; key.out.done_counter = nullptr;
  key.a = scratch_cell; key.b = scratch_sched; key.c = scratch_counts; key.d = tile_flags; key.use_tma = use_tma;
  auto fallback = [&]() {
    G.broken = true;
    cudaGetLastError();
    return enqueue_tiled_env_step(p, s, actions, out, inj, flags, scratch_cell, scratch_sched, scratch_counts, tile_flags,
                                  use_tma, st);
  };
  if (!G.valid || memcmp(&G.key, &key, sizeof(key)) != 0) {
    if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
    if (G.graph) { cudaGraphDestroy(G.graph); G.graph = nullptr; }
    G.valid = false;
    if (!G.cap && cudaStreamCreateWithFlags(&G.cap, cudaStreamNonBlocking) != cudaSuccess) return fallback();
    if (cudaStreamBeginCapture(G.cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return fallback();
    const cudaError_t e1 = enqueue_tiled_env_step(p, s, actions, out, inj, flags, scratch_cell, scratch_sched, scratch_counts,
                                                  tile_flags, use_tma, G.cap);
    const cudaError_t e2 = cudaStreamEndCapture(G.cap, &G.graph);
    if (e1 != cudaSuccess || e2 != cudaSuccess || !G.graph) return fallback();
    size_t nn = 0;
    if (cudaGraphGetNodes(G.graph, nullptr, &nn) != cudaSuccess || nn == 0) return fallback();
    cudaGraphNode_t* nodes = new cudaGraphNode_t[nn];
    cudaGraphGetNodes(G.graph, nodes, &nn);
    G.finish = nullptr;
    for (size_t i = 0; i < nn; ++i) {
      cudaGraphNodeType ty;
      if (cudaGraphNodeGetType(nodes[i], &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
      cudaKernelNodeParams kp;
      if (cudaGraphKernelNodeGetParams(nodes[i], &kp) == cudaSuccess && kp.func == (void*)tiled_finish_kernel) G.finish = nodes[i];
    }
    delete[] nodes;
    if (!G.finish || cudaGraphInstantiate(&G.exec, G.graph, 0) != cudaSuccess) return fallback();
    G.key = key;
    G.valid = true;
  }
  // patch the one argument that changes from step to step
  {
    gca_params pp = p; gca_state ss = s; gca_step_out oo = out;
    const int32_t* act = actions; const int32_t* cnt = scratch_counts; uint32_t fl = flags;
    void* args[6] = {&pp, &ss, &act, &oo, &cnt, &fl};
    cudaKernelNodeParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.func = (void*)tiled_finish_kernel;
    kp.gridDim = dim3((s.N + 127) / 128);
    kp.blockDim = dim3(128);
    kp.sharedMemBytes = 0;
    kp.kernelParams = args;
    if (cudaGraphExecKernelNodeSetParams(G.exec, G.finish, &kp) != cudaSuccess) { G.valid = false; return fallback(); }
  }
  if (cudaGraphLaunch(G.exec, st) != cudaSuccess) { G.valid = false; return fallback(); }
  return cudaSuccess;
}

}  // namespace gca
