// gca_tiled.cu -- environment step for grids of any size (one 4096x4096 grid, odd widths, ...): 2-D tiles of
// 32 x 64 cells with a halo of R cells, one CA sub-step per tile-kernel launch, the whole env step one CUDA graph.
//
// Same rule, same lazy counter-based draws and the same enclosure / exact-fallback logic as the 64x64 kernel
// (gca_step64.cu); what differs is the decomposition:
//   * only tiles that can change are worked on.  A dense pass per ENV STEP (tile_count_kernel, a warp per tile; it also
//     counts the env's tree / fire cells for the reward) counts the burning cells of every tile; the step's tile LIST
//     holds the tiles within two tiles of one that holds fire -- fire moves one cell per sub-step, a tile is 32 x 64
//     cells, so nothing else can become active during the K <= 8 sub-steps -- (with regrowth: every tile), and every
//     sub-step is ONE launch of a small fixed grid whose CTAs loop over that list (one 4096x4096 grid has 8192 tiles
//     of which a young fire touches a few dozen);
//   * two grid buffers, S.cell and the scratch grid, are read and written alternately (sub-step j reads the tiles +
//     halos from one and writes the tiles' new cells to the other: neighbouring tiles still read the old cells), so no
//     copy-back pass separates the sub-steps; the ring of tiles around the list is copied into the scratch grid once
//     (it is only ever read as halo); after an odd number of sub-steps the listed tiles are copied back;
//   * tile + halo of the u8 grid is staged into shared memory -- by TMA (cp.async.bulk.tensor, 3-D map (W, H, N);
//     out-of-bounds coordinates are zero-filled, which IS the reference's jnp.pad(constant_values=0) boundary,
//     ca_alexandridis_jax.py:26) when W % 16 == 0, by plain bounds-checked loads otherwise -- and turned into tree /
//     fire bit rows (warp ballots): front cells are found bit-parallel (a thread per tile row), the (2R+1)^2 heat
//     window is nested box popcounts of those rows, the doused rows in reach are staged as 64-bit words;
//   * front cells in rounds of 256: a thread per cell computes the enclosure of the burn probability, the (cell,
//     burning direction) draws are then dealt one per thread -- one threefry block addressed by the GLOBAL linear
//     index ((r W + c) 9 + d), so results do not depend on the tiling;
//   * write phase: unless regrowth is on only burning cells (burn-out tick reached?) and igniting cells (age draw) are
//     looked at; the tile's interior leaves as 128-bit stores; burn-out ticks are updated in place (own cells only);
//   * the key schedules of all K sub-steps are walked by spare warps of the count kernel (a warp per env, lane pairs
//     per split), the per-env scalars (clock, move, douse, reward, done) come from the epilogue kernel.
// Reference lines as in gca_step64.cu.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>
#include <cstring>

#include "gca_common.cuh"

namespace gca {

cudaError_t launch_auto_reset(const gca_params&, const gca_state&, const gca_state&, const float*, float*, const uint8_t*,
                              cudaStream_t);  // gca_aux.cu

constexpr int T_TH = 32, T_TW = 64, T_THREADS = 256;
constexpr int T_MAXR = GCA_MAX_R;
constexpr int T_HC = 16;                              // column halo: TMA needs the box start 16-byte aligned,
                                                      // so the left halo is always 16 columns (>= R)
constexpr int T_PITCH_MAX = T_TW + 2 * T_HC;          // 96
constexpr int T_ROWS_MAX = T_TH + 2 * T_MAXR;         // 52
#define T_LO 0.9998779296875f  /* 1 - 2^-13: (2R+1)^2 <= 441 terms -> |err| <= 441 u |sum| */
#define T_HI 1.0001220703125f  /* 1 + 2^-13 */

// sched[e][j][*] of env e, sub-step j (written by tiled_sched_warp)
enum { SC_BURN0 = 0, SC_BURN1, SC_GROW0, SC_GROW1, SC_AK10, SC_AK11, SC_AK20, SC_AK21, SC_WIND, SC_N = 12 };

struct TileSmem {
  alignas(128) uint8_t tile[T_ROWS_MAX * T_PITCH_MAX];  // cells incl. halo, pitch = params
  uint16_t list[T_TH * T_TW];                            // front cells of the tile: (lr << 6) | lc
  unsigned long long ign[T_TH];                          // cells of the tile that ignite this sub-step (bit = column)
  uint32_t tbits[T_ROWS_MAX][4];                         // tree bit-board of tile + halo (see fbits)
  int nlist2;                                            // entries of the write phase's change-candidate list
  float2 bnd[T_THREADS];                                 // per front cell of the round: enclosure (lo, hi), direction-independent part
  uint16_t pairs[T_THREADS * 8];                         // draws of the round: (front cell of the round << 4) | direction
  int npairs[2];                                         // their number, double-buffered by round parity
  unsigned long long dw[T_TH + 4][3];                     // doused rows r0-2 .. r0+TH+1, words (c0 >> 6) - 1 .. + 1
  int any_doused;
  uint32_t fbits[T_ROWS_MAX][4];                         // fire bit-board of tile + halo: bit c of the row <-> tile column c (3 words + a zero pad)
  alignas(8) unsigned long long mbar;
  int nfront;
  int cnt_tree, cnt_fire;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// copies tile (e, ty, tx) from one grid buffer to the other (first 128 threads of the CTA)
__device__ __forceinline__ void copy_tile(int H, int W, int e, int ty, int tx, const uint8_t* __restrict__ src,
                                          uint8_t* __restrict__ dst) {
  const int tid = threadIdx.x;
  if (tid >= T_TH * 4) return;
  const int r = ty * T_TH + (tid >> 2), c = tx * T_TW + (tid & 3) * 16;
  if (r >= H || c >= W) return;
  const size_t off = ((size_t)e * H + r) * W + c;
  if ((W & 15) == 0) {
    *reinterpret_cast<uint4*>(dst + off) = *reinterpret_cast<const uint4*>(src + off);
  } else {
    for (int k = 0; k < 16 && c + k < W; ++k) dst[off + k] = src[off + k];
  }
}

// 5x5 doused window of tile cell (lr, lcc) from the staged rows: bit (5 i + j) <-> (r-2+i, c-2+j); outside the grid = 0
__device__ __forceinline__ uint32_t tile_dous_window(const TileSmem& sm, int lr, int lcc) {
  uint32_t dwin = 0u;
  if (sm.any_doused) {
    const int pos = 62 + lcc, wi = pos >> 6, sh = pos & 63;  // bit offset of column c-2 in the three staged words
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const unsigned long long lo = sm.dw[lr + i][wi], hi = sm.dw[lr + i][wi + 1 < 3 ? wi + 1 : 2];
      const unsigned long long bits = (lo >> sh) | (sh && wi + 1 < 3 ? hi << (64 - sh) : 0ull);
      dwin |= ((uint32_t)bits & 31u) << (5 * i);
    }
  }
  return dwin;
}

// One CA sub-step of tile (e, ty, tx) by the whole CTA: staged from `cell_in` (TMA map `tmap`), new cells to `cell_out`
// (`it` = how many tiles this CTA has staged before: the parity of the TMA barrier's phase).
template <bool USE_TMA>
__device__ __forceinline__ void ca_tile_body(TileSmem& sm, const gca_params& P, const gca_state& S, const gca_inject& J,
                                            const CUtensorMap* tmap, const uint8_t* __restrict__ cell_in,
                                            uint8_t* __restrict__ cell_out, const uint32_t* __restrict__ sched,
                                            int32_t* __restrict__ counts, unsigned long long* stats, int substep, int pitch,
                                            int e, int ty, int tx, int it, int sched_env_stride, int sched_off) {
  const int H = P.H, W = P.W, R = P.R, mode = P.rng_mode;
  const int WW = (W + 63) >> 6;
  const int r0 = ty * T_TH, c0 = tx * T_TW;
  const int tid = threadIdx.x, lane = tid & 31;
  const int rows = T_TH + 2 * R;
  const size_t env_off = (size_t)e * H * W;
  const int win = 2 * R + 1;

  if (tid == 0) { sm.nfront = 0; sm.cnt_tree = 0; sm.cnt_fire = 0; sm.any_doused = 0; }
  // doused rows within the 5x5 window's reach of the tile (requested before the tile is staged)
  unsigned long long dword = 0ull;
  if (tid < (T_TH + 4) * 3) {
    const int gr = r0 - 2 + tid / 3, wc = (c0 >> 6) - 1 + tid % 3;
    if (gr >= 0 && gr < H && wc >= 0 && wc < WW)
      dword = reinterpret_cast<const unsigned long long*>(S.doused)[((size_t)e * H + gr) * WW + wc];
  }
  // ---- stage tile + halo -------------------------------------------------------------------------
  if (USE_TMA) {
    if (tid == 0) {
      const uint32_t bar = smem_u32(&sm.mbar);
      if (it == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      const uint32_t bytes = (uint32_t)(rows * pitch);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      // box (pitch, rows, 1) at (c0 - 16, r0 - R, e); negative / beyond-edge coordinates are zero-filled.
      // The innermost start coordinate must keep the global address 16-byte aligned.
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"(smem_u32(sm.tile)), "l"(tmap), "r"(c0 - T_HC), "r"(r0 - R), "r"(e), "r"(bar)
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
            : "r"(bar), "r"((uint32_t)(it & 1))
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

  if (tid < (T_TH + 4) * 3) {
    sm.dw[tid / 3][tid % 3] = dword;
    if (dword) sm.any_doused = 1;  // (benign race: every writer stores 1; read after the next barrier)
  }
  // fire bit-board of the staged tile (a warp ballots 32 cells of a row into one word): the heat window becomes
  // popcounts of masked row words instead of a walk over (2R+1)^2 bytes
  for (int wi = tid >> 5; wi < rows * 3; wi += T_THREADS / 32) {
    const int row = wi / 3, q = wi % 3;
    const int v = sm.tile[row * pitch + 32 * q + lane];
    const uint32_t bf = __ballot_sync(GCA_FULL, v == 2), bt = __ballot_sync(GCA_FULL, v == 1);
    if (lane == 0) {
      sm.fbits[row][q] = bf;
      sm.tbits[row][q] = bt;
      if (q == 0) { sm.fbits[row][3] = 0u; sm.tbits[row][3] = 0u; }
    }
  }
  if (tid < T_TH) sm.ign[tid] = 0ull;
  if (tid == 0) { sm.nlist2 = 0; sm.npairs[0] = 0; sm.npairs[1] = 0; }
  __syncthreads();
  const uint32_t* sc = sched + (size_t)e * sched_env_stride + sched_off;
  const uint32_t tick = S.tick[e] + (uint32_t)substep;  // S.tick advances by K in the epilogue
  const uint32_t half_cell = (uint32_t)(((size_t)H * W) >> 1);
  const uint32_t half_burn = (uint32_t)((9ull * H * W) >> 1);

  // ---- find the tile's front cells: tree cells with a burning Moore neighbour, bit-parallel (a thread per tile row:
  //      64-bit views of the row's interior columns, the fire rows also shifted by one column either way) -------------
  if (tid < T_TH) {
    const int lr = tid;
    unsigned long long fh = 0ull;  // cells of row lr with fire in columns c-1..c+1 of rows lr-1..lr+1
#pragma unroll
    for (int d = -1; d <= 1; ++d) {
      const uint32_t* fr = sm.fbits[lr + R + d];
      const unsigned long long lo = fr[0] | ((unsigned long long)fr[1] << 32), hi = fr[2] | ((unsigned long long)fr[3] << 32);
      fh |= ((lo >> (T_HC - 1)) | (hi << (64 - (T_HC - 1)))) | ((lo >> T_HC) | (hi << (64 - T_HC))) |
            ((lo >> (T_HC + 1)) | (hi << (64 - (T_HC + 1))));
    }
    const uint32_t* tr = sm.tbits[lr + R];
    const unsigned long long tlo = tr[0] | ((unsigned long long)tr[1] << 32), thi = tr[2] | ((unsigned long long)tr[3] << 32);
    unsigned long long front = ((tlo >> T_HC) | (thi << (64 - T_HC))) & fh;  // (cells outside the grid are zero-filled: no trees)
    if (front) {
      int idx = atomicAdd(&sm.nfront, __popcll(front));
      while (front) {
        const int c = __ffsll((long long)front) - 1;
        front &= front - 1;
        sm.list[idx++] = (uint16_t)((lr << 6) | c);
      }
    }
  }
  __syncthreads();

  // ---- the front cells, in rounds of T_THREADS: a thread per cell computes the enclosure of the burn probability's
  //      direction-independent part and lists the cell's burning directions; the (cell, direction) draws are then dealt
  //      one per thread (a cell with five burning neighbours does not make its warp wait for five threefry blocks) -----
  const int nfront = sm.nfront;
  const TfKey kburn = tf_key(sc[SC_BURN0], sc[SC_BURN1]);
  const float* wind = P.winds + 9 * (int)sc[SC_WIND];
  uint32_t n_draws = 0, n_thresh = 0;
  int par = 0;
  for (int base = 0; base < nfront; base += T_THREADS, par ^= 1) {
    {
      const int i = base + tid;
      uint32_t nb = 0u;
      float blo = 0.0f, bhi = 0.0f;
      if (i < nfront) {
        const int lr = sm.list[i] >> 6, lcc = sm.list[i] & 63;
        const size_t gcell = (size_t)(r0 + lr) * W + (c0 + lcc);
        // the cell's hidden byte: requested now, used after the heat sum
        int hid = 3 | (3 << 3);
        if (S.hidden != nullptr) hid = S.hidden[env_off + gcell];
        // heat from the fire bit rows: H = sum_k w_k (C_k - C_{k-1}) = sum_k (w_k - w_{k+1}) C_k, C_k = burning cells
        // within Chebyshev distance k = popcounts of the 2k+1 window rows under a (2k+1)-bit mask (the centre is a tree).
        // Any summation order is inside the enclosure (T_LO / T_HI leave 2048 u, this sum is off by a few dozen u).
        uint32_t wr[2 * T_MAXR + 1];
        {
          const int start = lcc + T_HC - R;  // first tile column of the window
          const int wq = start >> 5, sh = start & 31;
#pragma unroll
          for (int q = 0; q <= 2 * T_MAXR; ++q) {
            const int di = q - T_MAXR;
            wr[q] = 0u;
            if (di >= -R && di <= R) {
              const uint32_t* fr = sm.fbits[lr + R + di];
              wr[q] = __funnelshift_r(fr[wq], fr[wq + 1], sh);  // bit b <-> column offset b - R
            }
          }
        }
        float Hf = 0.0f;
#pragma unroll
        for (int k = T_MAXR; k >= 1; --k) {
          if (k <= R) {
            const uint32_t box = ((2u << (2 * k)) - 1u) << (R - k);  // |dj| <= k
            int Ck = 0;
#pragma unroll
            for (int di = -k; di <= k; ++di) Ck += __popc(wr[di + T_MAXR] & box);
            const float dk = k < R ? __fsub_rn(P.ring_w[k], P.ring_w[k + 1]) : P.ring_w[k];
            Hf = fmaf((float)Ck, dk, Hf);
          }
        }
        float Dlo = 0.0f, Dhi = 0.0f;
        const uint32_t dwin = tile_dous_window(sm, lr, lcc);
        if (dwin) {
          const int ni = __popc(dwin & ((0x0Eu << 5) | (0x0Eu << 10) | (0x0Eu << 15)));
          const int nbd = __popc(dwin) - ni;
          const float Df = fmaf((float)nbd, P.dous_border, (float)ni * P.dous_inner);
          Dlo = __fmul_rn(Df, T_LO);
          Dhi = __fmul_rn(Df, T_HI);
        }
        const float a = P.onep_veg[clip15(hid & 7)], b = P.onep_den[clip15((hid >> 3) & 7)];
        blo = __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(Hf, T_LO), Dhi), a), b);
        bhi = __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(Hf, T_HI), Dlo), a), b);
        if (bhi > 0.0f)  // burning Moore neighbours: bits R-1 .. R+1 of the three middle window rows; d = 3 i + j
          nb = ((((wr[T_MAXR - 1] >> (R - 1)) & 7u)) | (((wr[T_MAXR] >> (R - 1)) & 7u) << 3) |
                (((wr[T_MAXR + 1] >> (R - 1)) & 7u) << 6)) & ~16u;
      }
      sm.bnd[tid] = make_float2(blo, bhi);
      if (nb) {
        int at = atomicAdd(&sm.npairs[par], __popc(nb));
        while (nb) {
          const uint32_t d = (uint32_t)__ffs((int)nb) - 1u;
          nb &= nb - 1u;
          sm.pairs[at++] = (uint16_t)(((uint32_t)tid << 4) | d);
        }
      }
    }
    __syncthreads();
    const int npairs = sm.npairs[par];
    if (tid == 0) sm.npairs[par ^ 1] = 0;  // the next round's counter (its last readers passed the barrier above)
    for (int q = tid; q < npairs; q += T_THREADS) {
      const uint32_t rec = sm.pairs[q];
      const int t = (int)(rec >> 4), d = (int)(rec & 15u);
      const int lr = sm.list[base + t] >> 6, lcc = sm.list[base + t] & 63;
      const size_t gcell = (size_t)(r0 + lr) * W + (c0 + lcc);
      const float2 bd = sm.bnd[t];
      const float s = S.pslope ? S.pslope[(env_off + gcell) * 8 + dir_slot(d)] : 1.0f;
      float u;
      if (J.u_burn) u = J.u_burn[(((size_t)substep * S.N + e) * H * W + gcell) * 9 + d];
      else u = bits_to_uniform(bits_at(kburn, (uint32_t)(gcell * 9 + d), half_burn, mode));
      ++n_draws;
      const float w = wind[d];
      const float plo = __fmul_rn(__fmul_rn(bd.x, w), s), phi = __fmul_rn(__fmul_rn(bd.y, w), s);
      bool ig = u < plo;
      if (!ig && u < phi) {
        // threshold cell: reference-order (row-major, float32, from +0) sums
        ++n_thresh;
        const uint8_t* ctr = sm.tile + (lr + R) * pitch + (lcc + T_HC);
        float heat = 0.0f;
        for (int di = 0; di < win; ++di)
          for (int dj = 0; dj < win; ++dj)
            if (ctr[(di - R) * pitch + (dj - R)] == 2) heat = __fadd_rn(heat, P.ring_w[max(abs(di - R), abs(dj - R))]);
        const uint32_t dwin = tile_dous_window(sm, lr, lcc);
        float dous = 0.0f;
        for (int qq = 0; qq < 25; ++qq)
          if ((dwin >> qq) & 1u) {
            const int qi = qq / 5, qj = qq % 5;
            const bool inner = qi >= 1 && qi <= 3 && qj >= 1 && qj <= 3;
            dous = __fadd_rn(dous, inner ? P.dous_inner : P.dous_border);
          }
        int hid = 3 | (3 << 3);
        if (S.hidden != nullptr) hid = S.hidden[env_off + gcell];
        const float a = P.onep_veg[clip15(hid & 7)], b = P.onep_den[clip15((hid >> 3) & 7)];
        const float base_exact = __fmul_rn(__fmul_rn(__fsub_rn(heat, dous), a), b);
        ig = u < __fmul_rn(__fmul_rn(base_exact, w), s);
      }
      if (ig) atomicOr(&sm.ign[lr], 1ull << lcc);
    }
    __syncthreads();
  }

  // ---- write the new grid, burn-out ticks, counts -------------------------------------------------
  const TfKey ka1 = tf_key(sc[SC_AK10], sc[SC_AK11]), ka2 = tf_key(sc[SC_AK20], sc[SC_AK21]);
  const TfKey kg = tf_key(sc[SC_GROW0], sc[SC_GROW1]);
  int nt = 0, nfire = 0;
  uint32_t n_ign = 0, n_ext = 0;
  const bool dense_rule = P.p_tree > 0.0f;  // regrowth draws for every empty cell: the per-cell loop
  const int lc = tid & 63, rg = tid >> 6;
  if (!dense_rule) {
    // only burning cells (burn-out) and igniting cells can change: a thread per row lists them, all threads work the
    // list off (age draws / burn-out ticks), the changed bytes are patched in the staged tile and the tile's interior
    // leaves as 128-bit stores
    if (tid < T_TH) {
      const int lr = tid;
      const uint32_t* fr = sm.fbits[lr + R];
      const unsigned long long lo = fr[0] | ((unsigned long long)fr[1] << 32), hi = fr[2] | ((unsigned long long)fr[3] << 32);
      unsigned long long fire = (lo >> T_HC) | (hi << (64 - T_HC));
      unsigned long long ig = sm.ign[lr];
      const int n = __popcll(fire) + __popcll(ig);
      if (n) {
        int idx = atomicAdd(&sm.nlist2, n);
        while (fire) {
          const int c = __ffsll((long long)fire) - 1;
          fire &= fire - 1;
          sm.list[idx++] = (uint16_t)((lr << 6) | c);
        }
        while (ig) {
          const int c = __ffsll((long long)ig) - 1;
          ig &= ig - 1;
          sm.list[idx++] = (uint16_t)(0x8000 | (lr << 6) | c);
        }
      }
    }
    __syncthreads();
    const int n2 = sm.nlist2;
    for (int i = tid; i < n2; i += T_THREADS) {
      const uint32_t ent = sm.list[i];
      const int lr = (ent >> 6) & 31, lcc = ent & 63;
      const size_t gcell = (size_t)(r0 + lr) * W + (c0 + lcc);
      uint8_t* cellp = sm.tile + (lr + R) * pitch + (lcc + T_HC);
      if (ent & 0x8000u) {  // tree -> fire: age draw, burn-out tick
        int age;
        if (J.age_new) age = J.age_new[((size_t)substep * S.N + e) * H * W + gcell];
        else age = randint_from_bits(bits_at_ni(ka1, (uint32_t)gcell, half_cell, mode),
                                     bits_at_ni(ka2, (uint32_t)gcell, half_cell, mode), P.age_lo, P.age_span, P.age_mult);
        S.death[env_off + gcell] = (uint16_t)(tick + (uint32_t)age);
        *cellp = 2;
        ++n_ign;
      } else if ((((uint32_t)S.death[env_off + gcell] - tick) & 0xFFFFu) == 0u) {  // fire_age <= 1 -> empty, age ends at 0
        S.death[env_off + gcell] = 0;
        *cellp = 0;
        ++n_ext;
      }
    }
    nt = -(int)n_ign;
    nfire = (int)n_ign - (int)n_ext;
    __syncthreads();
    if ((W & 15) == 0) {
      if (tid < T_TH * 4) {
        const int lr = tid >> 2, q = tid & 3;
        const int gr = r0 + lr, gc = c0 + 16 * q;
        if (gr < H && gc < W)
          *reinterpret_cast<uint4*>(cell_out + env_off + (size_t)gr * W + gc) =
              *reinterpret_cast<const uint4*>(sm.tile + (lr + R) * pitch + T_HC + 16 * q);
      }
    } else {
      for (int k = 0; k < T_TH / 4; ++k) {
        const int lr = rg + 4 * k, gr = r0 + lr, gc = c0 + lc;
        if (gr < H && gc < W) cell_out[env_off + (size_t)gr * W + gc] = sm.tile[(lr + R) * pitch + (lc + T_HC)];
      }
    }
  } else {
  // the burn-out ticks of this thread's burning cells first (independent loads), then the rule
  uint16_t dth[T_TH / 4];
#pragma unroll
  for (int k = 0; k < T_TH / 4; ++k) {
    const int lr = rg + 4 * k;
    const int gr = r0 + lr, gc = c0 + lc;
    dth[k] = 0;
    if (gr < H && gc < W && sm.tile[(lr + R) * pitch + (lc + T_HC)] == 2) dth[k] = S.death[env_off + (size_t)gr * W + gc];
  }
#pragma unroll
  for (int k = 0; k < T_TH / 4; ++k) {
    const int lr = rg + 4 * k;
    const int gr = r0 + lr, gc = c0 + lc;
    if (gr >= H || gc >= W) continue;
    const size_t gcell = (size_t)gr * W + gc;
    const int old = sm.tile[(lr + R) * pitch + (lc + T_HC)];
    int nw = old;
    if (old == 1 && ((sm.ign[lr] >> lc) & 1ull)) {
      nw = 2;
      int age;
      if (J.age_new) age = J.age_new[((size_t)substep * S.N + e) * H * W + gcell];
      else age = randint_from_bits(bits_at_ni(ka1, (uint32_t)gcell, half_cell, mode),
                                   bits_at_ni(ka2, (uint32_t)gcell, half_cell, mode), P.age_lo, P.age_span, P.age_mult);
      S.death[env_off + gcell] = (uint16_t)(tick + (uint32_t)age);  // burns out at tick + age
      ++n_ign;
    } else if (old == 0) {
      if (P.p_tree > 0.0f) {
        float u;
        if (J.u_grow) u = J.u_grow[((size_t)substep * S.N + e) * H * W + gcell];
        else u = bits_to_uniform(bits_at_ni(kg, (uint32_t)gcell, half_cell, mode));
        if (u < P.p_tree) nw = 1;
      }
    } else if (old == 2) {
      if ((((uint32_t)dth[k] - tick) & 0xFFFFu) == 0u) {  // fire_age <= 1 -> empty, age ends at 0
        nw = 0;
        S.death[env_off + gcell] = 0;
        ++n_ext;
      }
    }
    cell_out[env_off + gcell] = (uint8_t)nw;
    nt += (nw == 1) - (old == 1);     // change of the env's tree / fire counts
    nfire += (nw == 2) - (old == 2);
  }
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

// The tile list of an env step (built once, from the fire counts at its start).  A burning cell moves at most one cell
// per sub-step and a tile is 32 x 64 cells, so for K <= 32 sub-steps every tile that can become active during the step
// -- fire in its 3x3 tile neighbourhood at that time -- lies within Chebyshev tile distance 2 of a tile that holds fire
// NOW: those tiles are listed for computation in every sub-step.  The ring at distance 3 is listed "copy only"
// (T_COPY_ONLY): its tiles cannot change, but their cells are the halo of the computed tiles, which read the two grid
// buffers alternately (see ca_tiled_list_kernel).  With regrowth every tile is computed.  All 32 lanes of a warp call.
constexpr uint32_t T_COPY_ONLY = 0x80000000u;
__device__ __forceinline__ void list_append(long long gid, long long total, int TX, int TY, const uint32_t* __restrict__ hasfire,
                                            uint32_t* __restrict__ list, int* __restrict__ nactive, int all_active) {
  int dist = 4;  // 4 = not listed
  if (gid < total) {
    if (all_active) dist = 0;
    else {
      // `hasfire`: one bit per tile (set by tile_count_kernel), tile t = bit t & 31 of word t >> 5; the 7 x 7 tile
      // neighbourhood is 7 bit fields of up to 7 bits
      const int tx = (int)(gid % TX), ty = (int)((gid / TX) % TY);
      const long long env0 = gid - (long long)ty * TX - tx;  // the env's tile (0, 0)
      const int x0 = max(tx - 3, 0), x1 = min(tx + 3, TX - 1), n = x1 - x0 + 1;
      for (int dy = -3; dy <= 3; ++dy) {
        const int y = ty + dy;
        if (y < 0 || y >= TY) continue;
        const long long p = env0 + (long long)y * TX + x0;
        const int sh = (int)(p & 31);
        const uint32_t lo = hasfire[p >> 5], hi = sh + n > 32 ? hasfire[(p >> 5) + 1] : 0u;
        const uint32_t bits = (__funnelshift_r(lo, hi, sh) & ((1u << n) - 1u)) << (x0 - (tx - 3));  // bit k <-> dx = k - 3
        if (!bits) continue;
        const int mdx = (bits & 0x08u) ? 0 : ((bits & 0x14u) ? 1 : ((bits & 0x22u) ? 2 : 3));
        dist = min(dist, max(mdx, abs(dy)));
      }
    }
  }
  const bool act = dist <= 3;
  const uint32_t bal = __ballot_sync(GCA_FULL, act);
  if (bal) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(&nactive[0], __popc(bal));
    base = __shfl_sync(GCA_FULL, base, 0);
    if (act) list[base + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)gid | (dist == 3 ? T_COPY_ONLY : 0u);
  }
}

// The key schedules of ALL K sub-steps of env e by one warp (the 3K split levels of the key chain are sequential, a lane
// pair runs the two blocks of a split; then lane pair 2j derives Sburn / Sgrow / the randint keys of sub-step j and pair
// 2j+1 its wind draws, 4 more levels) into sched[e][j][12]; the env's key and wind index advance.
__device__ __forceinline__ void tiled_sched_warp(const gca_params& P, const gca_state& S, const gca_inject& J, uint32_t* sched,
                                                 int e, int lane) {
  const int K = P.K, mode = P.rng_mode, N = S.N;
  uint32_t* se = sched + (size_t)e * SC_N * GCA_MAX_K;
  const int pair = lane >> 1, j = pair >> 1;
  const uint32_t w = lane & 1;
  const bool burn_role = (pair & 1) == 0;
  uint32_t k0 = S.key[2 * e], k1 = S.key[2 * e + 1];
  uint32_t c0 = 0, c1 = 0, sw0 = 0, sw1 = 0;  // burn role: S1 ; wind role: Sidx (c) and Swind (sw) of sub-step j
  for (int q = 0; q < K; ++q) {
    uint32_t n0, n1, s0, s1;
    split_pair(k0, k1, mode, lane, n0, n1, s0, s1);  // K1, S1
    if (q == j && burn_role) { c0 = s0; c1 = s1; }
    k0 = n0; k1 = n1;
    split_pair(k0, k1, mode, lane, n0, n1, s0, s1);  // K2, Swind
    if (q == j && !burn_role) { sw0 = s0; sw1 = s1; }
    k0 = n0; k1 = n1;
    split_pair(k0, k1, mode, lane, n0, n1, s0, s1);  // K3, Sidx
    if (q == j && !burn_role) { c0 = s0; c1 = s1; }
    k0 = n0; k1 = n1;
  }
  const uint32_t sc0 = (mode == GCA_RNG_LEGACY) ? w : 0u;  // split counters of this lane
  const uint32_t sc1 = (mode == GCA_RNG_LEGACY) ? w + 2u : w;
  uint32_t o0, o1, p0, p1, n0, n1, s0, s1;
  // level 1: burn: split(S1) -> Ka, Sburn ; wind: split(Sidx) -> wk1, wk2
  tf_exchange(c0, c1, sc0, sc1, o0, o1, p0, p1);
  assemble_split(mode, w, o0, o1, p0, p1, n0, n1, s0, s1);
  const uint32_t sburn0 = s0, sburn1 = s1;  // (wind role: wk2)
  uint32_t cur0 = n0, cur1 = n1;            // burn: Ka ; wind: wk1
  // level 2: burn: split(Ka) -> Kb, Sgrow ; wind: even lane bits(wk1, ()), odd lane bits(wk2, ())
  tf_exchange(burn_role ? cur0 : (w ? sburn0 : cur0), burn_role ? cur1 : (w ? sburn1 : cur1), burn_role ? sc0 : 0u,
              burn_role ? sc1 : 0u, o0, o1, p0, p1);
  assemble_split(mode, w, o0, o1, p0, p1, n0, n1, s0, s1);
  const uint32_t sgrow0 = s0, sgrow1 = s1;
  const uint32_t my_bits = (mode == GCA_RNG_LEGACY) ? o0 : (o0 ^ o1);
  const uint32_t pr_bits = (mode == GCA_RNG_LEGACY) ? p0 : (p0 ^ p1);
  const uint32_t hb = w ? pr_bits : my_bits, lb = w ? my_bits : pr_bits;
  cur0 = n0; cur1 = n1;  // burn: Kb
  // level 3: burn: split(Kb) -> Kc, Sage ; wind: bits(Swind, ())
  tf_exchange(burn_role ? cur0 : sw0, burn_role ? cur1 : sw1, burn_role ? sc0 : 0u, burn_role ? sc1 : 0u, o0, o1, p0, p1);
  assemble_split(mode, w, o0, o1, p0, p1, n0, n1, s0, s1);
  const uint32_t uw_bits = (mode == GCA_RNG_LEGACY) ? o0 : (o0 ^ o1);
  // level 4: burn: split(Sage) -> ak1, ak2
  tf_exchange(s0, s1, sc0, sc1, o0, o1, p0, p1);
  uint32_t a10, a11, a20, a21;
  assemble_split(mode, w, o0, o1, p0, p1, a10, a11, a20, a21);
  if (j < K && w == 0) {
    uint32_t* sc = se + j * SC_N;
    if (burn_role) {
      sc[SC_BURN0] = sburn0; sc[SC_BURN1] = sburn1; sc[SC_GROW0] = sgrow0; sc[SC_GROW1] = sgrow1;
      sc[SC_AK10] = a10; sc[SC_AK11] = a11; sc[SC_AK20] = a20; sc[SC_AK21] = a21;
    } else {
      float u = bits_to_uniform(uw_bits);
      int step = randint_from_bits(hb, lb, 1, 7u, 4u);
      if (J.u_wind) u = J.u_wind[(size_t)j * N + e];
      if (J.wind_step) step = J.wind_step[(size_t)j * N + e];
      sc[9] = (u < P.p_wind_change) ? 1u : 0u;
      sc[10] = (uint32_t)step;
    }
  }
  __syncwarp();
  if (lane == 0) {  // the wind index is threaded through the sub-steps; the chain's end is the env's new key
    int wi = S.wind_index[e];
    for (int q = 0; q < K; ++q) {
      uint32_t* sc = se + q * SC_N;
      sc[SC_WIND] = (uint32_t)wi;
      if (sc[9]) wi = (wi + (int)sc[10]) % 8;
    }
    S.wind_index[e] = wi;
    S.key[2 * e] = k0;
    S.key[2 * e + 1] = k1;
  }
}

// ---------------------------------------------------------------------------------------------
// tile activity (see the head of the file)
//   tile index t = (e * TY + ty) * TX + tx;  hasfire: bit t & 31 of word t >> 5;  list[] u32 (| T_COPY_ONLY);
//   nactive[0] = entries of the step's list
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_count_kernel(gca_params P, gca_state S, gca_inject J, uint32_t* sched, int TX,
                                                         int TY, uint32_t* __restrict__ hasfire, int32_t* __restrict__ counts) {
  const int N = S.N, H = P.H, W = P.W;
  const uint8_t* __restrict__ cell = S.cell;
  {
    // the envs' key schedules ride along: the LAST N warps of the grid walk them first (16 dependent threefry levels,
    // hidden behind the other warps' counting), so the step needs no schedule kernel
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (long long e = nw - 1 - gw; e < N; e += nw) tiled_sched_warp(P, S, J, sched, (int)e, threadIdx.x & 31);
  }
  // a warp per tile: lane l reads 16 cells of rows (l >> 2) + 8 q, q = 0..3 -- four independent 128-bit loads in flight
  const int lane = threadIdx.x & 31;
  const long long total = (long long)N * TY * TX;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  // the counts of the env of the CTA's first tile are summed in shared memory first
  __shared__ int s_cnt[2];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long t_cta = (long long)blockIdx.x * (blockDim.x >> 5);
  const int e_cta = t_cta < total ? (int)(t_cta / ((long long)TX * TY)) : N;
  int acc_t = 0, acc_f = 0;
  for (long long t = t_cta + (threadIdx.x >> 5); t < total; t += nwarps) {
    const int tx = (int)(t % TX), ty = (int)((t / TX) % TY), e = (int)(t / ((long long)TX * TY));
    const int c = tx * T_TW + (lane & 3) * 16;
    int nt = 0, nf = 0;
    if ((W & 15) == 0) {
      uint4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = ty * T_TH + (lane >> 2) + 8 * q;
        v[q] = make_uint4(0u, 0u, 0u, 0u);
        if (r < H && c < W) v[q] = *reinterpret_cast<const uint4*>(cell + ((size_t)e * H + r) * W + c);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // cell codes 0 / 1 / 2: bit 0 = tree, bit 1 = fire
          nt += __popc(w[k] & 0x01010101u);
          nf += __popc(w[k] & 0x02020202u);
        }
      }
    } else {
      for (int q = 0; q < 4; ++q) {
        const int r = ty * T_TH + (lane >> 2) + 8 * q;
        if (r >= H) continue;
        const uint8_t* src = cell + ((size_t)e * H + r) * W + c;
        for (int k = 0; k < 16 && c + k < W; ++k) {
          const int v = src[k];
          nt += v == 1;
          nf += v == 2;
        }
      }
    }
    nt = __reduce_add_sync(GCA_FULL, nt);
    nf = __reduce_add_sync(GCA_FULL, nf);
    if (lane == 0) {
      if (nf) atomicOr(&hasfire[t >> 5], 1u << (t & 31));
      if (e == e_cta) { acc_t += nt; acc_f += nf; }  // (thousands of tiles of one env: no global atomic per tile)
      else {
        if (nt) atomicAdd(&counts[2 * e], nt);
        if (nf) atomicAdd(&counts[2 * e + 1], nf);
      }
    }
  }
  if (lane == 0) {
    if (acc_t) atomicAdd(&s_cnt[0], acc_t);
    if (acc_f) atomicAdd(&s_cnt[1], acc_f);
  }
  __syncthreads();
  if (threadIdx.x == 0 && e_cta < N) {
    if (s_cnt[0]) atomicAdd(&counts[2 * e_cta], s_cnt[0]);
    if (s_cnt[1]) atomicAdd(&counts[2 * e_cta + 1], s_cnt[1]);
  }
}

// the step's tile list (thread per tile), once the counts of tile_count_kernel are complete
__global__ void __launch_bounds__(256) tile_list_kernel(int N, int TX, int TY, const uint32_t* __restrict__ hasfire,
                                                        uint32_t* __restrict__ list, int* __restrict__ nactive, int all_active) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  list_append(gid, (long long)N * TY * TX, TX, TY, hasfire, list, nactive, all_active);
}


// One CA sub-step of every listed tile.  The grid lives in two buffers, A = S.cell and B = the scratch grid: sub-step j
// reads tiles + halos from one (A when j is even) and writes the tiles' new cells to the other, so no copy-back pass
// separates the sub-steps -- neighbouring tiles still read the old cells while a tile writes its new ones.  Both buffers
// hold the current grid on the computed tiles after every sub-step (each is rewritten every other sub-step from the
// other); the copy-only ring around them is brought into B once, in sub-step 0; nothing else is ever read from B.
template <bool USE_TMA>
__global__ void __launch_bounds__(T_THREADS, 3)
ca_tiled_list_kernel(const __grid_constant__ gca_params P, const __grid_constant__ gca_state S,
                     const __grid_constant__ gca_inject J, const __grid_constant__ CUtensorMap tmap_a,
                     const __grid_constant__ CUtensorMap tmap_b, uint8_t* __restrict__ cell_b,
                     const uint32_t* __restrict__ sched, int32_t* __restrict__ counts, unsigned long long* stats,
                     const uint32_t* __restrict__ list, const int* __restrict__ nactive, int substep, int pitch, int TX,
                     int TY) {
  __shared__ TileSmem sm;
  const int n = nactive[0];
  const bool from_a = (substep & 1) == 0;
  const uint8_t* cell_in = from_a ? S.cell : cell_b;
  uint8_t* cell_out = from_a ? cell_b : S.cell;
  const CUtensorMap* tmap = from_a ? &tmap_a : &tmap_b;
  int it = 0;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const uint32_t ent = list[i], t = ent & ~T_COPY_ONLY;
    const int tx = (int)(t % (uint32_t)TX), ty = (int)((t / (uint32_t)TX) % (uint32_t)TY), e = (int)(t / (uint32_t)(TX * TY));
    if (ent & T_COPY_ONLY) {
      if (substep == 0) copy_tile(P.H, P.W, e, ty, tx, cell_in, cell_out);
      continue;
    }
    ca_tile_body<USE_TMA>(sm, P, S, J, tmap, cell_in, cell_out, sched, counts, stats, substep, pitch, e, ty, tx, it,
                          SC_N * GCA_MAX_K, SC_N * substep);
    ++it;
    __syncthreads();  // the shared tile is free for the next entry
  }
}

// K odd: the last sub-step wrote the scratch grid -- the computed tiles go back to S.cell
__global__ void __launch_bounds__(T_THREADS) tile_copy_back_kernel(int H, int W, int TX, int TY,
                                                                   const uint32_t* __restrict__ list,
                                                                   const int* __restrict__ nactive,
                                                                   const uint8_t* __restrict__ scratch, uint8_t* __restrict__ cell) {
  const int n = nactive[0];
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const uint32_t ent = list[i], t = ent & ~T_COPY_ONLY;
    if (ent & T_COPY_ONLY) continue;
    copy_tile(H, W, (int)(t / (uint32_t)(TX * TY)), (int)((t / (uint32_t)TX) % (uint32_t)TY), (int)(t % (uint32_t)TX), scratch, cell);
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
  const bool full = !(flags & GCA_FLAG_CA_ONLY);
  // every load first (independent, one DRAM round trip), the stores after: as written in program order the compiler must
  // keep each load behind the preceding store (they may alias), which makes the kernel a chain of cold misses
  const uint32_t tick = S.tick[e];
  const int t = counts[2 * e], f = counts[2 * e + 1];
  int a0 = 4, a1 = 0, row = 0, col = 0, ts = 0, night = 0;
  float tm = 0.0f, se = 0.0f, ra = 0.0f;
  if (full) {
    a0 = actions[3 * e]; a1 = actions[3 * e + 1];
    tm = S.time[e];
    row = S.position[2 * e]; col = S.position[2 * e + 1];
    ts = S.time_step[e];
    night = S.is_night[e];
    if (S.steps_elapsed) se = S.steps_elapsed[e];
    if (S.reward_accumulated) ra = S.reward_accumulated[e];
  }
  S.tick[e] = tick + (uint32_t)P.K;
  const float rew = award(t, f);
  const bool done = f == 0;
  if (full) {
    const int a0c = min(max(a0, 0), 8), a1c = min(max(a1, 0), 1);
    const float tt = __fadd_rn(__fadd_rn(P.t_move[a0c], P.t_shoot[a1c]), P.t_any);
    const float ntm = __fadd_rn(tm, tt);
    S.time[e] = __fsub_rn(ntm, truncf(ntm));
    move_position(a0, P.H, P.W, row, col);
    S.position[2 * e] = row;
    S.position[2 * e + 1] = col;
    if (a1 == 1)
      atomicOr(reinterpret_cast<unsigned long long*>(S.doused) + ((size_t)e * P.H + row) * WW + (col >> 6), 1ull << (col & 63));
    ts += 1;
    S.time_step[e] = ts;
    if (O.obs_night) O.obs_night[e] = (uint8_t)night;
    if (ts % P.day_length == 0) night = 1 - night;
    S.is_night[e] = night;
    if (S.steps_elapsed) S.steps_elapsed[e] = __fadd_rn(se, 1.0f);
    if (S.reward_accumulated) S.reward_accumulated[e] = __fadd_rn(ra, rew);
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

// One env step: one clear, one dense count, the key schedules of all sub-steps + the step's tile list, then ONE kernel
// with a small fixed grid per sub-step (+ a copy-back when K is odd) and the epilogue.  Layout of `aux` (words):
// nactive[16] | hasfire[tiles / 32 + 1] (one bit per tile) | list[tiles],
// with nactive directly behind the counts so that one clear covers both.
static cudaError_t enqueue_tiled_env_step(const gca_params& p, const gca_state& s, const int32_t* actions,
                                                 const gca_step_out& out, const gca_inject& inj, uint32_t flags,
                                          uint8_t* scratch_cell, uint32_t* scratch_sched, int32_t* scratch_counts,
                                          uint8_t* aux8, int use_tma, const gca_state* snap, const float* snap_reward,
                                          cudaStream_t st) {
  uint32_t* aux = reinterpret_cast<uint32_t*>(aux8);
  const int N = s.N, H = p.H, W = p.W, R = p.R;
  const int pitch = T_PITCH_MAX;
  const int rows = T_TH + 2 * R;
  const int TX = (W + T_TW - 1) / T_TW, TY = (H + T_TH - 1) / T_TH;
  const long long tiles = (long long)N * TX * TY;
  if (tiles >= (1ll << 31)) return cudaErrorInvalidValue;
  int* nactive = reinterpret_cast<int*>(aux);
  const long long bw = (tiles + 31) / 32 + 1;   // words of the one-bit-per-tile "holds fire" map
  uint32_t* hasfire = aux + 16;
  uint32_t* list = hasfire + bw;
  if (reinterpret_cast<int32_t*>(nactive) != scratch_counts + 2 * (size_t)N) return cudaErrorInvalidValue;
  cudaError_t err;
  const int all_active = p.p_tree > 0.0f ? 1 : 0;  // regrowth can change any empty cell
  if ((err = cudaMemsetAsync(scratch_counts, 0, sizeof(int32_t) * (2 * (size_t)N + 16 + (size_t)bw), st)) != cudaSuccess) return err;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  const int g_count = (int)((tiles + 7) / 8 < (long long)sms * 4 ? (tiles + 7) / 8 : (long long)sms * 4);  // 8 warps per CTA, a warp per tile
  const int g_tile = (int)(tiles < (long long)sms * 4 ? tiles : (long long)sms * 4);
  tile_count_kernel<<<g_count, 256, 0, st>>>(p, s, inj, scratch_sched, TX, TY, hasfire, scratch_counts);
  CUtensorMap tm_a, tm_b;
  memset(&tm_a, 0, sizeof(tm_a));
  memset(&tm_b, 0, sizeof(tm_b));
  const bool tma = use_tma && make_tmap(&tm_a, s.cell, N, H, W, pitch, rows) && make_tmap(&tm_b, scratch_cell, N, H, W, pitch, rows);
  tile_list_kernel<<<(unsigned)((tiles + 255) / 256), 256, 0, st>>>(N, TX, TY, hasfire, list, nactive, all_active);
  for (int j = 0; j < p.K; ++j) {
    if (tma)
      ca_tiled_list_kernel<true><<<g_tile, T_THREADS, 0, st>>>(p, s, inj, tm_a, tm_b, scratch_cell, scratch_sched,
                                                               scratch_counts, out.stats, list, nactive, j, pitch, TX, TY);
    else
      ca_tiled_list_kernel<false><<<g_tile, T_THREADS, 0, st>>>(p, s, inj, tm_a, tm_b, scratch_cell, scratch_sched,
                                                                scratch_counts, out.stats, list, nactive, j, pitch, TX, TY);
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
  }
  if (p.K & 1) tile_copy_back_kernel<<<g_tile, T_THREADS, 0, st>>>(H, W, TX, TY, list, nactive, scratch_cell, s.cell);
  tiled_finish_kernel<<<(N + 127) / 128, 128, 0, st>>>(p, s, actions, out, scratch_counts, flags);
  if ((err = cudaGetLastError()) != cudaSuccess) return err;
  // the fused conditional_reset rides in the same graph
  if (snap != nullptr) return launch_auto_reset(p, s, *snap, snap_reward, out.reward, out.terminated, st);
  return cudaSuccess;
}

// The K + 3 launches of one env step as ONE CUDA graph launch: at 4096x4096 with a small fire the step is bound by
// launch latency.  The graph is captured once per set of
// buffers (thread-local cache of one entry; a capture stream of its own: the caller's may be the legacy stream, which
// cannot capture) and replayed with only the action pointer of the last kernel patched.  GCA_TILED_GRAPH=0 disables it.
namespace {
struct TiledGraphKey {
  gca_params p; gca_state s; gca_step_out out; uint32_t flags; const void *a, *b, *c, *d; int use_tma;
  gca_state snap; const void* snap_reward; int has_snap;
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
                                  const gca_state* snap, const float* snap_reward, cudaStream_t st) {
  static const bool enabled = [] { const char* v = getenv("GCA_TILED_GRAPH"); return !(v && v[0] == '0'); }();
  TiledGraphCache& G = g_tiled_graph;
  const bool injected = inj.u_burn || inj.u_grow || inj.age_new || inj.u_wind || inj.wind_step;
  if (!enabled || G.broken || injected)
    return enqueue_tiled_env_step(p, s, actions, out, inj, flags, scratch_cell, scratch_sched, scratch_counts, tile_flags,
                                  use_tma, snap, snap_reward, st);
  TiledGraphKey key;
  memset(&key, 0, sizeof(key));
  key.p = p; key.s = s; key.out = out; key.flags = flags;
  key.out.done_token = 0; key.out.rgb = nullptr; key.out.rgb_u8 = 0;  // per-step values no tiled kernel reads
  key.out.host_done = nullptr; key.out.done_counter = nullptr;
  key.a = scratch_cell; key.b = scratch_sched; key.c = scratch_counts; key.d = tile_flags; key.use_tma = use_tma;
  if (snap != nullptr) { key.snap = *snap; key.snap_reward = snap_reward; key.has_snap = 1; }
  auto fallback = [&]() {
    G.broken = true;
    cudaGetLastError();
    return enqueue_tiled_env_step(p, s, actions, out, inj, flags, scratch_cell, scratch_sched, scratch_counts, tile_flags,
                                  use_tma, snap, snap_reward, st);
  };
  if (!G.valid || memcmp(&G.key, &key, sizeof(key)) != 0) {
    if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
    if (G.graph) { cudaGraphDestroy(G.graph); G.graph = nullptr; }
    G.valid = false;
    if (!G.cap && cudaStreamCreateWithFlags(&G.cap, cudaStreamNonBlocking) != cudaSuccess) return fallback();
    if (cudaStreamBeginCapture(G.cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return fallback();
    const cudaError_t e1 = enqueue_tiled_env_step(p, s, actions, out, inj, flags, scratch_cell, scratch_sched, scratch_counts,
                                                  tile_flags, use_tma, snap, snap_reward, G.cap);
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
