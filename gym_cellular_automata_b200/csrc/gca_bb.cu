// gca_bb.cu -- fused environment step for grids that fit one CTA's shared memory as bit-boards
// (W a multiple of 64, H * W <= 65536: 128x128, 192x192, 256x256 -- BASELINE config 3's grid): ONE launch per env
// step, one CTA per env, all K CA sub-steps on-chip (temporal blocking: the grid is read from HBM once and written
// once per env step; there is no halo -- the tile IS the grid, outside is the reference's zero padding).
//
// Same rule, same lazy counter-based draws addressed by the global linear index and the same enclosure /
// exact-fallback logic as the 64x64 kernel (gca_step64.cu) and the generic tiled kernel (gca_tiled.cu); reference
// lines as there (ca_alexandridis_jax.py:321-460, repeat_ca_jax.py:34-71, move_modify_jax.py:39-157,
// advanced_bulldozer.py:332-399,1103-1133).  What differs from the tiled kernel:
//   * tree / fire / doused masks are 64-bit words per 64 cells of a row (H * W / 8 bytes per mask: 8 KB at 256x256),
//     built from the u8 grid with 128-bit loads at the start of the step and written back as bytes at its end;
//   * the front (tree with a burning Moore neighbour) is three shifted ORs per word -- skipped for words whose three
//     rows hold no fire (per-row flags) --; front cells are compacted into a CTA-wide list and worked on in rounds of
//     256: a thread per cell cuts the (2R+1)^2 heat window out of the fire rows as 2R+1 bit fields (ring populations by
//     popc, ~250 instructions where the byte version walks 169 cells) and lists the cell's burning directions, then the
//     (cell, direction) draws are dealt one per thread over the CTA;
//   * burn-outs: once per ENV STEP the burn-out ticks of the words that hold fire are checked 64 cells at a time
//     (eight 128-bit loads); cells due inside the step go on a burn list with their sub-step;
//   * key schedule (jax.random.split chain of the K sub-steps), clock, move, douse, day/night, reward, done and the
//     info counters are part of the same launch.
#include "gca_common.cuh"

namespace gca {
namespace {

constexpr int BB_THREADS = 256;
constexpr int BB_LIST_CAP = 4096;  // front cells the balanced list holds; larger fronts are walked word by word
constexpr int BB_BURN_CAP = 1024;  // burn-outs of one env step the burn list holds (beyond: per-sub-step scans)
constexpr int BB_PAIR_CAP = BB_THREADS * 8;  // (front cell, burning direction) draws of one round of BB_THREADS front cells
#define BB_LO 0.9998779296875f     /* 1 - 2^-13: (2R+1)^2 <= 441 terms -> |err| <= 441 u |sum| */
#define BB_HI 1.0001220703125f     /* 1 + 2^-13 */

struct BbScalars {
  uint32_t sched[GCA_MAX_K][12];  // per sub-step: Sburn[2] Sgrow[2] ak1[2] ak2[2] wind change step pad
  int nfront;
  int npairs[2];                  // draws listed in the current round of front cells (double-buffered by round parity)
  unsigned long long rowfire[4];  // bit r: row r may hold a burning cell (set when fire is seen / ignites, never cleared in a step)
  int nburn;                      // entries on the burn list (> BB_BURN_CAP: overflow, the list is not used)
  int cnt_tree, cnt_fire;
  unsigned int n_draws, n_ign, n_ext, n_thresh, n_front;
};

// bits [c0, c0 + nbits) of a row of WW 64-bit words (bit c & 63 of word c >> 6); columns outside the row read 0
__device__ __forceinline__ uint32_t row_field(const unsigned long long* row, int WW, int c0, int nbits) {
  const int w = c0 >> 6;  // floor, also for negative c0
  const int s = c0 & 63;
  const unsigned long long lo = (w >= 0 && w < WW) ? row[w] : 0ull;
  const unsigned long long hi = (w + 1 >= 0 && w + 1 < WW) ? row[w + 1] : 0ull;
  const unsigned long long v = s ? ((lo >> s) | (hi << (64 - s))) : lo;
  return (uint32_t)v & ((1u << nbits) - 1u);
}

// 64 cells (codes 0 / 1 / 2) -> tree and fire masks
__device__ __forceinline__ void pack64(const uint8_t* src, unsigned long long& t, unsigned long long& f) {
  const uint4* p = reinterpret_cast<const uint4*>(src);
  t = 0ull; f = 0ull;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 v = p[q];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // bytes -> 4 bits: the bit 0 (tree) / bit 1 (fire) of the four bytes gathered by one multiply
      const uint32_t tn = (((w[k] & 0x01010101u) * 0x01020408u) >> 24) & 15u;
      const uint32_t fn = ((((w[k] >> 1) & 0x01010101u) * 0x01020408u) >> 24) & 15u;
      t |= (unsigned long long)tn << (16 * q + 4 * k);
      f |= (unsigned long long)fn << (16 * q + 4 * k);
    }
  }
}
// tree / fire masks -> 64 cell codes
__device__ __forceinline__ void unpack64(uint8_t* dst, unsigned long long t, unsigned long long f) {
  uint4* p = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t tn = (uint32_t)(t >> (16 * q + 4 * k)) & 15u, fn = (uint32_t)(f >> (16 * q + 4 * k)) & 15u;
      // 4 bits -> 4 bytes (bit i -> byte i)
      w[k] = ((tn * 0x00204081u) & 0x01010101u) | (((fn * 0x00204081u) & 0x01010101u) << 1);
    }
    p[q] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// jax.random key schedule of the K sub-steps (ca_alexandridis_jax.py:436-448, :352-368) by ONE warp: the chain
// K0 -> K1 -> K2 -> K3 per sub-step is walked by all lane pairs redundantly (3 dependent splits per sub-step), then
// pair 2j derives Sburn / Sgrow / the randint keys of sub-step j and pair 2j+1 its wind draws, in parallel.
__device__ void bb_key_schedule(BbScalars& sc, const gca_params& P, const gca_state& S, const gca_inject& J, int e, int lane) {
  const int K = P.K, mode = P.rng_mode, N = S.N;
  uint32_t k0 = S.key[2 * e], k1 = S.key[2 * e + 1];
  const int pair = lane >> 1;
  const uint32_t w = lane & 1;
  uint32_t my_s1[2] = {0, 0}, my_sw[2] = {0, 0}, my_si[2] = {0, 0};
  for (int j = 0; j < K; ++j) {
    uint32_t n0, n1, s0, s1;
    split_pair(k0, k1, mode, lane, n0, n1, s0, s1);  // K1, S1
    if (pair == 2 * j) { my_s1[0] = s0; my_s1[1] = s1; }
    k0 = n0; k1 = n1;
    split_pair(k0, k1, mode, lane, n0, n1, s0, s1);  // K2, Swind
    if (pair == 2 * j + 1) { my_sw[0] = s0; my_sw[1] = s1; }
    k0 = n0; k1 = n1;
    split_pair(k0, k1, mode, lane, n0, n1, s0, s1);  // K3, Sidx
    if (pair == 2 * j + 1) { my_si[0] = s0; my_si[1] = s1; }
    k0 = n0; k1 = n1;
  }
  if (lane == 0) { S.key[2 * e] = k0; S.key[2 * e + 1] = k1; }
  const bool burn_role = (pair & 1) == 0;
  const int j = pair >> 1;
  uint32_t n0, n1, s0, s1;
  // level 1: burn: split(S1) -> Ka, Sburn ; wind: split(Sidx) -> wk1, wk2
  split_pair(burn_role ? my_s1[0] : my_si[0], burn_role ? my_s1[1] : my_si[1], mode, lane, n0, n1, s0, s1);
  const uint32_t sburn0 = s0, sburn1 = s1;  // (wind role: wk2)
  const uint32_t c0 = n0, c1 = n1;          // burn: Ka ; wind: wk1
  // level 2: burn: split(Ka) -> Kb, Sgrow ; wind: even lane bits(wk1, ()), odd lane bits(wk2, ())
  uint32_t sgrow0 = 0, sgrow1 = 0, kb0 = 0, kb1 = 0, hb = 0, lb = 0;
  {
    split_pair(c0, c1, mode, lane, kb0, kb1, sgrow0, sgrow1);  // (meaningful for the burn role)
    const uint32_t mine = bits_scalar(tf_key(w ? sburn0 : c0, w ? sburn1 : c1), mode);
    const uint32_t other = __shfl_xor_sync(GCA_FULL, mine, 1);
    hb = w ? other : mine;
    lb = w ? mine : other;
  }
  // level 3: burn: split(Kb) -> Kc, Sage ; wind: bits(Swind, ())
  uint32_t sage0, sage1;
  split_pair(kb0, kb1, mode, lane, n0, n1, sage0, sage1);
  const uint32_t uw_bits = bits_scalar(tf_key(my_sw[0], my_sw[1]), mode);
  // level 4: burn: split(Sage) -> ak1, ak2
  uint32_t a10, a11, a20, a21;
  split_pair(sage0, sage1, mode, lane, a10, a11, a20, a21);
  if (j < K && w == 0) {
    uint32_t* row = sc.sched[j];
    if (burn_role) {
      row[0] = sburn0; row[1] = sburn1; row[2] = sgrow0; row[3] = sgrow1;
      row[4] = a10; row[5] = a11; row[6] = a20; row[7] = a21;
    } else {
      float u = bits_to_uniform(uw_bits);
      int step = randint_from_bits(hb, lb, 1, 7u, 4u);
      if (J.u_wind) u = J.u_wind[(size_t)j * N + e];
      if (J.wind_step) step = J.wind_step[(size_t)j * N + e];
      row[9] = (u < P.p_wind_change) ? 1u : 0u;
      row[10] = (uint32_t)step;
    }
  }
  __syncwarp();
  if (lane == 0) {
    int wi = S.wind_index[e];
    for (int q = 0; q < K; ++q) {
      sc.sched[q][8] = (uint32_t)wi;  // wind used by sub-step q
      if (sc.sched[q][9]) wi = (wi + (int)sc.sched[q][10]) % 8;
    }
    S.wind_index[e] = wi;
  }
}

struct BbView {
  unsigned long long *tree, *fire, *dous, *ign;
  uint16_t* list;
  uint32_t* burn;   // cells that burn out during this env step: (row << 8 | col) << 3 | sub-step
  float2* bnd;      // [BB_THREADS] enclosure (lo, hi) of the burn probability's direction-independent part, per front cell of the round
  uint16_t* pairs;  // [BB_PAIR_CAP] draws of the round: (front cell of the round << 4) | direction 0..8
  BbScalars* sc;
};

// Front cell (r, c): enclosure [blo, bhi] of (heat - dousing)(1 + p_veg)(1 + p_den) from ring populations, and which of
// its Moore neighbours burn (bit d = 3 i + j of the result; 0 when the cell cannot ignite).  `fld_out` / `dwin_out`
// (optional): the fire rows of the window and the 5x5 doused window, for the exact re-evaluation.
template <int R>
__device__ __forceinline__ uint32_t bb_front_bounds(const BbView& v, const gca_params& P, const gca_state& S, int e, int r,
                                                    int c, float& blo, float& bhi, float& a, float& b, uint32_t* fld,
                                                    uint32_t& dwin) {
  const int H = P.H, W = P.W, WW = W >> 6;
  const size_t env_off = (size_t)e * H * W;
  const size_t gcell = (size_t)r * W + c;
  constexpr int WIN = 2 * R + 1;
#pragma unroll
  for (int di = 0; di < WIN; ++di) {
    const int rr = r - R + di;
    fld[di] = (rr >= 0 && rr < H) ? row_field(v.fire + rr * WW, WW, c - R, WIN) : 0u;
  }
  // ring populations: C_k = burning cells within Chebyshev distance k
  float Hf = 0.0f;
  int prev = 0;
#pragma unroll
  for (int k = 1; k <= R; ++k) {
    const uint32_t mask = ((1u << (2 * k + 1)) - 1u) << (R - k);
    int cnt = 0;
#pragma unroll
    for (int di = R - k; di <= R + k; ++di) cnt += __popc(fld[di] & mask);
    Hf = fmaf((float)(cnt - prev), P.ring_w[k], Hf);
    prev = cnt;
  }
  dwin = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int rr = r - 2 + i;
    if (rr >= 0 && rr < H) dwin |= row_field(v.dous + rr * WW, WW, c - 2, 5) << (5 * i);
  }
  float Dlo = 0.0f, Dhi = 0.0f;
  if (dwin) {
    const int ni = __popc(dwin & ((0x0Eu << 5) | (0x0Eu << 10) | (0x0Eu << 15)));
    const int nb = __popc(dwin) - ni;
    const float Df = fmaf((float)nb, P.dous_border, (float)ni * P.dous_inner);
    Dlo = __fmul_rn(Df, BB_LO);
    Dhi = __fmul_rn(Df, BB_HI);
  }
  int hid = 3 | (3 << 3);
  if (S.hidden != nullptr) hid = S.hidden[env_off + gcell];
  a = P.onep_veg[clip15(hid & 7)];
  b = P.onep_den[clip15((hid >> 3) & 7)];
  blo = __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(Hf, BB_LO), Dhi), a), b);
  bhi = __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(Hf, BB_HI), Dlo), a), b);
  if (!(bhi > 0.0f)) return 0u;
  // burning Moore neighbours: bits R-1 .. R+1 of the three middle rows
  return (((fld[R - 1] >> (R - 1)) & 7u) | (((fld[R] >> (R - 1)) & 7u) << 3) | (((fld[R + 1] >> (R - 1)) & 7u) << 6)) & ~16u;
}

// reference-order (row-major, float32, from +0) value of (heat - dousing)(1 + p_veg)(1 + p_den): threshold cells
template <int R>
__device__ __forceinline__ float bb_exact_from(const gca_params& P, const uint32_t* fld, uint32_t dwin, float a, float b) {
  constexpr int WIN = 2 * R + 1;
  float heat = 0.0f;
#pragma unroll 1
  for (int di = 0; di < WIN; ++di) {
    uint32_t m = fld[di];
    const int a_di = di < R ? R - di : di - R;
    while (m) {
      const int dj = __ffs((int)m) - 1;
      m &= m - 1;
      const int a_dj = dj < R ? R - dj : dj - R;
      heat = __fadd_rn(heat, P.ring_w[max(a_di, a_dj)]);
    }
  }
  float dous = 0.0f;
  for (int q = 0; q < 25; ++q)
    if ((dwin >> q) & 1u) {
      const int qi = q / 5, qj = q % 5;
      const bool inner = qi >= 1 && qi <= 3 && qj >= 1 && qj <= 3;
      dous = __fadd_rn(dous, inner ? P.dous_inner : P.dous_border);
    }
  return __fmul_rn(__fmul_rn(__fsub_rn(heat, dous), a), b);
}
// ... recomputed from the bit-boards for one cell (the pooled draws keep only the bounds per cell)
template <int R>
__device__ __noinline__ float bb_exact_base(const BbView& v, const gca_params& P, const gca_state& S, int e, int r, int c) {
  uint32_t fld[2 * R + 1], dwin;
  float blo, bhi, a, b;
  bb_front_bounds<R>(v, P, S, e, r, c, blo, bhi, a, b, fld, dwin);
  return bb_exact_from<R>(P, fld, dwin, a, b);
}

// the uniform of draw (cell, direction d) of sub-step j
__device__ __forceinline__ float bb_draw(const gca_params& P, const gca_state& S, const gca_inject& J, int e, int j,
                                         size_t gcell, int d, const TfKey& kburn, uint32_t half_burn) {
  if (J.u_burn) return J.u_burn[(((size_t)j * S.N + e) * P.H * P.W + gcell) * 9 + d];
  return bits_to_uniform(bits_at(kburn, (uint32_t)(gcell * 9 + d), half_burn, P.rng_mode));
}

// One front cell (r, c) of sub-step j by one thread (fronts larger than the list): enclosure, one draw per burning
// direction, exact re-evaluation of undecided draws; an ignition sets the cell's bit in v.ign.
template <int R>
__device__ __forceinline__ void bb_front_cell(const BbView& v, const gca_params& P, const gca_state& S, const gca_inject& J,
                                              int e, int j, int r, int c, const TfKey& kburn, const float* wind,
                                              uint32_t half_burn, uint32_t& n_draws, uint32_t& n_thresh) {
  const int W = P.W, WW = W >> 6;
  const size_t env_off = (size_t)e * P.H * W;
  const size_t gcell = (size_t)r * W + c;
  uint32_t fld[2 * R + 1], dwin;
  float blo, bhi, a, b;
  const uint32_t nb = bb_front_bounds<R>(v, P, S, e, r, c, blo, bhi, a, b, fld, dwin);
  bool ig = false, have_exact = false;
  float base_exact = 0.0f;
#pragma unroll 1
  for (int d = 0; d < 9 && !ig; ++d) {
    if (!((nb >> d) & 1u)) continue;
    const float u = bb_draw(P, S, J, e, j, gcell, d, kburn, half_burn);
    ++n_draws;
    const float w = wind[d];
    const float s = S.pslope ? S.pslope[(env_off + gcell) * 8 + dir_slot(d)] : 1.0f;
    const float plo = __fmul_rn(__fmul_rn(blo, w), s), phi = __fmul_rn(__fmul_rn(bhi, w), s);
    if (u < plo) { ig = true; break; }
    if (u < phi) {
      if (!have_exact) {
        base_exact = bb_exact_from<R>(P, fld, dwin, a, b);
        have_exact = true;
        ++n_thresh;
      }
      if (u < __fmul_rn(__fmul_rn(base_exact, w), s)) ig = true;
    }
  }
  if (ig) atomicOr(v.ign + r * WW + (c >> 6), 1ull << (c & 63));
}

#ifndef GCA_BB_MINB
#define GCA_BB_MINB 4  /* 64 registers: measured 242 us per env step at 1024 envs of 256x256, K = 4 (325 us at 2 CTAs per SM, 277 at 3) */
#endif
template <int R>
__global__ void __launch_bounds__(BB_THREADS, GCA_BB_MINB)
env_step_bb_kernel(const __grid_constant__ gca_params P, const __grid_constant__ gca_state S,
                   const int32_t* __restrict__ actions, const __grid_constant__ gca_step_out O,
                   const __grid_constant__ gca_inject J, uint32_t flags) {
  extern __shared__ __align__(16) unsigned char bb_smem_raw[];
  const int H = P.H, W = P.W, WW = W >> 6, HW = H * WW, K = P.K, mode = P.rng_mode;
  BbView v;
  v.tree = reinterpret_cast<unsigned long long*>(bb_smem_raw);
  v.fire = v.tree + HW;
  v.dous = v.fire + HW;
  v.ign = v.dous + HW;
  v.list = reinterpret_cast<uint16_t*>(v.ign + HW);
  v.burn = reinterpret_cast<uint32_t*>(v.list + BB_LIST_CAP);
  v.bnd = reinterpret_cast<float2*>(v.burn + BB_BURN_CAP);
  v.pairs = reinterpret_cast<uint16_t*>(v.bnd + BB_THREADS);
  v.sc = reinterpret_cast<BbScalars*>(v.pairs + BB_PAIR_CAP);
  const uint32_t inv_ww = (65536u + (uint32_t)WW - 1u) / (uint32_t)WW;  // i / WW = (i * inv_ww) >> 16 for i < 1024
  BbScalars& sc = *v.sc;
  const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const size_t env_off = (size_t)e * H * W;
  const uint32_t tick0 = S.tick[e];
  const uint32_t half_cell = (uint32_t)(((size_t)H * W) >> 1);
  const uint32_t half_burn = (uint32_t)((9ull * H * W) >> 1);

  if (tid < 4) sc.rowfire[tid] = 0ull;
  if (tid == 0) {
    sc.cnt_tree = 0; sc.cnt_fire = 0; sc.nburn = 0; sc.npairs[0] = 0; sc.npairs[1] = 0;
    sc.n_draws = 0; sc.n_ign = 0; sc.n_ext = 0; sc.n_thresh = 0; sc.n_front = 0;
  }
  __syncthreads();  // (the row flags are clear before the first word with fire sets one)
  // ---- grid -> bit-boards (warp 0 walks the key chains meanwhile) --------------------------------------------------
  if (tid < 32) bb_key_schedule(sc, P, S, J, e, lane);
  for (int i = tid; i < HW; i += BB_THREADS) {
    unsigned long long t, f;
    pack64(S.cell + env_off + (size_t)i * 64, t, f);
    v.tree[i] = t;
    v.fire[i] = f;
    if (f) atomicOr(&sc.rowfire[(((uint32_t)i * inv_ww) >> 16) >> 6], 1ull << ((((uint32_t)i * inv_ww) >> 16) & 63u));
    v.dous[i] = reinterpret_cast<const unsigned long long*>(S.doused)[(size_t)e * HW + i];
    v.ign[i] = 0ull;
  }
  __syncthreads();

  // ---- burn-outs of the whole env step, found ONCE: the burn-out ticks of the words that hold fire are read 64 cells at
  //      a time; a burning cell whose tick falls inside [tick0, tick0 + K) goes on the burn list with its sub-step and its
  //      tick is cleared (fire_age ends at 0; nothing reads it before).  A list that overflows is dropped: the sub-steps
  //      then scan for their own burn-outs.
  for (int i = tid; i < HW; i += BB_THREADS) {
    const unsigned long long f_old = v.fire[i];
    if (!f_old) continue;
    const int r = (int)(((uint32_t)i * inv_ww) >> 16), w = i - r * WW;
    uint4* dp = reinterpret_cast<uint4*>(S.death + env_off + (size_t)i * 64);
    uint4 dv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) dv[q] = dp[q];
    unsigned long long due = 0ull;
    unsigned long long w0 = 0ull, w1 = 0ull, w2 = 0ull;  // bit planes of the sub-step (K <= 8)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t ww[4] = {dv[q].x, dv[q].y, dv[q].z, dv[q].w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int b = 8 * q + k;
        const uint32_t d = (ww[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
        const uint32_t rr = (d - tick0) & 0xFFFFu;
        const unsigned long long hit = (((f_old >> b) & 1ull) && rr < (uint32_t)K) ? 1ull : 0ull;
        due |= hit << b;
        w0 |= (hit & (unsigned long long)(rr & 1u)) << b;
        w1 |= (hit & (unsigned long long)((rr >> 1) & 1u)) << b;
        w2 |= (hit & (unsigned long long)((rr >> 2) & 1u)) << b;
      }
    }
    if (due) {
      const int n = __popcll(due);
      const int base = atomicAdd(&sc.nburn, n);
      int at = base;
      unsigned long long m = due;
      while (m) {
        const int b = __ffsll((long long)m) - 1;
        m &= m - 1;
        const uint32_t sub = (uint32_t)((w0 >> b) & 1ull) | ((uint32_t)((w1 >> b) & 1ull) << 1) | ((uint32_t)((w2 >> b) & 1ull) << 2);
        if (at < BB_BURN_CAP) v.burn[at] = ((((uint32_t)r << 8) | (uint32_t)(w * 64 + b)) << 3) | sub;
        ++at;
      }
    }
  }
  __syncthreads();
  const int nburn = sc.nburn;
  const bool burn_listed = nburn <= BB_BURN_CAP;
  if (burn_listed) {
    // the listed cells' ticks end at 0 (scattered 2-byte stores: a handful per env step)
    for (int i = tid; i < nburn; i += BB_THREADS) {
      const uint32_t cellrc = v.burn[i] >> 3;
      S.death[env_off + (size_t)(cellrc >> 8) * W + (cellrc & 255u)] = 0;
    }
  }

  uint32_t n_draws = 0, n_thresh = 0, n_ign = 0, n_ext = 0;
  for (int j = 0; j < K; ++j) {
    const uint32_t* sr = sc.sched[j];
    const TfKey kburn = tf_key(sr[0], sr[1]);
    const float* wind = P.winds + 9 * (int)sr[8];
    const uint32_t tick = tick0 + (uint32_t)j;
    if (tid == 0) { sc.nfront = 0; sc.npairs[0] = 0; }  // (the rounds of a sub-step start with counter 0)
    __syncthreads();
    // ---- front = tree AND dilate(fire) ------------------------------------------------------------------------------
    constexpr int MAXW = (65536 / 64 + BB_THREADS - 1) / BB_THREADS;  // words per thread
    unsigned long long fr[MAXW];
    int mine = 0;
#pragma unroll
    for (int q = 0; q < MAXW; ++q) {
      const int i = tid + q * BB_THREADS;
      fr[q] = 0ull;
      if (i < HW) {
        const int r = (int)(((uint32_t)i * inv_ww) >> 16), w = i - r * WW;
        // rows r-1 .. r+1 without a burning cell (most of a 256x256 grid): nothing to dilate
        const int rl = r > 0 ? r - 1 : 0, rh = r + 1 < H ? r + 1 : r;
        const unsigned long long near = ((sc.rowfire[rl >> 6] >> (rl & 63)) | (sc.rowfire[r >> 6] >> (r & 63)) |
                                         (sc.rowfire[rh >> 6] >> (rh & 63))) & 1ull;
        if (!near) continue;
        unsigned long long dil = 0ull;
#pragma unroll
        for (int dr = -1; dr <= 1; ++dr) {
          const int rr = r + dr;
          if (rr < 0 || rr >= H) continue;
          const unsigned long long* row = v.fire + rr * WW;
          const unsigned long long f = row[w];
          dil |= f | (f << 1) | (f >> 1);
          if (w > 0) dil |= row[w - 1] >> 63;
          if (w + 1 < WW) dil |= row[w + 1] << 63;
        }
        fr[q] = v.tree[i] & dil;
        mine += __popcll(fr[q]);
      }
    }
    int base = 0;
    if (mine) base = atomicAdd(&sc.nfront, mine);
    __syncthreads();
    const int nfront = sc.nfront;
    if (nfront <= BB_LIST_CAP) {
      // balanced: compact the front into a CTA-wide list, one cell per thread and round
#pragma unroll
      for (int q = 0; q < MAXW; ++q) {
        const int i = tid + q * BB_THREADS;
        unsigned long long m = fr[q];
        if (m) {
          const int r = (int)(((uint32_t)i * inv_ww) >> 16), w = i - r * WW;
          while (m) {
            const int b = __ffsll((long long)m) - 1;
            m &= m - 1;
            v.list[base++] = (uint16_t)((r << 8) | (w * 64 + b));
          }
        }
      }
      __syncthreads();
      // rounds of BB_THREADS front cells: a thread per cell computes the enclosure and lists the cell's burning
      // directions; the (cell, direction) draws are then dealt one per thread -- a cell with five burning neighbours
      // no longer makes its warp wait for five threefry blocks in a row
      int par = 0;
      for (int base = 0; base < nfront; base += BB_THREADS, par ^= 1) {
        {
          const int i = base + tid;
          uint32_t nb = 0u;
          float blo = 0.0f, bhi = 0.0f;
          if (i < nfront) {
            const int cellrc = v.list[i];
            uint32_t fld[2 * R + 1], dwin;
            float a, b;
            nb = bb_front_bounds<R>(v, P, S, e, cellrc >> 8, cellrc & 255, blo, bhi, a, b, fld, dwin);
          }
          v.bnd[tid] = make_float2(blo, bhi);
          if (nb) {
            int at = atomicAdd(&sc.npairs[par], __popc(nb));
            while (nb) {
              const uint32_t d = (uint32_t)__ffs((int)nb) - 1u;
              nb &= nb - 1u;
              v.pairs[at++] = (uint16_t)(((uint32_t)tid << 4) | d);
            }
          }
        }
        __syncthreads();
        const int npairs = sc.npairs[par];
        if (tid == 0) sc.npairs[par ^ 1] = 0;  // the next round's counter (its last readers passed the barrier above)
        for (int q = tid; q < npairs; q += BB_THREADS) {
          const uint32_t rec = v.pairs[q];
          const int t = (int)(rec >> 4), d = (int)(rec & 15u);
          const int cellrc = v.list[base + t];
          const int r = cellrc >> 8, c = cellrc & 255;
          const size_t gcell = (size_t)r * W + c;
          const float2 bd = v.bnd[t];
          const float s = S.pslope ? S.pslope[(env_off + gcell) * 8 + dir_slot(d)] : 1.0f;
          const float u = bb_draw(P, S, J, e, j, gcell, d, kburn, half_burn);
          ++n_draws;
          const float w = wind[d];
          const float plo = __fmul_rn(__fmul_rn(bd.x, w), s), phi = __fmul_rn(__fmul_rn(bd.y, w), s);
          bool ig = u < plo;
          if (!ig && u < phi) {  // threshold cell: reference-order sums
            ++n_thresh;
            ig = u < __fmul_rn(__fmul_rn(bb_exact_base<R>(v, P, S, e, r, c), w), s);
          }
          if (ig) atomicOr(v.ign + r * WW + (c >> 6), 1ull << (c & 63));
        }
        __syncthreads();
      }
    } else {
      // a front larger than the list (dense fires): every thread walks its own words
#pragma unroll 1
      for (int q = 0; q < MAXW; ++q) {
        const int i = tid + q * BB_THREADS;
        unsigned long long m = fr[q];
        const int r = (int)(((uint32_t)i * inv_ww) >> 16), w = i - r * WW;
        while (m) {
          const int b = __ffsll((long long)m) - 1;
          m &= m - 1;
          bb_front_cell<R>(v, P, S, J, e, j, r, w * 64 + b, kburn, wind, half_burn, n_draws, n_thresh);
        }
      }
    }
    if (tid == 0) sc.n_front += (unsigned)nfront;
    __syncthreads();
    // ---- apply: ignitions (fire-age draws), burn-outs, regrowth -----------------------------------------------------
    if (burn_listed) {
      // this sub-step's burn-outs join the ignition bits in v.ign (ignitions are tree cells, burn-outs fire cells)
      for (int i = tid; i < nburn; i += BB_THREADS) {
        const uint32_t ent = v.burn[i];
        if ((int)(ent & 7u) != j) continue;
        const uint32_t cellrc = ent >> 3, col = cellrc & 255u;
        atomicOr(v.ign + (cellrc >> 8) * WW + (col >> 6), 1ull << (col & 63));
      }
      __syncthreads();
    }
    const TfKey ka1 = tf_key(sr[4], sr[5]), ka2 = tf_key(sr[6], sr[7]), kg = tf_key(sr[2], sr[3]);
    for (int i = tid; i < HW; i += BB_THREADS) {
      const int r = (int)(((uint32_t)i * inv_ww) >> 16), w = i - r * WW;
      const unsigned long long t_old = v.tree[i], f_old = v.fire[i];
      const unsigned long long marks = v.ign[i];
      const unsigned long long I = marks & t_old;
      unsigned long long ext = marks & f_old, grow = 0ull;
      if (marks) v.ign[i] = 0ull;
      if (f_old && !burn_listed) {
        // (burn list overflow) burn-out ticks of the word's 64 cells: a burning cell whose tick is due goes out
        uint4* dp = reinterpret_cast<uint4*>(S.death + env_off + (size_t)i * 64);
        uint4 dv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) dv[q] = dp[q];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint32_t ww[4] = {dv[q].x, dv[q].y, dv[q].z, dv[q].w};
          bool touched = false;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int b = 8 * q + k;
            const uint32_t d = (ww[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
            if (((f_old >> b) & 1ull) && ((d - tick) & 0xFFFFu) == 0u) {
              ext |= 1ull << b;
              ww[k >> 1] &= ~(0xFFFFu << (16 * (k & 1)));
              touched = true;
            }
          }
          if (touched) dp[q] = make_uint4(ww[0], ww[1], ww[2], ww[3]);
        }
      }
      if (I) {
        atomicOr(&sc.rowfire[r >> 6], 1ull << (r & 63));
        unsigned long long m = I;
        while (m) {
          const int b = __ffsll((long long)m) - 1;
          m &= m - 1;
          const size_t gcell = (size_t)r * W + w * 64 + b;
          int age;
          if (J.age_new) age = J.age_new[((size_t)j * S.N + e) * H * W + gcell];
          else age = randint_from_bits(bits_at(ka1, (uint32_t)gcell, half_cell, mode),
                                       bits_at(ka2, (uint32_t)gcell, half_cell, mode), P.age_lo, P.age_span, P.age_mult);
          S.death[env_off + gcell] = (uint16_t)(tick + (uint32_t)age);  // burns out at tick + age
        }
        n_ign += __popcll(I);
      }
      if (P.p_tree > 0.0f) {
        unsigned long long m = ~(t_old | f_old);
        while (m) {
          const int b = __ffsll((long long)m) - 1;
          m &= m - 1;
          const size_t gcell = (size_t)r * W + w * 64 + b;
          float u;
          if (J.u_grow) u = J.u_grow[((size_t)j * S.N + e) * H * W + gcell];
          else u = bits_to_uniform(bits_at(kg, (uint32_t)gcell, half_cell, mode));
          if (u < P.p_tree) grow |= 1ull << b;
        }
      }
      n_ext += __popcll(ext);
      v.tree[i] = (t_old & ~I) | grow;
      v.fire[i] = (f_old & ~ext) | I;
    }
    __syncthreads();
  }

  // ---- write the grid back, counts ------------------------------------------------------------------------------------
  int nt = 0, nf = 0;
  for (int i = tid; i < HW; i += BB_THREADS) {
    const unsigned long long t = v.tree[i], f = v.fire[i];
    unpack64(S.cell + env_off + (size_t)i * 64, t, f);
    nt += __popcll(t);
    nf += __popcll(f);
  }
  nt = __reduce_add_sync(GCA_FULL, nt);
  nf = __reduce_add_sync(GCA_FULL, nf);
  n_draws = __reduce_add_sync(GCA_FULL, n_draws);
  n_thresh = __reduce_add_sync(GCA_FULL, n_thresh);
  n_ign = __reduce_add_sync(GCA_FULL, n_ign);
  n_ext = __reduce_add_sync(GCA_FULL, n_ext);
  if (lane == 0) {
    atomicAdd(&sc.cnt_tree, nt);
    atomicAdd(&sc.cnt_fire, nf);
    if (n_draws) atomicAdd(&sc.n_draws, n_draws);
    if (n_thresh) atomicAdd(&sc.n_thresh, n_thresh);
    if (n_ign) atomicAdd(&sc.n_ign, n_ign);
    if (n_ext) atomicAdd(&sc.n_ext, n_ext);
  }
  __syncthreads();
  if (tid != 0) return;
  // ---- per-env scalars: tick, clock, move, douse, day/night, reward, done, info counters -------------------------------
  S.tick[e] = tick0 + (uint32_t)K;
  const int t = sc.cnt_tree, f = sc.cnt_fire;
  const float rew = award(t, f);
  const bool done = f == 0;
  if (!(flags & GCA_FLAG_CA_ONLY)) {
    const int a0 = actions[3 * e], a1 = actions[3 * e + 1];
    const int a0c = min(max(a0, 0), 8), a1c = min(max(a1, 0), 1);
    const float tt = __fadd_rn(__fadd_rn(P.t_move[a0c], P.t_shoot[a1c]), P.t_any);
    const float ntm = __fadd_rn(S.time[e], tt);
    S.time[e] = __fsub_rn(ntm, truncf(ntm));
    int row = S.position[2 * e], col = S.position[2 * e + 1];
    move_position(a0, H, W, row, col);
    S.position[2 * e] = row;
    S.position[2 * e + 1] = col;
    if (a1 == 1) S.doused[((size_t)e * H + row) * WW + (col >> 6)] |= 1ull << (col & 63);
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
  if (O.stats != nullptr) {
    atomicAdd(&O.stats[0], (unsigned long long)sc.n_front);
    atomicAdd(&O.stats[1], (unsigned long long)sc.n_draws);
    atomicAdd(&O.stats[2], (unsigned long long)sc.n_ign);
    atomicAdd(&O.stats[3], (unsigned long long)sc.n_ext);
    if (sc.n_thresh) atomicAdd(&O.stats[4], (unsigned long long)sc.n_thresh);
    atomicAdd(&O.stats[5], 1ull);
  }
}

template <int R>
cudaError_t launch_bb_instance(const gca_params& p, const gca_state& s, const int32_t* actions, const gca_step_out& out,
                               const gca_inject& inj, uint32_t flags, size_t smem, cudaStream_t st) {
  static size_t configured = 0;
  auto kern = env_step_bb_kernel<R>;
  if (configured < smem) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    configured = smem;
  }
  kern<<<dim3((unsigned)s.N), dim3(BB_THREADS), smem, st>>>(p, s, actions, out, inj, flags);
  return cudaGetLastError();
}

}  // namespace

// the grids this kernel takes: rows of whole 64-bit words, at most 65536 cells, burn radius 4..6
bool bb_supported(const gca_params& p) {
  return (p.W & 63) == 0 && p.W <= 256 && p.H <= 256 && (long long)p.H * p.W <= 65536 && p.R >= 4 && p.R <= 6 &&
         !(p.H == 64 && p.W == 64);
}

cudaError_t launch_bb_env_step(const gca_params& p, const gca_state& s, const int32_t* actions, const gca_step_out& out,
                               const gca_inject& inj, uint32_t flags, cudaStream_t st) {
  const int HW = p.H * (p.W >> 6);
  const size_t smem = (size_t)HW * 8 * 4 + BB_LIST_CAP * sizeof(uint16_t) + BB_BURN_CAP * sizeof(uint32_t) +
                      BB_THREADS * sizeof(float2) + BB_PAIR_CAP * sizeof(uint16_t) + sizeof(BbScalars) + 16;
  switch (p.R) {
    case 4: return launch_bb_instance<4>(p, s, actions, out, inj, flags, smem, st);
    case 5: return launch_bb_instance<5>(p, s, actions, out, inj, flags, smem, st);
    case 6: return launch_bb_instance<6>(p, s, actions, out, inj, flags, smem, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace gca
