// gca_step64.cu -- fused environment step for 64x64 grids: one CTA of E warps steps E environments
// in lock-step; every warp OWNS one env (its bit-boards live in that warp's registers) and the two
// heavy phases of each CA sub-step -- per-front-cell burn probability and the per-(cell, direction)
// threefry draws -- are POOLED over the CTA, so a large fire front is worked on by all E warps.
//
// Replaces, for a batch of envs, jax.vmap(MDP.update) + _award + _is_done (+ conditional_reset)
// of /root/reference/gym_cellular_automata/forest_fire/bulldozer/advanced_bulldozer.py:332-518,
// 1103-1133 and the operators it calls (ca_alexandridis_jax.py:321-460, repeat_ca_jax.py:34-71,
// move_modify_jax.py:39-157).  Not a translation: the reference materialises (H,W,9,9) gathers and
// draws 12 random words per cell; this kernel works on bit-boards and draws lazily.
//
// Design:
//  * a 64-cell row is one 64-bit word per mask (tree, fire, doused).  Lane l of the owner warp holds
//    rows 2l, 2l+1 in registers; vertical halos come from warp shuffles.
//  * all K CA sub-steps of the env step are applied on-chip (temporal blocking); HBM sees one
//    coalesced read of the env's tree / fire bit-boards (1 KB, the packed twin of the u8 grid) and
//    sparse in-place writes of the cells (u8 grid) and bit-board rows that changed.
//  * front cells (tree with a burning Moore neighbour) are compacted into a per-env shared-memory
//    list (built once per env step, extended incrementally).  Work items of the pooled phase are
//    32-entry chunks of those lists, handed out by a shared-memory counter to whichever warp is
//    free: per front cell the 9x9 fire window is cut out of the bit-board, ring populations give an
//    upper bound of the float32 burn-probability chain; the (cell, burning direction) pairs go to
//    the warp's private pair buffer, and whenever it holds 32 pairs the warp draws them -- one
//    counter-based threefry2x32 block per pair reproduces exactly the uniform jax.random would have
//    drawn for that element.  u >= hi does not ignite, u < hi (1 - 2^-14) ignites, and the (rare)
//    in-between case is re-evaluated with the reference's exact row-major float32 summation
//    ("threshold cells").  (The first design gave every env to one warp alone; the kernel then lasted
//    as long as its heaviest env)
//  * fire ages are stored as burn-out ticks, so burning cells need no per-step decrement; a
//    per-row minimum tells which rows hold a cell that burns out in this step.
//  * the key chain of jax.random.split is evaluated by lane pairs; clock, move, douse, day/night,
//    reward (popc + warp reduce), done and the optional auto-reset are fused in the epilogue.
#include <cstddef>

#include "gca_common.cuh"

namespace gca {
namespace {

constexpr int S64_E = GCA_S64_WARPS;   // envs (= warps) per CTA
constexpr int S64_CAP = 256;           // front cells per pass
constexpr int S64_G = GCA_S64_GROUPS;  // independent lock-step groups inside a CTA (own barriers, own work-item counter)
constexpr int S64_GE = S64_E / S64_G;  // envs (= warps) per group
static_assert(S64_E % S64_G == 0 && S64_G >= 1 && S64_G <= 4, "groups must divide the CTA");
constexpr int S64_CHAIN_WARPS = (S64_E + 15) / 16;  // warps that walk the key chains (16 envs each)
// How an owner waits for its env's key chain: a named barrier per owner (ids 1 .. S64_E - 1: the chain warp arrives, the
// owner syncs -- a hardware wait, no polling) when there are enough barriers and no lock-step groups use them; else a
// shared-memory flag the owner polls (measured: the polling loop was 4 % of the kernel's instructions).
constexpr bool S64_CHAIN_BAR = S64_E <= 15 && S64_G == 1;
constexpr int S64_IGN_CAP = 160;       // deferred fire-age draws buffered per env (flushed early when full)
constexpr int S64_WP = 288;            // warp-private pair buffer: < 32 carried over + <= 256 of one chunk
constexpr uint32_t S64_HALF_BURN = 9u * 4096u / 2u;
constexpr uint32_t S64_HALF_CELL = 4096u / 2u;
// enclosure of the fast float32 path: |sequential sum - ring-count sum| <= 89 u |sum|
// (80 adds + 8 flops, u = 2^-24); 2^-16 = 256 u leaves a 2.8x margin.  The upper bound hi is
// chain(H (1 + 2^-16)); chain(H (1 - 2^-16)) >= hi (1 - 2^-15 (1 + 2^-6)) > hi (1 - 2^-14), so
// u < hi (1 - 2^-14) decides "ignites" without storing a second bound per cell.
#define S64_LO 0.9999847412109375f   /* 1 - 2^-16 */
#define S64_HI 1.0000152587890625f   /* 1 + 2^-16 */
#define S64_SURE 0.99993896484375f   /* 1 - 2^-14 */

struct __align__(16) EnvSmem {
  uint32_t fire32[72 * 4];          // fire rows -4..67, 4 overlapping 32-bit views per row
  unsigned long long dous64[68];    // doused rows -2..65 (sparse: the 5x5 window is cut out of the 64-bit rows)
  unsigned long long ign[64];       // ignition accumulator of the sub-step
  unsigned long long burn[64][4];   // rows with burn-outs in this env step: mask, sub-step bit planes 0..2
  unsigned long long listed[64];    // cells that are (or were) on the front list
  uint16_t list[S64_CAP];           // front cells: (row << 6) | col
  uint16_t ignlist[S64_IGN_CAP];    // cells ignited so far in this env step: (sub-step << 12) | cell; their fire-age
                                    //   draws are made in one batch at the end of the step (flush_ages)
  uint32_t sched[GCA_MAX_K][12];    // per sub-step: Sburn[2] Sgrow[2] ak1[2] ak2[2] wind change step pad
  uint4 hot;                        // current sub-step: Sburn k0, k1, k0^k1^C ; env index
  float wind[12];                   // current sub-step: wind matrix (9 used)
  int cnt;                          // list entries of the current pass
  uint32_t dous_even, dous_odd;     // bit l: some doused cell within 2 rows of row 2l / 2l+1
  int nign;                         // entries on ignlist
  int chain_done;                   // set (release) by the chain warp when sched rows 0..5 / hot.x, hot.y of this env are written
  uint32_t tick0;                   // CA tick at the start of this env step (scan items)
  int nscan;                        // rows whose burn-out ticks must be scanned in this env step: their indices are the
                                    //   first nscan BYTES of ignlist (which holds no ignitions before apply(0))
  int pad_;
};
static_assert(S64_E >= 1 && S64_E <= 32 && 28 % S64_E == 0, "envs per CTA: a divisor of 28 (28 warps of 72 registers fill an SM)");
static_assert(sizeof(EnvSmem) % 16 == 0 && offsetof(EnvSmem, burn) % 16 == 0 && offsetof(EnvSmem, hot) % 16 == 0, "128-bit shared accesses");

struct __align__(16) CtaSmem {
  EnvSmem env[S64_E];
  // warp-private buffer of (front cell, burning direction) draws waiting for a full round, of ANY env of the CTA:
  uint32_t pairs[S64_E][S64_WP];    // upper bound of the cell's (p_h (1+p_veg)) (1+p_den) (float32 bits rounded UP to a multiple
                                    //   of 32 ulp) | env slot in the 5 low bits; sign bit: dousing nearby (no cheap lower bound)
  uint16_t pidx[S64_E][S64_WP];     // (cell << 4) | direction 0..8: the pair is compared with element cell * 9 + direction of uniform(Sburn, (H,W,3,3))
  int nch[32];                      // chunks of each env in the current pass
  int next[4];                      // work-item counter of the pooled phase, per group
};

// barriers of one lock-step group (named barrier 1 + group, S64_GE warps); a single group uses barrier 0
__device__ __forceinline__ void group_sync(int group) {
  if (S64_G == 1) { __syncthreads(); return; }
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(S64_GE * 32) : "memory");
}
__device__ __forceinline__ int group_sync_or(int group, bool pred) {
  if (S64_G == 1) return __syncthreads_or(pred);
  int r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %3, 0;\n\tbar.red.or.pred q, %1, %2, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
      : "=r"(r) : "r"(group + 1), "r"(S64_GE * 32), "r"((uint32_t)pred) : "memory");
  return r;
}

// atomicAdd on a shared-memory int by ONE lane: the CUDA builtin compiles to the warp-aggregated pattern (vote, leader
// election, popc, shuffle: ~14 instructions); the plain instruction is what is wanted here
__device__ __forceinline__ int smem_add_ret(int* p, int v) {
  int old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void smem_st_release(int* p, int v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int smem_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}

__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// x % span with a precomputed magic = 0xFFFFFFFF / span (span < 2^16): quotient estimate is at
// most 2 too small
__device__ __forceinline__ uint32_t fastmod(uint32_t x, uint32_t span, uint32_t magic) {
  uint32_t r = x - __umulhi(x, magic) * span;
  if (r >= span) r -= span;
  if (r >= span) r -= span;
  return r;
}

// view k of a row covers columns 16k-4 .. 16k+27 (bit b <-> column 16k-4+b)
__device__ __forceinline__ void store_row_views(uint32_t* dst, unsigned long long x) {
  uint4 w;
  w.x = (uint32_t)(x << 4);
  w.y = (uint32_t)(x >> 12);
  w.z = (uint32_t)(x >> 28);
  w.w = (uint32_t)(x >> 44);
  *reinterpret_cast<uint4*>(dst) = w;
}

__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
  const uint32_t lo = __shfl_sync(GCA_FULL, (uint32_t)v, src);
  const uint32_t hi = __shfl_sync(GCA_FULL, (uint32_t)(v >> 32), src);
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long shfl64_up1(unsigned long long v, int lane) {
  const uint32_t lo = __shfl_up_sync(GCA_FULL, (uint32_t)v, 1);
  const uint32_t hi = __shfl_up_sync(GCA_FULL, (uint32_t)(v >> 32), 1);
  return lane == 0 ? 0ull : (((unsigned long long)hi << 32) | lo);
}
__device__ __forceinline__ unsigned long long shfl64_down1(unsigned long long v, int lane) {
  const uint32_t lo = __shfl_down_sync(GCA_FULL, (uint32_t)v, 1);
  const uint32_t hi = __shfl_down_sync(GCA_FULL, (uint32_t)(v >> 32), 1);
  return lane == 31 ? 0ull : (((unsigned long long)hi << 32) | lo);
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(GCA_FULL, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

// Key schedule of K successive PartiallyObservableForestFireJax.update calls
// (ca_alexandridis_jax.py:436-448 and :352-368), in two parts.  key_chain_pooled: K0 -> K1 -> K2 -> K3 is
// sequential (3K split levels of 2 threefry blocks), so ONE warp of the CTA walks the chains of all its envs,
// a lane pair per env, while the other warps unpack their grids.  key_sides (every owner warp, after the
// front data was requested): pair 2j derives Sburn / Sgrow / the randint keys of sub-step j and pair 2j+1
// its wind draws (4 more levels), then the wind index is threaded through the sub-steps.
// One warp walks the chains of up to 16 envs at once: lane pair p holds env p's key, its even lane leaves the
// side keys in that env's sched rows (S1 -> [0,1], Swind -> [2,3], Sidx -> [4,5]) and the final key in hot.x/y.
__device__ __noinline__ void key_chain_pooled(EnvSmem& ce, const gca_params& P, int lane, bool wr, uint32_t k0,
                                              uint32_t k1) {
  const int K = P.K, mode = P.rng_mode;
#pragma unroll 1
  for (int j = 0; j < K; ++j) {
    uint32_t* sc = ce.sched[j];
    uint32_t n0, n1, s0, s1;
    split_pair(k0, k1, mode, lane, n0, n1, s0, s1);  // K1, S1
    if (wr) { sc[0] = s0; sc[1] = s1; }
    k0 = n0; k1 = n1;
    split_pair(k0, k1, mode, lane, n0, n1, s0, s1);  // K2, Swind
    if (wr) { sc[2] = s0; sc[3] = s1; }
    k0 = n0; k1 = n1;
    split_pair(k0, k1, mode, lane, n0, n1, s0, s1);  // K3, Sidx
    if (wr) { sc[4] = s0; sc[5] = s1; }
    k0 = n0; k1 = n1;
  }
  if (wr) {
    ce.hot.x = k0; ce.hot.y = k1;
    if (!S64_CHAIN_BAR) smem_st_release(&ce.chain_done, 1);  // the owner warp of this env polls it before key_sides
  }
}
__device__ __noinline__ void key_sides(EnvSmem& sm, const gca_params& P, const gca_inject& J, int N, int e,
                                       int lane, int& widx) {
  const int K = P.K, mode = P.rng_mode;
  const int pair = lane >> 1;
  const uint32_t w = lane & 1;
  const bool burn_role = (pair & 1) == 0;
  const int j = pair >> 1;
  uint32_t c0 = 0, c1 = 0, sw0 = 0, sw1 = 0;  // burn role: S1 ; wind role: Sidx and Swind (left by key_chain_pooled)
  if (j < K) {
    const uint32_t* sc = sm.sched[j];
    if (burn_role) { c0 = sc[0]; c1 = sc[1]; }
    else { sw0 = sc[2]; sw1 = sc[3]; c0 = sc[4]; c1 = sc[5]; }
  }
  __syncwarp();  // every lane has its inputs before the rows are overwritten below
  const uint32_t sc0 = (mode == GCA_RNG_LEGACY) ? w : 0u;       // split counters of this lane
  const uint32_t sc1 = (mode == GCA_RNG_LEGACY) ? w + 2u : w;
  uint32_t o0, o1, p0, p1, n0, n1, s0, s1;
  // level 1: burn: split(S1) -> Ka, Sburn ; wind: split(Sidx) -> wk1, wk2
  tf_exchange(c0, c1, sc0, sc1, o0, o1, p0, p1);
  assemble_split(mode, w, o0, o1, p0, p1, n0, n1, s0, s1);
  const uint32_t sburn0 = s0, sburn1 = s1;  // (wind role: wk2)
  uint32_t cur0 = n0, cur1 = n1;            // burn: Ka ; wind: wk1
  // level 2: burn: split(Ka) -> Kb, Sgrow ; wind: even lane bits(wk1,()), odd lane bits(wk2,())
  {
    const uint32_t kk0 = burn_role ? cur0 : (w ? sburn0 : cur0);
    const uint32_t kk1 = burn_role ? cur1 : (w ? sburn1 : cur1);
    tf_exchange(kk0, kk1, burn_role ? sc0 : 0u, burn_role ? sc1 : 0u, o0, o1, p0, p1);
  }
  assemble_split(mode, w, o0, o1, p0, p1, n0, n1, s0, s1);
  const uint32_t sgrow0 = s0, sgrow1 = s1;
  // wind role: hb = even lane's word, lb = odd lane's word
  const uint32_t my_bits = (mode == GCA_RNG_LEGACY) ? o0 : (o0 ^ o1);
  const uint32_t pr_bits = (mode == GCA_RNG_LEGACY) ? p0 : (p0 ^ p1);
  const uint32_t hb = w ? pr_bits : my_bits, lb = w ? my_bits : pr_bits;
  cur0 = n0; cur1 = n1;  // burn: Kb
  // level 3: burn: split(Kb) -> Kc, Sage ; wind: bits(Swind, ())
  tf_exchange(burn_role ? cur0 : sw0, burn_role ? cur1 : sw1, burn_role ? sc0 : 0u, burn_role ? sc1 : 0u,
              o0, o1, p0, p1);
  assemble_split(mode, w, o0, o1, p0, p1, n0, n1, s0, s1);
  const uint32_t uw_bits = (mode == GCA_RNG_LEGACY) ? o0 : (o0 ^ o1);
  // level 4: burn: split(Sage) -> ak1, ak2
  tf_exchange(s0, s1, sc0, sc1, o0, o1, p0, p1);
  uint32_t a10, a11, a20, a21;
  assemble_split(mode, w, o0, o1, p0, p1, a10, a11, a20, a21);
  if (j < K && w == 0) {
    uint32_t* sc = sm.sched[j];
    if (burn_role) {
      sc[0] = sburn0; sc[1] = sburn1; sc[2] = sgrow0; sc[3] = sgrow1;
      sc[4] = a10; sc[5] = a11; sc[6] = a20; sc[7] = a21;
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
  int wi = widx;
  for (int q = 0; q < K; ++q) {
    if (lane == 0) sm.sched[q][8] = (uint32_t)wi;
    if (sm.sched[q][9]) wi = (wi + (int)sm.sched[q][10]) % 8;
  }
  widx = wi;
  __syncwarp();
}

// 9x9 fire window of cell (r, c): rows r-4..r+4 packed 3 per word (9 bits each)
__device__ __forceinline__ void fire_window(const EnvSmem& sm, int r, int c, uint32_t& A, uint32_t& B,
                                            uint32_t& C) {
  const uint32_t* fw = sm.fire32 + r * 4 + (c >> 4);
  const int o = c & 15;
  uint32_t wv[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wv[i] = (fw[i * 4] >> o) & 0x1FFu;
  A = wv[0] | (wv[1] << 9) | (wv[2] << 18);
  B = wv[3] | (wv[4] << 9) | (wv[5] << 18);
  C = wv[6] | (wv[7] << 9) | (wv[8] << 18);
}
// 5x5 doused window: rows r-2..r+2, 5 bits each (columns c-2..c+2; outside the grid = 0)
__device__ __forceinline__ uint32_t dous_window(const EnvSmem& sm, int r, int c) {
  uint32_t v = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const unsigned long long row = sm.dous64[r + i];  // index r - 2 + i + 2
    const uint32_t bits = (uint32_t)(c >= 2 ? (row >> (c - 2)) : (row << (2 - c))) & 0x1Fu;
    v |= bits << (5 * i);
  }
  return v;
}

// Exact (reference-order) value of (heat - dousing) (1+p_veg) (1+p_den) for one cell: row-major
// sequential float32 sums with the accumulator starting at +0 (oracle/alexandridis.py:_window_sum).
__device__ __noinline__ float exact_base(const EnvSmem& sm, const gca_params& P, int r, int c, float a, float b) {
  uint32_t A, B, C;
  fire_window(sm, r, c, A, B, C);
  const uint32_t rows3[3] = {A, B, C};
  float heat = 0.0f;
  for (int i = 0; i < 9; ++i) {
    const uint32_t bits = (rows3[i / 3] >> (9 * (i % 3))) & 0x1FFu;
    const int di = i < 4 ? 4 - i : i - 4;
    for (int jj = 0; jj < 9; ++jj) {
      if ((bits >> jj) & 1u) {
        const int dj = jj < 4 ? 4 - jj : jj - 4;
        const int ring = di > dj ? di : dj;
        heat = __fadd_rn(heat, P.ring_w[ring]);
      }
    }
  }
  const uint32_t dwin = dous_window(sm, r, c);
  float dous = 0.0f;
  for (int i = 0; i < 5; ++i)
    for (int jj = 0; jj < 5; ++jj)
      if ((dwin >> (5 * i + jj)) & 1u) {
        const bool inner = i >= 1 && i <= 3 && jj >= 1 && jj <= 3;
        dous = __fadd_rn(dous, inner ? P.dous_inner : P.dous_border);
      }
  const float ph = __fsub_rn(heat, dous);
  return __fmul_rn(__fmul_rn(ph, a), b);
}

// Compact the front cells [pass_base, pass_base + CAP) of the row masks fr0 (row 2*lane) and fr1
// (row 2*lane+1) into sm.list and return the total number of front cells.  The rows are first
// re-dealt in 16-column pieces (piece p = 4*row + quarter goes to lane p % 32) so that a long
// horizontal run of front cells -- the top/bottom edge of a burning blob -- is shared by several
// lanes instead of serialising one.  `scratch` = 512 bytes private to the warp.
__device__ __noinline__ int build_front_list(EnvSmem& sm, uint16_t* scratch, unsigned long long fr0,
                                             unsigned long long fr1, int lane, int pass_base) {
  reinterpret_cast<ulonglong2*>(scratch)[lane] = make_ulonglong2(fr0, fr1);
  __syncwarp();
  const uint16_t* q16 = scratch;
  // this lane's 8 pieces as two 64-bit words: bit 16*k + b of word h <-> piece (4h + k), column bit b
  unsigned long long w[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t lo = (uint32_t)q16[lane + 32 * (4 * h)] | ((uint32_t)q16[lane + 32 * (4 * h + 1)] << 16);
    const uint32_t hi = (uint32_t)q16[lane + 32 * (4 * h + 2)] | ((uint32_t)q16[lane + 32 * (4 * h + 3)] << 16);
    w[h] = ((unsigned long long)hi << 32) | lo;
  }
  const int n = __popcll(w[0]) + __popcll(w[1]);
  const int incl = warp_incl_scan(n, lane);
  const int T = __shfl_sync(GCA_FULL, incl, 31);
  int idx = incl - n - pass_base;
  // piece p = lane + 32 q  ->  row (lane >> 2) + 8 q, columns 16 (lane & 3) ..
  const uint32_t lane_base = ((uint32_t)(lane >> 2) << 6) | ((uint32_t)(lane & 3) << 4);
  unsigned long long m = w[0];
  uint32_t word_base = lane_base;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    while (m) {
      const uint32_t b = (uint32_t)__ffsll((long long)m) - 1u;
      m &= m - 1;
      // piece within the word = b >> 4 -> 8 rows further down per piece
      const uint32_t cell = word_base + ((b >> 4) << 9) + (b & 15u);
      if ((unsigned)idx < (unsigned)S64_CAP) sm.list[idx] = (uint16_t)cell;
      ++idx;
    }
    m = w[1];
    word_base = lane_base + (4u << 9);
  }
  __syncwarp();
  return T;
}

// Touch the hidden byte and the 32-byte slope-factor sector of listed front cells [from, to) so
// that they are in L2 (and, capacity permitting, L1) when the cell and draw phases ask for them.
__device__ __forceinline__ void prefetch_front(const EnvSmem& sm, const uint8_t* hidden, const float* pslope,
                                               size_t cell_base, int from, int to, int lane) {
  if (hidden == nullptr) return;
  for (int t = from + lane; t < to; t += 32) {
    const uint32_t cell = sm.list[t];
    prefetch_l1(hidden + cell_base + cell);
    if (pslope != nullptr) prefetch_l1(pslope + (cell_base + cell) * 8);
  }
}

// A buffered (front cell, burning direction) draw whose uniform fell below the stored upper bound of its burn
// probability (~2 % of the draws): u < hi (1 - 2^-14) ignites for sure; anything in between -- or any hit next to doused
// cells -- is re-evaluated with the reference's exact summation order ("threshold cells").  Ignitions are or-ed into the
// env's accumulator.
__device__ __forceinline__ void pair_hit(const gca_params& P, const uint8_t* hidden, EnvSmem& es, uint32_t ent,
                                         uint32_t cell, float w, float sl, float phi, float u, uint32_t& n_thresh) {
  bool ig = !(ent >> 31) && u < __fmul_rn(phi, S64_SURE);
  if (!ig) {
    // threshold cell: the float32 enclosure cannot decide -> reference-order evaluation
    const int r = cell >> 6, col = cell & 63;
    int hid = 3 | (3 << 3);
    if (hidden != nullptr) hid = hidden[(size_t)es.hot.w * 4096 + cell];
    const float a = P.onep_veg[clip15(hid & 7)];
    const float b = P.onep_den[clip15((hid >> 3) & 7)];
    const float base = exact_base(es, P, r, col, a, b);
    const float p = __fmul_rn(__fmul_rn(base, w), sl);
    ig = u < p;
    n_thresh++;
  }
  if (ig) atomicOr(reinterpret_cast<uint32_t*>(es.ign) + (cell >> 5), 1u << (cell & 31));
}

// empty -> tree with probability p_tree (0 in the reference env, so this is a cold path): a dense
// draw per empty cell of the two rows this lane owns.
__device__ __noinline__ void regrow_rows(const gca_params& P, const gca_inject& J, const TfKey kg, size_t inj_base,
                                         int lane, unsigned long long e0, unsigned long long e1,
                                         unsigned long long& g0, unsigned long long& g1) {
  unsigned long long m = e0;
  int row = 2 * lane;
  for (int half = 0; half < 2; ++half) {
    unsigned long long g = 0;
    while (m) {
      const int c = __ffsll((long long)m) - 1;
      m &= m - 1;
      const uint32_t cell = (uint32_t)(row * 64 + c);
      float u;
      if (J.u_grow) u = J.u_grow[inj_base + cell];
      else u = bits_to_uniform(bits_at_ni(kg, cell, S64_HALF_CELL, P.rng_mode));
      if (u < P.p_tree) g |= 1ull << c;
    }
    if (half == 0) g0 = g; else g1 = g;
    m = e1;
    row = 2 * lane + 1;
  }
}

// Fire ages of the n cells on sm.ignlist (ca_alexandridis_jax.py:367-370,394-398): two lanes per ignition --
// jax.random.randint needs two independent words -- with the randint keys of the sub-step the cell ignited
// in; the burn-out tick goes to S.death and, through a row-minimum scratch, into the owners' row minima.
// Nothing inside the env step reads these ticks (ages are >= 144 CA updates), so all sub-steps' ignitions
// are drawn in one batch.  `rowmin` = 64 words private to the warp.
__device__ __noinline__ void flush_ages(const EnvSmem& sm, const gca_params& P, const gca_state& S,
                                        const int32_t* j_age_new, int mode, int N, int e, int n, uint32_t tick0,
                                        uint32_t* rowmin, int lane) {
  const size_t cell_base = (size_t)e * 4096;
  const uint32_t age_magic = 0xFFFFFFFFu / P.age_span;
  rowmin[2 * lane] = 0xFFFFFFFFu;
  rowmin[2 * lane + 1] = 0xFFFFFFFFu;
  __syncwarp();
  for (int tb = 0; tb < 2 * n; tb += 32) {
    const int task = tb + lane, i = task >> 1;
    const bool valid = i < n;
    const uint32_t ent = sm.ignlist[valid ? i : 0];
    const uint32_t cell = ent & 0xFFFu, jj = ent >> 12;
    const uint32_t* sc = sm.sched[jj];
    uint32_t bits = 0;
    if (j_age_new == nullptr)
      bits = bits_at_ni((lane & 1) ? tf_key(sc[6], sc[7]) : tf_key(sc[4], sc[5]), cell, S64_HALF_CELL, mode);
    const uint32_t other = __shfl_xor_sync(GCA_FULL, bits, 1);
    if (valid && !(lane & 1)) {
      int age;
      if (j_age_new) age = j_age_new[((size_t)jj * N + e) * 4096 + cell];
      else {
        const uint32_t hm = fastmod(bits, P.age_span, age_magic), lm = fastmod(other, P.age_span, age_magic);
        age = P.age_lo + (int)fastmod(hm * P.age_mult + lm, P.age_span, age_magic);
      }
      const uint32_t dabs = tick0 + jj + (uint32_t)age;  // burn-out tick
      S.death[cell_base + cell] = (uint16_t)dabs;
      atomicMin(&rowmin[cell >> 6], dabs);
    }
  }
  __syncwarp();  // the caller folds rowmin[2 lane], rowmin[2 lane + 1] into its row minima
}

__device__ __forceinline__ void front_masks(unsigned long long t0, unsigned long long t1, unsigned long long f0,
                                            unsigned long long f1, int lane, unsigned long long& fr0,
                                            unsigned long long& fr1) {
  // front = tree cells with a burning Moore neighbour (row halos by warp shuffle)
  const unsigned long long fh0 = f0 | (f0 << 1) | (f0 >> 1);
  const unsigned long long fh1 = f1 | (f1 << 1) | (f1 >> 1);
  fr0 = t0 & (shfl64_up1(fh1, lane) | fh0 | fh1);
  fr1 = t1 & (fh0 | fh1 | shfl64_down1(fh0, lane));
}

// One round of the burn-out scan of env `es` (pooled work item of sub-step 0): rows rowlist[4 q .. 4 q + 3] -- lane
// group g = lane / 8 takes one row, each of its lanes 8 cells (one 128-bit load of their burn-out ticks).  Cells that
// burn out during this env step get their tick cleared (fire_age ends at 0), the row's masks (which cells, bit planes of
// the sub-step) are parked in es.burn[row] for the owner's apply phases, and the row's new minimum goes to S.row_min
// (L2: the owner re-reads it after the phase's closing barrier).
__device__ __noinline__ void scan_round(EnvSmem& es, const gca_state& S, int q, int K, int lane) {
  const int g = lane >> 3, li = lane & 7;
  const int nrows = es.nscan;
  const uint32_t tick0 = es.tick0;
  const size_t e = es.hot.w;
  const int ri = 4 * q + g;
  const bool have = ri < nrows;
  const int row = have ? (int)reinterpret_cast<const uint8_t*>(es.ignlist)[ri] : 0;
  uint16_t* const drow = S.death + e * 4096 + row * 64 + 8 * li;
  uint4 dv = make_uint4(0u, 0u, 0u, 0u);
  if (have) dv = *reinterpret_cast<const uint4*>(drow);
  // fire bits of this lane's 8 cells: view k of a row starts at column 16 k - 4
  const uint32_t fb = have ? (es.fire32[(row + 4) * 4 + (li >> 1)] >> (8 * (li & 1) + 4)) & 0xFFu : 0u;
  uint32_t w[4] = {dv.x, dv.y, dv.z, dv.w};
  uint32_t die8 = 0, p0 = 0, p1 = 0, p2 = 0, vmin = 0xFFFFFFFFu;
#pragma unroll
  for (int k = 0; k < 8; ++k) {  // branch-free: selects and masks only
    const uint32_t d = (w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
    const uint32_t r = (d - tick0) & 0xFFFFu;
    const uint32_t f = (fb >> k) & 1u;
    const uint32_t x = r < (uint32_t)K ? f : 0u;  // burns out during this env step (at sub-step r)
    die8 |= x << k;
    p0 |= (x & r) << k;
    p1 |= (x & (r >> 1)) << k;
    p2 |= (x & (r >> 2)) << k;
    w[k >> 1] &= ~((0u - x) & (0xFFFFu << (16 * (k & 1))));  // burnt-out cell: fire_age ends at 0
    vmin = min(vmin, (f ^ x) ? tick0 + r : 0xFFFFFFFFu);      // fires that keep burning: earliest burn-out
  }
  if (die8) *reinterpret_cast<uint4*>(drow) = make_uint4(w[0], w[1], w[2], w[3]);
  // the group's first lane collects the 8 packed byte-quadruples (die, plane 0..2) and transposes them into four
  // 64-bit row masks
  const uint32_t packed = die8 | (p0 << 8) | (p1 << 16) | (p2 << 24);
  uint32_t wq[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) wq[k] = __shfl_sync(GCA_FULL, packed, (lane & 24) + k);
  unsigned long long rowm[4];
#pragma unroll
  for (int qq = 0; qq < 4; ++qq) {
    // inner perm: byte 0 = x.byte[qq], byte 1 = y.byte[qq]; outer perm: bytes a0 a1 b0 b1
    const uint32_t lo = __byte_perm(__byte_perm(wq[0], wq[1], 0x0040u + qq + (qq << 4)),
                                    __byte_perm(wq[2], wq[3], 0x0040u + qq + (qq << 4)), 0x5410u);
    const uint32_t hi = __byte_perm(__byte_perm(wq[4], wq[5], 0x0040u + qq + (qq << 4)),
                                    __byte_perm(wq[6], wq[7], 0x0040u + qq + (qq << 4)), 0x5410u);
    rowm[qq] = ((unsigned long long)hi << 32) | lo;
  }
  uint32_t newmin = vmin;
#pragma unroll
  for (int d = 1; d < 8; d <<= 1) newmin = min(newmin, __shfl_xor_sync(GCA_FULL, newmin, d));
  if (have && li == 0) {
    reinterpret_cast<ulonglong2*>(es.burn[row])[0] = make_ulonglong2(rowm[0], rowm[1]);
    reinterpret_cast<ulonglong2*>(es.burn[row])[1] = make_ulonglong2(rowm[2], rowm[3]);
    __stcg(S.row_min + e * 64 + row, newmin);
  }
}

// n 32-bit words from device memory (read through L2: other SMs wrote them) to mapped host memory by one warp: 128-bit
// accesses, eight loads in flight per lane, so the copy costs a handful of L2 round trips and one PCIe burst
__device__ __noinline__ void burst_copy_u32(uint32_t* dst, const uint32_t* src, int n, int lane) {
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
    const int n4 = n >> 2;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int base = 0; base < n4; base += 32 * 8) {
      uint4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + k * 32 + lane;
        if (i < n4) v[k] = __ldcg(s4 + i);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + k * 32 + lane;
        if (i < n4) d4[i] = v[k];
      }
    }
    for (int i = (n4 << 2) + lane; i < n; i += 32) dst[i] = __ldcg(src + i);
  } else {
    for (int i = lane; i < n; i += 32) dst[i] = __ldcg(src + i);
  }
}

// Observation of one env by its owner warp (fused epilogue of the step): grid_to_rgb without extensions
// (advanced_bulldozer.py:1035-1101) -- palette by day / night, doused cells blended 0.25 rgb + 0.75 tint in float32,
// bulldozer pixel black -- from the tree / fire rows the lanes hold (lane l: rows 2l, 2l+1) and the doused rows in
// sm.dous64.  The rows are parked in shared memory (sm.burn / sm.ign: the caller makes sure they are free and clears
// sm.ign again if the pooled phases still need it).  `scratch`: the warp's pair buffer (>= 320 bytes): colour tables.
// colour tables of one env in `scratch` (320 bytes): [16] float4 r g b -- entry = doused << 2 | cell state (0 empty,
// 1 tree, 2 fire, 3 unused), entries 8..15 black (bulldozer) -- then [3][8] bytes: R of entries 0..7, G, B
__device__ __forceinline__ void render_lut(uint32_t* scratch, uint32_t night, int lane) {
  float4* lutf = reinterpret_cast<float4*>(scratch);
  uint8_t* lutb = reinterpret_cast<uint8_t*>(scratch) + 256;
  if (lane < 16) {
    float cr = 0.f, cg = 0.f, cb = 0.f;
    if (lane < 8) render_pixel(lane & 3, night != 0, lane >= 4, false, cr, cg, cb);
    lutf[lane] = make_float4(cr, cg, cb, 0.f);
    if (lane < 8) {
      lutb[lane] = (uint8_t)cr; lutb[8 + lane] = (uint8_t)cg; lutb[16 + lane] = (uint8_t)cb;
    }
  }
}

// The frame was drawn from the grid the step STARTED with (render_env64 in the prologue, so that its stores overlap
// the CA work); here the pixels of the cells that changed during the step (masks ch0 / ch1 of rows 2 lane, 2 lane + 1)
// are redrawn from the final rows.  The bulldozer pixel (prow, pcol) stays black.
template <bool U8>
__device__ __noinline__ void render_patch64(EnvSmem& sm, uint32_t* scratch, void* rgb, int e, unsigned long long t0,
                                            unsigned long long t1, unsigned long long f0, unsigned long long f1,
                                            unsigned long long ch0, unsigned long long ch1, int prow, int pcol,
                                            uint32_t night, int lane) {
  if (__ballot_sync(GCA_FULL, (ch0 | ch1) != 0ull) == 0u) return;
  render_lut(scratch, night, lane);
  __syncwarp();
  const float4* lutf = reinterpret_cast<const float4*>(scratch);
  unsigned long long ch = ch0, tt = t0, ff = f0;
  int row = 2 * lane;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const unsigned long long dd = sm.dous64[2 + row];
    while (ch) {
      const int c = __ffsll((long long)ch) - 1;
      ch &= ch - 1;
      if (row == prow && c == pcol) continue;
      const uint32_t code = (uint32_t)((tt >> c) & 1ull) | ((uint32_t)((ff >> c) & 1ull) << 1) | ((uint32_t)((dd >> c) & 1ull) << 2);
      const float4 px = lutf[code];
      const size_t o = ((size_t)e * 4096 + (size_t)row * 64 + c) * 3;
      if (U8) {
        uint8_t* p = reinterpret_cast<uint8_t*>(rgb) + o;
        p[0] = (uint8_t)px.x; p[1] = (uint8_t)px.y; p[2] = (uint8_t)px.z;
      } else {
        float* p = reinterpret_cast<float*>(rgb) + o;
        p[0] = px.x; p[1] = px.y; p[2] = px.z;
      }
    }
    ch = ch1; tt = t1; ff = f1;
    row = 2 * lane + 1;
  }
  __syncwarp();
}

// bit k of a 16-bit mask -> bit 4 k (one nibble per cell)
__device__ __forceinline__ unsigned long long spread16(uint32_t v) {
  unsigned long long x = v;
  x = (x | (x << 24)) & 0x000000FF000000FFull;
  x = (x | (x << 12)) & 0x000F000F000F000Full;
  x = (x | (x << 6)) & 0x0303030303030303ull;
  x = (x | (x << 3)) & 0x1111111111111111ull;
  return x;
}

template <bool U8>
__device__ __noinline__ void render_env64(EnvSmem& sm, uint32_t* scratch, void* rgb, int e, unsigned long long t0,
                                          unsigned long long t1, unsigned long long f0, unsigned long long f1, int prow,
                                          int pcol, uint32_t night, int lane) {
  unsigned long long* const flat = &sm.burn[0][0];   // 2 KB, free outside [barrier A of sub-step 0, last apply]
  reinterpret_cast<ulonglong2*>(flat)[lane] = make_ulonglong2(t0, t1);
  reinterpret_cast<ulonglong2*>(sm.ign)[lane] = make_ulonglong2(f0, f1);
  render_lut(scratch, night, lane);
  float4* lutf = reinterpret_cast<float4*>(scratch);
  uint8_t* lutb = reinterpret_cast<uint8_t*>(scratch) + 256;
  __syncwarp();
  const uint16_t* tq = reinterpret_cast<const uint16_t*>(flat);
  const uint16_t* fq = reinterpret_cast<const uint16_t*>(sm.ign);
  const uint16_t* dq = reinterpret_cast<const uint16_t*>(sm.dous64 + 2);
  const int q = lane & 3;
  uint32_t Rlo = 0, Rhi = 0, Glo = 0, Ghi = 0, Blo = 0, Bhi = 0;
  if (U8) {
    const uint2 R = reinterpret_cast<const uint2*>(lutb)[0], G = reinterpret_cast<const uint2*>(lutb)[1],
                B = reinterpret_cast<const uint2*>(lutb)[2];
    Rlo = R.x; Rhi = R.y; Glo = G.x; Ghi = G.y; Blo = B.x; Bhi = B.y;
  }
  // pixels are staged in shared memory (sm.burn: 2 KB, free by now) so that every global store of the warp covers 512
  // contiguous bytes: uint8 -- 8 rows (1536 B) per round, a lane renders 16 cells; float32 -- 2 rows (1536 B) per round,
  // a lane renders 4 cells
  uint4* const stage = reinterpret_cast<uint4*>(flat + 64);
  if (U8) {
    uint4* const out = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(rgb) + (size_t)e * 4096 * 3);
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
      const int r = 8 * it + (lane >> 2);
      const uint32_t t16 = tq[r * 4 + q], f16 = fq[r * 4 + q], d16 = dq[r * 4 + q];
      const int pk = (r == prow && (pcol >> 4) == q) ? (pcol & 15) : -1;  // the bulldozer's cell among this lane's 16
      uint32_t w[12];
      // 16 cells -> 16 nibbles: state | doused << 2 (tree and fire are disjoint, fire = 2)
      const unsigned long long sel64 = spread16(t16) | (spread16(f16) << 1) | (spread16(d16) << 2);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t sel = (uint32_t)(sel64 >> (16 * g)) & 0xFFFFu;
        const uint32_t R4 = __byte_perm(Rlo, Rhi, sel), G4 = __byte_perm(Glo, Ghi, sel), B4 = __byte_perm(Blo, Bhi, sel);
        // bytes R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
        const uint32_t rg = __byte_perm(R4, G4, 0x5140u);   // R0 G0 R1 G1
        const uint32_t rg2 = __byte_perm(R4, G4, 0x7362u);  // R2 G2 R3 G3
        w[3 * g] = __byte_perm(rg, B4, 0x2410u);            // R0 G0 B0 R1
        w[3 * g + 1] = __byte_perm(__byte_perm(rg, B4, 0x0053u), rg2, 0x5410u);  // G1 B1 R2 G2
        w[3 * g + 2] = __byte_perm(rg2, B4, 0x7326u);       // B2 R3 G3 B3
      }
      if (pk >= 0) {  // black pixel: bytes 3 pk .. 3 pk + 2 of the 48
#pragma unroll
        for (int jw = 0; jw < 12; ++jw) {
          uint32_t m = 0;
#pragma unroll
          for (int bb = 0; bb < 4; ++bb) {
            const int byte = 4 * jw + bb;
            if (byte >= 3 * pk && byte < 3 * pk + 3) m |= 0xFFu << (8 * bb);
          }
          w[jw] &= ~m;
        }
      }
      stage[3 * lane] = make_uint4(w[0], w[1], w[2], w[3]);
      stage[3 * lane + 1] = make_uint4(w[4], w[5], w[6], w[7]);
      stage[3 * lane + 2] = make_uint4(w[8], w[9], w[10], w[11]);
      __syncwarp();
      const uint4 a0 = stage[lane], a1 = stage[lane + 32], a2 = stage[lane + 64];
      // streaming stores: the frame is written once and read by another kernel (measured: float32 111 us per step
      // against 116 with plain stores)
      __stcs(&out[96 * it + lane], a0);
      __stcs(&out[96 * it + 32 + lane], a1);
      __stcs(&out[96 * it + 64 + lane], a2);
      __syncwarp();
    }
  } else {
    float4* const out = reinterpret_cast<float4*>(reinterpret_cast<float*>(rgb) + (size_t)e * 4096 * 3);
    float4* const stf = reinterpret_cast<float4*>(stage);
    const int h = lane >> 4, c4 = lane & 15;  // row of the pair, group of 4 cells
#pragma unroll 1
    for (int it = 0; it < 32; ++it) {
      const int r = 2 * it + h;
      const int sh = 4 * (c4 & 3);
      const uint32_t t4 = (uint32_t)tq[r * 4 + (c4 >> 2)] >> sh, f4 = (uint32_t)fq[r * 4 + (c4 >> 2)] >> sh,
                     d4 = (uint32_t)dq[r * 4 + (c4 >> 2)] >> sh;
      const int pk = (r == prow && (pcol >> 2) == c4) ? (pcol & 3) : -1;
      float4 c[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t code = ((t4 >> k) & 1u) | (((f4 >> k) & 1u) << 1) | (((d4 >> k) & 1u) << 2);
        if (k == pk) code = 8u;
        c[k] = lutf[code];
      }
      stf[3 * lane] = make_float4(c[0].x, c[0].y, c[0].z, c[1].x);
      stf[3 * lane + 1] = make_float4(c[1].y, c[1].z, c[2].x, c[2].y);
      stf[3 * lane + 2] = make_float4(c[2].z, c[3].x, c[3].y, c[3].z);
      __syncwarp();
      const float4 a0 = stf[lane], a1 = stf[lane + 32], a2 = stf[lane + 64];
      // streaming stores: the frame is written once and read by another kernel (measured: float32 111 us per step
      // against 116 with plain stores)
      __stcs(&out[96 * it + lane], a0);
      __stcs(&out[96 * it + 32 + lane], a1);
      __stcs(&out[96 * it + 64 + lane], a2);
      __syncwarp();
    }
  }
  __syncwarp();
}

// S64_TRACE (diagnostic builds only): per-env phase timestamps (SM clock, relative to the warp's start) go to
// O.stats[8 + 16 e ...]; the caller must have allocated stats with 8 + 32 N words.
#ifdef S64_TRACE
#define S64_STAMP(k) do { if (active && lane == 0 && O.stats) O.stats[8 + 32 * (size_t)e + (k)] = (unsigned long long)(clock64() - clk0); } while (0)
#else
#define S64_STAMP(k) do { } while (0)
#endif
// S64_ACC(k): add the cycles since the last S64_ACC / S64_MARK of this warp to trace slot k
#ifdef S64_TRACE
#define S64_MARK() do { trace_t = clock64(); } while (0)
#define S64_ACC(k) do { const long long now_ = clock64(); if (active && lane == 0 && O.stats) O.stats[8 + 32 * (size_t)e + (k)] += (unsigned long long)(now_ - trace_t); trace_t = now_; } while (0)
#else
#define S64_MARK() do { } while (0)
#define S64_ACC(k) do { } while (0)
#endif

#ifndef S64_MINB
#define S64_MINB (28 / S64_E > 0 ? 28 / S64_E : 1)  // 28 warps/SM x 72 registers = the whole register file; 4096 envs = one wave of 148 x 28
#endif
// MODE: GCA_RNG_* or -1 (read P.rng_mode); HP: 0 = no hidden layers, 1 = hidden + slope table present,
// -1 = test the pointers at run time; INJ: injected random fields may be present.  The launcher picks
// a fully specialised instance for the production cases and the generic one otherwise.
template <int MODE, int HP, bool INJ>
__global__ void __launch_bounds__(S64_E * 32, S64_MINB)
env_step64_kernel(const __grid_constant__ gca_params P, const __grid_constant__ gca_state S,
                  const int32_t* __restrict__ actions, const __grid_constant__ gca_step_out O,
                  const __grid_constant__ gca_inject J, const __grid_constant__ gca_state SNAP,
                  const float* __restrict__ snap_reward, uint32_t flags) {
  extern __shared__ __align__(16) unsigned char s64_smem_raw[];
  CtaSmem& cs = *reinterpret_cast<CtaSmem*>(s64_smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long clk0 = clock64();
#ifdef S64_TRACE
  long long trace_t = clk0;
#endif
  const int group = warp / S64_GE, gwarp = warp % S64_GE;  // lock-step group of this warp, index inside it
  const int slot = blockIdx.x * S64_E + warp;
  const int N = S.N;
  const bool active = slot < N;   // a warp without an env still joins the barriers and the pooled work
  // optional load-balancing indirection (gca_balance_order): which env this warp owns
  const int e = active ? (S.order != nullptr ? S.order[slot] : slot) : 0;
  EnvSmem& sm = cs.env[warp];
  uint32_t* const wp32 = cs.pairs[warp];  // warp-private pair buffer (and scratch of the owner phases)
  uint16_t* const wpi = cs.pidx[warp];
  uint16_t* const wp = reinterpret_cast<uint16_t*>(wp32);
  if (!S64_CHAIN_BAR) {
    if (lane == 0) sm.chain_done = 0;
    __syncthreads();  // (every warp is here within a few cycles of the launch) the flags are clear before a chain warp can set one
  }
  const int K = P.K, mode = MODE < 0 ? P.rng_mode : MODE;
  const size_t cell_base = (size_t)e * 4096;
  const uint8_t* const hidden = HP == 0 ? nullptr : S.hidden;
  const float* const pslope = HP == 0 ? nullptr : S.pslope;
  if (HP == 1) { __builtin_assume(hidden != nullptr); __builtin_assume(pslope != nullptr); }
  const float* const j_u_burn = INJ ? J.u_burn : nullptr;
  const int32_t* const j_age_new = INJ ? J.age_new : nullptr;

  unsigned long long t0 = 0, t1 = 0, f0 = 0, f1 = 0;
  unsigned long long ch0 = 0ull, ch1 = 0ull;  // cells whose state changed during this env step
  uint2 rm = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
  uint32_t tick0 = 0;
  int widx = 0;
  uint32_t burnrows = 0;  // bit 0 / 1: row 2*lane / 2*lane+1 has burn-outs in this env step (sm.burn valid)
  int T = 0, L = 0;
  bool dense = false;
  uint32_t n_draws = 0, n_thresh = 0, n_front = 0, n_ign = 0, n_ext = 0;
  uint32_t work = 0;  // warp-uniform cost estimate of this env step: front-list entries over its sub-steps

  ulonglong2 dz = make_ulonglong2(0ull, 0ull);
  if (active) {
    // ---- everything the step needs from HBM up front (the key is fetched by the warp that walks the chains): the
    //      tree / fire bit-boards of the grid (1 KB per env; lane l takes rows 2l, 2l+1), doused rows,
    //      row minima, scalars --------------------------------------------------------------------------
    S64_STAMP(16);
    {
      const ulonglong2* bbp = reinterpret_cast<const ulonglong2*>(S.bb + (size_t)e * 128);
      const ulonglong2 r0 = bbp[2 * lane], r1 = bbp[2 * lane + 1];
      t0 = r0.x; f0 = r0.y; t1 = r1.x; f1 = r1.y;
    }
    dz = reinterpret_cast<const ulonglong2*>(S.doused + (size_t)e * 64)[lane];
    rm = reinterpret_cast<const uint2*>(S.row_min + (size_t)e * 64)[lane];
    S64_STAMP(17);
    tick0 = S.tick[e];
    widx = S.wind_index[e];
    if (!(flags & GCA_FLAG_CA_ONLY)) {
      // the per-env scalars of the epilogue (lane 0, serial): get their lines on the way now
      // the action triple goes to the three spare words behind the wind matrix with asynchronous copies: when
      // `actions` is mapped host memory (gca_env_step_host, zero-copy) the bus round trip hides behind the step
      if (lane < 3) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sm.wind[9 + lane]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(actions + 3 * (size_t)e + lane) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      const void* pf = nullptr;
      switch (lane) {
        case 1: pf = S.time + e; break;
        case 2: pf = S.position + 2 * e; break;
        case 3: pf = S.time_step + e; break;
        case 4: pf = S.is_night + e; break;
        case 5: pf = S.steps_elapsed ? S.steps_elapsed + e : nullptr; break;
        case 6: pf = S.reward_accumulated ? S.reward_accumulated + e : nullptr; break;
        default: break;
      }
      if (pf != nullptr) prefetch_l1(pf);
    }
    if (lane < 16) sm.fire32[lane] = 0u; else sm.fire32[272 + lane - 16] = 0u;   // fire rows -4..-1, 64..67
    if (lane < 2) sm.dous64[lane] = 0ull; else if (lane < 4) sm.dous64[64 + lane] = 0ull;  // rows -2,-1,64,65
    S64_STAMP(0);
  }
  if (warp >= S64_E - S64_CHAIN_WARPS) {
    // ~3K dependent threefry blocks per env, for 16 envs at a time, while the grids stream in: lane pair p
    // fetches the key of the env in slot p itself (no barrier in front of the chain)
    const int p = (S64_E - 1 - warp) * 16 + (lane >> 1);
    const int pslot = blockIdx.x * S64_E + p;
    const bool pv = p < S64_E && pslot < N;
    uint32_t ck0 = 0, ck1 = 0;
    if (pv) {
      const int pe = S.order != nullptr ? S.order[pslot] : pslot;
      ck0 = S.key[2 * pe];
      ck1 = S.key[2 * pe + 1];
    }
    key_chain_pooled(cs.env[pv ? p : 0], P, lane, pv && !(lane & 1), ck0, ck1);
    if (S64_CHAIN_BAR) {
      // the sched rows / hot.x, hot.y written above are ordered before the arrival (producer side of bar.arrive / bar.sync)
#pragma unroll
      for (int w = 0; w < S64_E - 1; ++w) asm volatile("bar.arrive %0, 64;" ::"r"(w + 1) : "memory");
    }
  }
  if (active) {
    S64_STAMP(19);
    sm.ign[2 * lane] = 0ull;
    sm.ign[2 * lane + 1] = 0ull;
    reinterpret_cast<ulonglong2*>(sm.dous64 + 2)[lane] = dz;  // rows 2 lane, 2 lane + 1
    {
      // which rows have a doused cell within the 5x5 dousing window's reach (2 rows up / down)
      const uint32_t ev = __ballot_sync(GCA_FULL, dz.x != 0ull), od = __ballot_sync(GCA_FULL, dz.y != 0ull);
      if (lane == 0) {
        sm.dous_even = ev | (ev << 1) | (ev >> 1) | od | (od << 1);
        sm.dous_odd = od | (od << 1) | (od >> 1) | ev | (ev >> 1);
        sm.nign = 0;
      }
    }
    store_row_views(sm.fire32 + (2 * lane + 4) * 4, f0);
    store_row_views(sm.fire32 + (2 * lane + 5) * 4, f1);

    // ---- rows holding a cell that burns out during this env step: request their burn-out ticks now ----
    const bool need0 = f0 != 0ull && rm.x < tick0 + (uint32_t)K;
    const bool need1 = f1 != 0ull && rm.y < tick0 + (uint32_t)K;
    burnrows = (need0 ? 1u : 0u) | (need1 ? 2u : 0u);
    if (need0) prefetch_l1(S.death + cell_base + (2 * lane) * 64);
    if (need1) prefetch_l1(S.death + cell_base + (2 * lane + 1) * 64);

    // ---- front of sub-step 0: compact it and start fetching its hidden / slope-factor sectors.  The
    //      list is built ONCE per env step; later sub-steps append the cells that joined the front (a
    //      handful); entries that left it (ignited / no burning neighbour left) are recognised from
    //      their fire window.
    __syncwarp();
    {
      unsigned long long fr0, fr1;
      front_masks(t0, t1, f0, f1, lane, fr0, fr1);
      T = build_front_list(sm, wp, fr0, fr1, lane, 0);
      dense = T > S64_CAP;  // more front cells than the list holds: rebuild per sub-step, in passes
      L = min(T, S64_CAP);
      sm.listed[2 * lane] = fr0;
      sm.listed[2 * lane + 1] = fr1;
      n_front += (uint32_t)(__popcll(fr0) + __popcll(fr1));
    }
    S64_STAMP(1);

    // ---- rows to scan for burn-outs: the scan itself is POOLED work of sub-step 0 (scan items next to the front-cell
    //      chunks: every warp of the CTA takes rounds of four rows of whichever env), so an env with many burning
    //      rows does not hold its owner -- and with it the CTA's first barrier -- back.  The owner only lists the rows.
    {
      const uint32_t Me = __ballot_sync(GCA_FULL, need0), Mo = __ballot_sync(GCA_FULL, need1);
      const uint32_t lt = (1u << lane) - 1u;
      uint8_t* const rowlist = reinterpret_cast<uint8_t*>(sm.ignlist);
      if (need0) rowlist[__popc(Me & lt)] = (uint8_t)(2 * lane);
      if (need1) rowlist[__popc(Me) + __popc(Mo & lt)] = (uint8_t)(2 * lane + 1);
      if (lane == 0) { sm.nscan = __popc(Me) + __popc(Mo); sm.tick0 = tick0; }
    }
    S64_STAMP(2);
    S64_STAMP(3);
    // the front cells' hidden / slope-factor sectors: their DRAM latency hides behind the second half of the key schedule
    prefetch_front(sm, hidden, pslope, cell_base, 0, L, lane);
    S64_STAMP(18);
  }
  if (S64_CHAIN_BAR && warp < S64_E - 1) asm volatile("bar.sync %0, 64;" ::"r"(warp + 1) : "memory");
  if (active) {
    // the env's key chain is done (the chain warp walks it while the grids stream in); no CTA-wide barrier
    S64_MARK();
    if (!S64_CHAIN_BAR) {
      while (smem_ld_acquire(&sm.chain_done) == 0) __nanosleep(40);
      __syncwarp();
    }
    key_sides(sm, P, J, N, e, lane, widx);
    S64_ACC(28);   // trace: wait for the key chain + key sides
    if (lane == 0) {
      S.key[2 * e] = sm.hot.x;
      S.key[2 * e + 1] = sm.hot.y;
      S.wind_index[e] = widx;
    }
#ifdef S64_EARLY_RENDER  // measured: uint8 100 us per step against 94 with the frame drawn in the epilogue, float32 127 against 131
    if ((flags & GCA_FLAG_RENDER) && O.rgb != nullptr) {
      // the observation is drawn NOW, from the grid the step starts with (the frame's 12 / 48 KB of stores then overlap
      // the CA work instead of piling up at the end of the launch); the few cells that change are redrawn in the
      // epilogue.  New position (the action is known), pre-step dousing marks and day/night.
      uint32_t ri = 0;
      if (lane == 0) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        int row = S.position[2 * e], col = S.position[2 * e + 1];
        move_position(__float_as_int(sm.wind[9]), 64, 64, row, col);
        ri = (uint32_t)row | ((uint32_t)col << 8) | ((uint32_t)(S.is_night[e] != 0) << 16);
      }
      ri = __shfl_sync(GCA_FULL, ri, 0);
      if (O.rgb_u8) render_env64<true>(sm, wp32, O.rgb, e, t0, t1, f0, f1, (int)(ri & 255u), (int)((ri >> 8) & 255u), ri >> 16, lane);
      else render_env64<false>(sm, wp32, O.rgb, e, t0, t1, f0, f1, (int)(ri & 255u), (int)((ri >> 8) & 255u), ri >> 16, lane);
      sm.ign[2 * lane] = 0ull;   // (the render parked the fire rows there)
      sm.ign[2 * lane + 1] = 0ull;
      __syncwarp();
    }
#endif
  }

  const float lutreg = lane < 8 ? P.onep_veg[lane] : (lane < 16 ? P.onep_den[lane - 8] : 0.0f);
  const float w1 = P.ring_w[1], w2 = P.ring_w[2], w3 = P.ring_w[3], w4 = P.ring_w[4];

  // ================================ K CA sub-steps, all on-chip ===================================
  for (int j = 0; j < K; ++j) {
    if (active) {
      const uint32_t* sc = sm.sched[j];
      if (j > 0) {
        unsigned long long fr0, fr1;
        front_masks(t0, t1, f0, f1, lane, fr0, fr1);
        n_front += (uint32_t)(__popcll(fr0) + __popcll(fr1));
        if (!dense) {
          const unsigned long long listed0 = sm.listed[2 * lane], listed1 = sm.listed[2 * lane + 1];
          const unsigned long long nw0 = fr0 & ~listed0, nw1 = fr1 & ~listed1;
          const int nn = __popcll(nw0) + __popcll(nw1);
          const int incl_n = warp_incl_scan(nn, lane);
          const int tot_n = __shfl_sync(GCA_FULL, incl_n, 31);
          if (L + tot_n > S64_CAP) {
            dense = true;
          } else if (tot_n > 0) {
            int idx = L + incl_n - nn;
            unsigned long long m = nw0;
            int rowbits = (2 * lane) << 6;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
              while (m) {
                const uint32_t cell = (uint32_t)(rowbits | (__ffsll((long long)m) - 1));
                m &= m - 1;
                sm.list[idx++] = (uint16_t)cell;
                if (hidden != nullptr) {
                  prefetch_l1(hidden + cell_base + cell);
                  if (pslope != nullptr) prefetch_l1(pslope + (cell_base + cell) * 8);
                }
              }
              m = nw1;
              rowbits = (2 * lane + 1) << 6;
            }
            L += tot_n;
            if (nn) {
              sm.listed[2 * lane] = listed0 | nw0;
              sm.listed[2 * lane + 1] = listed1 | nw1;
            }
          }
        }
        if (dense) {
          T = build_front_list(sm, wp, fr0, fr1, lane, 0);
          prefetch_front(sm, hidden, pslope, cell_base, 0, min(T, S64_CAP), lane);
        }
      }
      if (j > 0) S64_ACC(29);
      if (lane == 0) sm.hot = make_uint4(sc[0], sc[1], sc[0] ^ sc[1] ^ 0x1BD11BDAu, (uint32_t)e);
      if (lane < 9) sm.wind[lane] = P.winds[(int)sc[8] * 9 + lane];
    }
    const int total = active ? (dense ? T : L) : 0;
    work += (uint32_t)total;

    for (int pass = 0;; ++pass) {
      if (active) {
        if (pass > 0 && total > pass * S64_CAP) {  // dense fires only: next slice of the list
          unsigned long long fr0, fr1;
          front_masks(t0, t1, f0, f1, lane, fr0, fr1);
          build_front_list(sm, wp, fr0, fr1, lane, pass * S64_CAP);
          prefetch_front(sm, hidden, pslope, cell_base, 0, min(S64_CAP, total - pass * S64_CAP), lane);
        }
      }
      S64_STAMP(4 + 3 * j);
      if (lane == 0) {
        const int cnt = max(0, min(S64_CAP, total - pass * S64_CAP));
        sm.cnt = cnt;
        // work items of this env: its 32-entry chunks, then (first pass of sub-step 0) its burn-out scan rounds of 4 rows
        const int nsc = (active && j == 0 && pass == 0) ? (sm.nscan + 3) >> 2 : 0;
        cs.nch[warp] = ((cnt + 31) >> 5) | (nsc << 16);
        if (gwarp == 0) cs.next[group] = 0;
      }
      S64_MARK();
      group_sync(group);
      S64_ACC(27);   // trace: wait at the barrier that opens the pooled phase (summed over the sub-steps)

      // ---------------- pooled phase: work items = 32-entry chunks of every env's front list -------
      {
        const int my_pk = lane < S64_GE ? cs.nch[group * S64_GE + lane] : 0;
        const int my_nc = my_pk & 0xFFFF;              // cell chunks of env slot `lane`
        const int my_n = my_nc + (my_pk >> 16);        // + its scan rounds
        int incl = my_n;
#pragma unroll
        for (int d = 1; d < S64_GE; d <<= 1) {
          const int o = __shfl_up_sync(GCA_FULL, incl, d);
          if (lane >= d) incl += o;
        }
        const int M = __shfl_sync(GCA_FULL, incl, S64_GE - 1);
        int PT = 0;
        bool more_items = M > 0;
        // the next work item is requested one item ahead, so the shared-memory atomic's round trip overlaps
        // with the work on the current item (lane 0 holds the ticket until it is needed)
#ifdef S64_STATIC_ITEMS
        int item_next = gwarp;  // item i goes to warp i mod S64_GE: no shared counter
#else
        int ticket = 0;
        if (lane == 0 && more_items) ticket = smem_add_ret(&cs.next[group], 1);
#endif
        for (;;) {
          if (PT < 32 && more_items) {
#ifdef S64_STATIC_ITEMS
            const int item = item_next;
            if (item >= M) { more_items = false; continue; }
            item_next += S64_GE;
#else
            const int item = __shfl_sync(GCA_FULL, ticket, 0);
            if (item >= M) { more_items = false; continue; }
            if (lane == 0) ticket = smem_add_ret(&cs.next[group], 1);
#endif
            const int gslot = __popc(__ballot_sync(GCA_FULL, lane < S64_GE && incl <= item));
            const int chunk = item - __shfl_sync(GCA_FULL, incl - my_n, gslot);
            const int es_slot = group * S64_GE + gslot;
            EnvSmem& es = cs.env[es_slot];
            const int nc_env = __shfl_sync(GCA_FULL, my_nc, gslot);
            if (chunk >= nc_env) {
              S64_MARK();
              scan_round(es, S, chunk - nc_env, K, lane);
              S64_ACC(30);
              continue;
            }
            S64_MARK();
            const int t = chunk * 32 + lane;
            const bool inrange = t < es.cnt;
            const uint32_t cell = inrange ? es.list[t] : 0u;
            const int r = cell >> 6, c = cell & 63;
            // the hidden byte is requested first: its (L2) latency hides behind the window work
            int hid = 3 | (3 << 3);
            if (hidden != nullptr && inrange) hid = hidden[(size_t)es.hot.w * 4096 + cell];
            uint32_t A, B, C;
            fire_window(es, r, c, A, B, C);
            uint32_t dirm = ((B >> 3) & 7u) | (((B >> 12) & 7u) << 3) | (((B >> 21) & 7u) << 6);
            // still a front cell in this sub-step?  (a listed tree leaves the front by igniting -- its own
            // fire bit is then set -- or by losing its last burning neighbour)
            const bool valid = inrange && !(dirm & 16u) && (dirm & ~16u) != 0u;
            dirm &= ~16u;
            // ring populations (Chebyshev rings 1..4 around the centre)
            const int S1 = __popc(B & 0x00E07038u);
            const int S2 = __popc(A & (0x07Cu << 18)) + __popc(B & 0x01F0F87Cu) + __popc(C & 0x07Cu);
            const int S3 = __popc(A & ((0x0FEu << 9) | (0x0FEu << 18))) + __popc(B & 0x03F9FCFEu) +
                           __popc(C & (0x0FEu | (0x0FEu << 9)));
            const int S4 = __popc(A) + __popc(B) + __popc(C);
            const float Hf = fmaf((float)(S4 - S3), w4,
                                  fmaf((float)(S3 - S2), w3, fmaf((float)(S2 - S1), w2, (float)S1 * w1)));
            float Dlo = 0.0f;
            bool near_doused = false;
            if ((((r & 1) ? es.dous_odd : es.dous_even) >> (r >> 1)) & 1u) {
              const uint32_t dwin = dous_window(es, r, c);
              if (dwin) {
                const int ni = __popc(dwin & ((0x0Eu << 5) | (0x0Eu << 10) | (0x0Eu << 15)));
                const int nb = __popc(dwin) - ni;
                const float Df = fmaf((float)nb, P.dous_border, (float)ni * P.dous_inner);
                Dlo = __fmul_rn(Df, S64_LO);
                near_doused = true;
              }
            }
            const float a = __shfl_sync(GCA_FULL, lutreg, clip15(hid & 7));
            const float b = __shfl_sync(GCA_FULL, lutreg, 8 + clip15((hid >> 3) & 7));
            const float ph_hi = __fsub_rn(__fmul_rn(Hf, S64_HI), Dlo);
            const float bhi = __fmul_rn(__fmul_rn(ph_hi, a), b);
            const int nd = (valid && bhi > 0.0f) ? __popc(dirm) : 0;
            const int incl2 = warp_incl_scan(nd, lane);
            int off = PT + incl2 - nd;
            const uint32_t em = nd ? dirm : 0u;
            // one record per (cell, burning direction): the cell's bound bhi rounded UP to a multiple of 32 ulp (2^-18
            // relative: inside the enclosure's margin) to make room for the env slot, sign bit = "dousing nearby";
            // and (cell << 4) | direction
            const uint32_t rec = ((__float_as_uint(bhi) + 31u) & 0x7FFFFFE0u) | (uint32_t)es_slot | (near_doused ? 0x80000000u : 0u);
            const uint32_t c16 = cell << 4;
#pragma unroll
            for (uint32_t d = 0; d < 9; ++d) {
              if (d == 4) continue;
              if (em & (1u << d)) {
                wp32[off] = rec;
                wpi[off] = (uint16_t)(c16 | d);
                ++off;
              }
            }
            PT += __shfl_sync(GCA_FULL, incl2, 31);
            __syncwarp();
            S64_ACC(31);
            continue;
          }
          if (PT == 0) break;
          // ---- draw up to 32 buffered pairs, one per lane: the record holds everything but the env's key.  (Two
          //      interleaved draws per lane were measured at the time the phase was bound by the integer pipe: no gain.)
          const int n = min(PT, 32);
          PT -= n;
          n_draws += (lane == 0) ? (uint32_t)n : 0u;
          const bool va = lane < n;
          const uint32_t ent = va ? wp32[PT + lane] : 0u;
          const uint32_t pk = va ? wpi[PT + lane] : 0u;
          EnvSmem& des = cs.env[ent & 31u];
          const uint4 hot = des.hot;
          const uint32_t cell = pk >> 4, d = pk & 15u;
          const uint32_t idx = cell * 9u + d;
          // direction's wind and slope factors: the global load's latency hides behind the threefry block
          float sl = 1.0f;
          if (pslope != nullptr && va) sl = pslope[((size_t)hot.w * 4096 + cell) * 8 + d - (d > 4u ? 1u : 0u)];
          const float wd = des.wind[d];
          float ua;
          if (j_u_burn) {
            const size_t inj0 = (size_t)j * N * 4096;
            ua = va ? j_u_burn[(inj0 + (size_t)hot.w * 4096) * 9 + idx] : 1.0f;
          } else {
            uint32_t o0, o1;
            TfKey key;
            key.k0 = hot.x; key.k1 = hot.y; key.k2 = hot.z;
            if (mode == GCA_RNG_LEGACY) {
              const bool first = idx < S64_HALF_BURN;
              const uint32_t c0 = first ? idx : idx - S64_HALF_BURN;
              threefry2x32(key, c0, c0 + S64_HALF_BURN, o0, o1);
              ua = bits_to_uniform(first ? o0 : o1);
            } else {
              threefry2x32(key, 0u, idx, o0, o1);
              ua = bits_to_uniform(o0 ^ o1);
            }
          }
          const float phi = __fmul_rn(__fmul_rn(__uint_as_float(ent & 0x7FFFFFE0u), wd), sl);
          if (va && ua < phi) pair_hit(P, hidden, des, ent, cell, wd, sl, phi, ua, n_thresh);
          __syncwarp();
        }
      }
      S64_STAMP(5 + 3 * j);
      const int more = group_sync_or(group, total > (pass + 1) * S64_CAP);
      S64_STAMP(6 + 3 * j);
      S64_MARK();
      if (!more) break;
    }

    // ---- apply: ignitions (with their fire-age draws), burn-outs, regrowth ------------------------
    if (active) {
      const uint32_t* sc = sm.sched[j];
      const size_t inj_base = ((size_t)j * N + e) * 4096;
      const unsigned long long I0 = sm.ign[2 * lane], I1 = sm.ign[2 * lane + 1];
      const int ni_l = __popcll(I0) + __popcll(I1);
      const int incl_i = warp_incl_scan(ni_l, lane);
      const int NI = __shfl_sync(GCA_FULL, incl_i, 31);
      if (NI > 0) {
        sm.ign[2 * lane] = 0ull;
        sm.ign[2 * lane + 1] = 0ull;
        // remember the ignited cells with their sub-step; the age draws are deferred to flush_ages
        int nign = sm.nign;
        int taken = 0;
        while (taken < NI) {
          const int take = min(S64_IGN_CAP - nign, NI - taken);
          int pos = nign + incl_i - ni_l - taken;  // list position of this lane's first ignition
          unsigned long long m = I0;
          uint32_t tag = ((uint32_t)j << 12) | ((uint32_t)(2 * lane) << 6);
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            while (m) {
              const uint32_t c = (uint32_t)__ffsll((long long)m) - 1u;
              m &= m - 1;
              if (pos >= nign && pos < nign + take) sm.ignlist[pos] = (uint16_t)(tag | c);
              ++pos;
            }
            m = I1;
            tag += 64u;
          }
          nign += take;
          taken += take;
          if (taken < NI) {  // the list is full (hundreds of ignitions in one env step): draw what it holds
            __syncwarp();
            flush_ages(sm, P, S, j_age_new, mode, N, e, nign, tick0, reinterpret_cast<uint32_t*>(wp), lane);
            rm.x = min(rm.x, reinterpret_cast<const uint32_t*>(wp)[2 * lane]);
            rm.y = min(rm.y, reinterpret_cast<const uint32_t*>(wp)[2 * lane + 1]);
            __syncwarp();
            nign = 0;
          }
        }
        __syncwarp();
        if (lane == 0) sm.nign = nign;
      }
      S64_ACC(27);
      if (j == 0) {  // the scan rounds of the pooled phase left the scanned rows' new minima in S.row_min
        if (burnrows & 1u) rm.x = __ldcg(S.row_min + (size_t)e * 64 + 2 * lane);
        if (burnrows & 2u) rm.y = __ldcg(S.row_min + (size_t)e * 64 + 2 * lane + 1);
      }
      unsigned long long ext0 = 0ull, ext1 = 0ull;
      if (burnrows & 1u) {
        const ulonglong2 a = reinterpret_cast<const ulonglong2*>(sm.burn[2 * lane])[0];
        const ulonglong2 b = reinterpret_cast<const ulonglong2*>(sm.burn[2 * lane])[1];
        ext0 = a.x & ((j & 1) ? a.y : ~a.y) & ((j & 2) ? b.x : ~b.x) & ((j & 4) ? b.y : ~b.y);
      }
      if (burnrows & 2u) {
        const ulonglong2 a = reinterpret_cast<const ulonglong2*>(sm.burn[2 * lane + 1])[0];
        const ulonglong2 b = reinterpret_cast<const ulonglong2*>(sm.burn[2 * lane + 1])[1];
        ext1 = a.x & ((j & 1) ? a.y : ~a.y) & ((j & 2) ? b.x : ~b.x) & ((j & 4) ? b.y : ~b.y);
      }
      unsigned long long g0 = 0, g1 = 0;
      if (P.p_tree > 0.0f) regrow_rows(P, J, tf_key(sc[2], sc[3]), inj_base, lane, ~(t0 | f0), ~(t1 | f1), g0, g1);
      n_ign += ni_l;
      n_ext += __popcll(ext0) + __popcll(ext1);
      ch0 |= I0 | ext0 | g0;
      ch1 |= I1 | ext1 | g1;
      t0 = (t0 & ~I0) | g0;
      t1 = (t1 & ~I1) | g1;
      f0 = (f0 & ~ext0) | I0;
      f1 = (f1 & ~ext1) | I1;
      if (j + 1 < K) {
        store_row_views(sm.fire32 + (2 * lane + 4) * 4, f0);
        store_row_views(sm.fire32 + (2 * lane + 5) * 4, f1);
      }
      __syncwarp();
    }
  }
  if (active && sm.nign > 0) {
    __syncwarp();
    flush_ages(sm, P, S, j_age_new, mode, N, e, sm.nign, tick0, reinterpret_cast<uint32_t*>(wp), lane);
    rm.x = min(rm.x, reinterpret_cast<const uint32_t*>(wp)[2 * lane]);
    rm.y = min(rm.y, reinterpret_cast<const uint32_t*>(wp)[2 * lane + 1]);
  }
  if (!active) {
    if (O.stats != nullptr) {
      const uint32_t b = __reduce_add_sync(GCA_FULL, n_draws), t = __reduce_add_sync(GCA_FULL, n_thresh);
      if (lane == 0) {
        if (b) atomicAdd(&O.stats[1], (unsigned long long)b);
        if (t) atomicAdd(&O.stats[4], (unsigned long long)t);
      }
    }
    return;
  }

#ifdef S64_TRACE
  const uint32_t trace_rows = __reduce_add_sync(GCA_FULL, __popc(burnrows));  // rows scanned for burn-outs
#endif
  // ---- sparse in-place write-back of the cells that changed --------------------------------------
  {
    unsigned long long ch = ch0;
    unsigned long long tt = t0, ff = f0;
    int row = 2 * lane;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      while (ch) {
        const int c = __ffsll((long long)ch) - 1;
        ch &= ch - 1;
        const uint8_t code = ((ff >> c) & 1ull) ? 2 : (((tt >> c) & 1ull) ? 1 : 0);
        S.cell[cell_base + row * 64 + c] = code;
      }
      ch = ch1;
      tt = t1; ff = f1;
      row = 2 * lane + 1;
    }
  }
  {
    // the bit-board copy of the grid: rows that changed
    ulonglong2* bbp = reinterpret_cast<ulonglong2*>(S.bb + (size_t)e * 128);
    if (ch0) bbp[2 * lane] = make_ulonglong2(t0, f0);
    if (ch1) bbp[2 * lane + 1] = make_ulonglong2(t1, f1);
  }
  reinterpret_cast<uint2*>(S.row_min + (size_t)e * 64)[lane] = rm;
  const int tcount = __reduce_add_sync(GCA_FULL, __popcll(t0) + __popcll(t1));
  const int fcount = __reduce_add_sync(GCA_FULL, __popcll(f0) + __popcll(f1));

  if (O.stats != nullptr) {
    const uint32_t a = __reduce_add_sync(GCA_FULL, n_front), b = __reduce_add_sync(GCA_FULL, n_draws);
    const uint32_t c = __reduce_add_sync(GCA_FULL, n_ign), d = __reduce_add_sync(GCA_FULL, n_ext);
    const uint32_t t = __reduce_add_sync(GCA_FULL, n_thresh);
    if (lane == 0) {
      atomicAdd(&O.stats[0], (unsigned long long)a);
      atomicAdd(&O.stats[1], (unsigned long long)b);
      atomicAdd(&O.stats[2], (unsigned long long)c);
      atomicAdd(&O.stats[3], (unsigned long long)d);
      if (t) atomicAdd(&O.stats[4], (unsigned long long)t);
      atomicAdd(&O.stats[5], 1ull);
    }
  }

  // ---- per-env scalars: clock, move, douse, day/night, reward, done (key / wind were stored above) --
  const bool ca_only = (flags & GCA_FLAG_CA_ONLY) != 0;
  const bool done = fcount == 0;
  if (!ca_only) {
    asm volatile("cp.async.wait_all;" ::: "memory");   // the action words (lanes 0..2 copied them)
    __syncwarp();
  }
  uint32_t rinfo = 0;  // observation inputs (lane 0): row | col << 8 | night << 16 | "add this step's dousing mark" << 17
  if (lane == 0) {
#ifdef S64_TRACE
    if (O.stats) {
      O.stats[8 + 32 * (size_t)e + 20] = work;
      O.stats[8 + 32 * (size_t)e + 21] = 0;
      O.stats[8 + 32 * (size_t)e + 22] = (unsigned long long)trace_rows;
      O.stats[8 + 32 * (size_t)e + 23] = (unsigned long long)(clock64() - clk0);
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      O.stats[8 + 32 * (size_t)e + 24] = smid;
      O.stats[8 + 32 * (size_t)e + 25] = gt;                       // end time, ns
      O.stats[8 + 32 * (size_t)e + 26] = (unsigned long long)blockIdx.x;
    }
#endif
    if (S.work != nullptr)
      S.work[e] = (flags & GCA_FLAG_WORK_CYCLES) ? (uint32_t)(clock64() - clk0) : work;
    S.tick[e] = tick0 + (uint32_t)K;
    const float rew = award(tcount, fcount);
    if (!ca_only) {
      // all loads first (they may alias the stores below as far as the compiler knows)
      const int a0 = __float_as_int(sm.wind[9]), a1 = __float_as_int(sm.wind[10]);
      const float t_old = S.time[e];
      int row = S.position[2 * e], col = S.position[2 * e + 1];
      const int ts = S.time_step[e] + 1;
      int night = S.is_night[e];
      const float se = S.steps_elapsed ? S.steps_elapsed[e] : 0.0f;
      const float ra = S.reward_accumulated ? S.reward_accumulated[e] : 0.0f;
      const int a0c = min(max(a0, 0), 8), a1c = min(max(a1, 0), 1);
      // clock (repeat_ca_jax.py:35-41): new = time + ((t_move + t_shoot) + t_any); keep the fraction
      const float tt = __fadd_rn(__fadd_rn(P.t_move[a0c], P.t_shoot[a1c]), P.t_any);
      const float nt = __fadd_rn(t_old, tt);
      move_position(a0, 64, 64, row, col);
      unsigned long long drow = 0ull;
      if (a1 == 1) drow = S.doused[(size_t)e * 64 + row];
      S.time[e] = __fsub_rn(nt, truncf(nt));
      S.position[2 * e] = row;
      S.position[2 * e + 1] = col;
      if (a1 == 1) S.doused[(size_t)e * 64 + row] = drow | (1ull << col);
      S.time_step[e] = ts;
      if (O.obs_night) O.obs_night[e] = (uint8_t)night;
      // the observation shows the NEW position with the PRE-step dousing marks and day/night (advanced_bulldozer.py:1120-1122)
      rinfo = (uint32_t)row | ((uint32_t)col << 8) | ((uint32_t)night << 16);
      if (ts % P.day_length == 0) night = 1 - night;
      if ((flags & GCA_FLAG_AUTO_RESET) && done)
        // ... unless the env resets now: conditional_reset redraws it from the restored grid and position with the
        // POST-step marks and day/night (:462-487); the mark of this step sits at the position just moved to
        rinfo = (uint32_t)row | ((uint32_t)col << 8) | ((uint32_t)night << 16) | (a1 == 1 ? 1u << 17 : 0u);
      S.is_night[e] = night;
      if (S.steps_elapsed) S.steps_elapsed[e] = __fadd_rn(se, 1.0f);
      if (S.reward_accumulated) S.reward_accumulated[e] = __fadd_rn(ra, rew);
    }
    if (O.step_reward) O.step_reward[e] = rew;
    if (O.terminated) O.terminated[e] = done ? 1 : 0;
    if (O.host_terminated && !O.host_done) O.host_terminated[e] = done ? 1 : 0;
    if (O.counts) { O.counts[2 * e] = tcount; O.counts[2 * e + 1] = fcount; }
    if (!((flags & GCA_FLAG_AUTO_RESET) && done)) {
      if (O.reward) O.reward[e] = rew;
      if (O.host_reward && !O.host_done) O.host_reward[e] = rew;
    }
  }

  // ---- observation, fused (GCA_FLAG_RENDER): MDP.grid_to_rgb (advanced_bulldozer.py:1035-1101) from the bit-boards --
  if ((flags & GCA_FLAG_RENDER) && O.rgb != nullptr) {
    rinfo = __shfl_sync(GCA_FULL, rinfo, 0);
    int prow = (int)(rinfo & 255u), pcol = (int)((rinfo >> 8) & 255u);
    const uint32_t night_obs = (rinfo >> 16) & 1u;
    unsigned long long rt0 = t0, rt1 = t1, rf0 = f0, rf1 = f1;
    const bool resets = (flags & GCA_FLAG_AUTO_RESET) && done;
    if (resets) {
      // this env resets now: the frame shows the restored grid and position (post-step marks and day/night)
      const ulonglong2* sb = reinterpret_cast<const ulonglong2*>(SNAP.bb + (size_t)e * 128);
      const ulonglong2 r0 = sb[2 * lane], r1 = sb[2 * lane + 1];
      rt0 = r0.x; rf0 = r0.y; rt1 = r1.x; rf1 = r1.y;
      if (lane == 0 && (rinfo & (1u << 17))) sm.dous64[2 + prow] |= 1ull << pcol;  // this step's mark (pre-restore position)
      prow = SNAP.position[2 * e];
      pcol = SNAP.position[2 * e + 1];
      __syncwarp();
    }
#ifdef S64_EARLY_RENDER
    if (!resets) {  // the prologue drew the frame from the grid the step started with: redraw the cells that changed
      if (O.rgb_u8) render_patch64<true>(sm, wp32, O.rgb, e, t0, t1, f0, f1, ch0, ch1, prow, pcol, night_obs, lane);
      else render_patch64<false>(sm, wp32, O.rgb, e, t0, t1, f0, f1, ch0, ch1, prow, pcol, night_obs, lane);
    } else
#endif
    {
      if (O.rgb_u8) render_env64<true>(sm, wp32, O.rgb, e, rt0, rt1, rf0, rf1, prow, pcol, night_obs, lane);
      else render_env64<false>(sm, wp32, O.rgb, e, rt0, rt1, rf0, rf1, prow, pcol, night_obs, lane);
    }
  }

  // ---- fused conditional_reset (advanced_bulldozer.py:422-518) ------------------------------------
  if ((flags & GCA_FLAG_AUTO_RESET) && done) {
    __syncwarp();
    const uint4* sc4 = reinterpret_cast<const uint4*>(SNAP.cell + cell_base);
    uint4* dc4 = reinterpret_cast<uint4*>(S.cell + cell_base);
#pragma unroll
    for (int i = 0; i < 8; ++i) dc4[i * 32 + lane] = sc4[i * 32 + lane];
    const uint4* sd4 = reinterpret_cast<const uint4*>(SNAP.death + cell_base);
    uint4* dd4 = reinterpret_cast<uint4*>(S.death + cell_base);
#pragma unroll
    for (int i = 0; i < 16; ++i) dd4[i * 32 + lane] = sd4[i * 32 + lane];
    reinterpret_cast<ulonglong2*>(S.doused + (size_t)e * 64)[lane] =
        reinterpret_cast<const ulonglong2*>(SNAP.doused + (size_t)e * 64)[lane];
    reinterpret_cast<uint2*>(S.row_min + (size_t)e * 64)[lane] =
        reinterpret_cast<const uint2*>(SNAP.row_min + (size_t)e * 64)[lane];
    {
      const ulonglong2* sb = reinterpret_cast<const ulonglong2*>(SNAP.bb + (size_t)e * 128);
      ulonglong2* db = reinterpret_cast<ulonglong2*>(S.bb + (size_t)e * 128);
      db[2 * lane] = sb[2 * lane];
      db[2 * lane + 1] = sb[2 * lane + 1];
    }
    if (lane == 0) {
      S.key[2 * e] = SNAP.key[2 * e];
      S.key[2 * e + 1] = SNAP.key[2 * e + 1];
      S.wind_index[e] = SNAP.wind_index[e];
      S.position[2 * e] = SNAP.position[2 * e];
      S.position[2 * e + 1] = SNAP.position[2 * e + 1];
      S.time[e] = SNAP.time[e];
      S.tick[e] = SNAP.tick[e];
      if (S.steps_elapsed) S.steps_elapsed[e] = 0.0f;
      if (S.reward_accumulated) S.reward_accumulated[e] = 0.0f;
      const float sr = snap_reward[e];
      if (O.reward) O.reward[e] = sr;
      if (O.host_reward && !O.host_done) O.host_reward[e] = sr;
    }
  }
  // ---- completion word (gca_env_step_host): the warp whose env ends last copies the step's results -- reward and
  //      terminated of ALL envs, from the device outputs -- to the mapped host mirrors in one burst, fences once at
  //      system scope and stores the token the host polls.  (Every env storing its own 5 bytes to host memory and
  //      fencing at system scope costs a PCIe round trip per warp on the kernel's tail.)
  if (O.host_done != nullptr) {
    uint32_t last = 0;
    if (lane == 0) {
      __threadfence();  // release: this env's device outputs
      last = atomicAdd(O.done_counter, 1u) == (uint32_t)N - 1u ? 1u : 0u;
    }
    last = __shfl_sync(GCA_FULL, last, 0);
    if (last) {
      __threadfence();  // acquire: the other envs' device outputs
      if (O.host_reward != nullptr && O.reward != nullptr)
        burst_copy_u32(reinterpret_cast<uint32_t*>(O.host_reward), reinterpret_cast<const uint32_t*>(O.reward), N, lane);
      if (O.host_terminated != nullptr && O.terminated != nullptr) {
        if ((N & 3) == 0 && ((reinterpret_cast<uintptr_t>(O.host_terminated) | reinterpret_cast<uintptr_t>(O.terminated)) & 3) == 0)
          burst_copy_u32(reinterpret_cast<uint32_t*>(O.host_terminated), reinterpret_cast<const uint32_t*>(O.terminated), N / 4, lane);
        else
          for (int i = lane; i < N; i += 32) O.host_terminated[i] = __ldcg(O.terminated + i);
      }
      __threadfence_system();
      __syncwarp();
      if (lane == 0) {
        *O.done_counter = 0u;
        *reinterpret_cast<volatile uint32_t*>(O.host_done) = O.done_token;
      }
    }
  }
}

template <int MODE, int HP, bool INJ>
cudaError_t launch_instance(const gca_params& p, const gca_state& s, const int32_t* actions, const gca_step_out& out,
                            const gca_inject& inj, const gca_state& snap, const float* snap_reward, uint32_t flags,
                            cudaStream_t st) {
  static bool configured = false;  // per instance; the attribute is per device function
  auto kern = env_step64_kernel<MODE, HP, INJ>;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaSmem));
    if (err != cudaSuccess) return err;
    err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (err != cudaSuccess) return err;
    configured = true;
  }
  const int blocks = (s.N + S64_E - 1) / S64_E;
  kern<<<dim3(blocks), dim3(S64_E * 32), sizeof(CtaSmem), st>>>(p, s, actions, out, inj, snap, snap_reward, flags);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_env_step64(const gca_params& p, const gca_state& s, const int32_t* actions,
                              const gca_step_out& out, const gca_inject& inj, const gca_state& snap,
                              const float* snap_reward, uint32_t flags, cudaStream_t st) {
  const bool injected = inj.u_burn || inj.u_grow || inj.age_new || inj.u_wind || inj.wind_step;
  const int hp = (!s.hidden && !s.pslope) ? 0 : ((s.hidden && s.pslope) ? 1 : -1);
#define GCA_LAUNCH64(M, H, I) return launch_instance<M, H, I>(p, s, actions, out, inj, snap, snap_reward, flags, st)
  if (injected || hp < 0) GCA_LAUNCH64(-1, -1, true);
  if (p.rng_mode == GCA_RNG_LEGACY) { if (hp) GCA_LAUNCH64(GCA_RNG_LEGACY, 1, false); else GCA_LAUNCH64(GCA_RNG_LEGACY, 0, false); }
  if (hp) GCA_LAUNCH64(GCA_RNG_PARTITIONABLE, 1, false);
  GCA_LAUNCH64(GCA_RNG_PARTITIONABLE, 0, false);
#undef GCA_LAUNCH64
}

}  // namespace gca
