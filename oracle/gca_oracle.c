/* ORACLE (test infrastructure, not product code) -- dense C restatement of the advanced
 * bulldozer environment step of frasermince/gym-cellular-automata.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  It is the fast twin of oracle/alexandridis.py (same dense,
 * literal data flow: every cell draws its 12 random words, every window is summed in
 * row-major order in float32) and is itself checked against that NumPy restatement in
 * tests/test_oracle.py.  Parity status: PRNG pinned to Random123 / public JAX vectors; the NumPy twin is
 * pinned to rollouts of the reference's own Python source run under oracle/ref_shim (jax itself is not
 * installable; XLA's float32 summation order and jax.random's bits beyond the known answers stay unpinned).
 *
 * Reference lines (relative to /root/reference/gym_cellular_automata/):
 *   forest_fire/operators/ca_alexandridis_jax.py:164-206,321-460   CA update
 *   forest_fire/operators/repeat_ca_jax.py:34-71                    clock
 *   forest_fire/operators/move_modify_jax.py:39-62,102-114          move / douse
 *   forest_fire/bulldozer/advanced_bulldozer.py:332-399,597-633,1103-1133  MDP, reward, done
 * jax.random (threefry2x32, split, bits, uniform, randint) is restated from its published
 * algorithm, see oracle/prng.py.
 *
 * Build: gcc -O3 -march=native -fno-fast-math -ffp-contract=off -fopenmp -shared -fPIC
 * (oracle/c_oracle.py:build).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXR 10
#define MAXWIN (2 * MAXR + 1)

typedef struct {
  int32_t H, W, R, K, rng_mode; /* rng_mode 0 = legacy layout, 1 = partitionable */
  int32_t age_lo, age_span, age_mult; /* randint(fire_age_min, fire_age_max) */
  int32_t day_length;
  float p_tree, p_wind_change, t_any;
  float t_move[9], t_shoot[2];
  float onep_veg[6], onep_den[6]; /* f32(1) + p_veg[i] */
  float winds[8 * 9];             /* wind_matrix of winds[k] (centre 0) */
  float dousing_weights[25];
  float burn_kernel[MAXWIN * MAXWIN]; /* (2R+1)^2 used, row-major */
} oracle_params;

/* optional injected random fields, each with a leading K axis; NULL = draw with threefry */
typedef struct {
  const float *u_burn;      /* (K,N,H,W,9) */
  const float *u_grow;      /* (K,N,H,W)   */
  const int32_t *age_new;   /* (K,N,H,W)   */
  const float *u_wind;      /* (K,N)       */
  const int32_t *wind_step; /* (K,N)       */
} oracle_inject;

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static inline void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t *o0,
                                uint32_t *o1) {
  static const int rot[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
  for (int g = 0; g < 5; ++g) {
    for (int q = 0; q < 4; ++q) {
      x0 += x1;
      x1 = rotl32(x1, rot[g & 1][q]);
      x1 ^= x0;
    }
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
  *o0 = x0;
  *o1 = x1;
}

/* 16 independent blocks per call, written round-by-round so gcc vectorises across lanes */
#define TFL 16
#define TF_ROUND(r)                                   \
  for (int l = 0; l < TFL; ++l) {                     \
    x0[l] += x1[l];                                   \
    x1[l] = (x1[l] << (r)) | (x1[l] >> (32 - (r)));   \
    x1[l] ^= x0[l];                                   \
  }
#define TF_INJECT(a, b, i)                            \
  for (int l = 0; l < TFL; ++l) {                     \
    x0[l] += (a);                                     \
    x1[l] += (b) + (uint32_t)(i);                     \
  }
static void threefry2x32_x16(uint32_t k0, uint32_t k1, uint32_t *restrict x0, uint32_t *restrict x1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  TF_INJECT(k0, k1, 0)
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6) TF_INJECT(k1, k2, 1)
  TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24) TF_INJECT(k2, k0, 2)
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6) TF_INJECT(k0, k1, 3)
  TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24) TF_INJECT(k1, k2, 4)
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6) TF_INJECT(k2, k0, 5)
}

/* jax.random.bits(key, (n,)) dense */
static void random_bits(const uint32_t key[2], int64_t n, int mode, uint32_t *out) {
  uint32_t x0[TFL], x1[TFL];
  if (mode == 0) {
    int64_t m = n + (n & 1), h = m / 2;
    for (int64_t b0 = 0; b0 < h; b0 += TFL) {
      for (int l = 0; l < TFL; ++l) {
        int64_t b = b0 + l;
        x0[l] = (uint32_t)b;
        x1[l] = (uint32_t)(b + h);
        if ((n & 1) && b + h == m - 1) x1[l] = 0; /* padded counter */
      }
      threefry2x32_x16(key[0], key[1], x0, x1);
      for (int l = 0; l < TFL; ++l) {
        int64_t b = b0 + l;
        if (b < h) {
          out[b] = x0[l];
          if (b + h < n) out[b + h] = x1[l];
        }
      }
    }
  } else {
    for (int64_t i0 = 0; i0 < n; i0 += TFL) {
      for (int l = 0; l < TFL; ++l) {
        uint64_t i = (uint64_t)(i0 + l);
        x0[l] = (uint32_t)(i >> 32);
        x1[l] = (uint32_t)i;
      }
      threefry2x32_x16(key[0], key[1], x0, x1);
      for (int l = 0; l < TFL; ++l)
        if (i0 + l < n) out[i0 + l] = x0[l] ^ x1[l];
    }
  }
}

/* key, subkey = jax.random.split(key) */
static void split2(const uint32_t key[2], int mode, uint32_t newkey[2], uint32_t sub[2]) {
  if (mode == 0) {
    uint32_t o[4];
    random_bits(key, 4, 0, o);
    newkey[0] = o[0]; newkey[1] = o[1]; sub[0] = o[2]; sub[1] = o[3];
  } else {
    threefry2x32(key[0], key[1], 0, 0, &newkey[0], &newkey[1]);
    threefry2x32(key[0], key[1], 0, 1, &sub[0], &sub[1]);
  }
}

static inline float bits_to_uniform(uint32_t b) {
  union { uint32_t u; float f; } v;
  v.u = (b >> 9) | 0x3F800000u;
  return v.f - 1.0f;
}

static inline int32_t randint_from_bits(uint32_t hb, uint32_t lb, int32_t lo, uint32_t span, uint32_t mult) {
  uint32_t off = (hb % span) * mult + (lb % span);
  off %= span;
  return lo + (int32_t)off;
}

typedef struct {
  float *fire_pad, *dous_pad, *heat, *dous, *ub, *ug, *new_grid, *new_age;
  uint32_t *bits, *hb, *lb;
} scratch_t;

static scratch_t scratch_alloc(int H, int W) {
  scratch_t s;
  size_t n = (size_t)H * W;
  size_t np_ = (size_t)(H + 2 * MAXR) * (W + 2 * MAXR);
  s.fire_pad = (float *)malloc(np_ * 4);
  s.dous_pad = (float *)malloc(np_ * 4);
  s.heat = (float *)malloc(n * 4);
  s.dous = (float *)malloc(n * 4);
  s.ub = (float *)malloc(n * 9 * 4);
  s.ug = (float *)malloc(n * 4);
  s.new_grid = (float *)malloc(n * 4);
  s.new_age = (float *)malloc(n * 4);
  s.bits = (uint32_t *)malloc(n * 9 * 4);
  s.hb = (uint32_t *)malloc(n * 4);
  s.lb = (uint32_t *)malloc(n * 4);
  return s;
}
static void scratch_free(scratch_t *s) {
  free(s->fire_pad); free(s->dous_pad); free(s->heat); free(s->dous); free(s->ub); free(s->ug);
  free(s->new_grid); free(s->new_age); free(s->bits); free(s->hb); free(s->lb);
}

/* row-major sequential float32 window sum, accumulator from +0 */
static void window_sum(const float *pad, int pw, int n, const float *wts, int H, int W, float *acc) {
  int size = 2 * n + 1;
  for (int i = 0; i < H * W; ++i) acc[i] = 0.0f;
  for (int i = 0; i < size; ++i)
    for (int j = 0; j < size; ++j) {
      float w = wts[i * size + j];
      for (int r = 0; r < H; ++r) {
        const float *src = pad + (size_t)(r + i) * pw + j;
        float *dst = acc + (size_t)r * W;
        for (int c = 0; c < W; ++c) dst[c] = dst[c] + src[c] * w;
      }
    }
}

/* one CA update of one env: PartiallyObservableForestFireJax.update */
static void ca_update_env(const oracle_params *P, scratch_t *S, float *grid, float *fire_age,
                          const int32_t *dousing, const int32_t *veg, const int32_t *den,
                          const float *pslope, int32_t *wind_index, uint32_t key[2],
                          const float *inj_ub, const float *inj_ug, const int32_t *inj_age,
                          const float *inj_uw, const int32_t *inj_ws, float *p_out) {
  const int H = P->H, W = P->W, R = P->R, mode = P->rng_mode;
  const int64_t n = (int64_t)H * W;
  uint32_t K1[2], S1[2], Ka[2], Sburn[2], Kb[2], Sgrow[2], Kc[2], Sage[2], K2[2], Swind[2], K3[2], Sidx[2];
  split2(key, mode, K1, S1);
  split2(S1, mode, Ka, Sburn);
  split2(Ka, mode, Kb, Sgrow);
  split2(Kb, mode, Kc, Sage);
  split2(K1, mode, K2, Swind);
  split2(K2, mode, K3, Sidx);
  const float *wind = P->winds + 9 * (*wind_index);

  /* padded masks (jnp.pad constant 0) */
  {
    int pw = W + 2 * R;
    memset(S->fire_pad, 0, (size_t)(H + 2 * R) * pw * 4);
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < W; ++c) S->fire_pad[(size_t)(r + R) * pw + c + R] = grid[r * W + c] == 2.0f ? 1.0f : 0.0f;
    window_sum(S->fire_pad, pw, R, P->burn_kernel, H, W, S->heat);
    int pd = W + 4;
    memset(S->dous_pad, 0, (size_t)(H + 4) * pd * 4);
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < W; ++c) S->dous_pad[(size_t)(r + 2) * pd + c + 2] = (float)dousing[r * W + c];
    window_sum(S->dous_pad, pd, 2, P->dousing_weights, H, W, S->dous);
  }
  /* dense draws */
  const float *ub = inj_ub;
  if (!ub) {
    random_bits(Sburn, 9 * n, mode, S->bits);
    for (int64_t i = 0; i < 9 * n; ++i) S->ub[i] = bits_to_uniform(S->bits[i]);
    ub = S->ub;
  }
  const float *ug = inj_ug;
  if (!ug) {
    random_bits(Sgrow, n, mode, S->bits);
    for (int64_t i = 0; i < n; ++i) S->ug[i] = bits_to_uniform(S->bits[i]);
    ug = S->ug;
  }
  if (!inj_age) {
    uint32_t k1[2], k2[2];
    split2(Sage, mode, k1, k2);
    random_bits(k1, n, mode, S->hb);
    random_bits(k2, n, mode, S->lb);
  }
  /* rule */
  for (int r = 0; r < H; ++r)
    for (int c = 0; c < W; ++c) {
      int idx = r * W + c;
      float g = grid[idx];
      int vi = veg[idx] < 1 ? 1 : (veg[idx] > 5 ? 5 : veg[idx]);
      int di = den[idx] < 1 ? 1 : (den[idx] > 5 ? 5 : den[idx]);
      float ph = S->heat[idx] - S->dous[idx];
      float base = ph * P->onep_veg[vi];
      base = base * P->onep_den[di];
      int hit = 0;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          float p = base * wind[i * 3 + j];
          p = p * pslope[(size_t)idx * 9 + i * 3 + j];
          if (p_out) p_out[(size_t)idx * 9 + i * 3 + j] = p;
          int rr = r - 1 + i, cc = c - 1 + j;
          int nb_fire = (rr >= 0 && rr < H && cc >= 0 && cc < W) ? (grid[rr * W + cc] == 2.0f) : 0;
          if (nb_fire && ub[(size_t)idx * 9 + i * 3 + j] < p) hit = 1;
        }
      float ng;
      if (g == 1.0f && hit) ng = 2.0f;
      else if (g == 0.0f && ug[idx] < P->p_tree) ng = 1.0f;
      else if (g == 2.0f && fire_age[idx] <= 1.0f) ng = 0.0f;
      else ng = g;
      float na = fire_age[idx];
      if (ng == 2.0f && g != 2.0f) {
        int32_t a = inj_age ? inj_age[idx]
                            : randint_from_bits(S->hb[idx], S->lb[idx], P->age_lo, (uint32_t)P->age_span,
                                                (uint32_t)P->age_mult);
        na = (float)a;
      }
      if (g == 2.0f) na = na - 1.0f;
      S->new_grid[idx] = ng;
      S->new_age[idx] = na;
    }
  memcpy(grid, S->new_grid, n * 4);
  memcpy(fire_age, S->new_age, n * 4);
  /* wind random walk */
  float uw;
  if (inj_uw) uw = *inj_uw;
  else {
    uint32_t b;
    random_bits(Swind, 1, mode, &b);
    uw = bits_to_uniform(b);
  }
  int32_t ws;
  if (inj_ws) ws = *inj_ws;
  else {
    uint32_t k1[2], k2[2], hb, lb;
    split2(Sidx, mode, k1, k2);
    random_bits(k1, 1, mode, &hb);
    random_bits(k2, 1, mode, &lb);
    ws = randint_from_bits(hb, lb, 1, 7, 4);
  }
  if (uw < P->p_wind_change) *wind_index = (*wind_index + ws) % 8;
  key[0] = K3[0];
  key[1] = K3[1];
}

/* One env step (MDP.update + reward + done) for N envs, reference-layout arrays, in place.
 * action: (N,3) int32 [move, shoot, ext].  p_out (optional): (N,H,W,9) burn probabilities of the
 * LAST sub-step.  Returns 0. */
int oracle_env_step(const oracle_params *P, int32_t N, float *grid, float *fire_age, int32_t *dousing,
                    const int32_t *veg, const int32_t *den, const float *pslope, int32_t *wind_index,
                    uint32_t *key, int32_t *is_night, int32_t *time_step, int32_t *position, float *time,
                    const int32_t *action, float *reward, uint8_t *terminated, int32_t *counts,
                    const oracle_inject *inj, float *p_out, int32_t nthreads) {
  const int H = P->H, W = P->W, K = P->K;
  const size_t n = (size_t)H * W;
  if (P->R > MAXR) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    static __thread scratch_t S;
    static __thread int sH = 0, sW = 0;
    if (sH != H || sW != W) {
      if (sH) scratch_free(&S);
      S = scratch_alloc(H, W);
      sH = H; sW = W;
    }
#pragma omp for schedule(dynamic, 1)
    for (int e = 0; e < N; ++e) {
      int a0 = action[e * 3 + 0], a1 = action[e * 3 + 1];
      /* clock */
      float t_action = P->t_move[a0] + P->t_shoot[a1];
      float t_taken = t_action + P->t_any;
      float nt = time[e] + t_taken;
      time[e] = nt - truncf(nt);
      for (int k = 0; k < K; ++k) {
        size_t o = (size_t)k * N + e;
        ca_update_env(P, &S, grid + e * n, fire_age + e * n, dousing + e * n, veg + e * n, den + e * n,
                      pslope + e * n * 9, wind_index + e, key + 2 * e,
                      (inj && inj->u_burn) ? inj->u_burn + o * n * 9 : NULL,
                      (inj && inj->u_grow) ? inj->u_grow + o * n : NULL,
                      (inj && inj->age_new) ? inj->age_new + o * n : NULL,
                      (inj && inj->u_wind) ? inj->u_wind + o : NULL,
                      (inj && inj->wind_step) ? inj->wind_step + o : NULL,
                      (p_out && k == K - 1) ? p_out + e * n * 9 : NULL);
      }
      /* move */
      int row = position[2 * e], col = position[2 * e + 1];
      int vu = row > 0, vd = row < H - 1, vl = col > 0, vr = col < W - 1;
      if ((a0 == 0 || a0 == 1 || a0 == 2) && vu) row -= 1;
      if ((a0 == 6 || a0 == 7 || a0 == 8) && vd) row += 1;
      if ((a0 == 0 || a0 == 3 || a0 == 6) && vl) col -= 1;
      if ((a0 == 2 || a0 == 5 || a0 == 8) && vr) col += 1;
      position[2 * e] = row;
      position[2 * e + 1] = col;
      /* douse */
      if (a1 == 1) dousing[e * n + (size_t)row * W + col] = 1;
      time_step[e] += 1;
      if (time_step[e] % P->day_length == 0) is_night[e] = 1 - is_night[e];
      int32_t t = 0, f = 0;
      for (size_t i = 0; i < n; ++i) {
        t += grid[e * n + i] == 1.0f;
        f += grid[e * n + i] == 2.0f;
      }
      float denom = (float)(t + f) + 1e-8f;
      reward[e] = -((float)f / denom);
      terminated[e] = (f == 0);
      if (counts) { counts[2 * e] = t; counts[2 * e + 1] = f; }
    }
  }
  return 0;
}

/* test hooks */
void oracle_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t *out) {
  threefry2x32(k0, k1, x0, x1, &out[0], &out[1]);
}
void oracle_random_bits(const uint32_t *key, int64_t n, int32_t mode, uint32_t *out) { random_bits(key, n, mode, out); }
void oracle_split(const uint32_t *key, int32_t mode, uint32_t *out4) { split2(key, mode, out4, out4 + 2); }
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
