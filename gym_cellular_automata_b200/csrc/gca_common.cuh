// gca_common.cuh -- device helpers shared by the libgca kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gca.h"

#define GCA_FULL 0xFFFFFFFFu
#ifndef GCA_S64_GROUPS
#define GCA_S64_GROUPS 1  /* independent lock-step groups inside a CTA of env_step64_kernel (own named barrier each) */
#endif
#ifndef GCA_S64_WARPS
#define GCA_S64_WARPS 14  /* warps (= envs) per CTA of env_step64_kernel: 2 CTAs x 14 warps x 72 registers fill an SM */
#endif

namespace gca {

// ---------------------------------------------------------------------------------------------
// threefry2x32-20 (Random123), the block function behind jax.random (see oracle/prng.py).
// The key-schedule word of each injection is pre-added by the caller-side struct so that the
// per-block work is 20 rounds + 6 injections.
// ---------------------------------------------------------------------------------------------
struct TfKey {
  uint32_t k0, k1, k2;  // k2 = k0 ^ k1 ^ 0x1BD11BDA
};

__device__ __forceinline__ TfKey tf_key(uint32_t k0, uint32_t k1) {
  TfKey k;
  k.k0 = k0;
  k.k1 = k1;
  k.k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  return k;
}

#define GCA_TF_ROUND(r)            \
  x0 += x1;                        \
  x1 = __funnelshift_l(x1, x1, r); \
  x1 ^= x0;

__device__ __forceinline__ void threefry2x32(const TfKey& k, uint32_t x0, uint32_t x1, uint32_t& o0,
                                             uint32_t& o1) {
  x0 += k.k0;
  x1 += k.k1;
  GCA_TF_ROUND(13) GCA_TF_ROUND(15) GCA_TF_ROUND(26) GCA_TF_ROUND(6)
  x0 += k.k1; x1 += k.k2 + 1u;
  GCA_TF_ROUND(17) GCA_TF_ROUND(29) GCA_TF_ROUND(16) GCA_TF_ROUND(24)
  x0 += k.k2; x1 += k.k0 + 2u;
  GCA_TF_ROUND(13) GCA_TF_ROUND(15) GCA_TF_ROUND(26) GCA_TF_ROUND(6)
  x0 += k.k0; x1 += k.k1 + 3u;
  GCA_TF_ROUND(17) GCA_TF_ROUND(29) GCA_TF_ROUND(16) GCA_TF_ROUND(24)
  x0 += k.k1; x1 += k.k2 + 4u;
  GCA_TF_ROUND(13) GCA_TF_ROUND(15) GCA_TF_ROUND(26) GCA_TF_ROUND(6)
  x0 += k.k2; x1 += k.k0 + 5u;
  o0 = x0;
  o1 = x1;
}

// Out-of-line copy for the places that are not throughput critical (key schedule, fire-age draws,
// regrowth): keeps the kernel's hot code inside the instruction cache.
static __device__ __noinline__ void threefry2x32_ni(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t& o0,
                                             uint32_t& o1) {
  threefry2x32(tf_key(k0, k1), x0, x1, o0, o1);
}

__device__ __forceinline__ uint32_t bits_at_ni(const TfKey& k, uint32_t idx, uint32_t half, int mode) {
  uint32_t o0, o1;
  if (mode == GCA_RNG_LEGACY) {
    const bool first = idx < half;
    threefry2x32_ni(k.k0, k.k1, first ? idx : idx - half, first ? idx + half : idx, o0, o1);
    return first ? o0 : o1;
  }
  threefry2x32_ni(k.k0, k.k1, 0u, idx, o0, o1);
  return o0 ^ o1;
}

// Element `idx` of jax.random.bits(key, (n,)) with n even and half = n / 2.
//   legacy:        idx <  half -> word 0 of block (idx, idx + half)
//                  idx >= half -> word 1 of block (idx - half, idx)
//   partitionable: word0 ^ word1 of block (0, idx)            (n < 2^32)
__device__ __forceinline__ uint32_t bits_at(const TfKey& k, uint32_t idx, uint32_t half, int mode) {
  uint32_t o0, o1;
  if (mode == GCA_RNG_LEGACY) {
    const bool first = idx < half;
    threefry2x32(k, first ? idx : idx - half, first ? idx + half : idx, o0, o1);
    return first ? o0 : o1;
  }
  threefry2x32(k, 0u, idx, o0, o1);
  return o0 ^ o1;
}

// jax.random.bits(key, ()) -- one word: legacy pads the odd count with counter 0 -> block (0,0) word 0
__device__ __forceinline__ uint32_t bits_scalar(const TfKey& k, int mode) {
  uint32_t o0, o1;
  threefry2x32(k, 0u, 0u, o0, o1);
  return mode == GCA_RNG_LEGACY ? o0 : (o0 ^ o1);
}

// float32 uniform in [0,1): (bits >> 9) * 2^-23 exactly (jax.random.uniform)
__device__ __forceinline__ float bits_to_uniform(uint32_t b) {
  return __uint_as_float((b >> 9) | 0x3F800000u) - 1.0f;
}

__device__ __forceinline__ int32_t randint_from_bits(uint32_t hb, uint32_t lb, int32_t lo, uint32_t span,
                                                     uint32_t mult) {
  uint32_t off = (hb % span) * mult + (lb % span);
  off %= span;
  return lo + (int32_t)off;
}

// key, subkey = jax.random.split(key), computed by ONE thread (2 blocks).
__device__ __forceinline__ void split_thread(uint32_t k0, uint32_t k1, int mode, uint32_t& n0, uint32_t& n1,
                                             uint32_t& s0, uint32_t& s1) {
  const TfKey k = tf_key(k0, k1);
  uint32_t a0, a1, b0, b1;
  if (mode == GCA_RNG_LEGACY) {
    threefry2x32(k, 0u, 2u, a0, a1);
    threefry2x32(k, 1u, 3u, b0, b1);
    n0 = a0; n1 = b0; s0 = a1; s1 = b1;
  } else {
    threefry2x32(k, 0u, 0u, a0, a1);
    threefry2x32(k, 0u, 1u, b0, b1);
    n0 = a0; n1 = a1; s0 = b0; s1 = b1;
  }
}

// key, subkey = jax.random.split(key) computed by a lane PAIR (lanes 2p, 2p+1 hold the same key):
// each lane runs one of the two blocks and they swap words.  All 32 lanes must call.
__device__ __forceinline__ void split_pair(uint32_t k0, uint32_t k1, int mode, int lane, uint32_t& n0,
                                           uint32_t& n1, uint32_t& s0, uint32_t& s1) {
  const uint32_t w = lane & 1;
  uint32_t o0, o1;
  threefry2x32_ni(k0, k1, mode == GCA_RNG_LEGACY ? w : 0u, mode == GCA_RNG_LEGACY ? w + 2u : w, o0, o1);
  const uint32_t p0 = __shfl_xor_sync(GCA_FULL, o0, 1);
  const uint32_t p1 = __shfl_xor_sync(GCA_FULL, o1, 1);
  const uint32_t a0 = w ? p0 : o0, a1 = w ? p1 : o1;  // block A (even lane)
  const uint32_t b0 = w ? o0 : p0, b1 = w ? o1 : p1;  // block B (odd lane)
  if (mode == GCA_RNG_LEGACY) { n0 = a0; n1 = b0; s0 = a1; s1 = b1; }
  else { n0 = a0; n1 = a1; s0 = b0; s1 = b1; }
}

// one threefry block per lane with lane-specific key/counters, words swapped inside the lane pair
__device__ __forceinline__ void tf_exchange(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t& o0,
                                            uint32_t& o1, uint32_t& p0, uint32_t& p1) {
  threefry2x32_ni(k0, k1, c0, c1, o0, o1);
  p0 = __shfl_xor_sync(GCA_FULL, o0, 1);
  p1 = __shfl_xor_sync(GCA_FULL, o1, 1);
}
__device__ __forceinline__ void assemble_split(int mode, uint32_t w, uint32_t o0, uint32_t o1, uint32_t p0,
                                               uint32_t p1, uint32_t& n0, uint32_t& n1, uint32_t& s0,
                                               uint32_t& s1) {
  const uint32_t a0 = w ? p0 : o0, a1 = w ? p1 : o1;
  const uint32_t b0 = w ? o0 : p0, b1 = w ? o1 : p1;
  if (mode == GCA_RNG_LEGACY) { n0 = a0; n1 = b0; s0 = a1; s1 = b1; }
  else { n0 = a0; n1 = a1; s0 = b0; s1 = b1; }
}

// Reward of _award: -(f / (t + f + 1e-8)) in float32 (advanced_bulldozer.py:627-630)
__device__ __forceinline__ float award(int t, int f) {
  const float denom = __fadd_rn((float)(t + f), 1e-8f);
  return -__fdiv_rn((float)f, denom);
}

// MoveJax.update (move_modify_jax.py:49-57)
__device__ __forceinline__ void move_position(int a0, int H, int W, int& row, int& col) {
  const bool vu = row > 0, vd = row < H - 1, vl = col > 0, vr = col < W - 1;
  if (a0 <= 2 && a0 >= 0 && vu) row -= 1;
  if (a0 >= 6 && a0 <= 8 && vd) row += 1;
  if ((a0 == 0 || a0 == 3 || a0 == 6) && vl) col -= 1;
  if ((a0 == 2 || a0 == 5 || a0 == 8) && vr) col += 1;
}

// One pixel: display value, palette, dousing tint, bulldozer -- the float32 arithmetic of grid_to_rgb
// (advanced_bulldozer.py:1035-1101), shared by both render kernels.
__device__ __forceinline__ void render_pixel(int disp, bool ng, bool ds, bool dozer, float& cr, float& cg, float& cb) {
  if (!ng) {
    if (disp == 1) { cr = 169.f; cg = 196.f; cb = 153.f; }       // #A9C499
    else if (disp == 2) { cr = 230.f; cg = 129.f; cb = 129.f; }  // #E68181
    else { cr = 221.f; cg = 209.f; cb = 211.f; }                 // #DDD1D3
  } else {
    if (disp == 1) { cr = 47.f; cg = 79.f; cb = 79.f; }          // #2F4F4F
    else if (disp == 2) { cr = 139.f; cg = 0.f; cb = 0.f; }      // #8B0000
    else { cr = 105.f; cg = 105.f; cb = 105.f; }                 // #696969
  }
  if (ds) {  // rgb * (1 - 0.75) + tint * 0.75, float32, tint blue by day / orange by night
    const float tr = ng ? 255.f : 0.f, tg = ng ? 165.f : 0.f, tb = ng ? 0.f : 200.f;
    cr = __fadd_rn(__fmul_rn(cr, 0.25f), __fmul_rn(tr, 0.75f));
    cg = __fadd_rn(__fmul_rn(cg, 0.25f), __fmul_rn(tg, 0.75f));
    cb = __fadd_rn(__fmul_rn(cb, 0.25f), __fmul_rn(tb, 0.75f));
  }
  if (dozer) { cr = 0.f; cg = 0.f; cb = 0.f; }
}

__device__ __forceinline__ int clip15(int v) { return v < 1 ? 1 : (v > 5 ? 5 : v); }

// direction index d = i*3+j (0..8, 4 = centre) -> slot in the 8-entry slope-factor table
__device__ __forceinline__ int dir_slot(int d) { return d < 4 ? d : d - 1; }

}  // namespace gca

// internal launchers (one per translation unit), called by gca_abi.cu
namespace gca {
cudaError_t launch_env_step64(const gca_params& p, const gca_state& s, const int32_t* actions,
                              const gca_step_out& out, const gca_inject& inj, const gca_state& snap,
                              const float* snap_reward, uint32_t flags, cudaStream_t st);
}  // namespace gca
