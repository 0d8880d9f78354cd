// Exact polynomial arithmetic over F_r (BN254 scalar field, 2-adicity 28) by
// number-theoretic transforms.  Replaces the reference's schoolbook
// PolynomialField.Mul (r1csqap/r1csqap.go:57-67, O(n^2)) and the long division
// PolynomialField.Div / DivisorPolynomial (r1csqap.go:70-84,213-216, O(n^3) as
// written) on the prove path  h(x) = p(x) / Z(x)  (groth16/groth16.go:266).
// Field-exact, so every coefficient equals the reference's (for sizes where the
// reference can run at all).
//
// Transforms: forward = decimation-in-frequency (natural in, bit-reversed out),
// inverse = decimation-in-time (bit-reversed in, natural out); products are
// taken in the bit-reversed domain, so no permutation pass is ever needed.
// One kernel launch per radix-2 stage (HBM-bound: N*32 B read + written per stage).
//
// Division a = q*b + rem by power-series inversion of the reversed divisor:
//   rev(q) = rev(a) * rev(b)^-1  mod x^(deg a - deg b + 1)
// with the inverse series from Newton iteration g <- g*(2 - f*g).  For a fixed
// divisor (pk.Z) the transformed inverse is cached, leaving two transforms and
// one pointwise product per proof.
#pragma once
#include <cuda_runtime.h>

#include "fp.cuh"
#include "msm.cuh"  // ld_fe

namespace b200 {

// twiddle[j] = w^j, j in [0, half), w = root of order 2*half (Montgomery form)
__global__ void k_twiddles(Fr* tw, uint32_t half, Fr w) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= half) return;
  Fr r = Fr::one(), b = w;
  for (uint32_t e = j; e; e >>= 1) {
    if (e & 1) r = r * b;
    b = b.sqr();
  }
  tw[j] = r;
}

// DIF stage: len = 2*half butterflies spaced `half` apart; twiddle stride = N/len.
__global__ void k_ntt_dif_stage(Fr* __restrict__ a, const Fr* __restrict__ tw, uint32_t n_half, uint32_t half,
                                uint32_t tw_stride) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_half) return;
  uint32_t j = t & (half - 1);
  uint32_t i = ((t - j) << 1) + j;
  Fr u = a[i], v = a[i + half];
  a[i] = u + v;
  Fr d = u - v;
  a[i + half] = j ? d * tw[j * tw_stride] : d;
}
// DIT stage (inverse transform; tw holds powers of w^-1)
__global__ void k_ntt_dit_stage(Fr* __restrict__ a, const Fr* __restrict__ tw, uint32_t n_half, uint32_t half,
                                uint32_t tw_stride) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_half) return;
  uint32_t j = t & (half - 1);
  uint32_t i = ((t - j) << 1) + j;
  Fr u = a[i], v = a[i + half];
  if (j) v = v * tw[j * tw_stride];
  a[i] = u + v;
  a[i + half] = u - v;
}

// ---- EXPERIMENT (off by default, B200_NTT_FUSED): several radix-2 stages per pass in shared memory ----------------
// The per-stage kernels above read and write all N elements once per stage (21 passes over 64 MB for a 2^21-point
// transform).  Here a CTA owns a tile of 1024 elements — R = 2^k rows spaced h_bot apart by C = 1024 / R adjacent
// columns — and runs the k stages with halves h_bot * 2^(k-1) ... h_bot on it in shared memory (32 KB), so a 2^21-point
// transform is 3 passes (8 + 3 strided stages, then the last 10 stages on contiguous 1024-element tiles: k = 10, C = 1).
// Butterflies and twiddles are exactly those of k_ntt_dif_stage / k_ntt_dit_stage; the field results are identical.
// dit = 0: DIF order (halves descending);  dit = 1: DIT order (halves ascending, twiddle applied before the butterfly).
constexpr uint32_t kNttTile = 1024;
__global__ void __launch_bounds__(256) k_ntt_fused(Fr* __restrict__ a, const Fr* __restrict__ tw, uint32_t n_half,
                                                   uint32_t log_hbot, uint32_t k, int dit) {
  __shared__ Fr tile[kNttTile];
  const uint32_t logC = 10 - k, C = 1u << logC, R = 1u << k, h_bot = 1u << log_hbot;
  const uint32_t tiles_per_group = h_bot >> logC;  // column tiles inside one group of R * h_bot elements
  const uint32_t g = blockIdx.x / tiles_per_group, ct = blockIdx.x % tiles_per_group;
  const size_t base = (size_t)g * ((size_t)R << log_hbot) + ((size_t)ct << logC);
  const uint32_t col0 = ct << logC;
  const uint32_t t = threadIdx.x;
  for (uint32_t e = t; e < kNttTile; e += blockDim.x) {
    uint32_t r = e >> logC, c = e & (C - 1);
    tile[e] = a[base + ((size_t)r << log_hbot) + c];
  }
  __syncthreads();
  for (uint32_t st = 0; st < k; st++) {
    const uint32_t s = dit ? k - 1 - st : st;            // DIF: s = 0 is the largest half; DIT: smallest half first
    const uint32_t lb = k - 1 - s;                        // log2 of the row distance of a butterfly pair
    const uint32_t half = h_bot << lb;
    const uint32_t tw_stride = n_half / half;
    for (uint32_t bf = t; bf < kNttTile / 2; bf += blockDim.x) {
      uint32_t c = bf & (C - 1), q = bf >> logC;
      uint32_t low = q & ((1u << lb) - 1u), high = q >> lb;
      uint32_t r = (high << (lb + 1)) | low;
      uint32_t i0 = (r << logC) | c, i1 = i0 + ((1u << lb) << logC);
      uint32_t j = (low << log_hbot) + col0 + c;          // position inside the half: the twiddle exponent
      Fr u = tile[i0], v = tile[i1];
      if (dit) {
        if (j) v = v * tw[(size_t)j * tw_stride];
        tile[i0] = u + v;
        tile[i1] = u - v;
      } else {
        tile[i0] = u + v;
        Fr d = u - v;
        tile[i1] = j ? d * tw[(size_t)j * tw_stride] : d;
      }
    }
    __syncthreads();
  }
  for (uint32_t e = t; e < kNttTile; e += blockDim.x) {
    uint32_t r = e >> logC, c = e & (C - 1);
    a[base + ((size_t)r << log_hbot) + c] = tile[e];
  }
}
// The passes of a fused 2^logn-point transform (logn >= 10): up to 8 strided stages per pass above the last ten, then one
// contiguous pass.  Calls launch(log_hbot, k) in execution order (DIF: large halves first; DIT: the reverse).
template <class Launch>
inline void ntt_fused_passes(int logn, int dit, Launch&& launch, int max_k = 8) {
  int upper = logn - 10;                 // stages with half >= 2^10
  int ks[24], hb[24], np = 0;
  int top = logn - 1;                    // log2 of the largest half still to do
  while (upper > 0) {
    int k = upper > max_k ? max_k : upper;   // max_k < 8 only in tests (several strided passes on small transforms)
    ks[np] = k;
    hb[np] = top - (k - 1);
    np++;
    top -= k;
    upper -= k;
  }
  ks[np] = 10;
  hb[np] = 0;
  np++;
  if (!dit)
    for (int i = 0; i < np; i++) launch((uint32_t)hb[i], (uint32_t)ks[i]);
  else
    for (int i = np - 1; i >= 0; i--) launch((uint32_t)hb[i], (uint32_t)ks[i]);
}

// c[i] = a[i] * b[i] (* scale)
__global__ void k_pointwise_mul(Fr* __restrict__ a, const Fr* __restrict__ b, uint32_t n, Fr scale, int use_scale) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr v = a[i] * b[i];
  if (use_scale) v = v * scale;
  a[i] = v;
}

__global__ void k_scale(Fr* a, uint32_t n, Fr scale) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = a[i] * scale;
}

// dst[0..n_dst) <- Montgomery(src coefficients), optionally reversed, zero padded.
//   reverse == 0: dst[i] = src[i]            for i < n_take
//   reverse == 1: dst[i] = src[n_src-1-i]    for i < n_take
// src_mont: source already in Montgomery form.  err |= 2 if a coefficient >= r.
__global__ void k_poly_load(const Fr* __restrict__ src, uint32_t n_src, uint32_t n_take, int reverse, int src_mont,
                            Fr* __restrict__ dst, uint32_t n_dst, int* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_dst) return;
  if (i >= n_take) {
    dst[i] = Fr::zero();
    return;
  }
  Fr v = src[reverse ? n_src - 1 - i : i];
  if (!src_mont) {
    if (v.geq_modulus()) {
      atomicOr(err, 2);
      v = Fr::zero();
    }
    v = v.to_mont();
  }
  dst[i] = v;
}
// dst[i] = standard-form(src[reverse ? n-1-i : i]) for i < n  (src Montgomery)
__global__ void k_poly_store(const Fr* __restrict__ src, uint32_t n, int reverse, int to_std, Fr* __restrict__ dst) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr v = src[reverse ? n - 1 - i : i];
  dst[i] = to_std ? v.from_mont() : v;
}
// Newton step helper: t[i] = -t[i] for i < n, t[0] += 2 ; zero the tail [n, n_pad)
__global__ void k_two_minus(Fr* t, uint32_t n, uint32_t n_pad) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  if (i >= n) {
    t[i] = Fr::zero();
    return;
  }
  Fr v = t[i].neg();
  if (i == 0) v = v + Fr::one() + Fr::one();
  t[i] = v;
}
// g[0] = 1 / f[0]   (f Montgomery); flag |= 4 if f[0] == 0
__global__ void k_series_inv0(const Fr* f, Fr* g, int* err) {
  if (threadIdx.x | blockIdx.x) return;
  Fr v = f[0];
  if (v.is_zero()) {
    atomicOr(err, 4);
    g[0] = Fr::zero();
  } else {
    g[0] = v.inverse();
  }
}
// rem[i] = a[i] - qb[i]  (all Montgomery) -> standard form
__global__ void k_poly_sub_store(const Fr* __restrict__ a, const Fr* __restrict__ qb, uint32_t n, Fr* __restrict__ dst) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[i] = (a[i] - qb[i]).from_mont();
}

}  // namespace b200
