// Bucket accumulation by BATCHED AFFINE ADDITION (the default hot path for large MSMs).
//
// An affine addition costs 1 inversion + 2M + 1S; with Montgomery's trick a batch
// of denominators shares one inversion at 3M each, so a bucket add costs ~6.5 field
// multiplications instead of the 10 (8M + 2S) of the XYZZ mixed add — and the
// multiply pipe (IMAD.WIDE, quarter rate on sm_100) is what bounds this kernel
// (DESIGN.md §4).
//
// Shape.  Buckets are cut into slices of exactly S = 2^R leaf slots (the tail of a
// bucket is padded with the point at infinity), so every slice is a perfect
// binary tree: round r (1..R) has S >> r independent pair additions per slice and
// pair p of the whole MSM is (slice p / q, position p % q), q = S >> r — no
// ragged bookkeeping.  Each round is three launches:
//   k_affine_forward   thread t walks its T pairs, recomputes each denominator
//                      d = x2 - x1 (2y for a doubling, 1 for the infinity cases),
//                      stores the running prefix product, and a block-wide
//                      shuffle/shared-memory scan gives every thread the product
//                      of all OTHER threads' totals; the block total goes to HBM
//   k_affine_invert    one inversion per block total (a few thousand per round —
//                      negligible work, pure latency: binary Euclid, a warp each)
//   k_affine_backward  thread t derives 1/total_t, walks its pairs backwards
//                      peeling 1/d off the running inverse, and writes the sums
// After R rounds every slice is one affine point; k_merge_slices_affine folds the
// slices of a bucket into the XYZZ bucket array consumed by the weighted reduction.
// Measured and dropped (profiles/r2_notes.md): register-staged and L2-prefetch software pipelines of the operand
// fetches, an inlined multiply, a backward pass ordered for short live ranges, a thread-per-slice variant with the
// rounds fused, fewer pairs per thread in the late rounds, an XYZZ tail after a few affine rounds.
#pragma once
#include "msm.cuh"

namespace b200 {

constexpr int kAffBlock = 128;    // threads per block
constexpr int kAffPairs = 32;     // pairs per thread per round

template <class F>
struct AffineRound {
  const Affine<F>* table;     // precomputed points (round 1 leaves)
  const uint32_t* entries;    // sorted entry ids (round 1)
  const uint32_t* slice_start;
  const uint32_t* slice_end;
  const uint32_t* nslices_ptr;  // device: slice_off[m]
  const Affine<F>* prev;      // nodes of the previous round (r > 1): nslices * 2q
  Affine<F>* out;             // nodes of this round: nslices * q
  F* pre;                     // prefix products, one per pair
  F* others;                  // per thread: product of the other threads' totals in its block
  F* btot;                    // per block: product of all denominators (then inverted in place)
  uint32_t q_log;             // log2(pairs per slice in this round)
  uint32_t round;             // 1-based
};

// operands of pair `p`; returns false when p is beyond the live range
template <class F>
__device__ __forceinline__ bool aff_operands(const AffineRound<F>& a, uint32_t p, uint32_t npairs, Affine<F>& P, Affine<F>& Q) {
  if (p >= npairs) return false;
  uint32_t slice = p >> a.q_log, j = p & ((1u << a.q_log) - 1u);
  if (a.round == 1) {
    uint32_t s = a.slice_start[slice], e = a.slice_end[slice];
    uint32_t i0 = s + 2 * j, i1 = i0 + 1;
    P = Affine<F>::inf();
    Q = Affine<F>::inf();
    if (i0 < e) {
      uint32_t en = a.entries[i0];
      P = ld_affine_gather(&a.table[en >> 1]);
      if ((en & 1) && !P.is_inf()) P.y = P.y.neg();
    }
    if (i1 < e) {
      uint32_t en = a.entries[i1];
      Q = ld_affine_gather(&a.table[en >> 1]);
      if ((en & 1) && !Q.is_inf()) Q.y = Q.y.neg();
    }
  } else {
    size_t base = ((size_t)slice << (a.q_log + 1)) + 2 * j;
    P = ld_affine(&a.prev[base]);
    Q = ld_affine(&a.prev[base + 1]);
  }
  return true;
}

// kind of the addition and its denominator (never zero)
//   0: P or Q at infinity / P = -Q  -> d = 1 ;  1: generic, d = x2 - x1 ;  2: doubling, d = 2 y1
template <class F>
__device__ __forceinline__ int aff_denominator(const Affine<F>& P, const Affine<F>& Q, F& d) {
  if (P.is_inf() || Q.is_inf()) {
    d = F::one();
    return 0;
  }
  F dx = Q.x - P.x;
  if (!dx.is_zero()) {
    d = dx;
    return 1;
  }
  if (P.y == Q.y && !P.y.is_zero()) {
    d = P.y.dbl();
    return 2;
  }
  d = F::one();
  return 0;
}

// x-coordinate only, for the forward pass (a G2 point is a full 128-byte line: its x is half of it)
template <class P, bool I>
__device__ __forceinline__ Fp<P, I> ld_x_gather(const Affine<Fp<P, I>>* p) { Fp<P, I> r; ld256_g64(&p->x, r.l); return r; }
template <bool I>
__device__ __forceinline__ Fq2T<I> ld_x_gather(const Affine<Fq2T<I>>* p) { return ld_fe(&p->x); }

// Denominator of pair `p` for the forward pass.  The generic case needs only the two x-coordinates, so only those
// are fetched (half the registers and load instructions of the full operands -> the forward kernel runs at twice the
// occupancy, which is what hides its dependent index -> point gathers); whenever an x is zero (infinity) or the x's
// coincide (doubling / P = -Q) the full operands decide, exactly as in the backward pass.
template <class F>
__device__ __forceinline__ bool aff_forward_denominator(const AffineRound<F>& a, uint32_t p, uint32_t npairs, F& d) {
  if (p >= npairs) return false;
  uint32_t slice = p >> a.q_log, j = p & ((1u << a.q_log) - 1u);
  const Affine<F>*pp, *qp;
  if (a.round == 1) {
    uint32_t s = a.slice_start[slice], e = a.slice_end[slice];
    uint32_t i0 = s + 2 * j, i1 = i0 + 1;
    if (i1 >= e) {  // at most one live operand: nothing to invert
      d = F::one();
      return true;
    }
    pp = &a.table[a.entries[i0] >> 1];
    qp = &a.table[a.entries[i1] >> 1];
  } else {
    size_t base = ((size_t)slice << (a.q_log + 1)) + 2 * j;
    pp = &a.prev[base];
    qp = &a.prev[base + 1];
  }
  F x1 = a.round == 1 ? ld_x_gather(pp) : ld_fe(&pp->x);
  F x2 = a.round == 1 ? ld_x_gather(qp) : ld_fe(&qp->x);
  F dx = x2 - x1;
  if (x1.is_zero() || x2.is_zero() || dx.is_zero()) {
    Affine<F> P, Q;
    aff_operands(a, p, npairs, P, Q);
    aff_denominator(P, Q, d);
  } else {
    d = dx;
  }
  return true;
}

template <class F>
__device__ __forceinline__ F shfl_up_fe(const F& v, int delta) {
  F r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(F) / 4); i++) d[i] = __shfl_up_sync(0xffffffffu, s[i], delta);
  return r;
}
template <class F>
__device__ __forceinline__ F shfl_down_fe(const F& v, int delta) {
  F r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(F) / 4); i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], delta);
  return r;
}
template <class F>
__device__ __forceinline__ F shfl_idx_fe(const F& v, int lane) {
  F r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(F) / 4); i++) d[i] = __shfl_sync(0xffffffffu, s[i], lane);
  return r;
}

template <class F, int kAffT, int MINB = 1>
__global__ void __launch_bounds__(kAffBlock, MINB) k_affine_forward(AffineRound<F> a) {
  __shared__ F wtot[kAffBlock / 32];
  const uint32_t nslices = *a.nslices_ptr;
  const uint32_t npairs = nslices << a.q_log;
  const uint32_t block_base = blockIdx.x * (kAffBlock * kAffT);
  if (block_base >= npairs) {  // whole block idle (grid sized for the worst case)
    if (threadIdx.x == 0) a.btot[blockIdx.x] = F::one();
    return;
  }
  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
  F run = F::one();
  for (int k = 0; k < kAffT; k++) {
    uint32_t p = block_base + k * kAffBlock + t;  // block-interleaved: coalesced across the warp
    F d;
    if (aff_forward_denominator(a, p, npairs, d)) {
      a.pre[p] = run;
      run = run * d;
    }
  }
  // Every thread needs the product of all OTHER threads' totals of the block (no division available):
  // inclusive prefix and suffix scans over the warp's lanes, shifted by one, times the other warps' totals.
  F incl = run;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    F up = shfl_up_fe(incl, off);
    if ((int)lane >= off) incl = incl * up;
  }
  F sincl = run;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    F dn = shfl_down_fe(sincl, off);
    if ((int)lane + off < 32) sincl = sincl * dn;
  }
  F pex = shfl_up_fe(incl, 1), sex = shfl_down_fe(sincl, 1);
  if (lane == 0) pex = F::one();
  if (lane == 31) sex = F::one();
  F warp_total = shfl_idx_fe(incl, 31);
  if (lane == 0) wtot[warp] = warp_total;
  __syncthreads();
  F other_warps = F::one();
#pragma unroll
  for (int w = 0; w < kAffBlock / 32; w++)
    if (w != (int)warp) other_warps = other_warps * wtot[w];
  uint32_t gthread = blockIdx.x * kAffBlock + t;
  a.others[gthread] = pex * sex * other_warps;
  if (t == 0) a.btot[blockIdx.x] = other_warps * wtot[0];
}

// One warp per block total, lane 0 working: the binary-Euclid inversion has data-dependent control flow, and this
// launch is pure latency (a few thousand inversions between two machine-filling passes) -- 182 us with a^(p-2).
template <class F>
__global__ void __launch_bounds__(128) k_affine_invert(F* btot, uint32_t nblocks_live) {
  uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if ((threadIdx.x & 31u) || i >= nblocks_live) return;
  btot[i] = btot[i].inverse_vartime();
}

template <class F, int kAffT, int MINB = 1>
__global__ void __launch_bounds__(kAffBlock, MINB) k_affine_backward(AffineRound<F> a) {
  const uint32_t nslices = *a.nslices_ptr;
  const uint32_t npairs = nslices << a.q_log;
  const uint32_t block_base = blockIdx.x * (kAffBlock * kAffT);
  if (block_base >= npairs) return;
  const uint32_t t = threadIdx.x;
  uint32_t gthread = blockIdx.x * kAffBlock + t;
  F inv_run = a.btot[blockIdx.x] * a.others[gthread];  // 1 / (product of this thread's denominators)
  for (int k = kAffT - 1; k >= 0; k--) {
    uint32_t p = block_base + k * kAffBlock + t;
    Affine<F> P, Q;
    if (!aff_operands(a, p, npairs, P, Q)) continue;
    F d;
    int kind = aff_denominator(P, Q, d);
    F inv_d = inv_run * a.pre[p];
    inv_run = inv_run * d;
    Affine<F> Rr;
    if (kind == 1) {
      F lam = (Q.y - P.y) * inv_d;
      F x3 = lam.sqr() - P.x - Q.x;
      Rr = Affine<F>{x3, lam * (P.x - x3) - P.y};
    } else if (kind == 2) {
      F xx = P.x.sqr();
      F lam = (xx.dbl() + xx) * inv_d;
      F x3 = lam.sqr() - P.x.dbl();
      Rr = Affine<F>{x3, lam * (P.x - x3) - P.y};
    } else {
      Rr = P.is_inf() ? Q : (Q.is_inf() ? P : Affine<F>::inf());
    }
    a.out[p] = Rr;
  }
}

// buckets[b-1] = sum of the (affine) slice results of bucket b.
template <class F>
__global__ void __launch_bounds__(128)
k_merge_slices_affine(const Affine<F>* __restrict__ slice_pts, SliceTables st, uint32_t nbuckets,
                      XYZZ<F>* __restrict__ buckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x + 1;
  uint32_t lane = threadIdx.x & 31;
  uint32_t first = 0, cnt = 0;
  if (b <= nbuckets) {
    first = st.slice_off[b];
    cnt = st.slice_off[b + 1] - first;
    if (cnt <= 12) {  // the common case: a handful of slices per bucket
      XYZZ<F> acc = XYZZ<F>::inf();
      for (uint32_t k = 0; k < cnt; k++) xyzz_madd(acc, ld_affine(&slice_pts[first + k]));
      buckets[b - 1] = acc;
    }
  }
  uint32_t multi = __ballot_sync(0xffffffffu, cnt > 12);  // skewed buckets: the whole warp sums them
  while (multi) {
    int j = __ffs(multi) - 1;
    multi &= multi - 1;
    uint32_t f = __shfl_sync(0xffffffffu, first, j), c = __shfl_sync(0xffffffffu, cnt, j);
    uint32_t bj = __shfl_sync(0xffffffffu, b, j);
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t k = lane; k < c; k += 32) xyzz_madd(acc, ld_affine(&slice_pts[f + k]));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      XYZZ<F> other = shfl_down_struct(acc, off, 32);
      xyzz_add(acc, other);
    }
    if (lane == 0) buckets[bj - 1] = acc;
  }
}

}  // namespace b200
