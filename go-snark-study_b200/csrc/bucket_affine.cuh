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
// (An explicit software pipeline of the operand fetches was tried and measured slower — the extra
// live registers cost more occupancy than the prefetch hides; profiles/r1_notes.md.)
#pragma once
#include "msm.cuh"

namespace b200 {

constexpr int kAffBlock = 128;    // threads per block; pairs per thread per round = template parameter T

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

// ---- L2 prefetch of the next iteration's operands -------------------------------------------------
// A thread walks its T pairs one after another and every pair starts with a dependent chain
// slice -> entry id -> 64/128-byte point gather from a >= 1 GB table.  The warp cannot run ahead of its own
// multiply chain, so the gather latency sits in front of every iteration.  Registers are the scarce resource
// (a register-staged software pipeline was measured slower), but `prefetch.global.L2` needs none: the entry ids of
// iteration k-2 are loaded during iteration k (two registers), the points of iteration k-1 are prefetched during
// iteration k, and by the time the warp gets there its gathers are L2 hits.
struct PairIdx { uint32_t e0, e1; };
constexpr uint32_t kNoEntry = 0xffffffffu;

template <class F>
__device__ __forceinline__ PairIdx aff_pair_idx(const AffineRound<F>& a, uint32_t p, uint32_t npairs) {
  PairIdx r{kNoEntry, kNoEntry};
  if (a.round == 1 && p < npairs) {
    uint32_t slice = p >> a.q_log, j = p & ((1u << a.q_log) - 1u);
    uint32_t s = a.slice_start[slice], e = a.slice_end[slice];
    uint32_t i0 = s + 2 * j, i1 = i0 + 1;
    if (i0 < e) r.e0 = a.entries[i0];
    if (i1 < e) r.e1 = a.entries[i1];
  }
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifdef __CUDA_ARCH__
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
template <int BYTES>
__device__ __forceinline__ void prefetch_span(const void* p) {
#pragma unroll
  for (int o = 0; o < BYTES; o += 32) prefetch_l2(reinterpret_cast<const char*>(p) + o);
}
// XONLY: the forward pass reads x-coordinates only (and nothing when the pair has a single live operand)
template <class F, bool XONLY>
__device__ __forceinline__ void aff_prefetch(const AffineRound<F>& a, uint32_t p, uint32_t npairs, PairIdx ix) {
  if (p >= npairs) return;
  constexpr int kBytes = XONLY ? (int)sizeof(F) : (int)sizeof(Affine<F>);
  if (a.round == 1) {
    if (XONLY && ix.e1 == kNoEntry) return;
    if (ix.e0 != kNoEntry) prefetch_span<kBytes>(&a.table[ix.e0 >> 1]);
    if (ix.e1 != kNoEntry) prefetch_span<kBytes>(&a.table[ix.e1 >> 1]);
  } else {
    uint32_t slice = p >> a.q_log, j = p & ((1u << a.q_log) - 1u);
    size_t base = ((size_t)slice << (a.q_log + 1)) + 2 * j;
    prefetch_span<kBytes>(&a.prev[base]);
    prefetch_span<kBytes>(&a.prev[base + 1]);
  }
  if (!XONLY) prefetch_span<(int)sizeof(F)>(&a.pre[p]);
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

template <class F, int kAffT, int MINB = 1, bool PF = false>
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
  PairIdx ix_next{kNoEntry, kNoEntry};
  if (PF) ix_next = aff_pair_idx(a, block_base + kAffBlock + t, npairs);
  for (int k = 0; k < kAffT; k++) {
    uint32_t p = block_base + k * kAffBlock + t;  // block-interleaved: coalesced across the warp
    if (PF) {
      if (k + 1 < kAffT) aff_prefetch<F, true>(a, p + kAffBlock, npairs, ix_next);
      if (k + 2 < kAffT) ix_next = aff_pair_idx(a, p + 2 * kAffBlock, npairs);
    }
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

template <class F, int kAffT, int MINB = 1, bool PF = false>
__global__ void __launch_bounds__(kAffBlock, MINB) k_affine_backward(AffineRound<F> a) {
  const uint32_t nslices = *a.nslices_ptr;
  const uint32_t npairs = nslices << a.q_log;
  const uint32_t block_base = blockIdx.x * (kAffBlock * kAffT);
  if (block_base >= npairs) return;
  const uint32_t t = threadIdx.x;
  uint32_t gthread = blockIdx.x * kAffBlock + t;
  F inv_run = a.btot[blockIdx.x] * a.others[gthread];  // 1 / (product of this thread's denominators)
  PairIdx ix_next{kNoEntry, kNoEntry};
  if (PF) ix_next = aff_pair_idx(a, block_base + (kAffT - 2) * kAffBlock + t, npairs);
  for (int k = kAffT - 1; k >= 0; k--) {
    uint32_t p = block_base + k * kAffBlock + t;
    if (PF) {
      if (k >= 1) aff_prefetch<F, false>(a, p - kAffBlock, npairs, ix_next);
      if (k >= 2) ix_next = aff_pair_idx(a, p - 2 * kAffBlock, npairs);
    }
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

// ---- EXPERIMENT (off by default, B200_AFF_SP): software-pipelined variants with the multiply INLINED ------------
// Hypothesis to measure in the next round: an out-of-line multiply is a scoreboard barrier (loads cannot stay in flight
// across a CALL), so in the kernels above every iteration pays its index -> point gather latency in front of its
// multiplies, and only other warps can hide it.  Here the field type is the inlined-multiply one (FqH, same memory
// layout) and the operand fetches are staged explicitly three deep: slice bounds of pair k+3, entry ids of pair k+2 and
// the point gathers of pair k+1 are issued before the multiplies of pair k.
template <class F>
struct AffStage {
  const AffineRound<F>& a;
  uint32_t npairs;
  __device__ __forceinline__ uint2 bounds(uint32_t p) const {
    if (a.round != 1 || p >= npairs) return make_uint2(0u, 0u);
    uint32_t slice = p >> a.q_log;
    return make_uint2(a.slice_start[slice], a.slice_end[slice]);
  }
  __device__ __forceinline__ PairIdx entries(uint32_t p, uint2 sb) const {
    PairIdx r{kNoEntry, kNoEntry};
    if (a.round == 1 && p < npairs) {
      uint32_t j = p & ((1u << a.q_log) - 1u);
      uint32_t i0 = sb.x + 2 * j;
      if (i0 < sb.y) r.e0 = a.entries[i0];
      if (i0 + 1 < sb.y) r.e1 = a.entries[i0 + 1];
    }
    return r;
  }
  __device__ __forceinline__ size_t node_base(uint32_t p) const {
    uint32_t slice = p >> a.q_log, j = p & ((1u << a.q_log) - 1u);
    return ((size_t)slice << (a.q_log + 1)) + 2 * j;
  }
};

template <class F, int kAffT, int MINB>
__global__ void __launch_bounds__(kAffBlock, MINB) k_affine_forward_sp(AffineRound<F> a) {
  __shared__ F wtot[kAffBlock / 32];
  const uint32_t nslices = *a.nslices_ptr;
  const uint32_t npairs = nslices << a.q_log;
  const uint32_t block_base = blockIdx.x * (kAffBlock * kAffT);
  if (block_base >= npairs) {
    if (threadIdx.x == 0) a.btot[blockIdx.x] = F::one();
    return;
  }
  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const AffStage<F> st{a, npairs};
  // gather the two x-coordinates of pair p: 0 = beyond the range, 1 = single live operand (d = 1), 2 = fetched
  auto gather = [&](uint32_t p, PairIdx ix, F& x1, F& x2) -> uint32_t {
    if (p >= npairs) return 0u;
    if (a.round == 1) {
      if (ix.e1 == kNoEntry) return 1u;
      x1 = ld_x_gather(&a.table[ix.e0 >> 1]);
      x2 = ld_x_gather(&a.table[ix.e1 >> 1]);
      return 2u;
    }
    size_t base = st.node_base(p);
    x1 = ld_fe(&a.prev[base].x);
    x2 = ld_fe(&a.prev[base + 1].x);
    return 2u;
  };
  const uint32_t p0 = block_base + t;
  uint2 sb = st.bounds(p0 + 2 * kAffBlock);                                   // bounds of pair k+2
  PairIdx ix = st.entries(p0 + kAffBlock, st.bounds(p0 + kAffBlock));           // entries of pair k+1
  F cx1 = F::zero(), cx2 = F::zero();
  uint32_t cst = gather(p0, st.entries(p0, st.bounds(p0)), cx1, cx2);
  F run = F::one();
#pragma unroll 1
  for (int k = 0; k < kAffT; k++) {
    const uint32_t p = p0 + k * kAffBlock;
    F nx1 = F::zero(), nx2 = F::zero();
    uint32_t nst = 0;
    if (k + 1 < kAffT) nst = gather(p + kAffBlock, ix, nx1, nx2);
    if (k + 2 < kAffT) ix = st.entries(p + 2 * kAffBlock, sb);
    if (k + 3 < kAffT) sb = st.bounds(p + 3 * kAffBlock);
    if (cst) {
      F d = F::one();
      if (cst == 2) {
        F dx = cx2 - cx1;
        if (cx1.is_zero() || cx2.is_zero() || dx.is_zero()) {  // infinity / doubling / P = -Q: the full operands decide
          Affine<F> P, Q;
          aff_operands(a, p, npairs, P, Q);
          aff_denominator(P, Q, d);
        } else {
          d = dx;
        }
      }
      a.pre[p] = run;
      run = run * d;
    }
    cx1 = nx1;
    cx2 = nx2;
    cst = nst;
  }
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
  a.others[blockIdx.x * kAffBlock + t] = pex * sex * other_warps;
  if (t == 0) a.btot[blockIdx.x] = other_warps * wtot[0];
}

template <class F, int kAffT, int MINB>
__global__ void __launch_bounds__(kAffBlock, MINB) k_affine_backward_sp(AffineRound<F> a) {
  const uint32_t nslices = *a.nslices_ptr;
  const uint32_t npairs = nslices << a.q_log;
  const uint32_t block_base = blockIdx.x * (kAffBlock * kAffT);
  if (block_base >= npairs) return;
  const uint32_t t = threadIdx.x;
  const AffStage<F> st{a, npairs};
  // fetch both operands and the prefix product of pair p (signs of round-1 entries are applied at use)
  auto gather = [&](uint32_t p, PairIdx ix, Affine<F>& P, Affine<F>& Q, F& pre) -> uint32_t {
    if (p >= npairs) return 0u;
    P = Affine<F>::inf();
    Q = Affine<F>::inf();
    if (a.round == 1) {
      if (ix.e0 != kNoEntry) P = ld_affine_gather(&a.table[ix.e0 >> 1]);
      if (ix.e1 != kNoEntry) Q = ld_affine_gather(&a.table[ix.e1 >> 1]);
    } else {
      size_t base = st.node_base(p);
      P = ld_affine(&a.prev[base]);
      Q = ld_affine(&a.prev[base + 1]);
    }
    pre = ld_fe(&a.pre[p]);
    return 1u;
  };
  const uint32_t pl = block_base + (kAffT - 1) * kAffBlock + t;   // last pair of this thread: processed first
  F inv_run = a.btot[blockIdx.x] * a.others[blockIdx.x * kAffBlock + t];
  uint2 sb = st.bounds(pl - 2 * kAffBlock);
  PairIdx ix = st.entries(pl - kAffBlock, st.bounds(pl - kAffBlock));
  PairIdx cix = st.entries(pl, st.bounds(pl));
  Affine<F> cP, cQ;
  F cpre = F::zero();
  uint32_t cst = gather(pl, cix, cP, cQ, cpre);
#pragma unroll 1
  for (int k = kAffT - 1; k >= 0; k--) {
    const uint32_t p = block_base + k * kAffBlock + t;
    Affine<F> nP, nQ;
    F npre = F::zero();
    uint32_t nst = 0;
    const PairIdx nix = ix;
    if (k >= 1) nst = gather(p - kAffBlock, ix, nP, nQ, npre);
    if (k >= 2) ix = st.entries(p - 2 * kAffBlock, sb);
    if (k >= 3) sb = st.bounds(p - 3 * kAffBlock);
    if (cst) {
      Affine<F> P = cP, Q = cQ;
      if (a.round == 1) {
        if (cix.e0 != kNoEntry && (cix.e0 & 1u) && !P.is_inf()) P.y = P.y.neg();
        if (cix.e1 != kNoEntry && (cix.e1 & 1u) && !Q.is_inf()) Q.y = Q.y.neg();
      }
      F d;
      int kind = aff_denominator(P, Q, d);
      F inv_d = inv_run * cpre;
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
    cP = nP;
    cQ = nQ;
    cpre = npre;
    cst = nst;
    cix = nix;
  }
}

// ---- EXPERIMENT (off by default, B200_AFF_LR): backward pass ordered for short live ranges -------------------------
// The backward kernel above fetches both operands (4 field elements) before its first multiply, which for G2 is 64
// registers of operands alone and makes the 128-register build spill (336 B).  Here the generic case is ordered so
// that at most ~5 field elements are live: prefix -> 1/d, then the x's -> running inverse, then the y's -> lambda, x3,
// y3.  Pairs with an absent operand, a zero x (infinity) or equal x's (doubling / P = -Q) take the out-of-line slow path
// with the original logic.
template <class F>
__device__ __noinline__ void aff_backward_slow(const AffineRound<F>& a, uint32_t p, uint32_t npairs, F& inv_run) {
  Affine<F> P, Q;
  if (!aff_operands(a, p, npairs, P, Q)) return;
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

template <class F, int kAffT, int MINB>
__global__ void __launch_bounds__(kAffBlock, MINB) k_affine_backward_lr(AffineRound<F> a) {
  const uint32_t nslices = *a.nslices_ptr;
  const uint32_t npairs = nslices << a.q_log;
  const uint32_t block_base = blockIdx.x * (kAffBlock * kAffT);
  if (block_base >= npairs) return;
  const uint32_t t = threadIdx.x;
  F inv_run = a.btot[blockIdx.x] * a.others[blockIdx.x * kAffBlock + t];
#pragma unroll 1
  for (int k = kAffT - 1; k >= 0; k--) {
    const uint32_t p = block_base + k * kAffBlock + t;
    if (p >= npairs) continue;
    const uint32_t slice = p >> a.q_log, j = p & ((1u << a.q_log) - 1u);
    const Affine<F>*pp = nullptr, *qp = nullptr;
    uint32_t neg1 = 0, neg2 = 0;
    if (a.round == 1) {
      uint32_t s = a.slice_start[slice], e = a.slice_end[slice];
      uint32_t i0 = s + 2 * j;
      if (i0 + 1 < e) {
        uint32_t e0 = a.entries[i0], e1 = a.entries[i0 + 1];
        pp = &a.table[e0 >> 1];
        qp = &a.table[e1 >> 1];
        neg1 = e0 & 1u;
        neg2 = e1 & 1u;
      }
    } else {
      size_t base = ((size_t)slice << (a.q_log + 1)) + 2 * j;
      pp = &a.prev[base];
      qp = &a.prev[base + 1];
    }
    bool fast = pp != nullptr;
    F x1, x2, dx;
    if (fast) {
      x1 = a.round == 1 ? ld_x_gather(pp) : ld_fe(&pp->x);
      x2 = a.round == 1 ? ld_x_gather(qp) : ld_fe(&qp->x);
      dx = x2 - x1;
      fast = !(x1.is_zero() || x2.is_zero() || dx.is_zero());
    }
    if (!fast) {
      aff_backward_slow(a, p, npairs, inv_run);
      continue;
    }
    F lam;
    {
      F inv_d = inv_run * ld_fe(&a.pre[p]);
      inv_run = inv_run * dx;
      F y2 = ld_fe(&qp->y);
      if (neg2) y2 = y2.neg();
      F y1 = ld_fe(&pp->y);
      if (neg1) y1 = y1.neg();
      lam = (y2 - y1) * inv_d;
    }
    F x3 = lam.sqr() - x1 - x2;
    F y1 = ld_fe(&pp->y);                     // re-read (L1 hit) instead of keeping it live across two multiplies
    if (neg1) y1 = y1.neg();
    a.out[p] = Affine<F>{x3, lam * (x1 - x3) - y1};
  }
}

// ---- EXPERIMENT (off by default, B200_AFF_TS): one thread per slice, rounds fused ----------------------------------
// In the layout above a round's pairs are dealt out block-interleaved, so the forward pass of round r+1 has to re-read the
// nodes round r just wrote, and the late rounds have too few pairs to fill the machine.  Here thread s owns slice s — the
// whole subtree — in every round: the kernel of round r walks the slice's 2^(R-r) pairs of that level, does the backward
// step (the addition) and, as soon as two sibling nodes exist, multiplies the denominator of THEIR addition into the
// prefix product of round r+1.  Only round 1 needs a stand-alone forward pass; every later round is one kernel + the
// inversion of the block totals, every thread is busy in every round (work per thread halves, the grid stays nslices / 128
// CTAs), and the forward pass's operands are registers instead of gathers.  The walk direction alternates per round
// (prefix products must be peeled off in the reverse order of their accumulation).  Node storage (slice-major, ping-pong)
// is the same as above, so k_merge_slices_affine is unchanged.
template <class F>
struct AffineRoundTS {
  const Affine<F>* table;
  const uint32_t* entries;
  const uint32_t* slice_start;
  const uint32_t* slice_end;
  const uint32_t* nslices_ptr;
  const Affine<F>* prev;   // nodes of the previous round (round > 1)
  Affine<F>* out;          // nodes of this round
  const F* pre;            // prefix products of this round's pairs (slice-major: (s << q_log) + j)
  const F* others;         // per thread: product of the other threads' totals of its CTA
  const F* btot;           // per CTA: inverse of the product of all its denominators
  F* pre_next;             // the same three for round + 1 (written unless `last`)
  F* others_next;
  F* btot_next;
  uint32_t q_log;          // log2(pairs per slice in this round)
  uint32_t round;          // 1-based
  uint32_t last;           // no round + 1
};

// others[thread] = product of the OTHER threads' `run` in the CTA, btot[CTA] = product of all (as in k_affine_forward)
template <class F>
__device__ __forceinline__ void aff_block_scan(const F& run, F* others, F* btot, F* wtot) {
  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
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
  others[blockIdx.x * kAffBlock + t] = pex * sex * other_warps;
  if (t == 0) btot[blockIdx.x] = other_warps * wtot[0];
}

// operands of pair j of slice s in this round
template <class F>
__device__ __forceinline__ void aff_ts_operands(const AffineRoundTS<F>& a, uint32_t s, uint32_t j, uint32_t s0, uint32_t s1,
                                                Affine<F>& P, Affine<F>& Q) {
  if (a.round == 1) {
    uint32_t i0 = s0 + 2 * j, i1 = i0 + 1;
    P = Affine<F>::inf();
    Q = Affine<F>::inf();
    if (i0 < s1) {
      uint32_t en = a.entries[i0];
      P = ld_affine_gather(&a.table[en >> 1]);
      if ((en & 1) && !P.is_inf()) P.y = P.y.neg();
    }
    if (i1 < s1) {
      uint32_t en = a.entries[i1];
      Q = ld_affine_gather(&a.table[en >> 1]);
      if ((en & 1) && !Q.is_inf()) Q.y = Q.y.neg();
    }
  } else {
    size_t base = ((size_t)s << (a.q_log + 1)) + 2 * j;
    P = ld_affine(&a.prev[base]);
    Q = ld_affine(&a.prev[base + 1]);
  }
}

// Round 1 only: prefix products of the leaf-pair denominators of slice s (ascending j), then the block scan.
template <class F, int MINB = 1>
__global__ void __launch_bounds__(kAffBlock, MINB) k_affine_ts_forward1(AffineRoundTS<F> a) {
  __shared__ F wtot[kAffBlock / 32];
  const uint32_t nslices = *a.nslices_ptr;
  const uint32_t s = blockIdx.x * kAffBlock + threadIdx.x;
  F run = F::one();
  if (s < nslices) {
    const uint32_t q = 1u << a.q_log, s0 = a.slice_start[s], s1 = a.slice_end[s];
    for (uint32_t j = 0; j < q; j++) {
      Affine<F> P, Q;
      aff_ts_operands(a, s, j, s0, s1, P, Q);
      F d;
      aff_denominator(P, Q, d);
      a.pre_next[((size_t)s << a.q_log) + j] = run;
      run = run * d;
    }
  }
  aff_block_scan(run, a.others_next, a.btot_next, wtot);
}

// one pair with the full logic (absent / infinity operands, doubling, P = -Q), out of line: the rare cases
template <class F>
__device__ __noinline__ void aff_ts_pair_slow(const AffineRoundTS<F>& a, uint32_t s, uint32_t j, uint32_t s0, uint32_t s1,
                                              F& inv_run, Affine<F>& Rr) {
  Affine<F> P, Q;
  aff_ts_operands(a, s, j, s0, s1, P, Q);
  F d;
  int kind = aff_denominator(P, Q, d);
  F inv_d = inv_run * a.pre[((size_t)s << a.q_log) + j];
  inv_run = inv_run * d;
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
}
// denominator of the addition of two sibling nodes this thread has just written (rare cases: re-read them coherently)
template <class F>
__device__ __forceinline__ void aff_ts_sibling_slow(const Affine<F>* out, size_t p_even, F& d2) {
  Affine<F> L = out[p_even], Rt = out[p_even + 1];
  aff_denominator(L, Rt, d2);
}

// SMEM: keep the two values that live across iterations (the next round's running product and the held sibling x) in
// shared memory instead of registers — for G2 they are 32 registers of a 128-register budget.
template <class F, int MINB = 1, bool SMEM = false>
__global__ void __launch_bounds__(kAffBlock, MINB) k_affine_ts_round(AffineRoundTS<F> a) {
  __shared__ F wtot[kAffBlock / 32];
  __shared__ F sm_state[SMEM ? 2 * kAffBlock : 1];
  const uint32_t nslices = *a.nslices_ptr;
  const uint32_t s = blockIdx.x * kAffBlock + threadIdx.x;
  F next_run = F::one();
  if (SMEM) sm_state[threadIdx.x] = next_run;
  if (s < nslices) {
    const uint32_t q = 1u << a.q_log;
    const bool desc = (a.round & 1u) != 0;        // round 1's stand-alone forward pass accumulated ascending
    uint32_t s0 = 0, s1 = 0;
    if (a.round == 1) {
      s0 = a.slice_start[s];
      s1 = a.slice_end[s];
    }
    F inv_run = a.btot[blockIdx.x] * a.others[blockIdx.x * kAffBlock + threadIdx.x];
    F held_x = F::zero();                         // x of the sibling computed one step earlier (zero: infinity)
    if (SMEM) sm_state[kAffBlock + threadIdx.x] = held_x;
#pragma unroll 1
    for (uint32_t i = 0; i < q; i++) {
      const uint32_t j = desc ? q - 1 - i : i;
      const size_t p = ((size_t)s << a.q_log) + j;
      // locate the operands; the generic case (both present, x's non-zero and distinct) runs with short live ranges
      const Affine<F>*pp = nullptr, *qp = nullptr;
      uint32_t neg1 = 0, neg2 = 0;
      if (a.round == 1) {
        uint32_t i0 = s0 + 2 * j;
        if (i0 + 1 < s1) {
          uint32_t e0 = a.entries[i0], e1 = a.entries[i0 + 1];
          pp = &a.table[e0 >> 1];
          qp = &a.table[e1 >> 1];
          neg1 = e0 & 1u;
          neg2 = e1 & 1u;
        }
      } else {
        size_t base = ((size_t)s << (a.q_log + 1)) + 2 * j;
        pp = &a.prev[base];
        qp = &a.prev[base + 1];
      }
      bool fast = pp != nullptr;
      F x1, x2, dx;
      if (fast) {
        x1 = a.round == 1 ? ld_x_gather(pp) : ld_fe(&pp->x);
        x2 = a.round == 1 ? ld_x_gather(qp) : ld_fe(&qp->x);
        dx = x2 - x1;
        fast = !(x1.is_zero() || x2.is_zero() || dx.is_zero());
      }
      F rx;                                        // x of the node just produced (zero when it is the point at infinity)
      if (fast) {
        F lam;
        {
          F inv_d = inv_run * ld_fe(&a.pre[p]);
          inv_run = inv_run * dx;
          F y2 = ld_fe(&qp->y);
          if (neg2) y2 = y2.neg();
          F y1 = ld_fe(&pp->y);
          if (neg1) y1 = y1.neg();
          lam = (y2 - y1) * inv_d;
        }
        rx = lam.sqr() - x1 - x2;
        F y1 = ld_fe(&pp->y);
        if (neg1) y1 = y1.neg();
        a.out[p] = Affine<F>{rx, lam * (x1 - rx) - y1};
      } else {
        Affine<F> Rr;
        aff_ts_pair_slow(a, s, j, s0, s1, inv_run, Rr);
        a.out[p] = Rr;
        rx = Rr.x;                                 // a zero x sends the sibling step to its exact slow path
      }
      if (!a.last) {
        const bool second = desc ? (j & 1u) == 0 : (j & 1u) == 1;   // both siblings of pair j >> 1 now exist
        if (second) {
          if (SMEM) held_x = sm_state[kAffBlock + threadIdx.x];
          F xl = desc ? rx : held_x, xr = desc ? held_x : rx;
          F d2 = xr - xl;
          if (xl.is_zero() || xr.is_zero() || d2.is_zero()) aff_ts_sibling_slow(a.out, p & ~(size_t)1, d2);
          if (SMEM) next_run = sm_state[threadIdx.x];
          a.pre_next[((size_t)s << (a.q_log - 1)) + (j >> 1)] = next_run;
          next_run = next_run * d2;
          if (SMEM) sm_state[threadIdx.x] = next_run;
        } else {
          if (SMEM) sm_state[kAffBlock + threadIdx.x] = rx;
          else held_x = rx;
        }
      }
    }
  }
  if (SMEM) next_run = sm_state[threadIdx.x];
  if (!a.last) aff_block_scan(next_run, a.others_next, a.btot_next, wtot);
}

// Tail of the tree (tuning knob B200_AFF_ROUNDS): after fewer than log2(S) affine rounds every slice
// still holds `q` nodes; LPB lanes per bucket add the (contiguous) nodes of all its slices with XYZZ mixed
// adds and merge through a shuffle tree.
template <class F, int LPB>
__global__ void __launch_bounds__(128)
k_accumulate_nodes(const Affine<F>* __restrict__ nodes, SliceTables st, uint32_t nbuckets, uint32_t q_log,
                   XYZZ<F>* __restrict__ buckets) {
  uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t b = gt / LPB + 1;
  uint32_t lane = gt % LPB;
  bool live = b <= nbuckets;
  size_t start = 0, end = 0;
  if (live) {
    start = (size_t)st.slice_off[b] << q_log;
    end = (size_t)st.slice_off[b + 1] << q_log;
  }
  XYZZ<F> acc = XYZZ<F>::inf();
  for (size_t k = start + lane; k < end; k += LPB) xyzz_madd(acc, ld_affine(&nodes[k]));
#pragma unroll
  for (int off = LPB / 2; off > 0; off >>= 1) {
    XYZZ<F> other = shfl_down_struct(acc, off, LPB);
    xyzz_add(acc, other);
  }
  if (live && lane == 0) buckets[b - 1] = acc;
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
