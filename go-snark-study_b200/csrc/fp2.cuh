// F_q^2 = F_q[u]/(u^2 + 1) — the coordinate field of G2.
// Replaces fields/fq2.go:37-133 (Add, Sub, Neg, Mul (Karatsuba), Square) with
// the non-residue fixed to -1 (bn128/bn128.go:86).  Both components are kept
// canonical, so results equal the reference's [2]*big.Int pair exactly.
//
// INL = false (default): Mul and Square are one out-of-line device function
// each, whose body holds the 3 (resp. 2) inlined F_q Montgomery products — the
// call overhead is amortised over ~600 instructions and G2 kernels stay small.
#pragma once
#include "fp.cuh"

namespace b200 {

template <bool INL>
struct Fq2T;
#ifdef __CUDACC__
template <bool INL>
__device__ __noinline__ Fq2T<INL> fq2_mul_outlined(Fq2T<INL> a, Fq2T<INL> b);
template <bool INL>
__device__ __noinline__ Fq2T<INL> fq2_sqr_outlined(Fq2T<INL> a);
#endif

template <bool INL = false>
struct alignas(32) Fq2T {
  using B = FqH;  // component type: inlined F_q products inside the F_q^2 routines
  B c0, c1;

  static HD Fq2T zero() { return Fq2T{B::zero(), B::zero()}; }
  static HD Fq2T one() { return Fq2T{B::one(), B::zero()}; }
  HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  HD bool operator==(const Fq2T& b) const { return c0 == b.c0 && c1 == b.c1; }
  HD bool operator!=(const Fq2T& b) const { return !(*this == b); }

  friend HD Fq2T operator+(const Fq2T& a, const Fq2T& b) { return Fq2T{a.c0 + b.c0, a.c1 + b.c1}; }
  friend HD Fq2T operator-(const Fq2T& a, const Fq2T& b) { return Fq2T{a.c0 - b.c0, a.c1 - b.c1}; }
  HD Fq2T neg() const { return Fq2T{c0.neg(), c1.neg()}; }
  HD Fq2T dbl() const { return Fq2T{c0.dbl(), c1.dbl()}; }

  // (a0 + a1 u)(b0 + b1 u) = (a0 b0 - a1 b1) + ((a0 + a1)(b0 + b1) - a0 b0 - a1 b1) u
  // Lazy reduction: the three Karatsuba products are kept as 512-bit integers and each output
  // component is reduced ONCE (3*64 + 2*64 = 320 wide multiplies instead of 3*128).  Bounds, with
  // p < 2^254: v0, v1 < p^2;  s = (a0+a1)(b0+b1) < 4p^2 < 2^510;  c1 = s - v0 - v1 = a0b1 + a1b0 in
  // [0, 2p^2);  c0 = v0 - v1 + p^2 in (0, 2p^2);  2p^2 < p*2^256, the Montgomery-reduction bound.
  static HD Fq2T mul_impl(const Fq2T& a, const Fq2T& b) {
    using namespace cc;
    uint32_t v0[16], v1[16], s[16], sa[8], sb[8];
    B::mul_full(v0, a.c0.l, b.c0.l);
    B::mul_full(v1, a.c1.l, b.c1.l);
    sa[0] = add_cc(a.c0.l[0], a.c1.l[0]);
    sb[0] = 0;
#pragma unroll
    for (int i = 1; i < 7; i++) sa[i] = addc_cc(a.c0.l[i], a.c1.l[i]);
    sa[7] = addc(a.c0.l[7], a.c1.l[7]);
    sb[0] = add_cc(b.c0.l[0], b.c1.l[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) sb[i] = addc_cc(b.c0.l[i], b.c1.l[i]);
    sb[7] = addc(b.c0.l[7], b.c1.l[7]);
    B::mul_full(s, sa, sb);
    // s <- s - v0 - v1
    s[0] = sub_cc(s[0], v0[0]);
#pragma unroll
    for (int i = 1; i < 15; i++) s[i] = subc_cc(s[i], v0[i]);
    s[15] = subc(s[15], v0[15]);
    s[0] = sub_cc(s[0], v1[0]);
#pragma unroll
    for (int i = 1; i < 15; i++) s[i] = subc_cc(s[i], v1[i]);
    s[15] = subc(s[15], v1[15]);
    // v0 <- v0 - v1 + p^2
    v0[0] = sub_cc(v0[0], v1[0]);
#pragma unroll
    for (int i = 1; i < 15; i++) v0[i] = subc_cc(v0[i], v1[i]);
    v0[15] = subc(v0[15], v1[15]);
    v0[0] = add_cc(v0[0], FqP2(0));
#pragma unroll
    for (int i = 1; i < 15; i++) v0[i] = addc_cc(v0[i], FqP2(i));
    v0[15] = addc(v0[15], FqP2(15));
    return Fq2T{B::redc_wide(v0), B::redc_wide(s)};
  }
  // (a0 + a1 u)^2 = (a0 + a1)(a0 - a1) + 2 a0 a1 u
  static HD Fq2T sqr_impl(const Fq2T& a) {
    B m = a.c0 * a.c1;
    return Fq2T{(a.c0 + a.c1) * (a.c0 - a.c1), m.dbl()};
  }
  friend HD Fq2T operator*(const Fq2T& a, const Fq2T& b) {
#ifdef __CUDA_ARCH__
    if constexpr (!INL) return fq2_mul_outlined<INL>(a, b);
#endif
    return mul_impl(a, b);
  }
  HD Fq2T sqr() const {
#ifdef __CUDA_ARCH__
    if constexpr (!INL) return fq2_sqr_outlined<INL>(*this);
#endif
    return sqr_impl(*this);
  }
  HD Fq2T to_mont() const { return Fq2T{c0.to_mont(), c1.to_mont()}; }
  HD Fq2T from_mont() const { return Fq2T{c0.from_mont(), c1.from_mont()}; }
  // fields/fq2.go:99-108
  HD Fq2T inverse() const {
    B t = (c0.sqr() + c1.sqr()).inverse();
    return Fq2T{c0 * t, (c1 * t).neg()};
  }
  HD Fq2T inverse_vartime() const {  // same value, Fp::inverse_vartime for the norm
    B t = (c0.sqr() + c1.sqr()).inverse_vartime();
    return Fq2T{c0 * t, (c1 * t).neg()};
  }
  HD bool geq_modulus() const { return c0.geq_modulus() || c1.geq_modulus(); }
};

#ifdef __CUDACC__
template <bool INL>
__device__ __noinline__ Fq2T<INL> fq2_mul_outlined(Fq2T<INL> a, Fq2T<INL> b) {
  return Fq2T<INL>::mul_impl(a, b);
}
template <bool INL>
__device__ __noinline__ Fq2T<INL> fq2_sqr_outlined(Fq2T<INL> a) {
  return Fq2T<INL>::sqr_impl(a);
}
#endif

using Fq2 = Fq2T<false>;
using Fq2H = Fq2T<true>;

}  // namespace b200
