// 254-bit prime-field arithmetic in Montgomery form, 8 x 32-bit limbs.
//
// Replaces the reference's big.Int field ops (fields/fq.go:32-98: Add, Sub,
// Neg, Mul, Square, Inverse — each a big.Int op followed by big.Int.Mod) for
// the two BN254 moduli of bn128/bn128.go:40,46.  Values are kept fully reduced
// in [0, p) so equality and zero tests are limb compares, and converting out
// of Montgomery form yields exactly the canonical residue the reference holds.
//
// Multiplication is operand-scanning Montgomery (CIOS) with the partial
// products of even- and odd-indexed limbs accumulated in two separate carry
// chains (each 32x32->64 product then occupies a private column pair, so a
// chain is mad.lo.cc / madc.hi.cc alternating with no carry conflicts; ptxas
// fuses each pair into one IMAD.WIDE.U32 with carry-in/out).
#pragma once
#include "constants.cuh"
#include "hd.cuh"

namespace b200 {

// INL = true : multiplication is force-inlined into the caller (hot kernels).
// INL = false: multiplication is one out-of-line device function per field
//              (register-passed, no stack traffic); keeps cold kernels and the
//              F_q^2 tower small and the build fast.  Same storage either way.
template <class P, bool INL>
struct Fp;
#ifdef __CUDACC__
template <class P>
__device__ __noinline__ Fp<P, false> fp_mul_outlined(Fp<P, false> a, Fp<P, false> b);
template <class P, bool INL>
__device__ __noinline__ Fp<P, INL> fp_inverse_outlined(Fp<P, INL> a);
template <class P, bool INL>
__device__ __noinline__ Fp<P, INL> fp_inverse_vartime_outlined(Fp<P, INL> a);
#endif

template <class P, bool INL = false>
struct alignas(32) Fp {
  uint32_t l[8];

  static HD Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = 0;
    return r;
  }
  static HD Fp one() {  // Montgomery form of 1
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = P::ONE(i);
    return r;
  }
  static HD Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = P::R2(i);
    return r;
  }

  HD bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= l[i];
    return o == 0;
  }
  HD bool operator==(const Fp& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= l[i] ^ b.l[i];
    return o == 0;
  }
  HD bool operator!=(const Fp& b) const { return !(*this == b); }

  // r = a - p if a >= p else a     (a < 2p)
  static HD void final_sub(uint32_t* a) {
    uint32_t t[8];
    t[0] = cc::sub_cc(a[0], P::MOD(0));
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = cc::subc_cc(a[i], P::MOD(i));
    uint32_t borrow = cc::subc(0u, 0u);  // 0 or 0xffffffff
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = borrow ? a[i] : t[i];
  }

  friend HD Fp operator+(const Fp& a, const Fp& b) {
    Fp r;
    r.l[0] = cc::add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.l[i] = cc::addc_cc(a.l[i], b.l[i]);
    r.l[7] = cc::addc(a.l[7], b.l[7]);  // < 2^255: no carry out
    final_sub(r.l);
    return r;
  }
  friend HD Fp operator-(const Fp& a, const Fp& b) {
    Fp r;
    r.l[0] = cc::sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.l[i] = cc::subc_cc(a.l[i], b.l[i]);
    uint32_t borrow = cc::subc(0u, 0u);  // all-ones if a < b
    r.l[0] = cc::add_cc(r.l[0], P::MOD(0) & borrow);
#pragma unroll
    for (int i = 1; i < 7; i++) r.l[i] = cc::addc_cc(r.l[i], P::MOD(i) & borrow);
    r.l[7] = cc::addc(r.l[7], P::MOD(7) & borrow);
    return r;
  }
  HD Fp neg() const { return zero() - *this; }
  HD Fp dbl() const { return *this + *this; }

  // ---- Montgomery multiplication ------------------------------------------
  // One operand-scanning step: (E, O) += a * bi, then += mi * p so that the
  // lowest column becomes zero.  E holds columns j (32-bit columns), O holds
  // columns j+1; on entry (non-first) the roles were swapped by the previous
  // step, whose zeroed low column is dropped here (the ">> 32" of Montgomery).
  static HD void mad_n_redc(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi, bool first) {
    using namespace cc;
    if (first) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) mul_wide(a[j + 1], bi, O[j], O[j + 1]);
#pragma unroll
      for (int j = 0; j < 8; j += 2) mul_wide(a[j], bi, E[j], E[j + 1]);
    } else {
      E[0] = add_cc(E[0], O[1]);
#pragma unroll
      for (int j = 0; j < 6; j += 2) {
        O[j] = madc_lo_cc(a[j + 1], bi, O[j + 2]);
        O[j + 1] = madc_hi_cc(a[j + 1], bi, O[j + 3]);
      }
      O[6] = madc_lo_cc(a[7], bi, 0u);
      O[7] = madc_hi_cc(a[7], bi, 0u);  // (.cc so the pair fuses; the carry out is always 0)
      E[0] = mad_lo_cc(a[0], bi, E[0]);
      E[1] = madc_hi_cc(a[0], bi, E[1]);
#pragma unroll
      for (int j = 2; j < 8; j += 2) {
        E[j] = madc_lo_cc(a[j], bi, E[j]);
        E[j + 1] = madc_hi_cc(a[j], bi, E[j + 1]);
      }
      O[7] = addc(O[7], 0u);
    }
    uint32_t mi = mul_lo(E[0], P::INV);
    O[0] = mad_lo_cc(P::MOD(1), mi, O[0]);
    O[1] = madc_hi_cc(P::MOD(1), mi, O[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
      O[j] = madc_lo_cc(P::MOD(j + 1), mi, O[j]);
      O[j + 1] = madc_hi_cc(P::MOD(j + 1), mi, O[j + 1]);
    }
    E[0] = mad_lo_cc(P::MOD(0), mi, E[0]);
    E[1] = madc_hi_cc(P::MOD(0), mi, E[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
      E[j] = madc_lo_cc(P::MOD(j), mi, E[j]);
      E[j + 1] = madc_hi_cc(P::MOD(j), mi, E[j + 1]);
    }
    O[7] = addc(O[7], 0u);
  }

  friend HD Fp operator*(const Fp& a, const Fp& b) {
#ifdef __CUDA_ARCH__
    if constexpr (!INL) return fp_mul_outlined<P>(a, b);
#endif
    return mul_impl(a, b);
  }
  static HD Fp mul_impl(const Fp& a, const Fp& b) {
    using namespace cc;
    uint32_t even[8], odd[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      mad_n_redc(even, odd, a.l, b.l[i], i == 0);
      mad_n_redc(odd, even, a.l, b.l[i + 1], false);
    }
    Fp r;
    r.l[0] = add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.l[i] = addc_cc(even[i], odd[i + 1]);
    r.l[7] = addc(even[7], 0u);
    final_sub(r.l);
    return r;
  }
  HD Fp sqr() const { return (*this) * (*this); }

  // ---- split multiply for lazy reduction (used by the F_q^2 tower) ---------------------------
  // w[0..16) = a * b as a plain 512-bit integer (operands may be unreduced, < 2^256).
  // Row i adds a * b[i] at column i; the partial products of even and odd columns are kept in two
  // accumulators E (pairs aligned at even columns) and O (aligned at odd columns) so every row is two
  // independent carry chains of fused IMAD.WIDE.U32.X, exactly as in mul_impl.
  static HD void mul_full(uint32_t* w, const uint32_t* a, const uint32_t* b) {
    using namespace cc;
    uint32_t E[16], O[16];  // O[k] sits at column k + 1
#pragma unroll
    for (int k = 0; k < 16; k++) E[k] = O[k] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint32_t bi = b[i];
      if ((i & 1) == 0) {
        // a[even] * bi -> columns i + j (even)  : E ;  a[odd] * bi -> columns i + j (odd) : O at index i + j - 1
        E[i] = mad_lo_cc(a[0], bi, E[i]);
        E[i + 1] = madc_hi_cc(a[0], bi, E[i + 1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
          E[i + j] = madc_lo_cc(a[j], bi, E[i + j]);
          E[i + j + 1] = madc_hi_cc(a[j], bi, E[i + j + 1]);
        }
        if (i + 8 < 16) E[i + 8] = addc(E[i + 8], 0u);
        O[i] = mad_lo_cc(a[1], bi, O[i]);
        O[i + 1] = madc_hi_cc(a[1], bi, O[i + 1]);
#pragma unroll
        for (int j = 3; j < 8; j += 2) {
          O[i + j - 1] = madc_lo_cc(a[j], bi, O[i + j - 1]);
          O[i + j] = madc_hi_cc(a[j], bi, O[i + j]);
        }
        if (i + 8 < 16) O[i + 8] = addc(O[i + 8], 0u);
      } else {
        // i odd: a[even] * bi -> odd columns : O at index i + j - 1 ; a[odd] * bi -> even columns : E
        O[i - 1] = mad_lo_cc(a[0], bi, O[i - 1]);
        O[i] = madc_hi_cc(a[0], bi, O[i]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
          O[i + j - 1] = madc_lo_cc(a[j], bi, O[i + j - 1]);
          O[i + j] = madc_hi_cc(a[j], bi, O[i + j]);
        }
        if (i + 7 < 16) O[i + 7] = addc(O[i + 7], 0u);
        E[i + 1] = mad_lo_cc(a[1], bi, E[i + 1]);
        E[i + 2] = madc_hi_cc(a[1], bi, E[i + 2]);
#pragma unroll
        for (int j = 3; j < 8; j += 2) {
          E[i + j] = madc_lo_cc(a[j], bi, E[i + j]);
          E[i + j + 1] = madc_hi_cc(a[j], bi, E[i + j + 1]);
        }
        if (i + 9 < 16) E[i + 9] = addc(E[i + 9], 0u);
      }
    }
    // w = E + (O << 32)
    w[0] = E[0];
    w[1] = add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 15; k++) w[k] = addc_cc(E[k], O[k - 1]);
    w[15] = addc(E[15], O[14]);
  }

  // Montgomery reduction of a 512-bit value t < p * 2^256:  returns t / 2^256 mod p, canonical.
  static HD Fp redc_wide(const uint32_t* t_in) {
    using namespace cc;
    uint32_t t[17];
#pragma unroll
    for (int k = 0; k < 16; k++) t[k] = t_in[k];
    t[16] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint32_t m = mul_lo(t[i], P::INV);
      // t += m * p << (32 i): even limbs of p in one chain, odd limbs in another, then fold the carries up
      uint32_t lo_c, hi_c;
      t[i] = mad_lo_cc(P::MOD(0), m, t[i]);
      t[i + 1] = madc_hi_cc(P::MOD(0), m, t[i + 1]);
#pragma unroll
      for (int j = 2; j < 8; j += 2) {
        t[i + j] = madc_lo_cc(P::MOD(j), m, t[i + j]);
        t[i + j + 1] = madc_hi_cc(P::MOD(j), m, t[i + j + 1]);
      }
      lo_c = addc(0u, 0u);
      t[i + 1] = mad_lo_cc(P::MOD(1), m, t[i + 1]);
      t[i + 2] = madc_hi_cc(P::MOD(1), m, t[i + 2]);
#pragma unroll
      for (int j = 3; j < 8; j += 2) {
        t[i + j] = madc_lo_cc(P::MOD(j), m, t[i + j]);
        t[i + j + 1] = madc_hi_cc(P::MOD(j), m, t[i + j + 1]);
      }
      hi_c = addc(0u, 0u);
      // carries: lo_c belongs to column i + 8, hi_c to column i + 9; ripple through the remaining limbs
      t[i + 8] = add_cc(t[i + 8], lo_c);
      if (i + 9 <= 16) t[i + 9] = addc_cc(t[i + 9], hi_c);
#pragma unroll
      for (int k = i + 10; k <= 16; k++) t[k] = addc_cc(t[k], 0u);
    }
    Fp r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.l[k] = t[8 + k];
    // t < p*2^256 + p*2^256  =>  result < 2p (t[16] is then 0 because 2p < 2^256)
    final_sub(r.l);
    return r;
  }

  HD Fp to_mont() const { return (*this) * r2(); }
  HD Fp from_mont() const {
    Fp o = zero();
    o.l[0] = 1;
    return (*this) * o;
  }

  // a^(p-2): Fermat inversion (0 -> 0).  Same value as fields/fq.go:66-68
  // (big.Int.ModInverse) for every non-zero a.
  HD Fp inverse() const {
#ifdef __CUDA_ARCH__
    return fp_inverse_outlined<P, INL>(*this);
#else
    return inverse_impl();
#endif
  }
  HD Fp inverse_impl() const {
    Fp res = one();
    Fp base = *this;
    for (int w = 0; w < 8; w++) {
      uint32_t e = P::PM2(0);
      // select limb w without dynamic constexpr indexing
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (k == w) e = P::PM2(k);
      for (int b = 0; b < 32; b++) {
        if ((e >> b) & 1) res = res * base;
        base = base.sqr();
      }
    }
    return res;
  }

  // Same value by the binary extended Euclid (right-shift) algorithm on the Montgomery representative:
  // ~250 subtract steps and ~500 halvings of 8-limb integers instead of ~380 dependent Montgomery products
  // -- a ~5x shorter latency chain, for the places where ONE inversion per warp sits on a critical path
  // (k_affine_invert between the forward and backward pass of every round, the final Jacobian -> affine).
  // Data-dependent control flow: give it a warp of its own (one active lane), or accept the divergence.
  // The inputs of those call sites are public group elements; nothing secret-dependent is timed here.
  HD Fp inverse_vartime() const {
#ifdef __CUDA_ARCH__
    return fp_inverse_vartime_outlined<P, INL>(*this);
#else
    return inverse_vartime_impl();
#endif
  }
  static HD void shr1(uint32_t* a) {
#pragma unroll
    for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[7] >>= 1;
  }
  // x <- x / 2 mod p for x in [0, p): (x + p) / 2 when x is odd (x + p < 2^255)
  static HD void halve_mod(uint32_t* x) {
    uint32_t m = 0u - (x[0] & 1u);
    x[0] = cc::add_cc(x[0], P::MOD(0) & m);
#pragma unroll
    for (int i = 1; i < 7; i++) x[i] = cc::addc_cc(x[i], P::MOD(i) & m);
    x[7] = cc::addc(x[7], P::MOD(7) & m);
    shr1(x);
  }
  static HD bool is_one_int(const uint32_t* a) {
    uint32_t o = a[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 8; i++) o |= a[i];
    return o == 0;
  }
  HD Fp inverse_vartime_impl() const {
    if (is_zero()) return zero();
    uint32_t u[8], v[8];
    Fp x1 = zero(), x2 = zero();
    x1.l[0] = 1;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      u[i] = l[i];
      v[i] = P::MOD(i);
    }
    // invariants: x1 * A = u, x2 * A = v (mod p); gcd(u, v) = 1; v odd
    while (!is_one_int(u) && !is_one_int(v)) {
      while (!(u[0] & 1u)) {
        shr1(u);
        halve_mod(x1.l);
      }
      while (!(v[0] & 1u)) {
        shr1(v);
        halve_mod(x2.l);
      }
      uint32_t t[8];
      t[0] = cc::sub_cc(u[0], v[0]);
#pragma unroll
      for (int i = 1; i < 8; i++) t[i] = cc::subc_cc(u[i], v[i]);
      uint32_t borrow = cc::subc(0u, 0u);
      if (!borrow) {  // u >= v
#pragma unroll
        for (int i = 0; i < 8; i++) u[i] = t[i];
        x1 = x1 - x2;
      } else {
        v[0] = cc::sub_cc(v[0], u[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) v[i] = cc::subc_cc(v[i], u[i]);
        x2 = x2 - x1;
      }
    }
    Fp r = is_one_int(u) ? x1 : x2;  // = A^-1 = a^-1 R^-1 (plain inverse of the representative)
    return r * (r2() * r2());       // * R^3 / R -> a^-1 R
  }

  // value >= p ?  (for validating standard-form inputs)
  HD bool geq_modulus() const {
    using namespace cc;
    sub_cc(l[0], P::MOD(0));
#pragma unroll
    for (int i = 1; i < 8; i++) subc_cc(l[i], P::MOD(i));
    uint32_t borrow = subc(0u, 0u);
    return borrow == 0;
  }
};

#ifdef __CUDACC__
template <class P>
__device__ __noinline__ Fp<P, false> fp_mul_outlined(Fp<P, false> a, Fp<P, false> b) {
  return Fp<P, false>::mul_impl(a, b);
}
template <class P, bool INL>
__device__ __noinline__ Fp<P, INL> fp_inverse_outlined(Fp<P, INL> a) {
  // always built on the out-of-line multiply: inversion is never on a hot path
  Fp<P, false> t;
#pragma unroll
  for (int i = 0; i < 8; i++) t.l[i] = a.l[i];
  t = t.inverse_impl();
  Fp<P, INL> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = t.l[i];
  return r;
}
template <class P, bool INL>
__device__ __noinline__ Fp<P, INL> fp_inverse_vartime_outlined(Fp<P, INL> a) {
  Fp<P, false> t;
#pragma unroll
  for (int i = 0; i < 8; i++) t.l[i] = a.l[i];
  t = t.inverse_vartime_impl();
  Fp<P, INL> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = t.l[i];
  return r;
}
#endif

using Fq = Fp<FqParams, false>;   // coordinates of G1 points      (fields over bn128.Q)
using FqH = Fp<FqParams, true>;   // same, multiplication inlined (hot kernels)
using Fr = Fp<FrParams, false>;   // scalars / polynomial coeffs   (fields over bn128.R)
using FrH = Fp<FrParams, true>;

}  // namespace b200
