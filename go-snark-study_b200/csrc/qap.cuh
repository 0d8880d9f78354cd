// Dense-API kernels for the QAP step upstream of GenerateProofs, and the small
// element-wise group / polynomial operations of the reference's API surface.
//
//   R1CSToQAP           r1csqap/r1csqap.go:161-188 (+ LagrangeInterpolation :150-158, NewPolZeroAt :129-147)
//   CombinePolynomials  r1csqap/r1csqap.go:191-210
//   Add / Sub / Eval    r1csqap/r1csqap.go:94-126
//   G1/G2 Add, Double, Affine  bn128/g1.go:32-170, g2.go:32-200 (reference formulas, X,Y,Z-exact)
//
// The reference interpolates every column with O(n^2) schoolbook products per
// point (O(m n^3) total) and overflows a native int for n > 21.  Here
// the n Lagrange basis polynomials over the domain {1..n} are built once
// (L_j = Z_n / ((x - j) d_j), synthetic division, one thread per j) and every
// column is a (sparse) combination of them; exact in F_r, hence equal to the
// reference's coefficients wherever the reference is correct (n <= 21), and the
// mathematically intended result beyond.  Dense storage limits this API to
// n <~ 2^12; the prove path itself never needs it (it consumes px).
#pragma once
#include <cuda_runtime.h>

#include "ec.cuh"

namespace b200 {

// zn[0..n] <- coefficients of prod_{i=1..n} (x - i), Montgomery.  One block.
__global__ void k_zero_poly(Fr* zn, uint32_t n) {
  uint32_t t = threadIdx.x, T = blockDim.x;
  for (uint32_t k = t; k <= n; k += T) zn[k] = k == 0 ? Fr::one() : Fr::zero();
  __syncthreads();
  // after step i the polynomial has degree i; new[k] = old[k-1] - i*old[k]
  for (uint32_t i = 1; i <= n; i++) {
    Fr iv = Fr::zero();
    iv.l[0] = i;
    iv = iv.to_mont();
    // process high -> low in chunks so each thread reads old values before they are overwritten
    Fr nv[8];
    uint32_t cnt = 0;
    for (uint32_t k = t; k <= i; k += T) {
      Fr lo = k ? zn[k - 1] : Fr::zero();
      Fr cur = k < i ? zn[k] : Fr::zero();
      if (cnt < 8) nv[cnt++] = lo - iv * cur;
    }
    __syncthreads();
    cnt = 0;
    for (uint32_t k = t; k <= i; k += T) zn[k] = nv[cnt++];
    __syncthreads();
  }
}

// L[j][0..n) <- coefficients of the Lagrange basis polynomial that is 1 at x = j+1 and 0 at the
// other points of {1..n}.  One thread per j.
__global__ void k_lagrange_basis(const Fr* __restrict__ zn, uint32_t n, Fr* __restrict__ L) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  Fr xj = Fr::zero();
  xj.l[0] = j + 1;
  xj = xj.to_mont();
  // d_j = prod_{i != j} (x_j - x_i)
  Fr d = Fr::one();
  for (uint32_t i = 1; i <= n; i++) {
    if (i == j + 1) continue;
    Fr xi = Fr::zero();
    xi.l[0] = i;
    d = d * (xj - xi.to_mont());
  }
  Fr dinv = d.inverse();
  // synthetic division of Z_n by (x - x_j): q_{n-1} = 1, q_{k-1} = zn_k + x_j q_k
  Fr q = zn[n];
  Fr* row = L + (size_t)j * n;
  row[n - 1] = q * dinv;
  for (uint32_t k = n - 1; k >= 1; k--) {
    q = zn[k] + xj * q;
    row[k - 1] = q * dinv;
  }
}

// out[i][k] = sum_j M[j][i] * L[j][k]   (M: n x m standard form, row-major; out: m x n standard form)
__global__ void k_qap_interpolate(const Fr* __restrict__ M, uint32_t n, uint32_t m, const Fr* __restrict__ L,
                                  Fr* __restrict__ out, int* err) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t i = blockIdx.y;
  if (k >= n || i >= m) return;
  Fr acc = Fr::zero();
  for (uint32_t j = 0; j < n; j++) {
    Fr v = M[(size_t)j * m + i];
    if (v.is_zero()) continue;  // R1CS columns are sparse
    if (v.geq_modulus()) {
      atomicOr(err, 2);
      continue;
    }
    acc = acc + v.to_mont() * L[(size_t)j * n + k];
  }
  out[(size_t)i * n + k] = acc.from_mont();
}

// out[k] = sum_i r[i] * P[i][k]   (P: m rows of n coefficients, standard form), Montgomery out
__global__ void k_combine(const Fr* __restrict__ r, uint32_t m, const Fr* __restrict__ P, uint32_t n,
                          Fr* __restrict__ out_mont, Fr* __restrict__ out_std, int* err) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  Fr acc = Fr::zero();
  for (uint32_t i = 0; i < m; i++) {
    Fr ri = r[i], v = P[(size_t)i * n + k];
    if (ri.is_zero() || v.is_zero()) continue;
    if (ri.geq_modulus() || v.geq_modulus()) {
      atomicOr(err, 2);
      continue;
    }
    acc = acc + ri.to_mont() * v.to_mont();
  }
  out_mont[k] = acc;
  out_std[k] = acc.from_mont();
}

// out[i] = a[i] (+/-) b[i] with the reference's max-length semantics (standard form in/out)
__global__ void k_poly_addsub(const Fr* a, uint32_t na, const Fr* b, uint32_t nb, int sub, Fr* out, int* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t n = na > nb ? na : nb;
  if (i >= n) return;
  Fr x = i < na ? a[i] : Fr::zero(), y = i < nb ? b[i] : Fr::zero();
  if (x.geq_modulus() || y.geq_modulus()) {
    atomicOr(err, 2);
    return;
  }
  out[i] = sub ? x - y : x + y;  // add/sub are form-agnostic (linear)
}

// out = sum_i v[i] x^i.  One block; each thread sums a strided subset with its own power chain.
__global__ void k_poly_eval(const Fr* __restrict__ v, uint32_t n, Fr x_std, Fr* out, int* err) {
  __shared__ Fr part[256];
  uint32_t t = threadIdx.x, T = blockDim.x;
  Fr x = x_std.to_mont();
  // x^t and x^T
  Fr xp = Fr::one(), base = x;
  for (uint32_t e = t; e; e >>= 1) {
    if (e & 1) xp = xp * base;
    base = base.sqr();
  }
  Fr xT = Fr::one();
  base = x;
  for (uint32_t e = T; e; e >>= 1) {
    if (e & 1) xT = xT * base;
    base = base.sqr();
  }
  Fr acc = Fr::zero();
  for (uint32_t i = t; i < n; i += T) {
    Fr c = v[i];
    if (c.geq_modulus()) atomicOr(err, 2);
    else acc = acc + c.to_mont() * xp;
    xp = xp * xT;
  }
  part[t] = acc;
  __syncthreads();
  for (uint32_t s = T / 2; s > 0; s >>= 1) {
    if (t < s) part[t] = part[t] + part[t + s];
    __syncthreads();
  }
  if (t == 0) out[0] = part[0].from_mont();
}

// out[i] = sum_k P[i][k] x^k for m polynomials of n coefficients (row-major, standard form): one block per
// polynomial.  Serves the trusted setup's At / Bt / Ct = Eval(alphas[i], tau) loops (groth16/groth16.go:164-205).
__global__ void k_poly_eval_batch(const Fr* __restrict__ P, uint32_t n, Fr x_std, Fr* __restrict__ out, int* err) {
  __shared__ Fr part[256];
  const Fr* v = P + (size_t)blockIdx.x * n;
  uint32_t t = threadIdx.x, T = blockDim.x;
  Fr x = x_std.to_mont();
  Fr xp = Fr::one(), base = x;
  for (uint32_t e = t; e; e >>= 1) {
    if (e & 1) xp = xp * base;
    base = base.sqr();
  }
  Fr xT = Fr::one();
  base = x;
  for (uint32_t e = T; e; e >>= 1) {
    if (e & 1) xT = xT * base;
    base = base.sqr();
  }
  Fr acc = Fr::zero();
  for (uint32_t i = t; i < n; i += T) {
    Fr c = v[i];
    if (c.geq_modulus()) atomicOr(err, 2);
    else if (!c.is_zero()) acc = acc + c.to_mont() * xp;
    xp = xp * xT;
  }
  part[t] = acc;
  __syncthreads();
  for (uint32_t s = T / 2; s > 0; s >>= 1) {
    if (t < s) part[t] = part[t] + part[t + s];
    __syncthreads();
  }
  if (t == 0) out[blockIdx.x] = part[0].from_mont();
}

// ---- element-wise group operations with the reference's own formulas --------
// op 0: Add(p, q)   op 1: Double(p)   op 2: Neg(p)     (Jacobian standard form in/out)
template <class F>
__global__ void k_group_op(int op, const F* __restrict__ p, const F* __restrict__ q, size_t n, F* __restrict__ out, int* err) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  F px = p[3 * i], py = p[3 * i + 1], pz = p[3 * i + 2];
  bool bad = px.geq_modulus() || py.geq_modulus() || pz.geq_modulus();
  Jacobian<F> P{px.to_mont(), py.to_mont(), pz.to_mont()}, r;
  if (op == 0) {
    F qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
    bad = bad || qx.geq_modulus() || qy.geq_modulus() || qz.geq_modulus();
    Jacobian<F> Q{qx.to_mont(), qy.to_mont(), qz.to_mont()};
    r = jac_add_ref(P, Q);
  } else if (op == 1) {
    r = jac_double_ref(P);
  } else {
    r = Jacobian<F>{P.X, P.Y.neg(), P.Z};
  }
  if (bad) atomicOr(err, 1);
  out[3 * i] = r.X.from_mont();
  out[3 * i + 1] = r.Y.from_mont();
  out[3 * i + 2] = r.Z.from_mont();
}
// Affine(p): (x, y) standard form; infinity -> (0, 0)   (g1.go:157-170)
template <class F>
__global__ void k_group_affine(const F* __restrict__ p, size_t n, F* __restrict__ out, int* err) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  F px = p[3 * i], py = p[3 * i + 1], pz = p[3 * i + 2];
  if (px.geq_modulus() || py.geq_modulus() || pz.geq_modulus()) atomicOr(err, 1);
  Affine<F> a = jac_to_affine(Jacobian<F>{px.to_mont(), py.to_mont(), pz.to_mont()});
  out[2 * i] = a.x.from_mont();
  out[2 * i + 1] = a.y.from_mont();
}

}  // namespace b200
