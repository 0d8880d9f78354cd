// Prove-path orchestration: groth16.GenerateProofs (groth16/groth16.go:225-278)
// and snark.GenerateProofs (snark.go:254-289) as a fixed sequence of MSM launches,
// one exact polynomial division and a tiny finalisation kernel — all enqueued on
// one stream, one host synchronisation per proof.
//
// Included inside capi.cu's anonymous namespace (uses its context globals).
//
// Groth16 algebra.  With pk base sets extended by the blinding
// points, the reference's
//     PiA  = sum w_i At_i + Alpha + r*Delta                       (:243,253-255)
//     piB1 = sum w_i B1_i + Beta1 + s*Delta                       (:244,259-262)
//     PiB  = sum w_i B2_i + Beta2 + s*Delta2                      (:245,260-264)
//     PiC  = sum_{i>l} w_i C_i + sum h_i PTD_i + s*PiA + r*piB1 - rs*Delta   (:248-275)
// become four MSMs:
//     A-set  = At ++ [Alpha, Delta]      scalars  w ++ [1, r]
//     B1-set = B1 ++ [Beta1, Delta]      scalars  w ++ [1, s]
//     B2-set = B2 ++ [Beta2, Delta2]     scalars  w ++ [1, s]
//     CH-set = C[l+1..m) ++ PTD ++ [Delta]   scalars  w[l+1..m) ++ h ++ 0.. ++ [-rs]
// plus the two variable-base products s*PiA and r*piB1 (k_groth16_finalize).

struct ProvingKey {
  int kind = 0;  // 1 Groth16, 2 Pinocchio
  size_t m = 0, npublic = 0, n_h_bases = 0;
  std::unique_ptr<Bases> g[8];  // Groth16: A, B1, B2(G2), CH   Pinocchio: A, Ap, B(G2), Bp, C, Cp, Kp, H
  Divisor Z;
  DevBuf s1, s2, s3;   // scalar vectors
  DevBuf px;           // px coefficients (device)
  DevBuf res;          // XYZZ results (8 x 256 B)
  DevBuf out_std;      // standard-form Jacobian outputs
  DevBuf rs;           // r, s on device (standard form)
};

std::map<uint64_t, std::unique_ptr<ProvingKey>> g_pks;
uint64_t g_next_pk = 1;

// p <- k * p for a 256-bit standard-form scalar (MSB-first double-and-add)
template <class F>
__device__ XYZZ<F> xyzz_mul_scalar(const XYZZ<F>& p, const Fr& k) {
  XYZZ<F> r = XYZZ<F>::inf();
  bool started = false;
  for (int w = 7; w >= 0; w--) {
    uint32_t limb = k.l[w];
    for (int b = 31; b >= 0; b--) {
      uint32_t bit = (limb >> b) & 1;
      if (!started && !bit) continue;
      started = true;
      r = xyzz_dbl(r);
      if (bit) xyzz_add(r, p);
    }
  }
  return r;
}

template <class F>
__device__ void store_jacobian_std(const XYZZ<F>& p, F* out) {
  Jacobian<F> j = xyzz_to_jacobian(p);
  out[0] = j.X.from_mont();
  out[1] = j.Y.from_mont();
  out[2] = j.Z.from_mont();
}

// res: [0] A (G1 XYZZ @ +0), [1] B1 (@ +256), [2] B2 (G2 XYZZ @ +512), [3] CH (@ +768)
// out: PiA (3 Fq) | PiC (3 Fq) | PiB (3 Fq2)
__global__ void k_groth16_finalize(const uint8_t* res, const Fr* rs, Fq* out_a, Fq* out_c, Fq2* out_b) {
  __shared__ XYZZ<Fq> prod[2];
  const XYZZ<Fq>& A = *reinterpret_cast<const XYZZ<Fq>*>(res);
  const XYZZ<Fq>& B1 = *reinterpret_cast<const XYZZ<Fq>*>(res + 256);
  const XYZZ<Fq2>& B2 = *reinterpret_cast<const XYZZ<Fq2>*>(res + 512);
  const XYZZ<Fq>& CH = *reinterpret_cast<const XYZZ<Fq>*>(res + 768);
  uint32_t t = threadIdx.x;
  if (t == 0) prod[0] = xyzz_mul_scalar(A, rs[1]);    // s * PiA     (groth16.go:272)
  if (t == 32) prod[1] = xyzz_mul_scalar(B1, rs[0]);  // r * piBG1   (groth16.go:273)
  if (t == 64) {
    store_jacobian_std(A, out_a);
    store_jacobian_std(B2, out_b);
  }
  __syncthreads();
  if (t == 0) {
    XYZZ<Fq> c = CH;
    xyzz_add(c, prod[0]);
    xyzz_add(c, prod[1]);
    store_jacobian_std(c, out_c);
  }
}

// out[k] = Jacobian(res[k]) for the 7 G1 results and the G2 result of Pinocchio
__global__ void k_pinocchio_finalize(const uint8_t* res, Fq* out_g1, Fq2* out_b) {
  uint32_t t = threadIdx.x;
  if (t < 7) store_jacobian_std(*reinterpret_cast<const XYZZ<Fq>*>(res + 256 * t), out_g1 + 3 * t);
  if (t == 32) store_jacobian_std(*reinterpret_cast<const XYZZ<Fq2>*>(res + 256 * 7), out_b);
}

// host: concatenate Jacobian point arrays (words per point = 12 or 24)
struct PointCat {
  std::vector<uint64_t> v;
  size_t words;
  explicit PointCat(size_t w) : words(w) {}
  void add(const uint64_t* p, size_t n) { v.insert(v.end(), p, p + n * words); }
  size_t count() const { return v.size() / words; }
};

int pk_common_init(ProvingKey& pk, const uint64_t* z, size_t nz, size_t m) {
  if (!z || nz == 0) return fail(B200_EINVAL, "pk_load: missing Z");
  int rc = divisor_init(pk.Z, z, nz);
  if (rc) return rc;
  CU(pk.s1.alloc((m + 4) * sizeof(Fr)));
  CU(pk.s2.alloc((m + 4) * sizeof(Fr)));
  CU(pk.res.alloc(8 * 256));
  CU(pk.out_std.alloc(64 * sizeof(Fq)));
  CU(pk.rs.alloc(4 * sizeof(Fr)));
  return check_err_flag<Fr>("pk_load(Z)");
}

int groth16_pk_load(const uint64_t* at, const uint64_t* b1, const uint64_t* b2, const uint64_t* bacdelta, size_t m,
                    const uint64_t* ptd, size_t n_ptd, const uint64_t* z, size_t nz, const uint64_t* alpha1,
                    const uint64_t* beta1, const uint64_t* delta1, const uint64_t* beta2, const uint64_t* delta2,
                    size_t npublic, int c, b200_pk_t* out) {
  if (!at || !b1 || !b2 || !bacdelta || !ptd || !alpha1 || !beta1 || !delta1 || !beta2 || !delta2 || !out)
    return fail(B200_EINVAL, "groth16_pk_load: null pointer");
  if (m == 0 || npublic + 1 > m || n_ptd == 0) return fail(B200_EINVAL, "groth16_pk_load: bad sizes");
  auto pk = std::make_unique<ProvingKey>();
  pk->kind = 1;
  pk->m = m;
  pk->npublic = npublic;
  pk->n_h_bases = n_ptd;
  int rc = pk_common_init(*pk, z, nz, m);
  if (rc) return rc;
  {
    PointCat cat(12);
    cat.add(at, m); cat.add(alpha1, 1); cat.add(delta1, 1);
    if ((rc = bases_create<Fq>(cat.v.data(), cat.count(), c, 1, pk->g[0]))) return rc;
  }
  {
    PointCat cat(12);
    cat.add(b1, m); cat.add(beta1, 1); cat.add(delta1, 1);
    if ((rc = bases_create<Fq>(cat.v.data(), cat.count(), c, 1, pk->g[1]))) return rc;
  }
  {
    PointCat cat(24);
    cat.add(b2, m); cat.add(beta2, 1); cat.add(delta2, 1);
    if ((rc = bases_create<Fq2>(cat.v.data(), cat.count(), c, 2, pk->g[2]))) return rc;
  }
  {
    PointCat cat(12);
    cat.add(bacdelta + 12 * (npublic + 1), m - npublic - 1); cat.add(ptd, n_ptd); cat.add(delta1, 1);
    if ((rc = bases_create<Fq>(cat.v.data(), cat.count(), c, 1, pk->g[3]))) return rc;
  }
  CU(pk->s3.alloc((m + n_ptd + 4) * sizeof(Fr)));
  uint64_t h = g_next_pk++;
  g_pks[h] = std::move(pk);
  *out = h;
  return B200_OK;
}

ProvingKey* find_pk(b200_pk_t h, int kind) {
  auto it = g_pks.find(h);
  if (it == g_pks.end() || it->second->kind != kind) return nullptr;
  return it->second.get();
}

// host F_r helpers for the blinding scalars (the reference does this with big.Int, groth16.go:274)
Fr fr_load_std(const uint64_t* v) {
  Fr r;
  memcpy(&r, v, sizeof(Fr));
  return r;
}

int groth16_prove(b200_pk_t h, const uint64_t* w, size_t nw, const uint64_t* px, size_t npx, const uint64_t* r,
                  const uint64_t* s, uint64_t* pi_a, uint64_t* pi_b, uint64_t* pi_c) {
  ProvingKey* pk = find_pk(h, 1);
  if (!pk) return fail(B200_EINVAL, "groth16_prove: bad proving-key handle");
  if (!w || !px || !r || !s || !pi_a || !pi_b || !pi_c) return fail(B200_EINVAL, "groth16_prove: null pointer");
  size_t m = pk->m, l1 = pk->npublic + 1;
  if (nw != m) return fail(B200_EINVAL, "groth16_prove: witness length %zu != NVars %zu", nw, m);
  if (npx < pk->Z.nb) return fail(B200_EINVAL, "groth16_prove: len(px) < len(Z)");
  size_t nq = npx - pk->Z.nb + 1;
  // groth16.go:269-270 indexes PowersTauDelta[i] for i < len(hx): out of range panics in the reference
  if (nq > pk->n_h_bases) return fail(B200_EINVAL, "groth16_prove: len(hx)=%zu exceeds len(PowersTauDelta)=%zu", nq, pk->n_h_bases);
  Fr fr_r = fr_load_std(r), fr_s = fr_load_std(s);
  if (fr_r.geq_modulus() || fr_s.geq_modulus()) return fail(B200_ERANGE, "groth16_prove: r or s >= field order");
  Fr neg_rs = (fr_r.to_mont() * fr_s.to_mont()).neg().from_mont();  // -(r*s) mod r  (groth16.go:274)
  Fr one = Fr::zero();
  one.l[0] = 1;
  cudaStream_t st = g_stream;
  Fr* sA = pk->s1.as<Fr>();
  Fr* sB = pk->s2.as<Fr>();
  Fr* sCH = pk->s3.as<Fr>();
  size_t n_c = m - l1, n_ch = n_c + pk->n_h_bases + 1;
  Fr tailA[2] = {one, fr_r}, tailB[2] = {one, fr_s}, rs_host[2] = {fr_r, fr_s};
  CU(pk->px.ensure(npx * sizeof(Fr)));
  CU(cudaMemcpyAsync(sA, w, m * sizeof(Fr), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(pk->px.p, px, npx * sizeof(Fr), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(sA + m, tailA, sizeof tailA, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(sB + m, tailB, sizeof tailB, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(pk->rs.p, rs_host, sizeof rs_host, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(sCH + n_ch - 1, &neg_rs, sizeof(Fr), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(sB, sA, m * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
  if (n_c) CU(cudaMemcpyAsync(sCH, sA + l1, n_c * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
  if (pk->n_h_bases > nq) CU(cudaMemsetAsync(sCH + n_c + nq, 0, (pk->n_h_bases - nq) * sizeof(Fr), st));
  uint8_t* res = pk->res.as<uint8_t>();
  int rc;
  if ((rc = msm_enqueue<Fq>(pk->g[0].get(), sA, m + 2, 0, reinterpret_cast<XYZZ<Fq>*>(res), st))) return rc;
  if ((rc = msm_enqueue<Fq>(pk->g[1].get(), sB, m + 2, 0, reinterpret_cast<XYZZ<Fq>*>(res + 256), st))) return rc;
  if ((rc = msm_enqueue<Fq2>(pk->g[2].get(), sB, m + 2, 0, reinterpret_cast<XYZZ<Fq2>*>(res + 512), st))) return rc;
  // hx = px / Z  (groth16.go:266) written straight into the CH scalar vector
  CU(poly_div_device(*g_poly, pk->Z, pk->px.as<Fr>(), npx, 0, sCH + n_c, nullptr, g_d_err, st));
  if ((rc = msm_enqueue<Fq>(pk->g[3].get(), sCH, n_ch, 0, reinterpret_cast<XYZZ<Fq>*>(res + 768), st))) return rc;
  Fq* o = pk->out_std.as<Fq>();
  k_groth16_finalize<<<1, 96, 0, st>>>(res, pk->rs.as<Fr>(), o, o + 3, reinterpret_cast<Fq2*>(o + 6));
  CU(cudaGetLastError());
  uint64_t host_out[48];
  CU(cudaMemcpyAsync(host_out, o, sizeof host_out, cudaMemcpyDeviceToHost, st));
  rc = check_err_flag<Fr>("groth16_prove");  // synchronises
  if (rc) return rc;
  memcpy(pi_a, host_out, 12 * 8);
  memcpy(pi_c, host_out + 12, 12 * 8);
  memcpy(pi_b, host_out + 24, 24 * 8);
  return B200_OK;
}

int pinocchio_pk_load(const uint64_t* a, const uint64_t* ap, const uint64_t* b2, const uint64_t* bp,
                      const uint64_t* c, const uint64_t* cp, const uint64_t* kp, size_t m, const uint64_t* g1t,
                      size_t n_g1t, const uint64_t* z, size_t nz, size_t npublic, int wb, b200_pk_t* out) {
  if (!a || !ap || !b2 || !bp || !c || !cp || !kp || !g1t || !out)
    return fail(B200_EINVAL, "pinocchio_pk_load: null pointer");
  if (m == 0 || npublic + 1 > m || n_g1t == 0) return fail(B200_EINVAL, "pinocchio_pk_load: bad sizes");
  auto pk = std::make_unique<ProvingKey>();
  pk->kind = 2;
  pk->m = m;
  pk->npublic = npublic;
  pk->n_h_bases = n_g1t;
  int rc = pk_common_init(*pk, z, nz, m);
  if (rc) return rc;
  size_t l1 = npublic + 1;
  // snark.go:265-268: PiA, PiAp run over i in [NPublic+1, NVars)
  if (m > l1) {
    if ((rc = bases_create<Fq>(a + 12 * l1, m - l1, wb, 1, pk->g[0]))) return rc;
    if ((rc = bases_create<Fq>(ap + 12 * l1, m - l1, wb, 1, pk->g[1]))) return rc;
  }
  if ((rc = bases_create<Fq2>(b2, m, wb, 2, pk->g[2]))) return rc;
  if ((rc = bases_create<Fq>(bp, m, wb, 1, pk->g[3]))) return rc;
  if ((rc = bases_create<Fq>(c, m, wb, 1, pk->g[4]))) return rc;
  if ((rc = bases_create<Fq>(cp, m, wb, 1, pk->g[5]))) return rc;
  if ((rc = bases_create<Fq>(kp, m, wb, 1, pk->g[6]))) return rc;
  if ((rc = bases_create<Fq>(g1t, n_g1t, wb, 1, pk->g[7]))) return rc;
  CU(pk->s3.alloc((n_g1t + 4) * sizeof(Fr)));
  uint64_t h = g_next_pk++;
  g_pks[h] = std::move(pk);
  *out = h;
  return B200_OK;
}

int pinocchio_prove(b200_pk_t h, const uint64_t* w, size_t nw, const uint64_t* px, size_t npx, uint64_t* out_g1,
                    uint64_t* pi_b) {
  ProvingKey* pk = find_pk(h, 2);
  if (!pk) return fail(B200_EINVAL, "pinocchio_prove: bad proving-key handle");
  if (!w || !px || !out_g1 || !pi_b) return fail(B200_EINVAL, "pinocchio_prove: null pointer");
  size_t m = pk->m, l1 = pk->npublic + 1;
  if (nw != m) return fail(B200_EINVAL, "pinocchio_prove: witness length %zu != NVars %zu", nw, m);
  if (npx < pk->Z.nb) return fail(B200_EINVAL, "pinocchio_prove: len(px) < len(Z)");
  size_t nq = npx - pk->Z.nb + 1;
  if (nq > pk->n_h_bases) return fail(B200_EINVAL, "pinocchio_prove: len(hx)=%zu exceeds len(G1T)=%zu", nq, pk->n_h_bases);
  cudaStream_t st = g_stream;
  Fr* dw = pk->s1.as<Fr>();
  Fr* dh = pk->s3.as<Fr>();
  CU(pk->px.ensure(npx * sizeof(Fr)));
  CU(cudaMemcpyAsync(dw, w, m * sizeof(Fr), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(pk->px.p, px, npx * sizeof(Fr), cudaMemcpyHostToDevice, st));
  uint8_t* res = pk->res.as<uint8_t>();
  auto R1 = [&](int k) { return reinterpret_cast<XYZZ<Fq>*>(res + 256 * k); };
  int rc;
  // result slots: 0 PiA, 1 PiAp, 2 PiBp, 3 PiC, 4 PiCp, 5 PiH, 6 PiKp, 7 PiB(G2)
  if (m > l1) {
    if ((rc = msm_enqueue<Fq>(pk->g[0].get(), dw + l1, m - l1, 0, R1(0), st))) return rc;
    if ((rc = msm_enqueue<Fq>(pk->g[1].get(), dw + l1, m - l1, 0, R1(1), st))) return rc;
  } else {
    CU(cudaMemsetAsync(res, 0, 512, st));
  }
  if ((rc = msm_enqueue<Fq2>(pk->g[2].get(), dw, m, 0, reinterpret_cast<XYZZ<Fq2>*>(res + 256 * 7), st))) return rc;
  if ((rc = msm_enqueue<Fq>(pk->g[3].get(), dw, m, 0, R1(2), st))) return rc;
  if ((rc = msm_enqueue<Fq>(pk->g[4].get(), dw, m, 0, R1(3), st))) return rc;
  if ((rc = msm_enqueue<Fq>(pk->g[5].get(), dw, m, 0, R1(4), st))) return rc;
  if ((rc = msm_enqueue<Fq>(pk->g[6].get(), dw, m, 0, R1(6), st))) return rc;
  CU(poly_div_device(*g_poly, pk->Z, pk->px.as<Fr>(), npx, 0, dh, nullptr, g_d_err, st));   // snark.go:280
  if ((rc = msm_enqueue<Fq>(pk->g[7].get(), dh, nq, 0, R1(5), st))) return rc;               // snark.go:284-286
  Fq* o = pk->out_std.as<Fq>();
  k_pinocchio_finalize<<<1, 64, 0, st>>>(res, o, reinterpret_cast<Fq2*>(o + 21));
  CU(cudaGetLastError());
  uint64_t host_out[(21 + 6) * 4];
  CU(cudaMemcpyAsync(host_out, o, sizeof host_out, cudaMemcpyDeviceToHost, st));
  rc = check_err_flag<Fr>("pinocchio_prove");
  if (rc) return rc;
  memcpy(out_g1, host_out, 84 * 8);
  memcpy(pi_b, host_out + 84, 24 * 8);
  return B200_OK;
}
