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
// become four MSMs, the first three over ONE scalar vector W = w ++ [1, r, s] (O = infinity):
//     A-set  = At ++ [Alpha, Delta,  O     ]
//     B1-set = B1 ++ [Beta1, O,      Delta ]
//     B2-set = B2 ++ [Beta2, O,      Delta2]
//     CH-set = C[l+1..m) ++ PTD ++ [Delta]   scalars  w[l+1..m) ++ h ++ 0.. ++ [-rs]
// so the digit recode + counting sort of W is done once and shared by A, B1 and B2
// (the sorted entry list depends only on the scalars).  The two variable-base
// products s*PiA and r*piB1 run on a side stream while B2 / CH are still accumulating.
//
// Streams: the four bucket phases are independent, so they are issued on four streams
// (caller's + 3 side streams): the latency-bound tails (bucket reduction, tree sum) of
// one MSM overlap the throughput-bound accumulation of the others.

struct ProvingKey {
  int kind = 0;  // 1 Groth16, 2 Pinocchio
  size_t m = 0, npublic = 0, n_h_bases = 0;
  // Work shard of this rank (world == 1: everything).  The four MSMs are laid end to end on a line weighted
  // by cost (a G2 term ~2.8 G1 terms) and rank g takes the g-th of `world` equal pieces, so a rank holds
  // whole MSMs where it can and index ranges where it must: fewer sorts / reduction tails per rank than
  // slicing every MSM `world` ways.  set k: 0 A, 1 B1, 2 B2, 3 C||PTD.
  int rank = 0, world = 1;
  int ctx = 0;  // prove context (B200_CFG_PK_CONTEXT at load): side streams + polynomial workspace this key's proofs run on
  cudaStream_t* side() const { return ctx ? g_side1 : g_side; }
  cudaStream_t main() const { return ctx ? g_stream1 : g_stream; }
  PolyCtx& poly() const { return ctx ? *g_poly1 : *g_poly; }
  size_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};  // index range inside set k
  bool tail[4] = {false, false, false, false};        // this rank also holds set k's blinding points
  size_t n_c_full = 0;                                // length of the C part of the C||PTD set (m - npublic - 1)
  int sort_src[3] = {0, 1, 2};                        // set k reuses the digit sort of set sort_src[k] (same scalars)
  DevBuf s4;                                          // extra scalar vector (non-shared case)
  DevBuf h_full;              // full quotient (sharded mode)
  DevBuf gather;              // all-gathered partial records (world x 1 KB), when a communicator is active
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // pinned landing area of the host-pointer entry points: 48 words of proof + the error flags.  Pinned, so the D2H copies
  // are truly asynchronous and the call can wait for them with the library mutex released (a pageable destination
  // makes cudaMemcpyAsync block until the proof is done — under the mutex).
  uint64_t* h_out = nullptr;
  ~ProvingKey() {
    for (auto e : ev) if (e) cudaEventDestroy(e);
    if (h_out) cudaFreeHost(h_out);
  }
  std::unique_ptr<Bases> g[8];  // Groth16: A, B1, B2(G2), CH   Pinocchio: A, Ap, B(G2), Bp, C, Cp, Kp, H
  Divisor Z;
  DevBuf s1, s2, s3;   // scalar vectors
  DevBuf px, w_stage;  // px coefficients / witness staging (host-pointer entry points)
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

// Partial-result record of one rank: A (G1 XYZZ @ +0), B1 (@ +256), B2 (G2 XYZZ @ +512),
// CH (@ +768) — 1 KB; `parts` holds nparts consecutive records (1 on a single GPU,
// world_size after the NCCL all-gather).  rs = {r, s} standard form.
// out: PiA (3 Fq) | PiC (3 Fq) | PiB (3 Fq2), Jacobian, standard form.
constexpr size_t kPartialBytes = 1024;
// Sharded mode.  Each rank has already folded s*A_part + r*B1_part into its C part (linear in the partial
// sums, so this is done per rank, overlapped with its other MSMs); what remains after the all-gather is
// three point sums: A (slot @+0), B2 (@+512), C (@+768).
__global__ void k_groth16_finalize(const uint8_t* parts, int nparts, Fq* out_a, Fq* out_c, Fq2* out_b) {
  uint32_t t = threadIdx.x;
  if (t == 0 || t == 32) {
    int off = t == 0 ? 0 : 768;
    XYZZ<Fq> acc = *reinterpret_cast<const XYZZ<Fq>*>(parts + off);
    for (int p = 1; p < nparts; p++) xyzz_add(acc, *reinterpret_cast<const XYZZ<Fq>*>(parts + kPartialBytes * p + off));
    store_jacobian_std(acc, t == 0 ? out_a : out_c);
  }
  if (t == 64) {
    XYZZ<Fq2> acc = *reinterpret_cast<const XYZZ<Fq2>*>(parts + 512);
    for (int p = 1; p < nparts; p++) xyzz_add(acc, *reinterpret_cast<const XYZZ<Fq2>*>(parts + kPartialBytes * p + 512));
    store_jacobian_std(acc, out_b);
  }
}
// in place: CH part (@+768) += prod[0] + prod[1]
__global__ void k_groth16_fold_products(uint8_t* res, const XYZZ<Fq>* prod) {
  if (threadIdx.x | blockIdx.x) return;
  XYZZ<Fq> c = *reinterpret_cast<const XYZZ<Fq>*>(res + 768);
  xyzz_add(c, prod[0]);
  xyzz_add(c, prod[1]);
  *reinterpret_cast<XYZZ<Fq>*>(res + 768) = c;
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
  pk.ctx = g_pk_ctx;
  int rc = divisor_init(pk.Z, z, nz);
  if (rc) return rc;
  CU(pk.s1.alloc((m + 4) * sizeof(Fr)));
  CU(pk.s2.alloc((m + 4) * sizeof(Fr)));
  CU(pk.res.alloc(8 * 256));
  CU(pk.out_std.alloc(64 * sizeof(Fq)));
  CU(pk.rs.alloc(8 * sizeof(Fr)));
  for (auto& e : pk.ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CU(cudaHostAlloc(reinterpret_cast<void**>(&pk.h_out), 64 * sizeof(uint64_t), cudaHostAllocDefault));
  return check_err_flag<Fr>("pk_load(Z)");
}

int groth16_pk_load(const uint64_t* at, const uint64_t* b1, const uint64_t* b2, const uint64_t* bacdelta, size_t m,
                    const uint64_t* ptd, size_t n_ptd, const uint64_t* z, size_t nz, const uint64_t* alpha1,
                    const uint64_t* beta1, const uint64_t* delta1, const uint64_t* beta2, const uint64_t* delta2,
                    size_t npublic, int c, int rank, int world, b200_pk_t* out) {
  if (!at || !b1 || !b2 || !bacdelta || !ptd || !alpha1 || !beta1 || !delta1 || !beta2 || !delta2 || !out)
    return fail(B200_EINVAL, "groth16_pk_load: null pointer");
  if (m == 0 || npublic + 1 > m || n_ptd == 0) return fail(B200_EINVAL, "groth16_pk_load: bad sizes");
  if (world < 1 || rank < 0 || rank >= world) return fail(B200_EINVAL, "groth16_pk_load: bad shard %d/%d", rank, world);
  auto pk = std::make_unique<ProvingKey>();
  pk->kind = 1;
  pk->m = m;
  pk->npublic = npublic;
  pk->n_h_bases = n_ptd;
  int rc = pk_common_init(*pk, z, nz, m);
  if (rc) return rc;
  pk->rank = rank;
  pk->world = world;
  const size_t l1 = npublic + 1;
  const size_t n_c_full = pk->n_c_full = m - l1;
  const size_t len[4] = {m, m, m, n_c_full + n_ptd};
  // Cost weights of the line partition, in G1 terms of the C||PTD set: a G2 term (B2), and a term of A / B1 — whose ranks
  // also pay the two GLV products s*A_part + r*B1_part (1 ms of single-warp latency) on their critical path when sharded.
  const double w_ab = world > 1 ? g_w_ab / 100.0 : 1.0;
  const double wgt[4] = {w_ab, w_ab, g_w_g2 / 100.0, 1.0};
  std::vector<ShardCut> cuts((size_t)world);
  shard_partition(len, wgt, world, cuts.data());   // equal pieces of the weighted line (shard_partition.h)
  for (int k = 0; k < 4; k++) {
    pk->lo[k] = cuts[(size_t)rank].lo[k];
    pk->hi[k] = cuts[(size_t)rank].hi[k];
    pk->tail[k] = pk->hi[k] == len[k] && (pk->lo[k] < pk->hi[k] || (rank == world - 1 && len[k] == 0));
  }
  // a set whose end falls exactly on a cut: the rank holding its last element owns the tail (checked above);
  // if no rank holds elements of it (len == 0) the last rank does.
  // Accumulation kernel per SET of this rank (profiles/r2_notes.md §8).  Stand-alone, G2 sets win with the batched-affine
  // tree from 2^18 terms (3.82 vs 4.21 ms; 5.20 vs 6.54 at 2^19) and G1 sets from ~2^19.5 (2.53 vs 2.45 ms at 2^19, 3.96 vs 4.12
  // at 2^20); inside a sharded proof a 318 k-term G2 shard was still faster on the XYZZ kernel (4.81 vs 5.69 ms per rank), a
  // 524 k-term one on the affine tree — hence 4e5 / 7e5.  A single-GPU key always takes the affine tree: its four MSMs cover
  // each other's gaps.
  auto set_terms = [&](int k) { return (double)(pk->hi[k] - pk->lo[k]); };
  auto affine_for = [&](int k) { return world == 1 || set_terms(k) >= (double)(k == 2 ? g_aff_min_g2 : g_aff_min_g1); };
  // sets that consume identical scalar slices share one digit sort
  for (int k = 1; k < 3; k++)
    for (int j = 0; j < k; j++)
      if (pk->sort_src[k] == k && pk->lo[k] == pk->lo[j] && pk->hi[k] == pk->hi[j] && pk->tail[k] == pk->tail[j] && affine_for(k) == affine_for(j) &&
          (pk->lo[k] < pk->hi[k] || pk->tail[k]))
        pk->sort_src[k] = pk->sort_src[j];
  static const uint64_t inf1[12] = {0}, inf2[24] = {0};
  auto has = [&](int k) { return pk->lo[k] < pk->hi[k] || pk->tail[k]; };
  if (has(0)) {
    PointCat cat(12);
    cat.add(at + 12 * pk->lo[0], pk->hi[0] - pk->lo[0]);
    if (pk->tail[0]) { cat.add(alpha1, 1); cat.add(delta1, 1); cat.add(inf1, 1); }
    if ((rc = bases_create<Fq>(cat.v.data(), cat.count(), c, 1, pk->g[0], affine_for(0)))) return rc;
  }
  if (has(1)) {
    PointCat cat(12);
    cat.add(b1 + 12 * pk->lo[1], pk->hi[1] - pk->lo[1]);
    if (pk->tail[1]) { cat.add(beta1, 1); cat.add(inf1, 1); cat.add(delta1, 1); }
    if ((rc = bases_create<Fq>(cat.v.data(), cat.count(), c, 1, pk->g[1], affine_for(1)))) return rc;
  }
  if (has(2)) {
    PointCat cat(24);
    cat.add(b2 + 24 * pk->lo[2], pk->hi[2] - pk->lo[2]);
    if (pk->tail[2]) { cat.add(beta2, 1); cat.add(inf2, 1); cat.add(delta2, 1); }
    if ((rc = bases_create<Fq2>(cat.v.data(), cat.count(), c, 2, pk->g[2], affine_for(2)))) return rc;
  }
  if (has(3)) {
    PointCat cat(12);
    size_t lo3 = pk->lo[3], hi3 = pk->hi[3];
    size_t c_lo = lo3 < n_c_full ? lo3 : n_c_full, c_hi = hi3 < n_c_full ? hi3 : n_c_full;
    size_t p_lo = lo3 > n_c_full ? lo3 - n_c_full : 0, p_hi = hi3 > n_c_full ? hi3 - n_c_full : 0;
    cat.add(bacdelta + 12 * (l1 + c_lo), c_hi - c_lo);
    cat.add(ptd + 12 * p_lo, p_hi - p_lo);
    if (pk->tail[3]) cat.add(delta1, 1);
    if ((rc = bases_create<Fq>(cat.v.data(), cat.count(), c, 1, pk->g[3], affine_for(3)))) return rc;
  }
  CU(pk->s3.alloc((m + n_ptd + 4) * sizeof(Fr)));
  CU(pk->s4.alloc((m + 4) * sizeof(Fr)));
  if (world > 1) CU(pk->h_full.alloc((m + n_ptd + 4) * sizeof(Fr)));
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

// ---- blinding products s*A and r*B1 (groth16.go:272-273): GLV half-warps (glv.cuh) -----------------------------
// One warp, all 32 lanes: lanes 0..15 -> prod[0] = s*A, lanes 16..31 -> prod[1] = r*B1 (res layout as in k_groth16_finalize).
__global__ void k_groth16_products(const uint8_t* res, GlvScalars g, XYZZ<Fq>* prod) {
  const uint32_t t = threadIdx.x & 31u;
  const uint32_t grp = t >> 4;
  Jacobian<Fq> p = xyzz_to_jacobian(*reinterpret_cast<const XYZZ<Fq>*>(res + (grp ? 256 : 0)));
  Jacobian<Fq> r = glv_mul_halfwarp(p, g.k[grp], g.neg[grp], t);
  if ((t & 15u) == 0) prod[grp] = jacobian_to_xyzz(r);
}
__global__ void k_groth16_combine(const uint8_t* res, const XYZZ<Fq>* prod, Fq* out_a, Fq* out_c, Fq2* out_b) {
  uint32_t t = threadIdx.x;
  if (t == 0) {
    XYZZ<Fq> c = *reinterpret_cast<const XYZZ<Fq>*>(res + 768);
    xyzz_add(c, prod[0]);
    xyzz_add(c, prod[1]);
    store_jacobian_std(c, out_c);
  }
  if (t == 32) store_jacobian_std(*reinterpret_cast<const XYZZ<Fq>*>(res), out_a);
  if (t == 64) store_jacobian_std(*reinterpret_cast<const XYZZ<Fq2>*>(res + 512), out_b);
}


#define EV_REC(e, stream) CU(cudaEventRecord((e), (stream)))
#define EV_WAIT(stream, e) CU(cudaStreamWaitEvent((stream), (e), 0))

// Enqueue one Groth16 proof on `st` from DEVICE-resident witness / px (standard
// form).  d_out receives PiA (3 Fq) | PiC (3 Fq) | PiB (3 Fq2) in standard form
// (or, for a sharded key, the 1 KB partial record).  No host synchronisation.
int groth16_enqueue(ProvingKey* pk, const Fr* d_w, size_t nw, const Fr* d_px, size_t npx, const uint64_t* r,
                    const uint64_t* s, Fq* d_out, cudaStream_t st, int px_mont = 0, const Fr* d_h_std = nullptr) {
  // d_h_std != nullptr: the quotient h (npx - len(Z) + 1 coefficients, standard form, device) is supplied by the caller
  // (witness path: qap_h_enqueue on side stream 3) and the division is skipped; d_px is then unused.
  size_t m = pk->m;
  if (nw != m) return fail(B200_EINVAL, "groth16_prove: witness length %zu != NVars %zu", nw, m);
  if (npx < pk->Z.nb) return fail(B200_EINVAL, "groth16_prove: len(px) < len(Z)");
  size_t nq = npx - pk->Z.nb + 1;
  // groth16.go:269-270 indexes PowersTauDelta[i] for i < len(hx): out of range panics in the reference
  if (nq > pk->n_h_bases)
    return fail(B200_EINVAL, "groth16_prove: len(hx)=%zu exceeds len(PowersTauDelta)=%zu", nq, pk->n_h_bases);
  Fr fr_r = fr_load_std(r), fr_s = fr_load_std(s);
  if (fr_r.geq_modulus() || fr_s.geq_modulus()) return fail(B200_ERANGE, "groth16_prove: r or s >= field order");
  Fr neg_rs = (fr_r.to_mont() * fr_s.to_mont()).neg().from_mont();  // -(r*s) mod r  (groth16.go:274)
  Fr one = Fr::zero();
  one.l[0] = 1;
  cudaStream_t s1 = pk->side()[0], s2 = pk->side()[1], s3 = pk->side()[2];
  if (g_serial) s1 = s2 = s3 = st;  // measurement mode (b200_profile bit 1): no overlap, exclusive kernel timings
  cudaEvent_t e_in = pk->ev[0], e_w = pk->ev[1], e_a = pk->ev[2], e_b1 = pk->ev[3], e_b2 = pk->ev[4],
              e_ch = pk->ev[5], e_prod = pk->ev[6];
  const size_t l1 = pk->npublic + 1, n_c_full = pk->n_c_full;
  uint8_t* res = pk->res.as<uint8_t>();
  int rc;
  Fr small[6] = {one, fr_r, fr_s, neg_rs, fr_r, fr_s};
  CU(cudaMemcpyAsync(pk->rs.p, small, sizeof small, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(res, 0, kPartialBytes, st));      // sets this rank does not hold contribute infinity
  // scalar vectors: set k in {A, B1, B2} uses w[lo_k, hi_k) (+ [1, r, s] when it owns the blinding points)
  Fr* sv[3] = {pk->s1.as<Fr>(), pk->s2.as<Fr>(), pk->s4.as<Fr>()};
  size_t nterm[4] = {0, 0, 0, 0};
  for (int k = 0; k < 3; k++) {
    if (!pk->g[k]) continue;
    if (pk->sort_src[k] != k) { nterm[k] = nterm[pk->sort_src[k]]; continue; }   // shares the sort of an identical slice
    size_t cnt = pk->hi[k] - pk->lo[k];
    if (cnt) CU(cudaMemcpyAsync(sv[k], d_w + pk->lo[k], cnt * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    if (pk->tail[k]) CU(cudaMemcpyAsync(sv[k] + cnt, pk->rs.as<Fr>(), 3 * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    nterm[k] = cnt + (pk->tail[k] ? 3 : 0);
  }
  // C || PTD set
  Fr* sCH = pk->s3.as<Fr>();
  size_t lo3 = pk->lo[3], hi3 = pk->hi[3];
  size_t c_lo = lo3 < n_c_full ? lo3 : n_c_full, c_hi = hi3 < n_c_full ? hi3 : n_c_full;
  size_t p_lo = lo3 > n_c_full ? lo3 - n_c_full : 0, p_hi = hi3 > n_c_full ? hi3 - n_c_full : 0;
  size_t n_c = c_hi - c_lo, n_p = p_hi - p_lo;
  if (pk->g[3]) {
    if (n_c) CU(cudaMemcpyAsync(sCH, d_w + l1 + c_lo, n_c * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    if (pk->tail[3]) CU(cudaMemcpyAsync(sCH + n_c + n_p, pk->rs.as<Fr>() + 3, sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    nterm[3] = n_c + n_p + (pk->tail[3] ? 1 : 0);
  }
  EV_REC(e_in, st);
  // --- side stream 3: hx = px / Z (groth16.go:266) when this rank holds PowersTauDelta points, then C||PTD.
  // Single GPU: h is written straight into the scalar vector.  Sharded: the rank(s) holding PTD indices repeat
  // the (cheap) division and keep h[p_lo, p_hi) — no inter-GPU traffic.
  EV_WAIT(s3, e_in);
  if (pk->g[3]) {
    if (n_p) {
      bool direct = p_lo == 0 && nq <= n_p;
      Fr* h_dst = direct ? sCH + n_c : pk->h_full.as<Fr>();
      if (!direct && !pk->h_full.p) CU(pk->h_full.alloc((pk->m + pk->n_h_bases + 4) * sizeof(Fr)));
      if (d_h_std) CU(cudaMemcpyAsync(h_dst, d_h_std, nq * sizeof(Fr), cudaMemcpyDeviceToDevice, s3));
      else CU(poly_div_device(pk->poly(), pk->Z, d_px, npx, px_mont, h_dst, nullptr, g_d_err, s3));
      size_t have = nq > p_lo ? (nq < p_hi ? nq - p_lo : n_p) : 0;   // valid h coefficients inside [p_lo, p_hi)
      if (!direct && have)
        CU(cudaMemcpyAsync(sCH + n_c, pk->h_full.as<Fr>() + p_lo, have * sizeof(Fr), cudaMemcpyDeviceToDevice, s3));
      if (n_p > have) CU(cudaMemsetAsync(sCH + n_c + have, 0, (n_p - have) * sizeof(Fr), s3));
    }
    if ((rc = msm_enqueue<Fq>(pk->g[3].get(), sCH, nterm[3], 0, reinterpret_cast<XYZZ<Fq>*>(res + 768), s3))) return rc;
  }
  EV_REC(e_ch, s3);
  // --- A, B1, B2.  Sets that consume identical scalar slices (all three on one GPU; A and B1 on a rank that
  // holds both whole) share ONE digit sort; the sort runs on the main stream, the bucket phases fan out.
  XYZZ<Fq>* rA = reinterpret_cast<XYZZ<Fq>*>(res);
  XYZZ<Fq>* rB1 = reinterpret_cast<XYZZ<Fq>*>(res + 256);
  XYZZ<Fq2>* rB2 = reinterpret_cast<XYZZ<Fq2>*>(res + 512);
  cudaStream_t sk[3] = {st, s1, s2};
  cudaEvent_t e_sorted[3] = {e_w, pk->ev[7], e_prod};   // e_prod is re-recorded later; safe: waits are enqueued first
  for (int k = 0; k < 3; k++) {
    if (!pk->g[k] || pk->sort_src[k] != k) continue;
    if ((rc = msm_sort(pk->g[k]->sort, pk->g[k]->sh, sv[k], nterm[k], 0, st))) return rc;
    EV_REC(e_sorted[k], st);
  }
  for (int k = 2; k >= 0; k--) {   // B2 first: it is the longest
    if (!pk->g[k]) continue;
    int src = pk->sort_src[k];
    if (sk[k] != st) EV_WAIT(sk[k], e_sorted[src]);
    size_t nt = nterm[src];
    if (k == 0) rc = msm_buckets<Fq>(pk->g[0].get(), pk->g[src]->sort, nt, rA, sk[0]);
    else if (k == 1) rc = msm_buckets<Fq>(pk->g[1].get(), pk->g[src]->sort, nt, rB1, sk[1]);
    else rc = msm_buckets<Fq2>(pk->g[2].get(), pk->g[src]->sort, nt, rB2, sk[2]);
    if (rc) return rc;
  }
  EV_REC(e_a, st);
  EV_REC(e_b1, s1);
  EV_REC(e_b2, s2);
  // s*A_part and r*B1_part (groth16.go:272-273; linear, so each rank does its own parts) on side stream 1
  // while B2 / C||PTD are still running
  XYZZ<Fq>* prod = reinterpret_cast<XYZZ<Fq>*>(res + 1024);
  EV_WAIT(s1, e_a);
  GlvScalars glv;
  if (glv_decompose(fr_s, glv.k[0], glv.neg[0]) || glv_decompose(fr_r, glv.k[1], glv.neg[1]))
    return fail(B200_EINVAL, "groth16_prove: GLV decomposition out of range");
  if (pk->g[0] || pk->g[1]) {
    k_groth16_products<<<1, 32, 0, s1>>>(res, glv, prod);
  } else {   // a rank that holds no part of A or B1: both products are the point at infinity (all-zero XYZZ) — skip the
    // 1 ms single-warp chain (it sat on the critical path of 5 of the 8 ranks of an 8-way proof)
    CU(cudaMemsetAsync(prod, 0, 2 * sizeof(XYZZ<Fq>), s1));
  }
  EV_REC(e_prod, s1);
  EV_WAIT(st, e_prod);
  EV_WAIT(st, e_b2);
  EV_WAIT(st, e_ch);
  if (pk->world == 1) {
    k_groth16_combine<<<1, 96, 0, st>>>(res, prod, d_out, d_out + 3, reinterpret_cast<Fq2*>(d_out + 6));
  } else {  // fold the products into this rank's C part; the 1 KB partial record goes to the all-gather
    k_groth16_fold_products<<<1, 32, 0, st>>>(res, prod);
    if (g_comm.active() && g_comm.world == pk->world) {
      if (g_comm.rank != pk->rank) return fail(B200_EINVAL, "groth16_prove: key shard %d on communicator rank %d", pk->rank, g_comm.rank);
      CU(pk->gather.ensure(kPartialBytes * (size_t)pk->world));
      ncclResult_t nrc = g_comm.api.AllGather(res, pk->gather.p, kPartialBytes, ncclUint8, g_comm.of(pk->ctx), st);
      if (nrc != ncclSuccess) return fail(B200_ECOMM, "ncclAllGather: %s", g_comm.api.GetErrorString(nrc));
      k_groth16_finalize<<<1, 96, 0, st>>>(pk->gather.as<uint8_t>(), pk->world, d_out, d_out + 3, reinterpret_cast<Fq2*>(d_out + 6));
      g_launches += 2;
    } else {
      CU(cudaMemcpyAsync(d_out, res, kPartialBytes, cudaMemcpyDeviceToDevice, st));
    }
  }
  g_launches += 3;
  CU(cudaGetLastError());
  return B200_OK;
}

// Sharded mode, after the all-gather: sum the `world` partial records and finish the proof.
int groth16_finalize_enqueue(ProvingKey* pk, const uint8_t* d_parts, int nparts, const uint64_t* r, const uint64_t* s,
                             Fq* d_out, cudaStream_t st) {
  (void)pk; (void)r; (void)s;  // r, s were consumed per rank (their products are already inside the C parts)
  k_groth16_finalize<<<1, 96, 0, st>>>(d_parts, nparts, d_out, d_out + 3, reinterpret_cast<Fq2*>(d_out + 6));
  g_launches += 1;
  CU(cudaGetLastError());
  return B200_OK;
}

int groth16_prove(b200_pk_t h, const uint64_t* w, size_t nw, const uint64_t* px, size_t npx, const uint64_t* r,
                  const uint64_t* s, uint64_t* pi_a, uint64_t* pi_b, uint64_t* pi_c, std::unique_lock<std::mutex>* lk = nullptr) {
  ProvingKey* pk = find_pk(h, 1);
  if (!pk) return fail(B200_EINVAL, "groth16_prove: bad proving-key handle");
  if (!w || !px || !r || !s || !pi_a || !pi_b || !pi_c) return fail(B200_EINVAL, "groth16_prove: null pointer");
  if (nw != pk->m) return fail(B200_EINVAL, "groth16_prove: witness length %zu != NVars %zu", nw, pk->m);
  const bool sharded = pk->world != 1;
  if (sharded && !(g_comm.active() && g_comm.world == pk->world))
    return fail(B200_EINVAL, "groth16_prove: sharded key (rank %d/%d) without a matching communicator: call b200_comm_init, "
                             "or use b200_groth16_prove_device + b200_groth16_finalize_device around your own all-gather",
                pk->rank, pk->world);
  if (npx < pk->Z.nb) return fail(B200_EINVAL, "groth16_prove: len(px) < len(Z)");
  cudaStream_t st = pk->main();   // context 1 keys: their own main stream (a second host thread's proof in flight)
  CU(pk->px.ensure(npx * sizeof(Fr)));
  CU(pk->w_stage.ensure(nw * sizeof(Fr)));
  bool need_px = true;
  if (!sharded) {
    CU(cudaMemcpyAsync(pk->w_stage.p, w, nw * sizeof(Fr), cudaMemcpyHostToDevice, st));
  } else {   // stage only what this rank reads: its witness index ranges, and px only if it holds PowersTauDelta points
    const size_t l1 = pk->npublic + 1, ncf = pk->n_c_full;
    size_t lo[4], hi[4];
    for (int k = 0; k < 3; k++) { lo[k] = pk->lo[k]; hi[k] = pk->hi[k]; }
    size_t c_lo = pk->lo[3] < ncf ? pk->lo[3] : ncf, c_hi = pk->hi[3] < ncf ? pk->hi[3] : ncf;
    lo[3] = l1 + c_lo;
    hi[3] = l1 + c_hi;
    need_px = pk->hi[3] > ncf;
    for (int k = 0; k < 4; k++) {
      if (lo[k] >= hi[k]) continue;
      bool dup = false;
      for (int j = 0; j < k; j++) dup = dup || (lo[j] <= lo[k] && hi[k] <= hi[j] && lo[j] < hi[j]);
      if (dup) continue;
      CU(cudaMemcpyAsync(pk->w_stage.as<Fr>() + lo[k], reinterpret_cast<const Fr*>(w) + lo[k], (hi[k] - lo[k]) * sizeof(Fr),
                         cudaMemcpyHostToDevice, st));
    }
  }
  // px is only needed by the division on side stream 3: copy it there so the transfer overlaps the sort of w
  // (same-stream order makes the division see it; the previous proof's use of pk->px finished before its combine)
  // (measurement mode runs the division on `st`, so the copy goes there too)
  // Only the TOP len(px) - len(Z) + 1 coefficients are staged: the quotient of the division depends on nothing else
  // (poly_div_device reads rev(px) mod x^nq; the remainder is never formed on the prove path) — at 2^20 constraints that
  // is 32 MB over PCIe instead of 64 MB.  They land at their own offset, so the device layout of px is unchanged.
  if (need_px) {
    const size_t px_off = pk->Z.nb - 1;
    CU(cudaMemcpyAsync(pk->px.as<Fr>() + px_off, reinterpret_cast<const Fr*>(px) + px_off, (npx - px_off) * sizeof(Fr),
                       cudaMemcpyHostToDevice, g_serial ? st : pk->side()[2]));
  }
  Fq* o = pk->out_std.as<Fq>();
  int rc = groth16_enqueue(pk, pk->w_stage.as<Fr>(), nw, pk->px.as<Fr>(), npx, r, s, o, st);
  if (rc) return rc;
  uint64_t* host_out = pk->h_out;
  CU(cudaMemcpyAsync(host_out, o, 48 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  rc = check_err_flag<Fr>("groth16_prove", st, lk, reinterpret_cast<int*>(host_out + 48));  // synchronises (the wait runs with the library mutex released)
  if (rc) return rc;
  memcpy(pi_a, host_out, 12 * 8);
  memcpy(pi_c, host_out + 12, 12 * 8);
  memcpy(pi_b, host_out + 24, 24 * 8);
  return B200_OK;
}

// Pinocchio result slots (XYZZ, 256 B each): 0 PiA, 1 PiAp, 2 PiBp, 3 PiC, 4 PiCp, 5 PiH, 6 PiKp, 7 PiB (G2).
// Base sets g[k]: 0 A, 1 Ap, 2 B (G2), 3 Bp, 4 C, 5 Cp, 6 Kp, 7 G1T.
constexpr int kPinSlot[8] = {0, 1, 7, 2, 3, 4, 6, 5};
constexpr size_t kPinRecordBytes = 8 * 256;

// Sharded mode (world > 1): the eight MSMs are independent objects, so whole MSMs are dealt to the ranks — greedy
// longest-first onto the least loaded rank by term count (a G2 term ~2.8 G1 terms) — and no MSM is split.  `owner[k]` is the
// same on every rank.  (Eight objects: more than 8 ranks leave the rest idle.)
void pinocchio_owners(size_t m, size_t l1, size_t n_g1t, int world, int owner[8]) {
  double cost[8] = {(double)(m - l1), (double)(m - l1), 2.8 * (double)m, (double)m, (double)m, (double)m, (double)m, (double)n_g1t};
  int order[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  std::stable_sort(order, order + 8, [&](int x, int y) { return cost[x] > cost[y]; });
  std::vector<double> load((size_t)world, 0.0);
  for (int q = 0; q < 8; q++) {
    int k = order[q], best = 0;
    for (int g = 1; g < world; g++)
      if (load[g] < load[best]) best = g;
    owner[k] = best;
    load[best] += cost[k];
  }
}

// sum the world records slot by slot (every slot is non-infinity on exactly one rank) -> standard-form outputs
__global__ void k_pinocchio_gather_finalize(const uint8_t* recs, int world, Fq* out_g1, Fq2* out_b) {
  uint32_t t = threadIdx.x;
  if (t < 7) {
    XYZZ<Fq> acc = XYZZ<Fq>::inf();
    for (int g = 0; g < world; g++) xyzz_add(acc, *reinterpret_cast<const XYZZ<Fq>*>(recs + kPinRecordBytes * g + 256 * t));
    store_jacobian_std(acc, out_g1 + 3 * t);
  }
  if (t == 32) {
    XYZZ<Fq2> acc = XYZZ<Fq2>::inf();
    for (int g = 0; g < world; g++) xyzz_add(acc, *reinterpret_cast<const XYZZ<Fq2>*>(recs + kPinRecordBytes * g + 256 * 7));
    store_jacobian_std(acc, out_b);
  }
}

int pinocchio_pk_load(const uint64_t* a, const uint64_t* ap, const uint64_t* b2, const uint64_t* bp,
                      const uint64_t* c, const uint64_t* cp, const uint64_t* kp, size_t m, const uint64_t* g1t,
                      size_t n_g1t, const uint64_t* z, size_t nz, size_t npublic, int wb, int rank, int world, b200_pk_t* out) {
  if (!a || !ap || !b2 || !bp || !c || !cp || !kp || !g1t || !out)
    return fail(B200_EINVAL, "pinocchio_pk_load: null pointer");
  if (m == 0 || npublic + 1 > m || n_g1t == 0) return fail(B200_EINVAL, "pinocchio_pk_load: bad sizes");
  if (world < 1 || rank < 0 || rank >= world) return fail(B200_EINVAL, "pinocchio_pk_load: bad shard %d/%d", rank, world);
  auto pk = std::make_unique<ProvingKey>();
  pk->kind = 2;
  pk->m = m;
  pk->npublic = npublic;
  pk->n_h_bases = n_g1t;
  int rc = pk_common_init(*pk, z, nz, m);
  if (rc) return rc;
  size_t l1 = npublic + 1;
  pk->rank = rank;
  pk->world = world;
  int owner[8];
  pinocchio_owners(m, l1, n_g1t, world, owner);
  auto mine = [&](int k) { return owner[k] == rank; };
  // snark.go:265-268: PiA, PiAp run over i in [NPublic+1, NVars)
  if (m > l1) {
    if (mine(0) && (rc = bases_create<Fq>(a + 12 * l1, m - l1, wb, 1, pk->g[0], true))) return rc;
    if (mine(1) && (rc = bases_create<Fq>(ap + 12 * l1, m - l1, wb, 1, pk->g[1], true))) return rc;
  }
  if (mine(2) && (rc = bases_create<Fq2>(b2, m, wb, 2, pk->g[2], true))) return rc;
  if (mine(3) && (rc = bases_create<Fq>(bp, m, wb, 1, pk->g[3], true))) return rc;
  if (mine(4) && (rc = bases_create<Fq>(c, m, wb, 1, pk->g[4], true))) return rc;
  if (mine(5) && (rc = bases_create<Fq>(cp, m, wb, 1, pk->g[5], true))) return rc;
  if (mine(6) && (rc = bases_create<Fq>(kp, m, wb, 1, pk->g[6], true))) return rc;
  if (mine(7) && (rc = bases_create<Fq>(g1t, n_g1t, wb, 1, pk->g[7], true))) return rc;
  CU(pk->s3.alloc((n_g1t + 4) * sizeof(Fr)));
  uint64_t h = g_next_pk++;
  g_pks[h] = std::move(pk);
  *out = h;
  return B200_OK;
}

// d_rec_out != nullptr: leave this rank's 2 KB record of XYZZ results there (device) and return without synchronising —
// the form the one-GPU tests use to emulate N ranks.
int pinocchio_prove(b200_pk_t h, const uint64_t* w, size_t nw, const uint64_t* px, size_t npx, uint64_t* out_g1,
                    uint64_t* pi_b, uint8_t* d_rec_out = nullptr) {
  ProvingKey* pk = find_pk(h, 2);
  if (!pk) return fail(B200_EINVAL, "pinocchio_prove: bad proving-key handle");
  if (!w || !px || (!d_rec_out && (!out_g1 || !pi_b))) return fail(B200_EINVAL, "pinocchio_prove: null pointer");
  size_t m = pk->m, l1 = pk->npublic + 1;
  if (nw != m) return fail(B200_EINVAL, "pinocchio_prove: witness length %zu != NVars %zu", nw, m);
  if (npx < pk->Z.nb) return fail(B200_EINVAL, "pinocchio_prove: len(px) < len(Z)");
  size_t nq = npx - pk->Z.nb + 1;
  if (nq > pk->n_h_bases) return fail(B200_EINVAL, "pinocchio_prove: len(hx)=%zu exceeds len(G1T)=%zu", nq, pk->n_h_bases);
  const bool sharded = pk->world > 1;
  const bool gather = sharded && !d_rec_out;
  if (gather && !(g_comm.active() && g_comm.world == pk->world && g_comm.rank == pk->rank))
    return fail(B200_EINVAL, "pinocchio_prove: sharded key (rank %d/%d) without a matching communicator (b200_comm_init)",
                pk->rank, pk->world);
  cudaStream_t st = g_stream, side[3] = {pk->side()[0], pk->side()[1], pk->side()[2]};
  cudaEvent_t e_in = pk->ev[0], e_sorted[2] = {pk->ev[1], pk->ev[2]}, e_done[3] = {pk->ev[3], pk->ev[4], pk->ev[5]};
  Fr* dw = pk->s1.as<Fr>();
  Fr* dh = pk->s3.as<Fr>();
  uint8_t* res = pk->res.as<uint8_t>();
  CU(cudaMemcpyAsync(dw, w, m * sizeof(Fr), cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(res, 0, kPinRecordBytes, st));        // slots this rank does not own stay at infinity
  if (pk->g[7]) {
    CU(pk->px.ensure(npx * sizeof(Fr)));
    const size_t px_off = pk->Z.nb - 1;   // the top nq coefficients are all the division reads (see groth16_prove)
    CU(cudaMemcpyAsync(pk->px.as<Fr>() + px_off, reinterpret_cast<const Fr*>(px) + px_off, nq * sizeof(Fr), cudaMemcpyHostToDevice, st));
  }
  EV_REC(e_in, st);
  int rc;
  int next_side = 0;
  auto slot_g1 = [&](int k) { return reinterpret_cast<XYZZ<Fq>*>(res + 256 * kPinSlot[k]); };
  // side stream: hx = px / Z (snark.go:280) and PiH (snark.go:284-286)
  bool used[3] = {false, false, false};
  if (pk->g[7]) {
    cudaStream_t s3 = side[2];
    used[2] = true;
    EV_WAIT(s3, e_in);
    CU(poly_div_device(pk->poly(), pk->Z, pk->px.as<Fr>(), npx, 0, dh, nullptr, g_d_err, s3));
    if ((rc = msm_enqueue<Fq>(pk->g[7].get(), dh, nq, 0, slot_g1(7), s3))) return rc;
  }
  // two scalar vectors: w[l+1..m) feeds PiA, PiAp (snark.go:265-268); w feeds PiB, PiBp, PiC, PiCp, PiKp (:270-278).
  // Per vector ONE digit sort (held by the first owned set of the group), the bucket phases fan out over the streams.
  const int groups[2][5] = {{0, 1, -1, -1, -1}, {2, 3, 4, 5, 6}};
  int holder[2] = {-1, -1};
  for (int gi = 0; gi < 2; gi++) {   // both sorts first (main stream), then the bucket phases fan out
    for (int q = 0; q < 5; q++)
      if (groups[gi][q] >= 0 && pk->g[groups[gi][q]]) { holder[gi] = groups[gi][q]; break; }
    if (holder[gi] < 0) continue;
    const Fr* sc = gi == 0 ? dw + l1 : dw;
    const size_t cnt = gi == 0 ? m - l1 : m;
    if ((rc = msm_sort(pk->g[holder[gi]]->sort, pk->g[holder[gi]]->sh, sc, cnt, 0, st))) return rc;
    EV_REC(e_sorted[gi], st);
  }
  for (int gi = 1; gi >= 0; gi--) {   // the larger group first
    if (holder[gi] < 0) continue;
    const size_t cnt = gi == 0 ? m - l1 : m;
    SortScratch& ss = pk->g[holder[gi]]->sort;
    for (int q = 0; q < 5; q++) {
      int k = groups[gi][q];
      if (k < 0 || !pk->g[k]) continue;
      cudaStream_t sk = st;
      if (!(gi == 1 && k == holder[1])) {   // one bucket phase stays on the main stream, the others rotate over the side streams
        int si = next_side++ % 3;
        sk = side[si];
        used[si] = true;
        EV_WAIT(sk, e_sorted[gi]);
      }
      if (k == 2) rc = msm_buckets<Fq2>(pk->g[2].get(), ss, cnt, reinterpret_cast<XYZZ<Fq2>*>(res + 256 * 7), sk);
      else rc = msm_buckets<Fq>(pk->g[k].get(), ss, cnt, slot_g1(k), sk);
      if (rc) return rc;
    }
  }
  for (int si = 0; si < 3; si++)
    if (used[si]) {
      EV_REC(e_done[si], side[si]);
      EV_WAIT(st, e_done[si]);
    }
  if (d_rec_out) {
    CU(cudaMemcpyAsync(d_rec_out, res, kPinRecordBytes, cudaMemcpyDeviceToDevice, st));
    CU(cudaGetLastError());
    return B200_OK;
  }
  Fq* o = pk->out_std.as<Fq>();
  if (gather) {
    CU(pk->gather.ensure(kPinRecordBytes * (size_t)pk->world));
    ncclResult_t nrc = g_comm.api.AllGather(res, pk->gather.p, kPinRecordBytes, ncclUint8, g_comm.of(pk->ctx), st);
    if (nrc != ncclSuccess) return fail(B200_ECOMM, "ncclAllGather: %s", g_comm.api.GetErrorString(nrc));
    k_pinocchio_gather_finalize<<<1, 64, 0, st>>>(pk->gather.as<uint8_t>(), pk->world, o, reinterpret_cast<Fq2*>(o + 21));
  } else {
    k_pinocchio_finalize<<<1, 64, 0, st>>>(res, o, reinterpret_cast<Fq2*>(o + 21));
  }
  CU(cudaGetLastError());
  uint64_t host_out[(21 + 6) * 4];
  CU(cudaMemcpyAsync(host_out, o, sizeof host_out, cudaMemcpyDeviceToHost, st));
  rc = check_err_flag<Fr>("pinocchio_prove");
  if (rc) return rc;
  memcpy(out_g1, host_out, 84 * 8);
  memcpy(pi_b, host_out + 84, 24 * 8);
  return B200_OK;
}

// one-GPU emulation of the collective: sum `world` gathered records (device) -> the proof (host)
int pinocchio_finalize_records(const uint8_t* d_recs, int world, uint64_t* out_g1, uint64_t* pi_b) {
  if (!d_recs || world < 1 || !out_g1 || !pi_b) return fail(B200_EINVAL, "pinocchio_finalize: bad arguments");
  DevBuf o;
  CU(o.alloc(64 * sizeof(Fq)));
  k_pinocchio_gather_finalize<<<1, 64, 0, g_stream>>>(d_recs, world, o.as<Fq>(), reinterpret_cast<Fq2*>(o.as<Fq>() + 21));
  CU(cudaGetLastError());
  uint64_t host_out[(21 + 6) * 4];
  CU(cudaMemcpyAsync(host_out, o.p, sizeof host_out, cudaMemcpyDeviceToHost, g_stream));
  int rc = check_err_flag<Fr>("pinocchio_finalize");
  if (rc) return rc;
  memcpy(out_g1, host_out, 84 * 8);
  memcpy(pi_b, host_out + 84, 24 * 8);
  return B200_OK;
}
