/* CPU oracle, C restatement (TEST INFRASTRUCTURE ONLY — never linked into or
 * called by the product library; used by tests/ and by bench.py's cpu_baseline /
 * --impl reference leg).
 *
 * Restates the reference's prove-path hot loops in their own operation order:
 *   fields/fq.go:32-98      Add/Sub/Mul/Square  (here Montgomery 4x64, canonical results)
 *   fields/fq2.go:37-133    Fq2 Add/Sub/Mul/Square (non-residue -1)
 *   bn128/g1.go:32-155      G1 Add (add-2007-bl), Double (dbl-2009-l), MulScalar (MSB-first)
 *   bn128/g2.go:32-181      the same over Fq2
 *   groth16/groth16.go:243-250   acc = Add(acc, MulScalar(P_i, w_i))   in index order
 * Pinned against oracle/ref_py.py (itself pinned bit-exactly against the Go
 * binary) by tests/test_oracle_c.py, including Jacobian X,Y,Z.
 *
 * All API values are standard-form little-endian 4 x uint64 limbs.
 */
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;
typedef struct { fe c0, c1; } fe2;

static const fe Q = {{0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};
static const uint64_t QINV = 0x87d20782e4866389ULL;   /* -q^-1 mod 2^64 */
static const fe R2 = {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}};
static const fe ONE_STD = {{1, 0, 0, 0}};

static int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static int fe_geq_q(const fe* a) {
  for (int i = 3; i >= 0; i--) {
    if (a->l[i] > Q.l[i]) return 1;
    if (a->l[i] < Q.l[i]) return 0;
  }
  return 1;
}
static void fe_sub_q(fe* a) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a->l[i] - Q.l[i] - (uint64_t)b;
    a->l[i] = (uint64_t)t;
    b = (t >> 64) & 1;
  }
}
static void fe_add(fe* r, const fe* a, const fe* b) {          /* fq.go:32-35 */
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a->l[i] + b->l[i];
    r->l[i] = (uint64_t)c;
    c >>= 64;
  }
  if (fe_geq_q(r)) fe_sub_q(r);
}
static void fe_sub(fe* r, const fe* a, const fe* b) {          /* fq.go:44-47 */
  u128 br = 0;
  fe t;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->l[i] - b->l[i] - (uint64_t)br;
    t.l[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
  if (br) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)t.l[i] + Q.l[i];
      t.l[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  *r = t;
}
static void fe_mul(fe* r, const fe* a, const fe* b) {          /* fq.go:56-59 (Montgomery CIOS) */
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a->l[j] * b->l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * QINV;
    c = (u128)m * Q.l[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * Q.l[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  fe o = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || fe_geq_q(&o)) fe_sub_q(&o);
  *r = o;
}
static void fe_sqr(fe* r, const fe* a) { fe_mul(r, a, a); }
static void fe_to_mont(fe* r, const fe* a) { fe_mul(r, a, &R2); }
static void fe_from_mont(fe* r, const fe* a) { fe_mul(r, a, &ONE_STD); }

/* ---- Fq2 (fq2.go), u^2 = -1 */
static void f2_add(fe2* r, const fe2* a, const fe2* b) { fe_add(&r->c0, &a->c0, &b->c0); fe_add(&r->c1, &a->c1, &b->c1); }
static void f2_sub(fe2* r, const fe2* a, const fe2* b) { fe_sub(&r->c0, &a->c0, &b->c0); fe_sub(&r->c1, &a->c1, &b->c1); }
static void f2_mul(fe2* r, const fe2* a, const fe2* b) {       /* fq2.go:63-76 */
  fe v0, v1, s0, s1, t;
  fe_mul(&v0, &a->c0, &b->c0);
  fe_mul(&v1, &a->c1, &b->c1);
  fe_add(&s0, &a->c0, &a->c1);
  fe_add(&s1, &b->c0, &b->c1);
  fe_mul(&t, &s0, &s1);
  fe_sub(&r->c0, &v0, &v1);
  fe_add(&s0, &v0, &v1);
  fe_sub(&r->c1, &t, &s0);
}
static void f2_sqr(fe2* r, const fe2* a) { fe2 t = *a; f2_mul(r, &t, &t); }
static int f2_is_zero(const fe2* a) { return fe_is_zero(&a->c0) && fe_is_zero(&a->c1); }

/* ---- generic Jacobian formulas via macros instantiated for fe and fe2 */
#define DEFINE_GROUP(NAME, F, ADD, SUB, MUL, SQR, ISZ)                                                   \
  typedef struct { F X, Y, Z; } NAME##_pt;                                                               \
  static void NAME##_add(NAME##_pt* o, const NAME##_pt* p1, const NAME##_pt* p2) { /* g1.go:32-89 */     \
    if (ISZ(&p1->Z)) { *o = *p2; return; }                                                               \
    if (ISZ(&p2->Z)) { *o = *p1; return; }                                                               \
    F z1z1, z2z2, u1, u2, t0, s1, t1, s2, h, t2, i, j, t3, r, v, t4, t5, t6, x3, t7, t8, t9, t10, y3,    \
        t11, t12, t13, t14, z3;                                                                          \
    SQR(&z1z1, &p1->Z); SQR(&z2z2, &p2->Z);                                                              \
    MUL(&u1, &p1->X, &z2z2); MUL(&u2, &p2->X, &z1z1);                                                    \
    MUL(&t0, &p2->Z, &z2z2); MUL(&s1, &p1->Y, &t0);                                                      \
    MUL(&t1, &p1->Z, &z1z1); MUL(&s2, &p2->Y, &t1);                                                      \
    SUB(&h, &u2, &u1); ADD(&t2, &h, &h); SQR(&i, &t2); MUL(&j, &h, &i);                                  \
    SUB(&t3, &s2, &s1); ADD(&r, &t3, &t3); MUL(&v, &u1, &i);                                             \
    SQR(&t4, &r); ADD(&t5, &v, &v); SUB(&t6, &t4, &j); SUB(&x3, &t6, &t5);                               \
    SUB(&t7, &v, &x3); MUL(&t8, &s1, &j); ADD(&t9, &t8, &t8); MUL(&t10, &r, &t7); SUB(&y3, &t10, &t9);   \
    ADD(&t11, &p1->Z, &p2->Z); SQR(&t12, &t11); SUB(&t13, &t12, &z1z1); SUB(&t14, &t13, &z2z2);          \
    MUL(&z3, &t14, &h);                                                                                  \
    o->X = x3; o->Y = y3; o->Z = z3;                                                                     \
  }                                                                                                      \
  static void NAME##_dbl(NAME##_pt* o, const NAME##_pt* p) { /* g1.go:101-138 */                         \
    if (ISZ(&p->Z)) { *o = *p; return; }                                                                 \
    F a, b, c, t0, t1, t2, t3, d, e, f, t4, x3, t5, c2, c4, t6, t7, y3, t8, z3;                           \
    SQR(&a, &p->X); SQR(&b, &p->Y); SQR(&c, &b);                                                         \
    ADD(&t0, &p->X, &b); SQR(&t1, &t0); SUB(&t2, &t1, &a); SUB(&t3, &t2, &c);                            \
    ADD(&d, &t3, &t3); ADD(&e, &a, &a); ADD(&e, &e, &a); SQR(&f, &e);                                    \
    ADD(&t4, &d, &d); SUB(&x3, &f, &t4); SUB(&t5, &d, &x3);                                              \
    ADD(&c2, &c, &c); ADD(&c4, &c2, &c2); ADD(&t6, &c4, &c4);                                            \
    MUL(&t7, &e, &t5); SUB(&y3, &t7, &t6); MUL(&t8, &p->Y, &p->Z); ADD(&z3, &t8, &t8);                   \
    o->X = x3; o->Y = y3; o->Z = z3;                                                                     \
  }                                                                                                      \
  static void NAME##_mul_scalar(NAME##_pt* o, const NAME##_pt* p, const uint64_t* e) { /* g1.go:140-155 */ \
    NAME##_pt q;                                                                                         \
    memset(&q, 0, sizeof q);                                                                             \
    int top = 255;                                                                                       \
    while (top >= 0 && !((e[top >> 6] >> (top & 63)) & 1)) top--;                                        \
    for (int i = top; i >= 0; i--) {                                                                     \
      NAME##_pt t;                                                                                       \
      NAME##_dbl(&t, &q);                                                                                \
      q = t;                                                                                             \
      if ((e[i >> 6] >> (i & 63)) & 1) { NAME##_add(&t, &q, p); q = t; }                                 \
    }                                                                                                    \
    *o = q;                                                                                              \
  }

DEFINE_GROUP(g1, fe, fe_add, fe_sub, fe_mul, fe_sqr, fe_is_zero)
DEFINE_GROUP(g2, fe2, f2_add, f2_sub, f2_mul, f2_sqr, f2_is_zero)

static void g1_load(g1_pt* p, const uint64_t* s) {
  fe t;
  for (int k = 0; k < 3; k++) { memcpy(&t, s + 4 * k, 32); fe_to_mont(k == 0 ? &p->X : k == 1 ? &p->Y : &p->Z, &t); }
}
static void g1_store(uint64_t* s, const g1_pt* p) {
  fe t;
  fe_from_mont(&t, &p->X); memcpy(s, &t, 32);
  fe_from_mont(&t, &p->Y); memcpy(s + 4, &t, 32);
  fe_from_mont(&t, &p->Z); memcpy(s + 8, &t, 32);
}
static void g2_load(g2_pt* p, const uint64_t* s) {
  fe t; fe* dst[6] = {&p->X.c0, &p->X.c1, &p->Y.c0, &p->Y.c1, &p->Z.c0, &p->Z.c1};
  for (int k = 0; k < 6; k++) { memcpy(&t, s + 4 * k, 32); fe_to_mont(dst[k], &t); }
}
static void g2_store(uint64_t* s, const g2_pt* p) {
  fe t; const fe* src[6] = {&p->X.c0, &p->X.c1, &p->Y.c0, &p->Y.c1, &p->Z.c0, &p->Z.c1};
  for (int k = 0; k < 6; k++) { fe_from_mont(&t, src[k]); memcpy(s + 4 * k, &t, 32); }
}

/* out = MulScalar(p, e)   (X,Y,Z-exact) */
void oc_g1_mul_scalar(const uint64_t* p, const uint64_t* e, uint64_t* out) {
  g1_pt a, r; g1_load(&a, p); g1_mul_scalar(&r, &a, e); g1_store(out, &r);
}
void oc_g2_mul_scalar(const uint64_t* p, const uint64_t* e, uint64_t* out) {
  g2_pt a, r; g2_load(&a, p); g2_mul_scalar(&r, &a, e); g2_store(out, &r);
}

/* The reference hot loop (groth16.go:243-250): acc = Add(acc, MulScalar(P_i, s_i)), i ascending.
 * threads <= 1: exactly the reference order (X,Y,Z-exact).  threads > 1: the index range is cut into
 * contiguous chunks run in parallel and the partial sums are added in chunk order (same group element;
 * used only to give the CPU baseline every host core, which the single-goroutine reference cannot use). */
#define DEFINE_LOOP(NAME, WORDS)                                                                        \
  void oc_##NAME##_msm_loop(const uint64_t* pts, const uint64_t* sc, long n, int threads, uint64_t* out) { \
    if (threads < 1) threads = 1;                                                                       \
    if (threads > 256) threads = 256;                                                                   \
    NAME##_pt part[256];                                                                                \
    memset(part, 0, sizeof(NAME##_pt) * threads);                                                       \
    _Pragma("omp parallel for num_threads(threads) schedule(static, 1)")                                \
    for (int t = 0; t < threads; t++) {                                                                 \
      long lo = n * t / threads, hi = n * (t + 1) / threads;                                            \
      NAME##_pt acc; memset(&acc, 0, sizeof acc);                                                       \
      for (long i = lo; i < hi; i++) {                                                                  \
        NAME##_pt p, m, s;                                                                              \
        NAME##_load(&p, pts + (size_t)WORDS * i);                                                       \
        NAME##_mul_scalar(&m, &p, sc + 4 * i);                                                          \
        NAME##_add(&s, &acc, &m);                                                                       \
        acc = s;                                                                                        \
      }                                                                                                 \
      part[t] = acc;                                                                                    \
    }                                                                                                   \
    NAME##_pt acc = part[0];                                                                            \
    for (int t = 1; t < threads; t++) { NAME##_pt s; NAME##_add(&s, &acc, &part[t]); acc = s; }         \
    NAME##_store(out, &acc);                                                                            \
  }
DEFINE_LOOP(g1, 12)
DEFINE_LOOP(g2, 24)

/* out[i] = MulScalar(p, sc[i]) for one base point p (threads in parallel): mints sample CRS points k_i*G for the CPU
 * baseline without touching the GPU library (groth16.go:139-219: every CRS entry is G.MulScalar(g, k)). */
#define DEFINE_MINT(NAME, WORDS)                                                                         \
  void oc_##NAME##_mul_batch_bcast(const uint64_t* p, const uint64_t* sc, long n, int threads, uint64_t* out) { \
    if (threads < 1) threads = 1;                                                                        \
    NAME##_pt base; NAME##_load(&base, p);                                                               \
    _Pragma("omp parallel for num_threads(threads) schedule(dynamic, 16)")                               \
    for (long i = 0; i < n; i++) {                                                                       \
      NAME##_pt m; NAME##_mul_scalar(&m, &base, sc + 4 * i); NAME##_store(out + (size_t)WORDS * i, &m);  \
    }                                                                                                    \
  }
DEFINE_MINT(g1, 12)
DEFINE_MINT(g2, 24)

int oc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

