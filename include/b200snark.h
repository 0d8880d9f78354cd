/* libb200snark — C ABI of the B200-native Groth16 / Pinocchio prove path.
 *
 * The reference (arnaucube/go-snark-study) has no FFI: its boundary is the
 * exported Go package API.  Each entry point below names the Go function(s)
 * whose body a cgo shim replaces (file:line in the upstream tree); the shim
 * itself is shown in INTEGRATION.md.  Conventions:
 *
 *   - field elements / scalars: 4 little-endian uint64 limbs, STANDARD form
 *     (not Montgomery), canonical (< modulus) — i.e. big.Int.Bits() on amd64,
 *     zero-padded.  Scalars must be reduced mod r (the shim sends |e| mod r:
 *     the reference's MulScalar consumes |e|, fields/fq.go:138-140).
 *   - G1 point  = [3]*big.Int  Jacobian (X, Y, Z)            -> 12 uint64
 *     G2 point  = [3][2]*big.Int Jacobian (X.c0, X.c1, Y.., Z..) -> 24 uint64
 *     Z == 0 (e.g. the (0,0,0) entries of Pk.BACDelta, groth16.go:177-180) is
 *     the point at infinity.  Outputs are normalised: (x, y, 1), infinity as
 *     all-zero — valid inputs to every reference function.
 *   - every function returns 0 (B200_OK) or a negative B200_E* code; no
 *     exceptions or panics cross the ABI; b200_last_error() gives the text.
 *   - all pointers are HOST pointers unless the name ends in _device.
 *   - thread-safety: calls are serialised per process by an internal mutex
 *     (cgo may migrate goroutines between OS threads).
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     B200_ENODEVICE.
 */
#ifndef B200SNARK_H
#define B200SNARK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ENODEVICE (-1) /* no CUDA device / init failed                         */
#define B200_ECUDA (-2)     /* CUDA runtime error (see b200_last_error)              */
#define B200_EINVAL (-3)    /* bad argument (null pointer, size mismatch, handle)    */
#define B200_ERANGE (-4)    /* a field element >= modulus or a scalar >= r           */
#define B200_EDIVZERO (-5)  /* polynomial division by a zero leading coefficient     */
#define B200_ENOMEM (-6)

typedef uint64_t b200_bases_t; /* device-resident, window-precomputed base-point set */
typedef uint64_t b200_pk_t;    /* device-resident proving key                        */

/* ---- context ------------------------------------------------------------ */
int b200_init(int device);        /* idempotent; selects the CUDA device of this process */
int b200_shutdown(void);
const char* b200_last_error(void);
int b200_version(void);

/* ---- base-point sets (the CRS arrays of groth16.Pk / snark.Pk) ---------- */
/* Upload n Jacobian points, normalise to affine Montgomery form on the device
 * and precompute 2^(c*w) * P for every window w (the CRS is static, so all
 * windows of one MSM share a single bucket set).
 * window_bits = 0 picks c from n.  Replaces nothing in the reference (its Pk
 * is consumed directly); this is the "load proving key once" step of H9.     */
int b200_g1_bases_load(const uint64_t* points_jac, size_t n, int window_bits, b200_bases_t* out);
int b200_g2_bases_load(const uint64_t* points_jac, size_t n, int window_bits, b200_bases_t* out);
int b200_bases_free(b200_bases_t h);
int b200_bases_info(b200_bases_t h, size_t* n, int* group, int* window_bits, int* n_windows);

/* ---- multi-scalar multiplication ---------------------------------------- */
/* out = sum_i scalars[i] * P_i over the first n bases of the set.
 * Replaces the hot loops  acc = G1.Add(acc, G1.MulScalar(P_i, w_i))  of
 * groth16/groth16.go:243-250,269-271 and snark.go:265-286 (G1: bn128/g1.go:32-155;
 * G2: bn128/g2.go:32-181).                                                    */
int b200_g1_msm(b200_bases_t h, const uint64_t* scalars, size_t n, uint64_t out_jac[12]);
int b200_g2_msm(b200_bases_t h, const uint64_t* scalars, size_t n, uint64_t out_jac[24]);

/* Same, scalars already resident in device memory (n * 32 bytes, standard form,
 * or Montgomery form when scalars_mont != 0), result left on the device as an
 * XYZZ record (4 field elements, Montgomery form; 128 B for G1, 256 B for G2)
 * on `stream` (a cudaStream_t, may be NULL).  No host synchronisation.  Used by
 * the prove pipeline, by bench.py's device-resident timing and by the
 * multi-GPU path (each rank's partial record is what the NCCL gather moves).  */
int b200_msm_device(b200_bases_t h, const void* d_scalars, size_t n, int scalars_mont,
                    void* d_out_xyzz, void* stream);
/* Sum `count` XYZZ partial records (device memory, e.g. the gathered per-rank
 * partials) and normalise: out_jac = (x, y, 1) in standard form on the host.  */
int b200_g1_sum_partials(const void* d_xyzz, size_t count, uint64_t out_jac[12], void* stream);
int b200_g2_sum_partials(const void* d_xyzz, size_t count, uint64_t out_jac[24], void* stream);

/* ---- batch scalar multiplication, reference operation order -------------- */
/* out[i] = MulScalar(points[i], scalars[i]) with the reference's MSB-first
 * double-and-add and its add-2007-bl / dbl-2009-l formulas, one thread per
 * term: X,Y,Z-exact drop-in for bn128.G1.MulScalar (bn128/g1.go:140-155) and
 * bn128.G2.MulScalar (bn128/g2.go:142-181).  Also mints CRSs on the GPU
 * (groth16/groth16.go:139-219: every Pk/Vk entry is G.MulScalar(g, k)).       */
int b200_g1_mul_batch(const uint64_t* points_jac, const uint64_t* scalars, size_t n, uint64_t* out_jac);
int b200_g2_mul_batch(const uint64_t* points_jac, const uint64_t* scalars, size_t n, uint64_t* out_jac);
/* points_jac may hold a single point broadcast to all scalars when n_points == 1 */
int b200_g1_mul_batch_bcast(const uint64_t* point_jac, const uint64_t* scalars, size_t n, uint64_t* out_jac);
int b200_g2_mul_batch_bcast(const uint64_t* point_jac, const uint64_t* scalars, size_t n, uint64_t* out_jac);

#ifdef __cplusplus
}
#endif
#endif /* B200SNARK_H */
