// Package b200 is the cgo binding of libb200snark (include/b200snark.h): the marshalling layer a
// maintainer of arnaucube/go-snark-study links under groth16.GenerateProofs / snark.GenerateProofs /
// r1csqap.PolynomialField / bn128.G1,G2 to run the prove path on a B200.
//
// NOT COMPILED IN THIS REPOSITORY: neither the build image nor the GPU box has a Go toolchain
// (`go version`: not found).  The identical marshalling is exercised through ctypes by the Python
// mirror (go-snark-study_b200/_lib.py) in tests/.  See INTEGRATION.md.
package b200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../go-snark-study_b200/lib -lb200snark
#include <stdlib.h>
#include "b200snark.h"
*/
import "C"

import (
	"crypto/sha256"
	"encoding/binary"
	"errors"
	"hash"
	"math/big"
	"sync"
	"unsafe"
)

// R is the BN254 scalar-field order (bn128/bn128.go:46).
var R, _ = new(big.Int).SetString("21888242871839275222246405745257275088548364400416034343698204186575808495617", 10)

func check(rc C.int) error {
	if rc == 0 {
		return nil
	}
	return errors.New("libb200snark: " + C.GoString(C.b200_last_error()))
}

// limbs appends the 4 little-endian 64-bit limbs of x (0 <= x < 2^256) to dst.
func limbs(dst []uint64, x *big.Int) []uint64 {
	w := x.Bits() // amd64: big.Word is 64-bit, little endian
	for i := 0; i < 4; i++ {
		if i < len(w) {
			dst = append(dst, uint64(w[i]))
		} else {
			dst = append(dst, 0)
		}
	}
	return dst
}

// Scalar is what crosses the ABI for a *big.Int scalar: |e| mod r (the reference's MulScalar
// consumes |e|: fields/fq.go:138-140).
func Scalar(e *big.Int) *big.Int { return new(big.Int).Mod(new(big.Int).Abs(e), R) }

// Coeff is what crosses the ABI for a polynomial coefficient: e mod r (Euclidean, like Fq.Add/Mul).
func Coeff(e *big.Int) *big.Int { return new(big.Int).Mod(e, R) }

func FlatFr(v []*big.Int, reduce func(*big.Int) *big.Int) []uint64 {
	out := make([]uint64, 0, 4*len(v))
	for _, x := range v {
		out = limbs(out, reduce(x))
	}
	return out
}
func FlatG1(pts [][3]*big.Int) []uint64 {
	out := make([]uint64, 0, 12*len(pts))
	for _, p := range pts {
		for k := 0; k < 3; k++ {
			out = limbs(out, p[k])
		}
	}
	return out
}
func FlatG2(pts [][3][2]*big.Int) []uint64 {
	out := make([]uint64, 0, 24*len(pts))
	for _, p := range pts {
		for k := 0; k < 3; k++ {
			out = limbs(limbs(out, p[k][0]), p[k][1])
		}
	}
	return out
}
func fromLimbs(w []uint64) *big.Int {
	words := make([]big.Word, len(w))
	for i, x := range w {
		words[i] = big.Word(x)
	}
	return new(big.Int).SetBits(words)
}
func G1FromLimbs(w []uint64) [3]*big.Int {
	return [3]*big.Int{fromLimbs(w[0:4]), fromLimbs(w[4:8]), fromLimbs(w[8:12])}
}
func G2FromLimbs(w []uint64) [3][2]*big.Int {
	var p [3][2]*big.Int
	for k := 0; k < 3; k++ {
		p[k] = [2]*big.Int{fromLimbs(w[8*k : 8*k+4]), fromLimbs(w[8*k+4 : 8*k+8])}
	}
	return p
}
func u64(v []uint64) *C.uint64_t { return (*C.uint64_t)(unsafe.Pointer(&v[0])) }

// Groth16Key is a device-resident proving key (b200_pk_t).
type Groth16Key struct{ h C.b200_pk_t }

// LoadGroth16 uploads groth16.Pk once (b200_groth16_pk_load).
func LoadGroth16(at, b1 [][3]*big.Int, b2 [][3][2]*big.Int, bacDelta, ptd [][3]*big.Int, z []*big.Int,
	alpha1, beta1, delta1 [3]*big.Int, beta2, delta2 [3][2]*big.Int, nVars, nPublic int) (*Groth16Key, error) {
	fa, fb1, fb2, fc, fp := FlatG1(at[:nVars]), FlatG1(b1[:nVars]), FlatG2(b2[:nVars]), FlatG1(bacDelta[:nVars]), FlatG1(ptd)
	fz := FlatFr(z, Coeff)
	a1, be1, d1 := FlatG1([][3]*big.Int{alpha1}), FlatG1([][3]*big.Int{beta1}), FlatG1([][3]*big.Int{delta1})
	be2, d2 := FlatG2([][3][2]*big.Int{beta2}), FlatG2([][3][2]*big.Int{delta2})
	var k Groth16Key
	rc := C.b200_groth16_pk_load(u64(fa), u64(fb1), u64(fb2), u64(fc), C.size_t(nVars), u64(fp), C.size_t(len(ptd)),
		u64(fz), C.size_t(len(z)), u64(a1), u64(be1), u64(d1), u64(be2), u64(d2), C.size_t(nPublic), 0, &k.h)
	return &k, check(rc)
}

// Prove is groth16.GenerateProofs after the randomness has been drawn (groth16.go:231-238).
func (k *Groth16Key) Prove(w, px []*big.Int, r, s *big.Int) (piA [3]*big.Int, piB [3][2]*big.Int, piC [3]*big.Int, err error) {
	fw, fpx := FlatFr(w, Scalar), FlatFr(px, Coeff)
	fr, fs := limbs(nil, r), limbs(nil, s)
	a, b, c := make([]uint64, 12), make([]uint64, 24), make([]uint64, 12)
	rc := C.b200_groth16_prove(k.h, u64(fw), C.size_t(len(w)), u64(fpx), C.size_t(len(px)), u64(fr), u64(fs), u64(a), u64(b), u64(c))
	if err = check(rc); err != nil {
		return
	}
	return G1FromLimbs(a), G2FromLimbs(b), G1FromLimbs(c), nil
}

func (k *Groth16Key) Free() { C.b200_pk_free(k.h) }

// LoadGroth16Pair uploads the same groth16.Pk under the library's two prove contexts (B200_CFG_PK_CONTEXT 0 and 1: own
// tables, scratch, streams).  Two goroutines — one per key — may then call Prove concurrently: a call enqueues under the
// library mutex and waits for its proof with the mutex released, so the sort / bucket-tail / blinding-product chains of one
// proof overlap the bucket accumulation of the other (a proof server's steady state: +4 % proofs/s at 2^20 constraints,
// +30 % at 2^16, bench.py `proofs_in_flight`).  cgo releases the goroutine's OS thread for the duration of the call.
func LoadGroth16Pair(at, b1 [][3]*big.Int, b2 [][3][2]*big.Int, bacDelta, ptd [][3]*big.Int, z []*big.Int,
	alpha1, beta1, delta1 [3]*big.Int, beta2, delta2 [3][2]*big.Int, nVars, nPublic int) (keys [2]*Groth16Key, err error) {
	for ctx := 0; ctx < 2; ctx++ {
		if err = check(C.b200_config(C.B200_CFG_PK_CONTEXT, C.int(ctx))); err != nil {
			break
		}
		if keys[ctx], err = LoadGroth16(at, b1, b2, bacDelta, ptd, z, alpha1, beta1, delta1, beta2, delta2, nVars, nPublic); err != nil {
			break
		}
	}
	C.b200_config(C.B200_CFG_PK_CONTEXT, 0)
	return
}

// PolyMul / PolyDiv back r1csqap.PolynomialField.Mul / Div (r1csqap/r1csqap.go:57-84).
func PolyMul(a, b []*big.Int) ([]*big.Int, error) {
	fa, fb := FlatFr(a, Coeff), FlatFr(b, Coeff)
	out := make([]uint64, 4*(len(a)+len(b)-1))
	if err := check(C.b200_poly_mul(u64(fa), C.size_t(len(a)), u64(fb), C.size_t(len(b)), u64(out))); err != nil {
		return nil, err
	}
	res := make([]*big.Int, len(a)+len(b)-1)
	for i := range res {
		res[i] = fromLimbs(out[4*i : 4*i+4])
	}
	return res, nil
}

// Fq12FromLimbs unpacks b200_pairing_batch's 48 words into the reference's [2][3][2]*big.Int (fields/fq12.go).
func Fq12FromLimbs(w []uint64) (r [2][3][2]*big.Int) {
	for h := 0; h < 2; h++ {
		for k := 0; k < 3; k++ {
			for c := 0; c < 2; c++ {
				o := 4 * (6*h + 2*k + c)
				r[h][k][c] = fromLimbs(w[o : o+4])
			}
		}
	}
	return
}

// Pairing backs bn128.Bn128.Pairing (bn128/bn128.go:179-186).
func Pairing(p1 [3]*big.Int, p2 [3][2]*big.Int) ([2][3][2]*big.Int, error) {
	out := make([]uint64, 48)
	g1, g2 := FlatG1([][3]*big.Int{p1}), FlatG2([][3][2]*big.Int{p2})
	err := check(C.b200_pairing_batch(u64(g1), u64(g2), 1, u64(out)))
	return Fq12FromLimbs(out), err
}

// Fq12Mul backs fields.Fq12.Mul in the verification equations (fields/fq12.go:72-84).
func Fq12Mul(a, b [2][3][2]*big.Int) ([2][3][2]*big.Int, error) {
	flat := func(x [2][3][2]*big.Int) []uint64 {
		out := make([]uint64, 0, 48)
		for h := 0; h < 2; h++ {
			for k := 0; k < 3; k++ {
				for c := 0; c < 2; c++ {
					out = limbs(out, x[h][k][c])
				}
			}
		}
		return out
	}
	fa, fb := flat(a), flat(b)
	out := make([]uint64, 48)
	err := check(C.b200_fq12_mul_batch(u64(fa), u64(fb), 1, u64(out)))
	return Fq12FromLimbs(out), err
}

// ---- proving-key cache for the drop-in GenerateProofs(circuit, pk, w, px) form --------------------------------------
// The reference passes Pk BY VALUE (groth16/groth16.go:225), so neither pointer identity nor slice headers identify a
// key across calls (a caller may rebuild an equal Pk, or reuse a backing array for a different one).  The cache key is
// a CONTENT fingerprint: SHA-256 over the array lengths, NVars/NPublic, the three blinding points, Z, and the first
// and last `fpSample` points of every CRS array.  (A maintainer who wants no hashing at all uses the explicit handle
// API instead: LoadGroth16 once, key.Prove per proof.)  At most `cap` keys stay resident; eviction frees device tables.
const fpSample = 8

type fingerprint [32]byte

func hashInt(h hash.Hash, x *big.Int) { h.Write(x.Bytes()); h.Write([]byte{0xff}) }
func hashG1(h hash.Hash, pts [][3]*big.Int) {
	binary.Write(h, binary.LittleEndian, uint64(len(pts)))
	for i, p := range pts {
		if i >= fpSample && i < len(pts)-fpSample {
			continue
		}
		for k := 0; k < 3; k++ {
			hashInt(h, p[k])
		}
	}
}
func hashG2(h hash.Hash, pts [][3][2]*big.Int) {
	binary.Write(h, binary.LittleEndian, uint64(len(pts)))
	for i, p := range pts {
		if i >= fpSample && i < len(pts)-fpSample {
			continue
		}
		for k := 0; k < 3; k++ {
			hashInt(h, p[k][0])
			hashInt(h, p[k][1])
		}
	}
}

// Groth16Fingerprint identifies a groth16.Pk by content (see above).
func Groth16Fingerprint(at, b1 [][3]*big.Int, b2 [][3][2]*big.Int, bacDelta, ptd [][3]*big.Int, z []*big.Int,
	alpha1, beta1, delta1 [3]*big.Int, beta2, delta2 [3][2]*big.Int, nVars, nPublic int) fingerprint {
	h := sha256.New()
	binary.Write(h, binary.LittleEndian, [2]uint64{uint64(nVars), uint64(nPublic)})
	hashG1(h, at)
	hashG1(h, b1)
	hashG2(h, b2)
	hashG1(h, bacDelta)
	hashG1(h, ptd)
	hashG1(h, [][3]*big.Int{alpha1, beta1, delta1})
	hashG2(h, [][3][2]*big.Int{beta2, delta2})
	for _, c := range z {
		hashInt(h, c)
	}
	var f fingerprint
	copy(f[:], h.Sum(nil))
	return f
}

// KeyCache keeps the most recently used device keys (safe for concurrent callers).
type KeyCache struct {
	mu    sync.Mutex
	cap   int
	order []fingerprint
	keys  map[fingerprint]*Groth16Key
}

func NewKeyCache(capacity int) *KeyCache {
	return &KeyCache{cap: capacity, keys: make(map[fingerprint]*Groth16Key)}
}

// Get returns the cached key for fp or loads it with load() and evicts (and frees) the least recently used one.
func (c *KeyCache) Get(fp fingerprint, load func() (*Groth16Key, error)) (*Groth16Key, error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	if k, ok := c.keys[fp]; ok {
		for i, f := range c.order {
			if f == fp {
				c.order = append(append(c.order[:i:i], c.order[i+1:]...), fp)
				break
			}
		}
		return k, nil
	}
	k, err := load()
	if err != nil {
		return nil, err
	}
	c.keys[fp] = k
	c.order = append(c.order, fp)
	for len(c.order) > c.cap {
		old := c.order[0]
		c.order = c.order[1:]
		c.keys[old].Free()
		delete(c.keys, old)
	}
	return k, nil
}

// ---- sparse R1CS, witness -> px, witness -> proof (include/b200snark.h: b200_r1cs_load, b200_qap_px, ..) ---------------
// The dense a, b, c [][]*big.Int of r1csqap.R1CSToQAP (r1csqap.go:161) cannot exist at 2^16+ constraints; a caller at
// that size keeps the R1CS in CSR form and calls CombinePolynomials directly on it.
type R1CS struct {
	h    C.b200_r1cs_t
	N, M int
}

// DenseToCSR converts the reference's dense matrix (rows = constraints) to CSR with coefficients mod r.
func DenseToCSR(m [][]*big.Int) (rowptr, col []uint32, val []uint64) {
	rowptr = append(rowptr, 0)
	for _, row := range m {
		for i, x := range row {
			if c := Coeff(x); c.Sign() != 0 {
				col = append(col, uint32(i))
				val = limbs(val, c)
			}
		}
		rowptr = append(rowptr, uint32(len(col)))
	}
	return
}

func u32(v []uint32) *C.uint32_t {
	if len(v) == 0 {
		return nil
	}
	return (*C.uint32_t)(unsafe.Pointer(&v[0]))
}
func u64n(v []uint64) *C.uint64_t {
	if len(v) == 0 {
		return nil
	}
	return u64(v)
}

// LoadR1CS uploads the three matrices (dense reference form) once.
func LoadR1CS(a, b, c [][]*big.Int) (*R1CS, error) {
	ar, ac, av := DenseToCSR(a)
	br, bc, bv := DenseToCSR(b)
	cr, cc, cv := DenseToCSR(c)
	r := &R1CS{N: len(a), M: len(a[0])}
	rc := C.b200_r1cs_load(C.size_t(r.N), C.size_t(r.M), u32(ar), u32(ac), u64n(av), u32(br), u32(bc), u64n(bv),
		u32(cr), u32(cc), u64n(cv), &r.h)
	return r, check(rc)
}

// CombinePolynomials == PolynomialField.CombinePolynomials(w, R1CSToQAP(a, b, c)...) (r1csqap.go:161-210).
func (r *R1CS) CombinePolynomials(w []*big.Int) (ax, bx, cx, px []*big.Int, err error) {
	fw := FlatFr(w, Coeff)
	oa, ob, oc, op := make([]uint64, 4*r.N), make([]uint64, 4*r.N), make([]uint64, 4*r.N), make([]uint64, 4*(2*r.N-1))
	if err = check(C.b200_qap_px(r.h, u64(fw), C.size_t(len(w)), u64(oa), u64(ob), u64(oc), u64(op))); err != nil {
		return
	}
	un := func(v []uint64) []*big.Int {
		out := make([]*big.Int, len(v)/4)
		for i := range out {
			out[i] = fromLimbs(v[4*i : 4*i+4])
		}
		return out
	}
	return un(oa), un(ob), un(oc), un(op), nil
}
func (r *R1CS) Free() { C.b200_r1cs_free(r.h) }

// ProveWitness is GenerateProofs fed from the witness: px never leaves the device (b200_groth16_prove_witness).
func (k *Groth16Key) ProveWitness(r1cs *R1CS, w []*big.Int, r, s *big.Int) (piA [3]*big.Int, piB [3][2]*big.Int, piC [3]*big.Int, err error) {
	fw := FlatFr(w, Scalar)
	fr, fs := limbs(nil, r), limbs(nil, s)
	a, b, c := make([]uint64, 12), make([]uint64, 24), make([]uint64, 12)
	rc := C.b200_groth16_prove_witness(k.h, r1cs.h, u64(fw), C.size_t(len(w)), u64(fr), u64(fs), u64(a), u64(b), u64(c))
	if err = check(rc); err != nil {
		return
	}
	return G1FromLimbs(a), G2FromLimbs(b), G1FromLimbs(c), nil
}

// ---- multi-GPU: one process per GPU, NCCL inside the library ----------------------------------------------------------
// Rank 0 calls CommUniqueID and hands the 128 bytes to the other ranks (any side channel); every rank calls CommInit after
// choosing its device, loads its shard with LoadGroth16Shard, and then calls key.Prove with the SAME w, px, r, s: the
// proof comes back on every rank.
func CommUniqueID() (id [128]byte, err error) {
	err = check(C.b200_comm_unique_id((*C.uint8_t)(unsafe.Pointer(&id[0]))))
	return
}
func CommInit(id [128]byte, rank, world int) error {
	return check(C.b200_comm_init((*C.uint8_t)(unsafe.Pointer(&id[0])), C.int(rank), C.int(world)))
}
func CommDestroy() error { return check(C.b200_comm_destroy()) }

// TuneShardsForTwoInFlight sets the partition of sharded keys loaded AFTERWARDS the way bench.py does when it keeps two
// proofs in flight (LoadGroth16Pair on every rank): every shard on the batched-affine tree, an A / B1 term 0.85 and a G2
// term 2.75 of a C||PTD term (profiles/r2_notes.md sections 11 and 16).  The defaults suit one proof at a time.
func TuneShardsForTwoInFlight() error {
	for _, kv := range [][2]C.int{{C.B200_CFG_SHARD_AFFINE_MIN_G1, 1}, {C.B200_CFG_SHARD_AFFINE_MIN_G2, 1},
		{C.B200_CFG_SHARD_W_AB, 85}, {C.B200_CFG_SHARD_W_G2, 275}} {
		if err := check(C.b200_config(kv[0], kv[1])); err != nil {
			return err
		}
	}
	return nil
}

func LoadGroth16Shard(at, b1 [][3]*big.Int, b2 [][3][2]*big.Int, bacDelta, ptd [][3]*big.Int, z []*big.Int,
	alpha1, beta1, delta1 [3]*big.Int, beta2, delta2 [3][2]*big.Int, nVars, nPublic, rank, world int) (*Groth16Key, error) {
	fa, fb1, fb2, fc, fp := FlatG1(at[:nVars]), FlatG1(b1[:nVars]), FlatG2(b2[:nVars]), FlatG1(bacDelta[:nVars]), FlatG1(ptd)
	fz := FlatFr(z, Coeff)
	a1, be1, d1 := FlatG1([][3]*big.Int{alpha1}), FlatG1([][3]*big.Int{beta1}), FlatG1([][3]*big.Int{delta1})
	be2, d2 := FlatG2([][3][2]*big.Int{beta2}), FlatG2([][3][2]*big.Int{delta2})
	var k Groth16Key
	rc := C.b200_groth16_pk_load_shard(u64(fa), u64(fb1), u64(fb2), u64(fc), C.size_t(nVars), u64(fp), C.size_t(len(ptd)),
		u64(fz), C.size_t(len(z)), u64(a1), u64(be1), u64(d1), u64(be2), u64(d2), C.size_t(nPublic), 0,
		C.int(rank), C.int(world), &k.h)
	return &k, check(rc)
}
