// Host/device portability layer.
//
// All arithmetic headers (fp.cuh, fp2.cuh, ec.cuh) are written once against the
// carry-chain primitives below.  Under nvcc device compilation they are single
// PTX instructions (add.cc / madc.lo.cc / madc.hi.cc ... -> IADD3.X /
// IMAD.WIDE.U32.X in SASS); compiled for the host (g++, used only by the CPU
// unit tests of the kernel arithmetic in tests/test_host_arith.py) they are
// emulated with an explicit carry flag so the *same* chain logic is tested
// without a GPU.  The host build is a test vehicle, not a product fallback: no
// product entry point reaches it (capi.cu is device-only).
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#define HDC __host__ __device__ __forceinline__ constexpr
#define HDN __host__ __device__ __noinline__  // big tower routines: one copy, called (pairing.cuh)
#define DEV __device__ __forceinline__
#else
#define HD inline
#define HDC inline constexpr
#define HDN inline
#define DEV inline
#endif

namespace b200 {
namespace cc {

#ifdef __CUDA_ARCH__
// asm volatile: keeps the relative order of the chain and forbids CSE of
// textually identical instructions that differ only in the incoming carry.
DEV uint32_t add_cc(uint32_t a, uint32_t b)  { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;"  : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV uint32_t addc(uint32_t a, uint32_t b)    { uint32_t r; asm volatile("addc.u32 %0, %1, %2;"    : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV uint32_t sub_cc(uint32_t a, uint32_t b)  { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;"  : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV uint32_t subc(uint32_t a, uint32_t b)    { uint32_t r; asm volatile("subc.u32 %0, %1, %2;"    : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV uint32_t mul_lo(uint32_t a, uint32_t b)  { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;"  : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV uint32_t mul_hi(uint32_t a, uint32_t b)  { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;"  : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV void mul_wide(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
  uint64_t t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a), "r"(b)); lo = (uint32_t)t; hi = (uint32_t)(t >> 32);
}
DEV uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c)  { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;"  : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEV uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c)  { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;"  : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEV uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEV uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEV uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c)    { uint32_t r; asm volatile("madc.lo.u32 %0, %1, %2, %3;"    : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEV uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c)    { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;"    : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// Host emulation of the PTX condition-code register (CC.CF): carry for add /
// mad chains, borrow for sub chains.
inline uint32_t& cf() { static thread_local uint32_t f = 0; return f; }
inline uint32_t add_cc(uint32_t a, uint32_t b)  { uint64_t t = (uint64_t)a + b;        cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + cf(); cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b)    { return a + b + cf(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b)  { uint64_t t = (uint64_t)a - b;        cf() = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - cf(); cf() = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b)    { return a - b - cf(); }
inline uint32_t mul_lo(uint32_t a, uint32_t b)  { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b)  { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline void mul_wide(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) { uint64_t t = (uint64_t)a * b; lo = (uint32_t)t; hi = (uint32_t)(t >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c)  { uint64_t t = (uint64_t)(a * b) + c;               cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c)  { uint64_t t = (uint64_t)mul_hi(a, b) + c;          cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(a * b) + c + cf();        cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)mul_hi(a, b) + c + cf();   cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c)    { return a * b + c + cf(); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c)    { return mul_hi(a, b) + c + cf(); }
#endif

}  // namespace cc
}  // namespace b200
