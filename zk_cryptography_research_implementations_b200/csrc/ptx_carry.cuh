// ptx_carry.cuh -- the extended-precision integer instructions of PTX as one-line wrappers.
//
// On the device every wrapper is a single `asm volatile` statement; ptxas fuses the
// mad.lo.cc / madc.hi.cc pairs that zk::fp emits on (even, odd) register pairs into
// IMAD.WIDE.U32(.X) -- one fma-pipe instruction per 32x32->64 product with the carry
// travelling in a predicate -- and the add/sub chains into IADD3(.X) on the alu pipe.
// The condition-code register is implicit PTX state that compiler-generated PTX never
// touches, and volatile asm statements keep their program order, so a chain may be split
// over several statements.
//
// With -DZK_HOST_EMU the same wrappers are emulated on the host with a software carry
// flag.  That build exists only so the limb gymnastics of fp.cuh can be unit-tested in
// a container without a GPU (tests/test_host_emu.py); it is not part of the library.
#pragma once
#include <stdint.h>

#if defined(ZK_HOST_EMU)
#define ZK_DEV inline
namespace zk { namespace ptx {
static thread_local uint32_t CC = 0;
inline uint32_t add_cc(uint32_t a, uint32_t b)  { uint64_t s = (uint64_t)a + b;      CC = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + CC; CC = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc(uint32_t a, uint32_t b)    { return a + b + CC; }
inline uint32_t sub_cc(uint32_t a, uint32_t b)  { uint64_t s = (uint64_t)a - b;      CC = (uint32_t)(s >> 63); return (uint32_t)s; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b - CC; CC = (uint32_t)(s >> 63); return (uint32_t)s; }
inline uint32_t subc(uint32_t a, uint32_t b)    { return a - b - CC; }
inline uint32_t mul_lo(uint32_t a, uint32_t b)  { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b)  { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c)  { return add_cc(a * b, c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(a * b, c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c)  { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c)    { return addc(mul_hi(a, b), c); }
// 64-bit slot forms: (lo, hi) (+)= a * b, carry in/out through CC
inline void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b)     { lo = mul_lo(a, b); hi = mul_hi(a, b); }
inline void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b)  { lo = mad_lo_cc(a, b, lo); hi = madc_hi_cc(a, b, hi); }
inline void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { lo = madc_lo_cc(a, b, lo); hi = madc_hi_cc(a, b, hi); }
}}  // namespace zk::ptx
#else
#define ZK_DEV __device__ __forceinline__
namespace zk { namespace ptx {
#define ZK_ASM2(name, ins)                                                        \
    ZK_DEV uint32_t name(uint32_t a, uint32_t b) {                                \
        uint32_t r;                                                               \
        asm volatile(ins " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));              \
        return r;                                                                 \
    }
#define ZK_ASM3(name, ins)                                                        \
    ZK_DEV uint32_t name(uint32_t a, uint32_t b, uint32_t c) {                    \
        uint32_t r;                                                               \
        asm volatile(ins " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));  \
        return r;                                                                 \
    }
ZK_ASM2(add_cc, "add.cc.u32")
ZK_ASM2(addc_cc, "addc.cc.u32")
ZK_ASM2(addc, "addc.u32")
ZK_ASM2(sub_cc, "sub.cc.u32")
ZK_ASM2(subc_cc, "subc.cc.u32")
ZK_ASM2(subc, "subc.u32")
ZK_ASM2(mul_lo, "mul.lo.u32")
ZK_ASM2(mul_hi, "mul.hi.u32")
ZK_ASM3(mad_lo_cc, "mad.lo.cc.u32")
ZK_ASM3(madc_lo_cc, "madc.lo.cc.u32")
ZK_ASM3(mad_hi_cc, "mad.hi.cc.u32")
ZK_ASM3(madc_hi_cc, "madc.hi.cc.u32")
ZK_ASM3(madc_hi, "madc.hi.u32")
#undef ZK_ASM2
#undef ZK_ASM3
// 64-bit slot forms.  The lo/hi halves MUST sit in one asm statement with in-place operands:
// only then does ptxas fuse them into a single IMAD.WIDE.U32(.X) (checked with cuobjdump -sass;
// split over two statements it emits IMAD + IMAD.HI.U32 + 2 x IADD3.X instead).
ZK_DEV void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
ZK_DEV void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
ZK_DEV void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
}}  // namespace zk::ptx
#endif
