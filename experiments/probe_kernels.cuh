// experiments/probe_kernels.cuh -- register-resident pipe-rate probes used for the round-1 A/B measurements
// (formerly kinds 3..9 of zk_arith_probe).  Not compiled into libzkb200; the product keeps kinds 0..2 (Montgomery
// product, fold by scalar, unreduced multiply-accumulate), which bench.py uses for the integer roofline.
#pragma once
#include "fp_columns.cuh"
#include "kernels.cuh"

namespace zk {

// Probe of the carry-chain-free unreduced product (Fp::mul_acc_cols): two independent chains per thread (kind 7 of
// zk_arith_probe), to set against mul_acc (kind 2).
template <int FID> __global__ void __launch_bounds__(kThreads) cols_probe_kernel(Fe* out, uint32_t iters) {
    Fe x[2], y[2];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x[c].v[k] = (threadIdx.x * 2654435761u + blockIdx.x + 977u * c + k) & 0x0fffffffu;
            y[c].v[k] = (threadIdx.x * 40503u + 31u * blockIdx.x + 13u * c + 7u * k) & 0x0fffffffu;
        }
    typename FpColumns<FID>::ColAcc acc[2];
    FpColumns<FID>::cols_init(acc[0]);
    FpColumns<FID>::cols_init(acc[1]);
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            FpColumns<FID>::mul_acc_cols(acc[c], x[c], y[c]);
            x[c].v[0] ^= acc[c].top[14];
        }
    }
    Fe r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = 0;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t t[17];
        FpColumns<FID>::cols_to_limbs(t, acc[c]);
#pragma unroll
        for (int k = 0; k < 8; ++k) r.v[k] += x[c].v[k] * (2 * c + 1) + t[k] + t[k + 8];
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// Probe of the flag-free radix-2^29 product (Fp::mul_cols29, operand conversion and the flush every six products
// included) against the chained mul_acc ON THE SAME OPERAND STREAM: two independent chains per thread, every limb of
// both operands changes from product to product (otherwise ptxas hoists the loop-invariant digit products out of the
// loop).  MODE 0: radix 2^29 (kind 8 of zk_arith_probe), MODE 1: chained mul_acc (kind 9).
template <int FID, int MODE> __global__ void __launch_bounds__(kThreads) cols29_probe_kernel(Fe* out, uint32_t iters) {
    Fe x[2], y[2];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x[c].v[k] = (threadIdx.x * 2654435761u + blockIdx.x + 977u * c + k) & 0x0fffffffu;
            y[c].v[k] = (threadIdx.x * 40503u + 31u * blockIdx.x + 13u * c + 7u * k) & 0x0fffffffu;
        }
    typename FpColumns<FID>::Cols29 cols[2];
    uint32_t acc[2][17];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        FpColumns<FID>::cols29_init(cols[c]);
#pragma unroll
        for (int k = 0; k < 17; ++k) acc[c][k] = 0;
    }
    for (uint32_t it = 0; it < iters; it += FpColumns<FID>::kCols29Budget) {
#pragma unroll
        for (int u = 0; u < FpColumns<FID>::kCols29Budget; ++u) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (MODE == 0) {
                    typename FpColumns<FID>::Digits29 dx, dy;
                    FpColumns<FID>::to_digits29(dx, x[c]);
                    FpColumns<FID>::to_digits29(dy, y[c]);
                    FpColumns<FID>::mul_cols29(cols[c], dx, dy);
                } else {
                    Fp<FID>::mul_acc(acc[c], x[c], y[c]);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {   // fresh operands for the next product (16 alu instructions in both modes)
                    x[c].v[k] ^= MODE == 0 ? (uint32_t)cols[c].c[k + 4] : acc[c][k + 4];
                    y[c].v[k] += x[c].v[k];
                }
            }
        }
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 2; ++c) FpColumns<FID>::cols29_flush(acc[c], cols[c]);
        }
    }
    Fe r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = 0;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) r.v[k] += x[c].v[k] * (2 * c + 1) + acc[c][k] + acc[c][k + 8];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// FP64 pipe probe: 8 independent DFMA chains per thread (kind 3 of zk_arith_probe).  Not used by any kernel;
// it answers whether a double-precision limb product (Emmart-style 52-bit limbs, 2 DFMA per product) could
// relieve the half-rate IMAD.WIDE pipe in a later round.
template <int UNUSED = 0> __global__ void __launch_bounds__(kThreads) dfma_probe_kernel(double* out, uint32_t iters) {
    double x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = 1.0 + 1e-9 * (threadIdx.x + 13 * c + blockIdx.x);
    const double a = 1.0000001, b = 1e-12;
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = fma(x[c], a, b);
    }
    double r = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) r += x[c];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// integer-multiply pipe probes (kinds 4, 5, 6 of zk_arith_probe): 8 independent chains per thread of
//   4: mad.wide.u32 (IMAD.WIDE.U32, 64-bit accumulate, no carry flag)
//   5: mad.lo.u32   (IMAD, 32-bit)
//   6: mad.lo.cc / madc.hi.cc pairs (IMAD.WIDE.U32.X, carry chained) -- what fp.cuh emits
template <int KIND> __global__ void __launch_bounds__(kThreads) imad_probe_kernel(uint64_t* out, uint32_t iters) {
    uint64_t acc[8];
    uint32_t a[8], b = threadIdx.x * 2654435761u + 12345u;
#pragma unroll
    for (int c = 0; c < 8; ++c) { acc[c] = blockIdx.x + c; a[c] = threadIdx.x * 40503u + 977u * c + 1u; }
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (KIND == 4) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a[c]), "r"(b));
            } else if (KIND == 5) {
                uint32_t lo = (uint32_t)acc[c];
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a[c]), "r"(b));
                acc[c] = lo;
            } else {
                uint32_t lo = (uint32_t)acc[c], hi = (uint32_t)(acc[c] >> 32);
                if (c == 0) ptx::mad_wide_cc(lo, hi, a[c], b);
                else ptx::madc_wide_cc(lo, hi, a[c], b);
                acc[c] = ((uint64_t)hi << 32) | lo;
            }
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) r += acc[c] * (2 * c + 1);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

}  // namespace zk
