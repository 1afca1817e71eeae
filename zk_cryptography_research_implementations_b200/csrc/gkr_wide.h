// gkr_wide.h -- the device-resident layered circuit shared by gkr_wide_circuit.cu (construction: validation, duplicate
// detection and the three CSR orderings, all built on the GPU) and gkr_wide.cu (prover and verifier).
#pragma once
#include <vector>
#include "internal.h"

// ---------------------------------------------------------------- device-resident circuit
struct GateCsr {            // gates of one layer grouped by a key (left / right / out index)
    uint64_t* off = nullptr;    // [n_keys + 1]
    uint32_t* x = nullptr;      // first other index per gate (see users)
    uint32_t* y = nullptr;      // second other index per gate
    uint8_t* op = nullptr;      // 0 add, 1 mul
};
struct WideLayer {
    uint64_t n_gates = 0;
    GateCsr by_left;    // x = out,  y = right
    GateCsr by_right;   // x = out,  y = left
    GateCsr by_out;     // x = left, y = right
};
struct DevBuf {   // RAII device allocation
    zk::Fe* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p) { o.p = nullptr; }
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(uint64_t n) { return cudaMalloc(&p, (size_t)(n ? n : 1) * sizeof(zk::Fe)); }
};
struct zk_wide_circuit {
    uint32_t L = 0;
    std::vector<uint32_t> bits;   // bits[li] = log2(#values of layer li), li = 0..L (L = inputs)
    std::vector<WideLayer> layers;
    int device = 0;
    uint64_t* pool_off = nullptr;   // storage of every layer's CSR arrays (the GateCsr pointers point into these)
    uint32_t *pool_x = nullptr, *pool_y = nullptr;
    uint8_t* pool_op = nullptr;
    // prover workspace, allocated with the circuit so a prove never calls cudaMalloc/cudaFree:
    std::vector<DevBuf> W;        // all layer values (Circuit::evaluate result), resident
    DevBuf wtab, eqa, h1, h2, Wc, half_hi, half_lo, half_hi2, half_lo2;
    DevBuf pre_pg, pre_eh;        // overlapped phase-2 precomputation: one product per gate (widest layer), eq over the known challenges
};

