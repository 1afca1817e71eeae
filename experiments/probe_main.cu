// experiments/probe_main.cu -- standalone runner of the round-1 pipe-rate probes (formerly zk_arith_probe kinds 3..9).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I zk_cryptography_research_implementations_b200/csrc \
//        -I experiments experiments/probe_main.cu -o /tmp/zk_probe && /tmp/zk_probe
// Prints G operations/s per kind at 1 and 2 blocks per SM.  Results of round 1: profiles/r01b/probe_products.json.
#include <cstdio>
#include "probe_kernels.cuh"

using namespace zk;

template <class Launch> static double timed(Launch launch, double ops) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0;
    for (int pass = 0; pass < 2; ++pass) {   // first pass warms up
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ops / (ms * 1e-3) / 1e9;
}

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
    const uint32_t iters = 1500;
    void* scratch = nullptr;
    cudaMalloc(&scratch, (size_t)prop.multiProcessorCount * 2 * kThreads * sizeof(Fe));
    for (int bps = 1; bps <= 2; ++bps) {
        const int grid = prop.multiProcessorCount * bps;
        const double threads = (double)grid * kThreads;
        printf("fp64 DFMA                    @%d blocks/SM: %8.1f G/s\n", bps, timed([&] { dfma_probe_kernel<0><<<grid, kThreads>>>((double*)scratch, iters); }, threads * 8 * iters));
        printf("IMAD.WIDE.U32 plain          @%d blocks/SM: %8.1f G/s\n", bps, timed([&] { imad_probe_kernel<4><<<grid, kThreads>>>((uint64_t*)scratch, iters); }, threads * 8 * iters));
        printf("IMAD 32-bit                  @%d blocks/SM: %8.1f G/s\n", bps, timed([&] { imad_probe_kernel<5><<<grid, kThreads>>>((uint64_t*)scratch, iters); }, threads * 8 * iters));
        printf("IMAD.WIDE.U32.X chained      @%d blocks/SM: %8.1f G/s\n", bps, timed([&] { imad_probe_kernel<6><<<grid, kThreads>>>((uint64_t*)scratch, iters); }, threads * 8 * iters));
        printf("mul_acc_cols carry-out slots @%d blocks/SM: %8.1f G products/s\n", bps, timed([&] { cols_probe_kernel<0><<<grid, kThreads>>>((Fe*)scratch, iters); }, threads * 2 * iters));
        printf("radix-2^29 flag-free         @%d blocks/SM: %8.1f G products/s\n", bps, timed([&] { cols29_probe_kernel<0, 0><<<grid, kThreads>>>((Fe*)scratch, iters); }, threads * 2 * iters));
        printf("mul_acc chained, same stream @%d blocks/SM: %8.1f G products/s\n", bps, timed([&] { cols29_probe_kernel<0, 1><<<grid, kThreads>>>((Fe*)scratch, iters); }, threads * 2 * iters));
    }
    cudaFree(scratch);
    return 0;
}
