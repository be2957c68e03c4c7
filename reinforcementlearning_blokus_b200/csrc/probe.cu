// probe.cu -- integer-pipe throughput microbenchmarks (measurement infrastructure for bench.py; not part
// of the product ABI).  Measures how many warp instructions per clock an SM sub-partition sustains for
//   mode 0: LOP3 + SHF only (both on the ALU pipe: expected 0.5 warp-instr/clk/SMSP, B300_MICROARCH.md)
//   mode 1: LOP3/SHF interleaved with IMAD (ALU + FMA pipes: expected up to 1 warp-instr/clk/SMSP)
// The row-sliced legality kernels are a LOP3/SHF/IMAD/POPC mix, so their issue-slot roofline sits between.
#include <stdint.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k_probe(uint32_t* out, int iters, uint32_t seed) {
    uint32_t x0 = threadIdx.x + seed, x1 = x0 * 3u, x2 = x0 * 5u, x3 = x0 * 7u;
    uint32_t x4 = x0 * 11u, x5 = x0 * 13u, x6 = x0 * 17u, x7 = x0 * 19u;
    const uint32_t k = seed | 1u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) {
                x0 = (x0 >> 1) ^ (x1 & k); x1 = (x1 >> 3) ^ (x2 & k); x2 = (x2 >> 5) ^ (x3 & k); x3 = (x3 >> 7) ^ (x0 & k);
                x4 = (x4 >> 1) ^ (x5 & k); x5 = (x5 >> 3) ^ (x6 & k); x6 = (x6 >> 5) ^ (x7 & k); x7 = (x7 >> 7) ^ (x4 & k);
            } else {
                x0 = (x0 >> 1) ^ (x1 & k); x1 = x1 * 5u + x2; x2 = x2 * 9u + x3; x3 = (x3 >> 7) ^ (x0 & k);
                x4 = x4 * 5u + x5; x5 = x5 * 9u + x6; x6 = (x6 >> 5) ^ (x7 & k); x7 = x7 * 5u + x4; x7 = x7 * 9u + x0;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

// Returns elapsed milliseconds of one launch of `blocks` x 256 threads running `iters` loop trips; the
// caller converts with the SASS instruction count per trip (see bench.py: 16 per trip for either mode,
// checked with cuobjdump at build time).
extern "C" int bk_probe_int(int mode, int blocks, int iters, float* ms_out) {
    uint32_t* out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * 256 * sizeof(uint32_t)) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {   // first launch warms up
        cudaEventRecord(e0);
        if (mode == 0) k_probe<0><<<blocks, 256>>>(out, iters, 12345u);
        else k_probe<1><<<blocks, 256>>>(out, iters, 12345u);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return -1; }
    }
    cudaEventElapsedTime(ms_out, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    return 0;
}
