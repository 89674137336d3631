// fp64_probe.cu -- measured FP64 instruction rate of the GPU (the roofline denominator of the RANSAC inlier ranking, which is
// FP64-compute-bound: 6 separately rounded DMUL / DADD + 2 DSETP per point-model test, no FMA contraction allowed because the
// reference rounds every product and sum, segmentation.rs:31-44).  Prints G instructions/s for DFMA, DMUL+DADD (non-fused) chains.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o benchmarks/build/fp64_probe benchmarks/fp64_probe.cu
#include <cuda_runtime.h>

#include <cstdio>

template <int MODE>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {  // 8 independent DFMA chains
            x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
            x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
        } else {          // 8 independent chains of DMUL then DADD (two instructions, two roundings)
            x0 = __dadd_rn(__dmul_rn(x0, a), b); x1 = __dadd_rn(__dmul_rn(x1, a), b); x2 = __dadd_rn(__dmul_rn(x2, a), b);
            x3 = __dadd_rn(__dmul_rn(x3, a), b); x4 = __dadd_rn(__dmul_rn(x4, a), b); x5 = __dadd_rn(__dmul_rn(x5, a), b);
            x6 = __dadd_rn(__dmul_rn(x6, a), b); x7 = __dadd_rn(__dmul_rn(x7, a), b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int MODE>
static void run(const char* name, double* out, int sms, int per_iter) {
    const int iters = 20000, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, 256>>>(out, 100, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double instr = (double)blocks * 256 * iters * per_iter;
    printf("{\"variant\": \"%s\", \"ms\": %.3f, \"G_fp64_instr_per_s\": %.1f, \"TFLOPs_if_fma\": %.2f}\n", name, ms, instr / ms / 1e6, 2 * instr / ms / 1e9);
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 8);
    run<0>("DFMA, 8 chains/thread", out, sms, 8);
    run<1>("DMUL + DADD (non-fused), 8 chains/thread", out, sms, 16);
    return 0;
}
