// gather_probe.cu -- what can B200 do on RANDOM 24-byte gathers over a 2.4 GB array?  (the floor of the voxel-grid
// reduce step: per-voxel sums need every point's Vec3f64 position in voxel order, voxel_grid.rs:339-386.)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o benchmarks/build/gather_probe benchmarks/gather_probe.cu
// Prints one JSON line per variant: ms per 100 M gathers, G gathers/s, effective GB/s on 24 B/gather.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
    return z;
}
__global__ void fill_idx(uint32_t* idx, uint64_t n, uint64_t window) {
    // window == n: uniformly random; smaller: random inside consecutive windows (locality like a pre-partitioned cloud)
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t w0 = i / window * window;
        const uint64_t wl = (n - w0) < window ? (n - w0) : window;
        idx[i] = (uint32_t)(w0 + mix(i * 0x9E3779B97F4A7C15ull + 1) % wl);
    }
}
__global__ void fill_pos(double* p, uint64_t n3) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n3; i += (uint64_t)gridDim.x * blockDim.x) p[i] = (double)(i & 1023);
}

template <int HINT>
__device__ __forceinline__ void ld24(const uint8_t* p, double& x, double& y, double& z) {
    if (HINT == 64) {
        if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
            asm volatile("ld.global.nc.L2::64B.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p));
            asm volatile("ld.global.nc.L2::64B.f64 %0, [%1];" : "=d"(z) : "l"(p + 16));
        } else {
            asm volatile("ld.global.nc.L2::64B.f64 %0, [%1];" : "=d"(x) : "l"(p));
            asm volatile("ld.global.nc.L2::64B.v2.f64 {%0, %1}, [%2];" : "=d"(y), "=d"(z) : "l"(p + 8));
        }
    } else {
        const double* d = reinterpret_cast<const double*>(p);
        x = __ldg(d); y = __ldg(d + 1); z = __ldg(d + 2);
    }
}

// every thread gathers U points per iteration (U independent gathers in flight), grid-stride over the sorted positions
template <int U, int HINT>
__global__ void __launch_bounds__(256) gather_kernel(const uint8_t* __restrict__ pos, uint32_t stride, const uint32_t* __restrict__ idx,
                                                     uint64_t n, double* __restrict__ out) {
    double sx = 0, sy = 0, sz = 0;
    const uint64_t per_block = (uint64_t)blockDim.x * U;
    for (uint64_t b = blockIdx.x * per_block; b < n; b += (uint64_t)gridDim.x * per_block) {
        uint32_t j[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const uint64_t i = b + (uint64_t)u * blockDim.x + threadIdx.x; j[u] = i < n ? idx[i] : 0u; }
        double x[U], y[U], z[U];
#pragma unroll
        for (int u = 0; u < U; ++u) ld24<HINT>(pos + (uint64_t)j[u] * stride, x[u], y[u], z[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) { sx += x[u]; sy += y[u]; sz += z[u]; }
    }
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    out[t] = sx + sy + sz;
}

// the same through shared memory staging: gather -> smem -> (nothing), measures the staging cost
template <int U>
__global__ void __launch_bounds__(256) gather_smem_kernel(const uint8_t* __restrict__ pos, uint32_t stride, const uint32_t* __restrict__ idx,
                                                          uint64_t n, double* __restrict__ out) {
    __shared__ double sx[U * 256], sy[U * 256], sz[U * 256];
    double acc = 0;
    const uint64_t per_block = (uint64_t)blockDim.x * U;
    for (uint64_t b = blockIdx.x * per_block; b < n; b += (uint64_t)gridDim.x * per_block) {
        uint32_t j[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const uint64_t i = b + (uint64_t)u * blockDim.x + threadIdx.x; j[u] = i < n ? idx[i] : 0u; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            double x, y, z;
            ld24<64>(pos + (uint64_t)j[u] * stride, x, y, z);
            sx[u * 256 + threadIdx.x] = x; sy[u * 256 + threadIdx.x] = y; sz[u * 256 + threadIdx.x] = z;
        }
        __syncthreads();
        const int k = (threadIdx.x * 7 + 3) & 255;
#pragma unroll
        for (int u = 0; u < U; ++u) acc += sx[u * 256 + k] + sy[u * 256 + k] + sz[u * 256 + k];
        __syncthreads();
    }
    out[blockIdx.x * (uint64_t)blockDim.x + threadIdx.x] = acc;
}

template <class F>
static float time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 100000000ull;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint8_t* pos; uint32_t* idx; double* out;
    CK(cudaMalloc(&pos, n * 32 + 64)); CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&out, (size_t)sms * 32 * 256 * 8));
    fill_pos<<<sms * 8, 256>>>((double*)pos, n * 4);
    const uint64_t windows[] = {n, 1u << 20, 1u << 16};
    for (uint64_t w : windows) {
        fill_idx<<<sms * 8, 256>>>(idx, n, w);
        CK(cudaDeviceSynchronize());
        auto report = [&](const char* name, int stride, float ms) {
            printf("{\"variant\": \"%s\", \"stride\": %d, \"window\": %llu, \"n\": %llu, \"ms\": %.4f, \"G_gathers_per_s\": %.2f, \"GBps_24B\": %.1f}\n", name, stride,
                   (unsigned long long)w, (unsigned long long)n, ms, n / ms / 1e6, 28.0 * n / ms / 1e6);
            fflush(stdout);
        };
        for (int stride : {24, 32}) {
            report("U1 hint64 8 CTA/SM", stride, time_ms([&] { gather_kernel<1, 64><<<sms * 8, 256>>>(pos, stride, idx, n, out); }));
            report("U4 hint64 8 CTA/SM", stride, time_ms([&] { gather_kernel<4, 64><<<sms * 8, 256>>>(pos, stride, idx, n, out); }));
            report("U8 hint64 8 CTA/SM", stride, time_ms([&] { gather_kernel<8, 64><<<sms * 8, 256>>>(pos, stride, idx, n, out); }));
            report("U8 plain 8 CTA/SM", stride, time_ms([&] { gather_kernel<8, 0><<<sms * 8, 256>>>(pos, stride, idx, n, out); }));
            report("U16 hint64 4 CTA/SM", stride, time_ms([&] { gather_kernel<16, 64><<<sms * 4, 256>>>(pos, stride, idx, n, out); }));
        }
        report("U8 hint64 via smem 4 CTA/SM", 24, time_ms([&] { gather_smem_kernel<8><<<sms * 4, 256>>>(pos, 24, idx, n, out); }));
    }
    return 0;
}
