// FP64 pipe microbenchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync f64) issue rates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
// Output: one JSON line with TFLOP/s per variant; used to fix the FP64 roofline denominator
// (MEASURED_PEAKS.json only carries HBM and bf16).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, double s) {
    double a0 = threadIdx.x, a1 = 1.0, a2 = 2.0, a3 = 3.0, a4 = 4., a5 = 5., a6 = 6., a7 = 7.;
    double b = s, c = 1e-9;
#pragma unroll 16
    for (int i = 0; i < ITERS; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void k_dmma884(double* out, double s) {
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = j; c[j][1] = threadIdx.x; }
    double a = s, b = 1.0 + 1e-9 * threadIdx.x;
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(c[j][0], c[j][1], a, b);
    }
    double r = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) r += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void k_dmma1688(double* out, double s) {
    double c[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) c[j][q] = j + q;
    double a[4] = {s, s, s, s}, b[2] = {1.0 + 1e-9 * threadIdx.x, 0.5};
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma1688(c[j], a, b);
    }
    double r = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) r += c[j][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void k_dmma16816(double* out, double s) {
    double c[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) c[j][q] = j + q;
    double a[8] = {s, s, s, s, s, s, s, s}, b[4] = {1.0 + 1e-9 * threadIdx.x, 0.5, 0.25, 0.125};
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma16816(c[j], a, b);
    }
    double r = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) r += c[j][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// shared-memory fed DMMA: a/b fragments re-read from smem each k-step (what a real GEMM does)
__global__ void k_dmma884_smem(double* out, double s) {
    __shared__ double sa[64 * 36];
    __shared__ double sb[64 * 36];
    for (int i = threadIdx.x; i < 64 * 36; i += blockDim.x) { sa[i] = s + i * 1e-9; sb[i] = 1.0 - i * 1e-9; }
    __syncthreads();
    int lane = threadIdx.x & 31;
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = 0; c[j][1] = 0; }
    // warp tile 16 (2 m-frags) x 32 (4 n-frags), k = 32 per pass
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
            double a0 = sa[(lane >> 2) * 36 + k + (lane & 3)];
            double a1 = sa[(8 + (lane >> 2)) * 36 + k + (lane & 3)];
            double b[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) b[n] = sb[(n * 8 + (lane >> 2)) * 36 + k + (lane & 3)];
#pragma unroll
            for (int n = 0; n < 4; ++n) { dmma884(c[n][0], c[n][1], a0, b[n]); dmma884(c[4 + n][0], c[4 + n][1], a1, b[n]); }
        }
    }
    double r = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) r += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    for (int threads : {128, 256, 512, 1024}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 2048) continue;
            int grid = sms * bps;
            double warps = double(grid) * threads / 32;
            double ms;
            ms = time_ms([&] { k_dfma<<<grid, threads>>>(out, 0.999); }, 20);
            printf(", \"dfma_t%d_b%d\": %.2f", threads, bps, double(grid) * threads * ITERS * 8 * 2 / ms / 1e9);
            ms = time_ms([&] { k_dmma884<<<grid, threads>>>(out, 0.999); }, 20);
            printf(", \"dmma884_t%d_b%d\": %.2f", threads, bps, warps * ITERS * 8 * (2.0 * 8 * 8 * 4) / ms / 1e9);
            ms = time_ms([&] { k_dmma1688<<<grid, threads>>>(out, 0.999); }, 20);
            printf(", \"dmma1688_t%d_b%d\": %.2f", threads, bps, warps * ITERS * 4 * (2.0 * 16 * 8 * 8) / ms / 1e9);
            ms = time_ms([&] { k_dmma16816<<<grid, threads>>>(out, 0.999); }, 20);
            printf(", \"dmma16816_t%d_b%d\": %.2f", threads, bps, warps * ITERS * 4 * (2.0 * 16 * 8 * 16) / ms / 1e9);
            ms = time_ms([&] { k_dmma884_smem<<<grid, threads>>>(out, 0.999); }, 20);
            printf(", \"dmma884smem_t%d_b%d\": %.2f", threads, bps, warps * (ITERS / 8) * 8 * 8 * (2.0 * 8 * 8 * 4) / ms / 1e9);
        }
    }
    printf("}\n");
    return 0;
}
