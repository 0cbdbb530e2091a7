// FP64 peak micro-benchmarks for B200 (sm_100a): DFMA vector pipe, DMMA (mma.sync f64) shapes, exp().
// Used to establish the roofline denominators that MEASURED_PEAKS.json lacks.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void dfma_kernel(double* out, int iters, double s) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    double b = s, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int NACC>
__global__ void dmma884_kernel(double* out, int iters, double s) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double a = s * (threadIdx.x & 7), b = s * (threadIdx.x & 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) r += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int NACC>
__global__ void dmma16816_kernel(double* out, int iters, double s) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; c[i][2] = 0; c[i][3] = 0; }
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = s * ((threadIdx.x + i) & 7);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = s * ((threadIdx.x + i) & 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void exp_kernel(double* out, int iters, double s) {
    double a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = -(threadIdx.x * 1e-3 + i) * s;
    double r = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { r += exp(a[i]); a[i] -= 1e-6; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 64 * 1024));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32; int blocks = sms * (warps >= 32 ? 2 : 4);
        float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 0.999); }, 5);
        double flops = 2.0 * 16 * iters * (double)threads * blocks;
        printf(", \"dfma_tflops_w%d\": %.2f", warps, flops / ms * 1e-9);
    }
    for (int warps : {4, 8, 16}) {
        int threads = warps * 32; int blocks = sms * 2;
        float ms = time_ms([&] { dmma884_kernel<16><<<blocks, threads>>>(out, iters / 4, 1e-3); }, 5);
        double flops = 2.0 * 256 * 16 * (iters / 4) * (double)warps * blocks;
        printf(", \"dmma884_tflops_w%d\": %.2f", warps, flops / ms * 1e-9);
        ms = time_ms([&] { dmma16816_kernel<8><<<blocks, threads>>>(out, iters / 16, 1e-3); }, 5);
        flops = 2.0 * 2048 * 8 * (iters / 16) * (double)warps * blocks;
        printf(", \"dmma16816_tflops_w%d\": %.2f", warps, flops / ms * 1e-9);
    }
    {
        int threads = 256, blocks = sms * 8;
        float ms = time_ms([&] { exp_kernel<<<blocks, threads>>>(out, 2000, 1.0); }, 5);
        double n = 4.0 * 2000 * (double)threads * blocks;
        printf(", \"exp_gops\": %.2f", n / ms * 1e-6);
    }
    printf("}\n");
    return 0;
}
