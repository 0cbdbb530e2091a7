// Micro-benchmark of the GEMM inner loop shape: per k-group a warp loads MI A-fragments and NI B-fragments from shared
// memory (conflict-free [k/4][row][4] layout) and issues MI*NI DMMA.8x8x4.  No global traffic, optional barrier per slab.
// Purpose: find which (warp tile, warps/SM) combination lets the DMMA pipe reach its 37 TFLOP/s peak.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MI, int NI, int WM, int WN, int BAR>
__global__ void __launch_bounds__(WM * WN * 32, 1) loop_kernel(double* out, int slabs) {
    extern __shared__ double sm[];
    constexpr int ROWS_A = WM * MI * 8, ROWS_B = WN * NI * 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3, wm = warp / WN, wn = warp % WN;
    for (int i = tid; i < 8 * (ROWS_A + ROWS_B) * 4; i += blockDim.x) sm[i] = 1e-3 * (i % 17);
    __syncthreads();
    double acc[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) { acc[mi][ni][0] = 0; acc[mi][ni][1] = 0; }
    const double* sA = sm;
    const double* sB = sm + 8 * ROWS_A * 4;
    const double* pa = sA + ((wm * MI * 8 + g) << 2) + t;
    const double* pb = sB + ((wn * NI * 8 + g) << 2) + t;
    for (int s = 0; s < slabs; ++s) {
        if (BAR) __syncthreads();
#pragma unroll
        for (int kg = 0; kg < 8; ++kg) {
            double a[MI], b[NI];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) a[mi] = pa[(kg * ROWS_A + mi * 8) << 2];
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) b[ni] = pb[(kg * ROWS_B + ni * 8) << 2];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    double r = 0;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) r += acc[mi][ni][0] + acc[mi][ni][1];
    out[blockIdx.x * blockDim.x + tid] = r;
}

template <int MI, int NI, int WM, int WN, int BAR>
void run(const char* name, double* out, int sms, int ctas_per_sm) {
    constexpr int ROWS_A = WM * MI * 8, ROWS_B = WN * NI * 8;
    int smem = 8 * (ROWS_A + ROWS_B) * 4 * 8;
    auto k = loop_kernel<MI, NI, WM, WN, BAR>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int slabs = 2000;
    int blocks = sms * ctas_per_sm;
    k<<<blocks, WM * WN * 32, smem>>>(out, 10);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); k<<<blocks, WM * WN * 32, smem>>>(out, slabs); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double flops = 2.0 * 256 * MI * NI * 8.0 * slabs * (WM * WN) * (double)blocks;
    printf("%s: %.2f TFLOP/s (smem %d B, %d CTA/SM)\n", name, flops / best * 1e-9, smem, ctas_per_sm);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 4 * 1024));
    run<8, 4, 2, 4, 0>("MI8 NI4 8 warps (64x32 warp tile) nobar", out, sms, 1);
    run<8, 4, 2, 4, 1>("MI8 NI4 8 warps (64x32 warp tile) bar  ", out, sms, 1);
    run<4, 4, 4, 4, 0>("MI4 NI4 16 warps (32x32)          nobar", out, sms, 1);
    run<4, 4, 4, 4, 1>("MI4 NI4 16 warps (32x32)          bar  ", out, sms, 1);
    run<4, 8, 4, 2, 0>("MI4 NI8 8 warps (32x64)           nobar", out, sms, 1);
    run<4, 4, 2, 4, 0>("MI4 NI4 8 warps x2 CTA (64x128 CTA) nobar", out, sms, 2);
    run<8, 4, 2, 2, 0>("MI8 NI4 4 warps (128x64 CTA)      nobar", out, sms, 1);
    run<8, 4, 2, 2, 0>("MI8 NI4 4 warps x2 CTA            nobar", out, sms, 2);
    run<2, 4, 4, 4, 0>("MI2 NI4 16 warps                  nobar", out, sms, 1);
    run<8, 8, 2, 2, 0>("MI8 NI8 4 warps (64x64 warp tile) nobar", out, sms, 1);
    return 0;
}
