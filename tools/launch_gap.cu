// Developer tool: cost of the diagonal-tile / spine kernels when launched back to back in one stream (launch gap included), against
// a trivial kernel with the same launch shape.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lcuda -o tools/bin/launch_gap tools/launch_gap.cu
#include <cstdio>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../gaussian_process_transportation_b200/csrc/factor.cuh"
using namespace gptb;
__global__ void __launch_bounds__(256, 1) empty_kernel(int* p) { if (p && threadIdx.x == 9999) *p = 1; }
template <typename F>
static float run(const char* name, int reps, F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int i = 0; i < 5; ++i) f(s);
    cudaStreamSynchronize(s);
    cudaEventRecord(a, s);
    for (int i = 0; i < reps; ++i) f(s);
    cudaEventRecord(b, s);
    cudaStreamSynchronize(s);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    printf("%-46s %8.2f us per launch (%s)\n", name, 1e3 * ms / reps, cudaGetErrorString(cudaGetLastError()));
    return ms;
}
int main() {
    const int Npad = 1024, p = 3;
    double *L, *D, *t1, *t2; int* info;
    cudaMalloc(&L, sizeof(double) * Npad * Npad); cudaMalloc(&D, sizeof(double) * Npad * 128);
    cudaMalloc(&t1, sizeof(double) * 4 * Npad); cudaMalloc(&t2, sizeof(double) * 4 * Npad); cudaMalloc(&info, 16);
    std::vector<double> hL((size_t)Npad * Npad, 0.0);
    for (int i = 0; i < Npad; ++i) for (int j = 0; j < Npad; ++j) hL[(size_t)i * Npad + j] = (i == j) ? 4.0 : 0.001 / (1 + abs(i - j));
    cudaMemset(t1, 0, sizeof(double) * 4 * Npad); cudaMemset(t2, 0, sizeof(double) * 4 * Npad); cudaMemset(info, 0, 16);
    cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM_BYTES);
    cudaFuncSetAttribute(potrf_spine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SPINE_SMEM_BYTES);
    cudaFuncSetAttribute(empty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM_BYTES);
    auto reset = [&]() { cudaMemcpy(L, hL.data(), sizeof(double) * Npad * Npad, cudaMemcpyHostToDevice); };
    reset();
    run("empty kernel, 1 CTA, no dynamic smem", 200, [&](cudaStream_t s) { empty_kernel<<<1, 256, 0, s>>>(nullptr); });
    run("empty kernel, 1 CTA, 205 KB dynamic smem", 200, [&](cudaStream_t s) { empty_kernel<<<1, 256, DIAG_SMEM_BYTES, s>>>(nullptr); });
    run("alternating 0 / 205 KB empty kernels", 200, [&](cudaStream_t s) { empty_kernel<<<1, 256, 0, s>>>(nullptr); empty_kernel<<<1, 256, DIAG_SMEM_BYTES, s>>>(nullptr); });
    // the factor of a factor is still positive definite enough: tile 0 is re-factorised in place each launch (values drift, timing does not)
    run("potrf_diag_kernel (with forward substitution)", 100, [&](cudaStream_t s) { cudaMemcpyAsync(L, hL.data(), 1, cudaMemcpyHostToDevice, s); potrf_diag_kernel<<<1, 256, DIAG_SMEM_BYTES, s>>>(L, Npad, 1, D, info, t1, t2, Npad, p); });
    reset();
    run("potrf_diag_kernel, tiles 1..6 in turn", 96, [&](cudaStream_t s) { static int k = 0; potrf_diag_kernel<<<1, 256, DIAG_SMEM_BYTES, s>>>(L, Npad, 1 + (k++ % 6), D, info, t1, t2, Npad, p); });
    reset();
    run("potrf_diag_kernel, no forward substitution", 96, [&](cudaStream_t s) { static int k = 0; potrf_diag_kernel<<<1, 256, DIAG_SMEM_BYTES, s>>>(L, Npad, 1 + (k++ % 6), D, info, nullptr, nullptr, Npad, 0); });
    reset();
    {
        long long* prof; cudaMalloc(&prof, 64 * 8); cudaMemset(prof, 0, 64 * 8);
        run("potrf_diag_kernel with phase stamps", 96, [&](cudaStream_t s) { static int k = 0; potrf_diag_kernel<<<1, 256, DIAG_SMEM_BYTES, s>>>(L, Npad, 1 + (k++ % 6), D, info, t1, t2, Npad, p, prof); });
        long long hp[64]; cudaMemcpy(hp, prof, sizeof(hp), cudaMemcpyDeviceToHost);
        printf("   stamped span %lld cycles = %.1f us wall;", hp[12] - hp[0], (hp[21] - hp[20]) / 1e3);
        for (int i = 0; i < 12; ++i) printf(" %lld", hp[i + 1] - hp[i]);
        printf("\n");
    }
    reset();
    run("potrf_spine_kernel (8 CTAs)", 100, [&](cudaStream_t s) { potrf_spine_kernel<<<SPINE_CTAS, 256, SPINE_SMEM_BYTES, s>>>(L, Npad, 2, D, t1, t2, Npad, p, info + 1, -1000000); });
    reset();
    run("diag + spine alternating", 100, [&](cudaStream_t s) { potrf_diag_kernel<<<1, 256, DIAG_SMEM_BYTES, s>>>(L, Npad, 1, D, info, t1, t2, Npad, p); potrf_spine_kernel<<<SPINE_CTAS, 256, SPINE_SMEM_BYTES, s>>>(L, Npad, 2, D, t1, t2, Npad, p, info + 1, -1000000); });
    return 0;
}
