// Operand-feed micro-benchmark for the FP64 tile engine: identical DMMA inner loop (8 warps, 64x32 warp tiles, BK=32),
// operands streamed from global memory by (0) all-thread cp.async, (1) a producer warp issuing 2D TMA boxes of {4 k x 128 rows}
// (8 per operand per slab), (2) a producer warp issuing one 3D TMA box {4, 128, 8} per operand per slab.
// Shared layout per operand slab is always [k/4][row][k%4] (conflict-free DMMA fragment loads).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int TS = 128, BK = 32, NST = 3, SLAB = TS * BK;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* s, const void* g) {
    unsigned a = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(a), "l"(g));
}
__device__ __forceinline__ void mbar_init(uint64_t* b, int cnt) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
                 "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(m), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
                 "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(m), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void compute_slab(const double* sA, const double* sB, double (&acc)[8][4][2], int wm, int wn, int g, int t) {
    const double* pa = sA + ((wm * 64 + g) << 2) + t;
    const double* pb = sB + ((wn * 32 + g) << 2) + t;
#pragma unroll
    for (int kg = 0; kg < BK / 4; ++kg) {
        double a[8], b[4];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) a[mi] = pa[(kg * TS + mi * 8) << 2];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) b[ni] = pb[(kg * TS + ni * 8) << 2];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
}

// mode 0: all 256 threads cp.async (current engine)
__global__ void __launch_bounds__(256, 1) k_cpasync(const double* A, const double* B, long long ld, int nslab, double* out) {
    extern __shared__ __align__(128) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3, wm = warp >> 2, wn = warp & 3;
    const double* Ab = A + (long long)(blockIdx.x % 64) * TS * ld;
    const double* Bb = B + (long long)(blockIdx.x % 7) * TS * ld;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0; acc[i][j][1] = 0; }
    auto load = [&](int st, int s) {
        const int row = tid >> 1, half = (tid & 1) << 1;
        const double* ga = Ab + (long long)row * ld + (long long)s * BK + half;
        const double* gb = Bb + (long long)row * ld + (long long)s * BK + half;
        double* sA = smem + st * 2 * SLAB; double* sB = sA + SLAB;
        const int soff = (row << 2) + half;
#pragma unroll
        for (int kg = 0; kg < BK / 4; ++kg) { cp_async16(sA + soff + kg * TS * 4, ga + kg * 4); cp_async16(sB + soff + kg * TS * 4, gb + kg * 4); }
    };
    for (int st = 0; st < NST - 1; ++st) { load(st, st); asm volatile("cp.async.commit_group;\n"); }
    for (int it = 0; it < nslab; ++it) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(NST - 2));
        __syncthreads();
        if (it + NST - 1 < nslab) load((it + NST - 1) % NST, it + NST - 1);
        asm volatile("cp.async.commit_group;\n");
        const double* sA = smem + (it % NST) * 2 * SLAB;
        compute_slab(sA, sA + SLAB, acc, wm, wn, g, t);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) r += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * 256 + tid] = r;
}

// mode 1/2: warp 8 = TMA producer, warps 0..7 consumers; full/empty mbarrier ring
template <int MODE>
__global__ void __launch_bounds__(288, 1) k_tma(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB, int nslab, double* out) {
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t full[NST], empty[NST];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int rowA = (blockIdx.x % 64) * TS, rowB = (blockIdx.x % 7) * TS;
    if (warp == 8) {
        if (lane == 0) {
            for (int it = 0; it < nslab; ++it) {
                const int st = it % NST;
                if (it >= NST) mbar_wait(&empty[st], ((it / NST) - 1) & 1);
                double* sA = smem + st * 2 * SLAB; double* sB = sA + SLAB;
                mbar_expect_tx(&full[st], 2 * SLAB * 8);
                if (MODE == 1) {
#pragma unroll
                    for (int kg = 0; kg < BK / 4; ++kg) {
                        tma_2d(sA + kg * TS * 4, &mA, it * BK + kg * 4, rowA, &full[st]);
                        tma_2d(sB + kg * TS * 4, &mB, it * BK + kg * 4, rowB, &full[st]);
                    }
                } else {
                    tma_3d(sA, &mA, 0, rowA, it * (BK / 4), &full[st]);
                    tma_3d(sB, &mB, 0, rowB, it * (BK / 4), &full[st]);
                }
            }
        }
        return;
    }
    const int g = lane >> 2, t = lane & 3, wm = warp >> 2, wn = warp & 3;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0; acc[i][j][1] = 0; }
    for (int it = 0; it < nslab; ++it) {
        const int st = it % NST;
        mbar_wait(&full[st], (it / NST) & 1);
        const double* sA = smem + st * 2 * SLAB;
        compute_slab(sA, sA + SLAB, acc, wm, wn, g, t);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) r += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * 256 + tid] = r;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const long long rows = 64 * TS, K = 16384, ld = K;
    double *A, *B, *out;
    CK(cudaMalloc(&A, rows * ld * 8)); CK(cudaMalloc(&B, rows * ld * 8)); CK(cudaMalloc(&out, sms * 256 * 8));
    CK(cudaMemset(A, 0, rows * ld * 8)); CK(cudaMemset(B, 0, rows * ld * 8));
    const int nslab = (int)(K / BK);
    const int smem = NST * 2 * SLAB * 8;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double flops = 2.0 * TS * TS * (double)K * sms;
    auto report = [&](const char* name, auto launch) {
        launch(); CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
        printf("%s: %.2f TFLOP/s (%.3f ms)\n", name, flops / best * 1e-9, best);
    };
    CK(cudaFuncSetAttribute(k_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    report("cp.async all threads      ", [&] { k_cpasync<<<sms, 256, smem>>>(A, B, ld, nslab, out); });

    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 0; }
    {
        CUtensorMap mA, mB;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
        cuuint32_t box[2] = {4, TS};
        cuuint32_t es[2] = {1, 1};
        CUresult r1 = encode(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, A, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = encode(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, B, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode 2D: %d %d\n", (int)r1, (int)r2);
        if (r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS) {
            CK(cudaFuncSetAttribute(k_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            report("TMA 2D boxes {4,128} x8    ", [&] { k_tma<1><<<sms, 288, smem>>>(mA, mB, nslab, out); });
        }
    }
    {
        CUtensorMap mA, mB;
        cuuint64_t dims[3] = {4, (cuuint64_t)rows, (cuuint64_t)K / 4};
        cuuint64_t strides[2] = {(cuuint64_t)ld * 8, 32};
        cuuint32_t box[3] = {4, TS, BK / 4};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r1 = encode(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, A, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = encode(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, B, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode 3D (4, rows, K/4): %d %d\n", (int)r1, (int)r2);
        if (r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS) {
            CK(cudaFuncSetAttribute(k_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            report("TMA 3D box {4,128,8}       ", [&] { k_tma<2><<<sms, 288, smem>>>(mA, mB, nslab, out); });
        }
    }
    return 0;
}
