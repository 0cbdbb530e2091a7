// Probe: one CTA computes D(128x128, int32) = A(128xK, int8) * B(128xK, int8)^T with tcgen05.mma kind::i8,
// operands staged by TMA (128-byte swizzle, K-major), accumulator in TMEM, read back with tcgen05.ld.
// Purpose: validate descriptors / layouts for an INT8-sliced (Ozaki-style) FP64 emulation of the variance product.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(su32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(su32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(su32(dst)), "l"(m), "r"(su32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, 128-byte swizzle: 8-row groups are 1024 B apart (SBO), LBO unused; version 1 (sm_100), layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(const void* smem_ptr) {
    uint64_t d = 0;
    d |= (uint64_t)((su32(smem_ptr) & 0x3FFFF) >> 4);        // start address [0,14)
    d |= (uint64_t)1 << 16;                                  // leading byte offset (ignored for swizzled K-major) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                        // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                                  // version [46,48)
    d |= (uint64_t)2 << 61;                                  // layout type [61,64): SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}\n" ::
                 "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(su32(bar)) : "memory");
}

constexpr int BM = 128, BN = 128, BKB = 128;   // k-chunk of 128 bytes (= 128 int8 = 4 MMAs of K=32)

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB, int K, int32_t* D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                 // 16 KB
    uint8_t* sB = smem + BM * BKB;      // 16 KB
    __shared__ __align__(8) uint64_t full_bar, mma_bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&full_bar, 1); mbar_init(&mma_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(su32(&tmem_base)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_d = tmem_base;
    // instruction descriptor: c=S32 (2<<4), a=INT8 (1<<7), b=INT8 (1<<10), K-major both, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    if (tid == 0) {
        const int nchunk = K / BKB;
        for (int c = 0; c < nchunk; ++c) {
            mbar_expect_tx(&full_bar, 2 * BM * BKB);
            tma_2d(sA, &mA, c * BKB, 0, &full_bar);
            tma_2d(sB, &mB, c * BKB, 0, &full_bar);
            mbar_wait(&full_bar, c & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            for (int k = 0; k < BKB / 32; ++k) {
                uint64_t ad = make_desc(sA) + (uint64_t)((k * 32) >> 4);
                uint64_t bd = make_desc(sB) + (uint64_t)((k * 32) >> 4);
                umma_i8(tmem_d, ad, bd, idesc, (c | k) ? 1u : 0u);
            }
            umma_commit(&mma_bar);
            mbar_wait(&mma_bar, c & 1);      // serialise (probe only): smem may be overwritten next iteration
        }
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // each warp reads its 32 lanes x 128 columns, 32 columns at a time
    for (int cb = 0; cb < BN / 32; ++cb) {
        uint32_t r[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
              "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
              "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        const int row = warp * 32 + lane;
        for (int j = 0; j < 32; ++j) D[row * BN + cb * 32 + j] = (int32_t)r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(128u) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int K = 1024;
    std::vector<int8_t> A(BM * K), B(BN * K);
    srand(1);
    for (auto& v : A) v = (int8_t)(rand() % 129 - 64);
    for (auto& v : B) v = (int8_t)(rand() % 129 - 64);
    int8_t *dA, *dB; int32_t* dD;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, BM * BN * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, BM * BN * 4));
    EncodeFn encode = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
    CUtensorMap mA, mB;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)BM}; cuuint64_t strides[1] = {(cuuint64_t)K};
    cuuint32_t box[2] = {BKB, BM}; cuuint32_t es[2] = {1, 1};
    CUresult r1 = encode(&mA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(&mB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d %d\n", (int)r1, (int)r2);
    const int smem = 2 * BM * BKB + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem>>>(mA, mB, K, dD);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> D(BM * BN);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    long long bad = 0; int first = -1;
    for (int m = 0; m < BM; ++m)
        for (int n = 0; n < BN; ++n) {
            int32_t s = 0;
            for (int k = 0; k < K; ++k) s += (int32_t)A[m * K + k] * (int32_t)B[n * K + k];
            if (s != D[m * BN + n]) { if (first < 0) first = m * BN + n; ++bad; }
        }
    printf("mismatches %lld of %d (first at %d: got %d)\n", bad, BM * BN, first, first >= 0 ? D[first] : 0);
    return bad != 0;
}
