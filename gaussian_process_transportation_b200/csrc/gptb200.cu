// libgptb200.so -- C ABI (include/gptb200.h) over the sm_100a kernels in gemm_engine.cuh / factor.cuh / query.cuh.
// Host code here is orchestration only: allocation, launch order, chunking of query batches, status mapping.
#include "../../include/gptb200.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>
#include <numeric>

#include <cub/device/device_radix_sort.cuh>   // library plumbing: radix sort of the query batch by Morton key (spatial mode)

#include "query.cuh"
#include "ozaki.cuh"

using namespace gptb;

namespace {

struct EvPair { cudaEvent_t a, b; };

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encoder() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 3-D TMA view (4, rows, K/4) of a row-major FP64 matrix (rows x K, leading dimension ld): box (4, 128, BK/4) lands in
// shared memory as [k/4][row][k%4], the conflict-free fragment layout of gemm_engine.cuh.
bool make_operand_map(CUtensorMap* m, const double* base, long long rows, long long K, long long ld, int box_rows = TS) {
    EncodeTiledFn enc = get_encoder();
    if (!enc) return false;
    cuuint64_t dims[3] = {4, (cuuint64_t)rows, (cuuint64_t)(K / 4)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8, 32};
    cuuint32_t box[3] = {4, (cuuint32_t)box_rows, (cuuint32_t)(BK / 4)};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 3-D TMA view (k bytes, rows, plane) of int8 digit planes [S][rows][ncols]; box (64, box_rows, box_planes), 64-byte swizzle
// (the K-major shared-memory layout tcgen05.mma expects).
bool make_plane_map(CUtensorMap* m, const int8_t* base, long long rows, long long ncols, int S, int box_rows, int box_planes, int box_k) {
    EncodeTiledFn enc = get_encoder();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)ncols, (cuuint64_t)rows, (cuuint64_t)S};
    cuuint64_t strides[2] = {(cuuint64_t)ncols, (cuuint64_t)rows * (cuuint64_t)ncols};
    cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               box_k == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// one view per box depth 1 .. S (oz::PlaneMaps): a chunk with leading zero planes loads only the planes it needs.
// box_k = 64: 64-byte k-chunks, 64-byte swizzle (dense product kernel); 128: 128-byte chunks, 128-byte swizzle (skipping kernel)
bool make_plane_maps(oz::PlaneMaps* pm, const int8_t* base, long long rows, long long ncols, int S, int box_rows, int box_k) {
    for (int n = 1; n <= S; ++n)
        if (!make_plane_map(&pm->m[n - 1], base, rows, ncols, S, box_rows, n, box_k)) return false;
    for (int n = S; n < 7; ++n) pm->m[n] = pm->m[S - 1];
    return true;
}

}  // namespace

struct gptb_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t aux = nullptr;               // high-priority stream for the look-ahead diagonal tile
    cudaStream_t gen = nullptr;               // lowest-priority stream: the generator of batch i+1 runs under the int8 products of batch i
    cudaEvent_t ev_gen[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_start = nullptr;
    // spatial mode (gptb_set_spatial): training points in Morton order, query batches sorted by Morton key, zero digit planes
    // skipped by the INT8-sliced product kernel
    int spatial = 0;
    int spatial_shuffle = 1;                  // chunk order shuffled (gptb_set_debug_option "spatial_shuffle" 0 keeps the plain Z-order: A/B only)
    int oz_force_skip = 0;                    // debug option "oz_force_skip_variant": dense case through the skipping loops (A/B of the two loop versions)
    int oz_whatif = 0;                        // debug option "oz_whatif" (only acts in a -DGPTB_OZ_WHATIF build)
    unsigned long long* oz_prof = nullptr;    // 64 cycle counters of CTA 0's roles (GPTB_OZ_WHATIF builds; gptb_debug_read_profile)
    unsigned long long exec_pairs_host = 0;   // plane-pair x chunk products issued by the dense product kernel (counted on the host)
    std::vector<int> perm;                    // perm[i] = caller's index of internal training row i (empty: natural order)
    bool perm_known = true;                   // false on a handle whose state arrived by broadcast (exports need the permutation)
    MortonBox mbox{};
    unsigned* flagsB = nullptr;               // [Npad/64][flags_stride] packed block masks of the digit planes of L^-1
    int flags_stride = 0;                     // 32-bit words per mask row
    int pipeline = 0;                         // 1 overlaps the generator of batch i+1 with the products of batch i (opt-in: measured no gain, the
                                              // int8 products run at the 1 kW power cap, so the two kernels share one energy budget)
    std::vector<cudaEvent_t> ev_diag, ev_col, ev_spine, ev_panel; // per-step dependencies between the factorisation's streams
    cudaStream_t head = nullptr;             // the two trailing tiles the next spine step needs (factorize_device)
    std::string err;
    long long N = 0, Npad = 0;
    int T = 0, d = 0, p = 0;
    double *X = nullptr, *Y = nullptr, *Xs = nullptr, *alpha = nullptr, *tmp1 = nullptr, *tmp2 = nullptr;
    double *Lbuf = nullptr, *Dinv = nullptr, *Minv = nullptr, *Wbuf = nullptr, *gradpart = nullptr, *scal = nullptr;
    double* header = nullptr;
    int* info = nullptr;                     // [0] LAPACK info, [1] product-kernel tile queue, [2..3] trailing queue + leavers, [4] spine barrier, [5..6] head queue
    int* chain_flags = nullptr;              // trsv_back_chain_kernel: flag[k] == chain_epoch once x_k is published
    int chain_epoch = 0, chain_cap = 0;
    double* splitk_ws = nullptr;             // partial tiles of the split-k variance product (small batches), splitk_cap = two per SM
    int splitk_cap = 0;
    int splitk_on = 1;                       // developer A/B ("variance_splitk")
    int back_variant = 1;                    // 1 = one chained launch, 0 = one launch per block (developer A/B)
    int spine_variant = 1;                   // 1 = diag -> spine kernel -> diag on the aux stream, 0 = round-1 schedule (developer A/B)
    CUtensorMap mapL, mapD, mapM, mapW;      // TMA views of Lbuf, Dinv, Minv, Wbuf
    CUtensorMap mapL64;                      // Lbuf with a 64-row box (half-tile trailing update)
    // INT8-sliced variance path (ozaki.cuh): digit planes of the inverse factor + per-row scales
    int var_mode = 0, var_slices = 6, var_bits = 7;   // var_mode 1 = INT8-sliced; var_bits = digit width (7 balanced | 8 full range)
    // run-time accuracy guard of the INT8-sliced path (gptb_set_variance_guard): what the caller asked for, what the probe decided
    int var_mode_req = 0, var_slices_req = 6;
    int var_extra = 0, var_extra_req = 0;     // 8-bit planes, S = 5: also form the first dropped diagonal a + b = S (19 products; ozaki.cuh)
    double guard_thresh = 2.0e-8;             // max |std_int8 - std_fp64| / sqrt(c + s2) on the probe set (a fifth of the 1e-7 tolerance); 0 = off
    bool guard_done = false, guard_busy = false;
    double guard_first_err = -1.0, guard_err = -1.0;
    double* probe = nullptr;                  // probe points + the two std vectors + the reduced error
    int trailing_variant = 1;                // 0: 128x128 tiles, one CTA/SM; 1: 128x64 half tiles, two CTAs/SM
    int8_t* Bplanes = nullptr;
    double* scaleB = nullptr;
    oz::PlaneMaps mapsBq;
    bool have_bplanes = false;
    KParams kp{};
    Affine af{};
    bool have_train = false, have_factor = false, have_alpha = false, have_minv = false, have_kinv = false;
    // query workspace
    double* ws = nullptr;
    size_t ws_bytes = 0;
    long long ws_limit = 16LL << 30;
    long long batch_cap = 524288;             // queries per batch (debug option "batch_cap"), where the workspace limit allows: c3 12.84 / 13.05 / 13.25 / 13.64 M q/s at 65536 / 131072 / 262144 / 524288 (fewer sort / finalize / launch tails per query; tools/batch_cap_ab.py)
    // staging for the host-pointer query: two device buffer sets, slices of HOST_SLICE queries; H2D of slice i+1 (s_h2d) and D2H of
    // slice i-1 (s_d2h) run under the kernels of slice i (main stream)
    double* stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    // scratch of the small epilogue entry points (joint covariance, orientation, stiffness): one buffer carved per call, kept
    // across calls up to SCRATCH_KEEP bytes
    char* scratch = nullptr;
    size_t scratch_bytes = 0;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_cd[2] = {nullptr, nullptr}, ev_oc[2] = {nullptr, nullptr};
    long long launches = 0;
    bool timing = false;
    std::vector<EvPair> ev[4];
};

#define GPTB_FAIL(h, code, ...)                               \
    do {                                                      \
        char _b[512];                                         \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                \
        (h)->err = _b;                                        \
        return (code);                                        \
    } while (0)

#define CU(h, call)                                                                                   \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) GPTB_FAIL(h, -2, "CUDA error %s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define MAKE_MAP(h, m, base, rows, K, ld)                                                              \
    do {                                                                                               \
        if (!make_operand_map((m), (base), (rows), (K), (ld))) GPTB_FAIL(h, -5, "cuTensorMapEncodeTiled failed at %s:%d", __FILE__, __LINE__); \
    } while (0)

#define LAUNCH_CHECK(h)                                                                               \
    do {                                                                                              \
        (h)->launches++;                                                                              \
        cudaError_t _e = cudaGetLastError();                                                          \
        if (_e != cudaSuccess) GPTB_FAIL(h, -3, "kernel launch failed: %s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

static void tic(gptb_handle* h, int cls, cudaStream_t st = nullptr) {
    if (!h->timing) return;
    EvPair e;
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    cudaEventRecord(e.a, st ? st : h->stream);
    h->ev[cls].push_back(e);
}
static void toc(gptb_handle* h, int cls, cudaStream_t st = nullptr) {
    if (!h->timing) return;
    cudaEventRecord(h->ev[cls].back().b, st ? st : h->stream);
}

static void free_model(gptb_handle* h) {
    double** ptrs[] = {&h->X, &h->Y, &h->Xs, &h->alpha, &h->tmp1, &h->tmp2, &h->Lbuf, &h->Dinv, &h->Minv, &h->Wbuf, &h->gradpart};
    for (auto pp : ptrs) {
        if (*pp) cudaFree(*pp);
        *pp = nullptr;
    }
    if (h->Bplanes) cudaFree(h->Bplanes);
    if (h->scaleB) cudaFree(h->scaleB);
    if (h->flagsB) cudaFree(h->flagsB);
    h->Bplanes = nullptr;
    h->scaleB = nullptr;
    h->flagsB = nullptr;
    h->have_train = h->have_factor = h->have_alpha = h->have_minv = h->have_kinv = h->have_bplanes = false;
}

// a new model (refit, new hyper-parameters, received state): digit planes are stale, the guard starts again from the requested mode
static void invalidate_planes(gptb_handle* h) {
    h->have_bplanes = false;
    h->guard_done = false;
    h->var_mode = h->var_mode_req;
    h->var_slices = h->var_slices_req;
    h->var_extra = h->var_extra_req;
}

static int set_kernel_attrs(gptb_handle* h) {
    CU(h, cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(potrf_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(potrf_trailing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(potrf_trailing64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(trtri_level_p1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(trtri_level_p2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(kinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(trmm_sumsq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(trmm_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(trsv_back_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BACKCHAIN_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(potrf_spine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SPINE_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(oz::ozaki_trmm_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<4>::SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(oz::ozaki_trmm_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<4>::SMEM_SKIP_BYTES));
    CU(h, cudaFuncSetAttribute(oz::ozaki_trmm_kernel<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<5>::SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(oz::ozaki_trmm_kernel<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<5>::SMEM_SKIP_BYTES));
    CU(h, cudaFuncSetAttribute(oz::ozaki_trmm_kernel<5, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<5>::SMEM_SKIP_BYTES));
    CU(h, cudaFuncSetAttribute(oz::ozaki_trmm_kernel<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<6>::SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(oz::ozaki_trmm_kernel<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<6>::SMEM_SKIP_BYTES));
    CU(h, cudaFuncSetAttribute(oz::ozaki_trmm_kernel<7, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<7>::SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(oz::ozaki_trmm_kernel<7, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<7>::SMEM_SKIP_BYTES));
    CU(h, cudaFuncSetAttribute(trmm_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(cov_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(cov_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(cov_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU(h, cudaFuncSetAttribute(cov_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    return 0;
}

extern "C" int gptb_version(void) { return 100; }

extern "C" int gptb_create(int device, gptb_handle** out) {
    if (!out) return -1;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return -4;   // no CUDA device: there is no fallback
    if (device < 0 || device >= ndev) return -1;
    gptb_handle* h = new gptb_handle();
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return -2; }
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    // numerically lower = higher priority: aux (look-ahead tile) > main > gen (the generator only fills what the products leave free)
    const int prio_main = (prio_lo - 1 >= prio_hi) ? prio_lo - 1 : prio_lo;
    if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_main) != cudaSuccess) { delete h; return -2; }
    if (cudaStreamCreateWithPriority(&h->aux, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { delete h; return -2; }
    if (cudaStreamCreateWithPriority(&h->gen, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { delete h; return -2; }
    if (cudaStreamCreateWithPriority(&h->head, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { delete h; return -2; }
    for (int i = 0; i < 2; ++i)
        if (cudaEventCreateWithFlags(&h->ev_gen[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming) != cudaSuccess) { delete h; return -2; }
    if (cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming) != cudaSuccess) { delete h; return -2; }
    if (cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking) != cudaSuccess) { delete h; return -2; }
    for (int i = 0; i < 2; ++i)
        if (cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_cd[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_oc[i], cudaEventDisableTiming) != cudaSuccess) { delete h; return -2; }
    if (cudaMalloc(&h->info, 8 * sizeof(int)) != cudaSuccess || cudaMalloc(&h->scal, 64 * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->header, 32 * sizeof(double)) != cudaSuccess) {
        delete h;
        return -2;
    }
    if (cudaMemset(h->scal, 0, 64 * sizeof(double)) != cudaSuccess) { delete h; return -2; }
    int rc = set_kernel_attrs(h);
    if (rc) { delete h; return rc; }
    h->af.on = 0;
    h->af.s = 1.0;
    for (int i = 0; i < MAXD; ++i)
        for (int j = 0; j < MAXD; ++j) h->af.R[i][j] = (i == j) ? 1.0 : 0.0;
    *out = h;
    return 0;
}

extern "C" void gptb_destroy(gptb_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_model(h);
    if (h->ws) cudaFree(h->ws);
    for (int i = 0; i < 2; ++i) {
        if (h->stage[i]) cudaFree(h->stage[i]);
        cudaEventDestroy(h->ev_in[i]); cudaEventDestroy(h->ev_cd[i]); cudaEventDestroy(h->ev_oc[i]);
    }
    cudaStreamSynchronize(h->s_h2d); cudaStreamSynchronize(h->s_d2h);
    cudaStreamDestroy(h->s_h2d); cudaStreamDestroy(h->s_d2h);
    if (h->oz_prof) cudaFree(h->oz_prof);
    if (h->scratch) cudaFree(h->scratch);
    if (h->probe) cudaFree(h->probe);
    cudaFree(h->info);
    cudaFree(h->chain_flags);
    cudaFree(h->splitk_ws);
    cudaFree(h->scal);
    cudaFree(h->header);
    for (auto& v : h->ev)
        for (auto& e : v) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto e : h->ev_diag) cudaEventDestroy(e);
    for (auto e : h->ev_col) cudaEventDestroy(e);
    for (auto e : h->ev_spine) cudaEventDestroy(e);
    for (auto e : h->ev_panel) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(h->ev_gen[i]); cudaEventDestroy(h->ev_done[i]); }
    cudaEventDestroy(h->ev_start);
    cudaStreamSynchronize(h->gen);
    cudaStreamDestroy(h->gen);
    cudaStreamDestroy(h->head);
    cudaStreamDestroy(h->aux);
    cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char* gptb_last_error(gptb_handle* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int64_t gptb_launch_count(gptb_handle* h) { return h ? h->launches : 0; }
extern "C" void* gptb_stream(gptb_handle* h) { return h ? (void*)h->stream : nullptr; }
extern "C" int gptb_set_trailing_variant(gptb_handle* h, int variant) {
    if (!h || variant < 0 || variant > 1) return -1;
    h->trailing_variant = variant;
    return 0;
}
extern "C" int gptb_set_spatial(gptb_handle* h, int on) {
    if (!h) return -1;
    if ((on != 0) != (h->spatial != 0) && h->have_train)
        GPTB_FAIL(h, -1, "gptb_set_spatial must be called before gptb_set_train (the training order is fixed there)");
    if (on < 0 || on > 2) GPTB_FAIL(h, -1, "gptb_set_spatial: mode must be 0, 1 or 2");
    h->spatial = on;          // 2 = Morton order and sorted batches, but every plane product issued (A/B of the skipping itself)
    return 0;
}
extern "C" int gptb_set_debug_option(gptb_handle* h, const char* name, int value) {
    if (!h || !name) return -1;
    if (!strcmp(name, "spatial_shuffle")) h->spatial_shuffle = value != 0;
    else if (!strcmp(name, "oz_force_skip_variant")) h->oz_force_skip = value != 0;
    else if (!strcmp(name, "oz_whatif")) h->oz_whatif = value;
    else if (!strcmp(name, "back_substitution_variant")) h->back_variant = value != 0;
    else if (!strcmp(name, "spine_variant")) h->spine_variant = value != 0;
    else if (!strcmp(name, "variance_splitk")) h->splitk_on = value != 0;
    else if (!strcmp(name, "batch_cap")) { if (value < 128 || value % 128) GPTB_FAIL(h, -1, "batch_cap must be a multiple of 128"); h->batch_cap = value; }
    else GPTB_FAIL(h, -1, "gptb_set_debug_option: unknown option '%s'", name);
    return 0;
}
// device counter of the skipping product kernel + the host-side count of the dense one
static unsigned long long* exec_counter(gptb_handle* h) { return reinterpret_cast<unsigned long long*>(h->scal + 40); }
extern "C" int gptb_executed_products(gptb_handle* h, int64_t* pairs, int reset) {
    if (!h || !pairs) return -1;
    CU(h, cudaSetDevice(h->device));
    unsigned long long dev = 0;
    CU(h, cudaMemcpyAsync(&dev, exec_counter(h), sizeof(dev), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    *pairs = (int64_t)(dev + h->exec_pairs_host);
    if (reset) {
        CU(h, cudaMemsetAsync(exec_counter(h), 0, sizeof(dev), h->stream));
        h->exec_pairs_host = 0;
    }
    return 0;
}
extern "C" int gptb_debug_read_profile(gptb_handle* h, int64_t* out64) {
    if (!h || !out64) return -1;
    CU(h, cudaSetDevice(h->device));
    if (!h->oz_prof) {
        CU(h, cudaMalloc(&h->oz_prof, 64 * sizeof(unsigned long long)));
        CU(h, cudaMemset(h->oz_prof, 0, 64 * sizeof(unsigned long long)));
    }
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaMemcpy(out64, h->oz_prof, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int gptb_set_query_pipeline(gptb_handle* h, int on) {
    if (!h) return -1;
    h->pipeline = on != 0;
    return 0;
}
extern "C" int gptb_set_workspace_limit(gptb_handle* h, int64_t bytes) {
    if (!h || bytes < (1 << 20)) return -1;
    h->ws_limit = bytes;
    return 0;
}
extern "C" int gptb_timing_enable(gptb_handle* h, int on) { if (!h) return -1; h->timing = on != 0; return 0; }
extern "C" int gptb_timing_reset(gptb_handle* h) {
    if (!h) return -1;
    cudaStreamSynchronize(h->stream);
    for (auto& v : h->ev) {
        for (auto& e : v) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
        v.clear();
    }
    return 0;
}
extern "C" int gptb_kernel_time(gptb_handle* h, int which, double* ms, int64_t* n) {
    if (!h || which < 0 || which > 3) return -1;
    CU(h, cudaStreamSynchronize(h->stream));
    double tot = 0.0;
    for (auto& e : h->ev[which]) {
        float t = 0.f;
        CU(h, cudaEventElapsedTime(&t, e.a, e.b));
        tot += t;
    }
    if (ms) *ms = tot;
    if (n) *n = (int64_t)h->ev[which].size();
    return 0;
}

static int alloc_model(gptb_handle* h, long long N, int d, int p, bool train) {
    if (N < 1 || d < 1 || d > MAXD || p < 1 || p > MAXP) GPTB_FAIL(h, -1, "unsupported shape N=%lld d=%d p=%d (need 1<=d<=%d, 1<=p<=%d)", N, d, p, MAXD, MAXP);
    CU(h, cudaSetDevice(h->device));
    long long Npad = (N + TS - 1) / TS * TS;
    if (Npad != h->Npad || d != h->d || p != h->p || (train && !h->Lbuf)) {
        CU(h, cudaStreamSynchronize(h->stream));
        free_model(h);
        h->Npad = Npad; h->T = (int)(Npad / TS); h->d = d; h->p = p;
        CU(h, cudaMalloc(&h->X, sizeof(double) * d * Npad));
        CU(h, cudaMalloc(&h->Xs, sizeof(double) * d * Npad));
        CU(h, cudaMalloc(&h->alpha, sizeof(double) * p * Npad));
        if (train) {
            CU(h, cudaMalloc(&h->Y, sizeof(double) * p * Npad));
            CU(h, cudaMalloc(&h->tmp1, sizeof(double) * p * Npad));
            CU(h, cudaMalloc(&h->tmp2, sizeof(double) * p * Npad));
            CU(h, cudaMalloc(&h->Lbuf, sizeof(double) * Npad * Npad));
            CU(h, cudaMalloc(&h->Dinv, sizeof(double) * Npad * TS));
            MAKE_MAP(h, &h->mapL, h->Lbuf, Npad, Npad, Npad);
            if (!make_operand_map(&h->mapL64, h->Lbuf, Npad, Npad, Npad, H_BN)) GPTB_FAIL(h, -5, "cuTensorMapEncodeTiled failed (64-row box)");
            MAKE_MAP(h, &h->mapD, h->Dinv, Npad, TS, TS);
        }
    }
    h->N = N;
    h->have_factor = h->have_alpha = h->have_minv = h->have_kinv = false;
    invalidate_planes(h);
    return 0;
}

extern "C" int gptb_set_variance_mode(gptb_handle* h, int mode, int slices) {
    if (!h) return -1;
    if (mode < 0 || mode > 3) GPTB_FAIL(h, -1, "unknown variance mode %d", mode);
    if (mode == 3 && slices != 5) GPTB_FAIL(h, -1, "the extra-diagonal variant exists for five 8-bit digit planes only (got %d)", slices);
    const int extra = (mode == 3) ? 1 : 0;
    if (mode == 3) mode = 2;
    if (mode == 1 && (slices < 5 || slices > 7)) GPTB_FAIL(h, -1, "the INT8-sliced variance path supports 5, 6 or 7 seven-bit digit planes (got %d)", slices);
    if (mode == 2 && (slices < 4 || slices > 6)) GPTB_FAIL(h, -1, "the INT8-sliced variance path supports 4, 5 or 6 eight-bit digit planes (got %d)", slices);
    const int bits = (mode == 2) ? 8 : 7;
    const bool same = (mode >= 1) == (h->var_mode == 1) && (mode == 0 || (slices == h->var_slices && bits == h->var_bits && extra == h->var_extra));
    h->var_mode_req = mode >= 1 ? 1 : 0;
    h->var_extra_req = extra;
    if (mode >= 1) { h->var_slices_req = slices; h->var_bits = bits; }
    if (!same || mode == 0) invalidate_planes(h);
    return 0;
}

extern "C" int gptb_set_kernel_kind(gptb_handle* h, int kind) {
    if (!h) return -1;
    if (kind < 0 || kind > 3) GPTB_FAIL(h, -1, "unknown kernel kind %d", kind);
    if (kind != h->kp.kind) { h->have_factor = h->have_alpha = h->have_minv = h->have_kinv = false; invalidate_planes(h); }
    h->kp.kind = kind;
    return 0;
}

extern "C" int gptb_set_train(gptb_handle* h, const double* X, const double* Y, int64_t N, int d, int p) {
    if (!h || !X || !Y) return -1;
    int rc = alloc_model(h, N, d, p, true);
    if (rc) return rc;
    const long long Npad = h->Npad;
    std::vector<double> xs((size_t)d * Npad, 0.0), ys((size_t)p * Npad, 0.0);
    h->perm.clear();
    h->perm_known = true;
    if (h->spatial) {
        // Morton (Z-order) permutation of the training points over their bounding box: neighbours in space become neighbours
        // in the factor, so 64-point chunks are compact and k(x*, chunk) is uniformly tiny for far-away query tiles.  Internal
        // order only: gptb_export_alpha / gptb_export_Kinv undo it; the posterior itself does not depend on the order.
        MortonBox& mb = h->mbox;
        mb.bits = (d <= 3) ? 10 : 8;
        const double top = (double)((1u << mb.bits) - 1u);
        for (int a = 0; a < MAXD; ++a) { mb.lo[a] = 0.0; mb.scale[a] = 0.0; }
        for (int a = 0; a < d; ++a) {
            double lo = X[a], hi = X[a];
            for (long long n = 1; n < N; ++n) { lo = std::min(lo, X[n * d + a]); hi = std::max(hi, X[n * d + a]); }
            mb.lo[a] = lo;
            mb.scale[a] = (hi > lo) ? top / (hi - lo) : 0.0;
        }
        std::vector<unsigned> code((size_t)N);
        for (long long n = 0; n < N; ++n) {
            unsigned cd = 0;
            for (int a = 0; a < d; ++a) {
                double t = (X[n * d + a] - mb.lo[a]) * mb.scale[a];
                t = (t > 0.0) ? ((t < top) ? t : top) : 0.0;
                const unsigned cell = (unsigned)t;
                for (int b = 0; b < mb.bits; ++b) cd |= ((cell >> b) & 1u) << (b * d + a);
            }
            code[(size_t)n] = cd;
        }
        h->perm.resize((size_t)N);
        std::iota(h->perm.begin(), h->perm.end(), 0);
        std::stable_sort(h->perm.begin(), h->perm.end(), [&](int i, int j) { return code[(size_t)i] < code[(size_t)j]; });
        // The skipping works per 64-point chunk, so only the chunks have to be compact -- their ORDER is free.  A fully
        // Z-ordered factor puts every point right behind its neighbours: the conditional variances L_ii shrink, the row
        // maxima of L^-1 (= the digit scales) grow exactly on the rows that carry the weight of near-training-point
        // queries, and the 40-bit products lose a decimal (std error 1.2e-7 at N = 32768).  Shuffling the chunk order
        // (fixed seed) keeps the compactness and gives back most of the conditioning of a random order.
        if (h->spatial_shuffle) {
            const long long nchunk = (N + 63) / 64;
            std::vector<long long> order((size_t)nchunk);
            std::iota(order.begin(), order.end(), 0LL);
            unsigned long long st = 0x9E3779B97F4A7C15ULL;
            for (long long i = nchunk - 1; i > 0; --i) {          // Fisher-Yates with a splitmix64 stream (deterministic)
                st += 0x9E3779B97F4A7C15ULL;
                unsigned long long z = st;
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
                z ^= z >> 31;
                std::swap(order[(size_t)i], order[(size_t)(z % (unsigned long long)(i + 1))]);
            }
            std::vector<int> shuffled;
            shuffled.reserve((size_t)N);
            for (long long cidx : order)
                for (long long n = cidx * 64; n < std::min<long long>(N, (cidx + 1) * 64); ++n) shuffled.push_back(h->perm[(size_t)n]);
            h->perm.swap(shuffled);
        }
    }
    for (long long n = 0; n < N; ++n) {
        const long long src = h->perm.empty() ? n : h->perm[(size_t)n];
        for (int a = 0; a < d; ++a) xs[(size_t)a * Npad + n] = X[src * d + a];
        for (int o = 0; o < p; ++o) ys[(size_t)o * Npad + n] = Y[src * p + o];
    }
    CU(h, cudaMemcpyAsync(h->X, xs.data(), sizeof(double) * d * Npad, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->Y, ys.data(), sizeof(double) * p * Npad, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->have_train = true;
    return 0;
}

static int set_params(gptb_handle* h, double c, const double* ell, double s2, double jitter) {
    if (!(c > 0.0) || !(s2 >= 0.0) || !(jitter >= 0.0)) GPTB_FAIL(h, -1, "invalid hyper-parameters c=%g s2=%g jitter=%g", c, s2, jitter);
    h->kp.c = c; h->kp.s2 = s2; h->kp.jitter = jitter;
    for (int a = 0; a < MAXD; ++a) {
        double e = (a < h->d) ? ell[a] : 1.0;
        if (!(e > 0.0)) GPTB_FAIL(h, -1, "invalid length-scale %g", e);
        h->kp.ell[a] = e;
        h->kp.inv_ell[a] = 1.0 / e;
    }
    return 0;
}

template <typename F>
static void dispatch_d(int d, F f) {
    switch (d) {
        case 1: f(std::integral_constant<int, 1>{}); break;
        case 2: f(std::integral_constant<int, 2>{}); break;
        case 3: f(std::integral_constant<int, 3>{}); break;
        default: f(std::integral_constant<int, 4>{}); break;
    }
}

static int launch_scale(gptb_handle* h) {
    scale_inputs_kernel<<<(unsigned)((h->Npad + 255) / 256), 256, 0, h->stream>>>(h->X, h->Xs, (int)h->N, (int)h->Npad, h->d, h->kp);
    LAUNCH_CHECK(h);
    return 0;
}

// Gram + blocked right-looking Cholesky with one step of look-ahead, forward substitution fused in.  Two schedules:
//
// spine_variant 1 (default): the serial chain lives on the aux stream as small kernels,
//   aux  stream : diag(k) -> [wait head(k-1)] -> spine(k) = tiles (k+1,k) and (k+1,k+1), eight CTAs, cooperative -> diag(k+1) -> ...
//   main stream : [wait diag(k)] panel(k), rows >= k+2 -> [wait spine(k)] trailing(k): every tile of columns >= k+1 but (k+1,k+1)
//                 and the two head tiles
//   head stream : [wait panel(k), spine(k)] tiles (k+2,k+1) and (k+2,k+2) of trailing(k) -- all the next spine step reads
// so one step of the chain is diag + spine (~42 + ~14 us) and the wide kernels only have to keep up with it.
//
// spine_variant 0 (round 1): the chain is diag(k+1) <- look-ahead column k+1 <- panel(k) <- diag(k), three launches of which two are
// full 128^3 tiles on one SM each (~17 us of DMMA time apiece):
//   main stream : panel(k) -> trailing column k+1 -> [event] -> rest of trailing(k)            (wide kernels)
//   aux  stream : [wait column event] -> diagonal tile k+1 (factor + inverse + z_{k+1}) -> [event]   (one CTA)
// In both, the persistent trailing kernel leaves one SM free for the diagonal-tile CTA.
// tiles [first, first + count) of the row-major lower-triangle enumeration with origin (base, base); count < 0: to the end
static int launch_trailing(gptb_handle* h, cudaStream_t st, int kt, int base, int first, int count, int nsm, int* counter) {
    const long long ld = h->Npad;
    const int r2 = h->T - base;
    const int njobs = count >= 0 ? count : r2 * (r2 + 1) / 2 - first;
    if (njobs <= 0) return 0;
    if (h->trailing_variant == 0) {
        const int grid = njobs < nsm - 1 ? njobs : nsm - 1;
        tic(h, 2, st);
        potrf_trailing_kernel<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(h->mapL, h->Lbuf, ld, kt, kt + 1, base, 0, njobs, first);
        toc(h, 2, st);
    } else {
        // half tiles, two CTAs per SM, dynamic queue.  The spine stream needs SMs of its own while this kernel runs: one for the diagonal
        // tile, eight for the spine kernel (which cannot share an SM with two of these CTAs).  A long update (njobs large) hides the
        // spine step anyway -- it can wait for this kernel to drain -- so only the diagonal tile's SM is kept free; a short one leaves
        // eight.  CTAs that land on a reserved SM exit at once (the queue hands their jobs to the others), so the grid carries spares.
        const int njobs2 = 2 * njobs;
        // The head launch (count >= 0) is itself part of the spine's chain: it takes no part in this and may use the reserved SMs.
        const int reserve = (h->spine_variant == 1 && njobs <= 640) ? SPINE_CTAS : 1;
        const int workers = 2 * (nsm - reserve);
        int grid, reserved_from;
        if (count < 0 && (njobs2 > workers || reserve > 1)) {
            grid = (njobs2 < workers ? njobs2 : workers) + 2 * reserve;
            reserved_from = nsm - reserve;
        } else {
            grid = njobs2;
            reserved_from = -1;
        }
        CU(h, cudaMemsetAsync(counter, 0, 2 * sizeof(int), st));            // [0] job queue, [1] CTAs that left a reserved SM
        tic(h, 2, st);
        potrf_trailing64_kernel<<<grid, H_THREADS, H_SMEM_BYTES, st>>>(h->mapL, h->mapL64, h->Lbuf, ld, kt, kt + 1, base, 0, njobs2, counter, reserved_from,
                                                                     2 * first);
        toc(h, 2, st);
    }
    LAUNCH_CHECK(h);
    return 0;
}

static int factorize_device(gptb_handle* h) {
    const int T = h->T;
    const long long ld = h->Npad;
    const int Npad = (int)h->Npad, p = h->p;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
    while ((int)h->ev_diag.size() < T + 1) {
        cudaEvent_t a, b, c, d;
        CU(h, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CU(h, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        CU(h, cudaEventCreateWithFlags(&c, cudaEventDisableTiming));
        CU(h, cudaEventCreateWithFlags(&d, cudaEventDisableTiming));
        h->ev_diag.push_back(a);
        h->ev_col.push_back(b);
        h->ev_spine.push_back(c);
        h->ev_panel.push_back(d);
    }
    CU(h, cudaMemsetAsync(h->info, 0, sizeof(int), h->stream));            // LAPACK info
    CU(h, cudaMemsetAsync(h->info + 4, 0, sizeof(int), h->stream));        // arrivals at the spine kernel's barrier
    CU(h, cudaMemcpyAsync(h->tmp1, h->Y, sizeof(double) * p * h->Npad, cudaMemcpyDeviceToDevice, h->stream));
    int rc = launch_scale(h);
    if (rc) return rc;
    const unsigned ntri = (unsigned)((long long)T * (T + 1) / 2);
    dispatch_d(h->d, [&](auto D) {
        gram_lower_kernel<decltype(D)::value><<<ntri, 256, 0, h->stream>>>(h->Xs, h->Lbuf, (int)h->N, (int)h->Npad, h->kp);
    });
    LAUNCH_CHECK(h);
    CU(h, cudaEventRecord(h->ev_col[T], h->stream));                 // "column 0 is ready"
    CU(h, cudaStreamWaitEvent(h->aux, h->ev_col[T], 0));
    potrf_diag_kernel<<<1, 256, DIAG_SMEM_BYTES, h->aux>>>(h->Lbuf, ld, 0, h->Dinv, h->info, h->tmp1, h->tmp2, Npad, p);
    LAUNCH_CHECK(h);
    CU(h, cudaEventRecord(h->ev_diag[0], h->aux));
    if (h->spine_variant == 1) {
        for (int kt = 0; kt + 1 < T; ++kt) {
            const int r = T - kt - 1;
            if (kt > 0) CU(h, cudaStreamWaitEvent(h->aux, h->ev_col[kt - 1], 0));          // head(kt-1) has updated tiles (kt+1, kt) and (kt+1, kt+1)
            {
                // cooperative launch: the eight CTAs meet at barriers, so they start only when all of them fit on the machine at once
                // (with several handles factorising on one GPU, half-started spine kernels must not be able to hold every SM)
                double* a_L = h->Lbuf; long long a_ld = ld; int a_kt = kt; const double* a_D = h->Dinv; double* a_rhs = h->tmp1; const double* a_sol = h->tmp2;
                int a_Npad = Npad, a_p = p; int* a_arr = h->info + 4; int a_target = 2 * SPINE_CTAS * (kt + 1);
                void* args[] = {&a_L, &a_ld, &a_kt, &a_D, &a_rhs, &a_sol, &a_Npad, &a_p, &a_arr, &a_target};
                CU(h, cudaLaunchCooperativeKernel((const void*)potrf_spine_kernel, dim3(SPINE_CTAS), dim3(256), args, SPINE_SMEM_BYTES, h->aux));
            }
            CU(h, cudaEventRecord(h->ev_spine[kt], h->aux));
            potrf_diag_kernel<<<1, 256, DIAG_SMEM_BYTES, h->aux>>>(h->Lbuf, ld, kt + 1, h->Dinv, h->info, h->tmp1, h->tmp2, Npad, p);
            LAUNCH_CHECK(h);
            CU(h, cudaEventRecord(h->ev_diag[kt + 1], h->aux));
            if (r > 1) {
                CU(h, cudaStreamWaitEvent(h->stream, h->ev_diag[kt], 0));
                potrf_panel_kernel<<<r - 1, GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(h->mapL, h->mapD, h->Lbuf, ld, kt, h->tmp1, h->tmp2, Npad, p, kt + 2);
                LAUNCH_CHECK(h);
                CU(h, cudaEventRecord(h->ev_panel[kt], h->stream));
                // head: tiles (kt+2, kt+1) and (kt+2, kt+2) -- jobs 1 and 2 of the triangle at (kt+1, kt+1) -- are all the next spine step
                // needs from this trailing update; they run on their own stream beside the rest
                CU(h, cudaStreamWaitEvent(h->head, h->ev_panel[kt], 0));
                CU(h, cudaStreamWaitEvent(h->head, h->ev_spine[kt], 0));
                if ((rc = launch_trailing(h, h->head, kt, kt + 1, 1, 2, nsm, h->info + 5))) return rc;
                CU(h, cudaEventRecord(h->ev_col[kt], h->head));
                CU(h, cudaStreamWaitEvent(h->stream, h->ev_spine[kt], 0));
                if ((rc = launch_trailing(h, h->stream, kt, kt + 1, 3, -1, nsm, h->info + 2))) return rc;
            }
        }
        CU(h, cudaStreamWaitEvent(h->stream, h->ev_diag[T - 1], 0));
        if (T > 2) CU(h, cudaStreamWaitEvent(h->stream, h->ev_col[T - 3], 0));             // the last head launch
    } else {
        for (int kt = 0; kt < T; ++kt) {
            CU(h, cudaStreamWaitEvent(h->stream, h->ev_diag[kt], 0));
            const int r = T - kt - 1;
            if (r <= 0) break;
            potrf_panel_kernel<<<r, GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(h->mapL, h->mapD, h->Lbuf, ld, kt, h->tmp1, h->tmp2, Npad, p, kt + 1);
            LAUNCH_CHECK(h);
            // look-ahead column: tiles (i, kt+1), i >= kt+1
            potrf_trailing_kernel<<<r, GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(h->mapL, h->Lbuf, ld, kt, kt + 1, kt + 1, 1, r, 0);
            LAUNCH_CHECK(h);
            CU(h, cudaEventRecord(h->ev_col[kt], h->stream));
            CU(h, cudaStreamWaitEvent(h->aux, h->ev_col[kt], 0));
            potrf_diag_kernel<<<1, 256, DIAG_SMEM_BYTES, h->aux>>>(h->Lbuf, ld, kt + 1, h->Dinv, h->info, h->tmp1, h->tmp2, Npad, p);
            LAUNCH_CHECK(h);
            CU(h, cudaEventRecord(h->ev_diag[kt + 1], h->aux));
            if (r > 1 && (rc = launch_trailing(h, h->stream, kt, kt + 2, 0, -1, nsm, h->info + 2))) return rc;
        }
    }
    int info = 0;
    CU(h, cudaMemcpyAsync(&info, h->info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (info > 0) {
        h->have_factor = false;
        GPTB_FAIL(h, info, "the kernel matrix is not positive definite: leading minor of order %d", info);
    }
    return 0;
}

// alpha = L^-T z, z = L^-1 Y already produced block by block inside factorize_device (tmp2).
static int solve_alpha(gptb_handle* h) {
    const int T = h->T;
    const long long ld = h->Npad;
    if (h->back_variant == 1) {
        if (T > h->chain_cap) {
            cudaFree(h->chain_flags);
            h->chain_flags = nullptr;
            h->chain_cap = 0;
            CU(h, cudaMalloc(&h->chain_flags, sizeof(int) * (size_t)(T + 64)));
            CU(h, cudaMemsetAsync(h->chain_flags, 0, sizeof(int) * (size_t)(T + 64), h->stream));
            h->chain_cap = T + 64;
            h->chain_epoch = 0;
        }
        h->chain_epoch += 1;
        trsv_back_chain_kernel<<<T, 256, BACKCHAIN_SMEM_BYTES, h->stream>>>(h->Lbuf, ld, h->Dinv, h->tmp2, h->alpha, (int)h->Npad, h->p, T, h->chain_flags,
                                                                          h->chain_epoch);
        LAUNCH_CHECK(h);
        return 0;
    }
    for (int kt = T - 1; kt >= 0; --kt) {
        trsv_back_step_kernel<<<kt > 0 ? kt : 1, 256, 0, h->stream>>>(h->Lbuf, ld, h->Dinv, h->tmp2, h->alpha, (int)h->Npad, h->p, kt);
        LAUNCH_CHECK(h);
    }
    return 0;
}

static int lml_value(gptb_handle* h, double* lml) {
    lml_terms_kernel<<<1, 1024, 0, h->stream>>>(h->Y, h->alpha, h->Lbuf, h->Npad, (int)h->N, (int)h->Npad, h->p, h->scal);
    LAUNCH_CHECK(h);
    double t[2];
    CU(h, cudaMemcpyAsync(t, h->scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    *lml = -0.5 * t[0] - (double)h->p * t[1] - (double)h->p * (double)h->N * 0.5 * std::log(2.0 * M_PI);
    return 0;
}

extern "C" int gptb_factorize(gptb_handle* h, double c, const double* ell, double s2, double jitter, double* lml) {
    if (!h || !ell) return -1;
    if (!h->have_train) GPTB_FAIL(h, -1, "gptb_factorize: no training data (call gptb_set_train)");
    CU(h, cudaSetDevice(h->device));
    int rc = set_params(h, c, ell, s2, jitter);
    if (rc) return rc;
    h->have_factor = h->have_alpha = h->have_minv = h->have_kinv = false;
    invalidate_planes(h);
    if ((rc = factorize_device(h))) return rc;
    if ((rc = solve_alpha(h))) return rc;
    h->have_factor = h->have_alpha = true;
    if (lml) {
        if ((rc = lml_value(h, lml))) return rc;
    } else {
        CU(h, cudaStreamSynchronize(h->stream));
    }
    return 0;
}

static int ensure_wbuf(gptb_handle* h) {
    if (!h->Wbuf) {
        CU(h, cudaMalloc(&h->Wbuf, sizeof(double) * h->Npad * h->Npad));
        MAKE_MAP(h, &h->mapW, h->Wbuf, h->Npad, h->Npad, h->Npad);
    }
    return 0;
}

static int build_minv(gptb_handle* h) {
    if (!h->splitk_ws) {                     // also for a model that arrived by gptb_state_commit (have_minv set, nothing built here)
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
        CU(h, cudaMalloc(&h->splitk_ws, sizeof(double) * TS * TS * (size_t)(2 * nsm)));
        h->splitk_cap = 2 * nsm;
    }
    if (h->have_minv) return 0;
    if (!h->have_factor) GPTB_FAIL(h, -1, "no factorisation available");
    const int T = h->T;
    const long long ld = h->Npad;
    if (!h->Minv) {
        CU(h, cudaMalloc(&h->Minv, sizeof(double) * h->Npad * h->Npad));
        MAKE_MAP(h, &h->mapM, h->Minv, h->Npad, h->Npad, h->Npad);
    }
    int rc = ensure_wbuf(h);
    if (rc) return rc;
    h->have_kinv = false;
    h->have_bplanes = false;
    trtri_init_kernel<<<T, 256, 0, h->stream>>>(h->Minv, ld, h->Dinv);
    LAUNCH_CHECK(h);
    for (int s = 1; s < T; s *= 2) {
        int pairs = (T + 2 * s - 1) / (2 * s);
        unsigned grid = (unsigned)((long long)pairs * s * s);
        trtri_level_p1_kernel<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(h->mapL, h->mapM, h->Wbuf, ld, T, s);
        LAUNCH_CHECK(h);
        trtri_level_p2_kernel<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(h->mapM, h->mapW, h->Minv, ld, T, s);
        LAUNCH_CHECK(h);
    }
    h->have_minv = true;
    return 0;
}

template <typename F>
static void dispatch_slices(int S, F f) {
    if (S == 4) f(std::integral_constant<int, 4>{});
    else if (S == 5) f(std::integral_constant<int, 5>{});
    else if (S == 6) f(std::integral_constant<int, 6>{});
    else f(std::integral_constant<int, 7>{});
}
// (planes, digit width) pairs with an instantiated slicer / fused generator: 5..7 x 7-bit, 4..6 x 8-bit
template <typename F>
static void dispatch_digits(int S, int bits, F f) {
    using std::integral_constant;
    if (bits == 8) {
        if (S == 4) f(integral_constant<int, 4>{}, integral_constant<int, 8>{});
        else if (S == 5) f(integral_constant<int, 5>{}, integral_constant<int, 8>{});
        else f(integral_constant<int, 6>{}, integral_constant<int, 8>{});
    } else {
        if (S == 5) f(integral_constant<int, 5>{}, integral_constant<int, 7>{});
        else if (S == 6) f(integral_constant<int, 6>{}, integral_constant<int, 7>{});
        else f(integral_constant<int, 7>{}, integral_constant<int, 7>{});
    }
}

// digit planes of the inverse factor for the INT8-sliced variance path
static int build_bplanes(gptb_handle* h) {
    if (h->have_bplanes && h->have_minv) return 0;      // a refit invalidates the inverse factor and with it the planes
    h->have_bplanes = false;
    int rc = build_minv(h);
    if (rc) return rc;
    const long long Npad = h->Npad;
    const int S = h->var_slices;
    // exactness of the int32 accumulators: a diagonal sums up to S digit-plane products of K terms, each |digit product| <= 2^12
    // (2^14 for 8-bit digits).  When that worst-case bound fails for 8-bit digits the slicer's data-dependent bound decides:
    // |sum_k a_k b_k| <= 128 * sum_k |b_k| over the actual digits of each inverse-factor row (most of them are far below 128).
    const bool static_ok = (long long)S * Npad * (h->var_bits == 8 ? 16384 : 4096) <= 2147483647LL;
    if (!static_ok && h->var_bits != 8)
        GPTB_FAIL(h, -1, "INT8-sliced variance path: N=%lld with %d 7-bit digit planes could overflow the int32 accumulators", (long long)h->N, S);
    if (h->Bplanes) { cudaFree(h->Bplanes); h->Bplanes = nullptr; }
    CU(h, cudaMalloc(&h->Bplanes, (size_t)S * Npad * Npad));
    if (!h->scaleB) CU(h, cudaMalloc(&h->scaleB, sizeof(double) * Npad));
    unsigned long long* l1max = reinterpret_cast<unsigned long long*>(h->scal + 32);
    CU(h, cudaMemsetAsync(l1max, 0, sizeof(unsigned long long), h->stream));
    // block masks of the non-zero digit planes (used by the product kernel in spatial mode; 8-bit planes only)
    h->flags_stride = (int)(((Npad / 64 + 3) / 4 + 3) / 4 * 4);     // 32-bit words per mask row, padded to 16 bytes (one uint4 per lane in the product kernel)
    if (h->flagsB) { cudaFree(h->flagsB); h->flagsB = nullptr; }
    if (h->var_bits == 8) {
        if (h->flags_stride > FLAG_WORDS) GPTB_FAIL(h, -1, "INT8-sliced variance path: N=%lld exceeds the %d-chunk limit of the block masks", (long long)h->N, FLAG_WORDS * 4);
        CU(h, cudaMalloc(&h->flagsB, sizeof(unsigned) * (size_t)(Npad / 64) * h->flags_stride));
        CU(h, cudaMemsetAsync(h->flagsB, 0, sizeof(unsigned) * (size_t)(Npad / 64) * h->flags_stride, h->stream));
    }
    dispatch_digits(S, h->var_bits, [&](auto SS, auto BB) {
        oz::slice_rows_kernel<decltype(SS)::value, decltype(BB)::value><<<(unsigned)Npad, 256, 0, h->stream>>>(h->Minv, Npad, Npad, (int)Npad, 1, h->Bplanes, Npad * Npad, h->scaleB, l1max,
                                                                                                               h->flagsB, h->flags_stride);
    });
    LAUNCH_CHECK(h);
    if (!static_ok) {
        unsigned long long l1 = 0;
        CU(h, cudaMemcpyAsync(&l1, l1max, sizeof(l1), cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        if (l1 * 128ULL > 2147483647ULL)
            GPTB_FAIL(h, -1, "INT8-sliced variance path: N=%lld with %d 8-bit digit planes could overflow the int32 accumulators "
                             "(largest row sum of |digits| = %llu); use the 7-bit planes (mode 1)", (long long)h->N, S, l1);
    }
    if (!make_plane_maps(&h->mapsBq, h->Bplanes, Npad, Npad, S, oz::ON, oz::OKB))
        GPTB_FAIL(h, -5, "cuTensorMapEncodeTiled failed for the digit planes");
    h->have_bplanes = true;
    return 0;
}

static int build_kinv(gptb_handle* h) {
    if (h->have_kinv) return 0;
    int rc = build_minv(h);
    if (rc) return rc;
    const int T = h->T;
    kinv_kernel<<<(unsigned)((long long)T * (T + 1) / 2), GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(h->mapM, h->Wbuf, h->Npad, T);
    LAUNCH_CHECK(h);
    h->have_kinv = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Run-time accuracy guard of the INT8-sliced variance path.  The digit scales are a-priori bounds (k* <= c, |dk*/dx| <= c/ell,
// row maxima of L^-1), so how many of the S x 8 bits are significant depends on the data: strongly correlated models (large c/s2,
// long length-scales, a Z-ordered factor) lose bits exactly where the variance nearly cancels.  Instead of trusting a fixed S, the
// engine measures: PROBE_N probe queries -- half of them a hair next to training points (where the std is most sensitive), half within
// a few length-scales of them -- are evaluated on the INT8 path and on the FP64 DMMA path of the same handle.  If the std differs by
// more than guard_thresh (default 2e-8 = a fifth of the 1e-7 tolerance, relative to sqrt(c + s2)) one more digit plane is
// added and the probe repeated; when the plane count is exhausted the handle falls back to the FP64 path for this model.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PROBE_N = 2048;

__global__ void probe_points_kernel(const double* __restrict__ X, int N, int Npad, int d, KParams kp, double* __restrict__ xq) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= PROBE_N) return;
    const int n = (int)(((long long)q * N) / PROBE_N);                  // evenly spread over the (possibly Morton-ordered) training set
    unsigned long long st = 0x9E3779B97F4A7C15ULL * (unsigned long long)(q + 1);
    for (int a = 0; a < d; ++a) {
        st += 0x9E3779B97F4A7C15ULL;
        unsigned long long z = st;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        const double u = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;     // [-1, 1)
        const double reach = (q & 1) ? ((q & 2) ? 1e-2 : 1e-3) : ((q & 2) ? 3.0 : 0.5);    // in length-scales: next to the point (x2) | nearby | around it
        xq[(long long)q * d + a] = X[(long long)a * Npad + n] + reach * kp.ell[a] * u;
    }
}

__global__ void probe_error_kernel(const double* __restrict__ a, const double* __restrict__ b, int n, int stride, double* __restrict__ out) {
    __shared__ double red[256];
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double e = fabs(a[(long long)i * stride] - b[(long long)i * stride]);
        m = (e == e) ? fmax(m, e) : 1e300;                                          // a NaN on either path counts as a failure
    }
    red[threadIdx.x] = m;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

static int build_bplanes(gptb_handle* h);

static int run_variance_guard(gptb_handle* h) {
    if (h->guard_done || h->guard_busy || h->var_mode != 1) return 0;
    if (!(h->guard_thresh > 0.0)) { h->guard_done = true; h->guard_first_err = h->guard_err = -1.0; return 0; }
    const int d = h->d, p = h->p;
    if (!h->probe) CU(h, cudaMalloc(&h->probe, sizeof(double) * ((size_t)PROBE_N * (d + 2 * p) + 8)));
    double* xq = h->probe;
    double* s8 = xq + (size_t)PROBE_N * d;
    double* s64 = s8 + (size_t)PROBE_N * p;
    double* err_dev = s64 + (size_t)PROBE_N * p;
    h->guard_busy = true;
    h->guard_first_err = -1.0;
    int rc = 0;
    probe_points_kernel<<<(PROBE_N + 255) / 256, 256, 0, h->stream>>>(h->X, (int)h->N, (int)h->Npad, d, h->kp, xq);
    h->launches++;
    bool have64 = false;
    for (;;) {
        if ((rc = gptb_query_dev(h, xq, PROBE_N, GPTB_STD, nullptr, nullptr, s8, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr))) break;
        if (!have64) {
            h->var_mode = 0;
            rc = gptb_query_dev(h, xq, PROBE_N, GPTB_STD, nullptr, nullptr, s64, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
            h->var_mode = 1;
            if (rc) break;
            have64 = true;
        }
        probe_error_kernel<<<1, 256, 0, h->stream>>>(s8, s64, PROBE_N, p, err_dev);
        h->launches++;
        double err = 0.0;
        if (cudaMemcpyAsync(&err, err_dev, sizeof(double), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
            cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = -2; h->err = "CUDA error in the variance guard"; break; }
        err /= std::sqrt(h->kp.c + h->kp.s2);
        if (h->guard_first_err < 0.0) h->guard_first_err = err;
        h->guard_err = err;
        if (err <= h->guard_thresh) break;
        const int smax = (h->var_bits == 8) ? 6 : 7;
        if (h->var_bits == 8 && h->var_slices == 5 && !h->var_extra) {
            h->var_extra = 1;                   // first the dropped diagonal a + b = S: 4 more products on the SAME planes (no rebuild)
            continue;
        }
        if (h->var_slices < smax) {
            h->var_slices += 1;                 // one more digit plane per operand
            h->var_extra = 0;
            h->have_bplanes = false;
            if ((rc = build_bplanes(h))) break;
            continue;
        }
        h->var_mode = 0;                        // plane count exhausted: this model is served by the FP64 DMMA path
        h->guard_err = 0.0;
        break;
    }
    h->guard_busy = false;
    h->guard_done = (rc == 0);
    return rc;
}

extern "C" int gptb_set_variance_guard(gptb_handle* h, double threshold) {
    if (!h || !(threshold >= 0.0)) return -1;
    h->guard_thresh = threshold;
    invalidate_planes(h);
    return 0;
}

extern "C" int gptb_variance_guard_report(gptb_handle* h, int* requested_slices, int* used_slices, int* used_extra_diagonal, double* probe_err,
                                          double* first_err, double* threshold) {
    if (!h) return -1;
    if (requested_slices) *requested_slices = h->var_mode_req ? h->var_slices_req : 0;
    if (used_slices) *used_slices = h->var_mode ? h->var_slices : 0;
    if (used_extra_diagonal) *used_extra_diagonal = (h->var_mode && h->var_slices == 5 && h->var_bits == 8) ? h->var_extra : 0;
    if (probe_err) *probe_err = h->guard_done ? h->guard_err : -1.0;
    if (first_err) *first_err = h->guard_done ? h->guard_first_err : -1.0;
    if (threshold) *threshold = h->guard_thresh;
    return 0;
}

extern "C" int gptb_prepare_variance(gptb_handle* h) {
    if (!h) return -1;
    CU(h, cudaSetDevice(h->device));
    int rc = (h->var_mode == 1) ? build_bplanes(h) : build_minv(h);
    if (rc) return rc;
    if ((rc = run_variance_guard(h))) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int gptb_lml(gptb_handle* h, double c, const double* ell, double s2, double jitter, int want_grad, double* lml, double* grad) {
    if (!h || !lml) return -1;
    int rc = gptb_factorize(h, c, ell, s2, jitter, lml);
    if (rc) return rc;
    if (!want_grad) return 0;
    if (!grad) return -1;
    if ((rc = build_kinv(h))) return rc;
    const int T = h->T;
    const long long ntri = (long long)T * (T + 1) / 2;
    if (!h->gradpart) CU(h, cudaMalloc(&h->gradpart, sizeof(double) * ntri * (2 + MAXD)));
    dispatch_d(h->d, [&](auto D) {
        lml_grad_kernel<decltype(D)::value><<<(unsigned)ntri, 256, 0, h->stream>>>(h->Xs, h->alpha, h->Wbuf, (int)h->N, (int)h->Npad, h->p, h->kp, h->gradpart);
    });
    LAUNCH_CHECK(h);
    lml_grad_reduce_kernel<<<1, 256, 0, h->stream>>>(h->gradpart, ntri, h->d, h->kp, h->scal + 8);
    LAUNCH_CHECK(h);
    double g[2 + MAXD];
    CU(h, cudaMemcpyAsync(g, h->scal + 8, sizeof(double) * (2 + h->d), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 2 + h->d; ++i) grad[i] = g[i];
    return 0;
}

extern "C" int gptb_set_affine(gptb_handle* h, const double* R, double s, const double* Sbar, const double* Tbar) {
    if (!h) return -1;
    if (!R) {   // reset to identity
        h->af.on = 0;
        h->af.s = 1.0;
        for (int i = 0; i < MAXD; ++i)
            for (int j = 0; j < MAXD; ++j) h->af.R[i][j] = (i == j) ? 1.0 : 0.0;
        return 0;
    }
    if (h->d < 1) GPTB_FAIL(h, -1, "gptb_set_affine before the model shape is known");
    if (!Sbar || !Tbar) return -1;
    h->af.on = 1;
    h->af.s = s;
    for (int i = 0; i < MAXD; ++i) {
        h->af.Sbar[i] = i < h->d ? Sbar[i] : 0.0;
        h->af.Tbar[i] = i < h->d ? Tbar[i] : 0.0;
        for (int j = 0; j < MAXD; ++j) h->af.R[i][j] = (i < h->d && j < h->d) ? R[i * h->d + j] : (i == j ? 1.0 : 0.0);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// queries
// ---------------------------------------------------------------------------------------------------------------
static bool oz_fused_supported(int d, int p) { return d == p && (d == 2 || d == 3); }

// scratch of the spatial mode for one batch: Morton keys / batch positions (radix-sort double buffers) and the block masks
struct SpatialWs {
    unsigned *keys_in, *keys_out, *vals_in;      // used by the generator stream only: shared by the two pipeline slots
    unsigned* vals_out[2];                       // qperm of the batch in slot b (read by the generator AND by finalize)
    void* cub_tmp;
    size_t cub_bytes;
    unsigned* flagsA[2];                         // block masks of the batch in slot b (written by the generator, read by the products)
};

// Split-k plan of the FP64 variance product for small batches (query.cuh: trmm_splitk_kernel).  In units of one 128^3 tile product the
// one-CTA-per-tile kernel lasts max(T, W / #SM) with W = rowtiles * T(T+1)/2 the whole work, the split one max(KC, W / #SM) plus the
// reduction: split when the longest k-loop, not the work, sets the time, with the chunk KC = ceil(W / #SM) (or larger if the
// partial-tile buffer, two tiles per SM, cannot hold that many jobs).  KC = 0: do not split.
static void trmm_splitk_plan(gptb_handle* h, int rowtiles, int T, int* KC, int* J) {
    *KC = 0;
    *J = 0;
    if (!h->splitk_ws || !h->splitk_on || T < 2) return;
    const int nsm = h->splitk_cap / 2;
    const long long work = (long long)rowtiles * T * (T + 1) / 2;
    int kc = (int)((work + nsm - 1) / nsm);
    if (kc < 1) kc = 1;
    for (; 4 * kc <= 3 * T; ++kc) {                  // a chunk above 3/4 of the longest loop is not worth the second kernel
        int j = 0;
        for (int t = 0; t < T; ++t) j += (t + kc) / kc;
        if ((long long)rowtiles * j <= h->splitk_cap) {
            *KC = kc;
            *J = j;
            return;
        }
    }
}

template <int D, int P>
static int query_chunk(gptb_handle* h, const double* x_dev, const double* vel_dev, int B, int Bpad, unsigned flags, int nrhs,
                       unsigned genflags, const QueryOut& out, long long q_off, long long Mtot, double* rhs, double* part,
                       double* macc, double* xr, int nsplit, const CUtensorMap* mapR, void* oz_planes, double* oz_scale, int pipe_slot,
                       const SpatialWs* sp) {
    const int T = h->T;
    // pipe_slot >= 0: this batch's generator runs on the low-priority stream into buffer set `pipe_slot`; the products and
    // the epilogue follow on the main stream once it has finished, so the NEXT batch's generator (FP64 pipe) overlaps them
    // (tensor pipe).  The caller orders buffer reuse through ev_done[].
    cudaStream_t gs = (pipe_slot >= 0) ? h->gen : h->stream;
    const long long rows_total = (long long)nrhs * Bpad;
    Affine af = h->af;
    af.on = (flags & GPTB_AFFINE_IN) ? h->af.on : 0;
    dim3 grid(Bpad / QPB, nsplit);
    tic(h, 1, gs);
    // digit planes straight from the generator when the INT8-sliced path is on and this (d,p) has a fused instantiation
    const bool fused = (nrhs > 0 && h->var_mode == 1 && oz_fused_supported(D, P));
    int8_t* Aplanes = reinterpret_cast<int8_t*>(oz_planes);
    const bool spatial = fused && sp != nullptr;
    const int slot = pipe_slot >= 0 ? pipe_slot : 0;
    unsigned* const vals_out_slot = spatial ? sp->vals_out[slot] : nullptr;
    unsigned* const flagsA_slot = spatial ? sp->flagsA[slot] : nullptr;
    if (fused) {
        DigitScales ds{};
        if (spatial) {
            // Morton keys of the (transformed) queries -> radix sort -> qperm; block masks start from zero
            morton_keys_kernel<D><<<(unsigned)((B + 255) / 256), 256, 0, gs>>>(x_dev, B, af, h->mbox, sp->keys_in, sp->vals_in);
            LAUNCH_CHECK(h);
            size_t tmp_bytes = sp->cub_bytes;
            CU(h, cub::DeviceRadixSort::SortPairs(sp->cub_tmp, tmp_bytes, sp->keys_in, sp->keys_out, sp->vals_in, vals_out_slot, B, 0, h->mbox.bits * D, gs));
            CU(h, cudaMemsetAsync(flagsA_slot, 0, sizeof(unsigned) * (size_t)(rows_total / 128) * h->flags_stride, gs));
            ds.qperm = vals_out_slot;
            ds.flags = flagsA_slot;
            ds.flags_stride = h->flags_stride;
        }
        auto set = [&](int idx, double bound) {
            if (h->var_bits == 8) {                     // tightest scale; digits8_pack4 wants 256^S / scale
                ds.scale[idx] = digit_scale8(bound);
                ds.down[idx] = std::ldexp(1.0, 8 * h->var_slices) / ds.scale[idx];
            } else {
                const int ex = digit_scale_exp(bound, h->var_bits);
                ds.down[idx] = std::ldexp(1.0, -ex);
                ds.scale[idx] = std::ldexp(1.0, ex);
            }
        };
        set(0, h->kp.c);                                               // k* <= c
        for (int a = 0; a < D; ++a) {
            set(1 + a, h->kp.c * h->kp.inv_ell[a]);                    // |dk*/dx_a| = k |dx_a|/ell_a^2 <= c/ell_a (r e^{-r^2/2} < 1)
            set(1 + D + a, h->kp.c * (1.0 + h->kp.inv_ell[a]));
        }
        if constexpr (D == P && (D == 2 || D == 3)) {
            dispatch_digits(h->var_slices, h->var_bits, [&](auto SS, auto BB) {
                kstar_kernel<D, P, 2, decltype(SS)::value, decltype(BB)::value><<<grid, 256, 0, gs>>>(x_dev, h->Xs, h->alpha, (int)h->N, (int)h->Npad, B, Bpad, h->kp, af, genflags,
                                                                                       nullptr, xr, macc, nsplit, Aplanes, rows_total * h->Npad, oz_scale, ds);
            });
        }
    } else if (genflags) {
        kstar_kernel<D, P, 1, 5><<<grid, 256, 0, h->stream>>>(x_dev, h->Xs, h->alpha, (int)h->N, (int)h->Npad, B, Bpad, h->kp, af, genflags, rhs, xr, macc, nsplit,
                                                             nullptr, 0, nullptr, DigitScales{});
    } else {
        kstar_kernel<D, P, 0, 5><<<grid, 256, 0, h->stream>>>(x_dev, h->Xs, h->alpha, (int)h->N, (int)h->Npad, B, Bpad, h->kp, af, 0u, rhs, xr, macc, nsplit, nullptr,
                                                             0, nullptr, DigitScales{});
    }
    toc(h, 1, gs);
    LAUNCH_CHECK(h);
    if (pipe_slot >= 0) {
        CU(h, cudaEventRecord(h->ev_gen[pipe_slot], gs));
        CU(h, cudaStreamWaitEvent(h->stream, h->ev_gen[pipe_slot], 0));
    }
    int Tpart = T;
    if (nrhs > 0 && h->var_mode == 1) {
        // INT8-sliced path: split the right-hand-side rows into digit planes, then S(S+1)/2 exact int8 GEMMs per tile
        const int S = h->var_slices;
        const int rowtiles = (int)(rows_total / TS);
        const int T64 = (int)(h->Npad / oz::ON);
        oz::PlaneMaps mapsAq;
        const bool use_masks = spatial && h->spatial == 1;
        const bool extra = h->var_extra && S == 5 && h->var_bits == 8;        // S = 5 + the first dropped diagonal (skipping kernel only)
        // spatial mode 2 (every plane product issued: the A/B of the skipping itself) runs the same kernel with null masks, so that the two
        // differ in nothing but the skipped all-zero products
        const bool skipping = use_masks || h->oz_force_skip || extra || spatial;
        if (!make_plane_maps(&mapsAq, Aplanes, rows_total, h->Npad, S, oz::OM, oz::OKB)) GPTB_FAIL(h, -5, "cuTensorMapEncodeTiled failed for the digit planes");
        if (!fused) {
            tic(h, 3);
            dispatch_digits(S, h->var_bits, [&](auto SS, auto BB) {
                oz::slice_rows_kernel<decltype(SS)::value, decltype(BB)::value><<<(unsigned)rows_total, 256, 0, h->stream>>>(rhs, h->Npad, rows_total, (int)h->Npad, 0, Aplanes, rows_total * h->Npad, oz_scale);
            });
            toc(h, 3);
            h->launches++;
        }
        tic(h, 0);
        dispatch_slices(S, [&](auto SS) {
            constexpr int SV = decltype(SS)::value;
            const long long ntiles = (long long)rowtiles * T64;
            int nsm = 148;
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
            cudaMemsetAsync(h->info + 1, 0, sizeof(int), h->stream);      // dynamic tile counter
            const unsigned grid_oz = (unsigned)(ntiles < nsm ? ntiles : nsm);
            if (skipping && extra) {
                if constexpr (SV == 5)
                    oz::ozaki_trmm_kernel<5, true, 1><<<grid_oz, oz::OTHREADS_SKIP, oz::Cfg<5>::SMEM_SKIP_BYTES, h->stream>>>(
                        mapsAq, h->mapsBq, oz_scale, h->scaleB, T64, rowtiles, rows_total, part, h->info + 1, h->var_bits, use_masks ? flagsA_slot : nullptr, use_masks ? h->flagsB : nullptr, h->flags_stride,
                        exec_counter(h), h->oz_whatif, h->oz_prof);
            } else if (skipping) {
                oz::ozaki_trmm_kernel<SV, true><<<grid_oz, oz::OTHREADS_SKIP, oz::Cfg<SV>::SMEM_SKIP_BYTES, h->stream>>>(
                    mapsAq, h->mapsBq, oz_scale, h->scaleB, T64, rowtiles, rows_total, part, h->info + 1, h->var_bits, use_masks ? flagsA_slot : nullptr, use_masks ? h->flagsB : nullptr, h->flags_stride,
                    exec_counter(h), h->oz_whatif, h->oz_prof);
            } else {
                oz::ozaki_trmm_kernel<SV, false><<<grid_oz, oz::OTHREADS, oz::Cfg<SV>::SMEM_BYTES, h->stream>>>(
                    mapsAq, h->mapsBq, oz_scale, h->scaleB, T64, rowtiles, rows_total, part, h->info + 1, h->var_bits, nullptr, nullptr, h->flags_stride, nullptr, 0, nullptr);
                h->exec_pairs_host += (unsigned long long)rowtiles * ((unsigned long long)T64 * (T64 + 1) / 2) * (unsigned long long)(SV * (SV + 1) / 2);
            }
        });
        toc(h, 0);
        LAUNCH_CHECK(h);
        Tpart = skipping ? 2 * T64 : T64;           // the skipping kernel writes two partial sums per (tile, row): one per epilogue warp group
    } else if (nrhs > 0) {
        const int rowtiles = (int)(rows_total / TS);
        int KC = 0, J = 0;
        trmm_splitk_plan(h, rowtiles, T, &KC, &J);
        tic(h, 0);
        int KV = T / 16 < 1 ? 1 : (T / 16 > 8 ? 8 : T / 16), JV = 0;            // matrix-vector form: chunk of k-tiles per job
        for (int t = 0; t < T; ++t) JV += (t + KV) / KV;
        if (h->splitk_ws && h->splitk_on && (long long)B * nrhs <= 8 && (long long)JV * 8 <= (long long)h->splitk_cap * TS) {
            trmv_partial_kernel<<<(unsigned)JV, 512, 0, h->stream>>>(rhs, h->Minv, h->Npad, B, nrhs, Bpad, KV, h->splitk_ws);
            LAUNCH_CHECK(h);
            trmv_reduce_kernel<<<(unsigned)T, 128, 0, h->stream>>>(h->splitk_ws, B, nrhs, Bpad, KV, rows_total, part);
        } else if (KC > 0) {
            trmm_splitk_kernel<<<(unsigned)(rowtiles * J), GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(*mapR, h->mapM, T, KC, J, h->splitk_ws);
            LAUNCH_CHECK(h);
            trmm_splitk_reduce_kernel<<<(unsigned)(rowtiles * T), 256, 0, h->stream>>>(h->splitk_ws, T, KC, J, rows_total, part);
        } else {
            trmm_sumsq_kernel<<<(unsigned)((long long)rowtiles * T), GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(*mapR, h->mapM, T, rowtiles, rows_total, part);
        }
        toc(h, 0);
        LAUNCH_CHECK(h);
    }
    finalize_kernel<D, P><<<(B + 127) / 128, 128, 0, h->stream>>>(macc, nsplit, part, Tpart, B, Bpad, rows_total, xr, vel_dev, h->kp, h->af, flags, out, q_off, Mtot,
                                                                  vals_out_slot);
    LAUNCH_CHECK(h);
    return 0;
}

typedef int (*chunk_fn)(gptb_handle*, const double*, const double*, int, int, unsigned, int, unsigned, const QueryOut&, long long,
                        long long, double*, double*, double*, double*, int, const CUtensorMap*, void*, double*, int, const SpatialWs*);

static chunk_fn pick_chunk_fn(int d, int p) {
    static const chunk_fn table[4][4] = {
        {query_chunk<1, 1>, query_chunk<1, 2>, query_chunk<1, 3>, query_chunk<1, 4>},
        {query_chunk<2, 1>, query_chunk<2, 2>, query_chunk<2, 3>, query_chunk<2, 4>},
        {query_chunk<3, 1>, query_chunk<3, 2>, query_chunk<3, 3>, query_chunk<3, 4>},
        {query_chunk<4, 1>, query_chunk<4, 2>, query_chunk<4, 3>, query_chunk<4, 4>}};
    return table[d - 1][p - 1];
}

extern "C" int gptb_query_dev(gptb_handle* h, const double* x_dev, int64_t M, uint32_t flags, const double* vel_dev, double* mean_dev,
                              double* std_dev, double* jac_dev, double* jacvar_dev, double* xhat_dev, double* vhat_dev,
                              double* vvar_dev, double* jphi_dev, double* dvar_dev) {
    if (!h || M < 0) return -1;
    if (M == 0) return 0;
    if (!x_dev) return -1;
    CU(h, cudaSetDevice(h->device));
    if (!h->have_alpha) GPTB_FAIL(h, -1, "gptb_query: model is not fitted");
    const int d = h->d, p = h->p;
    if ((flags & GPTB_MEAN) && !mean_dev) GPTB_FAIL(h, -1, "GPTB_MEAN without output buffer");
    if ((flags & GPTB_STD) && !std_dev) GPTB_FAIL(h, -1, "GPTB_STD without output buffer");
    if ((flags & GPTB_JAC) && !jac_dev) GPTB_FAIL(h, -1, "GPTB_JAC without output buffer");
    if ((flags & GPTB_JACVAR) && !jacvar_dev) GPTB_FAIL(h, -1, "GPTB_JACVAR without output buffer");
    if ((flags & GPTB_TRANSPORT) && (!xhat_dev || d != p)) GPTB_FAIL(h, -1, "GPTB_TRANSPORT needs xhat and d == p");
    if ((flags & GPTB_VELOCITY) && (!vhat_dev || !vel_dev || d != p)) GPTB_FAIL(h, -1, "GPTB_VELOCITY needs vel, vhat and d == p");
    if ((flags & GPTB_VELOCITY) && (flags & GPTB_JACVAR) && !vvar_dev) GPTB_FAIL(h, -1, "GPTB_VELOCITY|GPTB_JACVAR needs vvar");
    if ((flags & GPTB_JPHI) && (!jphi_dev || d != p)) GPTB_FAIL(h, -1, "GPTB_JPHI needs jphi and d == p");
    if ((flags & GPTB_DVAR) && !dvar_dev) GPTB_FAIL(h, -1, "GPTB_DVAR without output buffer");

    int nrhs = 0;
    unsigned genflags = 0;
    if (flags & (GPTB_STD | GPTB_DVAR)) { nrhs = 1; genflags |= 1u; }
    if (flags & (GPTB_JACVAR | GPTB_DVAR)) { nrhs = 1 + d; genflags |= 2u; }
    if (flags & GPTB_DVAR) { nrhs = 1 + 2 * d; genflags |= 4u | 1u; }
    // A handful of right-hand-side rows (a control-loop query, a rollout step): the matrix-vector form of the exact FP64 product
    // (trmv_partial_kernel) streams L^-1 once and beats the tiled INT8-sliced products, which are built for throughput -- so such a
    // call runs as if the handle were in "fp64" mode.  L^-1 in FP64 is kept in INT8 mode anyway (the guard's reference path).
    struct ModeRestore {
        gptb_handle* h;
        int saved;
        ~ModeRestore() { h->var_mode = saved; }
    } mode_restore{h, h->var_mode};
    if (nrhs > 0 && h->var_mode == 1 && h->splitk_on && h->have_minv && h->splitk_ws && M * nrhs <= 8) h->var_mode = 0;
    if (nrhs > 0 && h->var_mode == 1) {
        int rc = build_bplanes(h);
        if (rc) return rc;
        if ((rc = run_variance_guard(h))) return rc;       // may add digit planes or switch this model to the FP64 path
    }
    const bool ozaki = (nrhs > 0 && h->var_mode == 1);
    if (nrhs > 0 && !ozaki) {
        int rc = build_minv(h);
        if (rc) return rc;
    }
    // batch size: bounded by the workspace for the right-hand-side rows (nrhs * Bpad * Npad doubles / digit bytes)
    const bool fused_planes = ozaki && oz_fused_supported(d, p);
    const long long per_elem = ozaki ? (fused_planes ? h->var_slices : (long long)sizeof(double) + h->var_slices) : (long long)sizeof(double);
    auto batch_cap = [&](int nbuf) {
        long long bm = h->ws_limit / (nbuf * per_elem * nrhs * h->Npad);
        bm = bm / TS * TS;
        if (bm < TS) bm = TS;
        if (bm > h->batch_cap) bm = h->batch_cap;
        return bm;
    };
    long long Bmax = (nrhs > 0) ? batch_cap(1) : (1LL << 20);
    // INT8-sliced path with the fused generator: double-buffered batches so the generator of batch i+1 (FP64 pipe, gen stream)
    // runs under the digit-plane products of batch i (tensor pipe, main stream).  The stream is cut into >= 8 batches of
    // >= 8192 queries (enough 128 x 64 tiles to fill the persistent product kernel many times over).
    // spatial mode (8-bit planes, fused generator): sorted batches + block masks; with the overlap pipeline the per-batch sort
    // permutation and block masks are double-buffered like the digit planes
    const bool spatial_q = fused_planes && h->spatial && h->var_bits == 8;
    int nbuf = 1;
    if (fused_planes && h->pipeline && M >= 2 * 8192) {
        long long bp = (M / 8 + TS - 1) / TS * TS;
        if (bp < 8192) bp = 8192;
        const long long cap2 = batch_cap(2);
        if (bp > cap2) bp = cap2;
        if (bp < M) { nbuf = 2; Bmax = bp; }
    }
    long long Bfirst = (M < Bmax) ? (M + TS - 1) / TS * TS : Bmax;
    const int T = h->T;
    const int NACC = p + p * d;
    // split the training range when the batch is too small to fill the GPU
    int nsplit = 1;
    {
        long long ctas = Bfirst / QPB;
        while (ctas * nsplit < 296 && nsplit * 2 <= h->Npad / 128 && nsplit < 64) nsplit *= 2;
    }
    size_t need = 0;
    auto carve = [&](size_t doubles) { size_t off = need; need += (doubles * sizeof(double) + 255) / 256 * 256; return off; };
    size_t o_rhs = carve(fused_planes ? 0 : (size_t)nrhs * Bfirst * h->Npad);
    size_t o_part[2], o_ozp[2], o_ozs[2], o_macc[2], o_xr[2];
    for (int b = 0; b < nbuf; ++b) {
        o_part[b] = carve((size_t)(nrhs > 0 ? (ozaki ? 4 * T : T) : 0) * nrhs * Bfirst);
        o_ozp[b] = carve(ozaki ? ((size_t)h->var_slices * nrhs * Bfirst * h->Npad + 7) / 8 : 0);
        o_ozs[b] = carve(ozaki ? (size_t)nrhs * Bfirst : 0);
        o_macc[b] = carve((size_t)nsplit * Bfirst * NACC);
        o_xr[b] = carve((size_t)Bfirst * d);
    }
    SpatialWs sp{};
    size_t o_sp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (spatial_q) {
        for (int i = 0; i < 5; ++i) o_sp[i] = carve(((size_t)Bfirst + 1) / 2);                 // Bfirst 32-bit words each (3 shared + vals_out per slot)
        sp.cub_bytes = 0;
        CU(h, cub::DeviceRadixSort::SortPairs(nullptr, sp.cub_bytes, (unsigned*)nullptr, (unsigned*)nullptr, (unsigned*)nullptr, (unsigned*)nullptr,
                                              (int)Bfirst, 0, h->mbox.bits * d, h->stream));
        o_sp[5] = carve((sp.cub_bytes + 7) / 8);
        for (int b = 0; b < 2; ++b) o_sp[6 + b] = carve(((size_t)(nrhs * Bfirst / 128) * h->flags_stride + 1) / 2);
    }
    if (need > h->ws_bytes) {
        CU(h, cudaStreamSynchronize(h->stream));
        CU(h, cudaStreamSynchronize(h->gen));
        if (h->ws) cudaFree(h->ws);
        h->ws = nullptr;
        h->ws_bytes = 0;
        CU(h, cudaMalloc(&h->ws, need));
        h->ws_bytes = need;
    }
    char* base = reinterpret_cast<char*>(h->ws);
    double* rhs = reinterpret_cast<double*>(base + o_rhs);
    if (spatial_q) {
        sp.keys_in = reinterpret_cast<unsigned*>(base + o_sp[0]);
        sp.keys_out = reinterpret_cast<unsigned*>(base + o_sp[1]);
        sp.vals_in = reinterpret_cast<unsigned*>(base + o_sp[2]);
        sp.vals_out[0] = reinterpret_cast<unsigned*>(base + o_sp[3]);
        sp.vals_out[1] = reinterpret_cast<unsigned*>(base + o_sp[4]);
        sp.cub_tmp = base + o_sp[5];
        sp.flagsA[0] = reinterpret_cast<unsigned*>(base + o_sp[6]);
        sp.flagsA[1] = reinterpret_cast<unsigned*>(base + o_sp[7]);
    }
    QueryOut out{mean_dev, std_dev, jac_dev, jacvar_dev, xhat_dev, vhat_dev, vvar_dev, jphi_dev, dvar_dev};
    chunk_fn fn = pick_chunk_fn(d, p);
    CUtensorMap mapR_full, mapR_tail;
    int Bpad_mapped = -1;
    if (nbuf == 2) {   // the generator stream starts behind everything already queued on the main stream (inputs, digit planes of L^-1)
        CU(h, cudaEventRecord(h->ev_start, h->stream));
        CU(h, cudaStreamWaitEvent(h->gen, h->ev_start, 0));
    }
    long long ib = 0;
    for (long long q0 = 0; q0 < M; q0 += Bfirst, ++ib) {
        int B = (int)((M - q0 < Bfirst) ? (M - q0) : Bfirst);
        int Bpad = (B + TS - 1) / TS * TS;
        const CUtensorMap* mapR = nullptr;
        if (nrhs > 0 && !fused_planes) {
            // the right-hand-side rows of this batch: (nrhs * Bpad) x Npad, row-major in the workspace
            CUtensorMap* m = (Bpad == Bfirst) ? &mapR_full : &mapR_tail;
            if (Bpad != Bpad_mapped) {
                MAKE_MAP(h, m, rhs, (long long)nrhs * Bpad, h->Npad, h->Npad);
                Bpad_mapped = Bpad;
            }
            mapR = m;
        }
        const int b = (nbuf == 2) ? (int)(ib & 1) : 0;
        if (nbuf == 2 && ib >= 2) CU(h, cudaStreamWaitEvent(h->gen, h->ev_done[b], 0));     // buffer set b is free again
        int rc = fn(h, x_dev + q0 * d, vel_dev ? vel_dev + q0 * d : nullptr, B, Bpad, flags, nrhs, genflags, out, q0, M, rhs,
                    reinterpret_cast<double*>(base + o_part[b]), reinterpret_cast<double*>(base + o_macc[b]),
                    reinterpret_cast<double*>(base + o_xr[b]), nsplit, mapR, base + o_ozp[b], reinterpret_cast<double*>(base + o_ozs[b]),
                    nbuf == 2 ? b : -1, spatial_q ? &sp : nullptr);
        if (rc) return rc;
        if (nbuf == 2) CU(h, cudaEventRecord(h->ev_done[b], h->stream));
    }
    return 0;
}

// Host-pointer query: staged through two device buffer sets in slices of at most 2^19 queries, so that arbitrarily long query streams
// (BASELINE configs 4/5: 2^26 .. 2^29 points) need ~150 MB of staging, and pipelined: the H2D copy of slice i+1 and the D2H copy of
// slice i-1 run on their own streams under the kernels of slice i.  (Round 1 copied a 4M-point slice in, ran it, copied it out and
// synchronised: the copies were 3 % of a step on one GPU and the only thing bending the 8-GPU scaling curve, where eight ranks share
// the host path.)  A call that fits one slice uses the main stream only (no event hops on the latency path of small queries).
extern "C" int gptb_query(gptb_handle* h, const double* x, int64_t M, uint32_t flags, const double* vel, double* mean, double* std,
                          double* jac, double* jacvar, double* xhat, double* vhat, double* vvar, double* jphi, double* dvar) {
    if (!h || M < 0) return -1;
    if (M == 0) return 0;
    if (!x) return -1;
    CU(h, cudaSetDevice(h->device));
    const int d = h->d, p = h->p;
    if (d < 1) GPTB_FAIL(h, -1, "gptb_query: model is not fitted");
    if ((flags & GPTB_VELOCITY) && !vel) GPTB_FAIL(h, -1, "GPTB_VELOCITY without vel");
    const int64_t HOST_SLICE = 1 << 17;
    struct Seg { const double* hin; double* hout; size_t per; size_t off; };
    Seg segs[11] = {{x, nullptr, (size_t)d, 0},
                    {(flags & GPTB_VELOCITY) ? vel : nullptr, nullptr, (size_t)d, 0},
                    {nullptr, (flags & GPTB_MEAN) ? mean : nullptr, (size_t)p, 0},
                    {nullptr, (flags & GPTB_STD) ? std : nullptr, (size_t)p, 0},
                    {nullptr, (flags & GPTB_JAC) ? jac : nullptr, (size_t)p * d, 0},
                    {nullptr, (flags & GPTB_JACVAR) ? jacvar : nullptr, (size_t)p * d, 0},
                    {nullptr, (flags & GPTB_TRANSPORT) ? xhat : nullptr, (size_t)d, 0},
                    {nullptr, (flags & GPTB_VELOCITY) ? vhat : nullptr, (size_t)d, 0},
                    {nullptr, ((flags & GPTB_VELOCITY) && (flags & GPTB_JACVAR)) ? vvar : nullptr, (size_t)p, 0},
                    {nullptr, (flags & GPTB_JPHI) ? jphi : nullptr, (size_t)d * d, 0},
                    {nullptr, (flags & GPTB_DVAR) ? dvar : nullptr, (size_t)d, 0}};
    // slice plan: large slices (one device batch each: the kernels run 4 % faster on 2^19 queries than on 2^17) with a small LAST one,
    // because what the pipeline cannot hide is the H2D copy of the first slice (inputs only: small) and the D2H copy of the last
    const int64_t BIG = (h->batch_cap > HOST_SLICE) ? ((h->batch_cap < (1 << 19)) ? h->batch_cap : (int64_t)(1 << 19)) : HOST_SLICE;
    std::vector<std::pair<int64_t, int64_t>> plan;       // (first query, count)
    if (M <= HOST_SLICE) {
        plan.emplace_back(0, M);
    } else {
        const int64_t TAIL = HOST_SLICE / 2;
        int64_t q = 0, rem = M;
        while (rem > BIG + TAIL) { plan.emplace_back(q, BIG); q += BIG; rem -= BIG; }
        if (rem > TAIL) { plan.emplace_back(q, rem - TAIL); q += rem - TAIL; rem = TAIL; }
        plan.emplace_back(q, rem);
    }
    int64_t slice = 0;
    for (auto& pl : plan) slice = pl.second > slice ? pl.second : slice;
    const int64_t nslices = (int64_t)plan.size();
    const int nsets = nslices > 1 ? 2 : 1;
    size_t tot = 0;
    for (auto& sg : segs) {
        if (sg.hin || sg.hout) {
            sg.off = tot;
            tot += (sg.per * (size_t)slice * sizeof(double) + 255) / 256 * 256;
        }
    }
    if (tot > h->stage_bytes || (nsets == 2 && !h->stage[1])) {
        CU(h, cudaStreamSynchronize(h->stream));
        const size_t want = tot > h->stage_bytes ? tot : h->stage_bytes;
        for (int i = 0; i < 2; ++i) {
            if (h->stage[i]) cudaFree(h->stage[i]);
            h->stage[i] = nullptr;
        }
        h->stage_bytes = 0;
        for (int i = 0; i < nsets; ++i) CU(h, cudaMalloc(&h->stage[i], want));
        h->stage_bytes = want;
    }
    auto dp = [&](int set, int i) -> double* {
        return (segs[i].hin || segs[i].hout) ? reinterpret_cast<double*>(reinterpret_cast<char*>(h->stage[set]) + segs[i].off) : nullptr;
    };
    const bool piped = nslices > 1;
    cudaStream_t sin = piped ? h->s_h2d : h->stream, sout = piped ? h->s_d2h : h->stream;
    if (piped) {       // the copy streams start behind whatever is already queued on the main stream
        CU(h, cudaEventRecord(h->ev_start, h->stream));
        CU(h, cudaStreamWaitEvent(sin, h->ev_start, 0));
    }
    for (int64_t is = 0; is < nslices; ++is) {
        const int64_t q0 = plan[is].first;
        const int64_t m = plan[is].second;
        const int set = piped ? (int)(is & 1) : 0;
        if (piped && is >= 2) CU(h, cudaStreamWaitEvent(sin, h->ev_cd[set], 0));          // the kernels of slice is-2 have consumed this set's inputs
        for (int i = 0; i < 2; ++i)
            if (segs[i].hin) CU(h, cudaMemcpyAsync(dp(set, i), segs[i].hin + (size_t)q0 * segs[i].per, segs[i].per * m * sizeof(double), cudaMemcpyHostToDevice, sin));
        if (piped) {
            CU(h, cudaEventRecord(h->ev_in[set], sin));
            CU(h, cudaStreamWaitEvent(h->stream, h->ev_in[set], 0));
            if (is >= 2) CU(h, cudaStreamWaitEvent(h->stream, h->ev_oc[set], 0));         // this set's outputs of slice is-2 have left
        }
        int rc = gptb_query_dev(h, dp(set, 0), m, flags, dp(set, 1), dp(set, 2), dp(set, 3), dp(set, 4), dp(set, 5), dp(set, 6), dp(set, 7), dp(set, 8),
                                dp(set, 9), dp(set, 10));
        if (rc) {
            cudaStreamSynchronize(sin); cudaStreamSynchronize(h->stream); cudaStreamSynchronize(sout);
            return rc;
        }
        if (piped) {
            CU(h, cudaEventRecord(h->ev_cd[set], h->stream));
            CU(h, cudaStreamWaitEvent(sout, h->ev_cd[set], 0));
        }
        for (int i = 2; i < 10; ++i)
            if (segs[i].hout) CU(h, cudaMemcpyAsync(segs[i].hout + (size_t)q0 * segs[i].per, dp(set, i), segs[i].per * m * sizeof(double), cudaMemcpyDeviceToHost, sout));
        if (segs[10].hout)   // dvar is (d, M): this slice fills columns [q0, q0 + m) of each row
            CU(h, cudaMemcpy2DAsync(dvar + q0, sizeof(double) * M, dp(set, 10), sizeof(double) * m, sizeof(double) * m, d, cudaMemcpyDeviceToHost, sout));
        if (piped) CU(h, cudaEventRecord(h->ev_oc[set], sout));
    }
    if (piped) {
        CU(h, cudaStreamSynchronize(sin));
        CU(h, cudaStreamSynchronize(sout));
    }
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// one scratch buffer, carved in 256-byte aligned pieces (sizes in doubles); kept while it stays below SCRATCH_KEEP
constexpr size_t SCRATCH_KEEP = 256u << 20;
struct Carver {
    size_t need = 0;
    size_t add(size_t doubles) { const size_t off = need; need += (doubles * sizeof(double) + 255) / 256 * 256; return off; }
};
static int ensure_scratch(gptb_handle* h, size_t bytes) {
    if (bytes <= h->scratch_bytes) return 0;
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->scratch) cudaFree(h->scratch);
    h->scratch = nullptr;
    h->scratch_bytes = 0;
    CU(h, cudaMalloc(&h->scratch, bytes));
    h->scratch_bytes = bytes;
    return 0;
}
static void release_big_scratch(gptb_handle* h) {
    if (h->scratch_bytes > SCRATCH_KEEP) {
        cudaFree(h->scratch);
        h->scratch = nullptr;
        h->scratch_bytes = 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// rank-1 append of a training point at the current hyper-parameters
// ---------------------------------------------------------------------------------------------------------------
extern "C" int gptb_append_point(gptb_handle* h, const double* x, const double* y, double* lml) {
    if (!h || !x || !y) return -1;
    CU(h, cudaSetDevice(h->device));
    if (!h->have_train || !h->have_factor || !h->have_alpha) GPTB_FAIL(h, -1, "gptb_append_point: no fitted model (gptb_set_train + gptb_factorize first)");
    const int d = h->d, p = h->p;
    const long long N = h->N, Npad = h->Npad;
    int rc = 0;
    if (N == Npad) {
        // no room in the last 128-row tile: the N x N buffers have to grow -- re-factorise once with the point included (every 128th append)
        std::vector<double> xs((size_t)d * Npad), ys((size_t)p * Npad), X2((size_t)(N + 1) * d), Y2((size_t)(N + 1) * p);
        CU(h, cudaMemcpyAsync(xs.data(), h->X, sizeof(double) * d * Npad, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaMemcpyAsync(ys.data(), h->Y, sizeof(double) * p * Npad, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        const std::vector<int> perm = h->perm;
        for (long long n = 0; n < N; ++n) {
            const long long dst = perm.empty() ? n : perm[(size_t)n];           // back to the caller's order
            for (int a = 0; a < d; ++a) X2[(size_t)dst * d + a] = xs[(size_t)a * Npad + n];
            for (int o = 0; o < p; ++o) Y2[(size_t)dst * p + o] = ys[(size_t)o * Npad + n];
        }
        for (int a = 0; a < d; ++a) X2[(size_t)N * d + a] = x[a];
        for (int o = 0; o < p; ++o) Y2[(size_t)N * p + o] = y[o];
        const KParams kp = h->kp;
        if ((rc = gptb_set_train(h, X2.data(), Y2.data(), N + 1, d, p))) return rc;
        if ((rc = gptb_factorize(h, kp.c, kp.ell, kp.s2, kp.jitter, lml))) return rc;
        return build_minv(h);
    }
    if ((rc = build_minv(h))) return rc;
    NewPoint np{};
    for (int a = 0; a < d; ++a) { np.x[a] = x[a]; np.xs[a] = x[a] / h->kp.ell[a]; }
    for (int o = 0; o < p; ++o) np.y[o] = y[o];
    Carver cv;
    const size_t o_kv = cv.add((size_t)Npad), o_l = cv.add((size_t)Npad), o_r = cv.add((size_t)Npad);
    if ((rc = ensure_scratch(h, cv.need))) return rc;
    double *kv = reinterpret_cast<double*>(h->scratch + o_kv), *l = reinterpret_cast<double*>(h->scratch + o_l), *r = reinterpret_cast<double*>(h->scratch + o_r);
    double* scal = h->scal + 48;
    append_kvec_kernel<<<(unsigned)((Npad + 255) / 256), 256, 0, h->stream>>>(h->Xs, (int)N, (int)Npad, d, np, h->kp, kv);
    LAUNCH_CHECK(h);
    gemv_lower_rows_kernel<<<(unsigned)((N + 7) / 8), 256, 0, h->stream>>>(h->Minv, Npad, (int)N, kv, l);
    LAUNCH_CHECK(h);
    append_pivot_kernel<<<1, 1024, 0, h->stream>>>(l, (int)N, h->kp.c + h->kp.s2 + h->kp.jitter, scal);
    LAUNCH_CHECK(h);
    double dd = 0.0;
    CU(h, cudaMemcpyAsync(&dd, scal, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (!(dd > 0.0)) GPTB_FAIL(h, (int)(N + 1), "the kernel matrix is not positive definite: leading minor of order %lld (appended point)", N + 1);
    gemv_mirror_rows_kernel<<<(unsigned)((N + 7) / 8), 256, 0, h->stream>>>(h->Minv, Npad, (int)N, l, scal, r);
    LAUNCH_CHECK(h);
    append_write_kernel<<<1, 1024, 0, h->stream>>>(h->Lbuf, h->Minv, h->Dinv, Npad, (int)N, (int)Npad, d, p, l, r, scal, np, h->X, h->Xs, h->Y, h->alpha);
    LAUNCH_CHECK(h);
    h->N = N + 1;
    if (!h->perm.empty()) h->perm.push_back((int)N);           // spatial mode: appended points follow the Morton-ordered ones
    h->have_kinv = false;
    invalidate_planes(h);
    if (lml) return lml_value(h, lml);
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// minimum-variance stabilised rollouts, device-resident (one CUDA graph per step shape, replayed `steps` times)
// ---------------------------------------------------------------------------------------------------------------
extern "C" int gptb_rollout_min_variance(gptb_handle* h, const double* start, int64_t K, int steps, double gain, double* traj) {
    if (!h || !start || !traj || K < 1 || steps < 1) return -1;
    CU(h, cudaSetDevice(h->device));
    if (!h->have_alpha) GPTB_FAIL(h, -1, "gptb_rollout_min_variance: model is not fitted");
    const int d = h->d, p = h->p;
    if (d != p || d < 2 || d > 3) GPTB_FAIL(h, -1, "rollouts need a square dynamics model with d = p in {2, 3}, got d=%d p=%d", d, p);
    Carver cv;
    const size_t o_pos = cv.add((size_t)K * d), o_mean = cv.add((size_t)K * p), o_std = cv.add((size_t)K * p), o_dv = cv.add((size_t)K * d),
                 o_slot = cv.add((size_t)K * d), o_traj = cv.add((size_t)steps * K * d);
    int rc = ensure_scratch(h, cv.need);
    if (rc) return rc;
    auto at = [&](size_t off) { return reinterpret_cast<double*>(h->scratch + off); };
    double *pos = at(o_pos), *md = at(o_mean), *sd = at(o_std), *dv = at(o_dv), *slot = at(o_slot), *tr = at(o_traj);
    CU(h, cudaMemcpyAsync(pos, start, sizeof(double) * K * d, cudaMemcpyHostToDevice, h->stream));
    const uint32_t fl = GPTB_MEAN | GPTB_STD | GPTB_DVAR;
    const unsigned grid = (unsigned)((K + 127) / 128);
    auto one_step = [&](int t) -> int {
        int r = gptb_query_dev(h, pos, K, fl, nullptr, md, sd, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, dv);
        if (r) return r;
        double* dst = tr + (size_t)t * K * d;
        if (d == 2) rollout_step_kernel<2><<<grid, 128, 0, h->stream>>>(pos, md, sd, dv, K, gain, dst);
        else rollout_step_kernel<3><<<grid, 128, 0, h->stream>>>(pos, md, sd, dv, K, gain, dst);
        h->launches++;
        return 0;
    };
    // step 0 eagerly: builds the variance operands, grows the workspace (both synchronise, which a capture does not allow)
    if ((rc = one_step(0))) { release_big_scratch(h); return rc; }
    const bool was_timing = h->timing;
    h->timing = false;
    int t = 1;
    // the steps are launch-bound for small K (~20 kernels of a few microseconds each): capture one step as a graph and replay it; the
    // trajectory slot is the only thing that changes, so the captured step writes into a fixed slot and a D2D copy files it
    if (steps > 2 && !h->pipeline) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const long long before = h->launches;
        bool ok = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            int r = gptb_query_dev(h, pos, K, fl, nullptr, md, sd, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, dv);
            if (d == 2) rollout_step_kernel<2><<<grid, 128, 0, h->stream>>>(pos, md, sd, dv, K, gain, slot);
            else rollout_step_kernel<3><<<grid, 128, 0, h->stream>>>(pos, md, sd, dv, K, gain, slot);
            h->launches++;
            ok = (cudaStreamEndCapture(h->stream, &graph) == cudaSuccess) && r == 0 && graph != nullptr;
        }
        const long long per_step = h->launches - before;
        if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
        if (ok) {
            for (; t < steps; ++t) {
                if (cudaGraphLaunch(exec, h->stream) != cudaSuccess) { ok = false; break; }
                CU(h, cudaMemcpyAsync(tr + (size_t)t * K * d, slot, sizeof(double) * K * d, cudaMemcpyDeviceToDevice, h->stream));
                h->launches += per_step;
            }
        } else {
            cudaGetLastError();               // a failed capture leaves a sticky-free error: fall through to eager steps
        }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        if (!ok && t < steps) { h->timing = was_timing; release_big_scratch(h); GPTB_FAIL(h, -3, "graph replay of the rollout step failed"); }
    }
    for (; t < steps; ++t)
        if ((rc = one_step(t))) { h->timing = was_timing; release_big_scratch(h); return rc; }
    h->timing = was_timing;
    CU(h, cudaMemcpyAsync(traj, tr, sizeof(double) * (size_t)steps * K * d, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    release_big_scratch(h);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// dense lattice queries with on-device generation and reduction (BASELINE config 5)
// ---------------------------------------------------------------------------------------------------------------
extern "C" int gptb_query_grid(gptb_handle* h, const double* origin, const double* step, const int64_t* dims, int64_t first, int64_t count,
                               uint32_t flags, double* stats, int64_t sample_stride, double* sample_out) {
    if (!h || !origin || !step || !dims || !stats || first < 0 || count < 0) return -1;
    CU(h, cudaSetDevice(h->device));
    if (!h->have_alpha) GPTB_FAIL(h, -1, "gptb_query_grid: model is not fitted");
    const int d = h->d, p = h->p;
    if (flags & ~(GPTB_MEAN | GPTB_STD | GPTB_JAC | GPTB_AFFINE_IN)) GPTB_FAIL(h, -1, "gptb_query_grid supports GPTB_MEAN | GPTB_STD | GPTB_JAC (| GPTB_AFFINE_IN)");
    GridSpec g{};
    long long total = 1;
    for (int a = 0; a < d; ++a) {
        if (dims[a] < 1) GPTB_FAIL(h, -1, "gptb_query_grid: dims[%d] = %lld", a, (long long)dims[a]);
        g.origin[a] = origin[a]; g.step[a] = step[a]; g.dims[a] = dims[a];
        total *= dims[a];
    }
    if (first + count > total) GPTB_FAIL(h, -1, "gptb_query_grid: points [%lld, %lld) exceed the lattice of %lld", (long long)first, (long long)(first + count), total);
    const int ncol = ((flags & GPTB_MEAN) ? p : 0) + ((flags & GPTB_STD) ? 1 : 0) + ((flags & GPTB_JAC) ? p * d : 0);
    if (ncol == 0) GPTB_FAIL(h, -1, "gptb_query_grid: no output requested");
    for (int i = 0; i < 4 * ncol; ++i) stats[i] = 0.0;
    if (count == 0) return 0;
    const long long B = 65536;
    // sample j = lattice point j * stride (global numbering, so shards of one lattice produce disjoint pieces of one sample set)
    const long long j_first = sample_stride > 0 ? (first + sample_stride - 1) / sample_stride : 0;
    const long long j_end = sample_stride > 0 ? (first + count + sample_stride - 1) / sample_stride : 0;
    const long long nsample = j_end - j_first;
    if (nsample > 0 && !sample_out) GPTB_FAIL(h, -1, "gptb_query_grid: sample_stride without sample_out");
    Carver cv;
    const size_t o_x = cv.add((size_t)B * d), o_mean = cv.add((size_t)B * p), o_std = cv.add((size_t)B * p), o_jac = cv.add((size_t)B * p * d),
                 o_pack = cv.add((size_t)B * ncol), o_part = cv.add((size_t)GRID_STAT_BLOCKS * ncol * 4), o_acc = cv.add((size_t)ncol * 4),
                 o_smp = cv.add((size_t)(nsample > 0 ? nsample : 1) * ncol);
    int rc = ensure_scratch(h, cv.need);
    if (rc) return rc;
    auto at = [&](size_t off) { return reinterpret_cast<double*>(h->scratch + off); };
    double *xd = at(o_x), *md = at(o_mean), *sd = at(o_std), *jd = at(o_jac), *pack = at(o_pack), *part = at(o_part), *acc = at(o_acc), *smp = at(o_smp);
    bool firstb = true;
    for (long long q0 = 0; q0 < count; q0 += B) {
        const int m = (int)((count - q0 < B) ? (count - q0) : B);
        grid_points_kernel<<<(m + 255) / 256, 256, 0, h->stream>>>(g, d, first + q0, m, xd);
        LAUNCH_CHECK(h);
        rc = gptb_query_dev(h, xd, m, flags, nullptr, (flags & GPTB_MEAN) ? md : nullptr, (flags & GPTB_STD) ? sd : nullptr, (flags & GPTB_JAC) ? jd : nullptr,
                            nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (rc) { release_big_scratch(h); return rc; }
        grid_pack_kernel<<<(m + 255) / 256, 256, 0, h->stream>>>(md, sd, jd, m, p, d, flags, pack);
        LAUNCH_CHECK(h);
        grid_stats_partial_kernel<<<GRID_STAT_BLOCKS, 256, 0, h->stream>>>(pack, m, ncol, part);
        LAUNCH_CHECK(h);
        grid_stats_final_kernel<<<(ncol + 31) / 32, 32, 0, h->stream>>>(part, ncol, acc, firstb ? 1 : 0);
        LAUNCH_CHECK(h);
        firstb = false;
        if (nsample > 0) {
            const long long ja = (first + q0 + sample_stride - 1) / sample_stride, jb = (first + q0 + m + sample_stride - 1) / sample_stride;
            if (jb > ja) {
                grid_sample_kernel<<<(unsigned)((jb - ja + 127) / 128), 128, 0, h->stream>>>(pack, ncol, first + q0, m, sample_stride, j_first, smp);
                LAUNCH_CHECK(h);
            }
        }
    }
    CU(h, cudaMemcpyAsync(stats, acc, sizeof(double) * 4 * ncol, cudaMemcpyDeviceToHost, h->stream));
    if (nsample > 0) CU(h, cudaMemcpyAsync(sample_out, smp, sizeof(double) * nsample * ncol, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    release_big_scratch(h);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// joint covariance
// ---------------------------------------------------------------------------------------------------------------
template <int D, int P>
static int cov_generate(gptb_handle* h, const double* x_dev, int M, int Mpad, double* rhs, double* xr, double* macc, double* mean_dev) {
    Affine af = h->af;
    af.on = 0;
    dim3 grid(Mpad / QPB, 1);
    kstar_kernel<D, P, 1, 5><<<grid, 256, 0, h->stream>>>(x_dev, h->Xs, h->alpha, (int)h->N, (int)h->Npad, M, Mpad, h->kp, af, 1u, rhs, xr, macc, 1, nullptr, 0, nullptr, DigitScales{});
    LAUNCH_CHECK(h);
    QueryOut out{mean_dev, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    finalize_kernel<D, P><<<(M + 127) / 128, 128, 0, h->stream>>>(macc, 1, nullptr, h->T, M, Mpad, Mpad, xr, nullptr, h->kp, h->af, GPTB_MEAN, out, 0, M, nullptr);
    LAUNCH_CHECK(h);
    return 0;
}

typedef int (*covgen_fn)(gptb_handle*, const double*, int, int, double*, double*, double*, double*);
static covgen_fn pick_covgen(int d, int p) {
    static const covgen_fn table[4][4] = {
        {cov_generate<1, 1>, cov_generate<1, 2>, cov_generate<1, 3>, cov_generate<1, 4>},
        {cov_generate<2, 1>, cov_generate<2, 2>, cov_generate<2, 3>, cov_generate<2, 4>},
        {cov_generate<3, 1>, cov_generate<3, 2>, cov_generate<3, 3>, cov_generate<3, 4>},
        {cov_generate<4, 1>, cov_generate<4, 2>, cov_generate<4, 3>, cov_generate<4, 4>}};
    return table[d - 1][p - 1];
}

extern "C" int gptb_query_cov(gptb_handle* h, const double* x, int64_t M, double* mean, double* cov) {
    if (!h || M < 0) return -1;
    if (M == 0) return 0;
    if (!x || !mean || !cov) return -1;
    CU(h, cudaSetDevice(h->device));
    if (!h->have_alpha) GPTB_FAIL(h, -1, "gptb_query_cov: model is not fitted");
    if (M > 32768) GPTB_FAIL(h, -1, "gptb_query_cov: M=%lld exceeds the 32768-point limit of the joint covariance", (long long)M);
    int rc = build_minv(h);
    if (rc) return rc;
    const int d = h->d, p = h->p, T = h->T;
    const int Mpad = (int)((M + TS - 1) / TS * TS);
    const long long Npad = h->Npad;
    Carver cv;
    const size_t o_xd = cv.add((size_t)M * d), o_rhs = cv.add((size_t)Mpad * Npad), o_W = cv.add((size_t)Mpad * Npad), o_xr = cv.add((size_t)Mpad * d),
                 o_macc = cv.add((size_t)Mpad * (p + p * d)), o_mean = cv.add((size_t)M * p), o_cov = cv.add((size_t)M * M);
    if ((rc = ensure_scratch(h, cv.need))) return rc;
    auto at = [&](size_t off) { return reinterpret_cast<double*>(h->scratch + off); };
    double *xd = at(o_xd), *rhs = at(o_rhs), *W = at(o_W), *xr = at(o_xr), *macc = at(o_macc), *mean_d = at(o_mean), *cov_d = at(o_cov);
    auto cleanup = [&]() { release_big_scratch(h); };
#define CUX(call)                                                                                           \
    do {                                                                                                    \
        cudaError_t _e = (call);                                                                            \
        if (_e != cudaSuccess) { cleanup(); GPTB_FAIL(h, -2, "CUDA error %s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__); } \
    } while (0)
    CUX(cudaMemcpyAsync(xd, x, sizeof(double) * M * d, cudaMemcpyHostToDevice, h->stream));
    rc = pick_covgen(d, p)(h, xd, (int)M, Mpad, rhs, xr, macc, mean_d);
    if (rc) { cleanup(); return rc; }
    CUtensorMap mapR, mapW;
    if (!make_operand_map(&mapR, rhs, Mpad, Npad, Npad) || !make_operand_map(&mapW, W, Mpad, Npad, Npad)) {
        cleanup();
        GPTB_FAIL(h, -5, "cuTensorMapEncodeTiled failed");
    }
    const int mt = Mpad / TS;
    trmm_store_kernel<<<(unsigned)(mt * T), GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(mapR, h->mapM, T, mt, W, Npad);
    h->launches++;
    dispatch_d(d, [&](auto D) {
        cov_kernel<decltype(D)::value><<<(unsigned)(mt * mt), GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(mapW, T, mt, xr, h->kp, (int)M, cov_d);
    });
    h->launches++;
    CUX(cudaGetLastError());
    CUX(cudaMemcpyAsync(mean, mean_d, sizeof(double) * M * p, cudaMemcpyDeviceToHost, h->stream));
    CUX(cudaMemcpyAsync(cov, cov_d, sizeof(double) * M * M, cudaMemcpyDeviceToHost, h->stream));
    CUX(cudaStreamSynchronize(h->stream));
#undef CUX
    cleanup();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// orientation transport
// ---------------------------------------------------------------------------------------------------------------
static int transport_orientation_impl(gptb_handle* h, const double* pos, const double* ori, int64_t M, double* ori_out, double* jphi, int mode) {
    if (!h || M < 0) return -1;
    if (M == 0) return 0;
    if (!pos || !ori || !ori_out) return -1;
    CU(h, cudaSetDevice(h->device));
    if (!h->have_alpha) GPTB_FAIL(h, -1, "gptb_transport_orientation: model is not fitted");
    if (h->d != 3 || h->p != 3) GPTB_FAIL(h, -1, "orientation transport needs a 3-D map (d = p = 3), got d=%d p=%d", h->d, h->p);
    Carver cv;
    const size_t o_pd = cv.add((size_t)M * 3), o_od = cv.add((size_t)M * 4), o_jd = cv.add((size_t)M * 9), o_jp = cv.add((size_t)M * 9), o_qd = cv.add((size_t)M * 4);
    if (int rcs = ensure_scratch(h, cv.need)) return rcs;
    auto at = [&](size_t off) { return reinterpret_cast<double*>(h->scratch + off); };
    double *pd = at(o_pd), *od = at(o_od), *jd = at(o_jd), *jp = at(o_jp), *qd = at(o_qd);
    auto cleanup = [&]() { release_big_scratch(h); };
    auto fail = [&](cudaError_t e, int line) {
        cleanup();
        char b[256];
        snprintf(b, sizeof(b), "CUDA error %s at %s:%d", cudaGetErrorString(e), __FILE__, line);
        h->err = b;
        return -2;
    };
    cudaError_t e;
    if ((e = cudaMemcpyAsync(pd, pos, sizeof(double) * M * 3, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) return fail(e, __LINE__);
    if ((e = cudaMemcpyAsync(od, ori, sizeof(double) * M * 4, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) return fail(e, __LINE__);
    // mode 0: Jphi = R + Jpsi(pos) R at the UN-rotated positions (quirk Q7); mode 1 (diffeomorphic variant): Jpsi at gamma(pos)
    int rc = gptb_query_dev(h, pd, M, mode == 1 ? (GPTB_JAC | GPTB_JPHI | GPTB_AFFINE_IN) : (GPTB_JAC | GPTB_JPHI), nullptr, nullptr, nullptr, jd, nullptr, nullptr,
                            nullptr, nullptr, jp, nullptr);
    if (rc) { cleanup(); return rc; }
    quat_transport_kernel<<<(unsigned)((M + 127) / 128), 128, 0, h->stream>>>(mode == 1 ? jd : jp, od, M, qd, mode, h->af);
    h->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, __LINE__);
    if ((e = cudaMemcpyAsync(ori_out, qd, sizeof(double) * M * 4, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) return fail(e, __LINE__);
    if (jphi && (e = cudaMemcpyAsync(jphi, jp, sizeof(double) * M * 9, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) return fail(e, __LINE__);
    if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return fail(e, __LINE__);
    cleanup();
    return 0;
}

extern "C" int gptb_transport_orientation(gptb_handle* h, const double* pos, const double* ori, int64_t M, double* ori_out, double* jphi) {
    return transport_orientation_impl(h, pos, ori, M, ori_out, jphi, 0);
}
extern "C" int gptb_transport_orientation_diffeo(gptb_handle* h, const double* pos, const double* ori, int64_t M, double* ori_out) {
    return transport_orientation_impl(h, pos, ori, M, ori_out, nullptr, 1);
}

extern "C" int gptb_transport_stiffness(gptb_handle* h, const double* pos, const double* stiff, int64_t M, double* stiff_out, double* jphi) {
    if (!h || M < 0) return -1;
    if (M == 0) return 0;
    if (!pos || !stiff || !stiff_out) return -1;
    CU(h, cudaSetDevice(h->device));
    if (!h->have_alpha) GPTB_FAIL(h, -1, "gptb_transport_stiffness: model is not fitted");
    const int d = h->d;
    if (d != h->p || d < 2) GPTB_FAIL(h, -1, "stiffness transport needs a square map (d = p >= 2), got d=%d p=%d", h->d, h->p);
    const size_t dd = (size_t)d * d;
    Carver cv;
    const size_t o_pd = cv.add((size_t)M * d), o_kd = cv.add((size_t)M * dd), o_jd = cv.add((size_t)M * dd), o_jp = cv.add((size_t)M * dd), o_od = cv.add((size_t)M * dd);
    if (int rcs = ensure_scratch(h, cv.need)) return rcs;
    auto at = [&](size_t off) { return reinterpret_cast<double*>(h->scratch + off); };
    double *pd = at(o_pd), *kd = at(o_kd), *jd = at(o_jd), *jp = at(o_jp), *od = at(o_od);
    auto cleanup = [&]() { release_big_scratch(h); };
    auto fail = [&](cudaError_t e, int line) {
        cleanup();
        char b[256];
        snprintf(b, sizeof(b), "CUDA error %s at %s:%d", cudaGetErrorString(e), __FILE__, line);
        h->err = b;
        return -2;
    };
    cudaError_t e;
    if ((e = cudaMemcpyAsync(pd, pos, sizeof(double) * M * d, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) return fail(e, __LINE__);
    if ((e = cudaMemcpyAsync(kd, stiff, sizeof(double) * M * dd, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) return fail(e, __LINE__);
    // Jphi(x) = (I + Jpsi(gamma(x))) R: the Jacobian of the whole map at x, as for the velocity (policy_transportation.py:37-46)
    int rc = gptb_query_dev(h, pd, M, GPTB_JAC | GPTB_JPHI | GPTB_AFFINE_IN, nullptr, nullptr, nullptr, jd, nullptr, nullptr, nullptr, nullptr, jp, nullptr);
    if (rc) { cleanup(); return rc; }
    const unsigned grid = (unsigned)((M + 127) / 128);
    if (d == 2) stiffness_transport_kernel<2><<<grid, 128, 0, h->stream>>>(jp, kd, M, od);
    else if (d == 3) stiffness_transport_kernel<3><<<grid, 128, 0, h->stream>>>(jp, kd, M, od);
    else stiffness_transport_kernel<4><<<grid, 128, 0, h->stream>>>(jp, kd, M, od);
    h->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, __LINE__);
    if ((e = cudaMemcpyAsync(stiff_out, od, sizeof(double) * M * dd, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) return fail(e, __LINE__);
    if (jphi && (e = cudaMemcpyAsync(jphi, jp, sizeof(double) * M * dd, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) return fail(e, __LINE__);
    if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return fail(e, __LINE__);
    cleanup();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// exports
// ---------------------------------------------------------------------------------------------------------------
static int export_square(gptb_handle* h, const double* src, double* dst, bool lower_only) {
    const long long N = h->N, Npad = h->Npad;
    CU(h, cudaMemcpy2DAsync(dst, sizeof(double) * N, src, sizeof(double) * Npad, sizeof(double) * N, N, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (lower_only)
        for (long long i = 0; i < N; ++i)
            for (long long j = i + 1; j < N; ++j) dst[i * N + j] = 0.0;
    return 0;
}

extern "C" int gptb_export_L(gptb_handle* h, double* L) {
    if (!h || !L) return -1;
    if (!h->have_factor) GPTB_FAIL(h, -1, "gptb_export_L: model is not fitted");
    if (h->spatial) GPTB_FAIL(h, -1, "gptb_export_L: in spatial mode the factor belongs to the Morton-ordered system; factorise on a handle without gptb_set_spatial for the caller's order");
    CU(h, cudaSetDevice(h->device));
    return export_square(h, h->Lbuf, L, true);
}

extern "C" int gptb_export_alpha(gptb_handle* h, double* alpha) {
    if (!h || !alpha) return -1;
    if (!h->have_alpha) GPTB_FAIL(h, -1, "gptb_export_alpha: model is not fitted");
    CU(h, cudaSetDevice(h->device));
    std::vector<double> tmp((size_t)h->p * h->Npad);
    CU(h, cudaMemcpyAsync(tmp.data(), h->alpha, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->spatial && !h->perm_known) GPTB_FAIL(h, -1, "gptb_export_alpha: this handle received its state by broadcast; export on the fitting rank");
    for (long long n = 0; n < h->N; ++n) {
        const long long dst = h->perm.empty() ? n : h->perm[(size_t)n];
        for (int o = 0; o < h->p; ++o) alpha[dst * h->p + o] = tmp[(size_t)o * h->Npad + n];
    }
    return 0;
}

extern "C" int gptb_export_Kinv(gptb_handle* h, double* Kinv) {
    if (!h || !Kinv) return -1;
    CU(h, cudaSetDevice(h->device));
    if (!h->have_factor) GPTB_FAIL(h, -1, "gptb_export_Kinv: model is not fitted");
    int rc = build_kinv(h);
    if (rc) return rc;
    if (h->perm.empty()) {
        if (h->spatial && !h->perm_known) GPTB_FAIL(h, -1, "gptb_export_Kinv: this handle received its state by broadcast; export on the fitting rank");
        return export_square(h, h->Wbuf, Kinv, false);
    }
    std::vector<double> tmp((size_t)h->N * h->N);
    if ((rc = export_square(h, h->Wbuf, tmp.data(), false))) return rc;
    for (long long i = 0; i < h->N; ++i)
        for (long long j = 0; j < h->N; ++j) Kinv[(long long)h->perm[(size_t)i] * h->N + h->perm[(size_t)j]] = tmp[(size_t)(i * h->N + j)];
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// state exchange for sharded queries
// ---------------------------------------------------------------------------------------------------------------
static int write_header(gptb_handle* h) {
    double hd[32] = {0};
    hd[0] = (double)h->N; hd[1] = h->d; hd[2] = h->p; hd[3] = h->kp.c; hd[4] = h->kp.s2; hd[5] = h->kp.jitter;
    for (int a = 0; a < MAXD; ++a) hd[6 + a] = h->kp.ell[a];
    hd[10] = h->have_minv ? 1.0 : 0.0;
    hd[11] = (double)h->kp.kind;
    hd[12] = (double)h->spatial;
    hd[13] = (double)h->mbox.bits;
    for (int a = 0; a < MAXD; ++a) { hd[14 + a] = h->mbox.lo[a]; hd[18 + a] = h->mbox.scale[a]; }
    CU(h, cudaMemcpyAsync(h->header, hd, sizeof(hd), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int gptb_state_alloc(gptb_handle* h, int64_t N, int d, int p, int with_variance) {
    if (!h) return -1;
    int rc = alloc_model(h, N, d, p, false);
    if (rc) return rc;
    if (with_variance && !h->Minv) {
        CU(h, cudaMalloc(&h->Minv, sizeof(double) * h->Npad * h->Npad));
        MAKE_MAP(h, &h->mapM, h->Minv, h->Npad, h->Npad, h->Npad);
    }
    return 0;
}

extern "C" int gptb_state_buffer(gptb_handle* h, int which, void** dev_ptr, int64_t* bytes) {
    if (!h || !dev_ptr || !bytes) return -1;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    switch (which) {
        case 0:
            if (h->have_alpha && h->have_train) { int rc = write_header(h); if (rc) return rc; }
            *dev_ptr = h->header; *bytes = 32 * sizeof(double); return 0;
        case 1: *dev_ptr = h->X; *bytes = (int64_t)sizeof(double) * h->d * h->Npad; return 0;
        case 2: *dev_ptr = h->alpha; *bytes = (int64_t)sizeof(double) * h->p * h->Npad; return 0;
        case 3: *dev_ptr = h->Minv; *bytes = h->Minv ? (int64_t)sizeof(double) * h->Npad * h->Npad : 0; return 0;
        default: return -1;
    }
}

extern "C" int gptb_state_commit(gptb_handle* h) {
    if (!h) return -1;
    CU(h, cudaSetDevice(h->device));
    double hd[32];
    CU(h, cudaMemcpy(hd, h->header, sizeof(hd), cudaMemcpyDeviceToHost));
    if ((long long)hd[0] != h->N || (int)hd[1] != h->d || (int)hd[2] != h->p) GPTB_FAIL(h, -1, "state header does not match the allocated shape");
    h->kp.kind = (int)hd[11];
    h->spatial = (int)hd[12];
    h->mbox.bits = (int)hd[13];
    for (int a = 0; a < MAXD; ++a) { h->mbox.lo[a] = hd[14 + a]; h->mbox.scale[a] = hd[18 + a]; }
    h->perm.clear();
    h->perm_known = !h->spatial;
    int rc = set_params(h, hd[3], &hd[6], hd[4], hd[5]);
    if (rc) return rc;
    if ((rc = launch_scale(h))) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    h->have_alpha = true;                                   // mean / Jacobian queries are served from (X, alpha)
    h->have_minv = (hd[10] != 0.0) && (h->Minv != nullptr);  // variance queries need the inverse factor
    h->have_kinv = false;
    invalidate_planes(h);                                    // digit planes of a previous model's inverse factor are stale
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// unit-test hooks
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) test_gemm_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                   const __grid_constant__ CUtensorMap mapB, double* C, int nt, int K,
                                                                   int maskA, int maskB) {
    extern __shared__ __align__(128) double smem[];
    __shared__ PipeBarriers pipe;
    pipe_init(&pipe);
    const int tm = blockIdx.x / nt, tn = blockIdx.x % nt;
    double acc[8][4][2];
    acc_clear(acc);
    Operand a{&mapA, tm * TS, 0, maskA, tm};
    Operand b{&mapB, tn * TS, 0, maskB, tn};
    gemm_nt_tile(a, b, 0, K / TS, acc, smem, &pipe);
    double* out = C + (long long)tm * TS * ((long long)nt * TS) + (long long)tn * TS;
    acc_foreach(acc, [&](int r, int c, double v0, double v1) {
        out[(long long)r * nt * TS + c] = v0;
        out[(long long)r * nt * TS + c + 1] = v1;
    });
}

extern "C" int gptb_test_gemm_nt(gptb_handle* h, const double* A, const double* B, double* C, int mt, int nt, int K, int maskA, int maskB) {
    if (!h || !A || !B || !C || mt < 1 || nt < 1 || K < TS || K % TS) return -1;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaFuncSetAttribute(test_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    double *dA, *dB, *dC;
    size_t sa = sizeof(double) * mt * TS * K, sb = sizeof(double) * nt * TS * K, sc = sizeof(double) * mt * TS * nt * TS;
    CU(h, cudaMalloc(&dA, sa));
    CU(h, cudaMalloc(&dB, sb));
    CU(h, cudaMalloc(&dC, sc));
    // stream-ordered copies: the handle's stream is non-blocking, a legacy-stream cudaMemcpy would not order with it
    CU(h, cudaMemcpyAsync(dA, A, sa, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(dB, B, sb, cudaMemcpyHostToDevice, h->stream));
    CUtensorMap mapA, mapB;
    MAKE_MAP(h, &mapA, dA, (long long)mt * TS, K, K);
    MAKE_MAP(h, &mapB, dB, (long long)nt * TS, K, K);
    test_gemm_kernel<<<mt * nt, GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(mapA, mapB, dC, nt, K, maskA, maskB);
    LAUNCH_CHECK(h);
    CU(h, cudaMemcpyAsync(C, dC, sc, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return 0;
}

extern "C" int gptb_test_potrf_tile(gptb_handle* h, const double* A128, double* L128, double* Linv128, int* info) {
    if (!h || !A128 || !L128 || !Linv128 || !info) return -1;
    CU(h, cudaSetDevice(h->device));
    double *dA, *dI;
    CU(h, cudaMalloc(&dA, sizeof(double) * TS * TS));
    CU(h, cudaMalloc(&dI, sizeof(double) * TS * TS));
    CU(h, cudaMemcpyAsync(dA, A128, sizeof(double) * TS * TS, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemsetAsync(h->info, 0, sizeof(int), h->stream));
    long long* prof = nullptr;                   // phase clocks of the diagonal-tile kernel (developer record: gptb_debug_read_profile)
    if (h->oz_prof) prof = reinterpret_cast<long long*>(h->oz_prof);
    potrf_diag_kernel<<<1, 256, DIAG_SMEM_BYTES, h->stream>>>(dA, TS, 0, dI, h->info, nullptr, nullptr, TS, 0, prof);
    LAUNCH_CHECK(h);
    CU(h, cudaMemcpyAsync(L128, dA, sizeof(double) * TS * TS, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(Linv128, dI, sizeof(double) * TS * TS, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(info, h->info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    cudaFree(dA); cudaFree(dI);
    return 0;
}
