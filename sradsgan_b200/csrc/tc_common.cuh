// Shared tcgen05 / TMA / mbarrier PTX wrappers and tensor-map builders (sm_100a).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace sr {

// ------------------------------------------------------------------------------------------------
// driver entry points (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;

static int load_driver_fns() {
    if (g_encode_tiled && g_encode_im2col) return SR_OK;
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return SR_ERR_CUDA;
    }
    g_encode_tiled = (EncodeTiledFn)f;
    f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
        set_error("cuTensorMapEncodeIm2col not available from the driver");
        return SR_ERR_CUDA;
    }
    g_encode_im2col = (EncodeIm2colFn)f;
    cudaDriverGetVersion(&g_driver_version);
    return SR_OK;
}

// NHWC bf16 activation tensor -> im2col tensor map: box = `pixels` output positions x 64 channels.
static int make_im2col_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, const int* lower_wh,
                           const int* upper_wh, int stride, int pixels) {
    // lower_wh / upper_wh: {W, H} offsets of the first / last filter-window origin from the near / far edge
    // (forward 3x3 pad 1: lower = {-1,-1}, upper = {-1,-1})
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    int lower[2] = {lower_wh[0], lower_wh[1]};
    int upper[2] = {upper_wh[0], upper_wh[1]};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = g_encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides,
                                 lower, upper, 64, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeIm2col failed (%d) N=%d H=%d W=%d C=%d lower=(%d,%d) upper=(%d,%d) stride=%d", (int)r, N, H, W, C,
                  lower[0], lower[1], upper[0], upper[1], stride);
        return SR_ERR_CUDA;
    }
    // Known driver issue with im2col descriptors of tensors smaller than 128 KiB on drivers <= 13.1
    // (same workaround as CUTLASS make_im2col_tma_copy_desc): clear bit 21 of the second descriptor word.
    if (g_driver_version <= 13010 && (size_t)N * H * W * C * 2 < 131072)
        reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
    return SR_OK;
}

// NHWC bf16 activation tensor -> plain tiled 4-d map with box [1][box_h][box_w][64 ch], 128B swizzle
static int make_tiled4d_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int box_w, int box_h) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(4d) failed (%d) N=%d H=%d W=%d C=%d box=(%d,%d)", (int)r, N, H, W, C, box_w, box_h);
        return SR_ERR_CUDA;
    }
    return SR_OK;
}

// row-major bf16 matrix [rows][cols] -> tiled map with box [box_rows][64], 128B swizzle.
static int make_tiled2d_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box_rows=%u", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, box_rows);
        return SR_ERR_CUDA;
    }
    return SR_OK;
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_im2col(void* dst, const CUtensorMap* map, uint64_t* bar, int c, int w, int h,
                                                int n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        :
        : "r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// 2-d tiled load whose box lands at the SAME shared-memory offset in every CTA of `cta_mask` (thread-block cluster), each
// destination CTA's mbarrier at `bar`'s offset receiving the complete_tx
__device__ __forceinline__ void tma_load_2d_mcast(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
        :
        : "r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "h"(cta_mask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// all threads of all CTAs of the cluster
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the four K=16 steps of one 64-deep K-major SW128 k-block (descriptor start address +32 B per step) in ONE asm
// statement: a single operand hand-off to the uniform datapath instead of four
__device__ __forceinline__ void umma_f16_x4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, t;\n\t.reg .b64 a1, b1, a2, b2, a3, b3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, 2;\n\t"
        "add.s64 a2, %1, 4;\n\tadd.s64 b2, %2, 4;\n\t"
        "add.s64 a3, %1, 6;\n\tadd.s64 b3, %2, 6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// four consecutive K=16 steps of MN-major SW128 operands (16 pixel rows = 2048 B = 128 descriptor units per step)
__device__ __forceinline__ void umma_f16_x4_mn(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, t;\n\t.reg .b64 a1, b1, a2, b2, a3, b3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "add.s64 a1, %1, 128;\n\tadd.s64 b1, %2, 128;\n\t"
        "add.s64 a2, %1, 256;\n\tadd.s64 b2, %2, 256;\n\t"
        "add.s64 a3, %1, 384;\n\tadd.s64 b3, %2, 384;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the same arrival on the mbarrier at `bar`'s offset in EVERY CTA of cta_mask (releases a multicast weight stage in both CTAs)
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile ([rows][64 bf16], 8-row groups 1024 B apart): UMMA smem descriptor
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);    // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}


// 4-d tiled TMA load (coordinates may be negative / out of range: zero filled)
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :
        : "r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// waits for every outstanding tcgen05.ld of this thread; the registers are passed through so that the
// compiler cannot move their first use above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                   "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                   "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

template <int ACT> __device__ __forceinline__ float tc_act(float x, float slope) {
    if (ACT == SR_ACT_LRELU) return fmaxf(x, x * slope);          // slope in [0, 1]
    if (ACT == SR_ACT_RELU) return fmaxf(x, 0.f);
    if (ACT == SR_ACT_SIGMOID) return 1.f / (1.f + __expf(-x));
    return x;
}

// bias (shared memory, already in packed-column order) -> activation -> (+ residual) -> 16-byte stores of one
// 32-column accumulator chunk of one output row
// 16-byte load from SHARED memory through an explicit shared-space instruction (a generic pointer would compile to
// a generic LD with global-memory-like scoreboard latency in the epilogue's critical path)
__device__ __forceinline__ float4 lds128(const float* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}

// `mask` (nullable, bf16, same indexing as the output): f *= (mask > 0 ? 1 : mask_slope) — the derivative of the
// LeakyReLU / ReLU that produced `mask`, fused into the epilogue of the input-gradient convolution behind it.
template <typename OutT, int ACT>
__device__ __forceinline__ void tc_store_chunk(const uint32_t (&v)[32], const float* bias_s, float slope, const OutT* res, OutT* o,
                                               const __nv_bfloat16* mask = nullptr, float mask_slope = 0.f) {
    float f[32];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const float4 b = lds128(bias_s + g * 4);
        f[g * 4 + 0] = tc_act<ACT>(__uint_as_float(v[g * 4 + 0]) + b.x, slope);
        f[g * 4 + 1] = tc_act<ACT>(__uint_as_float(v[g * 4 + 1]) + b.y, slope);
        f[g * 4 + 2] = tc_act<ACT>(__uint_as_float(v[g * 4 + 2]) + b.z, slope);
        f[g * 4 + 3] = tc_act<ACT>(__uint_as_float(v[g * 4 + 3]) + b.w, slope);
    }
    if (mask) {
        const uint4* mp = reinterpret_cast<const uint4*>(mask);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint4 mv = mp[g];
            const __nv_bfloat162* m2 = reinterpret_cast<const __nv_bfloat162*>(&mv);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                if (!(__low2float(m2[h]) > 0.f)) f[g * 8 + h * 2] *= mask_slope;
                if (!(__high2float(m2[h]) > 0.f)) f[g * 8 + h * 2 + 1] *= mask_slope;
            }
        }
    }
    if (sizeof(OutT) == 2) {
        if (res) {
            const uint4* rp = reinterpret_cast<const uint4*>(res);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint4 rv = rp[g];
                const __nv_bfloat162* r2p = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    f[g * 8 + h * 2] += __low2float(r2p[h]);
                    f[g * 8 + h * 2 + 1] += __high2float(r2p[h]);
                }
            }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint4 w;
            __nv_bfloat162 b0 = __floats2bfloat162_rn(f[g * 8 + 0], f[g * 8 + 1]);
            __nv_bfloat162 b1 = __floats2bfloat162_rn(f[g * 8 + 2], f[g * 8 + 3]);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(f[g * 8 + 4], f[g * 8 + 5]);
            __nv_bfloat162 b3 = __floats2bfloat162_rn(f[g * 8 + 6], f[g * 8 + 7]);
            w.x = *reinterpret_cast<uint32_t*>(&b0); w.y = *reinterpret_cast<uint32_t*>(&b1);
            w.z = *reinterpret_cast<uint32_t*>(&b2); w.w = *reinterpret_cast<uint32_t*>(&b3);
            reinterpret_cast<uint4*>(o)[g] = w;
        }
    } else {
        if (res) {
            const float4* rp = reinterpret_cast<const float4*>(res);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const float4 rv = rp[g];
                f[g * 4] += rv.x; f[g * 4 + 1] += rv.y; f[g * 4 + 2] += rv.z; f[g * 4 + 3] += rv.w;
            }
        }
#pragma unroll
        for (int g = 0; g < 8; ++g)
            reinterpret_cast<float4*>(o)[g] = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
    }
}

// Row-coalesced variant of tc_store_chunk for bf16 outputs.  With lane = accumulator row, a 16-byte store instruction of the
// plain variant touches 32 different 128-byte lines (one per row); the L1/LSU path handles one line per cycle and shares its
// data path with the tensor core's shared-memory operand fetch, so those stores both dominated the epilogue and halved the
// MMA rate next to it (profiles/r02_halo_trace.txt).  Here the four 16-byte units of a lane's row chunk are transposed inside
// each lane quad (two butterfly steps, 16 shuffles), after which store instruction i writes, for every quad, the 64
// contiguous bytes of row 4k+i from its four lanes: 8 lines per instruction instead of 32.
//   q_ptr[i]  : address of this chunk's first column in row (lane & ~3) + i, already offset by (lane & 3) * 8 elements
//   q_ok      : bit i = that row is stored
//   res / mask: as in tc_store_chunk, for the lane's OWN row (nullptr when the row is not stored)
// Every lane of the warp must call (shuffles).
template <int ACT>
__device__ __forceinline__ void tc_store_chunk_quads(const uint32_t (&v)[32], const float* bias_s, float slope, const __nv_bfloat16* res,
                                                     const __nv_bfloat16* mask, float mask_slope, __nv_bfloat16* const (&q_ptr)[4],
                                                     unsigned q_ok, int lane) {
    float f[32];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const float4 b = lds128(bias_s + g * 4);
        f[g * 4 + 0] = tc_act<ACT>(__uint_as_float(v[g * 4 + 0]) + b.x, slope);
        f[g * 4 + 1] = tc_act<ACT>(__uint_as_float(v[g * 4 + 1]) + b.y, slope);
        f[g * 4 + 2] = tc_act<ACT>(__uint_as_float(v[g * 4 + 2]) + b.z, slope);
        f[g * 4 + 3] = tc_act<ACT>(__uint_as_float(v[g * 4 + 3]) + b.w, slope);
    }
    if (mask) {
        const uint4* mp = reinterpret_cast<const uint4*>(mask);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint4 mv = mp[g];
            const __nv_bfloat162* m2 = reinterpret_cast<const __nv_bfloat162*>(&mv);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                if (!(__low2float(m2[h]) > 0.f)) f[g * 8 + h * 2] *= mask_slope;
                if (!(__high2float(m2[h]) > 0.f)) f[g * 8 + h * 2 + 1] *= mask_slope;
            }
        }
    }
    if (res) {
        const uint4* rp = reinterpret_cast<const uint4*>(res);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint4 rv = rp[g];
            const __nv_bfloat162* r2p = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                f[g * 8 + h * 2] += __low2float(r2p[h]);
                f[g * 8 + h * 2 + 1] += __high2float(r2p[h]);
            }
        }
    }
    uint32_t w[16];                                   // unit u (16 bytes = columns 8u .. 8u+7) = w[4u .. 4u+3]
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&b);
    }
    // step 1 (partner = lane ^ 2): afterwards R[b][e] = unit 2*((lane >> 1) & 1) + e of the quad row whose bit 1 is b
    const bool hi = (lane & 2) != 0, odd = (lane & 1) != 0;
    uint32_t R[2][2][4];
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t lo_u = w[e * 4 + k], hi_u = w[(2 + e) * 4 + k];
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, hi ? lo_u : hi_u, 2);
            R[0][e][k] = hi ? recv : lo_u;
            R[1][e][k] = hi ? hi_u : recv;
        }
    // step 2 (partner = lane ^ 1): T[2b + c] = unit (lane & 3) of quad row 2b + c
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        uint32_t t0[4], t1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? R[b][0][k] : R[b][1][k], 1);
            const uint32_t keep = odd ? R[b][1][k] : R[b][0][k];
            t0[k] = odd ? recv : keep;
            t1[k] = odd ? keep : recv;
        }
        if ((q_ok >> (2 * b)) & 1) *reinterpret_cast<uint4*>(q_ptr[2 * b]) = make_uint4(t0[0], t0[1], t0[2], t0[3]);
        if ((q_ok >> (2 * b + 1)) & 1) *reinterpret_cast<uint4*>(q_ptr[2 * b + 1]) = make_uint4(t1[0], t1[1], t1[2], t1[3]);
    }
}

// MN-major, 128B-swizzled operand tile ([k rows][64 bf16 of M/N], 8-row groups 1024 B apart, further
// 64-wide M/N panels `lbo_bytes` apart): UMMA smem descriptor (cute::UMMA canonical MN-major SW128 layout)
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

}  // namespace sr
