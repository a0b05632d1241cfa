// tcgen05 implicit-GEMM convolution for sm_100a (B200): forward and stride-1 dgrad.
//
//   D[128 pixels x BLOCK_N channels] (fp32, TMEM)  +=  A[128 x 64] (bf16, smem)  *  B[BLOCK_N x 64]^T
//
// * A tiles are produced by TMA in IM2COL mode straight from the NHWC bf16 activation tensor: one
//   instruction loads 128 consecutive output pixels (flattened over n,oy,ox — rows and images wrap in
//   hardware) x 64 input channels for one filter tap, zero-filling the padding halo.  No im2col buffer,
//   no tile waste on 54x54 / 27x27 / 14x14 feature maps.
// * B tiles are TMA tiled loads of the packed weights [tap][Cout][Cin] (K-major), 128B-swizzled.
// * One elected thread issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=BLOCK_N, K=16);
//   accumulators are double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the
//   MMAs of tile i+1; the kernel is persistent (grid = #SMs) with a static round-robin tile schedule.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue
//   (tcgen05.ld 32x32b -> bias / LeakyReLU|ReLU / residual / PixelShuffle scatter -> 16B global stores).
//
// dgrad (stride 1) is the same kernel on dy with weights packed [tap][Cin][Cout] and mirrored tap
// offsets (flip=1).  Reference call sites: every 3x3/1x1 nn.Conv2d with Cin % 64 == 0 in
// model/sradsgan.py (RAB :222-223, GAB_UP :375,:381, Discriminator :476, VGG19 features).
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
static int make_im2col_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int kh, int kw, int pad,
                           int stride, int pixels) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    int lower[2] = {-pad, -pad};                       // {W, H}: first filter-window origin
    int upper[2] = {pad - (kw - 1), pad - (kh - 1)};   // last origin, relative to the far edge
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = g_encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides,
                                 lower, upper, 64, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeIm2col failed (%d) N=%d H=%d W=%d C=%d k=%d pad=%d stride=%d", (int)r, N, H, W, C, kh, pad, stride);
        return SR_ERR_CUDA;
    }
    // Known driver issue with im2col descriptors of tensors smaller than 128 KiB on drivers <= 13.1
    // (same workaround as CUTLASS make_im2col_tma_copy_desc): clear bit 21 of the second descriptor word.
    if (g_driver_version <= 13010 && (size_t)N * H * W * C * 2 < 131072)
        reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
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
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
struct TcParams {
    int M_total, Ho, Wo, Cout;
    int block_n, n_blocks, num_tiles;
    int kh, kw, stride, pad;
    int c_blocks;      // Cin / 64
    int flip;          // mirror tap offsets (dgrad)
    int act;
    float slope;
    int shuffle_r;     // weights rows are packed subpixel-major when > 1
    int num_stages;
    const float* bias;
    const void* residual;
    void* out;
};

constexpr int TC_THREADS = 192;
constexpr int TC_A_BYTES = 128 * 128;   // 128 pixels x 64 bf16
constexpr int TC_ACC_STRIDE = 256;      // TMEM columns per accumulator stage
constexpr int TC_MAX_STAGES = 8;

template <typename OutT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = p.block_n * 128;
    const int stage_bytes = TC_A_BYTES + b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.num_stages * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + TC_MAX_STAGES;
    uint64_t* acc_full = bars + 2 * TC_MAX_STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k_blocks = p.kh * p.kw * p.c_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < p.num_stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(acc_full + a, 1); mbar_init(acc_empty + a, 128); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int m_tile = tile / p.n_blocks, nb = tile - m_tile * p.n_blocks;
                const int m0 = m_tile * 128;
                const int ox = m0 % p.Wo; const int q = m0 / p.Wo;
                const int oy = q % p.Ho; const int n = q / p.Ho;
                const int w0 = ox * p.stride - p.pad, h0 = oy * p.stride - p.pad;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    const int tap = kb / p.c_blocks, cb = kb - tap * p.c_blocks;
                    const int ky = tap / p.kw, kx = tap - ky * p.kw;
                    const int offw = p.flip ? (p.kw - 1 - kx) : kx, offh = p.flip ? (p.kh - 1 - ky) : ky;
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    mbar_expect_tx(full + stage, (uint32_t)stage_bytes);
                    tma_load_im2col(sa, &map_a, full + stage, cb * 64, w0, h0, n, (uint16_t)offw, (uint16_t)offh);
                    tma_load_2d(sa + TC_A_BYTES, &map_b, full + stage, cb * 64, tap * p.Cout + nb * p.block_n);
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N=block_n, M=128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(acc_empty + acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_ACC_STRIDE);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                    const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + TC_A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)   // 4 x K=16 per 64-channel block: +32 B per step
                        umma_f16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                    umma_commit(empty + stage);
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(acc_full + acc);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int r = p.shuffle_r > 1 ? p.shuffle_r : 1;
        const int r2 = r * r;
        const int cq = p.Cout / r2;
        OutT* out = reinterpret_cast<OutT*>(p.out);
        const OutT* res = reinterpret_cast<const OutT*>(p.residual);
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int m_tile = tile / p.n_blocks, nb = tile - m_tile * p.n_blocks;
            const long long m = (long long)m_tile * 128 + row;
            const bool valid = m < p.M_total;
            int ox = 0, oy = 0, n = 0;
            if (r > 1 && valid) {
                ox = (int)(m % p.Wo); const long long q = m / p.Wo;
                oy = (int)(q % p.Ho); n = (int)(q / p.Ho);
            }
            mbar_wait(acc_full + acc, acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TC_ACC_STRIDE);
            for (int c0 = 0; c0 < p.block_n; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(t_base + (uint32_t)c0, v);
                if (!valid) continue;
                const int col = nb * p.block_n + c0;     // packed (GEMM) output column of v[0]
                long long idx;
                int sub = 0, ch0 = col;
                if (r > 1) {
                    sub = col / cq; ch0 = col - sub * cq;
                    const int si = sub / r, sj = sub - si * r;
                    idx = ((((long long)n * p.Ho * r + (oy * r + si)) * ((long long)p.Wo * r)) + (ox * r + sj)) * cq + ch0;
                } else {
                    idx = m * p.Cout + col;
                }
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = __uint_as_float(v[j]);
                    if (p.bias) x += p.bias[r > 1 ? (ch0 + j) * r2 + sub : col + j];
                    f[j] = apply_act(x, p.act, p.slope);
                }
                if (sizeof(OutT) == 2) {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + idx;
                    if (res) {
                        const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(res) + idx);
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
                    float* o = reinterpret_cast<float*>(out) + idx;
                    if (res) {
                        const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(res) + idx);
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
            tc_fence_before();
            mbar_arrive(acc_empty + acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int g_num_sms = 0;

static int pick_block_n(int Cout) {
    if (Cout % 256 == 0) return 128;     // 2+ column blocks per pixel tile -> better wave balance
    if (Cout % 192 == 0) return 192;
    if (Cout % 128 == 0) return 128;
    if (Cout % 64 == 0) return 64;
    return 0;
}

bool conv_tc_supported(const sr_conv_desc* d, bool dgrad) {
    if (d->in_dtype != SR_BF16) return false;
    const int Cs = dgrad ? d->Cout : d->Cin, Cd = dgrad ? d->Cin : d->Cout;
    if (Cs % 64 != 0 || pick_block_n(Cd) == 0) return false;
    if (d->kh != d->kw || d->kh > 7) return false;
    if (dgrad && d->stride != 1) return false;
    if (d->stride < 1 || d->stride > 2) return false;
    const int r = d->shuffle_r > 1 ? d->shuffle_r : 1;
    if (r > 1 && (Cd % (r * r) != 0 || (Cd / (r * r)) % 32 != 0)) return false;
    if ((long long)d->N * d->Ho * d->Wo >= (1ll << 31)) return false;
    return true;
}

// Runs y = epilogue(conv(src, w)) on the tensor cores.  For dgrad the caller passes the forward desc;
// src = dy (N,Ho,Wo,Cout), dst = dx (N,H,W,Cin), weights packed [tap][Cin][Cout].
int conv_tc_run(const sr_conv_desc* d, bool dgrad, const void* src, const void* w, const float* bias,
                const void* residual, void* dst, cudaStream_t st) {
    int rc = load_driver_fns();
    if (rc != SR_OK) return rc;
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int Hs = dgrad ? d->Ho : d->H, Ws = dgrad ? d->Wo : d->W, Cs = dgrad ? d->Cout : d->Cin;
    const int Hd = dgrad ? d->H : d->Ho, Wd = dgrad ? d->W : d->Wo, Cd = dgrad ? d->Cin : d->Cout;
    const int pad = dgrad ? (d->kh - 1 - d->pad) : d->pad;
    const int stride = dgrad ? 1 : d->stride;
    const int taps = d->kh * d->kw;

    TcParams p;
    p.M_total = d->N * Hd * Wd; p.Ho = Hd; p.Wo = Wd; p.Cout = Cd;
    p.block_n = pick_block_n(Cd);
    p.n_blocks = Cd / p.block_n;
    p.num_tiles = (int)cdiv(p.M_total, 128) * p.n_blocks;
    p.kh = d->kh; p.kw = d->kw; p.stride = stride; p.pad = pad;
    p.c_blocks = Cs / 64;
    p.flip = dgrad ? 1 : 0;
    p.act = dgrad ? SR_ACT_NONE : d->act; p.slope = d->slope;
    p.shuffle_r = dgrad ? 0 : d->shuffle_r;
    p.bias = bias; p.residual = residual; p.out = dst;
    const int stage_bytes = TC_A_BYTES + p.block_n * 128;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    p.num_stages = stages;
    const size_t smem = 1024 + (size_t)stages * stage_bytes + 256;

    alignas(64) CUtensorMap map_a, map_b;
    rc = make_im2col_map(&map_a, src, d->N, Hs, Ws, Cs, d->kh, d->kw, pad, stride, 128);
    if (rc != SR_OK) return rc;
    rc = make_tiled2d_map(&map_b, w, (uint64_t)taps * Cd, (uint64_t)Cs, (uint32_t)p.block_n);
    if (rc != SR_OK) return rc;

    const int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
    const bool out_bf16 = d->out_dtype == SR_BF16;
    static bool attr_set[2] = {false, false};
    if (out_bf16) {
        if (!attr_set[0]) { cudaFuncSetAttribute(conv_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_set[0] = true; }
        conv_tc_kernel<__nv_bfloat16><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, p);
    } else {
        if (!attr_set[1]) { cudaFuncSetAttribute(conv_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_set[1] = true; }
        conv_tc_kernel<float><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, p);
    }
    count_launch();
    return check_launch("conv_tc_kernel");
}

}  // namespace sr
