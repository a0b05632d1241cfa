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
#include <string.h>

#include "tc_common.cuh"

namespace sr {

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
struct TcParams {
    int M_total, Ho, Wo, Cout;
    int block_n, n_blocks, m_tiles, num_tiles;
    int stride, pad_w, pad_h;   // im2col base coordinate of output pixel (oy,ox) = (oy*stride - pad_h, ox*stride - pad_w)
    int c_blocks;      // Cin / 64
    int ntaps;         // filter taps visited (a subset for the parity classes of a strided dgrad)
    unsigned char tap_w[49], tap_ow[49], tap_oh[49];   // weight-tap index / im2col offsets per visited tap
    int os, py, px, Hfull, Wfull;   // os > 1: row (n,oy,ox) is written to pixel (oy*os+py, ox*os+px) of an Hfull x Wfull map
    int act;
    float slope;
    int shuffle_r;     // weights rows are packed subpixel-major when > 1
    int num_stages;
    int resident;      // 1: every CTA keeps the weights of ONE column block in shared memory for all of its tiles
    const float* bias;
    const void* residual;
    void* out;
};

constexpr int TC_EPI_WARPS = 8;                      // 2 per TMEM lane quarter
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int TC_A_BYTES = 128 * 128;                // 128 pixels x 64 bf16
constexpr int TC_ACC_STRIDE = 256;                   // TMEM columns per accumulator stage
constexpr int TC_MAX_STAGES = 12;
constexpr int TC_MAX_RES_KB = 36;                    // resident mode: at most this many 64-deep k-blocks
constexpr int TC_BIAS_MAX = 1024;                    // floats of bias staged in shared memory
constexpr int TC_TILE_BUDGET = 232448 - 1024 - 1024 - TC_BIAS_MAX * 4;   // bytes available for operand tiles

// tile `it` of this CTA -> (pixel tile, column block).  Resident mode pins the column block to the CTA.
__device__ __forceinline__ bool tc_tile_at(const TcParams& p, int it, int& m_tile, int& nb) {
    if (p.resident) {
        nb = blockIdx.x % p.n_blocks;
        m_tile = blockIdx.x / p.n_blocks + it * (gridDim.x / p.n_blocks);
        return m_tile < p.m_tiles;
    }
    const int t = blockIdx.x + it * gridDim.x;
    if (t >= p.num_tiles) return false;
    m_tile = t / p.n_blocks;
    nb = t - m_tile * p.n_blocks;
    return true;
}

// Epilogue warp: TMEM lane quarter `quarter`, every second 32-column chunk starting at `half`; the tcgen05.ld of
// the next chunk is in flight while the current one is converted and stored.
template <typename OutT, int ACT>
__device__ __forceinline__ void tc_epilogue(const TcParams& p, uint32_t tmem_base, uint64_t* acc_full, uint64_t* acc_empty,
                                            const float* bias_s, int quarter, int half, int lane) {
    const int row = quarter * 32 + lane;
    const int r = p.shuffle_r > 1 ? p.shuffle_r : 1;
    const int cq = p.Cout / (r * r);
    const int chunks = p.block_n >> 5;
    OutT* out = reinterpret_cast<OutT*>(p.out);
    const OutT* res = reinterpret_cast<const OutT*>(p.residual);
    int acc = 0; uint32_t acc_phase = 0;
    int m_tile, nb;
    for (int it = 0; tc_tile_at(p, it, m_tile, nb); ++it) {
        const long long m = (long long)m_tile * 128 + row;
        const bool valid = m < p.M_total;
        int ox = 0, oy = 0, n = 0;
        if ((r > 1 || p.os > 1) && valid) {
            ox = (int)(m % p.Wo); const long long q = m / p.Wo;
            oy = (int)(q % p.Ho); n = (int)(q / p.Ho);
        }
        long long row_idx;     // element index of column 0 of this row (plain / strided-scatter modes)
        if (p.os > 1) row_idx = (((long long)n * p.Hfull + (oy * p.os + p.py)) * p.Wfull + (ox * p.os + p.px)) * p.Cout;
        else row_idx = m * p.Cout;
        // bf16 outputs: row-coalesced stores (tc_store_chunk_quads, tc_common.cuh) — the row bases of the lane's quad, once per tile
        const long long row_base = r > 1 ? ((((long long)n * p.Ho * r + (long long)oy * r) * ((long long)p.Wo * r)) + (long long)ox * r) * cq : row_idx;
        long long q_base[4] = {0, 0, 0, 0};
        unsigned q_ok = 0;
        if (sizeof(OutT) == 2) {
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            q_ok = (vm >> (lane & ~3)) & 0xFu;
#pragma unroll
            for (int i = 0; i < 4; ++i) q_base[i] = __shfl_sync(0xffffffffu, row_base, (lane & ~3) + i) + (lane & 3) * 8;
        }
        mbar_wait(acc_full + acc, acc_phase);
        tc_fence_after();
        const uint32_t t_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TC_ACC_STRIDE);
        auto emit = [&](const uint32_t (&v)[32], int c) {
            const int col = nb * p.block_n + c * 32;     // packed (GEMM) output column of v[0]
            int chunk_off = col;                         // offset of this chunk's first column from the row base (the same for every row)
            if (r > 1) {
                const int sub = col / cq, ch0 = col - sub * cq;
                const int si = sub / r, sj = sub - si * r;
                chunk_off = (si * p.Wo * r + sj) * cq + ch0;
            }
            const long long idx = row_base + chunk_off;
            if (sizeof(OutT) == 2) {                     // every lane takes part (shuffles); invalid rows store nothing
                __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out);
                __nv_bfloat16* const qp[4] = {ob + q_base[0] + chunk_off, ob + q_base[1] + chunk_off, ob + q_base[2] + chunk_off, ob + q_base[3] + chunk_off};
                tc_store_chunk_quads<ACT>(v, bias_s + col, p.slope, (res && valid) ? reinterpret_cast<const __nv_bfloat16*>(res) + idx : nullptr,
                                          nullptr, 0.f, qp, q_ok, lane);
                return;
            }
            if (!valid) return;
            tc_store_chunk<OutT, ACT>(v, bias_s + col, p.slope, res ? res + idx : nullptr, out + idx);
        };
        uint32_t va[32], vb[32];
        int c = half;
        if (c < chunks) tmem_ld32_nowait(t_base + (uint32_t)(c * 32), va);
        while (c < chunks) {
            tmem_ld_wait(va);
            if (c + 2 < chunks) tmem_ld32_nowait(t_base + (uint32_t)((c + 2) * 32), vb);
            emit(va, c);
            c += 2;
            if (c >= chunks) break;
            tmem_ld_wait(vb);
            if (c + 2 < chunks) tmem_ld32_nowait(t_base + (uint32_t)((c + 2) * 32), va);
            emit(vb, c);
            c += 2;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
}

template <typename OutT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = p.block_n * 128;
    const int k_blocks = p.ntaps * p.c_blocks;
    // resident: [k_blocks x B][stages x A] ; streaming: [stages x (A + B)]
    const int stage_bytes = p.resident ? TC_A_BYTES : TC_A_BYTES + b_bytes;
    uint8_t* stage_base = smem + (p.resident ? (size_t)k_blocks * b_bytes : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + (size_t)p.num_stages * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + TC_MAX_STAGES;
    uint64_t* acc_full = bars + 2 * TC_MAX_STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* b_full = acc_empty + 2;                                    // [TC_MAX_RES_KB]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + TC_MAX_RES_KB);
    float* bias_s = reinterpret_cast<float*>(bars) + 256;                 // 1024 B after the barrier block

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < p.num_stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(acc_full + a, 1); mbar_init(acc_empty + a, TC_EPI_WARPS); }
        if (p.resident) for (int kb = 0; kb < k_blocks; ++kb) mbar_init(b_full + kb, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    pdl_trigger();
    pdl_wait();                    // (common.cuh) global memory is read from here on
    {   // bias in packed-column order (subpixel-major when the epilogue does the PixelShuffle)
        const int r2 = p.shuffle_r > 1 ? p.shuffle_r * p.shuffle_r : 1;
        const int cq = p.Cout / r2;
        for (int i = threadIdx.x; i < p.Cout; i += TC_THREADS) {
            float b = 0.f;
            if (p.bias) { const int sub = i / cq, ch = i - sub * cq; b = p.bias[r2 > 1 ? ch * r2 + sub : i]; }
            bias_s[i] = b;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int m_tile, nb;
            for (int it = 0; tc_tile_at(p, it, m_tile, nb); ++it) {
                const int m0 = m_tile * 128;
                const int ox = m0 % p.Wo; const int q = m0 / p.Wo;
                const int oy = q % p.Ho; const int n = q / p.Ho;
                const int w0 = ox * p.stride - p.pad_w, h0 = oy * p.stride - p.pad_h;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    const int ti = kb / p.c_blocks, cb = kb - ti * p.c_blocks;
                    const int tap = p.tap_w[ti], offw = p.tap_ow[ti], offh = p.tap_oh[ti];
                    if (p.resident && it == 0) {      // weights of this CTA's column block: loaded once, k-block by k-block
                        mbar_expect_tx(b_full + kb, (uint32_t)b_bytes);
                        tma_load_2d(smem + (size_t)kb * b_bytes, &map_b, b_full + kb, cb * 64, tap * p.Cout + nb * p.block_n);
                    }
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t* sa = stage_base + (size_t)stage * stage_bytes;
                    mbar_expect_tx(full + stage, (uint32_t)stage_bytes);
                    tma_load_im2col(sa, &map_a, full + stage, cb * 64, w0, h0, n, (uint16_t)offw, (uint16_t)offh);
                    if (!p.resident) tma_load_2d(sa + TC_A_BYTES, &map_b, full + stage, cb * 64, tap * p.Cout + nb * p.block_n);
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the schedule converged, lane 0 issues =====
        // instruction descriptor: D=f32, A=B=bf16, both K-major, N=block_n, M=128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((128u >> 4) << 24);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        int m_tile, nb;
        for (int it = 0; tc_tile_at(p, it, m_tile, nb); ++it) {
            mbar_wait(acc_empty + acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_ACC_STRIDE);
            for (int kb = 0; kb < k_blocks; ++kb) {
                if (p.resident && it == 0) mbar_wait(b_full + kb, 0);
                mbar_wait(full + stage, phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(stage_base + (size_t)stage * stage_bytes);
                    const uint32_t b_addr = p.resident ? smem_u32(smem + (size_t)kb * b_bytes) : a_addr + TC_A_BYTES;
                    umma_f16_x4(d_tmem, make_kmajor_sw128_desc(a_addr), make_kmajor_sw128_desc(b_addr), idesc, kb ? 1u : 0u);
                    umma_commit(empty + stage);
                }
                __syncwarp();
                if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
            }
            if (lane == 0) umma_commit(acc_full + acc);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===== epilogue: warps 2..9; TMEM lane quarter = warp % 4, column-chunk parity = (warp - 2) / 4 =====
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        switch (p.act) {
            case SR_ACT_LRELU: tc_epilogue<OutT, SR_ACT_LRELU>(p, tmem_base, acc_full, acc_empty, bias_s, quarter, half, lane); break;
            case SR_ACT_RELU: tc_epilogue<OutT, SR_ACT_RELU>(p, tmem_base, acc_full, acc_empty, bias_s, quarter, half, lane); break;
            case SR_ACT_SIGMOID: tc_epilogue<OutT, SR_ACT_SIGMOID>(p, tmem_base, acc_full, acc_empty, bias_s, quarter, half, lane); break;
            default: tc_epilogue<OutT, SR_ACT_NONE>(p, tmem_base, acc_full, acc_empty, bias_s, quarter, half, lane); break;
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
    if (Cs % 64 != 0 || pick_block_n(Cd) == 0 || Cd > TC_BIAS_MAX) return false;
    if (d->kh != d->kw || d->kh > 7) return false;
    if (d->stride < 1 || d->stride > 2) return false;
    if (dgrad && d->stride == 2 && !(d->kh == 3 && d->pad == 1)) return false;   // parity decomposition: 3x3/pad 1 only
    const int r = d->shuffle_r > 1 ? d->shuffle_r : 1;
    if (!dgrad && r > 1 && (Cd % (r * r) != 0 || (Cd / (r * r)) % 32 != 0)) return false;
    if ((long long)d->N * d->Ho * d->Wo >= (1ll << 31) || (long long)d->N * d->H * d->W >= (1ll << 31)) return false;
    return true;
}

// Completes the schedule of one launch (tiles, resident-weights mode, pipeline depth, grid) and launches.
static int launch_tc(TcParams p, const CUtensorMap& map_a, const CUtensorMap& map_b, bool out_bf16, cudaStream_t st) {
    const int b_bytes = p.block_n * 128;
    const int k_blocks = p.ntaps * p.c_blocks;
    p.m_tiles = (int)cdiv(p.M_total, 128);
    p.num_tiles = p.m_tiles * p.n_blocks;
    int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
    // Resident weights: when the [taps x Cin] x block_n slab of one column block fits beside >= 3 activation
    // stages, each CTA loads it ONCE and streams only activation tiles (halves the L2->SM traffic of Cin=64 convs).
    const long long res_bytes = (long long)k_blocks * b_bytes;
    p.resident = (k_blocks <= TC_MAX_RES_KB && res_bytes + 3 * TC_A_BYTES <= TC_TILE_BUDGET && p.num_tiles > grid) ? 1 : 0;
    int stages;
    size_t tile_bytes;
    if (p.resident) {
        grid -= grid % p.n_blocks;
        stages = (int)((TC_TILE_BUDGET - res_bytes) / TC_A_BYTES);
        if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
        tile_bytes = (size_t)res_bytes + (size_t)stages * TC_A_BYTES;
    } else {
        stages = TC_TILE_BUDGET / (TC_A_BYTES + b_bytes);
        if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
        tile_bytes = (size_t)stages * (TC_A_BYTES + b_bytes);
    }
    p.num_stages = stages;
    const size_t smem = 1024 + tile_bytes + 1024 + TC_BIAS_MAX * 4;
    static bool attr_set[2] = {false, false};
    if (out_bf16) {
        if (!attr_set[0]) { cudaFuncSetAttribute(conv_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448); attr_set[0] = true; }
        launch_pdl(conv_tc_kernel<__nv_bfloat16>, dim3(grid), dim3(TC_THREADS), smem, st, option("SR_PDL", 0) != 0, map_a, map_b, p);
    } else {
        if (!attr_set[1]) { cudaFuncSetAttribute(conv_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448); attr_set[1] = true; }
        launch_pdl(conv_tc_kernel<float>, dim3(grid), dim3(TC_THREADS), smem, st, option("SR_PDL", 0) != 0, map_a, map_b, p);
    }
    count_launch();
    return check_launch("conv_tc_kernel");
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
    const int k = d->kh, taps = d->kh * d->kw;
    const bool out_bf16 = d->out_dtype == SR_BF16;

    TcParams p;
    memset(&p, 0, sizeof(p));
    p.Cout = Cd;
    p.block_n = pick_block_n(Cd);
    p.n_blocks = Cd / p.block_n;
    p.c_blocks = Cs / 64;
    p.act = dgrad ? SR_ACT_NONE : d->act; p.slope = d->slope;
    p.shuffle_r = dgrad ? 0 : d->shuffle_r;
    p.bias = bias; p.residual = residual; p.out = dst;
    p.os = 1; p.Hfull = Hd; p.Wfull = Wd;

    alignas(64) CUtensorMap map_a, map_b;
    rc = make_tiled2d_map(&map_b, w, (uint64_t)taps * Cd, (uint64_t)Cs, (uint32_t)p.block_n);
    if (rc != SR_OK) return rc;

    if (!(dgrad && d->stride == 2)) {
        // forward (any stride) or stride-1 dgrad (= forward over dy with mirrored taps, pad' = k-1-pad)
        const int pad = dgrad ? (k - 1 - d->pad) : d->pad;
        const int stride = dgrad ? 1 : d->stride;
        p.M_total = d->N * Hd * Wd; p.Ho = Hd; p.Wo = Wd;
        p.stride = stride; p.pad_w = pad; p.pad_h = pad;
        p.ntaps = taps;
        for (int t = 0; t < taps; ++t) {
            const int ky = t / k, kx = t - ky * k;
            p.tap_w[t] = (unsigned char)t;
            p.tap_ow[t] = (unsigned char)(dgrad ? (k - 1 - kx) : kx);
            p.tap_oh[t] = (unsigned char)(dgrad ? (k - 1 - ky) : ky);
        }
        const int lower[2] = {-pad, -pad};
        const int upper[2] = {pad - (k - 1), pad - (k - 1)};
        rc = make_im2col_map(&map_a, src, d->N, Hs, Ws, Cs, lower, upper, stride, 128);
        if (rc != SR_OK) return rc;
        return launch_tc(p, map_a, map_b, out_bf16, st);
    }

    // stride-2 dgrad of a 3x3 / pad-1 conv: the input pixels split into 4 parity classes (iy%2, ix%2); each
    // class is a stride-1 convolution of dy with the subset of taps whose (iy + pad - ky) is even, written
    // to every second pixel of dx (the transposed-convolution analogue of the PixelShuffle epilogue).
    // The four class launches are independent (disjoint output pixels) and, on the small maps of the deeper layers, far
    // too small to fill the GPU one at a time (D.8: 7 tiles per class) — classes 1..3 are forked onto the caller's auxiliary
    // streams (event fork / join: stream-ordered for the caller, capturable) and run next to class 0.
    const AuxStreams& aux = aux_streams();                 // handed in by the caller (sr_set_aux_streams); none: one stream
    const bool fork_on = aux.n >= 3 && option("SR_S2_STREAMS", 1);
    cudaStream_t const* cls_stream = aux.stream;
    cudaEvent_t ev_fork = aux.fork;
    cudaEvent_t const* ev_join = aux.join;
    if (fork_on) cudaEventRecord(ev_fork, st);
    int cls = 0;
    bool forked[3] = {false, false, false};
    for (int py = 0; py < 2; ++py) {
        for (int px = 0; px < 2; ++px, ++cls) {
            const int Hp = (d->H - py + 1) / 2, Wp = (d->W - px + 1) / 2;     // pixels of this class
            if (Hp <= 0 || Wp <= 0) continue;
            cudaStream_t cst = st;
            if (fork_on && cls > 0) { cst = cls_stream[cls - 1]; cudaStreamWaitEvent(cst, ev_fork, 0); forked[cls - 1] = true; }
            TcParams q = p;
            q.M_total = d->N * Hp * Wp; q.Ho = Hp; q.Wo = Wp;
            q.stride = 1; q.pad_w = 0; q.pad_h = 0;
            q.os = 2; q.py = py; q.px = px; q.Hfull = d->H; q.Wfull = d->W;
            int nt = 0;
            for (int ky = 0; ky < 3; ++ky) {
                const int ty = py + d->pad - ky;
                if (ty < 0 || (ty & 1)) continue;
                for (int kx = 0; kx < 3; ++kx) {
                    const int tx = px + d->pad - kx;
                    if (tx < 0 || (tx & 1)) continue;
                    q.tap_w[nt] = (unsigned char)(ky * 3 + kx);
                    q.tap_ow[nt] = (unsigned char)(tx / 2);
                    q.tap_oh[nt] = (unsigned char)(ty / 2);
                    ++nt;
                }
            }
            q.ntaps = nt;
            const int lower[2] = {0, 0};
            const int upper[2] = {Wp - d->Wo, Hp - d->Ho};     // exactly Hp x Wp base positions over the Ho x Wo map
            rc = make_im2col_map(&map_a, src, d->N, Hs, Ws, Cs, lower, upper, 1, 128);
            if (rc != SR_OK) return rc;
            rc = launch_tc(q, map_a, map_b, out_bf16, cst);
            if (fork_on && cls > 0) cudaEventRecord(ev_join[cls - 1], cst);
            if (rc != SR_OK) break;
        }
        if (rc != SR_OK) break;
    }
    for (int i = 0; i < 3; ++i)
        if (forked[i]) cudaStreamWaitEvent(st, ev_join[i], 0);      // always re-join (also on the error path: a capture must not dangle)
    return rc;
}

}  // namespace sr
