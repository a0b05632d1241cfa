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
    int block_n, n_blocks, num_tiles;
    int stride, pad_w, pad_h;   // im2col base coordinate of output pixel (oy,ox) = (oy*stride - pad_h, ox*stride - pad_w)
    int c_blocks;      // Cin / 64
    int ntaps;         // filter taps visited (a subset for the parity classes of a strided dgrad)
    unsigned char tap_w[49], tap_ow[49], tap_oh[49];   // weight-tap index / im2col offsets per visited tap
    int os, py, px, Hfull, Wfull;   // os > 1: row (n,oy,ox) is written to pixel (oy*os+py, ox*os+px) of an Hfull x Wfull map
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
    const int k_blocks = p.ntaps * p.c_blocks;

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
                const int w0 = ox * p.stride - p.pad_w, h0 = oy * p.stride - p.pad_h;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    const int ti = kb / p.c_blocks, cb = kb - ti * p.c_blocks;
                    const int tap = p.tap_w[ti], offw = p.tap_ow[ti], offh = p.tap_oh[ti];
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
            if ((r > 1 || p.os > 1) && valid) {
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
                } else if (p.os > 1) {
                    idx = (((long long)n * p.Hfull + (oy * p.os + p.py)) * p.Wfull + (ox * p.os + p.px)) * p.Cout + col;
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
    if (d->stride < 1 || d->stride > 2) return false;
    if (dgrad && d->stride == 2 && !(d->kh == 3 && d->pad == 1)) return false;   // parity decomposition: 3x3/pad 1 only
    const int r = d->shuffle_r > 1 ? d->shuffle_r : 1;
    if (!dgrad && r > 1 && (Cd % (r * r) != 0 || (Cd / (r * r)) % 32 != 0)) return false;
    if ((long long)d->N * d->Ho * d->Wo >= (1ll << 31) || (long long)d->N * d->H * d->W >= (1ll << 31)) return false;
    return true;
}

static int launch_tc(const TcParams& p, const CUtensorMap& map_a, const CUtensorMap& map_b, bool out_bf16, cudaStream_t st) {
    const int stage_bytes = TC_A_BYTES + p.block_n * 128;
    const size_t smem = 1024 + (size_t)p.num_stages * stage_bytes + 256;
    const int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
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
    const int stage_bytes = TC_A_BYTES + p.block_n * 128;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    p.num_stages = stages;

    alignas(64) CUtensorMap map_a, map_b;
    rc = make_tiled2d_map(&map_b, w, (uint64_t)taps * Cd, (uint64_t)Cs, (uint32_t)p.block_n);
    if (rc != SR_OK) return rc;

    if (!(dgrad && d->stride == 2)) {
        // forward (any stride) or stride-1 dgrad (= forward over dy with mirrored taps, pad' = k-1-pad)
        const int pad = dgrad ? (k - 1 - d->pad) : d->pad;
        const int stride = dgrad ? 1 : d->stride;
        p.M_total = d->N * Hd * Wd; p.Ho = Hd; p.Wo = Wd;
        p.num_tiles = (int)cdiv(p.M_total, 128) * p.n_blocks;
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
    for (int py = 0; py < 2; ++py) {
        for (int px = 0; px < 2; ++px) {
            const int Hp = (d->H - py + 1) / 2, Wp = (d->W - px + 1) / 2;     // pixels of this class
            if (Hp <= 0 || Wp <= 0) continue;
            TcParams q = p;
            q.M_total = d->N * Hp * Wp; q.Ho = Hp; q.Wo = Wp;
            q.num_tiles = (int)cdiv(q.M_total, 128) * q.n_blocks;
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
            rc = launch_tc(q, map_a, map_b, out_bf16, st);
            if (rc != SR_OK) return rc;
        }
    }
    return SR_OK;
}

}  // namespace sr
