// tcgen05 weight-gradient kernel for sm_100a (B200).
//
//   dW[tap][co][ci] = sum over output pixels  dY[pix][co] * X[pix @ tap][ci]
//
// GEMM view per filter tap: D[M = co][N = ci] += A[M][K = pixels] * B[N][K]^T, where BOTH operands are
// "MN-major" — the reduction index (pixel) is the slow index of the NHWC tensors, the channel index the
// fast one — so the tiles are consumed exactly as TMA lands them, with no transposition pass:
//   * A: dY viewed as a [pixels][Cout] matrix, tiled TMA boxes of [BK pixels][64 channels];
//   * B: X through the SAME im2col tensor map as the forward kernel (BK output pixels x 64 input
//        channels at one tap offset, zero-filled halo) — the forward A tile is the wgrad B tile.
// Accumulators live in TMEM (fp32, M x N <= 128 x 256); the pixel dimension is split across CTAs
// (split-K) so that ~one wave of CTAs covers the problem, and partial sums are combined with fp32
// atomics straight into the OIHW master-gradient buffer (which is also the all-reduce bucket).
// Reference: autograd of every nn.Conv2d of model/sradsgan.py (K18 of SURVEY.md §2b).
#include <stdlib.h>

#include "tc_common.cuh"

namespace sr {

struct WgParams {
    int M_total, Ho, Wo;          // output pixel space (the reduction dimension)
    int Cout, Cin, kh, kw, stride, pad;
    int mt, nt, bk;               // tile: mt output channels x nt input channels, bk pixels per stage
    int tg, tap_groups;           // filter taps stacked along the MMA N dimension (N = tg * nt <= 256), groups of taps
    int co_blocks, ci_blocks, splits, ptiles, ptiles_per_split;
    int num_stages;
    float* dw;
    float* partial;               // non-NULL: split-K partial tiles [item][mt][tg*nt] (plain stores; reduced by a second kernel)
};

constexpr int WG_THREADS = 192;
constexpr int WG_MAX_STAGES = 8;

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int panel = p.bk * 128;                           // one [bk pixels][64 ch] bf16 panel
    const int a_panels = p.mt / 64, ci_panels = p.nt / 64;
    const int b_panels = ci_panels * p.tg;                  // stage layout: [dY panels][tap 0 ci panels][tap 1 ci panels]...
    const int stage_bytes = (a_panels + b_panels) * panel;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.num_stages * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + WG_MAX_STAGES;
    uint64_t* acc_full = bars + 2 * WG_MAX_STAGES;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int taps = p.kh * p.kw;
    const int items = p.splits * p.tap_groups * p.ci_blocks * p.co_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_dy);
        tma_prefetch_desc(&map_x);
        for (int s = 0; s < p.num_stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 128);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work item -> (split, first tap of the group, ci block, co block); co fastest so that neighbouring CTAs share X
    // tiles in L2.  A group stacks up to tg taps along N: one dY tile feeds all of them.
    auto decode = [&](int w, int& split, int& tap, int& cib, int& cob) {
        cob = w % p.co_blocks; w /= p.co_blocks;
        cib = w % p.ci_blocks; w /= p.ci_blocks;
        tap = (w % p.tap_groups) * p.tg; split = w / p.tap_groups;
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x) {
                int split, tap, cib, cob;
                decode(w, split, tap, cib, cob);
                const int ntap = min(p.tg, taps - tap);
                const int pt0 = split * p.ptiles_per_split;
                const int pt1 = min(p.ptiles, pt0 + p.ptiles_per_split);
                for (int pt = pt0; pt < pt1; ++pt) {
                    const int m0 = pt * p.bk;
                    const int ox = m0 % p.Wo; const int q = m0 / p.Wo;
                    const int oy = q % p.Ho; const int n = q / p.Ho;
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    uint8_t* sb = sa + a_panels * panel;
                    mbar_expect_tx(full + stage, (uint32_t)((a_panels + ntap * ci_panels) * panel));
                    for (int j = 0; j < a_panels; ++j)
                        tma_load_2d(sa + j * panel, &map_dy, full + stage, cob * p.mt + j * 64, m0);
                    for (int t = 0; t < ntap; ++t) {
                        const int ky = (tap + t) / p.kw, kx = (tap + t) - ky * p.kw;
                        for (int j = 0; j < ci_panels; ++j)
                            tma_load_im2col(sb + (t * ci_panels + j) * panel, &map_x, full + stage, cib * p.nt + j * 64,
                                            ox * p.stride - p.pad, oy * p.stride - p.pad, n, (uint16_t)kx, (uint16_t)ky);
                    }
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: the warp walks the schedule converged, lane 0 issues 4 K-steps per asm statement
        int stage = 0; uint32_t phase = 0; uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x) {
            int split, tap, cib, cob;
            decode(w, split, tap, cib, cob);
            const int ntap = min(p.tg, taps - tap);
            // D=f32, A=B=bf16, A and B MN-major (bits 15/16), N = ntap * nt, M = mt
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)((ntap * p.nt) >> 3) << 17) | ((uint32_t)(p.mt >> 4) << 24);
            const int pt0 = split * p.ptiles_per_split;
            const int pt1 = min(p.ptiles, pt0 + p.ptiles_per_split);
            mbar_wait(acc_empty, acc_phase ^ 1);
            tc_fence_after();
            for (int pt = pt0; pt < pt1; ++pt) {
                mbar_wait(full + stage, phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint32_t b_addr = a_addr + a_panels * panel;
                    const uint64_t adesc = make_mnmajor_sw128_desc(a_addr, (uint32_t)panel);
                    const uint64_t bdesc = make_mnmajor_sw128_desc(b_addr, (uint32_t)panel);
                    for (int k = 0; k < p.bk / 64; ++k)     // 64 pixels (4 MMAs of K = 16) per statement
                        umma_f16_x4_mn(tmem_base, adesc + (uint64_t)(k * 512), bdesc + (uint64_t)(k * 512), idesc, (pt > pt0 || k > 0) ? 1u : 0u);
                    umma_commit(empty + stage);
                }
                __syncwarp();
                if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
            }
            if (lane == 0) umma_commit(acc_full);
            __syncwarp();
            acc_phase ^= 1;
        }
    } else {
        // epilogue: TMEM -> fp32 atomics into dW (OIHW). M=128: row = lane id; M=64: rows sit in the lower
        // 16 lanes of each 32-lane sub-partition (row = 16*(lane/32) + lane%32).
        const int quarter = warp & 3;
        int row;
        bool row_ok;
        if (p.mt == 128) { row = quarter * 32 + lane; row_ok = true; }
        else { row = quarter * 16 + lane; row_ok = lane < 16; }
        uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x) {
            int split, tap, cib, cob;
            decode(w, split, tap, cib, cob);
            const int ntap = min(p.tg, taps - tap);
            mbar_wait(acc_full, acc_phase);
            tc_fence_after();
            const int co = cob * p.mt + row;
            const uint32_t t_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
            const int ncols = p.tg * p.nt;
            for (int c0 = 0; c0 < ntap * p.nt; c0 += 32) {      // column = t_local * nt + ci_local
                uint32_t v[32];
                tmem_ld32(t_base + (uint32_t)c0, v);
                if (row_ok && co < p.Cout) {
                    if (p.partial) {
                        float4* dst = reinterpret_cast<float4*>(p.partial + ((long long)w * p.mt + row) * ncols + c0);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                                 __uint_as_float(v[4 * j + 3]));
                    } else {
                        const int t_local = c0 / p.nt, ci0 = c0 - t_local * p.nt;
                        float* dst = p.dw + ((long long)co * p.Cin + (cib * p.nt + ci0)) * taps + tap + t_local;
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(dst + (long long)j * taps, __uint_as_float(v[j]));
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty);
            acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// dW[co][ci][tap] += sum over splits of the partial tiles written by conv_tc_wgrad_kernel (one thread per element)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, WgParams p, int tiles, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ncols = p.tg * p.nt;
    const int col = (int)(i % ncols); long long r = i / ncols;
    const int row = (int)(r % p.mt); const int tile = (int)(r / p.mt);      // tile = (group * ci_blocks + cib) * co_blocks + cob
    const int cob = tile % p.co_blocks; int q = tile / p.co_blocks;
    const int cib = q % p.ci_blocks; const int g = q / p.ci_blocks;
    const int taps = p.kh * p.kw;
    const int t_local = col / p.nt, ci = col - t_local * p.nt;
    const int tap = g * p.tg + t_local;
    const int co = cob * p.mt + row;
    if (tap >= taps || co >= p.Cout) return;
    const long long stride = (long long)tiles * p.mt * ncols;
    const float* src = partial + ((long long)tile * p.mt + row) * ncols + col;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // independent chains: the loads of four splits are in flight together
    int s = 0;
    for (; s + 4 <= p.splits; s += 4) {
        a0 += src[(long long)s * stride]; a1 += src[(long long)(s + 1) * stride];
        a2 += src[(long long)(s + 2) * stride]; a3 += src[(long long)(s + 3) * stride];
    }
    for (; s < p.splits; ++s) a0 += src[(long long)s * stride];
    p.dw[((long long)co * p.Cin + cib * p.nt + ci) * taps + tap] += (a0 + a1) + (a2 + a3);
}

static int g_wg_sms = 0;
static float* g_wg_workspace = nullptr;
static size_t g_wg_workspace_bytes = 0;

void conv_tc_wgrad_set_workspace(void* ptr, size_t bytes) { g_wg_workspace = (float*)ptr; g_wg_workspace_bytes = bytes; }

bool conv_tc_wgrad_supported(const sr_conv_desc* d) {
    if (d->in_dtype != SR_BF16) return false;
    if (d->Cin % 64 != 0 || d->Cout % 64 != 0) return false;
    if (d->kh != d->kw || d->kh > 7) return false;
    if (d->stride < 1 || d->stride > 2) return false;
    if ((long long)d->N * d->Ho * d->Wo >= (1ll << 31)) return false;
    return true;
}

// dw must have been zero-filled (or hold the value to accumulate onto).
int conv_tc_wgrad_run(const sr_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    int rc = load_driver_fns();
    if (rc != SR_OK) return rc;
    if (!g_wg_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_wg_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    WgParams p;
    p.M_total = d->N * d->Ho * d->Wo; p.Ho = d->Ho; p.Wo = d->Wo;
    p.Cout = d->Cout; p.Cin = d->Cin; p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad;
    p.mt = (d->Cout % 128 == 0) ? 128 : 64;
    p.nt = (d->Cin % 256 == 0) ? 256 : (d->Cin % 128 == 0 ? 128 : 64);
    p.co_blocks = d->Cout / p.mt;
    p.ci_blocks = d->Cin / p.nt;
    const int taps_total = d->kh * d->kw;
    p.tg = 256 / p.nt;                                              // taps stacked along N (N = tg * nt <= 256)
    if (p.tg > taps_total) p.tg = taps_total;
    if (taps_total == 9 && p.tg >= 3) p.tg = 3;                     // one filter row per group: 3 equal groups
    static int tg_override = -1, bk_override = -1;                  // tuning knobs (environment, read once)
    if (tg_override < 0) { const char* e = getenv("SR_WG_TG"); tg_override = e ? atoi(e) : 0; e = getenv("SR_WG_BK"); bk_override = e ? atoi(e) : 0; }
    if (tg_override > 0 && tg_override < p.tg) p.tg = tg_override;
    p.tap_groups = (int)cdiv(taps_total, p.tg);
    const int panels = p.mt / 64 + p.tg * (p.nt / 64);
    // 128-pixel stages whenever two of them fit: large TMA boxes and few barrier round trips beat pipeline depth
    // (measured: K2 62 -> 47 us, V.3 136 -> 89 us, profiles/r01_wgrad_knobs.txt)
    p.bk = (panels * 128 * 128 * 2 <= 200 * 1024) ? 128 : 64;
    if (bk_override == 64 || (bk_override == 128 && panels * 128 * 128 * 2 <= 200 * 1024)) p.bk = bk_override;
    const int stage_bytes = panels * p.bk * 128;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
    p.num_stages = stages;
    p.ptiles = (int)cdiv(p.M_total, p.bk);
    const int tiles = p.tap_groups * p.co_blocks * p.ci_blocks;
    int splits = g_wg_sms / tiles;                                  // one wave of CTAs, no ragged second round
    const int max_splits = (int)cdiv(p.ptiles, 4);                  // at least 4 pixel tiles per split
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.ptiles_per_split = (int)cdiv(p.ptiles, splits);
    p.splits = (int)cdiv(p.ptiles, p.ptiles_per_split);
    p.dw = dw;

    alignas(64) CUtensorMap map_dy, map_x;
    rc = make_tiled2d_map(&map_dy, dy, (uint64_t)p.M_total, (uint64_t)d->Cout, (uint32_t)p.bk);
    if (rc != SR_OK) return rc;
    const int lower[2] = {-d->pad, -d->pad};
    const int upper[2] = {d->pad - (d->kw - 1), d->pad - (d->kh - 1)};
    rc = make_im2col_map(&map_x, x, d->N, d->H, d->W, d->Cin, lower, upper, d->stride, p.bk);
    if (rc != SR_OK) return rc;

    const int items = p.splits * tiles;
    const int grid = items < g_wg_sms ? items : g_wg_sms;
    // split-K partials go to the registered workspace (plain 16-byte stores + one reduce kernel) when it is large
    // enough; otherwise they are combined with fp32 atomics straight into dW
    const size_t need = (size_t)items * p.mt * (p.tg * p.nt) * sizeof(float);
    static int ws_off = -1;
    if (ws_off < 0) { const char* e = getenv("SR_WG_NOWS"); ws_off = (e && atoi(e)) ? 1 : 0; }
    p.partial = (!ws_off && g_wg_workspace && need <= g_wg_workspace_bytes && p.splits > 1) ? g_wg_workspace : nullptr;
    const size_t smem = 1024 + (size_t)stages * stage_bytes + 256;
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_set = true; }
    conv_tc_wgrad_kernel<<<grid, WG_THREADS, smem, st>>>(map_dy, map_x, p);
    count_launch();
    if (p.partial) {
        const long long total = (long long)tiles * p.mt * (p.tg * p.nt);
        wgrad_reduce_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(p.partial, p, tiles, total);
        count_launch();
    }
    return check_launch("conv_tc_wgrad_kernel");
}

}  // namespace sr
