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
    // halo variant (conv_wgrad_halo_kernel): a pixel tile is R full image rows in a padded-linear space of pitch P
    int P, R, row_groups;         // P = roundup8(W + 2), row_groups = ceil(H / R); bk = R * P
    float* dbias;                 // nullable: bias gradient (column sums of dY), accumulated with fp32 atomics
};

constexpr int WG_THREADS = 192;
constexpr int WG_MAX_STAGES = 8;

// Epilogue warps: TMEM -> split-K partial tile (plain 16 B stores) or fp32 atomics into dW (OIHW).  M=128: row = lane id;
// M=64: rows sit in the lower 16 lanes of each 32-lane sub-partition (row = 16*(lane/32) + lane%32).
__device__ __forceinline__ void wg_epilogue(const WgParams& p, uint32_t tmem_base, uint64_t* acc_full, uint64_t* acc_empty,
                                            int warp, int lane, int items) {
    const int taps = p.kh * p.kw;
    const int quarter = warp & 3;
    int row;
    bool row_ok;
    if (p.mt == 128) { row = quarter * 32 + lane; row_ok = true; }
    else { row = quarter * 16 + lane; row_ok = lane < 16; }
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x) {
        int cob = w % p.co_blocks; int q = w / p.co_blocks;
        const int cib = q % p.ci_blocks; q /= p.ci_blocks;
        const int tap = (q % p.tap_groups) * p.tg;
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
        wg_epilogue(p, tmem_base, acc_full, acc_empty, warp, lane, items);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------
// Halo variant for 3x3 / stride 1 / pad 1 layers: the three taps of one filter ROW are three shifted views of ONE
// activation tile (as conv_halo.cu does for the forward operand), so a pixel tile costs one X box + the dY box(es)
// instead of three im2col boxes + dY (K1: 640 -> 384 bytes per pixel and CTA).
//   * reduction index k runs over a padded-linear pixel space of pitch P = roundup8(W + 2): tile = R image rows,
//     k = r * P + w.  dY arrives through a TILED 4-d box {64 ch, P cols from w = 0, R rows}: columns w >= W are out
//     of range and therefore ZERO, which kills every product whose shifted X view wrapped into the next row;
//   * X arrives through ONE tiled 4-d box {64 ch, P cols from w = -1, R rows from h0 + ky - 1} (zero-filled halo);
//     B operand of tap kx = the same tile seen through a descriptor whose start is shifted by kx pixel rows
//     (128 B each): the three taps are stacked along N by giving the MN-major descriptor a leading-dimension
//     byte offset of 128 B (panel j = view j).  Stage layout [X panel][dY panels]: the 2 pixel rows a shifted
//     view reads past the X panel land in the dY panel of the same stage (finite data x zero dY columns);
//   * the bias gradient (column sums of dY) is accumulated by four otherwise idle warps straight from the dY
//     stages of the centre-row items (no separate colsum launch).
// Work items, split-K partial tiles and the reduce kernel are those of conv_tc_wgrad_kernel with nt = 64, tg = 3.
// ------------------------------------------------------------------------------------------------
constexpr int WH_THREADS = 320;      // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue, warps 6-9 bias column sums

__global__ void __launch_bounds__(WH_THREADS, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int panel = p.bk * 128;                           // one [bk pixels][64 ch] bf16 panel
    const int a_panels = p.mt / 64;
    const int stage_bytes = (1 + a_panels) * panel;         // [X panel][dY panels]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.num_stages * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + WG_MAX_STAGES;
    uint64_t* acc_full = bars + 2 * WG_MAX_STAGES;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int items = p.splits * p.tap_groups * p.ci_blocks * p.co_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_dy);
        tma_prefetch_desc(&map_x);
        for (int s = 0; s < p.num_stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 5); }   // MMA commit + 4 bias warps
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 128);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work item -> (split, filter row ky, ci block, co block); co fastest so that neighbouring CTAs share X tiles in L2
    auto decode = [&](int w, int& split, int& ky, int& cib, int& cob) {
        cob = w % p.co_blocks; w /= p.co_blocks;
        cib = w % p.ci_blocks; w /= p.ci_blocks;
        ky = w % p.tap_groups; split = w / p.tap_groups;
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x) {
                int split, ky, cib, cob;
                decode(w, split, ky, cib, cob);
                const int pt0 = split * p.ptiles_per_split;
                const int pt1 = min(p.ptiles, pt0 + p.ptiles_per_split);
                for (int pt = pt0; pt < pt1; ++pt) {
                    const int n = pt / p.row_groups, h0 = (pt - n * p.row_groups) * p.R;
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t* sx = smem + (size_t)stage * stage_bytes;
                    uint8_t* sa = sx + panel;
                    mbar_expect_tx(full + stage, (uint32_t)stage_bytes);
                    tma_load_4d(sx, &map_x, full + stage, cib * 64, -1, h0 + ky - 1, n);
                    for (int j = 0; j < a_panels; ++j)
                        tma_load_4d(sa + j * panel, &map_dy, full + stage, cob * p.mt + j * 64, 0, h0, n);
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: D[M = co][N = 3 taps x 64 ci] += dY^T X, both operands MN-major
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(192 >> 3) << 17) | ((uint32_t)(p.mt >> 4) << 24);
        const int ksteps = p.bk >> 4;
        int stage = 0; uint32_t phase = 0; uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x) {
            int split, ky, cib, cob;
            decode(w, split, ky, cib, cob);
            const int pt0 = split * p.ptiles_per_split;
            const int pt1 = min(p.ptiles, pt0 + p.ptiles_per_split);
            mbar_wait(acc_empty, acc_phase ^ 1);
            tc_fence_after();
            for (int pt = pt0; pt < pt1; ++pt) {
                mbar_wait(full + stage, phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t x_addr = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t bdesc = make_mnmajor_sw128_desc(x_addr, 128u);                       // panel j = view shifted by j pixels
                    const uint64_t adesc = make_mnmajor_sw128_desc(x_addr + (uint32_t)panel, (uint32_t)panel);
                    int k = 0;
                    for (; k + 4 <= ksteps; k += 4)     // 16 pixel rows = 2048 B = 128 descriptor units per K step
                        umma_f16_x4_mn(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (pt > pt0 || k > 0) ? 1u : 0u);
                    for (; k < ksteps; ++k)
                        umma_f16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (pt > pt0 || k > 0) ? 1u : 0u);
                    umma_commit(empty + stage);
                }
                __syncwarp();
                if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
            }
            if (lane == 0) umma_commit(acc_full);
            __syncwarp();
            acc_phase ^= 1;
        }
    } else if (warp < 6) {
        wg_epilogue(p, tmem_base, acc_full, acc_empty, warp, lane, items);
    } else {
        // bias-gradient warps: walk the stage schedule; on centre-row items of ci block 0 add up the dY tile from shared
        // memory.  Thread = (16-byte channel chunk cc, row phase j): rows j, j + tpc, ... of the tile; chunk cc of pixel
        // row r sits at chunk position cc ^ (r & 7) of its 128-byte line (128B swizzle), and r & 7 == j & 7 is constant.
        const int t = threadIdx.x - 192;
        const int tpc = 128 / (a_panels * 8);               // threads per channel chunk: 8 (mt = 128) or 16 (mt = 64)
        const int cc = t / tpc, j = t - cc * tpc;
        const uint32_t chunk_off = (uint32_t)((cc >> 3) * panel + (((cc & 7) ^ (j & 7)) << 4));
        int stage = 0; uint32_t phase = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x) {
            int split, ky, cib, cob;
            decode(w, split, ky, cib, cob);
            const bool do_bias = p.dbias != nullptr && ky == 1 && cib == 0;
            const int pt0 = split * p.ptiles_per_split;
            const int pt1 = min(p.ptiles, pt0 + p.ptiles_per_split);
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            for (int pt = pt0; pt < pt1; ++pt) {
                mbar_wait(full + stage, phase);
                if (do_bias) {
                    const uint8_t* sa = smem + (size_t)stage * stage_bytes + panel + chunk_off;
                    for (int r = j; r < p.bk; r += tpc) {
                        const uint4 v = *reinterpret_cast<const uint4*>(sa + (size_t)r * 128);
                        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                        for (int i = 0; i < 4; ++i) { acc[2 * i] += __low2float(h[i]); acc[2 * i + 1] += __high2float(h[i]); }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + stage);
                if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
            }
            if (do_bias) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float v = acc[i];
                    for (int o = tpc >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    acc[i] = v;
                }
                const int co = cob * p.mt + cc * 8;
                if (j == 0 && co < p.Cout) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) atomicAdd(p.dbias + co + i, acc[i]);
                }
            }
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

// Split-K reduction of the halo variant (3x3: tg = 3, nt = 64): one block per (output channel, 64-wide ci block) gathers the
// row's 3 x 192 partial columns (one thread per column, the splits summed in registers), transposes them through shared
// memory into OIHW order ([ci][ky][kx]) and adds the 576 contiguous floats of dW with coalesced accesses — the one-thread-
// per-element kernel above scatters its read-modify-writes 36 bytes apart.
__global__ void __launch_bounds__(576)
wgrad_reduce_rows_kernel(const float* __restrict__ partial, WgParams p, int tiles) {
    __shared__ float row_s[576];
    const int co = blockIdx.x / p.ci_blocks, cib = blockIdx.x - co * p.ci_blocks;
    const int cob = co / p.mt, row = co - cob * p.mt;
    const int e = threadIdx.x;
    const int g = e / 192, col = e - g * 192;
    const int t_local = col >> 6, ci = col & 63;
    const int tile = (g * p.ci_blocks + cib) * p.co_blocks + cob;
    const long long stride = (long long)tiles * p.mt * 192;
    const float* src = partial + ((long long)tile * p.mt + row) * 192 + col;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int sp = 0;
    for (; sp + 8 <= p.splits; sp += 8) {
        const float v0 = src[(long long)sp * stride], v1 = src[(long long)(sp + 1) * stride], v2 = src[(long long)(sp + 2) * stride],
                    v3 = src[(long long)(sp + 3) * stride], v4 = src[(long long)(sp + 4) * stride], v5 = src[(long long)(sp + 5) * stride],
                    v6 = src[(long long)(sp + 6) * stride], v7 = src[(long long)(sp + 7) * stride];
        a0 += v0 + v4; a1 += v1 + v5; a2 += v2 + v6; a3 += v3 + v7;
    }
    for (; sp < p.splits; ++sp) a0 += src[(long long)sp * stride];
    row_s[ci * 9 + g * 3 + t_local] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    p.dw[((long long)co * p.Cin + cib * 64) * 9 + e] += row_s[e];
}

static int g_wg_sms = 0;


bool conv_tc_wgrad_supported(const sr_conv_desc* d) {
    if (d->in_dtype != SR_BF16) return false;
    if (d->Cin % 64 != 0 || d->Cout % 64 != 0) return false;
    if (d->kh != d->kw || d->kh > 7) return false;
    if (d->stride < 1 || d->stride > 2) return false;
    if ((long long)d->N * d->Ho * d->Wo >= (1ll << 31)) return false;
    return true;
}

// dw must have been zero-filled (or hold the value to accumulate onto).
// upper bound of the split-K scratch of either weight-gradient kernel: one [mt <= 128] x [tg * nt <= 256] fp32 tile per work item,
// at most one item per SM
size_t conv_tc_wgrad_workspace_bytes(const sr_conv_desc*) { return (size_t)160 * 128 * 256 * sizeof(float); }

int conv_tc_wgrad_run(const sr_conv_desc* d, const void* x, const void* dy, float* dw, float* ws, size_t ws_bytes, cudaStream_t st) {
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
    const int tg_override = option("SR_WG_TG", 0), bk_override = option("SR_WG_BK", 0);      // tuning options (sr_set_option)
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
    p.P = p.R = p.row_groups = 0; p.dbias = nullptr;

    alignas(64) CUtensorMap map_dy, map_x;
    rc = make_tiled2d_map(&map_dy, dy, (uint64_t)p.M_total, (uint64_t)d->Cout, (uint32_t)p.bk);
    if (rc != SR_OK) return rc;
    const int lower[2] = {-d->pad, -d->pad};
    const int upper[2] = {d->pad - (d->kw - 1), d->pad - (d->kh - 1)};
    rc = make_im2col_map(&map_x, x, d->N, d->H, d->W, d->Cin, lower, upper, d->stride, p.bk);
    if (rc != SR_OK) return rc;

    const int items = p.splits * tiles;
    const int grid = items < g_wg_sms ? items : g_wg_sms;
    // split-K partials go to the caller's workspace (plain 16-byte stores + one reduce kernel) when it is large
    // enough; otherwise they are combined with fp32 atomics straight into dW
    const size_t need = (size_t)items * p.mt * (p.tg * p.nt) * sizeof(float);
    p.partial = (!option("SR_WG_NOWS", 0) && ws && need <= ws_bytes && p.splits > 1) ? ws : nullptr;
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

// ---- halo variant: host side ----
bool conv_wgrad_halo_supported(const sr_conv_desc* d) {
    if (!option("SR_WG_HALO", 1)) return false;
    if (d->in_dtype != SR_BF16) return false;
    if (d->kh != 3 || d->kw != 3 || d->stride != 1 || d->pad != 1) return false;
    if (d->Cin % 64 != 0 || d->Cout % 64 != 0) return false;
    if (d->W + 2 > 256 || d->Ho != d->H || d->Wo != d->W) return false;
    if ((long long)d->N * d->H * d->W >= (1ll << 31)) return false;
    return true;
}

// dw (and dbias, when given) must have been zero-filled or hold the value to accumulate onto.
int conv_wgrad_halo_run(const sr_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, float* ws, size_t ws_bytes, cudaStream_t st) {
    int rc = load_driver_fns();
    if (rc != SR_OK) return rc;
    if (!g_wg_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_wg_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    WgParams p;
    p.M_total = d->N * d->Ho * d->Wo; p.Ho = d->Ho; p.Wo = d->Wo;
    p.Cout = d->Cout; p.Cin = d->Cin; p.kh = 3; p.kw = 3; p.stride = 1; p.pad = 1;
    p.mt = (d->Cout % 128 == 0) ? 128 : 64;
    p.nt = 64; p.tg = 3; p.tap_groups = 3;
    p.co_blocks = d->Cout / p.mt;
    p.ci_blocks = d->Cin / 64;
    p.P = (int)cdiv(d->W + 2, 8) * 8;
    // rows per tile: R * P must be a multiple of 16 pixels (one K step); aim at ~128-pixel stages (few, large TMA boxes)
    const int r_mult = option("SR_WG_RMULT", 2);            // measured: 2 (224-pixel stages at 54^2) beats 1 on every layer but D.7
    const int r0 = (p.P % 16 == 0) ? 1 : 2;
    int m = 128 / (r0 * p.P); if (m < 1) m = 1;
    p.R = r0 * m * (r_mult > 0 ? r_mult : 1);
    const int a_panels = p.mt / 64;
    while (p.R > r0 && (size_t)p.R * p.P * 128 * (1 + a_panels) * 2 > 200 * 1024) p.R -= r0;
    if (p.R > 256) p.R = 256 - 256 % r0;
    p.bk = p.R * p.P;
    p.row_groups = (int)cdiv(d->H, p.R);
    p.ptiles = d->N * p.row_groups;
    const int stage_bytes = (1 + a_panels) * p.bk * 128;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
    if (stages < 2) { set_error("conv_wgrad_halo: stage of %d bytes does not fit twice (W=%d Cout=%d)", stage_bytes, d->W, d->Cout); return SR_ERR_UNSUPPORTED; }
    p.num_stages = stages;
    const int tiles = p.tap_groups * p.co_blocks * p.ci_blocks;
    int splits = g_wg_sms / tiles;                                  // one wave of CTAs, no ragged second round
    const int max_splits = (int)cdiv(p.ptiles, 4);                  // at least 4 pixel tiles per split
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.ptiles_per_split = (int)cdiv(p.ptiles, splits);
    p.splits = (int)cdiv(p.ptiles, p.ptiles_per_split);
    p.dw = dw;
    p.dbias = dbias;

    alignas(64) CUtensorMap map_dy, map_x;
    rc = make_tiled4d_map(&map_dy, dy, d->N, d->Ho, d->Wo, d->Cout, p.P, p.R);
    if (rc != SR_OK) return rc;
    rc = make_tiled4d_map(&map_x, x, d->N, d->H, d->W, d->Cin, p.P, p.R);
    if (rc != SR_OK) return rc;

    const int items = p.splits * tiles;
    const int grid = items < g_wg_sms ? items : g_wg_sms;
    const size_t need = (size_t)items * p.mt * (p.tg * p.nt) * sizeof(float);
    p.partial = (!option("SR_WG_NOWS", 0) && ws && need <= ws_bytes && p.splits > 1) ? ws : nullptr;
    const size_t smem = 1024 + (size_t)stages * stage_bytes + 256;
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_set = true; }
    conv_wgrad_halo_kernel<<<grid, WH_THREADS, smem, st>>>(map_dy, map_x, p);
    count_launch();
    if (p.partial) {
        wgrad_reduce_rows_kernel<<<(unsigned)(d->Cout * p.ci_blocks), 576, 0, st>>>(p.partial, p, tiles);
        count_launch();
    }
    return check_launch("conv_wgrad_halo_kernel");
}

}  // namespace sr
