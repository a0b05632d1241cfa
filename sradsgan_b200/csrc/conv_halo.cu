// tcgen05 halo-tile convolution for sm_100a (B200): 3x3 / stride 1 / pad 1, forward and dgrad.
//
// The im2col kernel (conv_tc.cu) re-reads every activation once per filter tap (9x).  Here the activation tile is
// loaded ONCE per 64-channel block:
//   * one TILED 4-d TMA box {64 ch, TWp columns, TR+2 rows, 1 image} lands the tile INCLUDING its one-pixel halo
//     (out-of-range coordinates are zero-filled by the hardware = the convolution padding) as consecutive 128-byte
//     pixel rows, 128B-swizzled, i.e. directly as a K-major UMMA operand with pixel = M row;
//   * output positions are numbered in the padded-linear space of the tile, q = r*TWp + cp (cp = 0 / TWp-1 are the
//     halo columns), MMA row j <-> q = j+1, and the A operand of filter tap (ky,kx) is the SAME shared-memory tile
//     seen through a descriptor whose start address is shifted by (ky*TWp + kx) pixel rows — the tensor core applies
//     the 128B swizzle to absolute shared-memory addresses, so any 128 B-aligned start works
//     (profiles/r01_umma_descriptor_probe.txt).  The 2 halo columns per row (and the rows >= TR*TWp) produce junk
//     accumulator rows that are never stored: ~84 % of the MMA rows are useful, for 4.5x less L2->SMEM traffic;
//   * weights: resident in shared memory for the whole kernel when the [9*Cin] x block_n slab fits (Cin = 64),
//     else streamed tap by tap through their own TMA ring.
//   * 64 output channels (RAB conv2, the input gradient of conv1): an N = 64 instruction keeps the tensor core busy for 32
//     cycles but cannot be issued faster than every ~50 (profiles/r01_umma_issue_rate.txt) — the kernel was issue-bound at
//     ~70 cycles per instruction (profiles/r02_halo_trace.txt).  `stack` mode: the weight stage of a filter ROW holds its three
//     kx taps as 192 consecutive B rows, ONE N = 192 instruction per k-step multiplies the row-shifted (ky only) view with
//     all three, and the column shift moves to the epilogue: out[j] = R_0[j] + R_1[j + 1] + R_2[j + 2] (accumulator rows are
//     TMEM lanes = threads of the epilogue warp: two shuffles per value, the two rows past a warp's quarter come from the
//     next warp through shared memory).  3x fewer instructions, each long enough for a single issuer.
// Warp roles, TMEM double buffering and the epilogue are those of conv_tc.cu.
// Reference call sites: the 3x3 stride-1 nn.Conv2d layers of model/sradsgan.py (RAB :222-223, GAB_UP :381,
// Discriminator :476 odd blocks, VGG19 features) and their input gradients.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace sr {

struct HaloParams {
    int N, H, W, Cout;               // Cout = channels actually stored (bias length); narrow: Cout <= 4 computed as a 64-wide block
    int narrow;
    int TWp, TW, TR;                 // padded tile width, valid columns (TWp - 2), output rows per tile
    int tiles_x, tiles_y, p_tiles;   // pixel tiles = N * tiles_y * tiles_x
    int block_n, n_blocks, c_blocks;
    int flip;                        // dgrad: filter taps mirrored
    int quads;                       // bf16 outputs: row-coalesced (quad-transposed) stores
    int act; float slope; int shuffle_r;
    int a_stage_bytes, a_box_bytes, num_a_stages, num_b_stages, resident;
    int dual;                        // 1: two MMA-issuer warps work on two pixel tiles at once (shared weights)
    int n_pair_items, n_items;       // per CTA lane: items [0, n_pair_items) are tile pairs, the rest single tiles
    int split_ok;                    // dual + streamed weights + >= 2 channel blocks: a SINGLE tile is split along K between the two issuers
    int stack;                       // Cout = 64: the three kx taps of a filter row stacked along N (one N = 192 instruction instead of three N = 64
                                     // ones, which are issue-bound); accumulator = [R_kx0 | R_kx1 | R_kx2], out[j] = sum_kx R_kx[j + shift(kx)] in the
                                     // epilogue.  Tile pairs (dual = 1) on ONE issuer, single-buffered accumulators (2 x 192 TMEM columns)
    int mcast;                       // stack mode on single tiles: clusters of TWO CTAs share every streamed weight stage (each loads half of
                                     // it and multicasts it into both; the two weight rings advance in lockstep, the CTA with fewer
                                     // tiles running "ghost" items that only load and release weights)
    const float* bias;
    const void* residual;
    const void* mask;                // nullable bf16 tensor shaped like the output: out *= act'(mask)
    float mask_slope;
    void* out;
    // CLAM pooling partials of the (bf16) output, emitted by the epilogue for the local-attention chain that follows conv2 of a
    // RAB (la_band.cu): [N][pool_rows][64] channel sums and packed (max, first arg-max pixel) keys, one row per pixel tile of
    // the image; nullable
    float* pool_sum; unsigned int* pool_key; int pool_rows;
#ifdef SR_WITH_PROBES
    int dbg;                         // diagnostics build: 1 = epilogue issues no global stores, 2 = epilogue reads no accumulators either
    long long* trace;                // diagnostics build: [grid][HL_TRACE_SLOTS] SM clock stamps of the roles' milestones (scripts/halo_trace.py)
#endif
};

#ifdef SR_WITH_PROBES
constexpr int HL_TRACE_SLOTS = 128;
long long* g_hl_trace = nullptr;
int g_hl_dbg = 0;
// slots: 0 start, 1 setup done, 2 end; producer 8+it*2+w = A load issued; issuer w: 32+w*24+it*3+{0 accumulator free, 1 A landed,
// 2 MMAs committed}; epilogue group g (quarter 0): 80+g*16+it*2+{0 accumulator full, 1 stores issued}
#define HL_TRACE(slot) do { if (p.trace && (slot) < HL_TRACE_SLOTS) p.trace[(long long)blockIdx.x * HL_TRACE_SLOTS + (slot)] = clock64(); } while (0)
#else
#define HL_TRACE(slot) do { } while (0)
#endif

__device__ __forceinline__ unsigned int hl_bf16_key(__nv_bfloat16 v) {
    const unsigned int b = __bfloat16_as_ushort(v);
    return (b & 0x8000u) ? (~b & 0xFFFFu) : (b | 0x8000u);
}

// Epilogue of one 32-row x 32-column accumulator chunk (lane = row) for launches that also emit the CLAM pooling partials:
// bias + activation -> bf16 -> 16-byte stores, then the column maxima (packed with the pixel index) and column sums of the
// ROUNDED values (what the chain will read back) by a butterfly transpose-reduce — 31 shuffles per quantity, after which lane L
// holds column L.  Rows that are not output pixels store nothing and contribute nothing.  The value array is reused in place
// (keys first, then sums) so that the extra live state is one 32-register array.
template <int ACT>
__device__ __forceinline__ void hl_store_pool_chunk(const uint32_t (&v)[32], const float* bias_s, float slope, bool valid, __nv_bfloat16* o,
                                                    unsigned int ptag, int lane, float& col_sum, unsigned int& col_key) {
    float f[32];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const float4 b = lds128(bias_s + g * 4);
        f[g * 4 + 0] = __bfloat162float(__float2bfloat16_rn(tc_act<ACT>(__uint_as_float(v[g * 4 + 0]) + b.x, slope)));
        f[g * 4 + 1] = __bfloat162float(__float2bfloat16_rn(tc_act<ACT>(__uint_as_float(v[g * 4 + 1]) + b.y, slope)));
        f[g * 4 + 2] = __bfloat162float(__float2bfloat16_rn(tc_act<ACT>(__uint_as_float(v[g * 4 + 2]) + b.z, slope)));
        f[g * 4 + 3] = __bfloat162float(__float2bfloat16_rn(tc_act<ACT>(__uint_as_float(v[g * 4 + 3]) + b.w, slope)));
    }
    if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint4 w;                                  // f is bf16-representable: its upper 16 bits ARE the bf16 value
            w.x = (__float_as_uint(f[g * 8 + 0]) >> 16) | (__float_as_uint(f[g * 8 + 1]) & 0xFFFF0000u);
            w.y = (__float_as_uint(f[g * 8 + 2]) >> 16) | (__float_as_uint(f[g * 8 + 3]) & 0xFFFF0000u);
            w.z = (__float_as_uint(f[g * 8 + 4]) >> 16) | (__float_as_uint(f[g * 8 + 5]) & 0xFFFF0000u);
            w.w = (__float_as_uint(f[g * 8 + 6]) >> 16) | (__float_as_uint(f[g * 8 + 7]) & 0xFFFF0000u);
            reinterpret_cast<uint4*>(o)[g] = w;
        }
    }
    {
        unsigned int k[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const unsigned int bits = __float_as_uint(f[i]) >> 16;
            const unsigned int key = (bits & 0x8000u) ? (~bits & 0xFFFFu) : (bits | 0x8000u);       // orderable bf16 key
            k[i] = valid ? ((key << 16) | ptag) : 0u;
        }
#pragma unroll
        for (int half = 16; half >= 1; half >>= 1) {
            const bool up = (lane & half) != 0;
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const unsigned int recv = __shfl_xor_sync(0xffffffffu, up ? k[i] : k[i + half], half);
                const unsigned int keep = up ? k[i + half] : k[i];
                k[i] = keep > recv ? keep : recv;
            }
        }
        col_key = k[0];
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = valid ? f[i] : 0.f;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float recv = __shfl_xor_sync(0xffffffffu, up ? f[i] : f[i + half], half);
            f[i] = (up ? f[i + half] : f[i]) + recv;
        }
    }
    col_sum = f[0];
}

constexpr int HL_EPI_WARPS = 8;
constexpr int HL_THREADS = 96 + 32 * HL_EPI_WARPS;   // warp 0 TMA, warps 1-2 MMA issuers, warps 3..10 epilogue
constexpr int HL_MAX_A = 6, HL_MAX_B = 16, HL_MAX_RES = 36;
constexpr int HL_BIAS_MAX = 1024;
constexpr int HL_TILE_BUDGET = 232448 - 1024 - 1024 - HL_BIAS_MAX * 4;

// Work item `it` of this CTA -> pixel tiles (t0, t1; -1 = none) and column block.  Resident mode pins the column
// block to the CTA (lane = CTAs sharing a column block); dual mode hands out PAIRS of pixel tiles first and the
// ragged remainder as single tiles, so that the last round costs one tile time, not two.
__device__ __forceinline__ bool hl_item_at(const HaloParams& p, int it, int& t0, int& t1, int& nb) {
    int rank, lane_ctas;
    if (p.resident) { nb = blockIdx.x % p.n_blocks; rank = blockIdx.x / p.n_blocks; lane_ctas = gridDim.x / p.n_blocks; }
    else { nb = 0; rank = blockIdx.x; lane_ctas = gridDim.x; }
    const int idx = rank + it * lane_ctas;
    if (idx >= p.n_items) return false;
    if (p.dual) {
        if (idx < p.n_pair_items) { t0 = 2 * idx; t1 = 2 * idx + 1; if (t1 >= p.p_tiles) t1 = -1; }
        else { t0 = 2 * p.n_pair_items + (idx - p.n_pair_items); t1 = -1; }
    } else if (p.resident) {
        t0 = idx; t1 = -1;
    } else {
        t0 = idx / p.n_blocks; nb = idx - t0 * p.n_blocks; t1 = -1;
    }
    return true;
}

__device__ __forceinline__ void hl_tile_origin(const HaloParams& p, int p_tile, int& n, int& y0, int& x0) {
    const int tx = p_tile % p.tiles_x; const int q = p_tile / p.tiles_x;
    const int ty = q % p.tiles_y; n = q / p.tiles_y;
    y0 = ty * p.TR; x0 = tx * p.TW;
}

// Epilogue warp.  dual: serves issuer `grp` (its own double-buffered accumulators), all 32-column chunks.
// single: both groups serve the one issuer, chunk parity = grp.
// split (dual, a single tile whose channel blocks were divided between the two issuers): both groups wait for BOTH
// accumulators, group g sums and stores the chunks of parity g, and after a barrier among the eight epilogue warps group g
// releases issuer g's buffer.  Every warp tracks both issuers' buffer indices so the three item kinds can interleave.
template <typename OutT, int ACT, bool POOL, bool STACK>
__device__ __forceinline__ void hl_epilogue(const HaloParams& p, uint32_t tmem_base, uint64_t* acc_full, uint64_t* acc_empty,
                                            const float* bias_s, float* xch_s, int quarter, int grp, int lane) {
    const int j = quarter * 32 + lane;           // MMA row
    const int q = j + 1;                         // padded-linear position inside the tile
    const int tr = q / p.TWp, cp = q - tr * p.TWp;
    const int r = p.shuffle_r > 1 ? p.shuffle_r : 1;
    const int cq = p.Cout / (r * r);
    const int chunks = p.block_n >> 5;
    const int acc_stride = p.dual ? 128 : 256;
    OutT* out = reinterpret_cast<OutT*>(p.out);
    const OutT* res = reinterpret_cast<const OutT*>(p.residual);
    const __nv_bfloat16* mask = reinterpret_cast<const __nv_bfloat16*>(p.mask);
    int accs[2] = {0, 0}; uint32_t phs[2] = {0, 0};
    const int nbuf = (p.stack && p.dual) ? 1 : 2;   // accumulator buffers per issuer / tile slot (stack mode on tile PAIRS: one each)
    int xq = 0;                                     // stack mode: which half of this group's exchange area the next chunk uses
    auto advance = [&](int w) { if (++accs[w] == nbuf) { accs[w] = 0; phs[w] ^= 1; } };
    int t0, t1, nb;
    for (int it = 0; hl_item_at(p, it, t0, t1, nb); ++it) {
        const bool split = p.split_ok && t1 < 0;
        const int own = p.dual ? grp : 0;        // issuer whose buffer this group reads (and releases)
        const int p_tile = split ? t0 : ((p.dual && grp == 1) ? t1 : t0);
        if (p_tile >= 0) {
            int n, y0, x0;
            hl_tile_origin(p, p_tile, n, y0, x0);
            const int oy = y0 + tr, ox = x0 + cp - 1;
            const bool valid = tr < p.TR && cp >= 1 && cp <= p.TW && oy < p.H && ox < p.W;
            const long long row_idx = (((long long)n * p.H + oy) * p.W + ox) * p.Cout;
            const int tile_in_img = p_tile - n * (p.tiles_y * p.tiles_x);
            float pool_s[2] = {0.f, 0.f};
            unsigned int pool_k[2] = {0u, 0u};
            // POOL: the four quarter warps of this group (32 tile rows each) combine their column results through shared memory
            // (fixed order) and quarter 0 writes ONE partial row per tile: [N][tiles per image][64]
            auto pool_flush = [&](int cmask) {
                float* sc_s = const_cast<float*>(bias_s) + 64 + grp * 384;                 // [3 quarters][2 chunks][32 lanes] sums, then keys
                unsigned int* sc_k = reinterpret_cast<unsigned int*>(sc_s + 192);
                if (quarter > 0) {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        if ((cmask >> c) & 1) { sc_s[((quarter - 1) * 2 + c) * 32 + lane] = pool_s[c]; sc_k[((quarter - 1) * 2 + c) * 32 + lane] = pool_k[c]; }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
                if (quarter == 0) {
                    const long long prow = ((long long)n * p.pool_rows + tile_in_img) * p.Cout;
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        if ((cmask >> c) & 1) {
                            float sv = pool_s[c]; unsigned int kv = pool_k[c];
#pragma unroll
                            for (int qq = 0; qq < 3; ++qq) {
                                sv += sc_s[(qq * 2 + c) * 32 + lane];
                                const unsigned int ko = sc_k[(qq * 2 + c) * 32 + lane];
                                kv = kv > ko ? kv : ko;
                            }
                            p.pool_sum[prow + c * 32 + lane] = sv;
                            p.pool_key[prow + c * 32 + lane] = kv;
                        }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
            };
            // bf16 outputs: row-coalesced stores (tc_store_chunk_quads) — the row bases of the lane's quad, fetched once per tile
            const bool quads = sizeof(OutT) == 2 && !POOL && !p.narrow && p.quads;
            const long long row_base = r > 1 ? ((((long long)n * p.H * r + (long long)oy * r) * ((long long)p.W * r)) + (long long)ox * r) * cq : row_idx;
            long long q_base[4] = {0, 0, 0, 0};
            unsigned q_ok = 0;
            if (quads) {
                const unsigned vm = __ballot_sync(0xffffffffu, valid);
                q_ok = (vm >> (lane & ~3)) & 0xFu;
#pragma unroll
                for (int i = 0; i < 4; ++i) q_base[i] = __shfl_sync(0xffffffffu, row_base, (lane & ~3) + i) + (lane & 3) * 8;
            }
            auto emit = [&](const uint32_t (&v)[32], int c) {
#ifdef SR_WITH_PROBES
                if (p.dbg & 1) return;
#endif
                if (POOL) {                      // (bf16 output, Cout = 64, no residual / mask / shuffle: checked on the host) every lane takes part
                    if (sizeof(OutT) == 2) {
                        const int col = c * 32;
                        float cs; unsigned int ck;
                        hl_store_pool_chunk<ACT>(v, bias_s + col, p.slope, valid, reinterpret_cast<__nv_bfloat16*>(out) + row_idx + col,
                                                 (unsigned int)(0xFFFF - (oy * p.W + ox)) & 0xFFFFu, lane, cs, ck);
                        if (c == 0) { pool_s[0] = cs; pool_k[0] = ck; } else { pool_s[1] = cs; pool_k[1] = ck; }
                    }
                    return;
                }
                if (p.narrow) {              // thin output (RGB / 1-channel critic map): only the first Cout accumulator columns are real
                    if (valid && c == 0) {
                        OutT* o = out + ((((long long)n * p.H + oy) * p.W) + ox) * p.Cout;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (jj < p.Cout) o[jj] = from_f32<OutT>(tc_act<ACT>(__uint_as_float(v[jj]) + bias_s[jj], p.slope));
                    }
                    return;
                }
                const int col = nb * p.block_n + c * 32;
                int chunk_off = col;             // offset of this chunk's first column from the row base (the same for every row)
                if (r > 1) {
                    const int sub = col / cq, ch0 = col - sub * cq;
                    const int si = sub / r, sj = sub - si * r;
                    chunk_off = (si * p.W * r + sj) * cq + ch0;
                }
                if (sizeof(OutT) == 2 && quads) {
                    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out);
                    __nv_bfloat16* const qp[4] = {ob + q_base[0] + chunk_off, ob + q_base[1] + chunk_off, ob + q_base[2] + chunk_off, ob + q_base[3] + chunk_off};
                    const long long idx = row_base + chunk_off;
                    tc_store_chunk_quads<ACT>(v, bias_s + col, p.slope, (res && valid) ? reinterpret_cast<const __nv_bfloat16*>(res) + idx : nullptr,
                                              (mask && valid) ? mask + idx : nullptr, p.mask_slope, qp, q_ok, lane);
                    return;
                }
                if (!valid) return;
                const long long idx = row_base + chunk_off;
                tc_store_chunk<OutT, ACT>(v, bias_s + col, p.slope, res ? res + idx : nullptr, out + idx,
                                          mask ? mask + idx : nullptr, p.mask_slope);
            };
            uint32_t va[32], vb[32];
            if (split) {
                mbar_wait(acc_full + accs[0], phs[0]);
                mbar_wait(acc_full + 2 + accs[1], phs[1]);
                tc_fence_after();
                const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
                const uint32_t base0 = lane_base + (uint32_t)(accs[0] * acc_stride), base1 = lane_base + (uint32_t)((2 + accs[1]) * acc_stride);
                for (int c = grp; c < chunks; c += 2) {
                    tmem_ld32_nowait(base0 + (uint32_t)(c * 32), va);
                    tmem_ld32_nowait(base1 + (uint32_t)(c * 32), vb);
                    tmem_ld_wait(va);
                    tmem_ld_wait(vb);
#pragma unroll
                    for (int k = 0; k < 32; ++k) va[k] = __float_as_uint(__uint_as_float(va[k]) + __uint_as_float(vb[k]));
                    emit(va, c);
                }
                if (POOL) pool_flush(1 << grp);                         // split: this group produced the chunks of parity grp
                tc_fence_before();
                asm volatile("bar.sync 1, 256;" ::: "memory");          // both groups are done reading both buffers
                if (lane == 0) mbar_arrive(acc_empty + grp * 2 + accs[grp]);
            } else {
                const int c_first = p.dual ? 0 : grp, c_step = p.dual ? 1 : 2;
                const int buf = own * 2 + accs[own];
                mbar_wait(acc_full + buf, phs[own]);
                tc_fence_after();
                if (quarter == 0 && lane == 0) HL_TRACE(80 + grp * 16 + it * 2);
                const uint32_t t_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(STACK ? (p.dual ? own : buf) * 192 : buf * acc_stride);
                int c = c_first;
                if (STACK) {
                    // out[j] = S0[j] + S1[j + 1] + S2[j + 2], S_s = the 64-column block whose taps sit s pixels to the right (flip mirrors
                    // the order).  Rows j + 1, j + 2 of lanes 30 / 31 belong to the next quarter's warp: every warp publishes its first
                    // rows of S1 (row 0) and S2 (rows 0, 1) in shared memory, one half of the exchange area per 32-column chunk (so one
                    // barrier per chunk orders writes against the previous reads of the same half).
                    const uint32_t col_s0 = p.flip ? 128u : 0u, col_s2 = p.flip ? 0u : 128u;
                    // tile pairs: group = tile, both chunks; single tiles: both groups on the one tile, chunk = group
#pragma unroll 1
                    for (int cc = p.dual ? 0 : grp; cc < (p.dual ? 2 : grp + 1); ++cc) {
                        float* xch = xch_s + (grp * 2 + xq) * (4 * 3 * 32);              // [quarter][3 rows][32 columns]
                        xq ^= 1;
                        float f[32];
                        tmem_ld32_nowait(t_base + col_s0 + (uint32_t)(cc * 32), va);
                        tmem_ld32_nowait(t_base + 64u + (uint32_t)(cc * 32), vb);
                        tmem_ld_wait(va);
                        tmem_ld_wait(vb);
                        if (lane == 0) {
#pragma unroll
                            for (int k = 0; k < 32; k += 4)
                                *reinterpret_cast<uint4*>(xch + (quarter * 3 + 0) * 32 + k) = make_uint4(vb[k], vb[k + 1], vb[k + 2], vb[k + 3]);
                        }
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            const float nx = __shfl_down_sync(0xffffffffu, __uint_as_float(vb[k]), 1);
                            f[k] = __uint_as_float(va[k]) + (lane < 31 ? nx : 0.f);
                        }
                        tmem_ld32_nowait(t_base + col_s2 + (uint32_t)(cc * 32), va);
                        tmem_ld_wait(va);
                        if (lane < 2) {
#pragma unroll
                            for (int k = 0; k < 32; k += 4)
                                *reinterpret_cast<uint4*>(xch + (quarter * 3 + 1 + lane) * 32 + k) = make_uint4(va[k], va[k + 1], va[k + 2], va[k + 3]);
                        }
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            const float nx = __shfl_down_sync(0xffffffffu, __uint_as_float(va[k]), 2);
                            f[k] += (lane < 30 ? nx : 0.f);
                        }
                        asm volatile("bar.sync %0, 128;" ::"r"(4 + grp) : "memory");
                        if (lane >= 30 && quarter < 3) {
                            // lane 31: S1 row 0 + S2 row 1 of the next warp; lane 30: S2 row 0
                            const float* nx = xch + (quarter + 1) * 3 * 32;
                            const float* a = nx + (lane == 31 ? 0 : 32);
#pragma unroll
                            for (int k = 0; k < 32; ++k) {
                                float add = a[k];
                                if (lane == 31) add += nx[64 + k];
                                f[k] += add;
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 32; ++k) va[k] = __float_as_uint(f[k]);
                        emit(va, cc);
                    }
                    c = chunks;
                }
#ifdef SR_WITH_PROBES
                if (p.dbg & 2) c = chunks;
#endif
                if (c < chunks) tmem_ld32_nowait(t_base + (uint32_t)(c * 32), va);
                while (c < chunks) {
                    tmem_ld_wait(va);
                    if (c + c_step < chunks) tmem_ld32_nowait(t_base + (uint32_t)((c + c_step) * 32), vb);
                    emit(va, c);
                    c += c_step;
                    if (c >= chunks) break;
                    tmem_ld_wait(vb);
                    if (c + c_step < chunks) tmem_ld32_nowait(t_base + (uint32_t)((c + c_step) * 32), va);
                    emit(vb, c);
                    c += c_step;
                }
                if (POOL) pool_flush((STACK && !p.dual) ? (1 << grp) : 3);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + buf);
                if (quarter == 0 && lane == 0) HL_TRACE(80 + grp * 16 + it * 2 + 1);
            }
        }
        // buffer bookkeeping of BOTH issuers (identical in every epilogue warp)
        if (!p.dual) { advance(0); }
        else if (split) { advance(0); advance(1); }
        else { if (t0 >= 0) advance(0); if (t1 >= 0) advance(1); }
    }
}

// POOL: the instantiation whose epilogue also emits the CLAM pooling partials (RAB conv2 launches only) — a separate
// kernel so that its extra register pressure (spills) never touches the other convolutions
template <typename OutT, bool POOL, bool STACK>
__global__ void __launch_bounds__(HL_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = (p.stack ? 192 : p.block_n) * 128;
    const int k_blocks = 9 * p.c_blocks;
    // [resident weights: k_blocks x B] [A ring(s)] [B ring] [barriers] [bias]
    uint8_t* a_base = smem + (p.resident ? (size_t)k_blocks * b_bytes : 0);
    uint8_t* b_base = a_base + (size_t)p.num_a_stages * p.a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)p.num_b_stages * b_bytes);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + HL_MAX_A;
    uint64_t* b_full = a_empty + HL_MAX_A;
    uint64_t* b_empty = b_full + HL_MAX_B;
    uint64_t* acc_full = b_empty + HL_MAX_B;      // [4]: dual -> [issuer][buffer], single -> [buffer]
    uint64_t* acc_empty = acc_full + 4;
    uint64_t* r_full = acc_empty + 4;             // [HL_MAX_RES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_full + HL_MAX_RES);
    float* bias_s = reinterpret_cast<float*>(bars) + 256;                 // 1024 B after the barrier block
    float* xch_s = bias_s + HL_BIAS_MAX;                                  // stack mode only: [2 groups][2 chunks][4 quarters][3 rows][32] floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int issuers = p.dual ? 2 : 1;           // tile slots of an item (stack mode: both served by the ONE issuer warp)
    const int ring_a = p.num_a_stages / issuers;  // each slot owns a private ring of activation stages
    const int b_taps = p.stack ? 3 : 9;           // weight stages per channel block (stack: one per filter row)
    if (threadIdx.x == 0) {
        HL_TRACE(0);
#ifdef SR_WITH_PROBES
        if (p.trace) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); p.trace[(long long)blockIdx.x * HL_TRACE_SLOTS + 3] = gt; }
#endif
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < p.num_a_stages; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
        for (int s = 0; s < p.num_b_stages; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, p.mcast ? 2 : (p.stack ? 1 : issuers)); }
        for (int a = 0; a < 4; ++a) { mbar_init(acc_full + a, 1); mbar_init(acc_empty + a, p.dual ? 4 : HL_EPI_WARPS); }
        if (p.resident) for (int kb = 0; kb < k_blocks; ++kb) mbar_init(r_full + kb, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    pdl_trigger();                 // the next kernel of the stream may be scheduled as this grid's blocks exit
    pdl_wait();                    // everything above overlapped the predecessor's tail; global memory is read from here on
    {
        const int r2 = p.shuffle_r > 1 ? p.shuffle_r * p.shuffle_r : 1;
        const int cq = p.Cout / r2;
        for (int i = threadIdx.x; i < p.Cout; i += HL_THREADS) {
            float b = 0.f;
            if (p.bias) { const int sub = i / cq, ch = i - sub * cq; b = p.bias[r2 > 1 ? ch * r2 + sub : i]; }
            bias_s[i] = b;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) HL_TRACE(1);
    // multicast mode: items of the cluster's longer lane (its even CTA); the barriers of BOTH CTAs are initialised before any
    // remote arrival / multicast box can land
    int mc_rounds = 0;
    uint32_t cta_rank = 0;
    if (STACK && p.mcast) {
        cta_rank = cluster_ctarank();
        const int first = (int)(blockIdx.x & ~1u);
        mc_rounds = first < p.n_items ? (p.n_items - first + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        cluster_sync_all();
    }

    if (warp == 0 && STACK && p.mcast) {
        // ===== TMA producer, multicast mode: own activation tiles; HALF of every weight stage, multicast into both CTAs =====
        if (lane == 0) {
            int sa = 0; uint32_t pa = 0;
            int sb = 0; uint32_t pb = 0;
            for (int it = 0; it < mc_rounds; ++it) {
                int t0, t1, nb;
                const bool real = hl_item_at(p, it, t0, t1, nb);
                int n = 0, y0 = 0, x0 = 0;
                if (real) hl_tile_origin(p, t0, n, y0, x0);
                for (int cb = 0; cb < p.c_blocks; ++cb) {
                    if (real) {
                        mbar_wait(a_empty + sa, pa ^ 1);
                        mbar_expect_tx(a_full + sa, (uint32_t)p.a_box_bytes);
                        tma_load_4d(a_base + (size_t)sa * p.a_stage_bytes, &map_a, a_full + sa, cb * 64, x0 - 1, y0 - 1, n);
                        if (cb == 0) HL_TRACE(8 + it * 2);
                        if (++sa == ring_a) { sa = 0; pa ^= 1; }
                    }
                    for (int ky = 0; ky < 3; ++ky) {
                        mbar_wait(b_empty + sb, pb ^ 1);                       // released by BOTH CTAs' issuers
                        mbar_expect_tx(b_full + sb, (uint32_t)b_bytes);        // both halves land here (the peer's may arrive first)
                        tma_load_2d_mcast(b_base + (size_t)sb * b_bytes + (size_t)cta_rank * (96 * 128), &map_b, b_full + sb, cb * 64,
                                          ky * 192 + (int)cta_rank * 96, (uint16_t)3);
                        if (++sb == p.num_b_stages) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int sa[2] = {0, 0}; uint32_t pa[2] = {0, 0};
            int sb = 0; uint32_t pb = 0;
            int t0, t1, nb;
            for (int it = 0; hl_item_at(p, it, t0, t1, nb); ++it) {
                const bool split = p.split_ok && t1 < 0;
                for (int cb = 0; cb < p.c_blocks; ++cb) {
                    for (int w = 0; w < issuers; ++w) {
                        const int tile = split ? ((cb & 1) == w ? t0 : -1) : (w ? t1 : t0);     // split: channel block cb belongs to issuer cb & 1
                        if (tile < 0) continue;
                        int n, y0, x0;
                        hl_tile_origin(p, tile, n, y0, x0);
                        const int st = w * ring_a + sa[w];
                        mbar_wait(a_empty + st, pa[w] ^ 1);
                        mbar_expect_tx(a_full + st, (uint32_t)p.a_box_bytes);
                        tma_load_4d(a_base + (size_t)st * p.a_stage_bytes, &map_a, a_full + st, cb * 64, x0 - 1, y0 - 1, n);
                        if (cb == 0) HL_TRACE(8 + it * 2 + w);
                        if (++sa[w] == ring_a) { sa[w] = 0; pa[w] ^= 1; }
                    }
                    for (int tap = 0; tap < b_taps; ++tap) {
                        const int kb = cb * 9 + tap;
                        if (p.stack) {
                            mbar_wait(b_empty + sb, pb ^ 1);
                            mbar_expect_tx(b_full + sb, (uint32_t)b_bytes);
                            tma_load_2d(b_base + (size_t)sb * b_bytes, &map_b, b_full + sb, cb * 64, tap * 192);
                            if (++sb == p.num_b_stages) { sb = 0; pb ^= 1; }
                        } else if (p.resident) {
                            if (it == 0) {
                                mbar_expect_tx(r_full + kb, (uint32_t)b_bytes);
                                tma_load_2d(smem + (size_t)kb * b_bytes, &map_b, r_full + kb, cb * 64, tap * p.Cout + nb * p.block_n);
                            }
                        } else {
                            mbar_wait(b_empty + sb, pb ^ 1);
                            mbar_expect_tx(b_full + sb, (uint32_t)b_bytes);
                            tma_load_2d(b_base + (size_t)sb * b_bytes, &map_b, b_full + sb, cb * 64, tap * p.Cout + nb * p.block_n);
                            if (++sb == p.num_b_stages) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp <= 2) {
        // ===== MMA issuers (warp 1, and warp 2 in dual mode).  One thread issuing short dependent groups of tcgen05.mma
        // sustains ~100 clk per instruction; two issuers on independent accumulators reach the 64 clk floor of an
        // M=128 x N=128 instruction (profiles/r01_umma_issue_rate.txt).  The warp walks the schedule converged. =====
        const int w = warp - 1;
        if (STACK) {
            // ===== stack mode: ONE issuer, N = 192.  Tile PAIRS (p.dual): both tiles of an item per weight stage, one accumulator
            // each (2 x 192 columns) — half the weight traffic, but the epilogue cannot overlap the next item.  SINGLE tiles: the
            // accumulator is double buffered (2 x 192), the epilogue of tile i runs under the MMAs of tile i + 1. =====
            if (w == 0) {
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(192 >> 3) << 17) | ((128u >> 4) << 24);
                int sa[2] = {0, 0}; uint32_t pa[2] = {0, 0};
                int sb = 0; uint32_t pb = 0;
                uint32_t acc_phase[2] = {0, 0};
                int acc = 0;                     // single tiles: accumulator buffer of the next item
                int t0, t1, nb;
                for (int it = 0;; ++it) {
                    // multicast mode: `mc_rounds` items in both CTAs of the cluster; a CTA without a tile in a round still waits for and
                    // releases every weight stage
                    const bool real = hl_item_at(p, it, t0, t1, nb);
                    if (p.mcast ? it >= mc_rounds : !real) break;
                    const bool has[2] = {real && t0 >= 0, real && p.dual && t1 >= 0};
                    // barrier index / TMEM column of slot g: pairs -> (g * 2, g * 192); single tiles -> (acc, acc * 192)
                    const int bar_of[2] = {p.dual ? 0 : acc, 2};
                    const uint32_t col_of[2] = {p.dual ? 0u : (uint32_t)acc * 192u, 192u};
#pragma unroll
                    for (int g = 0; g < 2; ++g)
                        if (has[g]) mbar_wait(acc_empty + bar_of[g], acc_phase[p.dual ? g : acc] ^ 1);
                    tc_fence_after();
                    if (lane == 0) HL_TRACE(32 + it * 3);
                    for (int cb = 0; cb < p.c_blocks; ++cb) {
                        uint64_t adesc[2] = {0, 0};
#pragma unroll
                        for (int g = 0; g < 2; ++g)
                            if (has[g]) {
                                const int st = g * ring_a + sa[g];
                                mbar_wait(a_full + st, pa[g]);
                                adesc[g] = make_kmajor_sw128_desc(smem_u32(a_base + (size_t)st * p.a_stage_bytes));
                            }
                        if (lane == 0 && cb == 0) HL_TRACE(32 + it * 3 + 1);
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            mbar_wait(b_full + sb, pb);
                            tc_fence_after();
                            if (lane == 0) {
                                const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(b_base + (size_t)sb * b_bytes));
                                const uint32_t row_off = (uint32_t)((p.flip ? 2 - ky : ky) * p.TWp) * 8u;
                                const uint32_t accum = (cb | ky) ? 1u : 0u;
                                if (has[0]) umma_f16_x4(tmem_base + col_of[0], adesc[0] + row_off, bdesc, idesc, accum);
                                if (has[1]) umma_f16_x4(tmem_base + col_of[1], adesc[1] + row_off, bdesc, idesc, accum);
                                if (p.mcast) umma_commit_mcast(b_empty + sb, (uint16_t)3); else umma_commit(b_empty + sb);
                            }
                            __syncwarp();
                            if (++sb == p.num_b_stages) { sb = 0; pb ^= 1; }
                        }
#pragma unroll
                        for (int g = 0; g < 2; ++g)
                            if (has[g]) {
                                if (lane == 0) umma_commit(a_empty + g * ring_a + sa[g]);
                                if (++sa[g] == ring_a) { sa[g] = 0; pa[g] ^= 1; }
                            }
                        __syncwarp();
                    }
#pragma unroll
                    for (int g = 0; g < 2; ++g)
                        if (has[g]) {
                            if (lane == 0) umma_commit(acc_full + bar_of[g]);
                            acc_phase[p.dual ? g : acc] ^= 1;
                        }
                    if (!p.dual && has[0]) acc ^= 1;
                    if (lane == 0) HL_TRACE(32 + it * 3 + 2);
                    __syncwarp();
                }
            }
        } else if (w < issuers) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((128u >> 4) << 24);
            uint32_t tap_off[9];      // descriptor start-address offsets (16-byte units) of the nine tap views
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap % 3;
                const int oh = p.flip ? 2 - ky : ky, ow = p.flip ? 2 - kx : kx;
                tap_off[tap] = (uint32_t)(oh * p.TWp + ow) * 8u;
            }
            const uint32_t b_step = (uint32_t)b_bytes >> 4;
            const int acc_stride = p.dual ? 128 : 256;
            int sa = 0; uint32_t pa = 0;
            int sb = 0; uint32_t pb = 0;
            int acc = 0; uint32_t acc_phase = 0;
            bool res_ready = false;
            int t0, t1, nb;
            for (int it = 0; hl_item_at(p, it, t0, t1, nb); ++it) {
                const bool split = p.split_ok && t1 < 0;
                const bool has = split || (w ? t1 : t0) >= 0;
                bool first = true;            // first MMA group of this issuer in this item overwrites the accumulator
                uint32_t d_tmem = 0;
                if (has) {
                    mbar_wait(acc_empty + w * 2 + acc, acc_phase ^ 1);
                    tc_fence_after();
                    d_tmem = tmem_base + (uint32_t)((w * 2 + acc) * acc_stride);
                    if (lane == 0) HL_TRACE(32 + w * 24 + it * 3);
                }
                for (int cb = 0; cb < p.c_blocks; ++cb) {
                    uint64_t adesc0 = 0;
                    const int st = w * ring_a + sa;
                    const bool mine = split ? ((cb & 1) == w) : has;       // this issuer multiplies channel block cb
                    if (mine) {
                        mbar_wait(a_full + st, pa);
                        if (lane == 0 && cb == 0) HL_TRACE(32 + w * 24 + it * 3 + 1);
                        adesc0 = make_kmajor_sw128_desc(smem_u32(a_base + (size_t)st * p.a_stage_bytes));
                    }
                    if (p.resident) {
                        if (has) {
                            const uint64_t bdesc0 = make_kmajor_sw128_desc(smem_u32(smem + (size_t)cb * 9 * b_bytes));
                            if (!res_ready) {
#pragma unroll
                                for (int tap = 0; tap < 9; ++tap) {
                                    mbar_wait(r_full + cb * 9 + tap, 0);
                                    tc_fence_after();
                                    if (lane == 0) umma_f16_x4(d_tmem, adesc0 + tap_off[tap], bdesc0 + (uint64_t)(tap * b_step), idesc, (cb | tap) ? 1u : 0u);
                                    __syncwarp();
                                }
                            } else {
                                tc_fence_after();
                                if (lane == 0) {
#pragma unroll
                                    for (int tap = 0; tap < 9; ++tap)
                                        umma_f16_x4(d_tmem, adesc0 + tap_off[tap], bdesc0 + (uint64_t)(tap * b_step), idesc, (cb | tap) ? 1u : 0u);
                                }
                                __syncwarp();
                            }
                        }
                    } else {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(b_full + sb, pb);
                            tc_fence_after();
                            if (lane == 0) {
                                if (mine) {
                                    const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(b_base + (size_t)sb * b_bytes));
                                    umma_f16_x4(d_tmem, adesc0 + tap_off[tap], bdesc, idesc, first ? 0u : 1u);
                                }
                                umma_commit(b_empty + sb);      // both issuers release every weight stage (count = issuers)
                            }
                            if (mine) first = false;
                            __syncwarp();
                            if (++sb == p.num_b_stages) { sb = 0; pb ^= 1; }
                        }
                    }
                    if (mine) {
                        if (lane == 0) umma_commit(a_empty + st);
                        __syncwarp();
                        if (++sa == ring_a) { sa = 0; pa ^= 1; }
                    }
                }
                if (has) {
                    if (p.resident) res_ready = true;
                    if (lane == 0) { umma_commit(acc_full + w * 2 + acc); HL_TRACE(32 + w * 24 + it * 3 + 2); }
                    __syncwarp();
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        const int e = warp - 3;                    // 0..7
        const int quarter = warp & 3, grp = e >> 2;
        switch (p.act) {
            case SR_ACT_LRELU: hl_epilogue<OutT, SR_ACT_LRELU, POOL, STACK>(p, tmem_base, acc_full, acc_empty, bias_s, xch_s, quarter, grp, lane); break;
            case SR_ACT_RELU: hl_epilogue<OutT, SR_ACT_RELU, POOL, STACK>(p, tmem_base, acc_full, acc_empty, bias_s, xch_s, quarter, grp, lane); break;
            case SR_ACT_SIGMOID: hl_epilogue<OutT, SR_ACT_SIGMOID, false, false>(p, tmem_base, acc_full, acc_empty, bias_s, xch_s, quarter, grp, lane); break;
            default: hl_epilogue<OutT, SR_ACT_NONE, POOL, STACK>(p, tmem_base, acc_full, acc_empty, bias_s, xch_s, quarter, grp, lane); break;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (STACK && p.mcast) cluster_sync_all();      // the peer may still signal this CTA's barriers / this CTA the peer's
    if (threadIdx.x == 0) {
        HL_TRACE(2);
#ifdef SR_WITH_PROBES
        if (p.trace) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); p.trace[(long long)blockIdx.x * HL_TRACE_SLOTS + 4] = gt; }
#endif
    }
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int g_hl_sms = 0;

static bool hl_cout_ok(int Cout) { return Cout % 64 == 0 || Cout <= 4; }

bool conv_halo_supported(const sr_conv_desc* d, bool dgrad) {
    if (d->in_dtype != SR_BF16) return false;
    if (d->kh != 3 || d->kw != 3 || d->stride != 1 || d->pad != 1) return false;
    const int Cs = dgrad ? d->Cout : d->Cin, Cd = dgrad ? d->Cin : d->Cout;
    if (Cs % 64 != 0 || !hl_cout_ok(Cd) || Cd > HL_BIAS_MAX) return false;
    const int r = d->shuffle_r > 1 ? d->shuffle_r : 1;
    if (!dgrad && r > 1 && (Cd % (r * r) != 0 || (Cd / (r * r)) % 32 != 0 || Cd <= 4)) return false;
    if ((long long)d->N * d->H * d->W >= (1ll << 31)) return false;
    return true;
}

// geometry of the pixel tiling (depends on the map size only)
static void hl_tiling(int H, int W, int& tiles_x, int& TW, int& TR, int& tiles_y) {
    tiles_x = (int)cdiv(W, 62);
    TW = (int)cdiv(W, tiles_x);
    TR = 129 / (TW + 2);
    if (TR > H) TR = H;
    tiles_y = (int)cdiv(H, TR);
}

// rows per image of the pooling partials a forward launch emits (0: this convolution cannot emit them)
int conv_halo_pool_rows(const sr_conv_desc* d) {
    if (!conv_halo_supported(d, false) || d->Cout != 64 || d->shuffle_r > 1 || d->out_dtype != SR_BF16) return 0;
    if ((long long)d->H * d->W > 65535) return 0;
    int tx, TW, TR, ty;
    hl_tiling(d->H, d->W, tx, TW, TR, ty);
    return tx * ty;
}

int conv_halo_run(const sr_conv_desc* d, bool dgrad, const void* src, const void* w, const float* bias,
                  const void* residual, void* dst, cudaStream_t st, const void* mask, float mask_slope) {
    int rc = load_driver_fns();
    if (rc != SR_OK) return rc;
    if (!g_hl_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_hl_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int Cs = dgrad ? d->Cout : d->Cin, Cd = dgrad ? d->Cin : d->Cout;
    HaloParams p;
    memset(&p, 0, sizeof(p));
    p.N = d->N; p.H = d->H; p.W = d->W; p.Cout = Cd;
    // tile = TR rows x TW columns of one image with TR * (TW + 2) <= 129 MMA rows; column strips of <= 62 pixels keep
    // the halo tile (and its shared-memory stage) small for every map width
    hl_tiling(d->H, d->W, p.tiles_x, p.TW, p.TR, p.tiles_y);
    p.TWp = p.TW + 2;
    p.p_tiles = d->N * p.tiles_y * p.tiles_x;
    p.c_blocks = Cs / 64;
    p.flip = dgrad ? 1 : 0;
    p.quads = option("SR_HALO_QUADS", 1) ? 1 : 0;
    p.act = dgrad ? SR_ACT_NONE : d->act; p.slope = d->slope;
    p.shuffle_r = dgrad ? 0 : d->shuffle_r;
    p.bias = bias; p.residual = residual; p.out = dst;
    p.mask = mask; p.mask_slope = mask_slope;
#ifdef SR_WITH_PROBES
    p.trace = g_hl_trace; p.dbg = g_hl_dbg;
#endif
    if (!dgrad && d->pool_sum && d->pool_key) {
        const int rows = conv_halo_pool_rows(d);
        if (!rows || residual || mask) { set_error("conv_halo: pooling partials need a plain bf16 forward convolution with 64 output channels"); return SR_ERR_UNSUPPORTED; }
        p.pool_sum = (float*)d->pool_sum; p.pool_key = (unsigned int*)d->pool_key; p.pool_rows = rows;
    }
    p.a_box_bytes = (p.TR + 2) * p.TWp * 128;
    // the tap views of junk rows reach up to pixel row 129 + 2*TWp of the stage: keep that inside the stage
    const int rows_needed = 130 + 2 * p.TWp, rows_loaded = (p.TR + 2) * p.TWp;
    p.a_stage_bytes = (int)cdiv((rows_needed > rows_loaded ? rows_needed : rows_loaded) * 128, 1024) * 1024;
    const int k_blocks = 9 * p.c_blocks;

    // Column-block width and issue mode (profiles/r01_umma_issue_rate.txt):
    //   resident + dual : the [9*Cin] x 128 weight slab stays in smem (Cin = 64), two issuer warps on two pixel tiles
    //   N = 256 single  : one issuer already sustains the 128 clk floor of an N=256 instruction (weights streamed)
    //   dual streamed   : Cout = 64 / 128 (one column block): two pixel tiles share every streamed weight stage
    //   single streamed : everything else (Cout = 192 * k, 128 * odd)
    // Thin outputs (Cd <= 4: conv3 64->3, the critic's 512->1 map, the input gradients of the RGB-side convolutions) run as
    // ONE 64-wide column block: the weight box of a tap then also covers rows of the following taps (or zero fill past
    // the end) — junk accumulator columns that are simply never stored.
    p.narrow = Cd <= 4 ? 1 : 0;
    if (p.narrow && (residual || mask || p.shuffle_r > 1)) { set_error("conv_halo: thin outputs support bias / activation only"); return SR_ERR_UNSUPPORTED; }
    // stack mode (64 output channels fed by >= 128 input channels: RAB conv2 and the input gradient of conv1)
    // ... and only when the nine-tap mode could NOT keep the [9*Cin] x 64 weight slab resident (Cin = 128: 147 KB fits, and loading the
    // weights once beats streaming them per tile pair — 128->64 @108^2 input gradient: 41.7 us resident vs 49.8 us stacked)
    const bool slab_fits = (long long)k_blocks * 64 * 128 + 2 * p.a_stage_bytes <= HL_TILE_BUDGET && k_blocks <= HL_MAX_RES;
    const bool stack = Cd == 64 && Cs >= 128 && !slab_fits && d->out_dtype == SR_BF16 && p.shuffle_r <= 1 && p.TR * p.TWp <= 128 &&
                       option("SR_HALO_STACK", 1);
    int bn;
    if (p.narrow) bn = 64; else if (Cd % 128 == 0) bn = 128; else if (Cd % 192 == 0) bn = 192; else bn = 64;
    long long res_bytes = (long long)k_blocks * bn * 128;
    bool resident = !stack && k_blocks <= HL_MAX_RES && bn <= 128 && res_bytes + 2 * p.a_stage_bytes <= HL_TILE_BUDGET;
    if (!resident && Cd % 256 == 0) bn = 256;
    p.block_n = bn;
    p.n_blocks = p.narrow ? 1 : Cd / bn;
    const int b_bytes = (stack ? 192 : bn) * 128;
    res_bytes = (long long)k_blocks * b_bytes;
    int grid = g_hl_sms;
    if (resident) {
        grid -= grid % p.n_blocks;
        if (p.p_tiles * p.n_blocks <= grid) resident = false;       // every CTA would run a single tile: nothing to keep
    }
    p.resident = resident ? 1 : 0;
    p.stack = stack ? 1 : 0;
    p.dual = (bn <= 128 && (resident || p.n_blocks == 1)) ? 1 : 0;
    // stack mode on single tiles (double-buffered accumulator) instead of tile pairs: measured SLOWER (K2 forward 37.5k vs 35.9k
    // cycles per CTA, profiles/r02_halo_stack_variants.txt) — every SM has to take in the whole 295 KB weight stream per tile,
    // and that intake (~48 B/clk per SM), not the L2, is the limit, so sharing the stream between two tiles of the same CTA wins
    if (stack && !option("SR_HALO_STACK_PAIRS", 1)) p.dual = 0;
    // experiment option: resident-weights layers on ONE issuer with a two-deep activation ring
    if (option("SR_HALO_NODUAL_RES", 0) && resident) p.dual = 0;
    const int lane_ctas = resident ? grid / p.n_blocks : grid;
    if (p.dual) {
        const int pairs_total = p.p_tiles / 2;
        const int full = (pairs_total / lane_ctas) * lane_ctas;      // pair items of the complete rounds
        const int rem_tiles = p.p_tiles - 2 * full;
        if (rem_tiles <= lane_ctas) { p.n_pair_items = full; p.n_items = full + rem_tiles; }
        else { p.n_pair_items = (int)cdiv(p.p_tiles, 2); p.n_items = p.n_pair_items; }
    } else {
        p.n_pair_items = 0;
        p.n_items = resident ? p.p_tiles : p.p_tiles * p.n_blocks;
    }
    p.split_ok = (option("SR_HALO_SPLITK", 1) && p.dual && !resident && !stack && p.c_blocks >= 2) ? 1 : 0;
    const int total_items = resident ? p.n_items * p.n_blocks : p.n_items;
    if (total_items < grid) { grid = total_items; if (resident) grid -= grid % p.n_blocks; }
    if (grid < 1) grid = p.n_blocks;

    size_t tile_bytes;
    const int xch_bytes = stack ? 2 * 2 * 4 * 3 * 32 * 4 : 0;      // [group][chunk][quarter][3 rows][32] floats
    if (stack) {
        // the junk rows' views may read past the rows a stage loads (into the next stage / the weight ring): tight stages
        p.a_stage_bytes = (int)cdiv(rows_loaded * 128, 1024) * 1024;
        const int na = 4;
        int nbs = (HL_TILE_BUDGET - xch_bytes - na * p.a_stage_bytes) / b_bytes;
        if (nbs > HL_MAX_B) nbs = HL_MAX_B;
        if (nbs < 2) { set_error("conv_halo: tile does not fit shared memory (stack mode, TWp=%d TR=%d)", p.TWp, p.TR); return SR_ERR_UNSUPPORTED; }
        p.num_a_stages = na; p.num_b_stages = nbs;
        tile_bytes = (size_t)na * p.a_stage_bytes + (size_t)nbs * b_bytes;
    } else if (resident) {
        int na = (int)((HL_TILE_BUDGET - res_bytes) / p.a_stage_bytes);
        if (na > HL_MAX_A) na = HL_MAX_A;
        if (p.dual) na -= na % 2;
        p.num_a_stages = na; p.num_b_stages = 0;
        tile_bytes = (size_t)res_bytes + (size_t)na * p.a_stage_bytes;
    } else {
        int na = p.dual ? 4 : 3;
        while (na > 2 && (HL_TILE_BUDGET - na * p.a_stage_bytes) / b_bytes < 4) na -= p.dual ? 2 : 1;
        int nbs = (HL_TILE_BUDGET - na * p.a_stage_bytes) / b_bytes;
        if (nbs > HL_MAX_B) nbs = HL_MAX_B;
        if (nbs < 2) { set_error("conv_halo: tile does not fit shared memory (TWp=%d TR=%d block_n=%d)", p.TWp, p.TR, bn); return SR_ERR_UNSUPPORTED; }
        p.num_a_stages = na; p.num_b_stages = nbs;
        tile_bytes = (size_t)na * p.a_stage_bytes + (size_t)nbs * b_bytes;
    }
    const size_t smem = 1024 + tile_bytes + 1024 + HL_BIAS_MAX * 4 + xch_bytes;
    const bool out_bf16 = d->out_dtype == SR_BF16;
    if (stack && !out_bf16) { set_error("conv_halo: internal: stack mode needs a bf16 output"); return SR_ERR_UNSUPPORTED; }
    const bool pdl = option("SR_PDL", 0) != 0;
    static bool attr_set[5] = {false, false, false, false, false};
#define HL_ATTR(slot, ...)                                                                                                \
    do {                                                                                                                  \
        if (!attr_set[slot]) { cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448); attr_set[slot] = true; } \
    } while (0)
    // multicast mode (stack mode on single tiles): clusters of two CTAs, as many as can be resident at once
    int cluster = 1;
    // (measured: no gain — 42.5k cycles; the multicast halves the L2 reads but not what each SM must take in; kept as an option)
    if (stack && !p.dual && option("SR_HALO_MCAST", 0) && p.n_items >= 2 * 2) {
        static int max_clusters[2] = {-1, -1};
        const int which = p.pool_sum ? 1 : 0;
        if (max_clusters[which] < 0) {
            if (which) { HL_ATTR(4, conv_halo_kernel<__nv_bfloat16, true, true>); max_clusters[1] = max_active_clusters(conv_halo_kernel<__nv_bfloat16, true, true>, HL_THREADS, smem, 2, g_hl_sms); }
            else { HL_ATTR(3, conv_halo_kernel<__nv_bfloat16, false, true>); max_clusters[0] = max_active_clusters(conv_halo_kernel<__nv_bfloat16, false, true>, HL_THREADS, smem, 2, g_hl_sms); }
        }
        int g2 = 2 * max_clusters[which];
        if (g2 > p.n_items) g2 = p.n_items & ~1;
        if (g2 >= 2) { cluster = 2; grid = g2; p.mcast = 1; }
    }
    alignas(64) CUtensorMap map_a, map_b;
    rc = make_tiled4d_map(&map_a, src, d->N, d->H, d->W, Cs, p.TWp, p.TR + 2);
    if (rc != SR_OK) return rc;
    rc = make_tiled2d_map(&map_b, w, (uint64_t)9 * Cd, (uint64_t)Cs, (uint32_t)(stack ? (p.mcast ? 96 : 192) : bn));
    if (rc != SR_OK) return rc;
#define HL_LAUNCH(slot, ...)                                                                                              \
    do {                                                                                                                  \
        HL_ATTR(slot, __VA_ARGS__);                                                                                       \
        launch_ex(__VA_ARGS__, dim3(grid), dim3(HL_THREADS), smem, st, pdl, cluster, map_a, map_b, p);                    \
    } while (0)
    if (p.pool_sum && stack) HL_LAUNCH(4, conv_halo_kernel<__nv_bfloat16, true, true>);
    else if (p.pool_sum) HL_LAUNCH(2, conv_halo_kernel<__nv_bfloat16, true, false>);
    else if (stack) HL_LAUNCH(3, conv_halo_kernel<__nv_bfloat16, false, true>);
    else if (out_bf16) HL_LAUNCH(0, conv_halo_kernel<__nv_bfloat16, false, false>);
    else HL_LAUNCH(1, conv_halo_kernel<float, false, false>);
#undef HL_LAUNCH
#undef HL_ATTR
    count_launch();
    return check_launch("conv_halo_kernel");
}

}  // namespace sr
