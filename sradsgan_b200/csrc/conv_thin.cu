// "Thin" convolutions: one side of the layer has <= 4 channels (RGB in / RGB out / 1-channel critic map).
// These layers carry <1 % of the FLOPs but touch the largest tensors (216x216 x batch), so they are HBM
// bound: an implicit-GEMM tile wastes >90 % of its lanes on them.  Direct kernels:
//   thin_cin  : Cs <= 4  -> Cd wide   (conv1, MSB.conv1/2.0/3, D.model.0, VGG.0 forward; conv3 dgrad)
//   thin_cout : Cs wide  -> Cd <= 4   (conv3, D.model.25 forward; D.model.0 / VGG.0 dgrad)
//   thin_wgrad_ci / thin_wgrad_co : the matching weight gradients (fp32 atomics into OIHW)
// stride 1 only; k in {1,3,5,7}.  Weights use the same packed layouts as the implicit-GEMM kernels
// ([tap][Cd][Cs]); `flip` mirrors the tap offsets so that a stride-1 dgrad is the forward kernel on dy.
#include <algorithm>

#include "common.cuh"

namespace sr {

struct ThinParams {
    int N, H, W;          // spatial size (same for source and destination: stride 1, "same" padding handled by pad)
    int Ho, Wo;
    int Cs, Cd;           // source / destination channels
    int k, pad, flip;
    int act;
    float slope;
};

// ---------------------------------------------------------------------------------------------
// Cs <= 4 -> Cd (multiple of 64).  Register tile: one thread = 4 consecutive pixels of a row x 8 channels per pass, so
// every 16-byte weight read from shared memory feeds 16 FMAs (a 1-pixel tile is shared-memory-bandwidth bound).
// block = 64 pixel groups x 4 channel groups; a pass covers 32 channels (channel = pass*32 + cg*8 + j*4 + k).
// The weights of a pass are laid out [j][cg][k] in shared memory, so the four channel groups of a warp read four
// ADJACENT 16-byte chunks (conflict-free + broadcast; the natural [channel] order put cg 0/2 and 1/3 on the same banks:
// 50 % of the shared wavefronts were replays), and the 32-accumulator tile keeps the kernel under 128 registers =
// two resident blocks per SM (the former 4 x 16 tile: 173 registers, one block, 12 % of the warp slots, FMA pipe 28 %).
// ---------------------------------------------------------------------------------------------
constexpr int TCI_PX = 4;

template <int ACT> __device__ __forceinline__ float thin_act(float x, float slope) {
    if (ACT == SR_ACT_LRELU) return fmaxf(x, x * slope);
    if (ACT == SR_ACT_RELU) return fmaxf(x, 0.f);
    if (ACT == SR_ACT_SIGMOID) return 1.f / (1.f + __expf(-x));
    return x;
}

template <typename TIn, typename TOut, int K, int CS, int ACT>
__global__ void __launch_bounds__(256, 2)
thin_cin_kernel(ThinParams p, const TIn* __restrict__ src, const TIn* __restrict__ wpk, const float* __restrict__ bias,
                TOut* __restrict__ dst) {
    extern __shared__ __align__(16) float thin_smem[];     // ws[tap][cs][pass][j][cg][k], then bias[Cd]
    constexpr int TAPS = K * K, NIN = TAPS * CS, WIN = TCI_PX + K - 1;
    for (int i = threadIdx.x; i < NIN * p.Cd; i += 256) {
        const int cd = i % p.Cd; const int r = i / p.Cd; const int cs = r % CS; const int tap = r / CS;
        const int ky = tap / K, kx = tap - ky * K;
        const int wt = p.flip ? (K - 1 - ky) * K + (K - 1 - kx) : tap;      // window offset (ky,kx) reads weight tap wt
        const int ps = cd >> 5, q = cd & 31;                                // channel cd = ps*32 + cgi*8 + j*4 + k
        const int cgi = q >> 3, j = (q >> 2) & 1, k = q & 3;
        thin_smem[r * p.Cd + ps * 32 + (j * 4 + cgi) * 4 + k] = to_f32<TIn>(wpk[((long long)wt * p.Cd + cd) * CS + cs]);
    }
    float* bias_s = thin_smem + NIN * p.Cd;
    for (int i = threadIdx.x; i < p.Cd; i += 256) bias_s[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int groups_x = (p.Wo + TCI_PX - 1) / TCI_PX;
    const long long G = (long long)p.N * p.Ho * groups_x;
    const long long grp = (long long)blockIdx.x * 64 + (threadIdx.x >> 2);
    if (grp >= G) return;
    const int cg = threadIdx.x & 3;
    const int gx = (int)(grp % groups_x); const long long q = grp / groups_x;
    const int oy = (int)(q % p.Ho); const int n = (int)(q / p.Ho);
    const int ox0 = gx * TCI_PX;
    // input window: K rows x (4 + K - 1) columns x CS channels (zero outside the image)
    float in[K][WIN][CS];
#pragma unroll
    for (int r = 0; r < K; ++r) {
        const int sy = oy + r - p.pad;
#pragma unroll
        for (int c = 0; c < WIN; ++c) {
            const int sx = ox0 + c - p.pad;
            const bool ok = sy >= 0 && sy < p.H && sx >= 0 && sx < p.W;
            const TIn* sp = src + (((long long)n * p.H + (ok ? sy : 0)) * p.W + (ok ? sx : 0)) * CS;
#pragma unroll
            for (int cs = 0; cs < CS; ++cs) in[r][c][cs] = ok ? to_f32<TIn>(sp[cs]) : 0.f;
        }
    }
    for (int ps = 0; ps * 32 < p.Cd; ++ps) {
        const int c0 = ps * 32 + cg * 8;
        float acc[TCI_PX][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float b = bias_s[c0 + j];
#pragma unroll
            for (int px = 0; px < TCI_PX; ++px) acc[px][j] = b;
        }
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx)
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) {
                    const float4* wr = reinterpret_cast<const float4*>(thin_smem + (size_t)((ky * K + kx) * CS + cs) * p.Cd + ps * 32) + cg;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float4 w = wr[j * 4];
#pragma unroll
                        for (int px = 0; px < TCI_PX; ++px) {
                            const float v = in[ky][px + kx][cs];
                            acc[px][j * 4] = fmaf(v, w.x, acc[px][j * 4]); acc[px][j * 4 + 1] = fmaf(v, w.y, acc[px][j * 4 + 1]);
                            acc[px][j * 4 + 2] = fmaf(v, w.z, acc[px][j * 4 + 2]); acc[px][j * 4 + 3] = fmaf(v, w.w, acc[px][j * 4 + 3]);
                        }
                    }
                }
#pragma unroll
        for (int px = 0; px < TCI_PX; ++px) {
            if (ox0 + px < p.Wo) {
                TOut* o = dst + ((((long long)n * p.Ho + oy) * p.Wo) + ox0 + px) * p.Cd + c0;
#pragma unroll
                for (int j = 0; j < 8; j += 4)
                    store4<TOut>(o + j, thin_act<ACT>(acc[px][j], p.slope), thin_act<ACT>(acc[px][j + 1], p.slope),
                                 thin_act<ACT>(acc[px][j + 2], p.slope), thin_act<ACT>(acc[px][j + 3], p.slope));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Cs (multiple of 64) -> Cd <= 4.  block = 8 warps, tile = 8 rows x 32 pixels, halo tile of one 64-channel
// block of the source in shared memory; a warp owns one tile row, lanes split the 64 channels.
// ---------------------------------------------------------------------------------------------
constexpr int TC_TH = 8, TC_TW = 32;

template <typename TIn, typename TOut, int K>
__global__ void __launch_bounds__(256)
thin_cout_kernel(ThinParams p, const TIn* __restrict__ src, const TIn* __restrict__ wpk, const float* __restrict__ bias,
                 TOut* __restrict__ dst) {
    extern __shared__ __align__(16) unsigned char thin_raw[];
    TIn* tile = reinterpret_cast<TIn*>(thin_raw);                       // [(TH+K-1)][(TW+K-1)][64]
    constexpr int TAPS = K * K;
    constexpr int th = TC_TH + K - 1, tw = TC_TW + K - 1;
    const int tiles_x = (p.Wo + TC_TW - 1) / TC_TW, tiles_y = (p.Ho + TC_TH - 1) / TC_TH;
    int b = blockIdx.x;
    const int tx = b % tiles_x; b /= tiles_x;
    const int ty = b % tiles_y; const int n = b / tiles_y;
    const int x0 = tx * TC_TW, y0 = ty * TC_TH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float res[4] = {0.f, 0.f, 0.f, 0.f};                                // outputs of pixel (row = warp, col = lane)
    for (int cb = 0; cb < p.Cs; cb += 64) {
        __syncthreads();
        for (int i = threadIdx.x; i < th * tw * 8; i += 256) {         // 8 channels per thread-iteration
            const int v = i & 7; const int pp = i >> 3;
            const int c = pp % tw, r = pp / tw;
            const int sy = y0 + r - p.pad, sx = x0 + c - p.pad;
            TIn* d = tile + ((size_t)r * tw + c) * 64 + v * 8;
            if (sy >= 0 && sy < p.H && sx >= 0 && sx < p.W) {
                const TIn* sp = src + (((long long)n * p.H + sy) * p.W + sx) * p.Cs + cb + v * 8;
#pragma unroll
                for (int j = 0; j < 8; j += 4) { float t4[4]; load4<TIn>(sp + j, t4); store4<TIn>(d + j, t4[0], t4[1], t4[2], t4[3]); }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j += 4) store4<TIn>(d + j, 0.f, 0.f, 0.f, 0.f);
            }
        }
        // this lane's weights for channels (2*lane, 2*lane+1) of the block: w[tap][cd][cs]
        float w0[TAPS][4], w1[TAPS][4];
#pragma unroll
        for (int tap = 0; tap < TAPS; ++tap)
#pragma unroll
            for (int cd = 0; cd < 4; ++cd) {
                if (cd < p.Cd) {
                    const TIn* wp = wpk + ((long long)tap * p.Cd + cd) * p.Cs + cb + lane * 2;
                    w0[tap][cd] = to_f32<TIn>(wp[0]); w1[tap][cd] = to_f32<TIn>(wp[1]);
                } else { w0[tap][cd] = 0.f; w1[tap][cd] = 0.f; }
            }
        __syncthreads();
        for (int px = 0; px < TC_TW; ++px) {
            float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
                const int ky = tap / K, kx = tap - ky * K;
                const int oyy = p.flip ? (K - 1 - ky) : ky, oxx = p.flip ? (K - 1 - kx) : kx;
                const TIn* tp = tile + ((size_t)(warp + oyy) * tw + (px + oxx)) * 64 + lane * 2;
                const float v0 = to_f32<TIn>(tp[0]), v1 = to_f32<TIn>(tp[1]);
#pragma unroll
                for (int cd = 0; cd < 4; ++cd) a[cd] = fmaf(v0, w0[tap][cd], fmaf(v1, w1[tap][cd], a[cd]));
            }
#pragma unroll
            for (int cd = 0; cd < 4; ++cd) {
                float v = a[cd];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == px) res[cd] += v;
            }
        }
    }
    const int oy = y0 + warp, ox = x0 + lane;
    if (oy < p.Ho && ox < p.Wo) {
        TOut* o = dst + (((long long)n * p.Ho + oy) * p.Wo + ox) * p.Cd;
        for (int cd = 0; cd < p.Cd; ++cd) o[cd] = from_f32<TOut>(apply_act(res[cd] + (bias ? bias[cd] : 0.f), p.act, p.slope));
    }
}

// ---------------------------------------------------------------------------------------------
// weight gradients
// ---------------------------------------------------------------------------------------------
// thin input (Cin <= 4), wide output (Cout multiple of 64): dw[co][ci][tap] += sum_pix dy[pix][co] * x[pix@tap][ci]
// block = 16 output-channel quads x 16 pixel streams; a thread keeps a [4 co][28] accumulator tile in registers so
// each 16-byte shared-memory read feeds 16 FMAs.  Pixel chunks of 64 are double-buffered in shared memory: the dy
// tile of chunk k+1 arrives by cp.async and the im2col'ed x values of chunk k+1 sit in registers while chunk k is
// being multiplied, so global-memory latency is hidden with one block (8 warps) per SM.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = pred ? 16 : 0;       // src-size 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <typename TIn>
__global__ void __launch_bounds__(256)
thin_wgrad_ci_kernel(ThinParams p, const TIn* __restrict__ x, const TIn* __restrict__ dy, float* __restrict__ dw, int chunks_per_block) {
    // p.Cs = Cin (thin), p.Cd = Cout (wide); blockIdx.y = 64-wide block of output channels
    constexpr int VEC = 16 / (int)sizeof(TIn);          // dy elements per 16-byte cp.async
    __shared__ __align__(16) float xs[2][64][28];       // per pixel: taps*Cin (<= 27) source values, zero padded
    __shared__ __align__(16) TIn gs[2][64][64];         // dy of the chunk
    float (*red)[28] = xs[0];                           // reused after the main loop for the cross-stream reduction
    const int taps = p.k * p.k, nin = taps * p.Cs;
    const int s = threadIdx.x >> 4, cq = threadIdx.x & 15;
    const int co0 = blockIdx.y * 64;
    const int M = p.N * p.Ho * p.Wo;                    // < 2^31 (checked by the caller)
    // gather role: thread = (pixel gl, 7 consecutive im2col columns starting at gr0)
    const int gl = threadIdx.x >> 2, gr0 = (threadIdx.x & 3) * 7;
    int g_dy[7], g_dx[7], g_ci[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const int r = gr0 + j;
        const int tap = r / p.Cs, ci = r - tap * p.Cs;
        const int ky = tap / p.k, kx = tap - ky * p.k;
        g_dy[j] = ky - p.pad; g_dx[j] = kx - p.pad; g_ci[j] = r < nin ? ci : -1;
    }
    float acc[4][28];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < 28; ++i) acc[a][i] = 0.f;

    float xr[7];
    auto prefetch = [&](int ch, int buf) {
        const int p0 = (blockIdx.x * chunks_per_block + ch) * 64;
        const int pix = p0 + gl;
        const bool okp = pix < M;
        int ox = 0, oy = 0, base = 0;
        if (okp) {
            ox = pix % p.Wo; const int q = pix / p.Wo;
            oy = q % p.Ho; base = (q / p.Ho) * p.H * p.W;
        }
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int sy = oy + g_dy[j], sx = ox + g_dx[j];
            float v = 0.f;
            if (okp && g_ci[j] >= 0 && sy >= 0 && sy < p.H && sx >= 0 && sx < p.W)
                v = to_f32<TIn>(x[(long long)(base + sy * p.W + sx) * p.Cs + g_ci[j]]);
            xr[j] = v;
        }
        for (int i = threadIdx.x; i < 64 * (64 / VEC); i += 256) {
            const int pl = i / (64 / VEC), v0 = (i % (64 / VEC)) * VEC;
            const int px = p0 + pl;
            cp_async16(&gs[buf][pl][v0], dy + (long long)(px < M ? px : 0) * p.Cd + co0 + v0, px < M);
        }
        cp_async_commit();
    };
    const int first_p0 = blockIdx.x * chunks_per_block * 64;
    int nch = 0;
    if (first_p0 < M) nch = min(chunks_per_block, (M - first_p0 + 63) / 64);
    if (nch > 0) prefetch(0, 0);
    for (int ch = 0; ch < nch; ++ch) {
        const int buf = ch & 1;
#pragma unroll
        for (int j = 0; j < 7; ++j) xs[buf][gl][gr0 + j] = xr[j];
        cp_async_wait_all();
        __syncthreads();                                 // chunk `ch` is complete in buffer `buf`; buffer buf^1 is free
        if (ch + 1 < nch) prefetch(ch + 1, buf ^ 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pl = s * 4 + j;
            float gv[4];
            load4<TIn>(&gs[buf][pl][cq * 4], gv);
            const float4* xrow = reinterpret_cast<const float4*>(&xs[buf][pl][0]);
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const float4 v = xrow[i];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    acc[a][i * 4] = fmaf(gv[a], v.x, acc[a][i * 4]); acc[a][i * 4 + 1] = fmaf(gv[a], v.y, acc[a][i * 4 + 1]);
                    acc[a][i * 4 + 2] = fmaf(gv[a], v.z, acc[a][i * 4 + 2]); acc[a][i * 4 + 3] = fmaf(gv[a], v.w, acc[a][i * 4 + 3]);
                }
            }
        }
    }
    // combine the 16 pixel streams in shared memory, then one global atomic per output element
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 28; i += 256) (&red[0][0])[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < 28; ++i)
            if (i < nin) atomicAdd(&red[cq * 4 + a][i], acc[a][i]);
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * nin; i += 256) {
        const int c = i / nin, r = i - c * nin;
        const int tap = r / p.Cs, ci = r - tap * p.Cs;
        atomicAdd(dw + ((long long)(co0 + c) * p.Cs + ci) * taps + tap, red[c][r]);
    }
}

// wide input (Cin multiple of 64), thin output (Cout <= 4): same halo tile as thin_cout; thread = (pixel stream, ci)
template <typename TIn>
__global__ void __launch_bounds__(256)
thin_wgrad_co_kernel(ThinParams p, const TIn* __restrict__ x, const TIn* __restrict__ dy, float* __restrict__ dw) {
    // p.Cs = Cin (wide), p.Cd = Cout (thin); blockIdx.y = 64-wide block of input channels; 3x3 only
    extern __shared__ __align__(16) unsigned char thin_raw[];
    constexpr int k = 3, taps = 9;
    constexpr int th = TC_TH + k - 1, tw = TC_TW + k - 1;
    TIn* tile = reinterpret_cast<TIn*>(thin_raw);                                  // [th][tw][64]
    float* gs = reinterpret_cast<float*>(thin_raw + (size_t)th * tw * 64 * sizeof(TIn));   // [TH*TW][4] dy of the tile
    float* red = gs + TC_TH * TC_TW * 4;                                            // [4][64]
    const int tiles_x = (p.Wo + TC_TW - 1) / TC_TW, tiles_y = (p.Ho + TC_TH - 1) / TC_TH;
    const int total_tiles = tiles_x * tiles_y * p.N;
    const int cb = blockIdx.y * 64;
    const int s = threadIdx.x >> 6, c = threadIdx.x & 63;
    float acc[9][4];
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int b = blockIdx.x; b < total_tiles; b += gridDim.x) {
        int bb = b;
        const int tx = bb % tiles_x; bb /= tiles_x;
        const int ty = bb % tiles_y; const int n = bb / tiles_y;
        const int x0 = tx * TC_TW, y0 = ty * TC_TH;
        __syncthreads();
        const int vec_per_pix = 8;
        for (int i = threadIdx.x; i < th * tw * vec_per_pix; i += 256) {
            const int v = i % vec_per_pix; const int pp = i / vec_per_pix;
            const int cc = pp % tw, r = pp / tw;
            const int sy = y0 + r - p.pad, sx = x0 + cc - p.pad;
            TIn* d = tile + ((size_t)r * tw + cc) * 64 + v * 8;
            if (sy >= 0 && sy < p.H && sx >= 0 && sx < p.W) {
                const TIn* sp = x + (((long long)n * p.H + sy) * p.W + sx) * p.Cs + cb + v * 8;
#pragma unroll
                for (int j = 0; j < 8; j += 4) { float t4[4]; load4<TIn>(sp + j, t4); store4<TIn>(d + j, t4[0], t4[1], t4[2], t4[3]); }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j += 4) store4<TIn>(d + j, 0.f, 0.f, 0.f, 0.f);
            }
        }
        for (int i = threadIdx.x; i < TC_TH * TC_TW; i += 256) {
            const int r = i / TC_TW, cc = i % TC_TW;
            const int oy = y0 + r, ox = x0 + cc;
            for (int cd = 0; cd < 4; ++cd)
                gs[i * 4 + cd] = (cd < p.Cd && oy < p.Ho && ox < p.Wo) ? to_f32<TIn>(dy[(((long long)n * p.Ho + oy) * p.Wo + ox) * p.Cd + cd]) : 0.f;
        }
        __syncthreads();
        // stream s handles tile rows 2s, 2s+1
        for (int r = 2 * s; r < 2 * s + 2; ++r) {
            for (int cc = 0; cc < TC_TW; ++cc) {
                const float4 g = *reinterpret_cast<const float4*>(gs + (r * TC_TW + cc) * 4);
#pragma unroll
                for (int tap = 0; tap < taps; ++tap) {
                    const int ky = tap / k, kx = tap - ky * k;
                    const float v = to_f32<TIn>(tile[((size_t)(r + ky) * tw + (cc + kx)) * 64 + c]);
                    acc[tap][0] = fmaf(g.x, v, acc[tap][0]); acc[tap][1] = fmaf(g.y, v, acc[tap][1]);
                    acc[tap][2] = fmaf(g.z, v, acc[tap][2]); acc[tap][3] = fmaf(g.w, v, acc[tap][3]);
                }
            }
        }
    }
#pragma unroll
    for (int tap = 0; tap < taps; ++tap)
#pragma unroll
        for (int cd = 0; cd < 4; ++cd) {
            if (cd < p.Cd) {       // uniform across the block
                __syncthreads();
                red[s * 64 + c] = acc[tap][cd];
                __syncthreads();
                if (s == 0) atomicAdd(dw + ((long long)cd * p.Cs + cb + c) * taps + tap, red[c] + red[64 + c] + red[128 + c] + red[192 + c]);
            }
        }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static bool thin_geom_ok(const sr_conv_desc* d) {
    return d->stride == 1 && d->kh == d->kw && (d->kh == 1 || d->kh == 3) && (d->shuffle_r <= 1);
}

static bool thin_cin_combo(int k, int cs) { return (k == 3 && (cs == 3 || cs == 1)) || (k == 1 && cs == 3); }

bool thin_fwd_supported(const sr_conv_desc* d, bool dgrad) {
    if (!thin_geom_ok(d)) return false;
    const int Cs = dgrad ? d->Cout : d->Cin, Cd = dgrad ? d->Cin : d->Cout;
    if (Cs <= 4 && Cd % 64 == 0 && Cd <= 512 && thin_cin_combo(d->kh, Cs)) return true;
    if (Cd <= 4 && Cs % 64 == 0) return true;
    return false;
}

bool thin_wgrad_supported(const sr_conv_desc* d) {
    if (!thin_geom_ok(d)) return false;
    if (d->Cin <= 4 && d->Cout % 64 == 0 && d->kh * d->kw * d->Cin <= 27 && (long long)d->N * d->Ho * d->Wo < (1ll << 31) - 64) return true;
    if (d->Cout <= 4 && d->Cin % 64 == 0 && d->kh == 3) return true;
    return false;
}

template <typename TIn, typename TOut, int K, int CS>
static void thin_cin_launch(const ThinParams& p, const void* src, const void* w, const float* bias, void* dst, cudaStream_t st) {
    const long long G = (long long)p.N * p.Ho * cdiv(p.Wo, TCI_PX);
    const size_t smem = sizeof(float) * (K * K * CS * p.Cd + p.Cd);
    const unsigned grid = (unsigned)cdiv(G, 64);
#define SR_THIN_CIN(A) thin_cin_kernel<TIn, TOut, K, CS, A><<<grid, 256, smem, st>>>(p, (const TIn*)src, (const TIn*)w, bias, (TOut*)dst)
    switch (p.act) {
        case SR_ACT_LRELU: SR_THIN_CIN(SR_ACT_LRELU); break;
        case SR_ACT_RELU: SR_THIN_CIN(SR_ACT_RELU); break;
        case SR_ACT_SIGMOID: SR_THIN_CIN(SR_ACT_SIGMOID); break;
        default: SR_THIN_CIN(SR_ACT_NONE); break;
    }
#undef SR_THIN_CIN
}

template <typename TIn, typename TOut, int K>
static void thin_cout_launch(const ThinParams& p, const void* src, const void* w, const float* bias, void* dst, cudaStream_t st) {
    constexpr int th = TC_TH + K - 1, tw = TC_TW + K - 1;
    const size_t smem = (size_t)th * tw * 64 * sizeof(TIn);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(thin_cout_kernel<TIn, TOut, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr = true; }
    const int tiles = (int)(cdiv(p.Wo, TC_TW) * cdiv(p.Ho, TC_TH) * p.N);
    thin_cout_kernel<TIn, TOut, K><<<tiles, 256, smem, st>>>(p, (const TIn*)src, (const TIn*)w, bias, (TOut*)dst);
}

template <typename TIn, typename TOut>
static int thin_fwd_t(const ThinParams& p, const void* src, const void* w, const float* bias, void* dst, cudaStream_t st) {
    if (p.Cs <= 4) {
        if (p.k == 3 && p.Cs == 3) thin_cin_launch<TIn, TOut, 3, 3>(p, src, w, bias, dst, st);
        else if (p.k == 3 && p.Cs == 1) thin_cin_launch<TIn, TOut, 3, 1>(p, src, w, bias, dst, st);
        else thin_cin_launch<TIn, TOut, 1, 3>(p, src, w, bias, dst, st);
    } else {
        if (p.k == 3) thin_cout_launch<TIn, TOut, 3>(p, src, w, bias, dst, st);
        else thin_cout_launch<TIn, TOut, 1>(p, src, w, bias, dst, st);
    }
    count_launch();
    return check_launch("thin conv kernel");
}

int thin_fwd_run(const sr_conv_desc* d, bool dgrad, const void* src, const void* w, const float* bias, void* dst, cudaStream_t st) {
    ThinParams p;
    p.N = d->N;
    p.H = dgrad ? d->Ho : d->H; p.W = dgrad ? d->Wo : d->W;
    p.Ho = dgrad ? d->H : d->Ho; p.Wo = dgrad ? d->W : d->Wo;
    p.Cs = dgrad ? d->Cout : d->Cin; p.Cd = dgrad ? d->Cin : d->Cout;
    p.k = d->kh; p.pad = dgrad ? (d->kh - 1 - d->pad) : d->pad; p.flip = dgrad ? 1 : 0;
    p.act = dgrad ? SR_ACT_NONE : d->act; p.slope = d->slope;
    const bool in_bf = d->in_dtype == SR_BF16, out_bf = d->out_dtype == SR_BF16;
    if (in_bf && out_bf) return thin_fwd_t<__nv_bfloat16, __nv_bfloat16>(p, src, w, bias, dst, st);
    if (in_bf) return thin_fwd_t<__nv_bfloat16, float>(p, src, w, bias, dst, st);
    if (out_bf) return thin_fwd_t<float, __nv_bfloat16>(p, src, w, bias, dst, st);
    return thin_fwd_t<float, float>(p, src, w, bias, dst, st);
}

template <typename TIn>
static int thin_wgrad_t(const ThinParams& p, bool thin_ci, const void* x, const void* dy, float* dw, cudaStream_t st) {
    const long long M = (long long)p.N * p.Ho * p.Wo;
    if (thin_ci) {
        const long long chunks = cdiv(M, 64);
        const int cpb = (int)std::max<long long>(1, cdiv(chunks, std::max(1, 148 / std::max(1, p.Cd / 64))));   // one wave, 1 block / SM
        dim3 grid((unsigned)cdiv(chunks, cpb), (unsigned)(p.Cd / 64));
        thin_wgrad_ci_kernel<TIn><<<grid, 256, 0, st>>>(p, (const TIn*)x, (const TIn*)dy, dw, cpb);
    } else {
        const int th = TC_TH + p.k - 1, tw = TC_TW + p.k - 1;
        const size_t smem = (size_t)th * tw * 64 * sizeof(TIn) + sizeof(float) * (TC_TH * TC_TW * 4 + 256);
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(thin_wgrad_co_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024); attr = true; }
        const int tiles = (int)(cdiv(p.Wo, TC_TW) * cdiv(p.Ho, TC_TH) * p.N);
        dim3 grid((unsigned)std::min(tiles, 148 * 2), (unsigned)(p.Cs / 64));
        thin_wgrad_co_kernel<TIn><<<grid, 256, smem, st>>>(p, (const TIn*)x, (const TIn*)dy, dw);
    }
    count_launch();
    return check_launch("thin wgrad kernel");
}

int thin_wgrad_run(const sr_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    ThinParams p;
    p.N = d->N; p.H = d->H; p.W = d->W; p.Ho = d->Ho; p.Wo = d->Wo;
    p.Cs = d->Cin; p.Cd = d->Cout; p.k = d->kh; p.pad = d->pad; p.flip = 0; p.act = 0; p.slope = 0.f;
    const bool thin_ci = d->Cin <= 4;
    if (d->in_dtype == SR_BF16) return thin_wgrad_t<__nv_bfloat16>(p, thin_ci, x, dy, dw, st);
    return thin_wgrad_t<float>(p, thin_ci, x, dy, dw, st);
}

}  // namespace sr
