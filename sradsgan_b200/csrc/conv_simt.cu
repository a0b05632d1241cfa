// Generic SIMT implicit-GEMM convolution (forward / dgrad / wgrad), fp32 accumulate.
//
// Role in the design: this is the exact-arithmetic path — fp32 mode (<=1e-4
// parity), thin-channel layers (Cin = 3/2, Cout = 3/1: K5/K6/SLAM 7x7 of SURVEY.md §2b, all HBM bound),
// strided dgrad — and the safety net under the tcgen05 kernels in conv_tc.cu.  NHWC activations,
// weights packed [tap][Cd][Cs] so that the reduction index (tap, cs) is contiguous for both operands.
#include "common.cuh"

namespace sr {

struct IgemmParams {
    int N, Hs, Ws, Cs;   // source tensor (x for fwd, dy for dgrad)
    int Hd, Wd, Cd;      // destination pixel grid and channel count
    int kh, kw, stride, pad;
    int act;
    float slope;
    int shuffle_r;
    long long M;         // N*Hd*Wd
    int K;               // kh*kw*Cs
};

// maps destination row/col + tap to the source row/col; returns false when the tap falls outside
template <bool DGRAD>
__device__ __forceinline__ bool src_coord(const IgemmParams& p, int dy_, int dx_, int ky, int kx, int& sy, int& sx) {
    if (!DGRAD) {
        sy = dy_ * p.stride + ky - p.pad;
        sx = dx_ * p.stride + kx - p.pad;
        return sy >= 0 && sy < p.Hs && sx >= 0 && sx < p.Ws;
    } else {
        int ty = dy_ + p.pad - ky, tx = dx_ + p.pad - kx;
        if (ty < 0 || tx < 0) return false;
        if (p.stride > 1) {
            if ((ty % p.stride) || (tx % p.stride)) return false;
            ty /= p.stride; tx /= p.stride;
        }
        sy = ty; sx = tx;
        return sy < p.Hs && sx < p.Ws;
    }
}

template <typename TIn, typename TOut, bool DGRAD, int BN>
__global__ void __launch_bounds__(256)
igemm_simt_kernel(IgemmParams p, const TIn* __restrict__ src, const TIn* __restrict__ wpk,
                  const float* __restrict__ bias, const TOut* __restrict__ residual, TOut* __restrict__ dst) {
    constexpr int BM = 64, BK = 16, TN = BN / 16;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int t = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    const int l_row = t >> 2, l_k = (t & 3) * 4;
    const long long am = m0 + l_row;
    const bool a_valid = am < p.M;
    int an = 0, ay = 0, ax = 0;
    if (a_valid) {
        ax = (int)(am % p.Wd);
        long long r = am / p.Wd;
        ay = (int)(r % p.Hd);
        an = (int)(r / p.Hd);
    }
    const bool b_thread = l_row < BN;
    const int b_cd = n0 + l_row;
    const bool b_valid = b_thread && b_cd < p.Cd;
    const bool vec = (p.Cs % 4) == 0;

    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    const int ty = t >> 4, tx = t & 15;
    float ra[4], rb[4];

    auto load_tiles = [&](int k0) {
        const int k = k0 + l_k;
#pragma unroll
        for (int j = 0; j < 4; ++j) { ra[j] = 0.f; rb[j] = 0.f; }
        if (k >= p.K) return;
        if (vec) {
            const int tap = k / p.Cs, cs = k - tap * p.Cs;
            const int ky = tap / p.kw, kx = tap - ky * p.kw;
            int sy, sx;
            if (a_valid && src_coord<DGRAD>(p, ay, ax, ky, kx, sy, sx))
                load4<TIn>(src + (((long long)an * p.Hs + sy) * p.Ws + sx) * p.Cs + cs, ra);
            if (b_valid) load4<TIn>(wpk + ((long long)tap * p.Cd + b_cd) * p.Cs + cs, rb);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kk = k + j;
                if (kk >= p.K) break;
                const int tap = kk / p.Cs, cs = kk - tap * p.Cs;
                const int ky = tap / p.kw, kx = tap - ky * p.kw;
                int sy, sx;
                if (a_valid && src_coord<DGRAD>(p, ay, ax, ky, kx, sy, sx))
                    ra[j] = to_f32<TIn>(src[(((long long)an * p.Hs + sy) * p.Ws + sx) * p.Cs + cs]);
                if (b_valid) rb[j] = to_f32<TIn>(wpk[((long long)tap * p.Cd + b_cd) * p.Cs + cs]);
            }
        }
    };

    load_tiles(0);
    for (int k0 = 0; k0 < p.K; k0 += BK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[l_k + j][l_row] = ra[j];
            if (b_thread) Bs[l_k + j][l_row] = rb[j];
        }
        __syncthreads();
        if (k0 + BK < p.K) load_tiles(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float b[TN];
            if constexpr (TN == 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                b[0] = b4.x; b[1] = b4.y; b[2] = b4.z; b[3] = b4.w;
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    const int r = p.shuffle_r > 1 ? p.shuffle_r : 1;
    const int r2 = r * r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        int x = 0, y = 0, n = 0;
        if (r > 1) {
            x = (int)(m % p.Wd);
            long long q = m / p.Wd;
            y = (int)(q % p.Hd);
            n = (int)(q / p.Hd);
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int cd = n0 + tx * TN + j;
            if (cd >= p.Cd) continue;
            float v = acc[i][j];
            long long idx;
            if (r > 1) {
                // weights are packed subpixel-major (sr_pack_weights shuffle_r): column cd = sub*cq + c
                const int cq = p.Cd / r2;
                const int sub = cd / cq, c = cd - sub * cq, si = sub / r, sj = sub - si * r;
                if (bias) v += bias[c * r2 + sub];
                idx = ((((long long)n * p.Hd * r + (y * r + si)) * ((long long)p.Wd * r)) + (x * r + sj)) * cq + c;
            } else {
                if (bias) v += bias[cd];
                idx = m * p.Cd + cd;
            }
            v = apply_act(v, p.act, p.slope);
            if (residual) v += to_f32<TOut>(residual[idx]);
            dst[idx] = from_f32<TOut>(v);
        }
    }
}

struct WgradParams {
    int N, H, W, Cin, Ho, Wo, Cout, kh, kw, stride, pad;
    long long M;      // N*Ho*Wo
    int co_tiles, ci_tiles, taps;
    long long chunk;  // pixels per split
};

// dw[co][ci][ky][kx] += sum_pix dy[pix][co] * x[pix@tap][ci]; split over pixel chunks, fp32 atomics.
template <typename TIn, int BN>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(WgradParams p, const TIn* __restrict__ x, const TIn* __restrict__ dy, float* __restrict__ dw) {
    constexpr int BM = 64, BK = 16, TN = BN / 16;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int t = threadIdx.x;
    int bx = blockIdx.x;
    const int tap = bx % p.taps; bx /= p.taps;
    const int ci_t = bx % p.ci_tiles;
    const int co_t = bx / p.ci_tiles;
    const int ky = tap / p.kw, kx = tap - ky * p.kw;
    const int co0 = co_t * BM, ci0 = ci_t * BN;
    const long long p0 = (long long)blockIdx.y * p.chunk;
    const long long p1 = min(p.M, p0 + p.chunk);

    const int l_pix = t >> 4, l_c = t & 15;
    const bool a_vec = (p.Cout % 4) == 0;
    const bool b_vec = (p.Cin % 4) == 0 && BN == 64;
    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    const int ty = t >> 4, tx = t & 15;

    for (long long pk = p0; pk < p1; pk += BK) {
        const long long pix = pk + l_pix;
        float ra[4] = {0.f, 0.f, 0.f, 0.f};
        float rb[4] = {0.f, 0.f, 0.f, 0.f};
        if (pix < p1) {
            const int co = co0 + l_c * 4;
            if (a_vec) {
                if (co < p.Cout) load4<TIn>(dy + pix * p.Cout + co, ra);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (co + j < p.Cout) ra[j] = to_f32<TIn>(dy[pix * p.Cout + co + j]);
            }
            const int ox = (int)(pix % p.Wo);
            const long long q = pix / p.Wo;
            const int oy = (int)(q % p.Ho);
            const int n = (int)(q / p.Ho);
            const int iy = oy * p.stride + ky - p.pad, ix = ox * p.stride + kx - p.pad;
            if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
                const TIn* xp = x + (((long long)n * p.H + iy) * p.W + ix) * p.Cin;
                if (BN == 64) {
                    const int ci = ci0 + l_c * 4;
                    if (b_vec) {
                        if (ci < p.Cin) load4<TIn>(xp + ci, rb);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (ci + j < p.Cin) rb[j] = to_f32<TIn>(xp[ci + j]);
                    }
                } else {
                    const int ci = ci0 + l_c;
                    if (ci < p.Cin) rb[0] = to_f32<TIn>(xp[ci]);
                }
            }
        }
        *reinterpret_cast<float4*>(&As[l_pix][l_c * 4]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
        if (BN == 64) *reinterpret_cast<float4*>(&Bs[l_pix][l_c * 4]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
        else Bs[l_pix][l_c] = rb[0];
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float b[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        if (co >= p.Cout) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int ci = ci0 + tx * TN + j;
            if (ci >= p.Cin) continue;
            atomicAdd(dw + (((long long)co * p.Cin + ci) * p.kh + ky) * p.kw + kx, acc[i][j]);
        }
    }
}

template <typename TIn, typename TOut, bool DGRAD>
static int launch_igemm(const IgemmParams& p, const void* src, const void* wpk, const float* bias,
                        const void* residual, void* dst, cudaStream_t st) {
    dim3 block(256);
    if (p.Cd > 16) {
        dim3 grid((unsigned)cdiv(p.M, 64), (unsigned)cdiv(p.Cd, 64));
        igemm_simt_kernel<TIn, TOut, DGRAD, 64><<<grid, block, 0, st>>>(
            p, (const TIn*)src, (const TIn*)wpk, bias, (const TOut*)residual, (TOut*)dst);
    } else {
        dim3 grid((unsigned)cdiv(p.M, 64), (unsigned)cdiv(p.Cd, 16));
        igemm_simt_kernel<TIn, TOut, DGRAD, 16><<<grid, block, 0, st>>>(
            p, (const TIn*)src, (const TIn*)wpk, bias, (const TOut*)residual, (TOut*)dst);
    }
    count_launch();
    return check_launch("igemm_simt_kernel");
}

template <bool DGRAD>
static int dispatch_igemm(int in_dtype, int out_dtype, const IgemmParams& p, const void* src, const void* wpk,
                          const float* bias, const void* residual, void* dst, cudaStream_t st) {
    if (in_dtype == SR_F32 && out_dtype == SR_F32) return launch_igemm<float, float, DGRAD>(p, src, wpk, bias, residual, dst, st);
    if (in_dtype == SR_BF16 && out_dtype == SR_BF16) return launch_igemm<__nv_bfloat16, __nv_bfloat16, DGRAD>(p, src, wpk, bias, residual, dst, st);
    if (in_dtype == SR_BF16 && out_dtype == SR_F32) return launch_igemm<__nv_bfloat16, float, DGRAD>(p, src, wpk, bias, residual, dst, st);
    if (in_dtype == SR_F32 && out_dtype == SR_BF16) return launch_igemm<float, __nv_bfloat16, DGRAD>(p, src, wpk, bias, residual, dst, st);
    set_error("conv: unsupported dtype combination %d/%d", in_dtype, out_dtype);
    return SR_ERR_ARG;
}

int conv_fwd_simt(const sr_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual,
                  void* y, cudaStream_t st) {
    IgemmParams p;
    p.N = d->N; p.Hs = d->H; p.Ws = d->W; p.Cs = d->Cin;
    p.Hd = d->Ho; p.Wd = d->Wo; p.Cd = d->Cout;
    p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad;
    p.act = d->act; p.slope = d->slope; p.shuffle_r = d->shuffle_r;
    p.M = (long long)d->N * d->Ho * d->Wo;
    p.K = d->kh * d->kw * d->Cin;
    return dispatch_igemm<false>(d->in_dtype, d->out_dtype, p, x, w, bias, residual, y, st);
}

int conv_dgrad_simt(const sr_conv_desc* d, const void* dy, const void* wt, void* dx, cudaStream_t st) {
    IgemmParams p;
    p.N = d->N; p.Hs = d->Ho; p.Ws = d->Wo; p.Cs = d->Cout;
    p.Hd = d->H; p.Wd = d->W; p.Cd = d->Cin;
    p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad;
    p.act = SR_ACT_NONE; p.slope = 0.f; p.shuffle_r = 0;
    p.M = (long long)d->N * d->H * d->W;
    p.K = d->kh * d->kw * d->Cout;
    return dispatch_igemm<true>(d->in_dtype, d->out_dtype, p, dy, wt, nullptr, nullptr, dx, st);
}

int conv_wgrad_simt(const sr_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    WgradParams p;
    p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Ho = d->Ho; p.Wo = d->Wo; p.Cout = d->Cout;
    p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad;
    p.M = (long long)d->N * d->Ho * d->Wo;
    const int BN = d->Cin > 16 ? 64 : 16;
    p.co_tiles = (int)cdiv(d->Cout, 64);
    p.ci_tiles = (int)cdiv(d->Cin, BN);
    p.taps = d->kh * d->kw;
    const long long tiles = (long long)p.co_tiles * p.ci_tiles * p.taps;
    long long split = cdiv(148 * 6, tiles);
    const long long max_split = cdiv(p.M, 256);
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    p.chunk = cdiv(cdiv(p.M, split), 16) * 16;
    split = cdiv(p.M, p.chunk);
    dim3 grid((unsigned)tiles, (unsigned)split), block(256);
    if (d->in_dtype == SR_F32) {
        if (BN == 64) wgrad_simt_kernel<float, 64><<<grid, block, 0, st>>>(p, (const float*)x, (const float*)dy, dw);
        else wgrad_simt_kernel<float, 16><<<grid, block, 0, st>>>(p, (const float*)x, (const float*)dy, dw);
    } else {
        if (BN == 64) wgrad_simt_kernel<__nv_bfloat16, 64><<<grid, block, 0, st>>>(p, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw);
        else wgrad_simt_kernel<__nv_bfloat16, 16><<<grid, block, 0, st>>>(p, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw);
    }
    count_launch();
    return check_launch("wgrad_simt_kernel");
}

}  // namespace sr
