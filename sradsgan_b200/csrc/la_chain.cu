// Fused local-attention chain of SRADSGAN's RAB / ResGroup tails (C = 64 channels):
//
//     z = Conv1x1( SLAM( CLAM(x) ) ) + t
//       = W . ( m[n,p] * s[n,c] * x[n,p,c] ) + b + t[n,p,:]
//     s = sigmoid( MLP(avgpool_p x) + MLP(maxpool_p x) )        (CLAM, reference model/sradsgan.py:117-127)
//     m = sigmoid( conv7x7( [mean_c(s*x), max_c(s*x)] ) )       (SLAM, :141-151)
//     + 1x1 conv (:233,:262 / :297,:311) + residual (:274 / :323)
//
// The reference issues ~25 ATen kernels and ~15 full-tensor passes per chain (48 chains per generator
// forward, x3 in backward).  Here: 5 small kernels forward (x is read 3 times, z written once in fp32
// for the residual trunk and once in the compute dtype for the next 3x3 conv) and 6 backward.
// All reductions are warp-shuffle / shared-memory based with fp32 accumulation; every kernel is HBM/L2
// bound.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "la_common.cuh"

namespace sr {

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------

// per (image, pixel slice): channel sums / maxima / first arg-max.  block = 8 warps, lane = 2 channels.
template <typename T>
__global__ void __launch_bounds__(256)
la_pool_partial_kernel(const T* __restrict__ x, int P, int S, float* __restrict__ psum, float* __restrict__ pmax, int* __restrict__ pidx) {
    const int n = blockIdx.y, sl = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (P + S - 1) / S;
    const int p0 = sl * per, p1 = min(P, p0 + per);
    float s0 = 0.f, s1 = 0.f, m0 = -INFINITY, m1 = -INFINITY;
    int i0 = 0x7fffffff, i1 = 0x7fffffff;
    const T* base = x + (long long)n * P * LA_C + lane * 2;
    for (int p = p0 + warp; p < p1; p += 8) {
        const float a = to_f32<T>(base[(long long)p * LA_C]), b = to_f32<T>(base[(long long)p * LA_C + 1]);
        s0 += a; s1 += b;
        if (a > m0) { m0 = a; i0 = p; }
        if (b > m1) { m1 = b; i1 = p; }
    }
    __shared__ float sh_s[8][LA_C], sh_m[8][LA_C];
    __shared__ int sh_i[8][LA_C];
    sh_s[warp][lane * 2] = s0; sh_s[warp][lane * 2 + 1] = s1;
    sh_m[warp][lane * 2] = m0; sh_m[warp][lane * 2 + 1] = m1;
    sh_i[warp][lane * 2] = i0; sh_i[warp][lane * 2 + 1] = i1;
    __syncthreads();
    if (threadIdx.x < LA_C) {
        const int c = threadIdx.x;
        float s = 0.f, m = -INFINITY; int idx = 0x7fffffff;
        for (int w = 0; w < 8; ++w) {
            s += sh_s[w][c];
            const float mv = sh_m[w][c]; const int iv = sh_i[w][c];
            if (mv > m || (mv == m && iv < idx)) { m = mv; idx = iv; }
        }
        const long long o = ((long long)n * S + sl) * LA_C + c;
        psum[o] = s; pmax[o] = m; pidx[o] = idx;
    }
}

// per image: finish the pooling, run the bias-free MLP on both pooled vectors, gate = sigmoid(sum).
// 256 threads: four threads per channel walk the slice partials (S is up to 32), combined through shared memory.
__global__ void __launch_bounds__(256)
la_gate_fwd_kernel(const float* __restrict__ psum, const float* __restrict__ pmax, const int* __restrict__ pidx, int P, int S,
                   const float* __restrict__ fc1, const float* __restrict__ fc2, int Cr,
                   float* __restrict__ s_out, float* __restrict__ avg_out, float* __restrict__ max_out, int* __restrict__ pstar) {
    const int n = blockIdx.x, c = threadIdx.x & (LA_C - 1), part = threadIdx.x >> 6;
    __shared__ float a[LA_C], mx[LA_C], ha[16], hm[16];
    __shared__ float ps[4][LA_C], pm[4][LA_C];
    __shared__ int pi[4][LA_C];
    float s = 0.f, m = -INFINITY; int idx = 0x7fffffff;
    for (int sl = part; sl < S; sl += 4) {
        const long long o = ((long long)n * S + sl) * LA_C + c;
        s += psum[o];
        const float mv = pmax[o]; const int iv = pidx[o];
        if (mv > m || (mv == m && iv < idx)) { m = mv; idx = iv; }
    }
    ps[part][c] = s; pm[part][c] = m; pi[part][c] = idx;
    __syncthreads();
    if (part == 0) {
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            s += ps[k][c];
            const float mv = pm[k][c]; const int iv = pi[k][c];
            if (mv > m || (mv == m && iv < idx)) { m = mv; idx = iv; }
        }
        a[c] = s / (float)P; mx[c] = m;
        avg_out[n * LA_C + c] = a[c]; max_out[n * LA_C + c] = m; pstar[n * LA_C + c] = idx;
    }
    __syncthreads();
    if (threadIdx.x < Cr) {
        const int r = threadIdx.x;
        float u = 0.f, v = 0.f;
        for (int k = 0; k < LA_C; ++k) { u += fc1[r * LA_C + k] * a[k]; v += fc1[r * LA_C + k] * mx[k]; }
        ha[r] = fmaxf(u, 0.f); hm[r] = fmaxf(v, 0.f);
    }
    __syncthreads();
    if (part == 0) {
        float o = 0.f;
        for (int j = 0; j < Cr; ++j) o += fc2[c * Cr + j] * (ha[j] + hm[j]);
        s_out[n * LA_C + c] = 1.f / (1.f + __expf(-o));
    }
}

// per pixel (one warp): mean / max / first arg-max over channels of u = s*x
template <typename T>
__global__ void __launch_bounds__(256)
la_stats_kernel(const T* __restrict__ x, const float* __restrict__ s, int P, long long NP, float* __restrict__ q, unsigned char* __restrict__ cstar) {
    const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (pix >= NP) return;
    const int lane = threadIdx.x & 31;
    const int n = (int)(pix / P);
    const float u0 = to_f32<T>(x[pix * LA_C + lane * 2]) * s[n * LA_C + lane * 2];
    const float u1 = to_f32<T>(x[pix * LA_C + lane * 2 + 1]) * s[n * LA_C + lane * 2 + 1];
    float sum = u0 + u1;
    float mv = u0; int mi = lane * 2;
    if (u1 > mv) { mv = u1; mi = lane * 2 + 1; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float ov = __shfl_xor_sync(0xffffffffu, mv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (ov > mv || (ov == mv && oi < mi)) { mv = ov; mi = oi; }
    }
    if (lane == 0) {
        q[pix * 2] = sum * (1.f / LA_C);
        q[pix * 2 + 1] = mv;
        cstar[pix] = (unsigned char)mi;
    }
}

// m = sigmoid(conv7x7(q)), q = [mean, max] planes stored interleaved [N][H][W][2]
__global__ void __launch_bounds__(256)
la_conv7_fwd_kernel(const float* __restrict__ q, const float* __restrict__ w7, int N, int H, int W, float* __restrict__ m) {
    __shared__ float ws[98];
    if (threadIdx.x < 98) ws[threadIdx.x] = w7[threadIdx.x];     // [ch][ky][kx]
    __syncthreads();
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (long long)N * H * W) return;
    const int x = (int)(pix % W); const long long r = pix / W;
    const int y = (int)(r % H); const int n = (int)(r / H);
    float e = 0.f;
    for (int ky = 0; ky < 7; ++ky) {
        const int yy = y + ky - 3;
        if (yy < 0 || yy >= H) continue;
        for (int kx = 0; kx < 7; ++kx) {
            const int xx = x + kx - 3;
            if (xx < 0 || xx >= W) continue;
            const float2 v = *reinterpret_cast<const float2*>(q + (((long long)n * H + yy) * W + xx) * 2);
            e += ws[ky * 7 + kx] * v.x + ws[49 + ky * 7 + kx] * v.y;
        }
    }
    m[pix] = 1.f / (1.f + __expf(-e));
}

// m = sigmoid(conv7x7(q)) at one pixel, computed by the FOUR threads (sub = 0..3, consecutive lanes) that stage this pixel in
// the apply kernels: 12-13 taps each, two shuffles.  q is the [mean, max] map written by la_stats_kernel (L2 resident);
// w7s = the 98 filter values in shared memory.  Replaces the separate la_conv7_fwd_kernel launch of the forward chain.
__device__ __forceinline__ float la_conv7_quad(const float* __restrict__ q, const float* w7s, int n, int y, int x, int H, int W, int sub, bool valid) {
    float e = 0.f;
    if (valid) {
        for (int tap = sub; tap < 49; tap += 4) {
            const int ky = tap / 7, kx = tap - ky * 7;
            const int yy = y + ky - 3, xx = x + kx - 3;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const float2 v = *reinterpret_cast<const float2*>(q + (((long long)n * H + yy) * W + xx) * 2);
            e += w7s[tap] * v.x + w7s[49 + tap] * v.y;
        }
    }
    e += __shfl_xor_sync(0xffffffffu, e, 1);
    e += __shfl_xor_sync(0xffffffffu, e, 2);
    return 1.f / (1.f + __expf(-e));
}

// z = W.(m*s*x) + b + t   (64 pixels x 64 channels per block; 4x4 register tile per thread)
template <typename T>
__global__ void __launch_bounds__(256)
la_apply_kernel(const T* __restrict__ x, const float* __restrict__ s, const float* __restrict__ q, const float* __restrict__ w7,
                float* __restrict__ m_out, const float* __restrict__ t_res,
                const float* __restrict__ Wm, const float* __restrict__ bias, int P, int H, int Wd, long long NP,
                float* __restrict__ z32, T* __restrict__ z16) {
    __shared__ __align__(16) float ws[LA_C][LA_C + 4];   // ws[ci][co] = W[co][ci]
    __shared__ __align__(16) float vs[LA_C][LA_C + 4];   // vs[ci][pixel]
    __shared__ float w7s[98];
    const int t = threadIdx.x;
    for (int i = t; i < LA_C * LA_C; i += 256) ws[i % LA_C][i / LA_C] = Wm[i];
    if (t < 98) w7s[t] = w7[t];
    __syncthreads();
    const long long p0 = (long long)blockIdx.x * 64;
    {
        const int pl = t >> 2, cb = (t & 3) * 4;
        const long long pix = p0 + pl;
        const bool okp = pix < NP;
        const int n_ = okp ? (int)(pix / P) : 0, pp = okp ? (int)(pix - (long long)n_ * P) : 0;
        float mp;                                                                               // SLAM gate of this pixel
        if (q) { mp = la_conv7_quad(q, w7s, n_, pp / Wd, pp % Wd, H, Wd, t & 3, okp); if (okp && (t & 3) == 0) m_out[pix] = mp; }
        else mp = okp ? m_out[pix] : 0.f;                                                       // computed by la_conv7_fwd_kernel
        if (pix < NP) {
            const int n = (int)(pix / P);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int c = cb + 16 * jj;
                float v[4];
                load4<T>(x + pix * LA_C + c, v);
                const float4 sv = *reinterpret_cast<const float4*>(s + n * LA_C + c);
                vs[c][pl] = v[0] * mp * sv.x; vs[c + 1][pl] = v[1] * mp * sv.y;
                vs[c + 2][pl] = v[2] * mp * sv.z; vs[c + 3][pl] = v[3] * mp * sv.w;
            }
        } else {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int k = 0; k < 4; ++k) vs[cb + 16 * jj + k][pl] = 0.f;
        }
    }
    __syncthreads();
    const int tp = t >> 4, tc = t & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
    for (int ci = 0; ci < LA_C; ++ci) {
        const float4 a = *reinterpret_cast<const float4*>(&vs[ci][tp * 4]);
        const float4 w = *reinterpret_cast<const float4*>(&ws[ci][tc * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    const float4 bv = make_float4(bias[tc * 4], bias[tc * 4 + 1], bias[tc * 4 + 2], bias[tc * 4 + 3]);   // caller pointer: no alignment assumed
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long pix = p0 + tp * 4 + i;
        if (pix >= NP) continue;
        const float4 r = *reinterpret_cast<const float4*>(t_res + pix * LA_C + tc * 4);
        float4 o = make_float4(acc[i][0] + bv.x + r.x, acc[i][1] + bv.y + r.y, acc[i][2] + bv.z + r.z, acc[i][3] + bv.w + r.w);
        *reinterpret_cast<float4*>(z32 + pix * LA_C + tc * 4) = o;
        if (z16) store4<T>(z16 + pix * LA_C + tc * 4, o.x, o.y, o.z, o.w);
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------

// Persistent over 64-pixel tiles.  dz = gz32 + gz16 (either may be null);  dv = W^T dz;  g = m*dv;
// dm = sum_c dv*u (u = s*x);  dW += (m*dz) (x) u;  db += sum dz;  optionally writes dz (the residual gradient).
template <typename T>
__global__ void __launch_bounds__(256)
la_bwd_apply_kernel(const float* __restrict__ gz32, const T* __restrict__ gz16, const T* __restrict__ x, const float* __restrict__ s,
                    const float* __restrict__ m, const float* __restrict__ Wm, int P, long long NP, int tiles,
                    float* __restrict__ g, float* __restrict__ dm, float* __restrict__ dW, float* __restrict__ db,
                    float* __restrict__ dz_out) {
    extern __shared__ __align__(16) float la_smem[];
    float (*ws)[LA_C + 4] = reinterpret_cast<float (*)[LA_C + 4]>(la_smem);                        // ws[co][ci] = W[co][ci]
    float (*dzs)[LA_C + 4] = reinterpret_cast<float (*)[LA_C + 4]>(la_smem + LA_C * (LA_C + 4));   // dzs[co][pixel]
    float (*us)[LA_C + 4] = reinterpret_cast<float (*)[LA_C + 4]>(la_smem + 2 * LA_C * (LA_C + 4)); // us[ci][pixel] = s*x
    float* ms = la_smem + 3 * LA_C * (LA_C + 4);                                                   // m per pixel of the tile
    const int t = threadIdx.x;
    for (int i = t; i < LA_C * LA_C; i += 256) ws[i / LA_C][i % LA_C] = Wm[i];
    const int tp = t >> 4, tc = t & 15;
    float wacc[4][4];      // dW[co = tp*4+i][ci = tc + 16*j]
    float bacc = 0.f;      // db[co = t] for t < 64
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) wacc[i][j] = 0.f;

    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long p0 = (long long)tile * 64;
        __syncthreads();
        {
            const int pl = t >> 2, cb = (t & 3) * 4;
            const long long pix = p0 + pl;
            if (pix < NP) {
                const int n = (int)(pix / P);
                if ((t & 3) == 0) ms[pl] = m[pix];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int c = cb + 16 * jj;
                    float xv[4], d[4] = {0.f, 0.f, 0.f, 0.f};
                    load4<T>(x + pix * LA_C + c, xv);
                    if (gz32) { const float4 a = *reinterpret_cast<const float4*>(gz32 + pix * LA_C + c); d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; }
                    if (gz16) { float b[4]; load4<T>(gz16 + pix * LA_C + c, b); d[0] += b[0]; d[1] += b[1]; d[2] += b[2]; d[3] += b[3]; }
                    if (dz_out) *reinterpret_cast<float4*>(dz_out + pix * LA_C + c) = make_float4(d[0], d[1], d[2], d[3]);
                    const float4 sv = *reinterpret_cast<const float4*>(s + n * LA_C + c);
                    const float svv[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        dzs[c + k][pl] = d[k];
                        us[c + k][pl] = xv[k] * svv[k];
                    }
                }
            } else {
                if ((t & 3) == 0) ms[pl] = 0.f;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int k = 0; k < 4; ++k) { dzs[cb + 16 * jj + k][pl] = 0.f; us[cb + 16 * jj + k][pl] = 0.f; }
            }
        }
        __syncthreads();
        // GEMM 1: dv[pixel = tp*4+i][ci = tc*4+j] = sum_co dz[pixel][co] * W[co][ci]
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
        for (int co = 0; co < LA_C; ++co) {
            const float4 a = *reinterpret_cast<const float4*>(&dzs[co][tp * 4]);
            const float4 w = *reinterpret_cast<const float4*>(&ws[co][tc * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long pix = p0 + tp * 4 + i;
            const bool ok = pix < NP;
            const float mp = ms[tp * 4 + i];
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) part += acc[i][j] * us[tc * 4 + j][tp * 4 + i];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (ok) {
                *reinterpret_cast<float4*>(g + pix * LA_C + tc * 4) = make_float4(mp * acc[i][0], mp * acc[i][1], mp * acc[i][2], mp * acc[i][3]);
                if (tc == 0) dm[pix] = part;
            }
        }
        // GEMM 2: dW[co = tp*4+i][ci = tc+16*j] += sum_pixel (m*dz)[pixel][co] * u[pixel][ci]
#pragma unroll 4
        for (int p = 0; p < 64; p += 4) {
            const float4 mv = *reinterpret_cast<const float4*>(&ms[p]);
            float4 dzv[4], uv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dzv[i] = *reinterpret_cast<const float4*>(&dzs[tp * 4 + i][p]);
                dzv[i].x *= mv.x; dzv[i].y *= mv.y; dzv[i].z *= mv.z; dzv[i].w *= mv.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) uv[j] = *reinterpret_cast<const float4*>(&us[tc + 16 * j][p]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    wacc[i][j] += dzv[i].x * uv[j].x + dzv[i].y * uv[j].y + dzv[i].z * uv[j].z + dzv[i].w * uv[j].w;
        }
        if (t < LA_C) {
            float sacc = 0.f;
#pragma unroll 8
            for (int p = 0; p < 64; ++p) sacc += dzs[t][p];
            bacc += sacc;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(dW + (tp * 4 + i) * LA_C + tc + 16 * j, wacc[i][j]);
    if (t < LA_C) atomicAdd(db + t, bacc);
}

// ------------------------------------------------------------------------------------------------
// bf16 mode: the two 64x64x64 products of the chain on the tensor cores (warp-level mma.sync m16n8k16, fp32
// accumulation).  The fp32 SIMT kernels above spend about half of their time in the register-tiled GEMM; with
// the products on the tensor pipe both kernels are bound by their tensor traffic.  Operands are staged in shared
// memory as bf16 rows of 72 elements (144 B pitch: ldmatrix rows and fragment loads hit 32 distinct banks).
// fp32 mode keeps the SIMT kernels (<=1e-4 parity).
// ------------------------------------------------------------------------------------------------
// z = W.(m*s*x) + b + t, persistent over 64-pixel tiles.  warp = (16-pixel row tile mt, 32-channel half nh).
__global__ void __launch_bounds__(256)
la_apply_mma_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ s, const float* __restrict__ q,
                    const float* __restrict__ w7, float* __restrict__ m_out,
                    const float* __restrict__ t_res, const float* __restrict__ Wm, const float* __restrict__ bias, int P, int H, int Wd,
                    long long NP, int tiles, float* __restrict__ z32, __nv_bfloat16* __restrict__ z16) {
    __shared__ __align__(16) __nv_bfloat16 Ws[LA_C * LA_LD], Wl[LA_C * LA_LD];     // [co][ci], hi / lo
    __shared__ __align__(16) __nv_bfloat16 Vs[LA_C * LA_LD], Vl[LA_C * LA_LD];     // [pixel][ci] = m*s*x, hi / lo
    __shared__ float bias_s[LA_C], w7s[98];
    const int t = threadIdx.x;
    la_stage_w_split(Ws, Wl, Wm, t);
    if (t < LA_C) bias_s[t] = bias[t];
    if (t < 98) w7s[t] = w7[t];
    const int warp = t >> 5, lane = t & 31, mt = warp & 3, nh = warp >> 2, g = lane >> 2, tq = lane & 3;
    const int a_row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, a_col = ((lane >> 4) & 1) * 8;
    const int b_row = (lane & 7) + ((lane >> 4) & 1) * 8, b_col = ((lane >> 3) & 1) * 8;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long p0 = (long long)tile * 64;
        __syncthreads();
        {
            const int pl = t >> 2, cb = (t & 3) * 4;
            const long long pix = p0 + pl;
            const bool okp = pix < NP;
            const int n_ = okp ? (int)(pix / P) : 0, pp = okp ? (int)(pix - (long long)n_ * P) : 0;
            float mp;                                                                               // SLAM gate of this pixel
            if (q) { mp = la_conv7_quad(q, w7s, n_, pp / Wd, pp % Wd, H, Wd, t & 3, okp); if (okp && (t & 3) == 0) m_out[pix] = mp; }
            else mp = okp ? m_out[pix] : 0.f;                                                       // computed by la_conv7_fwd_kernel
            if (pix < NP) {
                const int n = (int)(pix / P);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int c = cb + 16 * jj;
                    float v[4];
                    load4<__nv_bfloat16>(x + pix * LA_C + c, v);
                    const float4 sv = *reinterpret_cast<const float4*>(s + n * LA_C + c);
                    st_split4(Vs + pl * LA_LD + c, Vl + pl * LA_LD + c, v[0] * mp * sv.x, v[1] * mp * sv.y, v[2] * mp * sv.z, v[3] * mp * sv.w);
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) st_split4(Vs + pl * LA_LD + cb + 16 * jj, Vl + pl * LA_LD + cb + 16 * jj, 0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a[4], al[4];
            ldsm_x4(a, Vs + a_row * LA_LD + ks * 16 + a_col);
            ldsm_x4(al, Vl + a_row * LA_LD + ks * 16 + a_col);
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                uint32_t b[4], bl[4];   // B[k = ci][n = co] = W[co][ci]: rows of Ws are already the "col" fragments
                ldsm_x4(b, Ws + (nh * 32 + np * 16 + b_row) * LA_LD + ks * 16 + b_col);
                ldsm_x4(bl, Wl + (nh * 32 + np * 16 + b_row) * LA_LD + ks * 16 + b_col);
                mma_bf16(acc[np * 2], a, b[0], b[1]);
                mma_bf16(acc[np * 2 + 1], a, b[2], b[3]);
                mma_bf16(acc[np * 2], al, b[0], b[1]);
                mma_bf16(acc[np * 2 + 1], al, b[2], b[3]);
                mma_bf16(acc[np * 2], a, bl[0], bl[1]);
                mma_bf16(acc[np * 2 + 1], a, bl[2], bl[3]);
            }
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const long long pix = p0 + mt * 16 + g + rr * 8;
            if (pix >= NP) continue;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int co = nh * 32 + nt * 8 + 2 * tq;
                const float2 r = *reinterpret_cast<const float2*>(t_res + pix * LA_C + co);
                const float o0 = acc[nt][rr * 2] + bias_s[co] + r.x, o1 = acc[nt][rr * 2 + 1] + bias_s[co + 1] + r.y;
                *reinterpret_cast<float2*>(z32 + pix * LA_C + co) = make_float2(o0, o1);
                if (z16) *reinterpret_cast<__nv_bfloat162*>(z16 + pix * LA_C + co) = __floats2bfloat162_rn(o0, o1);
            }
        }
    }
}

// Backward of the same tile: with e = m*dz staged once as bf16,
//   g = m*dv = W^T e (GEMM 1, stored fp32),  dm = (sum_ci g*u) / m,  dW += e^T u (GEMM 2, K = pixels, both operands
//   read transposed through ldmatrix.trans),  db += sum dz (fp32, from the loaded values),  dz_out = dz.
__global__ void __launch_bounds__(256, 3)
la_bwd_apply_mma_kernel(const float* __restrict__ gz32, const __nv_bfloat16* __restrict__ gz16, const float* __restrict__ gacc,
                        const __nv_bfloat16* __restrict__ x,
                        const float* __restrict__ s, const float* __restrict__ m, const float* __restrict__ Wm, int P, long long NP,
                        int tiles, float* __restrict__ g_out, float* __restrict__ dm, float* __restrict__ dW, float* __restrict__ db,
                        float* __restrict__ dz_out, float* __restrict__ wpart) {
    pdl_trigger();
    pdl_wait();                    // (common.cuh) the band path launches this kernel with the programmatic-serialization attribute
    extern __shared__ __align__(16) unsigned char la_mma_smem[];
    __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(la_mma_smem);      // [co][ci] hi
    __nv_bfloat16* Wl = Ws + LA_C * LA_LD;                                   //          lo
    __nv_bfloat16* Es = Wl + LA_C * LA_LD;                                   // [pixel][co] = m*dz hi (dz lives on the fp32 trunk)
    __nv_bfloat16* El = Es + LA_C * LA_LD;                                   //                     lo
    __nv_bfloat16* Us = El + LA_C * LA_LD;                                   // [pixel][ci] = s*x hi
    __nv_bfloat16* Ul = Us + LA_C * LA_LD;                                   //                   lo
    __shared__ float ms[LA_C], dm_part[2][LA_C], bsum[LA_C];
    const int t = threadIdx.x;
    la_stage_w_split(Ws, Wl, Wm, t);
    if (t < LA_C) bsum[t] = 0.f;
    const int warp = t >> 5, lane = t & 31, mt = warp & 3, nh = warp >> 2, g = lane >> 2, tq = lane & 3;
    const int r8a = (lane & 7) + ((lane >> 3) & 1) * 8, c8a = ((lane >> 4) & 1) * 8;     // matrices 1/2 = rows+8 / cols+8
    const int r8b = (lane & 7) + ((lane >> 4) & 1) * 8, c8b = ((lane >> 3) & 1) * 8;     // matrices 1/2 = cols+8 / rows+8
    float wacc[4][4];      // dW[co = mt*16 + g (+8)][ci = nh*32 + nt*8 + 2*tq (+1)]
    float bacc[16];        // db partial of channels cb + 16*jj + k over this thread's pixels
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { wacc[i][j] = 0.f; bacc[i * 4 + j] = 0.f; }

    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long p0 = (long long)tile * 64;
        __syncthreads();
        {
            const int pl = t >> 2, cb = (t & 3) * 4;
            const long long pix = p0 + pl;
            if (pix < NP) {
                const int n = (int)(pix / P);
                const float mp = m[pix];
                if ((t & 3) == 0) ms[pl] = mp;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int c = cb + 16 * jj;
                    float xv[4], d[4] = {0.f, 0.f, 0.f, 0.f};
                    load4<__nv_bfloat16>(x + pix * LA_C + c, xv);
                    if (gz32) { const float4 a = *reinterpret_cast<const float4*>(gz32 + pix * LA_C + c); d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; }
                    if (gz16) { float b[4]; load4<__nv_bfloat16>(gz16 + pix * LA_C + c, b); d[0] += b[0]; d[1] += b[1]; d[2] += b[2]; d[3] += b[3]; }
                    if (gacc) { const float4 a = *reinterpret_cast<const float4*>(gacc + pix * LA_C + c); d[0] += a.x; d[1] += a.y; d[2] += a.z; d[3] += a.w; }
                    if (dz_out) *reinterpret_cast<float4*>(dz_out + pix * LA_C + c) = make_float4(d[0], d[1], d[2], d[3]);
                    const float4 sv = *reinterpret_cast<const float4*>(s + n * LA_C + c);
                    st_split4(Es + pl * LA_LD + c, El + pl * LA_LD + c, mp * d[0], mp * d[1], mp * d[2], mp * d[3]);
                    st_split4(Us + pl * LA_LD + c, Ul + pl * LA_LD + c, xv[0] * sv.x, xv[1] * sv.y, xv[2] * sv.z, xv[3] * sv.w);
                    bacc[jj * 4] += d[0]; bacc[jj * 4 + 1] += d[1]; bacc[jj * 4 + 2] += d[2]; bacc[jj * 4 + 3] += d[3];
                }
            } else {
                if ((t & 3) == 0) ms[pl] = 0.f;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    st_split4(Es + pl * LA_LD + cb + 16 * jj, El + pl * LA_LD + cb + 16 * jj, 0.f, 0.f, 0.f, 0.f);
                    st_split4(Us + pl * LA_LD + cb + 16 * jj, Ul + pl * LA_LD + cb + 16 * jj, 0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        __syncthreads();
        // GEMM 1: g[pixel][ci] = sum_co e[pixel][co] * W[co][ci]      (M = pixels, N = ci, K = co)
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a[4], al[4];
            ldsm_x4(a, Es + (mt * 16 + r8a) * LA_LD + ks * 16 + c8a);
            ldsm_x4(al, El + (mt * 16 + r8a) * LA_LD + ks * 16 + c8a);
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                uint32_t b[4], bl[4];   // B[k = co][n = ci] = Ws[co][ci] read transposed
                ldsm_x4_t(b, Ws + (ks * 16 + r8a) * LA_LD + nh * 32 + np * 16 + c8a);
                ldsm_x4_t(bl, Wl + (ks * 16 + r8a) * LA_LD + nh * 32 + np * 16 + c8a);
                mma_bf16(acc[np * 2], a, b[0], b[1]);
                mma_bf16(acc[np * 2 + 1], a, b[2], b[3]);
                mma_bf16(acc[np * 2], al, b[0], b[1]);
                mma_bf16(acc[np * 2 + 1], al, b[2], b[3]);
                mma_bf16(acc[np * 2], a, bl[0], bl[1]);
                mma_bf16(acc[np * 2 + 1], a, bl[2], bl[3]);
            }
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int px = mt * 16 + g + rr * 8;
            const long long pix = p0 + px;
            float part = 0.f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int ci = nh * 32 + nt * 8 + 2 * tq;
                const __nv_bfloat162 u2 = *reinterpret_cast<const __nv_bfloat162*>(Us + px * LA_LD + ci);
                const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(Ul + px * LA_LD + ci);
                part += acc[nt][rr * 2] * (__low2float(u2) + __low2float(v2)) + acc[nt][rr * 2 + 1] * (__high2float(u2) + __high2float(v2));
                if (pix < NP) *reinterpret_cast<float2*>(g_out + pix * LA_C + ci) = make_float2(acc[nt][rr * 2], acc[nt][rr * 2 + 1]);
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (tq == 0) dm_part[nh][px] = part;
        }
        // GEMM 2: dW[co][ci] += sum_pixel e[pixel][co] * u[pixel][ci]      (M = co, N = ci, K = pixels)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a[4], al[4];       // A[m = co][k = pixel] = Es[pixel][co] read transposed
            ldsm_x4_t(a, Es + (ks * 16 + r8b) * LA_LD + mt * 16 + c8b);
            ldsm_x4_t(al, El + (ks * 16 + r8b) * LA_LD + mt * 16 + c8b);
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                uint32_t b[4], bl[4];   // B[k = pixel][n = ci] = Us[pixel][ci] read transposed
                ldsm_x4_t(b, Us + (ks * 16 + r8a) * LA_LD + nh * 32 + np * 16 + c8a);
                ldsm_x4_t(bl, Ul + (ks * 16 + r8a) * LA_LD + nh * 32 + np * 16 + c8a);
                mma_bf16(wacc[np * 2], a, b[0], b[1]);
                mma_bf16(wacc[np * 2 + 1], a, b[2], b[3]);
                mma_bf16(wacc[np * 2], al, b[0], b[1]);
                mma_bf16(wacc[np * 2 + 1], al, b[2], b[3]);
                mma_bf16(wacc[np * 2], a, bl[0], bl[1]);
                mma_bf16(wacc[np * 2 + 1], a, bl[2], bl[3]);
            }
        }
        __syncthreads();
        if (t < LA_C) {
            const long long pix = p0 + t;
            const float mp = ms[t];
            if (pix < NP) dm[pix] = mp > 0.f ? (dm_part[0][t] + dm_part[1][t]) / mp : 0.f;     // m = 0: de = dm*m*(1-m) = 0 anyway
        }
    }
    // per-block results: either added to dW / db with fp32 atomics, or (band path) written as ONE partial row per block that
    // la_bwd_band_kernel sums in a fixed order — 4160 plain stores instead of 4160 contended atomics per block
    float* wrow = wpart ? wpart + (long long)blockIdx.x * (LA_C * LA_C + LA_C) : nullptr;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int e = (mt * 16 + g + rr * 8) * LA_C + nh * 32 + nt * 8 + 2 * tq;
            if (wrow) *reinterpret_cast<float2*>(wrow + e) = make_float2(wacc[nt][rr * 2], wacc[nt][rr * 2 + 1]);
            else { atomicAdd(dW + e, wacc[nt][rr * 2]); atomicAdd(dW + e + 1, wacc[nt][rr * 2 + 1]); }
        }
    {
        const int cb = (t & 3) * 4;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int k = 0; k < 4; ++k) atomicAdd(&bsum[cb + 16 * jj + k], bacc[jj * 4 + k]);
    }
    __syncthreads();
    if (t < LA_C) {
        if (wrow) wrow[LA_C * LA_C + t] = bsum[t];
        else atomicAdd(db + t, bsum[t]);
    }
}

static int la_mma_enabled() {
    return option("SR_LA_MMA", 1);
}

// dq[p][ch] = sum_taps w7[ch][tap] * de[p - off],  de = dm * m * (1 - m)
__global__ void __launch_bounds__(256)
la_conv7_dgrad_kernel(const float* __restrict__ dm, const float* __restrict__ m, const float* __restrict__ w7, int N, int H, int W,
                      float* __restrict__ dq) {
    __shared__ float ws[98];
    if (threadIdx.x < 98) ws[threadIdx.x] = w7[threadIdx.x];
    __syncthreads();
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (long long)N * H * W) return;
    const int x = (int)(pix % W); const long long r = pix / W;
    const int y = (int)(r % H); const int n = (int)(r / H);
    float a = 0.f, b = 0.f;
    for (int ky = 0; ky < 7; ++ky) {
        const int yy = y - (ky - 3);
        if (yy < 0 || yy >= H) continue;
        for (int kx = 0; kx < 7; ++kx) {
            const int xx = x - (kx - 3);
            if (xx < 0 || xx >= W) continue;
            const long long o = ((long long)n * H + yy) * W + xx;
            const float mv = m[o];
            const float de = dm[o] * mv * (1.f - mv);
            a += ws[ky * 7 + kx] * de;
            b += ws[49 + ky * 7 + kx] * de;
        }
    }
    dq[pix * 2] = a; dq[pix * 2 + 1] = b;
}

// dw7[ch][ky][kx] += sum_p de[p] * q[p + off][ch],  de = dm * m * (1 - m).
// block = one band of LA_WG_ROWS image rows: the band's de values and the q rows it touches (+-3 halo) are staged in
// shared memory once, thread (f, h) accumulates filter element f over half h of the band, one atomic per element.
constexpr int LA_WG_ROWS = 8;

__global__ void __launch_bounds__(256)
la_conv7_wgrad_kernel(const float* __restrict__ dm, const float* __restrict__ m, const float* __restrict__ q, int N, int H, int W,
                      float* __restrict__ dw7) {
    extern __shared__ __align__(16) float la_wg_smem[];
    const int Wp = W + 6;
    float* qs = la_wg_smem;                                  // [(ROWS+6)][Wp][2], zero outside the image
    float* des = qs + (LA_WG_ROWS + 6) * Wp * 2;             // [ROWS][W]
    float* red = des + LA_WG_ROWS * W;                       // [2][98]
    const int bands = (H + LA_WG_ROWS - 1) / LA_WG_ROWS;
    const int n = blockIdx.x / bands, y0 = (blockIdx.x % bands) * LA_WG_ROWS;
    const int rows = min(LA_WG_ROWS, H - y0);
    for (int i = threadIdx.x; i < (LA_WG_ROWS + 6) * Wp; i += 256) {
        const int r = i / Wp, c = i - r * Wp;
        const int yy = y0 + r - 3, xx = c - 3;
        float2 v = make_float2(0.f, 0.f);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = *reinterpret_cast<const float2*>(q + (((long long)n * H + yy) * W + xx) * 2);
        qs[i * 2] = v.x; qs[i * 2 + 1] = v.y;
    }
    for (int i = threadIdx.x; i < LA_WG_ROWS * W; i += 256) {
        const int r = i / W, c = i - r * W;
        float v = 0.f;
        if (r < rows) {
            const long long pix = ((long long)n * H + y0 + r) * W + c;
            const float mv = m[pix];
            v = dm[pix] * mv * (1.f - mv);
        }
        des[i] = v;
    }
    __syncthreads();
    const int f = threadIdx.x % 98, h = threadIdx.x / 98;    // threads 196..255 idle in the main loop
    if (h < 2) {
        const int ch = f / 49, ky = (f % 49) / 7, kx = f % 7;
        const int r0 = h * (LA_WG_ROWS / 2), r1 = r0 + LA_WG_ROWS / 2;
        float acc = 0.f;
        for (int r = r0; r < r1; ++r) {
            const float* qrow = qs + ((r + ky) * Wp + kx) * 2 + ch;
            const float* drow = des + r * W;
            for (int c = 0; c < W; ++c) acc = fmaf(drow[c], qrow[c * 2], acc);
        }
        red[h * 98 + f] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 98) atomicAdd(dw7 + threadIdx.x, red[threadIdx.x] + red[98 + threadIdx.x]);
}

__global__ void __launch_bounds__(LA_C)
la_gate_bwd_kernel(const float* __restrict__ ds, const float* __restrict__ s, const float* __restrict__ avg, const float* __restrict__ mx,
                   const float* __restrict__ fc1, const float* __restrict__ fc2, int Cr,
                   float* __restrict__ d_fc1, float* __restrict__ d_fc2, float* __restrict__ da, float* __restrict__ dmx) {
    la_gate_bwd_body(blockIdx.x, ds, s, avg, mx, fc1, fc2, Cr, d_fc1, d_fc2, da, dmx);
}

// per (image, slice): du = g + dq_avg/C + dq_max*[c==c*];  dx_pre = s*du;  ds[n][c] += sum_p du*x
template <typename T>
__global__ void __launch_bounds__(256)
la_bwd_stats_kernel(const float* __restrict__ g, const float* __restrict__ dq, const unsigned char* __restrict__ cstar,
                    const T* __restrict__ x, const float* __restrict__ s, int P, int S, T* __restrict__ dx, float* __restrict__ ds,
                    int* __restrict__ done, const float* __restrict__ avg, const float* __restrict__ mx, const float* __restrict__ fc1,
                    const float* __restrict__ fc2, int Cr, float* __restrict__ d_fc1, float* __restrict__ d_fc2, float* __restrict__ da,
                    float* __restrict__ dmx) {
    const int n = blockIdx.y, sl = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (P + S - 1) / S;
    const int p0 = sl * per, p1 = min(P, p0 + per);
    const float s0 = s[n * LA_C + lane * 2], s1 = s[n * LA_C + lane * 2 + 1];
    float a0 = 0.f, a1 = 0.f;
    for (int p = p0 + warp; p < p1; p += 8) {
        const long long pix = (long long)n * P + p;
        const float2 gv = *reinterpret_cast<const float2*>(g + pix * LA_C + lane * 2);
        const float2 dqv = *reinterpret_cast<const float2*>(dq + pix * 2);
        const int cs = cstar[pix];
        const float du0 = gv.x + dqv.x * (1.f / LA_C) + (cs == lane * 2 ? dqv.y : 0.f);
        const float du1 = gv.y + dqv.x * (1.f / LA_C) + (cs == lane * 2 + 1 ? dqv.y : 0.f);
        const float x0 = to_f32<T>(x[pix * LA_C + lane * 2]), x1 = to_f32<T>(x[pix * LA_C + lane * 2 + 1]);
        a0 += du0 * x0; a1 += du1 * x1;
        dx[pix * LA_C + lane * 2] = from_f32<T>(s0 * du0);
        dx[pix * LA_C + lane * 2 + 1] = from_f32<T>(s1 * du1);
    }
    __shared__ float sh[8][LA_C];
    sh[warp][lane * 2] = a0; sh[warp][lane * 2 + 1] = a1;
    __syncthreads();
    if (threadIdx.x < LA_C) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += sh[w][threadIdx.x];
        atomicAdd(ds + n * LA_C + threadIdx.x, v);
        __threadfence();
    }
    // The LAST slice block of an image to get here finishes the image: gate backward (formerly its own launch between this
    // kernel and la_fix_kernel).  `done` is zeroed together with ds by the caller's memset.
    if (done == nullptr) return;              // SR_LA_GATE_TAIL=0: the gate backward runs as its own kernel
    __shared__ int is_last;
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(done + n, 1) == S - 1) ? 1 : 0;
    __syncthreads();
    if (is_last) {
        __threadfence();
        la_gate_bwd_body(n, ds, s, avg, mx, fc1, fc2, Cr, d_fc1, d_fc2, da, dmx);
    }
}
// (Computing dq inside this kernel — lanes splitting the 49 taps per pixel — was measured: chain backward 79 -> 96 us; the
// separate la_conv7_dgrad_kernel stays.)

// dx += da/P  (+ dmx at the arg-max pixel)
template <typename T>
__global__ void __launch_bounds__(256)
la_fix_kernel(T* __restrict__ dx, const float* __restrict__ da, const float* __restrict__ dmx, const int* __restrict__ pstar, int P, long long total) {
    pdl_trigger();
    pdl_wait();
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= total) return;
    const int c = (int)(i % LA_C);
    const long long pix = i / LA_C;
    const int n = (int)(pix / P), p = (int)(pix % P);
    const float invP = 1.f / (float)P;
    float v[4];
    load4<T>(dx + i, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] += da[n * LA_C + c + k] * invP;
        if (pstar[n * LA_C + c + k] == p) v[k] += dmx[n * LA_C + c + k];
        dx[i + k] = from_f32<T>(v[k]);
    }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
// pixel slices per image of the two per-image reductions (pool partials, ds): their kernels walk a slice with one warp per
// pixel and dependent global loads, so they want MANY blocks (measured, SR_LA_SLICE_PX = pixels per slice)
static int la_slices(int P) {
    int px = option("SR_LA_SLICE_PX", 96);
    if (px < 8) px = 8;
    int s = (P + px - 1) / px;
    return s < 1 ? 1 : (s > 32 ? 32 : s);
}

size_t la_workspace_bytes(int N, int H, int W) {
    const int P = H * W, S = la_slices(P);
    // fwd: psum, pmax, pidx ; bwd: g, dm, dq, ds, da, dmx
    size_t fwd = (size_t)N * (S > 32 ? S : 32) * LA_C * 12 + (size_t)N * 16 + 64;   // + the fused kernel's partials / barrier counters
    size_t bwd = (size_t)N * P * LA_C * 4 + ((size_t)N * P + 4) * 4 + ((size_t)N * P + 4) * 8 + (size_t)N * LA_C * 12 + ((size_t)N + 4) * 4;
    // band path: g, dm, <= 296 rows of dW | db partials, per-band ds partials, ds / da / dmx
    const size_t band = (size_t)N * P * LA_C * 4 + ((size_t)N * P + 4) * 4 + (size_t)444 * (LA_C * LA_C + LA_C) * 4 +
                        (size_t)N * (size_t)H * (LA_C + 100) * 4 + (size_t)N * LA_C * 12 + 64;
    if (band > bwd) bwd = band;
    return (fwd > bwd ? fwd : bwd) + 256;
}

template <typename T>
static int la_fwd_t(const void* x, const float* t_res, const float* fc1, const float* fc2, const float* w7, const float* Wm,
                    const float* bias, int N, int H, int W, int Cr, float* z32, void* z16, float* s_out, float* m_out,
                    float* avg_out, float* max_out, int* pstar, float* q, unsigned char* cstar, float* ws, cudaStream_t st) {
    const int P = H * W, S = la_slices(P);
    const long long NP = (long long)N * P;
    float* psum = ws; float* pmax = psum + (size_t)N * S * LA_C; int* pidx = reinterpret_cast<int*>(pmax + (size_t)N * S * LA_C);
    la_pool_partial_kernel<T><<<dim3(S, N), 256, 0, st>>>((const T*)x, P, S, psum, pmax, pidx);
    la_gate_fwd_kernel<<<N, 256, 0, st>>>(psum, pmax, pidx, P, S, fc1, fc2, Cr, s_out, avg_out, max_out, pstar);
    la_stats_kernel<T><<<(unsigned)cdiv(NP, 8), 256, 0, st>>>((const T*)x, s_out, P, NP, q, cstar);
    // The SLAM gate m = sigmoid(conv7x7(q)) is computed (and stored for the backward) by the apply kernel itself when the
    // chain is latency bound (training maps: a launch saved, 38.9 -> 37.5 us); on large inference batches the apply kernel
    // walks many tiles per block and the inline gate would sit on every tile's critical path (x9 tiled inference 394 -> 367
    // Mpix/s), so there the separate kernel runs first and the apply kernel reads m.
    const int tiles = (int)cdiv(NP, 64);
    const bool fuse_gate = tiles <= 4 * 296;
    if (!fuse_gate) la_conv7_fwd_kernel<<<(unsigned)cdiv(NP, 256), 256, 0, st>>>(q, w7, N, H, W, m_out);
    const float* qg = fuse_gate ? q : nullptr;
    if (sizeof(T) == 2 && la_mma_enabled()) {
        la_apply_mma_kernel<<<tiles < 296 ? tiles : 296, 256, 0, st>>>((const __nv_bfloat16*)x, s_out, qg, w7, m_out, t_res, Wm, bias, P, H, W, NP,
                                                                    tiles, z32, (__nv_bfloat16*)z16);
    } else {
        la_apply_kernel<T><<<(unsigned)cdiv(NP, 64), 256, 0, st>>>((const T*)x, s_out, qg, w7, m_out, t_res, Wm, bias, P, H, W, NP, z32, (T*)z16);
    }
    count_launch(fuse_gate ? 4 : 5);
    return check_launch("la_chain_fwd");
}

template <typename T>
static int la_bwd_t(const float* gz32, const void* gz16, const void* x, const float* s, const float* m, const float* avg,
                    const float* mx, const int* pstar, const float* q, const unsigned char* cstar, const float* fc1,
                    const float* fc2, const float* w7, const float* Wm, int N, int H, int W, int Cr, void* dx, float* d_fc1,
                    float* d_fc2, float* d_w7, float* dW, float* db, float* dz_out, float* ws, cudaStream_t st) {
    const int P = H * W, S = la_slices(P);
    const long long NP = (long long)N * P;
    const size_t npa = ((size_t)NP + 3) & ~(size_t)3;           // keep every sub-buffer 16-byte aligned
    float* g = ws; float* dm = g + (size_t)NP * LA_C; float* dq = dm + npa; float* ds = dq + npa * 2;
    const size_t n4 = ((size_t)N + 3) & ~(size_t)3;
    int* done = reinterpret_cast<int*>(ds + (size_t)N * LA_C);      // per-image block counters, zeroed with ds
    float* da = ds + (size_t)N * LA_C + n4; float* dmx = da + (size_t)N * LA_C;
    cudaMemsetAsync(ds, 0, sizeof(float) * ((size_t)N * LA_C + n4), st);
    const int tiles = (int)cdiv(NP, 64);
    const int grid = tiles < 296 ? tiles : 296;
    const size_t smem = sizeof(float) * (3 * LA_C * (LA_C + 4) + 64);
    static bool attr[2] = {false, false};
    const int ai = sizeof(T) == 2 ? 1 : 0;
    if (!attr[ai]) { cudaFuncSetAttribute(la_bwd_apply_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr[ai] = true; }
    if (sizeof(T) == 2 && la_mma_enabled()) {
        const size_t mma_smem = (size_t)6 * LA_C * LA_LD * sizeof(__nv_bfloat16);
        static bool mma_attr = false;
        if (!mma_attr) { cudaFuncSetAttribute(la_bwd_apply_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma_smem); mma_attr = true; }
        la_bwd_apply_mma_kernel<<<grid, 256, mma_smem, st>>>(gz32, (const __nv_bfloat16*)gz16, nullptr, (const __nv_bfloat16*)x, s, m, Wm, P, NP, tiles, g, dm, dW,
                                                     db, dz_out, nullptr);
    } else
        la_bwd_apply_kernel<T><<<grid, 256, smem, st>>>(gz32, (const T*)gz16, (const T*)x, s, m, Wm, P, NP, tiles, g, dm, dW, db, dz_out);
    // The 7x7 weight gradient only feeds the optimiser: it runs on the caller's auxiliary stream (forked / joined with events, so
    // the call stays stream-ordered for the caller and capturable) next to the rest of the chain.
    const AuxStreams& aux = aux_streams();                 // the caller's auxiliary stream 0 (sr_set_aux_streams), if any
    const bool side_on = aux.n >= 1 && option("SR_LA_SIDE", 1);
    cudaStream_t side = side_on ? aux.stream[0] : nullptr;
    cudaEvent_t ev_fork = aux.fork, ev_join = aux.join[0];
    {
        const int bands = (H + LA_WG_ROWS - 1) / LA_WG_ROWS;
        const size_t wg_smem = sizeof(float) * ((size_t)(LA_WG_ROWS + 6) * (W + 6) * 2 + (size_t)LA_WG_ROWS * W + 2 * 98);
        static bool wg_attr = false;
        if (!wg_attr) { cudaFuncSetAttribute(la_conv7_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); wg_attr = true; }
        cudaStream_t ws_st = st;
        if (side_on) { cudaEventRecord(ev_fork, st); cudaStreamWaitEvent(side, ev_fork, 0); ws_st = side; }
        la_conv7_wgrad_kernel<<<N * bands, 256, wg_smem, ws_st>>>(dm, m, q, N, H, W, d_w7);
        if (side_on) cudaEventRecord(ev_join, side);
    }
    la_conv7_dgrad_kernel<<<(unsigned)cdiv(NP, 256), 256, 0, st>>>(dm, m, w7, N, H, W, dq);
    const int gate_tail = option("SR_LA_GATE_TAIL", 1);
    la_bwd_stats_kernel<T><<<dim3(S, N), 256, 0, st>>>(g, dq, cstar, (const T*)x, s, P, S, (T*)dx, ds, gate_tail ? done : nullptr, avg, mx, fc1, fc2,
                                                       Cr, d_fc1, d_fc2, da, dmx);
    if (!gate_tail) la_gate_bwd_kernel<<<N, LA_C, 0, st>>>(ds, s, avg, mx, fc1, fc2, Cr, d_fc1, d_fc2, da, dmx);
    la_fix_kernel<T><<<(unsigned)cdiv(NP * LA_C / 4, 256), 256, 0, st>>>((T*)dx, da, dmx, pstar, P, NP * LA_C);
    if (side_on) cudaStreamWaitEvent(st, ev_join, 0);
    count_launch(5);
    return check_launch("la_chain_bwd");
}

int la_chain_fwd(const void* x, int dtype, const float* t_res, const float* fc1, const float* fc2, const float* w7, const float* Wm,
                 const float* bias, int N, int H, int W, int Cr, float* z32, void* z16, float* s_out, float* m_out, float* avg_out,
                 float* max_out, int* pstar, float* q, unsigned char* cstar, float* ws, cudaStream_t st) {
    if (dtype == SR_F32) return la_fwd_t<float>(x, t_res, fc1, fc2, w7, Wm, bias, N, H, W, Cr, z32, z16, s_out, m_out, avg_out, max_out, pstar, q, cstar, ws, st);
    return la_fwd_t<__nv_bfloat16>(x, t_res, fc1, fc2, w7, Wm, bias, N, H, W, Cr, z32, z16, s_out, m_out, avg_out, max_out, pstar, q, cstar, ws, st);
}

int la_chain_bwd(const float* gz32, const void* gz16, const void* x, int dtype, const float* s, const float* m, const float* avg,
                 const float* mx, const int* pstar, const float* q, const unsigned char* cstar, const float* fc1, const float* fc2,
                 const float* w7, const float* Wm, int N, int H, int W, int Cr, void* dx, float* d_fc1, float* d_fc2, float* d_w7,
                 float* dW, float* db, float* dz_out, float* ws, cudaStream_t st) {
    if (dtype == SR_F32) return la_bwd_t<float>(gz32, gz16, x, s, m, avg, mx, pstar, q, cstar, fc1, fc2, w7, Wm, N, H, W, Cr, dx, d_fc1, d_fc2, d_w7, dW, db, dz_out, ws, st);
    return la_bwd_t<__nv_bfloat16>(gz32, gz16, x, s, m, avg, mx, pstar, q, cstar, fc1, fc2, w7, Wm, N, H, W, Cr, dx, d_fc1, d_fc2, d_w7, dW, db, dz_out, ws, st);
}


// ------------------------------------------------------------------------------------------------
// struct-based entry points: band path (la_band.cu) when it applies, else the tile kernels above
// ------------------------------------------------------------------------------------------------
bool la_band_supported(int N, int H, int W);
int la_band_count(int N, int H, int W);
int la_pool_pack(const void* x, int N, int P, int S, float* psum, unsigned int* pkey, cudaStream_t st);
int la_band_fwd(LaBandFwd p, cudaStream_t st);
int la_band_bwd(LaBandBwd p, cudaStream_t st);

static int la_band_enabled() {
    return option("SR_LA_BAND", 1);
}

bool la_chain_band_path(int N, int H, int W, int dtype) { return dtype == SR_BF16 && la_band_enabled() && la_band_supported(N, H, W); }

int la_chain_forward(const sr_la_chain_args* a, cudaStream_t st) {
    const bool band = la_chain_band_path(a->N, a->H, a->W, a->x_dtype) && a->z16;
    if (!band) {
        if (a->acc_out || a->out_pool_sum) { set_error("la_chain_forward: accumulator / pooling outputs need the band path (bf16, H*W <= 65535)"); return SR_ERR_UNSUPPORTED; }
        return la_chain_fwd(a->x, a->x_dtype, a->t, a->fc1, a->fc2, a->w7, a->Wm, a->bias, a->N, a->H, a->W, a->Cr, a->z32, a->z16, a->s, a->m,
                            a->avg, a->max, a->pstar, a->q, a->cstar, (float*)a->workspace, st);
    }
    LaBandFwd p;
    memset(&p, 0, sizeof(p));
    p.x = (const __nv_bfloat16*)a->x; p.t = a->t; p.acc_in = a->acc_in; p.acc_out = a->acc_out;
    p.fc1 = a->fc1; p.fc2 = a->fc2; p.w7 = a->w7; p.Wm = a->Wm; p.bias = a->bias;
    p.N = a->N; p.H = a->H; p.W = a->W; p.Cr = a->Cr;
    p.z32 = a->z32; p.z16 = (__nv_bfloat16*)a->z16;
    p.s_out = a->s; p.m_out = a->m; p.avg_out = a->avg; p.max_out = a->max; p.pstar = a->pstar; p.q = a->q; p.cstar = a->cstar;
    p.out_psum = a->out_pool_sum; p.out_pkey = a->out_pool_key;
    if (a->pool_sum && a->pool_key && a->pool_rows > 0) {
        p.psum = a->pool_sum; p.pkey = a->pool_key; p.T = a->pool_rows;
    } else {                                   // no producer-side partials: one pooling kernel first
        const int P = a->H * a->W, S = la_slices(P);
        float* psum = (float*)a->workspace;
        unsigned int* pkey = reinterpret_cast<unsigned int*>(psum + (size_t)a->N * S * LA_C);
        int rc = la_pool_pack(a->x, a->N, P, S, psum, pkey, st);
        if (rc) return rc;
        p.psum = psum; p.pkey = pkey; p.T = S;
    }
    return la_band_fwd(p, st);
}

int la_chain_backward(const sr_la_chain_grad_args* a, cudaStream_t st) {
    const bool band = la_chain_band_path(a->N, a->H, a->W, a->x_dtype) && a->tickets;
    if (!band) {
        if (a->gacc) { set_error("la_chain_backward: the accumulator gradient needs the band path"); return SR_ERR_UNSUPPORTED; }
        return la_chain_bwd(a->gz32, a->gz16, a->x, a->x_dtype, a->s, a->m, a->avg, a->max, a->pstar, a->q, a->cstar, a->fc1, a->fc2, a->w7,
                            a->Wm, a->N, a->H, a->W, a->Cr, a->dx, a->d_fc1, a->d_fc2, a->d_w7, a->dW, a->db, a->dz_out, (float*)a->workspace, st);
    }
    const int N = a->N, P = a->H * a->W;
    const long long NP = (long long)N * P;
    const size_t npa = ((size_t)NP + 3) & ~(size_t)3;
    const int bands = la_band_count(N, a->H, a->W);
    float* ws = (float*)a->workspace;
    float* g = ws; float* dm = g + (size_t)NP * LA_C; float* wpart = dm + npa;
    const int tiles = (int)cdiv(NP, 64);
    const int grid = tiles < 444 ? tiles : 444;            // 3 resident blocks per SM (55 KB of shared memory, <= 85 registers x 256 threads)
    float* dspart = wpart + (size_t)grid * (LA_C * LA_C + LA_C);
    float* w7part = dspart + (size_t)N * bands * LA_C;
    float* ds = w7part + (size_t)N * bands * 100; float* da = ds + (size_t)N * LA_C; float* dmx = da + (size_t)N * LA_C;
    const size_t mma_smem = (size_t)6 * LA_C * LA_LD * sizeof(__nv_bfloat16);
    static bool mma_attr = false;
    if (!mma_attr) { cudaFuncSetAttribute(la_bwd_apply_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma_smem); mma_attr = true; }
    const bool pdl = option("SR_PDL", 0) != 0;
    launch_pdl(la_bwd_apply_mma_kernel, dim3(grid), dim3(256), mma_smem, st, pdl, a->gz32, (const __nv_bfloat16*)a->gz16, a->gacc,
               (const __nv_bfloat16*)a->x, a->s, a->m, a->Wm, P, NP, tiles, g, dm, a->dW, a->db, a->dz_out, wpart);
    count_launch();
    LaBandBwd p;
    memset(&p, 0, sizeof(p));
    p.g = g; p.dm = dm; p.m = a->m; p.q = a->q; p.cstar = a->cstar; p.x = (const __nv_bfloat16*)a->x; p.s = a->s; p.avg = a->avg; p.mx = a->max;
    p.fc1 = a->fc1; p.fc2 = a->fc2; p.w7 = a->w7; p.wpart = wpart; p.nparts = grid;
    p.N = N; p.H = a->H; p.W = a->W; p.Cr = a->Cr;
    p.dx = (__nv_bfloat16*)a->dx; p.d_w7 = a->d_w7; p.dW = a->dW; p.db = a->db; p.d_fc1 = a->d_fc1; p.d_fc2 = a->d_fc2;
    p.dspart = dspart; p.w7part = w7part; p.ds = ds; p.da = da; p.dmx = dmx; p.tickets = a->tickets;
    int rc = la_band_bwd(p, st);
    if (rc) return rc;
    launch_pdl(la_fix_kernel<__nv_bfloat16>, dim3((unsigned)cdiv(NP * LA_C / 4, 256)), dim3(256), 0, st, pdl, (__nv_bfloat16*)a->dx, (const float*)da,
               (const float*)dmx, (const int*)a->pstar, P, NP * LA_C);
    count_launch();
    return check_launch("la_chain_backward");
}

}  // namespace sr
