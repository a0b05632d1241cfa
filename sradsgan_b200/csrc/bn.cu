// Train-mode BatchNorm2d fused with LeakyReLU for the discriminator's first-order passes
// (reference model/sradsgan.py:476-479: Conv2d -> BatchNorm2d -> LeakyReLU(0.2), D always in train mode).
//
//   forward : one reduction pass (shifted sums -> mean / biased var, fp32), a finalize kernel that also
//             updates running_mean / running_var (momentum, unbiased var, like torch), one apply pass
//             y = lrelu(gamma * (x - mean) * rstd + beta)          -> 2 reads + 1 write of x
//   backward: one reduction pass (dbeta, dgamma), one apply pass
//             dx = gamma*rstd * (g' - dbeta/M - xhat*dgamma/M),  g' = gy * lrelu'(.)
// x is NHWC viewed as [M = N*H*W rows][C]; every kernel is HBM bound.
#include <algorithm>

#include "common.cuh"

namespace sr {

// block = 32 channels x 8 row-lanes; shifted by the first row (pivot) to avoid E[x^2]-E[x]^2 cancellation
template <typename T>
__global__ void __launch_bounds__(256)
bn_stats_kernel(const T* __restrict__ x, long long rows, int C, float* __restrict__ sum, float* __restrict__ sq) {
    __shared__ float s1[8][33], s2[8][33];
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + lane;
    float a = 0.f, b = 0.f;
    if (c < C) {
        const float pivot = to_f32<T>(x[c]);
        for (long long r = (long long)blockIdx.x * 8 + wy; r < rows; r += (long long)gridDim.x * 8) {
            const float v = to_f32<T>(x[r * C + c]) - pivot;
            a += v; b += v * v;
        }
    }
    s1[wy][lane] = a; s2[wy][lane] = b;
    __syncthreads();
    if (wy == 0 && c < C) {
#pragma unroll
        for (int i = 1; i < 8; ++i) { a += s1[i][lane]; b += s2[i][lane]; }
        atomicAdd(sum + c, a);
        atomicAdd(sq + c, b);
    }
}

template <typename T>
__global__ void bn_finalize_kernel(const T* __restrict__ x, const float* __restrict__ sum, const float* __restrict__ sq, long long rows,
                                   int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ save_mean,
                                   float* __restrict__ save_rstd, float* __restrict__ scale, float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float pivot = to_f32<T>(x[c]);
    const float inv = 1.f / (float)rows;
    const float ms = sum[c] * inv;
    const float mean = pivot + ms;
    float var = sq[c] * inv - ms * ms;
    var = fmaxf(var, 0.f);
    const float rstd = rsqrtf(var + eps);
    save_mean[c] = mean; save_rstd[c] = rstd;
    const float sc = gamma[c] * rstd;
    scale[c] = sc; shift[c] = beta[c] - mean * sc;
    if (running_mean) {
        const float unbiased = rows > 1 ? var * ((float)rows / (float)(rows - 1)) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
}

// The elementwise BatchNorm kernels walk the tensor with a grid stride that is a multiple of C (1024 elements per block,
// C | 1024 or C | grid stride — checked by the launcher, HOIST = true), so a thread meets the SAME four channels in every
// iteration: their per-channel coefficients are loaded once into registers instead of 2..11 global loads per element
// (bn_bwd_apply on the 108^2 x 64 layer: 74 us for 72 MB of traffic, LSU bound).  Same formulas.
template <typename T, bool HOIST>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const T* __restrict__ x, long long total, int C, const float* __restrict__ scale, const float* __restrict__ shift,
                float slope, T* __restrict__ y) {
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    float sc[4], sh[4];
    if (HOIST) {
        const int c = (int)(i0 % C);
#pragma unroll
        for (int k = 0; k < 4; ++k) { sc[k] = scale[c + k]; sh[k] = shift[c + k]; }
    }
    for (long long i = i0; i < total; i += (long long)gridDim.x * blockDim.x * 4) {
        if (!HOIST) {
            const int c = (int)(i % C);
#pragma unroll
            for (int k = 0; k < 4; ++k) { sc[k] = scale[c + k]; sh[k] = shift[c + k]; }
        }
        float v[4];
        load4<T>(x + i, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float z = v[k] * sc[k] + sh[k];
            v[k] = z > 0.f ? z : z * slope;
        }
        store4<T>(y + i, v[0], v[1], v[2], v[3]);
    }
}

// dbeta = sum g', dgamma = sum g' * xhat
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const T* __restrict__ gy, const T* __restrict__ x, long long rows, int C, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd, float slope,
                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ float s1[8][33], s2[8][33];
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + lane;
    float a = 0.f, b = 0.f;
    if (c < C) {
        const float sc = scale[c], sh = shift[c], mu = mean[c], rs = rstd[c];
        for (long long r = (long long)blockIdx.x * 8 + wy; r < rows; r += (long long)gridDim.x * 8) {
            const float xv = to_f32<T>(x[r * C + c]);
            float g = to_f32<T>(gy[r * C + c]);
            if (!(xv * sc + sh > 0.f)) g *= slope;
            a += g; b += g * (xv - mu) * rs;
        }
    }
    s1[wy][lane] = a; s2[wy][lane] = b;
    __syncthreads();
    if (wy == 0 && c < C) {
#pragma unroll
        for (int i = 1; i < 8; ++i) { a += s1[i][lane]; b += s2[i][lane]; }
        atomicAdd(dbeta + c, a);
        atomicAdd(dgamma + c, b);
    }
}

template <typename T, bool HOIST>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ gy, const T* __restrict__ x, long long total, long long rows, int C,
                    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ dgamma, const float* __restrict__ dbeta, float slope,
                    T* __restrict__ dx) {
    const float invM = 1.f / (float)rows;
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    float sc[4], sh[4], mu[4], rs[4], dg[4], db[4];
    auto load_params = [&](int c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sc[k] = scale[c + k]; sh[k] = shift[c + k]; mu[k] = mean[c + k]; rs[k] = rstd[c + k];
            dg[k] = dgamma[c + k] * invM; db[k] = dbeta[c + k] * invM;
        }
    };
    if (HOIST) load_params((int)(i0 % C));
    for (long long i = i0; i < total; i += (long long)gridDim.x * blockDim.x * 4) {
        if (!HOIST) load_params((int)(i % C));
        float xv[4], g[4];
        load4<T>(x + i, xv);
        load4<T>(gy + i, g);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!(xv[k] * sc[k] + sh[k] > 0.f)) g[k] *= slope;
            const float xh = (xv[k] - mu[k]) * rs[k];
            g[k] = sc[k] * (g[k] - db[k] - xh * dg[k]);
        }
        store4<T>(dx + i, g[0], g[1], g[2], g[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// Double backward (WGAN-GP: d/d(gy, x, gamma) of <u, dx> where dx = bn_act_bwd(gy, x, gamma), reference
// model/sradsgan.py:621 create_graph=True + :639/:886).  With N rows per channel, r = rstd, xh = (x-mean) r,
// m = lrelu'(z), gz = gy m, a = mean(gz), b = mean(gz xh), ub = mean(u), uxb = mean(u xh),
// S1 = sum u (gz - a - xh b):
//     t      = u - ub - xh uxb
//     d_gy   = m gamma r t
//     d_x    = -gamma r^2 [ (S1/N) xh + b t + uxb (gz - a - xh b) ]
//     d_gamma= r S1
// (the LeakyReLU mask is piecewise constant, so it contributes no second derivative).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd2_reduce_kernel(const T* __restrict__ u, const T* __restrict__ gy, const T* __restrict__ x, long long rows, int C,
                      const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                      const float* __restrict__ rstd, float slope, float* __restrict__ sums /* [3][C]: u, u*xh, u*gz */) {
    __shared__ float s1[8][33], s2[8][33], s3[8][33];
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + lane;
    float a = 0.f, b = 0.f, d = 0.f;
    if (c < C) {
        const float sc = scale[c], sh = shift[c], mu = mean[c], rs = rstd[c];
        for (long long r = (long long)blockIdx.x * 8 + wy; r < rows; r += (long long)gridDim.x * 8) {
            const float xv = to_f32<T>(x[r * C + c]);
            const float uv = to_f32<T>(u[r * C + c]);
            float g = to_f32<T>(gy[r * C + c]);
            if (!(xv * sc + sh > 0.f)) g *= slope;
            a += uv; b += uv * (xv - mu) * rs; d += uv * g;
        }
    }
    s1[wy][lane] = a; s2[wy][lane] = b; s3[wy][lane] = d;
    __syncthreads();
    if (wy == 0 && c < C) {
#pragma unroll
        for (int i = 1; i < 8; ++i) { a += s1[i][lane]; b += s2[i][lane]; d += s3[i][lane]; }
        atomicAdd(sums + c, a);
        atomicAdd(sums + C + c, b);
        atomicAdd(sums + 2 * C + c, d);
    }
}

template <typename T, bool HOIST>
__global__ void __launch_bounds__(256)
bn_bwd2_apply_kernel(const T* __restrict__ u, const T* __restrict__ gy, const T* __restrict__ x, long long total, long long rows, int C,
                     const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                     const float* __restrict__ sums, float slope, T* __restrict__ d_gy, T* __restrict__ d_x, float* __restrict__ d_gamma) {
    const float invN = 1.f / (float)rows;
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const float a = dbeta[c] * invN, b = dgamma[c] * invN;
            d_gamma[c] = rstd[c] * (sums[2 * C + c] - a * sums[c] - b * sums[C + c]);
        }
    }
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    float sc[4], sh[4], mu[4], rs[4], pa[4], pb[4], ub[4], uxb[4], s1n[4];
    auto load_params = [&](int c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ck = c + k;
            sc[k] = scale[ck]; sh[k] = shift[ck]; mu[k] = mean[ck]; rs[k] = rstd[ck];
            pa[k] = dbeta[ck] * invN; pb[k] = dgamma[ck] * invN;
            ub[k] = sums[ck] * invN; uxb[k] = sums[C + ck] * invN;
            s1n[k] = (sums[2 * C + ck] - pa[k] * sums[ck] - pb[k] * sums[C + ck]) * invN;
        }
    };
    if (HOIST) load_params((int)(i0 % C));
    for (long long i = i0; i < total; i += (long long)gridDim.x * blockDim.x * 4) {
        if (!HOIST) load_params((int)(i % C));
        float xv[4], g[4], uv[4], og[4], ox[4];
        load4<T>(x + i, xv);
        load4<T>(gy + i, g);
        load4<T>(u + i, uv);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a = pa[k], b = pb[k];
            const float m = (xv[k] * sc[k] + sh[k] > 0.f) ? 1.f : slope;
            const float xh = (xv[k] - mu[k]) * rs[k];
            const float gz = g[k] * m;
            const float t = uv[k] - ub[k] - xh * uxb[k];
            og[k] = m * sc[k] * t;
            ox[k] = -sc[k] * rs[k] * (s1n[k] * xh + b * t + uxb[k] * (gz - a - xh * b));
        }
        store4<T>(d_gy + i, og[0], og[1], og[2], og[3]);
        store4<T>(d_x + i, ox[0], ox[1], ox[2], ox[3]);
    }
}

template <typename T>
static int bn_bwd2_t(const void* u, const void* gy, const void* x, long long rows, int C, const float* save, const float* dgamma,
                     const float* dbeta, float slope, void* d_gy, void* d_x, float* d_gamma, float* ws, cudaStream_t st) {
    const float *mean = save, *rstd = save + C, *scale = save + 2 * C, *shift = save + 3 * C;
    cudaMemsetAsync(ws, 0, sizeof(float) * 3 * C, st);
    const int cy = (int)cdiv(C, 32);
    long long bx = std::min<long long>(cdiv(148 * 8, cy), cdiv(rows, 8));
    bn_bwd2_reduce_kernel<T><<<dim3((unsigned)bx, (unsigned)cy), 256, 0, st>>>((const T*)u, (const T*)gy, (const T*)x, rows, C, scale, shift,
                                                                               mean, rstd, slope, ws);
    const long long total = rows * C;
    const int blocks = (int)std::min<long long>(148 * 16, cdiv(total, 1024));
    if (((long long)blocks * 1024) % C == 0)
        bn_bwd2_apply_kernel<T, true><<<blocks, 256, 0, st>>>((const T*)u, (const T*)gy, (const T*)x, total, rows, C, scale, shift, mean, rstd,
                                                              dgamma, dbeta, ws, slope, (T*)d_gy, (T*)d_x, d_gamma);
    else
        bn_bwd2_apply_kernel<T, false><<<blocks, 256, 0, st>>>((const T*)u, (const T*)gy, (const T*)x, total, rows, C, scale, shift, mean, rstd,
                                                               dgamma, dbeta, ws, slope, (T*)d_gy, (T*)d_x, d_gamma);
    count_launch(2);
    return check_launch("bn_act_bwd_bwd");
}

template <typename T>
static int bn_fwd_t(const void* x, long long rows, int C, const float* gamma, const float* beta, float eps, float momentum,
                    float slope, float* rm, float* rv, void* y, float* save /* [4][C]: mean, rstd, scale, shift */, float* ws,
                    cudaStream_t st) {
    float* sum = ws; float* sq = ws + C;
    cudaMemsetAsync(ws, 0, sizeof(float) * 2 * C, st);
    const int cy = (int)cdiv(C, 32);
    long long bx = std::min<long long>(cdiv(148 * 8, cy), cdiv(rows, 8));
    bn_stats_kernel<T><<<dim3((unsigned)bx, (unsigned)cy), 256, 0, st>>>((const T*)x, rows, C, sum, sq);
    bn_finalize_kernel<T><<<(unsigned)cdiv(C, 128), 128, 0, st>>>((const T*)x, sum, sq, rows, C, gamma, beta, eps, momentum, rm, rv,
                                                                  save, save + C, save + 2 * C, save + 3 * C);
    const long long total = rows * C;
    const int blocks = (int)std::min<long long>(148 * 16, cdiv(total, 1024));
    if (((long long)blocks * 1024) % C == 0)
        bn_apply_kernel<T, true><<<blocks, 256, 0, st>>>((const T*)x, total, C, save + 2 * C, save + 3 * C, slope, (T*)y);
    else
        bn_apply_kernel<T, false><<<blocks, 256, 0, st>>>((const T*)x, total, C, save + 2 * C, save + 3 * C, slope, (T*)y);
    count_launch(3);
    return check_launch("bn_act_fwd");
}

template <typename T>
static int bn_bwd_t(const void* gy, const void* x, long long rows, int C, const float* save, float slope, void* dx, float* dgamma,
                    float* dbeta, cudaStream_t st) {
    const float *mean = save, *rstd = save + C, *scale = save + 2 * C, *shift = save + 3 * C;
    cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st);
    cudaMemsetAsync(dbeta, 0, sizeof(float) * C, st);
    const int cy = (int)cdiv(C, 32);
    long long bx = std::min<long long>(cdiv(148 * 8, cy), cdiv(rows, 8));
    bn_bwd_reduce_kernel<T><<<dim3((unsigned)bx, (unsigned)cy), 256, 0, st>>>((const T*)gy, (const T*)x, rows, C, scale, shift, mean, rstd,
                                                                              slope, dgamma, dbeta);
    const long long total = rows * C;
    const int blocks = (int)std::min<long long>(148 * 16, cdiv(total, 1024));
    if (((long long)blocks * 1024) % C == 0)
        bn_bwd_apply_kernel<T, true><<<blocks, 256, 0, st>>>((const T*)gy, (const T*)x, total, rows, C, scale, shift, mean, rstd, dgamma, dbeta,
                                                             slope, (T*)dx);
    else
        bn_bwd_apply_kernel<T, false><<<blocks, 256, 0, st>>>((const T*)gy, (const T*)x, total, rows, C, scale, shift, mean, rstd, dgamma, dbeta,
                                                              slope, (T*)dx);
    count_launch(2);
    return check_launch("bn_act_bwd");
}

int bn_act_fwd(const void* x, int dtype, long long rows, int C, const float* gamma, const float* beta, float eps, float momentum,
               float slope, float* rm, float* rv, void* y, float* save, float* ws, cudaStream_t st) {
    if (dtype == SR_F32) return bn_fwd_t<float>(x, rows, C, gamma, beta, eps, momentum, slope, rm, rv, y, save, ws, st);
    return bn_fwd_t<__nv_bfloat16>(x, rows, C, gamma, beta, eps, momentum, slope, rm, rv, y, save, ws, st);
}

int bn_act_bwd(const void* gy, const void* x, int dtype, long long rows, int C, const float* save, float slope, void* dx,
               float* dgamma, float* dbeta, cudaStream_t st) {
    if (dtype == SR_F32) return bn_bwd_t<float>(gy, x, rows, C, save, slope, dx, dgamma, dbeta, st);
    return bn_bwd_t<__nv_bfloat16>(gy, x, rows, C, save, slope, dx, dgamma, dbeta, st);
}

int bn_act_bwd_bwd(const void* u, const void* gy, const void* x, int dtype, long long rows, int C, const float* save,
                   const float* dgamma, const float* dbeta, float slope, void* d_gy, void* d_x, float* d_gamma, float* ws, cudaStream_t st) {
    if (dtype == SR_F32) return bn_bwd2_t<float>(u, gy, x, rows, C, save, dgamma, dbeta, slope, d_gy, d_x, d_gamma, ws, st);
    return bn_bwd2_t<__nv_bfloat16>(u, gy, x, rows, C, save, dgamma, dbeta, slope, d_gy, d_x, d_gamma, ws, st);
}

}  // namespace sr
