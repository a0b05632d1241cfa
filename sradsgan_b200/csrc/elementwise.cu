// Bandwidth-bound helper kernels: weight packing, column reductions, fused Adam(+clamp).
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace sr {

// OIHW fp32 -> [tap][Cout][Cin] (mode 0) or [tap][Cin][Cout] (mode 1)
// shuffle_r > 1 (mode 0 only): output rows are permuted subpixel-major, row n' = sub*(Cout/r^2) + c holds
// original output channel c*r^2 + sub, so that PixelShuffle becomes a contiguous store per sub-pixel.
template <typename T>
__global__ void pack_weights_kernel(const float* __restrict__ w, T* __restrict__ out, int Cout, int Cin, int taps, int mode, int shuffle_r) {
    const long long total = (long long)Cout * Cin * taps;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // i indexes the OUTPUT (coalesced writes)
        int tap, co, ci;
        if (mode == 0) {
            ci = (int)(i % Cin); long long q = i / Cin; co = (int)(q % Cout); tap = (int)(q / Cout);
        } else {
            co = (int)(i % Cout); long long q = i / Cout; ci = (int)(q % Cin); tap = (int)(q / Cin);
        }
        if (shuffle_r > 1) {
            const int r2 = shuffle_r * shuffle_r, cq = Cout / r2;
            const int sub = co / cq, c = co - sub * cq;
            co = c * r2 + sub;
        }
        out[i] = from_f32<T>(w[((long long)co * Cin + ci) * taps + tap]);
    }
}

// Batched variant: ONE launch re-packs every convolution weight of a network after its optimiser step.
// table[e] = {w (fp32 OIHW), out, Cout, Cin, taps, mode, shuffle_r, first block}; a block handles 1024 consecutive
// OUTPUT elements of one entry (found by binary search over the first-block column).
template <typename T>
__global__ void __launch_bounds__(256)
pack_weights_batched_kernel(const long long* __restrict__ table, int n_entries) {
    __shared__ int e_s;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_entries - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (table[mid * 8 + 7] <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        e_s = lo;
    }
    __syncthreads();
    const long long* t = table + (long long)e_s * 8;
    const float* __restrict__ w = reinterpret_cast<const float*>(t[0]);
    T* __restrict__ out = reinterpret_cast<T*>(t[1]);
    const int Cout = (int)t[2], Cin = (int)t[3], taps = (int)t[4], mode = (int)t[5], shuffle_r = (int)t[6];
    const long long total = (long long)Cout * Cin * taps;
    const long long base = ((long long)blockIdx.x - t[7]) * 1024;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long i = base + k * 256 + threadIdx.x;
        if (i >= total) break;
        int tap, co, ci;
        if (mode == 0) {
            ci = (int)(i % Cin); long long q = i / Cin; co = (int)(q % Cout); tap = (int)(q / Cout);
        } else {
            co = (int)(i % Cout); long long q = i / Cout; ci = (int)(q % Cin); tap = (int)(q / Cin);
        }
        if (shuffle_r > 1) {
            const int r2 = shuffle_r * shuffle_r, cq = Cout / r2;
            const int sub = co / cq, c = co - sub * cq;
            co = c * r2 + sub;
        }
        out[i] = from_f32<T>(w[((long long)co * Cin + ci) * taps + tap]);
    }
}

// sum (and sum of squares) over rows of x[rows][C]; block = 32 channels x 8 row-lanes
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, long long rows, int C, float* __restrict__ sum, float* __restrict__ sq) {
    __shared__ float s1[8][33], s2[8][33];
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + lane;
    float a = 0.f, b = 0.f;
    if (c < C) {
        for (long long r = (long long)blockIdx.x * 8 + wy; r < rows; r += (long long)gridDim.x * 8) {
            const float v = to_f32<T>(x[r * C + c]);
            a += v;
            b += v * v;
        }
    }
    s1[wy][lane] = a; s2[wy][lane] = b;
    __syncthreads();
    if (wy == 0 && c < C) {
#pragma unroll
        for (int i = 1; i < 8; ++i) { a += s1[i][lane]; b += s2[i][lane]; }
        atomicAdd(sum + c, a);
        if (sq) atomicAdd(sq + c, b);
    }
}

// Vectorised variant for C in {64, 128, 256, 512}: a thread owns 4 consecutive channels (one 8- or 16-byte load per row),
// C/4 threads span a row and the block's 256 threads cover 1024/C rows per iteration, so the per-thread channel is fixed
// and the accumulators stay in registers; rows of the same channel meet in shared memory, then one atomic per channel.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ x, long long rows, int C, float* __restrict__ sum, float* __restrict__ sq) {
    __shared__ float red[2][256][4];
    const int tpr = C >> 2;                       // threads per row
    const int rpb = 256 / tpr;                    // rows per block iteration
    const int tr = threadIdx.x / tpr, tc = (threadIdx.x - tr * tpr) * 4;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long r = (long long)blockIdx.x * rpb + tr; r < rows; r += (long long)gridDim.x * rpb) {
        float v[4];
        load4<T>(x + r * C + tc, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[k] += v[k]; b[k] += v[k] * v[k]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { red[0][threadIdx.x][k] = a[k]; red[1][threadIdx.x][k] = b[k]; }
    __syncthreads();
    if (threadIdx.x < tpr) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float s = 0.f, q = 0.f;
            for (int i = 0; i < rpb; ++i) { s += red[0][i * tpr + threadIdx.x][k]; q += red[1][i * tpr + threadIdx.x][k]; }
            atomicAdd(sum + tc + k, s);
            if (sq) atomicAdd(sq + tc + k, q);
        }
    }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt,
                            const int* __restrict__ step_dev, float grad_scale, float lo, float hi) {
    if (step_dev) {   // CUDA-graph replay: the step counter lives on the device
        const float t = (float)(*step_dev);
        bc1 = 1.f - powf(b1, t);
        bc2_sqrt = sqrtf(1.f - powf(b2, t));
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * grad_scale;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        float pi = p[i] - (lr / bc1) * (mi / denom);
        if (hi > lo) pi = fminf(fmaxf(pi, lo), hi);
        p[i] = pi;
    }
}

template <typename TG, typename TY, typename TO>
__global__ void act_bwd_kernel(const TG* __restrict__ gy, const TY* __restrict__ y, int act, float slope, int r, int N, int Ho, int Wo,
                               int C, TO* __restrict__ out) {
    const long long total = (long long)N * Ho * Wo * C;
    const int r2 = r * r, cq = C / r2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long src = i;
        if (r > 1) {      // i indexes out (n, oy, ox, c = cc*r2 + sub) <- shuffled (n, oy*r+si, ox*r+sj, cc)
            const int c = (int)(i % C); long long q = i / C;
            const int ox = (int)(q % Wo); q /= Wo;
            const int oy = (int)(q % Ho); const int n = (int)(q / Ho);
            const int cc = c / r2, sub = c - cc * r2, si = sub / r, sj = sub - si * r;
            src = ((((long long)n * Ho * r + (oy * r + si)) * ((long long)Wo * r)) + (ox * r + sj)) * cq + cc;
        }
        float g = to_f32<TG>(gy[src]);
        if (act == SR_ACT_LRELU) { if (!(to_f32<TY>(y[src]) > 0.f)) g *= slope; }
        else if (act == SR_ACT_RELU) { if (!(to_f32<TY>(y[src]) > 0.f)) g = 0.f; }
        out[i] = from_f32<TO>(g);
    }
}

// PixelShuffle(2) variant, vectorised: thread = (output pixel, pair of pre-shuffle channels cc, cc+1).  The 8 outputs
// c = cc*4 + sub are contiguous (one 16-byte store in bf16); each of the 4 sub-pixels contributes 2 contiguous source
// channels, and a warp's 32 pairs cover one whole 64-channel source pixel per load (the scalar kernel above gathers
// 2-byte elements from four different rows per thread: 396 us for the 216^2 x 64 gradient, ~6x its traffic time).
template <typename T> __device__ __forceinline__ float2 load2f(const T* p);
template <> __device__ __forceinline__ float2 load2f<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 load2f<__nv_bfloat16>(const __nv_bfloat16* p) {
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(p);
    return make_float2(__low2float(v), __high2float(v));
}

template <typename TG, typename TY, typename TO>
__global__ void __launch_bounds__(256)
act_bwd_ps2_kernel(const TG* __restrict__ gy, const TY* __restrict__ y, int act, float slope, int N, int Ho, int Wo, int C,
                   TO* __restrict__ out) {
    const int cq = C >> 2, pairs = cq >> 1;
    const long long total = (long long)N * Ho * Wo * pairs;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int pr = (int)(i % pairs); const long long pix = i / pairs;
        const int ox = (int)(pix % Wo); const long long q = pix / Wo;
        const int oy = (int)(q % Ho); const int n = (int)(q / Ho);
        float o[8];
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
            const int si = sub >> 1, sj = sub & 1;
            const long long src = ((((long long)n * Ho * 2 + (oy * 2 + si)) * ((long long)Wo * 2)) + (ox * 2 + sj)) * cq + pr * 2;
            float2 g = load2f<TG>(gy + src);
            const float2 yv = load2f<TY>(y + src);
            const float f = act == SR_ACT_LRELU ? slope : 0.f;
            if (act == SR_ACT_LRELU || act == SR_ACT_RELU) {
                if (!(yv.x > 0.f)) g.x *= f;
                if (!(yv.y > 0.f)) g.y *= f;
            }
            o[sub] = g.x; o[4 + sub] = g.y;
        }
        TO* dst = out + pix * C + pr * 8;
        store4<TO>(dst, o[0], o[1], o[2], o[3]);
        store4<TO>(dst + 4, o[4], o[5], o[6], o[7]);
    }
}

template <typename TG, typename TY>
static void act_bwd_launch(const void* gy, const void* y, int act, float slope, int r, int N, int Ho, int Wo, int C, void* out,
                           int out_dtype, int blocks, cudaStream_t st) {
    if (r == 2 && C % 8 == 0) {
        const long long total = (long long)N * Ho * Wo * (C / 8);
        const int b2 = (int)std::min<long long>(148 * 16, (long long)cdiv(total, 256));
        if (out_dtype == SR_F32) act_bwd_ps2_kernel<TG, TY, float><<<b2, 256, 0, st>>>((const TG*)gy, (const TY*)y, act, slope, N, Ho, Wo, C, (float*)out);
        else act_bwd_ps2_kernel<TG, TY, __nv_bfloat16><<<b2, 256, 0, st>>>((const TG*)gy, (const TY*)y, act, slope, N, Ho, Wo, C, (__nv_bfloat16*)out);
        return;
    }
    if (out_dtype == SR_F32) act_bwd_kernel<TG, TY, float><<<blocks, 256, 0, st>>>((const TG*)gy, (const TY*)y, act, slope, r, N, Ho, Wo, C, (float*)out);
    else act_bwd_kernel<TG, TY, __nv_bfloat16><<<blocks, 256, 0, st>>>((const TG*)gy, (const TY*)y, act, slope, r, N, Ho, Wo, C, (__nv_bfloat16*)out);
}

int act_bwd(const void* gy, int gy_dtype, const void* y, int y_dtype, int act, float slope, int r, int N, int Ho, int Wo, int C,
            void* out, int out_dtype, cudaStream_t st) {
    const long long total = (long long)N * Ho * Wo * C;
    if (total <= 0) return SR_OK;
    const int blocks = (int)std::min<long long>(148 * 16, (long long)cdiv(total, 256));
    if (r < 1) r = 1;
    if (gy_dtype == SR_F32 && y_dtype == SR_F32) act_bwd_launch<float, float>(gy, y, act, slope, r, N, Ho, Wo, C, out, out_dtype, blocks, st);
    else if (gy_dtype == SR_F32) act_bwd_launch<float, __nv_bfloat16>(gy, y, act, slope, r, N, Ho, Wo, C, out, out_dtype, blocks, st);
    else if (y_dtype == SR_F32) act_bwd_launch<__nv_bfloat16, float>(gy, y, act, slope, r, N, Ho, Wo, C, out, out_dtype, blocks, st);
    else act_bwd_launch<__nv_bfloat16, __nv_bfloat16>(gy, y, act, slope, r, N, Ho, Wo, C, out, out_dtype, blocks, st);
    count_launch();
    return check_launch("act_bwd_kernel");
}

// ------------------------------------------------------------------------------------------------
// MaxPool2d(2, 2) of VGG19 features[4] / [9] (reference model/sradsgan.py:92-94), NHWC.  thread = (output pixel, 8 channels):
// four vector loads, one vector store; the backward recomputes the arg-max from the saved input (first maximum in
// row-major window order, like torch's max_pool2d_with_indices) instead of storing an index tensor.  HBM bound.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
maxpool2_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C, T* __restrict__ y) {
    const int Ho = H >> 1, Wo = W >> 1, cv = C >> 3;
    const long long total = (long long)N * Ho * Wo * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % cv) * 8; const long long pix = i / cv;
        const int ox = (int)(pix % Wo); const long long q = pix / Wo;
        const int oy = (int)(q % Ho); const int n = (int)(q / Ho);
        const T* base = x + ((((long long)n * H + oy * 2) * W) + ox * 2) * C + c0;
        float m[8], v[8];
        load4<T>(base, *reinterpret_cast<float (*)[4]>(m)); load4<T>(base + 4, *reinterpret_cast<float (*)[4]>(m + 4));
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            const T* pk = base + ((long long)(k >> 1) * W + (k & 1)) * C;
            load4<T>(pk, *reinterpret_cast<float (*)[4]>(v)); load4<T>(pk + 4, *reinterpret_cast<float (*)[4]>(v + 4));
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = v[j] > m[j] ? v[j] : m[j];
        }
        T* o = y + pix * C + c0;
        store4<T>(o, m[0], m[1], m[2], m[3]);
        store4<T>(o + 4, m[4], m[5], m[6], m[7]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
maxpool2_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, int N, int H, int W, int C, T* __restrict__ dx) {
    const int Ho = H >> 1, Wo = W >> 1, cv = C >> 3;
    const long long total = (long long)N * Ho * Wo * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % cv) * 8; const long long pix = i / cv;
        const int ox = (int)(pix % Wo); const long long q = pix / Wo;
        const int oy = (int)(q % Ho); const int n = (int)(q / Ho);
        const long long off = ((((long long)n * H + oy * 2) * W) + ox * 2) * C + c0;
        float v[4][8], g[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const T* pk = x + off + ((long long)(k >> 1) * W + (k & 1)) * C;
            load4<T>(pk, *reinterpret_cast<float (*)[4]>(v[k])); load4<T>(pk + 4, *reinterpret_cast<float (*)[4]>(v[k] + 4));
        }
        load4<T>(dy + pix * C + c0, *reinterpret_cast<float (*)[4]>(g)); load4<T>(dy + pix * C + c0 + 4, *reinterpret_cast<float (*)[4]>(g + 4));
        int arg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float m = v[0][j]; int a = 0;
#pragma unroll
            for (int k = 1; k < 4; ++k) if (v[k][j] > m) { m = v[k][j]; a = k; }
            arg[j] = a;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            T* o = dx + off + ((long long)(k >> 1) * W + (k & 1)) * C;
            store4<T>(o, arg[0] == k ? g[0] : 0.f, arg[1] == k ? g[1] : 0.f, arg[2] == k ? g[2] : 0.f, arg[3] == k ? g[3] : 0.f);
            store4<T>(o + 4, arg[4] == k ? g[4] : 0.f, arg[5] == k ? g[5] : 0.f, arg[6] == k ? g[6] : 0.f, arg[7] == k ? g[7] : 0.f);
        }
    }
}

int maxpool2_fwd(const void* x, int dtype, int N, int H, int W, int C, void* y, cudaStream_t st) {
    const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
    if (total <= 0) return SR_OK;
    const int blocks = (int)std::min<long long>(148 * 16, (long long)cdiv(total, 256));
    if (dtype == SR_F32) maxpool2_fwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, N, H, W, C, (float*)y);
    else maxpool2_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, N, H, W, C, (__nv_bfloat16*)y);
    count_launch();
    return check_launch("maxpool2_fwd_kernel");
}

int maxpool2_bwd(const void* dy, const void* x, int dtype, int N, int H, int W, int C, void* dx, cudaStream_t st) {
    const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
    const size_t esz = dtype == SR_F32 ? 4 : 2;
    if ((H & 1) || (W & 1)) cudaMemsetAsync(dx, 0, (size_t)N * H * W * C * esz, st);      // the odd last row / column is in no window
    if (total <= 0) return SR_OK;
    const int blocks = (int)std::min<long long>(148 * 16, (long long)cdiv(total, 256));
    if (dtype == SR_F32) maxpool2_bwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)dy, (const float*)x, N, H, W, C, (float*)dx);
    else maxpool2_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, N, H, W, C, (__nv_bfloat16*)dx);
    count_launch();
    return check_launch("maxpool2_bwd_kernel");
}

int pack_weights(const float* w, void* out, int Cout, int Cin, int kh, int kw, int mode, int dtype, int shuffle_r, cudaStream_t st) {
    const long long total = (long long)Cout * Cin * kh * kw;
    const int blocks = (int)std::min<long long>(148 * 8, (long long)cdiv(total, 256));
    if (dtype == SR_F32) pack_weights_kernel<float><<<blocks, 256, 0, st>>>(w, (float*)out, Cout, Cin, kh * kw, mode, shuffle_r);
    else pack_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(w, (__nv_bfloat16*)out, Cout, Cin, kh * kw, mode, shuffle_r);
    count_launch();
    return check_launch("pack_weights_kernel");
}

int pack_weights_batched(const long long* table, int n_entries, int total_blocks, int dtype, cudaStream_t st) {
    if (n_entries <= 0 || total_blocks <= 0) return SR_OK;
    if (dtype == SR_F32) pack_weights_batched_kernel<float><<<total_blocks, 256, 0, st>>>(table, n_entries);
    else pack_weights_batched_kernel<__nv_bfloat16><<<total_blocks, 256, 0, st>>>(table, n_entries);
    count_launch();
    return check_launch("pack_weights_batched_kernel");
}

int colsum(const void* x, int dtype, long long rows, int C, float* sum, float* sq, int accumulate, cudaStream_t st) {
    if (!accumulate) {
        cudaMemsetAsync(sum, 0, sizeof(float) * C, st);
        if (sq) cudaMemsetAsync(sq, 0, sizeof(float) * C, st);
    }
    if (rows <= 0) return SR_OK;
    const int cy = (int)cdiv(C, 32);
    long long bx = cdiv(148 * 8, cy);
    if (bx > cdiv(rows, 8)) bx = cdiv(rows, 8);
    dim3 grid((unsigned)bx, (unsigned)cy);
    if (C == 64 || C == 128 || C == 256 || C == 512) {
        const int rpb = 1024 / C;
        const unsigned blocks = (unsigned)std::min<long long>(148 * 4, cdiv(rows, (long long)rpb * 4));
        if (dtype == SR_F32) colsum_vec_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, rows, C, sum, sq);
        else colsum_vec_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, rows, C, sum, sq);
        count_launch();
        return check_launch("colsum_vec_kernel");
    }
    if (dtype == SR_F32) colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)x, rows, C, sum, sq);
    else colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, rows, C, sum, sq);
    count_launch();
    return check_launch("colsum_kernel");
}

int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
              int step, const int* step_dev, float grad_scale, float lo, float hi, cudaStream_t st) {
    if (n <= 0) return SR_OK;
    const float bc1 = (float)(1.0 - pow((double)b1, (double)step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, (double)step));
    const int blocks = (int)std::min<long long>(148 * 8, (long long)cdiv(n, 256));
    adam_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, bc1, bc2_sqrt, step_dev, grad_scale, lo, hi);
    count_launch();
    return check_launch("adam_kernel");
}

// ------------------------------------------------------------------------------------------------
// One separable pass of PIL's 8-bit resampling (Pillow src/libImaging/Resample.c, ImagingResampleHorizontal_8bpc /
// ImagingResampleVertical_8bpc — what `Image.resize(..., Image.BICUBIC)` of the reference's datasets runs,
// data/dataset.py:403-438): out = clip8((2^21 + sum_k in[lo + k] * coeff[k]) >> 22) with the caller's fixed-point coefficient
// table (22 fractional bits, built on the host exactly like precompute_coeffs / normalize_coeffs_8bpc).  int32 arithmetic,
// arithmetic shift: bit-exact.  in: [planes][H][W] uint8; axis 0 resamples along W (out [planes][H][out_size]), axis 1 along H.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
resample_u8_kernel(const unsigned char* __restrict__ in, int planes, int H, int W, unsigned char* __restrict__ out, int out_size, int axis,
                   const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
    const int oh = axis ? out_size : H, ow = axis ? W : out_size;
    const long long total = (long long)planes * oh * ow;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % ow);
        const long long r = i / ow;
        const int y = (int)(r % oh), pl = (int)(r / oh);
        const int o = axis ? y : x;
        const int lo = bounds[2 * o], n = bounds[2 * o + 1];
        const int* k = kk + (long long)o * ksize;
        const unsigned char* src = in + (long long)pl * H * W + (axis ? (long long)lo * W + x : (long long)y * W + lo);
        const int stride = axis ? W : 1;
        int ss = 1 << 21;
        for (int t = 0; t < n; ++t) ss += (int)src[(long long)t * stride] * k[t];
        ss >>= 22;
        out[i] = (unsigned char)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
    }
}

int resample_u8(const unsigned char* in, int planes, int H, int W, unsigned char* out, int out_size, int axis, const int* bounds,
                const int* kk, int ksize, cudaStream_t st) {
    const long long total = (long long)planes * (axis ? out_size : H) * (axis ? W : out_size);
    long long grid = cdiv(total, 256);
    if (grid > 148 * 16) grid = 148 * 16;
    resample_u8_kernel<<<(int)grid, 256, 0, st>>>(in, planes, H, W, out, out_size, axis, bounds, kk, ksize);
    count_launch();
    return check_launch("resample_u8_kernel");
}

}  // namespace sr
