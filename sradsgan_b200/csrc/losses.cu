// Loss reductions and the small elementwise glue of one SRADSGAN iteration, sm_100a.
//
//   sr_diff_mean_fwd / _bwd : nn.L1Loss / nn.MSELoss between the generator output and the HR batch, and between the
//                             VGG19 feature maps (reference model/sradsgan.py:685-688, :834, :838)
//   sr_mean_fwd / _bwd      : GANLoss('wgan-gp'): +-mean(D(.)) (:46-52, :847-848, :876-877)
//   sr_gp_penalty_fwd / _bwd: ||grad||_p over the colour channels per pixel -> (n-1)^2 | relu(n-1) -> mean (:623-637)
//   sr_lerp_nhwc            : alpha*real + (1-alpha)*fake, the WGAN-GP interpolates (:611)
//   sr_nchw_to_nhwc         : the host framework's NCHW fp32 batch -> NHWC compute dtype (:821-823 staging copies)
//   sr_add_cast             : out = a + b with independent dtypes (sums of gradient branches)
//
// Every reduction is deterministic: vectorised, coalesced grid-stride loads -> warp-shuffle -> shared memory ->
// one partial per block; the LAST block to finish (threadfence + ticket) adds the partials in a fixed order, writes
// the scalar and re-arms the ticket, so the caller-owned workspace only has to be zero once, when it is created.
// All kernels are bandwidth bound: algorithmic bytes = each operand read once (+ the gradient written once).
#include "common.cuh"

namespace sr {

constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 592;           // 4 x 148 SMs

// workspace: [RED_MAX_BLOCKS] float partials, then one int ticket
size_t reduce_workspace_bytes() { return sizeof(float) * RED_MAX_BLOCKS + 16; }

__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (RED_THREADS >> 5) ? sh[threadIdx.x] : 0.f;
        t = warp_sum(t);
    }
    return t;       // valid in warp 0
}

// out[0] = scale * sum of partials (fixed order); total[0] (+)= weight * out[0]
__device__ __forceinline__ void finish_reduce(float part, float* ws, float scale, float* out) {
    __shared__ float sh[RED_THREADS >> 5];
    __shared__ int last;
    const float b = block_sum(part, sh);
    int* ticket = reinterpret_cast<int*>(ws + RED_MAX_BLOCKS);
    if (threadIdx.x == 0) {
        ws[blockIdx.x] = b;
        __threadfence();
        last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    float v = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS) v += __ldcg(ws + i);
    __syncthreads();
    const float t = block_sum(v, sh);
    if (threadIdx.x == 0) {
        out[0] = t * scale;
        *ticket = 0;                 // re-armed for the next launch on this workspace
    }
}

static int red_grid(long long work_items) {
    long long g = cdiv(work_items, RED_THREADS);
    if (g < 1) g = 1;
    return (int)(g > RED_MAX_BLOCKS ? RED_MAX_BLOCKS : g);
}

// index of NHWC element i in an NCHW tensor with C channels and HW pixels per plane
__device__ __forceinline__ long long nchw_index(long long i, int C, long long HW) {
    const int c = (int)(i % C);
    const long long pix = i / C, n = pix / HW, hw = pix - n * HW;
    return (n * C + c) * HW + hw;
}

// ------------------------------------------------------------------------------------------------
// mean |a - b|^p
// ------------------------------------------------------------------------------------------------
template <typename TA, typename TB>
__global__ void __launch_bounds__(RED_THREADS)
diff_mean_fwd_kernel(const TA* __restrict__ a, const TB* __restrict__ b, long long n, int p, int b_nchw_C, long long b_HW, float inv_n,
                     float* __restrict__ ws, float* __restrict__ out) {
    float acc = 0.f;
    if (b_nchw_C > 0) {
        for (long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS) {
            const float d = to_f32<TA>(a[i]) - to_f32<TB>(b[nchw_index(i, b_nchw_C, b_HW)]);
            acc += p == 1 ? fabsf(d) : d * d;
        }
    } else {
        const long long n4 = n >> 2;
        for (long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n4; i += (long long)gridDim.x * RED_THREADS) {
            float x[4], y[4];
            load4<TA>(a + i * 4, x);
            load4<TB>(b + i * 4, y);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float d = x[k] - y[k]; acc += p == 1 ? fabsf(d) : d * d; }
        }
        if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
            const long long i = (n4 << 2) + threadIdx.x;
            const float d = to_f32<TA>(a[i]) - to_f32<TB>(b[i]);
            acc += p == 1 ? fabsf(d) : d * d;
        }
    }
    finish_reduce(acc, ws, inv_n, out);
}

// da = g * scale * d|a-b|^p / n   (sign(0) = 0 as torch's L1Loss)
template <typename TA, typename TB, typename TO>
__global__ void __launch_bounds__(256)
diff_mean_bwd_kernel(const TA* __restrict__ a, const TB* __restrict__ b, long long n, int p, int b_nchw_C, long long b_HW,
                     const float* __restrict__ g, float coef, TO* __restrict__ da) {
    const float k = g[0] * coef;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float bv = b_nchw_C > 0 ? to_f32<TB>(b[nchw_index(i, b_nchw_C, b_HW)]) : to_f32<TB>(b[i]);
        const float d = to_f32<TA>(a[i]) - bv;
        const float r = p == 1 ? (d > 0.f ? k : (d < 0.f ? -k : 0.f)) : 2.f * k * d;
        da[i] = from_f32<TO>(r);
    }
}

template <typename TA, typename TB>
static int diff_fwd_t(const void* a, const void* b, long long n, int p, int bC, long long bHW, float* out, float* ws, cudaStream_t st) {
    const int grid = red_grid(bC > 0 ? n : (n >> 2) + 1);
    diff_mean_fwd_kernel<TA, TB><<<grid, RED_THREADS, 0, st>>>((const TA*)a, (const TB*)b, n, p, bC, bHW, 1.f / (float)n, ws, out);
    count_launch();
    return check_launch("diff_mean_fwd");
}

int diff_mean_fwd(const void* a, int a_dtype, const void* b, int b_dtype, long long n, int p, int b_nchw_C, long long b_HW, float* out,
                  float* ws, cudaStream_t st) {
    if (a_dtype == SR_F32 && b_dtype == SR_F32) return diff_fwd_t<float, float>(a, b, n, p, b_nchw_C, b_HW, out, ws, st);
    if (a_dtype == SR_BF16 && b_dtype == SR_BF16) return diff_fwd_t<__nv_bfloat16, __nv_bfloat16>(a, b, n, p, b_nchw_C, b_HW, out, ws, st);
    if (a_dtype == SR_F32 && b_dtype == SR_BF16) return diff_fwd_t<float, __nv_bfloat16>(a, b, n, p, b_nchw_C, b_HW, out, ws, st);
    return diff_fwd_t<__nv_bfloat16, float>(a, b, n, p, b_nchw_C, b_HW, out, ws, st);
}

template <typename TA, typename TB, typename TO>
static int diff_bwd_t(const void* a, const void* b, long long n, int p, int bC, long long bHW, const float* g, float coef, void* da,
                      cudaStream_t st) {
    long long grid = cdiv(n, 256 * 4);
    if (grid > 1184) grid = 1184;
    diff_mean_bwd_kernel<TA, TB, TO><<<(unsigned)grid, 256, 0, st>>>((const TA*)a, (const TB*)b, n, p, bC, bHW, g, coef, (TO*)da);
    count_launch();
    return check_launch("diff_mean_bwd");
}

int diff_mean_bwd(const void* a, int a_dtype, const void* b, int b_dtype, long long n, int p, int bC, long long bHW, const float* g,
                  float scale, void* da, int da_dtype, cudaStream_t st) {
    const float coef = scale / (float)n;
#define SR_DIFF_BWD(TA, TB)                                                                                              \
    return da_dtype == SR_F32 ? diff_bwd_t<TA, TB, float>(a, b, n, p, bC, bHW, g, coef, da, st)                         \
                              : diff_bwd_t<TA, TB, __nv_bfloat16>(a, b, n, p, bC, bHW, g, coef, da, st)
    if (a_dtype == SR_F32 && b_dtype == SR_F32) { SR_DIFF_BWD(float, float); }
    if (a_dtype == SR_BF16 && b_dtype == SR_BF16) { SR_DIFF_BWD(__nv_bfloat16, __nv_bfloat16); }
    if (a_dtype == SR_F32 && b_dtype == SR_BF16) { SR_DIFF_BWD(float, __nv_bfloat16); }
    SR_DIFF_BWD(__nv_bfloat16, float);
#undef SR_DIFF_BWD
}

// ------------------------------------------------------------------------------------------------
// scale * mean(x)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
mean_fwd_kernel(const T* __restrict__ x, long long n, float scale_over_n, float* __restrict__ ws, float* __restrict__ out) {
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS) acc += to_f32<T>(x[i]);
    finish_reduce(acc, ws, scale_over_n, out);
}

template <typename T>
__global__ void __launch_bounds__(256)
fill_scaled_kernel(const float* __restrict__ g, float coef, long long n, T* __restrict__ dx) {
    const T v = from_f32<T>(g[0] * coef);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dx[i] = v;
}

int mean_fwd(const void* x, int dtype, long long n, float scale, float* out, float* ws, cudaStream_t st) {
    const int grid = red_grid(n);
    if (dtype == SR_F32) mean_fwd_kernel<float><<<grid, RED_THREADS, 0, st>>>((const float*)x, n, scale / (float)n, ws, out);
    else mean_fwd_kernel<__nv_bfloat16><<<grid, RED_THREADS, 0, st>>>((const __nv_bfloat16*)x, n, scale / (float)n, ws, out);
    count_launch();
    return check_launch("mean_fwd");
}

int mean_bwd(const float* g, float scale, long long n, void* dx, int dtype, cudaStream_t st) {
    long long grid = cdiv(n, 256);
    if (grid > 592) grid = 592;
    if (dtype == SR_F32) fill_scaled_kernel<float><<<(unsigned)grid, 256, 0, st>>>(g, scale / (float)n, n, (float*)dx);
    else fill_scaled_kernel<__nv_bfloat16><<<(unsigned)grid, 256, 0, st>>>(g, scale / (float)n, n, (__nv_bfloat16*)dx);
    count_launch();
    return check_launch("mean_bwd");
}

// ------------------------------------------------------------------------------------------------
// WGAN-GP penalty over the colour channels: grad [pixels][C], C <= 4
// norm: 0 = L2, 1 = L1, 2 = Linf;  penalty: 0 = 'LS' (n-1)^2, 1 = 'hinge' relu(n-1)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gp_norm(const float (&v)[4], int C, int norm) {
    float n = 0.f;
    if (norm == 0) { for (int c = 0; c < C; ++c) n += v[c] * v[c]; n = sqrtf(n); }
    else if (norm == 1) { for (int c = 0; c < C; ++c) n += fabsf(v[c]); }
    else { for (int c = 0; c < C; ++c) n = fmaxf(n, fabsf(v[c])); }
    return n;
}

template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
gp_penalty_fwd_kernel(const T* __restrict__ grad, long long pixels, int C, int norm, int penalty, float inv_pixels, float* __restrict__ ws,
                      float* __restrict__ out) {
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < pixels; i += (long long)gridDim.x * RED_THREADS) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < C; ++c) v[c] = to_f32<T>(grad[i * C + c]);
        const float d = gp_norm(v, C, norm) - 1.f;
        acc += penalty == 0 ? d * d : fmaxf(d, 0.f);
    }
    finish_reduce(acc, ws, inv_pixels, out);
}

// d penalty / d grad, times g*scale/pixels.  torch's norm backward yields 0 where the norm is 0 (L2: x/||x|| -> 0 by its
// masked division; L1: sign(0) = 0); Linf routes to the (first) channel of largest magnitude like torch.max over dim 1.
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
gp_penalty_bwd_kernel(const T* __restrict__ grad, long long pixels, int C, int norm, int penalty, const float* __restrict__ g, float coef,
                      TO* __restrict__ dgrad) {
    const float k = g[0] * coef;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < C; ++c) v[c] = to_f32<T>(grad[i * C + c]);
        const float n = gp_norm(v, C, norm);
        const float d = n - 1.f;
        const float dn = k * (penalty == 0 ? 2.f * d : (d > 0.f ? 1.f : 0.f));
        int amax = 0;
        if (norm == 2) for (int c = 1; c < C; ++c) if (fabsf(v[c]) > fabsf(v[amax])) amax = c;
        for (int c = 0; c < C; ++c) {
            float r;
            if (norm == 0) r = n > 0.f ? dn * v[c] / n : 0.f;
            else if (norm == 1) r = v[c] > 0.f ? dn : (v[c] < 0.f ? -dn : 0.f);
            else r = c == amax ? (v[c] > 0.f ? dn : (v[c] < 0.f ? -dn : 0.f)) : 0.f;
            dgrad[i * C + c] = from_f32<TO>(r);
        }
    }
}

int gp_penalty_fwd(const void* grad, int dtype, long long pixels, int C, int norm, int penalty, float* out, float* ws, cudaStream_t st) {
    const int grid = red_grid(pixels);
    if (dtype == SR_F32) gp_penalty_fwd_kernel<float><<<grid, RED_THREADS, 0, st>>>((const float*)grad, pixels, C, norm, penalty, 1.f / (float)pixels, ws, out);
    else gp_penalty_fwd_kernel<__nv_bfloat16><<<grid, RED_THREADS, 0, st>>>((const __nv_bfloat16*)grad, pixels, C, norm, penalty, 1.f / (float)pixels, ws, out);
    count_launch();
    return check_launch("gp_penalty_fwd");
}

int gp_penalty_bwd(const void* grad, int dtype, long long pixels, int C, int norm, int penalty, const float* g, float scale, void* dgrad,
                   int out_dtype, cudaStream_t st) {
    long long grid = cdiv(pixels, 256);
    if (grid > 1184) grid = 1184;
    const float coef = scale / (float)pixels;
#define SR_GP_BWD(T, TO) gp_penalty_bwd_kernel<T, TO><<<(unsigned)grid, 256, 0, st>>>((const T*)grad, pixels, C, norm, penalty, g, coef, (TO*)dgrad)
    if (dtype == SR_F32 && out_dtype == SR_F32) SR_GP_BWD(float, float);
    else if (dtype == SR_F32) SR_GP_BWD(float, __nv_bfloat16);
    else if (out_dtype == SR_F32) SR_GP_BWD(__nv_bfloat16, float);
    else SR_GP_BWD(__nv_bfloat16, __nv_bfloat16);
#undef SR_GP_BWD
    count_launch();
    return check_launch("gp_penalty_bwd");
}

// ------------------------------------------------------------------------------------------------
// elementwise glue
// ------------------------------------------------------------------------------------------------
// out[n][hw][c] = alpha[n] * real + (1 - alpha[n]) * fake  (NHWC out; real may be NCHW (real_nchw_C > 0) or NHWC; fake NHWC)
template <typename TR, typename TF, typename TO>
__global__ void __launch_bounds__(256)
lerp_kernel(const TR* __restrict__ real, const TF* __restrict__ fake, const float* __restrict__ alpha, long long n, long long per_image,
            int real_nchw_C, long long HW, TO* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float a = alpha[i / per_image];
        const float r = real_nchw_C > 0 ? to_f32<TR>(real[nchw_index(i, real_nchw_C, HW)]) : to_f32<TR>(real[i]);
        // the reference's expression, evaluated in fp32 in the same order: alpha * real + ((1 - alpha) * fake)
        out[i] = from_f32<TO>(a * r + (1.f - a) * to_f32<TF>(fake[i]));
    }
}

int lerp_nhwc(const void* real, int real_dtype, int real_nchw_C, const void* fake, int fake_dtype, const float* alpha, long long n,
              long long per_image, long long HW, void* out, int out_dtype, cudaStream_t st) {
    long long grid = cdiv(n, 256 * 4);
    if (grid > 1184) grid = 1184;
    if (grid < 1) grid = 1;
#define SR_LERP(TR, TF, TO) lerp_kernel<TR, TF, TO><<<(unsigned)grid, 256, 0, st>>>((const TR*)real, (const TF*)fake, alpha, n, per_image, real_nchw_C, HW, (TO*)out)
    const bool rf = real_dtype == SR_F32, ff = fake_dtype == SR_F32, of = out_dtype == SR_F32;
    if (rf && ff && of) SR_LERP(float, float, float);
    else if (rf && ff) SR_LERP(float, float, __nv_bfloat16);
    else if (rf && !ff && of) SR_LERP(float, __nv_bfloat16, float);
    else if (rf && !ff) SR_LERP(float, __nv_bfloat16, __nv_bfloat16);
    else if (!rf && ff && of) SR_LERP(__nv_bfloat16, float, float);
    else if (!rf && ff) SR_LERP(__nv_bfloat16, float, __nv_bfloat16);
    else if (of) SR_LERP(__nv_bfloat16, __nv_bfloat16, float);
    else SR_LERP(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16);
#undef SR_LERP
    count_launch();
    return check_launch("lerp_nhwc");
}

// NCHW fp32 (the host framework's batch) -> NHWC in TO; thread = pixel, C <= 4: plane reads and pixel writes both coalesce
template <typename TO>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ x, long long pixels, int C, long long HW, TO* __restrict__ out) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
        const long long n = p / HW, hw = p - n * HW;
        for (int c = 0; c < C; ++c) out[p * C + c] = from_f32<TO>(x[(n * C + c) * HW + hw]);
    }
}

int nchw_to_nhwc(const float* x, long long N, int C, long long HW, void* out, int out_dtype, cudaStream_t st) {
    const long long pixels = N * HW;
    long long grid = cdiv(pixels, 256);
    if (grid > 2368) grid = 2368;
    if (out_dtype == SR_F32) nchw_to_nhwc_kernel<float><<<(unsigned)grid, 256, 0, st>>>(x, pixels, C, HW, (float*)out);
    else nchw_to_nhwc_kernel<__nv_bfloat16><<<(unsigned)grid, 256, 0, st>>>(x, pixels, C, HW, (__nv_bfloat16*)out);
    count_launch();
    return check_launch("nchw_to_nhwc");
}

// out = a + b (b nullable: plain cast), elementwise over n elements, independent dtypes
template <typename TA, typename TB, typename TO>
__global__ void __launch_bounds__(256)
add_cast_kernel(const TA* __restrict__ a, const TB* __restrict__ b, long long n4, long long n, TO* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float x[4], y[4] = {0.f, 0.f, 0.f, 0.f};
        load4<TA>(a + i * 4, x);
        if (b) load4<TB>(b + i * 4, y);
        store4<TO>(out + i * 4, x[0] + y[0], x[1] + y[1], x[2] + y[2], x[3] + y[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        out[i] = from_f32<TO>(to_f32<TA>(a[i]) + (b ? to_f32<TB>(b[i]) : 0.f));
    }
}

int add_cast(const void* a, int a_dtype, const void* b, int b_dtype, long long n, void* out, int out_dtype, cudaStream_t st) {
    const long long n4 = n >> 2;
    long long grid = cdiv(n4 > 0 ? n4 : 1, 256);
    if (grid > 2368) grid = 2368;
#define SR_ADD(TA, TB, TO) add_cast_kernel<TA, TB, TO><<<(unsigned)grid, 256, 0, st>>>((const TA*)a, (const TB*)b, n4, n, (TO*)out)
    const bool af = a_dtype == SR_F32, bf = b_dtype == SR_F32, of = out_dtype == SR_F32;
    if (af && bf && of) SR_ADD(float, float, float);
    else if (af && bf) SR_ADD(float, float, __nv_bfloat16);
    else if (af && !bf && of) SR_ADD(float, __nv_bfloat16, float);
    else if (af && !bf) SR_ADD(float, __nv_bfloat16, __nv_bfloat16);
    else if (!af && bf && of) SR_ADD(__nv_bfloat16, float, float);
    else if (!af && bf) SR_ADD(__nv_bfloat16, float, __nv_bfloat16);
    else if (of) SR_ADD(__nv_bfloat16, __nv_bfloat16, float);
    else SR_ADD(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16);
#undef SR_ADD
    count_launch();
    return check_launch("add_cast");
}

}  // namespace sr
