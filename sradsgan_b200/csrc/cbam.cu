// Channel + spatial attention of the discriminator (CBAM: ChannelAttention(256) -> SpatialAttention after block 6, reference
// model/base_networks.py:366-457, used by model/sradsgan.py:476-499) as a CLOSED family of memory-bound primitives, sm_100a.
//
// The discriminator is differentiated TWICE (WGAN-GP, model/sradsgan.py:611-639), so the attention cannot be one fused
// forward/backward pair: every primitive below has derivatives that are again primitives of the family, and the host side
// (ops.py) wires them as torch.autograd.Functions whose backward passes call each other — autograd then differentiates to any
// order while every pass over a [N][P][C] activation is one of these kernels (the reference runs ~25 ATen/cuBLAS launches per
// forward, ~30 per backward, on each of the four critic passes of an iteration).
//
//   full  = [N][P][C] activations (NHWC, bf16 or fp32), chan = [N][C] fp32, pix = [N][P] fp32
//   cbam_ew      : y = x*s*m  +  s2*(g0/C + g1*[c == cidx_p])  +  (a + b*[p == idx_c])  +  acc      (each term optional) -> full
//                  gate application, and the adjoints of the channel pooling / the spatial pooling
//   cbam_red_c   : out_c = scale * sum_p a*b*m  +  sum_p a*g1*[c == cidx_p]                                              -> chan
//   cbam_red_p   : out_p = scale * sum_c a*b*s                                                                           -> pix
//   cbam_pool_hw : avg / max / first arg-max over the pixels of each (image, channel)   (AdaptiveAvgPool2d + AdaptiveMaxPool2d)
//   cbam_gather_hw : out_c = x[idx_c][c]
//   cbam_cpool   : q0_p = mean_c s_c x_pc, q1_p = max_c s_c x_pc (+ first arg-max channel)   (torch.mean / torch.max over dim 1)
//   cbam_gather_c: out_p = x[p][cidx_p] * s[cidx_p]
//   small_gemm_nt: C[M][N] = A[M][K] B[N][K]^T for the 256 <-> 16 shared MLP (any strides; replaces cuBLAS gemv launches)
// Algorithmic traffic: each full operand read once, the full result written once; everything else is O(N*(P + C)).
#include "common.cuh"

namespace sr {

template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&o)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&o)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[2 * i] = __low2float(h[i]); o[2 * i + 1] = __high2float(h[i]); }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&o)[8]);
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&o)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&o)[8]) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
}

struct CbamEw {
    const void* x; const float* s; const float* m;                       // term 1 (x nullable; s, m nullable = 1)
    const float* s2; const float* g0; const float* g1; const int* cidx;  // term 2 (s2 nullable; g0, g1 nullable = 0)
    const float* a; const float* b; const int* idx;                      // term 3 (a, b nullable)
    const void* acc;                                                     // term 4 (nullable)
    void* y;
    int N, P, C;
    float inv_c;
};

template <typename T, typename TA>
__global__ void __launch_bounds__(256)
cbam_ew_kernel(const CbamEw q) {
    const int c8 = q.C >> 3;
    const long long total = (long long)q.N * q.P * c8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % c8);
        const long long np = i / c8;
        const int n = (int)(np / q.P), p = (int)(np - (long long)n * q.P), c = cg * 8;
        const long long off = np * q.C + c;
        float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (q.x) {
            float xv[8];
            ld8<T>(reinterpret_cast<const T*>(q.x) + off, xv);
            const float mp = q.m ? q.m[np] : 1.f;
            if (q.s) {
                float sv[8];
                ld8<float>(q.s + (long long)n * q.C + c, sv);
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = xv[k] * sv[k] * mp;
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = xv[k] * mp;
            }
        }
        if (q.s2) {
            float sv[8];
            ld8<float>(q.s2 + (long long)n * q.C + c, sv);
            const float w0 = q.g0 ? q.g0[np] * q.inv_c : 0.f;
            const float w1 = q.g1 ? q.g1[np] : 0.f;
            const int cs = q.g1 ? q.cidx[np] - c : -1;
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] += sv[k] * (w0 + (k == cs ? w1 : 0.f));
        }
        if (q.a || q.b) {
            if (q.a) {
                float av[8];
                ld8<float>(q.a + (long long)n * q.C + c, av);
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] += av[k];
            }
            if (q.b) {
                float bv[8];
                ld8<float>(q.b + (long long)n * q.C + c, bv);
                const int4 i0 = *reinterpret_cast<const int4*>(q.idx + (long long)n * q.C + c), i1 = *reinterpret_cast<const int4*>(q.idx + (long long)n * q.C + c + 4);
                const int iv[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] += (iv[k] == p ? bv[k] : 0.f);
            }
        }
        if (q.acc) {
            float av[8];
            ld8<TA>(reinterpret_cast<const TA*>(q.acc) + off, av);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] += av[k];
        }
        st8<T>(reinterpret_cast<T*>(q.y) + off, o);
    }
}

// ---- reductions over the pixels of an image: one block per (image, 64-channel group); thread = (pixel lane 0..31, channel octet) ----
struct CbamRedC {
    const void* a; const void* b; const float* m; const float* g1; const int* cidx;
    float* out; float* out_max; int* out_idx;      // pool mode: out = mean, out_max / out_idx = max and its FIRST pixel
    int N, P, C;
    float scale;
};

template <typename T, typename TB, bool POOL>
__global__ void __launch_bounds__(256)
cbam_red_c_kernel(const CbamRedC q) {
    __shared__ float sh[32][65];
    __shared__ float shm[POOL ? 32 : 1][65];
    __shared__ int shi[POOL ? 32 : 1][65];
    const int groups = q.C >> 6;
    const int n = blockIdx.x / groups, c0 = (blockIdx.x - n * groups) * 64;
    const int oct = threadIdx.x & 7, pl = threadIdx.x >> 3, c = c0 + oct * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float mx[8]; int mi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { mx[k] = -INFINITY; mi[k] = 0x7fffffff; }
    const T* a = reinterpret_cast<const T*>(q.a);
    const TB* b = reinterpret_cast<const TB*>(q.b);
    for (int p = pl; p < q.P; p += 32) {
        const long long np = (long long)n * q.P + p;
        float av[8];
        ld8<T>(a + np * q.C + c, av);
        if (POOL) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                acc[k] += av[k];
                if (av[k] > mx[k]) { mx[k] = av[k]; mi[k] = p; }       // pixels visited in increasing order: the first maximum stays
            }
        } else {
            float w = q.scale * (q.m ? q.m[np] : 1.f);
            if (q.b) {
                float bv[8];
                ld8<TB>(b + np * q.C + c, bv);
#pragma unroll
                for (int k = 0; k < 8; ++k) av[k] *= bv[k];
            }
            if (q.g1) {
                const float w1 = q.g1[np];
                const int cs = q.cidx[np] - c;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += av[k] * (w + (k == cs ? w1 : 0.f));
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += av[k] * w;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        sh[pl][oct * 8 + k] = acc[k];
        if (POOL) { shm[pl][oct * 8 + k] = mx[k]; shi[pl][oct * 8 + k] = mi[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        float t = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) t += sh[r][threadIdx.x];            // fixed order
        const long long o = (long long)n * q.C + c0 + threadIdx.x;
        if (POOL) {
            float bm = -INFINITY; int bi = 0x7fffffff;
            for (int r = 0; r < 32; ++r) {
                const float v = shm[r][threadIdx.x]; const int vi = shi[r][threadIdx.x];
                if (v > bm || (v == bm && vi < bi)) { bm = v; bi = vi; }
            }
            q.out[o] = t * q.scale; q.out_max[o] = bm; q.out_idx[o] = bi == 0x7fffffff ? 0 : bi;
        } else {
            q.out[o] = t;
        }
    }
}

// ---- reductions over the channels of a pixel: one warp per pixel ----
struct CbamRedP {
    const void* a; const void* b; const float* s;
    float* out;             // red_p: [N][P]; cpool: q [N][2][P] (mean plane, max plane)
    int* cidx;              // cpool: first arg-max channel
    long long NP; int P, C;
    float scale;
};

template <typename T, typename TB, bool CPOOL>
__global__ void __launch_bounds__(256)
cbam_red_p_kernel(const CbamRedP q) {
    const int lane = threadIdx.x & 31;
    const long long np = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (np >= q.NP) return;
    const int n = (int)(np / q.P);
    const T* a = reinterpret_cast<const T*>(q.a) + np * q.C;
    const TB* b = reinterpret_cast<const TB*>(q.b);
    float acc = 0.f, mx = -INFINITY; int mi = 0x7fffffff;
    for (int c = lane * 8; c < q.C; c += 256) {
        float av[8];
        ld8<T>(a + c, av);
        if (q.b) {
            float bv[8];
            ld8<TB>(b + np * q.C + c, bv);
#pragma unroll
            for (int k = 0; k < 8; ++k) av[k] *= bv[k];
        }
        if (q.s) {
            float sv[8];
            ld8<float>(q.s + (long long)n * q.C + c, sv);
#pragma unroll
            for (int k = 0; k < 8; ++k) av[k] *= sv[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            acc += av[k];
            if (CPOOL && av[k] > mx) { mx = av[k]; mi = c + k; }
        }
    }
    acc = warp_sum(acc);
    if (CPOOL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
            if (om > mx || (om == mx && oi < mi)) { mx = om; mi = oi; }
        }
        if (lane == 0) {
            const int p = (int)(np - (long long)n * q.P);
            q.out[((long long)n * 2) * q.P + p] = acc * q.scale;
            q.out[((long long)n * 2 + 1) * q.P + p] = mx;
            q.cidx[np] = mi == 0x7fffffff ? 0 : mi;
        }
    } else if (lane == 0) {
        q.out[np] = acc * q.scale;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
cbam_gather_hw_kernel(const T* __restrict__ x, const int* __restrict__ idx, int N, int P, int C, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * C) return;
    const int n = i / C, c = i - n * C;
    out[i] = to_f32<T>(x[((long long)n * P + idx[i]) * C + c]);
}

template <typename T>
__global__ void __launch_bounds__(256)
cbam_gather_c_kernel(const T* __restrict__ x, const float* __restrict__ s, const int* __restrict__ cidx, long long NP, int P, int C,
                     float* __restrict__ out) {
    const long long np = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (np >= NP) return;
    const int c = cidx[np];
    const float v = to_f32<T>(x[np * C + c]);
    out[np] = s ? v * s[(np / P) * C + c] : v;
}

// C[m][n] = sum_k A[m*lda_m + k*lda_k] * B[n*ldb_n + k*ldb_k]: one warp per output element (M*N <= a few thousand, K <= 1024)
__global__ void __launch_bounds__(256)
small_gemm_nt_kernel(const float* __restrict__ A, long long lda_m, long long lda_k, const float* __restrict__ B, long long ldb_n, long long ldb_k,
                     int M, int N, int K, float* __restrict__ C) {
    const int lane = threadIdx.x & 31;
    const long long e = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= (long long)M * N) return;
    const int m = (int)(e / N), n = (int)(e - (long long)m * N);
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc += A[m * lda_m + k * lda_k] * B[n * ldb_n + k * ldb_k];
    acc = warp_sum(acc);
    if (lane == 0) C[e] = acc;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int ew_grid(long long items) {
    long long g = cdiv(items, 256);
    return (int)(g > 148 * 8 ? 148 * 8 : (g < 1 ? 1 : g));
}

int cbam_ew(const CbamEw& q, int dtype, int acc_dtype, cudaStream_t st) {
    const int grid = ew_grid((long long)q.N * q.P * (q.C >> 3));
    if (dtype == SR_BF16) {
        if (acc_dtype == SR_BF16) cbam_ew_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(q);
        else cbam_ew_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(q);
    } else {
        if (acc_dtype == SR_BF16) cbam_ew_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(q);
        else cbam_ew_kernel<float, float><<<grid, 256, 0, st>>>(q);
    }
    count_launch();
    return check_launch("cbam_ew_kernel");
}

int cbam_red_c(const CbamRedC& q, int a_dtype, int b_dtype, bool pool, cudaStream_t st) {
    const int grid = q.N * (q.C >> 6);
    if (pool) {
        if (a_dtype == SR_BF16) cbam_red_c_kernel<__nv_bfloat16, __nv_bfloat16, true><<<grid, 256, 0, st>>>(q);
        else cbam_red_c_kernel<float, float, true><<<grid, 256, 0, st>>>(q);
    } else if (a_dtype == SR_BF16) {
        if (b_dtype == SR_BF16) cbam_red_c_kernel<__nv_bfloat16, __nv_bfloat16, false><<<grid, 256, 0, st>>>(q);
        else cbam_red_c_kernel<__nv_bfloat16, float, false><<<grid, 256, 0, st>>>(q);
    } else {
        if (b_dtype == SR_BF16) cbam_red_c_kernel<float, __nv_bfloat16, false><<<grid, 256, 0, st>>>(q);
        else cbam_red_c_kernel<float, float, false><<<grid, 256, 0, st>>>(q);
    }
    count_launch();
    return check_launch("cbam_red_c_kernel");
}

int cbam_red_p(const CbamRedP& q, int a_dtype, int b_dtype, bool cpool, cudaStream_t st) {
    const int grid = (int)cdiv(q.NP, 8);
    if (cpool) {
        if (a_dtype == SR_BF16) cbam_red_p_kernel<__nv_bfloat16, __nv_bfloat16, true><<<grid, 256, 0, st>>>(q);
        else cbam_red_p_kernel<float, float, true><<<grid, 256, 0, st>>>(q);
    } else if (a_dtype == SR_BF16) {
        if (b_dtype == SR_BF16) cbam_red_p_kernel<__nv_bfloat16, __nv_bfloat16, false><<<grid, 256, 0, st>>>(q);
        else cbam_red_p_kernel<__nv_bfloat16, float, false><<<grid, 256, 0, st>>>(q);
    } else {
        if (b_dtype == SR_BF16) cbam_red_p_kernel<float, __nv_bfloat16, false><<<grid, 256, 0, st>>>(q);
        else cbam_red_p_kernel<float, float, false><<<grid, 256, 0, st>>>(q);
    }
    count_launch();
    return check_launch("cbam_red_p_kernel");
}

int cbam_gather_hw(const void* x, int dtype, const int* idx, int N, int P, int C, float* out, cudaStream_t st) {
    const int grid = (int)cdiv((long long)N * C, 256);
    if (dtype == SR_BF16) cbam_gather_hw_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, idx, N, P, C, out);
    else cbam_gather_hw_kernel<float><<<grid, 256, 0, st>>>((const float*)x, idx, N, P, C, out);
    count_launch();
    return check_launch("cbam_gather_hw_kernel");
}

int cbam_gather_c(const void* x, int dtype, const float* s, const int* cidx, int N, int P, int C, float* out, cudaStream_t st) {
    const long long NP = (long long)N * P;
    const int grid = (int)cdiv(NP, 256);
    if (dtype == SR_BF16) cbam_gather_c_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, s, cidx, NP, P, C, out);
    else cbam_gather_c_kernel<float><<<grid, 256, 0, st>>>((const float*)x, s, cidx, NP, P, C, out);
    count_launch();
    return check_launch("cbam_gather_c_kernel");
}

int small_gemm_nt(const float* A, long long lda_m, long long lda_k, const float* B, long long ldb_n, long long ldb_k, int M, int N, int K,
                  float* C, cudaStream_t st) {
    small_gemm_nt_kernel<<<(int)cdiv((long long)M * N, 8), 256, 0, st>>>(A, lda_m, lda_k, B, ldb_n, ldb_k, M, N, K, C);
    count_launch();
    return check_launch("small_gemm_nt_kernel");
}

}  // namespace sr
