// CGAM — channel global attention of GAB_UP (reference model/sradsgan.py:178-213, light=False), C = 64, fp32 throughout:
//
//     E   = X X^T                     (64 x 64 gram over the P = H*W pixels of one image;  X[c][p] = x[n][p][c])
//     A   = softmax_j( max_j E_ij - E_ij )
//     y   = gamma * (A X) + x
//
// The reference runs two batched GEMMs, a row max, a subtraction and a softmax through cuBLAS / ATen (and as many again,
// transposed, in backward).  Here: the gram is split over pixel slices into deterministic partial tiles, a
// per-image finalize block adds them in a fixed order and runs the softmax, and one pass over the pixels applies A and the
// gamma-residual (optionally also writing the compute-dtype twin the next 1x1 convolutions read).  The logits and the softmax
// stay in fp32: the gram sums 2916 products and softmax(max - E) is sensitive to ABSOLUTE error in E (bf16 blows the 1e-2
// budget here).  Backward:
//     G = dY X^T;  dgamma = <A, G>;  dA = gamma G;  dE = -A o (dA - rowsum(dA o A))   (the row max is a per-row constant
//     shift of a softmax argument: its gradient is identically zero);  dx = dy + gamma A^T dy + (dE + dE^T) x.
// All kernels are bandwidth bound (x is read twice forward, x and dy twice backward).
#include "common.cuh"

namespace sr {

constexpr int CG_C = 64;
constexpr int CG_MAX_S = 16;

static int cg_slices(int P) {
    int s = (int)cdiv(P, 192);
    return s < 1 ? 1 : (s > CG_MAX_S ? CG_MAX_S : s);
}

// floats: gram partials [N][S][64][64], A [N][64][64], M [N][64][64], dgamma partials [N]
size_t cgam_workspace_bytes(int N, int P) {
    return sizeof(float) * ((size_t)N * cg_slices(P) * CG_C * CG_C + 2 * (size_t)N * CG_C * CG_C + (size_t)N + 64);
}

// partial[n][sl][i][j] = sum over the slice's pixels of a[p][i] * b[p][j]      (a == b: the gram E;  a = dy, b = x: G)
// block = 256 threads, thread (ti, tj) owns a 4 x 4 tile of the 64 x 64 result; pixels staged 32 at a time.
__global__ void __launch_bounds__(256)
cgam_gram_kernel(const float* a, const float* b, int P, int S, float* __restrict__ partial) {      // a may alias b (the gram)
    __shared__ __align__(16) float as[32][CG_C], bs[32][CG_C];
    const int n = blockIdx.y, sl = blockIdx.x, t = threadIdx.x;
    const int per = (P + S - 1) / S;
    const int p0 = sl * per, p1 = min(P, p0 + per);
    const int ti = t >> 4, tj = t & 15;
    const bool same = a == b;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* abase = a + (long long)n * P * CG_C;
    const float* bbase = b + (long long)n * P * CG_C;
    for (int c0 = p0; c0 < p1; c0 += 32) {
        const int cnt = min(32, p1 - c0);
        __syncthreads();
        for (int i = t; i < 32 * (CG_C / 4); i += 256) {
            const int r = i >> 4, q = (i & 15) * 4;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (r < cnt) {
                va = *reinterpret_cast<const float4*>(abase + (long long)(c0 + r) * CG_C + q);
                if (!same) vb = *reinterpret_cast<const float4*>(bbase + (long long)(c0 + r) * CG_C + q);
            }
            *reinterpret_cast<float4*>(&as[r][q]) = va;
            if (!same) *reinterpret_cast<float4*>(&bs[r][q]) = vb;
        }
        __syncthreads();
        const float (*bsel)[CG_C] = same ? as : bs;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const float4 u = *reinterpret_cast<const float4*>(&as[r][ti * 4]);
            const float4 v = *reinterpret_cast<const float4*>(&bsel[r][tj * 4]);
            const float uu[4] = {u.x, u.y, u.z, u.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(uu[i], vv[j], acc[i][j]);
        }
    }
    float* o = partial + ((long long)n * S + sl) * CG_C * CG_C;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(o + (ti * 4 + i) * CG_C + tj * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}

// per image: E = sum of the slice partials (fixed order), A = softmax_j(max_j E_ij - E_ij) exactly as torch evaluates it
// (e = rowmax - E, then exp(e - max_j e) / sum)
__global__ void __launch_bounds__(256)
cgam_softmax_kernel(const float* __restrict__ partial, int S, float* __restrict__ A) {
    __shared__ float E[CG_C][CG_C + 1];
    const int n = blockIdx.x, t = threadIdx.x;
    for (int e = t; e < CG_C * CG_C; e += 256) {
        float v = 0.f;
        for (int sl = 0; sl < S; ++sl) v += partial[((long long)n * S + sl) * CG_C * CG_C + e];
        E[e >> 6][e & 63] = v;
    }
    __syncthreads();
    if (t < CG_C) {
        float mx = -INFINITY;
        for (int j = 0; j < CG_C; ++j) mx = fmaxf(mx, E[t][j]);
        float m2 = -INFINITY;
        for (int j = 0; j < CG_C; ++j) { const float e = mx - E[t][j]; E[t][j] = e; m2 = fmaxf(m2, e); }
        float sum = 0.f;
        for (int j = 0; j < CG_C; ++j) { const float w = expf(E[t][j] - m2); E[t][j] = w; sum += w; }
        const float inv = 1.f / sum;
        for (int j = 0; j < CG_C; ++j) E[t][j] *= inv;
    }
    __syncthreads();
    for (int e = t; e < CG_C * CG_C; e += 256) A[(long long)n * CG_C * CG_C + e] = E[e >> 6][e & 63];
}

// y[p][i] = gamma * sum_j A[i][j] x[p][j] + x[p][i]     (64 pixels x 64 channels per block, 4 x 4 register tile)
template <typename T16>
__global__ void __launch_bounds__(256)
cgam_apply_kernel(const float* __restrict__ x, const float* __restrict__ A, const float* __restrict__ gamma, int P,
                  float* __restrict__ y32, T16* __restrict__ y16) {
    __shared__ __align__(16) float At[CG_C][CG_C + 4];   // At[j][i] = A[i][j]
    __shared__ __align__(16) float xs[CG_C][CG_C + 4];   // xs[j][pixel]
    const int n = blockIdx.y, t = threadIdx.x;
    const int p0 = blockIdx.x * 64;
    const float* An = A + (long long)n * CG_C * CG_C;
    for (int e = t; e < CG_C * CG_C; e += 256) At[e & 63][e >> 6] = An[e];
    {
        const int pl = t >> 2, cb = (t & 3) * 4;
        const int p = p0 + pl;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int c = cb + 16 * jj;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < P) v = *reinterpret_cast<const float4*>(x + ((long long)n * P + p) * CG_C + c);
            xs[c][pl] = v.x; xs[c + 1][pl] = v.y; xs[c + 2][pl] = v.z; xs[c + 3][pl] = v.w;
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
    for (int j = 0; j < CG_C; ++j) {
        const float4 a = *reinterpret_cast<const float4*>(&xs[j][tp * 4]);
        const float4 w = *reinterpret_cast<const float4*>(&At[j][tc * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(av[i], wv[k], acc[i][k]);
    }
    const float g = gamma[0];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = p0 + tp * 4 + i;
        if (p >= P) continue;
        const long long o = ((long long)n * P + p) * CG_C + tc * 4;
        const float4 xv = *reinterpret_cast<const float4*>(x + o);
        const float4 r = make_float4(g * acc[i][0] + xv.x, g * acc[i][1] + xv.y, g * acc[i][2] + xv.z, g * acc[i][3] + xv.w);
        *reinterpret_cast<float4*>(y32 + o) = r;
        if (y16) store4<T16>(y16 + o, r.x, r.y, r.z, r.w);
    }
}

// per image: G = sum of partials; dgamma_n = <A, G>; M = dE + dE^T with dE = -A o (gamma G - rowsum(gamma G o A))
__global__ void __launch_bounds__(256)
cgam_bwd_finalize_kernel(const float* __restrict__ partial, int S, const float* __restrict__ A, const float* __restrict__ gamma,
                         float* __restrict__ M, float* __restrict__ dgamma_part) {
    __shared__ float G[CG_C][CG_C + 1], As[CG_C][CG_C + 1], row[CG_C], red[8];
    const int n = blockIdx.x, t = threadIdx.x;
    float dg = 0.f;
    for (int e = t; e < CG_C * CG_C; e += 256) {
        float v = 0.f;
        for (int sl = 0; sl < S; ++sl) v += partial[((long long)n * S + sl) * CG_C * CG_C + e];
        const float a = A[(long long)n * CG_C * CG_C + e];
        G[e >> 6][e & 63] = v; As[e >> 6][e & 63] = a;
        dg += a * v;
    }
    dg = warp_sum(dg);
    if ((t & 31) == 0) red[t >> 5] = dg;
    __syncthreads();
    if (t == 0) { float s = 0.f; for (int w = 0; w < 8; ++w) s += red[w]; dgamma_part[n] = s; }
    const float g = gamma[0];
    if (t < CG_C) {
        float r = 0.f;
        for (int j = 0; j < CG_C; ++j) r += g * G[t][j] * As[t][j];
        row[t] = r;
    }
    __syncthreads();
    // dE[i][j] = -A[i][j] * (g G[i][j] - row[i]);  M = dE + dE^T
    for (int e = t; e < CG_C * CG_C; e += 256) {
        const int i = e >> 6, j = e & 63;
        const float dij = -As[i][j] * (g * G[i][j] - row[i]);
        const float dji = -As[j][i] * (g * G[j][i] - row[j]);
        M[(long long)n * CG_C * CG_C + e] = dij + dji;
    }
}

// dx[p][c] = dy[p][c] + sum_i gamma A[i][c] dy[p][i] + sum_j M[c][j] x[p][j];   block (0,0) also reduces dgamma over the images
__global__ void __launch_bounds__(256)
cgam_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ A, const float* __restrict__ M,
                      const float* __restrict__ gamma, const float* __restrict__ dgamma_part, int N, int P, float* __restrict__ dx,
                      float* __restrict__ dgamma, int accumulate) {
    extern __shared__ __align__(16) float cg_smem[];
    float (*Bt)[CG_C + 4] = reinterpret_cast<float (*)[CG_C + 4]>(cg_smem);                        // Bt[i][c] = gamma A[i][c]
    float (*Mt)[CG_C + 4] = reinterpret_cast<float (*)[CG_C + 4]>(cg_smem + CG_C * (CG_C + 4));    // Mt[j][c] = M[c][j] = M[j][c]
    float (*ds)[CG_C + 4] = reinterpret_cast<float (*)[CG_C + 4]>(cg_smem + 2 * CG_C * (CG_C + 4)); // ds[i][pixel]
    float (*xs)[CG_C + 4] = reinterpret_cast<float (*)[CG_C + 4]>(cg_smem + 3 * CG_C * (CG_C + 4)); // xs[j][pixel]
    const int n = blockIdx.y, t = threadIdx.x;
    const int p0 = blockIdx.x * 64;
    const float g = gamma[0];
    if (blockIdx.x == 0 && blockIdx.y == 0 && t == 0) {
        float s = 0.f;
        for (int i = 0; i < N; ++i) s += dgamma_part[i];
        dgamma[0] = accumulate ? dgamma[0] + s : s;
    }
    for (int e = t; e < CG_C * CG_C; e += 256) {
        Bt[e >> 6][e & 63] = g * A[(long long)n * CG_C * CG_C + e];
        Mt[e >> 6][e & 63] = M[(long long)n * CG_C * CG_C + e];
    }
    {
        const int pl = t >> 2, cb = (t & 3) * 4;
        const int p = p0 + pl;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int c = cb + 16 * jj;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f), w = v;
            if (p < P) {
                v = *reinterpret_cast<const float4*>(dy + ((long long)n * P + p) * CG_C + c);
                w = *reinterpret_cast<const float4*>(x + ((long long)n * P + p) * CG_C + c);
            }
            ds[c][pl] = v.x; ds[c + 1][pl] = v.y; ds[c + 2][pl] = v.z; ds[c + 3][pl] = v.w;
            xs[c][pl] = w.x; xs[c + 1][pl] = w.y; xs[c + 2][pl] = w.z; xs[c + 3][pl] = w.w;
        }
    }
    __syncthreads();
    const int tp = t >> 4, tc = t & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < CG_C; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&ds[k][tp * 4]);
        const float4 w = *reinterpret_cast<const float4*>(&Bt[k][tc * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&xs[k][tp * 4]);
        const float4 m = *reinterpret_cast<const float4*>(&Mt[k][tc * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
        const float bv[4] = {b.x, b.y, b.z, b.w}, mv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], fmaf(bv[i], mv[j], acc[i][j]));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = p0 + tp * 4 + i;
        if (p >= P) continue;
        const long long o = ((long long)n * P + p) * CG_C + tc * 4;
        const float4 d = *reinterpret_cast<const float4*>(dy + o);
        *reinterpret_cast<float4*>(dx + o) = make_float4(d.x + acc[i][0], d.y + acc[i][1], d.z + acc[i][2], d.w + acc[i][3]);
    }
}

// x, y32: [N][P][64] fp32 NHWC; y16 nullable (dtype y16_dtype); A_out [N][64][64] is saved for the backward
int cgam_fwd(const float* x, const float* gamma, int N, int P, float* y32, void* y16, int y16_dtype, float* A_out, float* ws, cudaStream_t st) {
    const int S = cg_slices(P);
    cgam_gram_kernel<<<dim3(S, N), 256, 0, st>>>(x, x, P, S, ws);
    cgam_softmax_kernel<<<N, 256, 0, st>>>(ws, S, A_out);
    const dim3 grid((unsigned)cdiv(P, 64), N);
    if (y16 && y16_dtype == SR_BF16) cgam_apply_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, A_out, gamma, P, y32, (__nv_bfloat16*)y16);
    else cgam_apply_kernel<float><<<grid, 256, 0, st>>>(x, A_out, gamma, P, y32, (float*)y16);
    count_launch(3);
    return check_launch("cgam_fwd");
}

int cgam_bwd(const float* dy, const float* x, const float* A, const float* gamma, int N, int P, float* dx, float* dgamma, int accumulate,
             float* ws, cudaStream_t st) {
    const int S = cg_slices(P);
    float* M = ws + (size_t)N * S * CG_C * CG_C;
    float* dgp = M + (size_t)N * CG_C * CG_C;
    cgam_gram_kernel<<<dim3(S, N), 256, 0, st>>>(dy, x, P, S, ws);
    cgam_bwd_finalize_kernel<<<N, 256, 0, st>>>(ws, S, A, gamma, M, dgp);
    const size_t smem = sizeof(float) * 4 * CG_C * (CG_C + 4);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(cgam_bwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    cgam_bwd_apply_kernel<<<dim3((unsigned)cdiv(P, 64), N), 256, smem, st>>>(dy, x, A, M, gamma, dgp, N, P, dx, dgamma, accumulate);
    count_launch(3);
    return check_launch("cgam_bwd");
}

}  // namespace sr
