// Single-kernel forward of the fused local-attention chain  z = Conv1x1(SLAM(CLAM(x))) + t  (see la_chain.cu for the
// math and the reference citations).  The five-kernel path in la_chain.cu reads x three times and pays five launches
// per chain (48 chains per generator forward); here the grid is ONE co-resident wave — `cpi` CTAs per image, each owning
// a contiguous pixel slice that it loads into shared memory ONCE — and the two global dependencies of the chain
// (the per-channel pooling over the whole image, and the 7x7 stencil over the per-pixel channel statistics) are crossed
// with two per-image barriers on global counters instead of kernel boundaries.  HBM/L2 traffic is the ideal one:
// x and t read once, z32 / z16 written once (+ the small tensors saved for the backward pass).
#include <algorithm>

#include "common.cuh"

namespace sr {

constexpr int LF_C = 64;

struct LaFusedParams {
    const void* x; const float* t; const float* fc1; const float* fc2; const float* w7; const float* Wm; const float* bias;
    int N, H, W, P, Cr, cpi, ppc;
    float* z32; void* z16;
    float* s_out; float* m_out; float* avg_out; float* max_out; int* pstar; float* q; unsigned char* cstar;
    float* psum; float* pmax; int* pidx;          // [N][cpi][64] partial pooling results
    unsigned int* bar;                             // [N][2] arrival counters (zeroed before the launch)
};

__device__ __forceinline__ void la_image_barrier(unsigned int* counter, unsigned int expected) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(counter, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < expected);
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(256, 1)
la_fused_fwd_kernel(const LaFusedParams p) {
    extern __shared__ __align__(16) unsigned char lf_smem[];
    T* xs = reinterpret_cast<T*>(lf_smem);                                         // [ppc][64]
    float* ws = reinterpret_cast<float*>(lf_smem + (((size_t)p.ppc * LF_C * sizeof(T) + 15) & ~(size_t)15));   // ws[ci][co], pitch 68
    float* vs = ws + LF_C * 68;                                                    // vs[ci][64 pixels], pitch 68
    float* ms = vs + LF_C * 68;                                                    // [ppc]
    float* s_s = ms + ((p.ppc + 3) & ~3);                                          // [64]
    float* red_f = s_s + LF_C;                                                     // [8][64] sums, [8][64] maxima
    int* red_i = reinterpret_cast<int*>(red_f + 2 * 8 * LF_C);                     // [8][64] arg-max pixel
    float* avg_s = reinterpret_cast<float*>(red_i + 8 * LF_C);                     // [64]
    float* max_s = avg_s + LF_C;                                                   // [64]
    float* hid = max_s + LF_C;                                                     // [32]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.x / p.cpi, part = blockIdx.x - n * p.cpi;
    const int p0 = part * p.ppc;
    const int np = max(0, min(p.ppc, p.P - p0));
    const T* x = reinterpret_cast<const T*>(p.x) + ((long long)n * p.P + p0) * LF_C;

    // ---- phase 1: pixel slice -> shared memory; partial channel pooling --------------------------------------------
    constexpr int VEC = 16 / (int)sizeof(T);
    for (int i = tid; i < np * (LF_C / VEC); i += 256)
        reinterpret_cast<uint4*>(xs)[i] = reinterpret_cast<const uint4*>(x)[i];
    {   // ws[ci][co] = W[co][ci]: 16 independent loads per thread first, then the transposed stores
        float wr[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) wr[k] = p.Wm[tid + k * 256];
#pragma unroll
        for (int k = 0; k < 16; ++k) { const int i = tid + k * 256; ws[(i % LF_C) * 68 + i / LF_C] = wr[k]; }
    }
    __syncthreads();
    {
        float s0 = 0.f, s1 = 0.f, m0 = -INFINITY, m1 = -INFINITY;
        int i0 = 0x7fffffff, i1 = 0x7fffffff;
        for (int pp = warp; pp < np; pp += 8) {
            const float a = to_f32<T>(xs[pp * LF_C + lane * 2]), b = to_f32<T>(xs[pp * LF_C + lane * 2 + 1]);
            s0 += a; s1 += b;
            if (a > m0) { m0 = a; i0 = p0 + pp; }
            if (b > m1) { m1 = b; i1 = p0 + pp; }
        }
        red_f[warp * LF_C + lane * 2] = s0; red_f[warp * LF_C + lane * 2 + 1] = s1;
        red_f[8 * LF_C + warp * LF_C + lane * 2] = m0; red_f[8 * LF_C + warp * LF_C + lane * 2 + 1] = m1;
        red_i[warp * LF_C + lane * 2] = i0; red_i[warp * LF_C + lane * 2 + 1] = i1;
    }
    __syncthreads();
    if (tid < LF_C) {
        float s = 0.f, m = -INFINITY; int idx = 0x7fffffff;
        for (int w = 0; w < 8; ++w) {
            s += red_f[w * LF_C + tid];
            const float mv = red_f[8 * LF_C + w * LF_C + tid]; const int iv = red_i[w * LF_C + tid];
            if (mv > m || (mv == m && iv < idx)) { m = mv; idx = iv; }
        }
        const long long o = ((long long)n * p.cpi + part) * LF_C + tid;
        p.psum[o] = s; p.pmax[o] = m; p.pidx[o] = idx;
    }
    la_image_barrier(p.bar + n * 2, (unsigned)p.cpi);

    // ---- phase 2: every CTA of the image finishes the pooling and evaluates the gate MLP ----------------------------
    // (the partials of the other CTAs are pulled from L2 with independent coalesced loads into vs, which is free here)
    {
        float* ps = vs; float* pm = vs + p.cpi * LF_C; int* pi = reinterpret_cast<int*>(vs + 2 * p.cpi * LF_C);   // cpi <= 16
        for (int i = tid; i < p.cpi * LF_C; i += 256) {
            const long long o = (long long)n * p.cpi * LF_C + i;
            ps[i] = __ldcg(p.psum + o); pm[i] = __ldcg(p.pmax + o); pi[i] = __ldcg(p.pidx + o);
        }
    }
    __syncthreads();
    if (tid < LF_C) {
        const float* ps = vs; const float* pm = vs + p.cpi * LF_C; const int* pi = reinterpret_cast<const int*>(vs + 2 * p.cpi * LF_C);
        float s = 0.f, m = -INFINITY; int idx = 0x7fffffff;
        for (int k = 0; k < p.cpi; ++k) {
            s += ps[k * LF_C + tid];
            const float mv = pm[k * LF_C + tid]; const int iv = pi[k * LF_C + tid];
            if (mv > m || (mv == m && iv < idx)) { m = mv; idx = iv; }
        }
        avg_s[tid] = s / (float)p.P; max_s[tid] = m;
        if (part == 0) { p.avg_out[n * LF_C + tid] = avg_s[tid]; p.max_out[n * LF_C + tid] = m; p.pstar[n * LF_C + tid] = idx; }
    }
    __syncthreads();
    if (tid < p.Cr) {
        float u = 0.f, v = 0.f;
        for (int k = 0; k < LF_C; ++k) { const float f = p.fc1[tid * LF_C + k]; u += f * avg_s[k]; v += f * max_s[k]; }
        hid[tid] = fmaxf(u, 0.f) + fmaxf(v, 0.f);
    }
    __syncthreads();
    if (tid < LF_C) {
        float o = 0.f;
        for (int j = 0; j < p.Cr; ++j) o += p.fc2[tid * p.Cr + j] * hid[j];
        const float s = 1.f / (1.f + __expf(-o));
        s_s[tid] = s;
        if (part == 0) p.s_out[n * LF_C + tid] = s;
    }
    __syncthreads();

    // ---- phase 3: per-pixel channel statistics of u = s * x ---------------------------------------------------------
    {
        const float sa = s_s[lane * 2], sb = s_s[lane * 2 + 1];
        for (int pp = warp; pp < np; pp += 8) {
            const float u0 = to_f32<T>(xs[pp * LF_C + lane * 2]) * sa, u1 = to_f32<T>(xs[pp * LF_C + lane * 2 + 1]) * sb;
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
                const long long pix = (long long)n * p.P + p0 + pp;
                *reinterpret_cast<float2*>(p.q + pix * 2) = make_float2(sum * (1.f / LF_C), mv);
                p.cstar[pix] = (unsigned char)mi;
            }
        }
    }
    la_image_barrier(p.bar + n * 2 + 1, (unsigned)p.cpi);

    // ---- phase 4: m = sigmoid(conv7x7(q)) for the pixels of this slice.  The rows of q the slice touches (+-3) are
    // pulled from L2 into shared memory with coalesced independent loads first (vs is free until phase 5). ----------
    {
        float* w7s = ms + ((p.ppc + 3) & ~3) + LF_C + 2 * 8 * LF_C;      // reuse red_i (8*64 ints >= 98 floats)
        if (tid < 98) w7s[tid] = p.w7[tid];
        const int y_lo = max(0, p0 / p.W - 3), y_hi = min(p.H - 1, (p0 + max(np, 1) - 1) / p.W + 3);
        const int nq = (y_hi - y_lo + 1) * p.W;                          // pixels of q staged: <= (rows + 7) * W
        float2* qs = reinterpret_cast<float2*>(vs);
        const bool staged = (size_t)nq * sizeof(float2) <= sizeof(float) * LF_C * 68;
        if (staged) {
            const float2* qg = reinterpret_cast<const float2*>(p.q) + ((long long)n * p.H + y_lo) * p.W;
            for (int i = tid; i < nq; i += 256) qs[i] = __ldcg(qg + i);
        }
        __syncthreads();
        for (int pp = tid; pp < np; pp += 256) {
            const int pl = p0 + pp;
            const int yy0 = pl / p.W, xx0 = pl - yy0 * p.W;
            float e = 0.f;
            for (int ky = 0; ky < 7; ++ky) {
                const int yy = yy0 + ky - 3;
                if (yy < 0 || yy >= p.H) continue;
                for (int kx = 0; kx < 7; ++kx) {
                    const int xx = xx0 + kx - 3;
                    if (xx < 0 || xx >= p.W) continue;
                    const float2 v = staged ? qs[(yy - y_lo) * p.W + xx]
                                            : __ldcg(reinterpret_cast<const float2*>(p.q + (((long long)n * p.H + yy) * p.W + xx) * 2));
                    e += w7s[ky * 7 + kx] * v.x + w7s[49 + ky * 7 + kx] * v.y;
                }
            }
            const float mv = 1.f / (1.f + __expf(-e));
            ms[pp] = mv;
            p.m_out[(long long)n * p.P + pl] = mv;
        }
        __syncthreads();
    }

    // ---- phase 5: z = W (m * s * x) + b + t, 64 pixels x 64 channels per pass, 4x4 register tile per thread ----------
    const int tp = tid >> 4, tc = tid & 15;
    const float4 bv = make_float4(p.bias[tc * 4], p.bias[tc * 4 + 1], p.bias[tc * 4 + 2], p.bias[tc * 4 + 3]);
    T* z16 = reinterpret_cast<T*>(p.z16);
    for (int base = 0; base < np; base += 64) {
        __syncthreads();
        {
            const int pl = tid >> 2, cb = (tid & 3) * 4;
            const int pp = base + pl;
            const float mp = pp < np ? ms[pp] : 0.f;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int c = cb + 16 * jj;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (pp < np) load4<T>(xs + pp * LF_C + c, v);
                vs[c * 68 + pl] = v[0] * mp * s_s[c]; vs[(c + 1) * 68 + pl] = v[1] * mp * s_s[c + 1];
                vs[(c + 2) * 68 + pl] = v[2] * mp * s_s[c + 2]; vs[(c + 3) * 68 + pl] = v[3] * mp * s_s[c + 3];
            }
        }
        __syncthreads();
        float4 rres[4];                               // residual t of this thread's 4 pixels: in flight during the GEMM
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int pp = base + tp * 4 + i;
            rres[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pp < np) rres[i] = __ldcs(reinterpret_cast<const float4*>(p.t + ((long long)n * p.P + p0 + pp) * LF_C + tc * 4));
        }
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
        for (int ci = 0; ci < LF_C; ++ci) {
            const float4 a = *reinterpret_cast<const float4*>(&vs[ci * 68 + tp * 4]);
            const float4 w = *reinterpret_cast<const float4*>(&ws[ci * 68 + tc * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int pp = base + tp * 4 + i;
            if (pp >= np) continue;
            const long long pix = (long long)n * p.P + p0 + pp;
            const float4 r = rres[i];
            const float4 o = make_float4(acc[i][0] + bv.x + r.x, acc[i][1] + bv.y + r.y, acc[i][2] + bv.z + r.z, acc[i][3] + bv.w + r.w);
            *reinterpret_cast<float4*>(p.z32 + pix * LF_C + tc * 4) = o;
            if (z16) store4<T>(z16 + pix * LF_C + tc * 4, o.x, o.y, o.z, o.w);
        }
    }
}

static int g_lf_sms = 0;

// CTAs per image / pixels per CTA of the single-wave schedule; cpi = 0 when the shape does not fit one wave
static size_t lf_smem_bytes(int pp, int elem_bytes) {
    return (((size_t)pp * LF_C * elem_bytes + 15) & ~(size_t)15) +
           sizeof(float) * (2 * LF_C * 68 + ((pp + 3) & ~3) + LF_C + 2 * 8 * LF_C + 8 * LF_C + 2 * LF_C + 32);
}

void la_fused_plan(int N, int P, int elem_bytes, int* cpi, int* ppc, size_t* smem) {
    if (!g_lf_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_lf_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    *cpi = 0; *ppc = 0; *smem = 0;
    if (N < 1 || N > g_lf_sms) return;
    // all CTAs must be co-resident (they meet at barriers): up to two 256-thread CTAs per SM (111 registers, < 100 KB smem)
    for (int per_sm = 2; per_sm >= 1; --per_sm) {
        int c = g_lf_sms * per_sm / N;
        if (c > 16) c = 16;                                    // 3 * cpi * 64 floats of partials are staged in the 64 x 68 tile buffer
        if (c < 1) continue;
        int pp = (int)cdiv(P, c);
        c = (int)cdiv(P, pp);                                  // drop CTAs that would own no pixel
        const size_t bytes = lf_smem_bytes(pp, elem_bytes);
        if (bytes * per_sm > 200 * 1024) continue;
        int resident = 0;                                      // the barriers need every CTA on an SM at the same time
        if (elem_bytes == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, la_fused_fwd_kernel<__nv_bfloat16>, 256, bytes);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, la_fused_fwd_kernel<float>, 256, bytes);
        if (resident < per_sm || N * c > g_lf_sms * per_sm) continue;
        *cpi = c; *ppc = pp; *smem = bytes;
        return;
    }
}

// returns 1 when the fused kernel was launched, 0 when the caller must take the multi-kernel path, <0 on error
int la_fused_fwd(const void* x, int dtype, const float* t, const float* fc1, const float* fc2, const float* w7, const float* Wm,
                 const float* bias, int N, int H, int W, int Cr, float* z32, void* z16, float* s_out, float* m_out, float* avg_out,
                 float* max_out, int* pstar, float* q, unsigned char* cstar, float* ws, cudaStream_t st) {
    const int P = H * W;
    int cpi, ppc; size_t smem;
    la_fused_plan(N, P, dtype == SR_BF16 ? 2 : 4, &cpi, &ppc, &smem);
    if (!cpi || Cr > 32) return 0;
    LaFusedParams p;
    p.x = x; p.t = t; p.fc1 = fc1; p.fc2 = fc2; p.w7 = w7; p.Wm = Wm; p.bias = bias;
    p.N = N; p.H = H; p.W = W; p.P = P; p.Cr = Cr; p.cpi = cpi; p.ppc = ppc;
    p.z32 = z32; p.z16 = z16; p.s_out = s_out; p.m_out = m_out; p.avg_out = avg_out; p.max_out = max_out; p.pstar = pstar;
    p.q = q; p.cstar = cstar;
    p.psum = ws; p.pmax = p.psum + (size_t)N * cpi * LF_C; p.pidx = reinterpret_cast<int*>(p.pmax + (size_t)N * cpi * LF_C);
    p.bar = reinterpret_cast<unsigned int*>(p.pidx + (size_t)N * cpi * LF_C);
    cudaMemsetAsync(p.bar, 0, sizeof(unsigned int) * 2 * N, st);
    static bool attr[2] = {false, false};
    if (dtype == SR_BF16) {
        if (!attr[1]) { cudaFuncSetAttribute(la_fused_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr[1] = true; }
        la_fused_fwd_kernel<__nv_bfloat16><<<N * cpi, 256, smem, st>>>(p);
    } else {
        if (!attr[0]) { cudaFuncSetAttribute(la_fused_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr[0] = true; }
        la_fused_fwd_kernel<float><<<N * cpi, 256, smem, st>>>(p);
    }
    count_launch();
    int rc = check_launch("la_fused_fwd_kernel");
    return rc == SR_OK ? 1 : rc;
}

}  // namespace sr
