// Flash-style position attention (SGAM, reference model/sradsgan.py:153-176) for sm_100a: q, k in R^8, v in R^64,
// N = H*W tokens.  The reference materialises the N x N energy and softmax (2 x 544 MB at 54^2 x 16, 2 x 8.7 GB per
// image at 108^2); here nothing N x N ever leaves the SM:
//   * pass 1 (sgam_stats_kernel, SIMT fp32): per query row max m and 1/sum of exp — the d_k = 8 logits are cheap to
//     recompute, so the softmax is NORMALISED BEFORE the value product and no online rescaling of the accumulator is needed;
//   * sgam_pv_kernel (tcgen05): out[r] = sum_c w(r,c) vals[c], w(r,c) = exp(a_r.b_c - m_row[r] - m_col[c]) s_row[r] s_col[c].
//     Four producer warps compute 128 x 64 weight blocks in fp32, round them to bf16 and write them straight into a
//     128B-swizzled K-major shared-memory tile (the A operand); the value block [64 tokens][64 ch] arrives by TMA and is
//     consumed as an MN-major B operand exactly as it lies in the NHWC tensor; fp32 accumulation in TMEM.
//     Forward: rows = queries (m_row = m, s_row = 1/l), vals = v, epilogue y = gamma*acc + x.
//     Backward dV: rows = keys, columns = queries (m_col = m, s_col = 1/l), vals = dO.
//   * sgam_ds_kernel (tcgen05 + SIMT): dP block = rowvals . colvals^T on the tensor core (fresh TMEM accumulator per block,
//     double buffered), read back with tcgen05.ld; dS = P (dP - D); out8[r] += dS b_c.  rows = queries gives dQ, rows = keys
//     (with the per-query statistics on the columns) gives dK — no cross-thread reduction in either.
// Logits, softmax statistics and all accumulations are fp32; only the MMA operands (weights, v, dO) are bf16.
#include "tc_common.cuh"

namespace sr {

constexpr int SG_DK = 8, SG_ROWS = 128, SG_COLS = 64;
constexpr float SG_LOG2E = 1.4426950408889634f;

template <typename T> __device__ __forceinline__ void sg_load8(const T* p, float (&o)[8]);
template <> __device__ __forceinline__ void sg_load8<float>(const float* p, float (&o)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <> __device__ __forceinline__ void sg_load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[2 * i] = __low2float(h[i]); o[2 * i + 1] = __high2float(h[i]); }
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// ------------------------------------------------------------------------------------------------
// pass 1: softmax row statistics.  grid (ceil(P/128), N), block 128, thread = query row.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
sgam_stats_kernel(const T* __restrict__ a, const T* __restrict__ b, int P, float* __restrict__ m_out, float* __restrict__ linv_out) {
    __shared__ __align__(16) float bs[256][SG_DK];
    const int n = blockIdx.y, r = blockIdx.x * 128 + threadIdx.x;
    float av[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (r < P) sg_load8<T>(a + ((long long)n * P + r) * SG_DK, av);
#pragma unroll
    for (int j = 0; j < 8; ++j) av[j] *= SG_LOG2E;             // work in base 2: exp(x) = exp2(x log2 e)
    float m = -INFINITY, l = 0.f;
    for (int c0 = 0; c0 < P; c0 += 256) {
        __syncthreads();
        for (int i = threadIdx.x; i < 256; i += 128) {
            float bv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (c0 + i < P) sg_load8<T>(b + ((long long)n * P + c0 + i) * SG_DK, bv);
#pragma unroll
            for (int j = 0; j < 8; ++j) bs[i][j] = bv[j];
        }
        __syncthreads();
        const int nc = min(256, P - c0);
        for (int cb = 0; cb < nc; cb += 16) {                  // 16 logits in registers: one dot product per column
            float sv[16];
            float cm = -INFINITY;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float4 b0 = *reinterpret_cast<const float4*>(&bs[cb + c][0]), b1 = *reinterpret_cast<const float4*>(&bs[cb + c][4]);
                float sdot = av[0] * b0.x + av[1] * b0.y + av[2] * b0.z + av[3] * b0.w + av[4] * b1.x + av[5] * b1.y + av[6] * b1.z + av[7] * b1.w;
                if (cb + c >= nc) sdot = -INFINITY;
                sv[c] = sdot;
                cm = fmaxf(cm, sdot);
            }
            const float mn = fmaxf(m, cm);
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c) acc += exp2f(sv[c] - mn);
            l = l * exp2f(m - mn) + acc;
            m = mn;
        }
    }
    if (r < P) {
        m_out[(long long)n * P + r] = m / SG_LOG2E;            // natural-log units
        linv_out[(long long)n * P + r] = 1.f / l;
    }
}

// ------------------------------------------------------------------------------------------------
// shared pieces of the two tensor-core kernels
// ------------------------------------------------------------------------------------------------
struct SgCommon {
    const void* a; const void* b;          // [N][P][8]: row / column 8-vectors (bf16 or fp32)
    int ab_f32;
    const float* row_m; const float* row_s; const float* row_d;    // [N][P], nullable: subtract / multiply / "D" of the row
    const float* col_m; const float* col_s; const float* col_d;    // [N][P], nullable: the same for the column
    int N, P;
};

// column block `jb` of image n -> registers of the 128 staging threads: thread t carries 4 of the 64 x 8 vector
// components (column t / 2, components 4 * (t & 1) ..) and, for t < 64, the three column statistics
struct SgColRegs { float v[4]; float m, s, d; };

__device__ __forceinline__ void sg_fetch_cols(const SgCommon& p, int n, int jb, int t, SgColRegs& r) {
    const int c = jb * SG_COLS + (t >> 1), j0 = (t & 1) * 4;
    r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.f;
    r.m = 0.f; r.s = 0.f; r.d = 0.f;                        // s = 0 kills out-of-range columns
    if (c < p.P) {
        const long long o = ((long long)n * p.P + c) * SG_DK + j0;
        if (p.ab_f32) {
            const float4 f = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.b) + o);
            r.v[0] = f.x; r.v[1] = f.y; r.v[2] = f.z; r.v[3] = f.w;
        } else {
            const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.b) + o);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
            r.v[0] = __low2float(h[0]); r.v[1] = __high2float(h[0]); r.v[2] = __low2float(h[1]); r.v[3] = __high2float(h[1]);
        }
    }
    if (t < SG_COLS) {
        const int cc = jb * SG_COLS + t;
        if (cc < p.P) {
            const long long o = (long long)n * p.P + cc;
            r.m = p.col_m ? p.col_m[o] : 0.f;
            r.s = p.col_s ? p.col_s[o] : 1.f;
            r.d = p.col_d ? p.col_d[o] : 0.f;
        }
    }
}

// bs: [64][8] column vectors, cst: [64][4] (m * log2e, s, d, unused)
__device__ __forceinline__ void sg_store_cols(const SgColRegs& r, int t, float* bs, float* cst) {
    *reinterpret_cast<float4*>(bs + (t >> 1) * SG_DK + (t & 1) * 4) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    if (t < SG_COLS) *reinterpret_cast<float4*>(cst + t * 4) = make_float4(r.m * SG_LOG2E, r.s, r.d, 0.f);
}

__device__ __forceinline__ void sg_load_row(const SgCommon& p, int n, int row, bool valid, float (&av)[8], float& rm, float& rs, float& rd) {
#pragma unroll
    for (int j = 0; j < 8; ++j) av[j] = 0.f;
    rm = 0.f; rs = 0.f; rd = 0.f;
    if (valid) {
        const long long o = (long long)n * p.P + row;
        if (p.ab_f32) sg_load8<float>(reinterpret_cast<const float*>(p.a) + o * SG_DK, av);
        else sg_load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(p.a) + o * SG_DK, av);
        rm = (p.row_m ? p.row_m[o] : 0.f) * SG_LOG2E;
        rs = p.row_s ? p.row_s[o] : 1.f;
        rd = p.row_d ? p.row_d[o] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) av[j] *= SG_LOG2E;
}

// A (K-major, +32 B per K step) x B (MN-major, +16 rows = 2048 B per K step): four K = 16 steps
__device__ __forceinline__ void umma_f16_x4_kmn(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, t;\n\t.reg .b64 a1, b1, a2, b2, a3, b3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, 128;\n\t"
        "add.s64 a2, %1, 4;\n\tadd.s64 b2, %2, 256;\n\t"
        "add.s64 a3, %1, 6;\n\tadd.s64 b3, %2, 384;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ------------------------------------------------------------------------------------------------
// out[r][0..63] = sum_c w(r,c) vals[c][0..63]
// ------------------------------------------------------------------------------------------------
struct SgPvParams {
    SgCommon c;
    __nv_bfloat16* o16;                 // [N][P][64] = acc (nullable)
    float* y32; const float* resid32; const float* gamma;     // y32 = gamma[0] * acc + resid32 (nullable)
};

constexpr int SG_THREADS = 320;          // warp 0 TMA, warp 1 MMA, warps 2..9: two per 32-row quarter, one 32-column half each
constexpr int SG_B_STAGES = 4;

__global__ void __launch_bounds__(SG_THREADS, 1)
sgam_pv_kernel(const __grid_constant__ CUtensorMap map_v, const SgPvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_tiles = smem;                                   // 2 x [128 rows][64 weights] bf16, K-major SW128
    uint8_t* b_tiles = smem + 2 * 16384;                       // SG_B_STAGES x [64 tokens][64 ch] bf16 (TMA, MN-major operand)
    float* bs = reinterpret_cast<float*>(b_tiles + SG_B_STAGES * 8192);   // 2 x [64][8]
    float* cst = bs + 2 * SG_COLS * SG_DK;                     // 2 x [64][4]
    uint64_t* bars = reinterpret_cast<uint64_t*>(cst + 2 * SG_COLS * 4);
    uint64_t* a_full = bars;            // [2] count 128
    uint64_t* a_empty = bars + 2;       // [2]
    uint64_t* b_full = bars + 4;        // [SG_B_STAGES]
    uint64_t* b_empty = b_full + SG_B_STAGES;
    uint64_t* acc_full = b_empty + SG_B_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.y, row0 = blockIdx.x * SG_ROWS;
    const int nblk = (p.c.P + SG_COLS - 1) / SG_COLS;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_v);
        for (int i = 0; i < 2; ++i) { mbar_init(a_full + i, 256); mbar_init(a_empty + i, 1); }
        for (int i = 0; i < SG_B_STAGES; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int jb = 0; jb < nblk; ++jb) {
                mbar_wait(b_empty + st, ph ^ 1);
                mbar_expect_tx(b_full + st, 8192);
                tma_load_2d(b_tiles + st * 8192, &map_v, b_full + st, 0, n * p.c.P + jb * SG_COLS);
                if (++st == SG_B_STAGES) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // D = f32, A = bf16 K-major, B = bf16 MN-major (bit 16), N = 64, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(SG_COLS >> 3) << 17) | ((128u >> 4) << 24);
        int st = 0; uint32_t ph = 0;
        for (int jb = 0; jb < nblk; ++jb) {
            const int buf = jb & 1;
            mbar_wait(a_full + buf, (uint32_t)((jb >> 1) & 1));
            mbar_wait(b_full + st, ph);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(a_tiles + buf * 16384));
                const uint64_t bdesc = make_mnmajor_sw128_desc(smem_u32(b_tiles + st * 8192), 8192);
                umma_f16_x4_kmn(tmem_base, adesc, bdesc, idesc, jb ? 1u : 0u);
                umma_commit(a_empty + buf);
                umma_commit(b_empty + st);
                if (jb == nblk - 1) umma_commit(acc_full);
            }
            __syncwarp();
            if (++st == SG_B_STAGES) { st = 0; ph ^= 1; }
        }
    } else {
        const int t = threadIdx.x - 64;                        // 0..255; threads 0..127 also stage the column blocks
        const int half = (warp - 2) >> 2;                      // which 32 columns of each 64-column block
        const int rl = (warp & 3) * 32 + lane;                 // row of the tile = TMEM lane
        const int row = row0 + rl;
        const bool valid = row < p.c.P;
        float av[8], rm, rs, rd;
        sg_load_row(p.c, n, row, valid, av, rm, rs, rd);
        SgColRegs cr;
        if (t < 128) { sg_fetch_cols(p.c, n, 0, t, cr); sg_store_cols(cr, t, bs, cst); }
        named_bar_sync(1, 256);
        for (int jb = 0; jb < nblk; ++jb) {
            const int buf = jb & 1;
            if (jb + 1 < nblk && t < 128) sg_fetch_cols(p.c, n, jb + 1, t, cr);
            mbar_wait(a_empty + buf, (uint32_t)(((jb >> 1) & 1) ^ 1));
            const float* bb = bs + buf * SG_COLS * SG_DK + half * 32 * SG_DK;
            const float* cc = cst + buf * SG_COLS * 4 + half * 32 * 4;
            uint32_t wp[16];                                   // 32 weights as bf16 pairs
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                float w2[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float4 b0 = *reinterpret_cast<const float4*>(bb + (c + h) * SG_DK), b1 = *reinterpret_cast<const float4*>(bb + (c + h) * SG_DK + 4);
                    const float4 cs = *reinterpret_cast<const float4*>(cc + (c + h) * 4);
                    const float s = av[0] * b0.x + av[1] * b0.y + av[2] * b0.z + av[3] * b0.w + av[4] * b1.x + av[5] * b1.y + av[6] * b1.z + av[7] * b1.w;
                    w2[h] = exp2f(fminf(s - rm - cs.x, 0.f)) * (rs * cs.y);      // exponent <= 0 by construction; the clamp keeps padding finite
                }
                const __nv_bfloat162 pk = __floats2bfloat162_rn(w2[0], w2[1]);
                wp[c >> 1] = *reinterpret_cast<const uint32_t*>(&pk);
            }
            uint8_t* arow = a_tiles + buf * 16384 + rl * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j)                         // 16-byte chunk (4*half + j) of the row lands at chunk ^ (row & 7)
                *reinterpret_cast<uint4*>(arow + (((half * 4 + j) ^ (rl & 7)) << 4)) = make_uint4(wp[4 * j], wp[4 * j + 1], wp[4 * j + 2], wp[4 * j + 3]);
            fence_proxy_async();
            mbar_arrive(a_full + buf);
            if (jb + 1 < nblk && t < 128) sg_store_cols(cr, t, bs + (buf ^ 1) * SG_COLS * SG_DK, cst + (buf ^ 1) * SG_COLS * 4);
            named_bar_sync(1, 256);
        }
        // epilogue
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const uint32_t t_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const float gm = p.gamma ? p.gamma[0] : 1.f;
#pragma unroll
        for (int c0 = half * 32; c0 < half * 32 + 32; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(t_base + (uint32_t)c0, v);
            if (valid) {
                const long long o = ((long long)n * p.c.P + row) * 64 + c0;
                if (p.o16) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 w;
                        __nv_bfloat162 b0 = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
                        __nv_bfloat162 b1 = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
                        __nv_bfloat162 b3 = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
                        w.x = *reinterpret_cast<uint32_t*>(&b0); w.y = *reinterpret_cast<uint32_t*>(&b1);
                        w.z = *reinterpret_cast<uint32_t*>(&b2); w.w = *reinterpret_cast<uint32_t*>(&b3);
                        reinterpret_cast<uint4*>(p.o16 + o)[g] = w;
                    }
                }
                if (p.y32) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float4 r = reinterpret_cast<const float4*>(p.resid32 + o)[g];
                        reinterpret_cast<float4*>(p.y32 + o)[g] = make_float4(gm * __uint_as_float(v[g * 4]) + r.x, gm * __uint_as_float(v[g * 4 + 1]) + r.y,
                                                                               gm * __uint_as_float(v[g * 4 + 2]) + r.z, gm * __uint_as_float(v[g * 4 + 3]) + r.w);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------------------------
// out8[r] = sum_c P(r,c) (dP(r,c) - D) b_c,   dP = rowvals[r] . colvals[c]
// ------------------------------------------------------------------------------------------------
struct SgDsParams {
    SgCommon c;
    float* out8;                        // [N][P][8]
};

__global__ void __launch_bounds__(SG_THREADS, 1)
sgam_ds_kernel(const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_c, const SgDsParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* r_tile = smem;                                    // [128 rows][64 ch] bf16 K-major (TMA)
    uint8_t* c_tiles = smem + 16384;                           // SG_B_STAGES x [64 tokens][64 ch] bf16 K-major (TMA)
    float* bs = reinterpret_cast<float*>(c_tiles + SG_B_STAGES * 8192);
    float* cst = bs + 2 * SG_COLS * SG_DK;
    uint64_t* bars = reinterpret_cast<uint64_t*>(cst + 2 * SG_COLS * 4);
    uint64_t* r_full = bars;            // [1]
    uint64_t* c_full = bars + 1;        // [SG_B_STAGES]
    uint64_t* c_empty = c_full + SG_B_STAGES;
    uint64_t* acc_full = c_empty + SG_B_STAGES;   // [2]
    uint64_t* acc_empty = acc_full + 2;           // [2] count 4 (one arrive per consumer warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.y, row0 = blockIdx.x * SG_ROWS;
    const int nblk = (p.c.P + SG_COLS - 1) / SG_COLS;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_r);
        tma_prefetch_desc(&map_c);
        mbar_init(r_full, 1);
        for (int i = 0; i < SG_B_STAGES; ++i) { mbar_init(c_full + i, 1); mbar_init(c_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(r_full, 16384);
            tma_load_2d(r_tile, &map_r, r_full, 0, n * p.c.P + row0);
            int st = 0; uint32_t ph = 0;
            for (int jb = 0; jb < nblk; ++jb) {
                mbar_wait(c_empty + st, ph ^ 1);
                mbar_expect_tx(c_full + st, 8192);
                tma_load_2d(c_tiles + st * 8192, &map_c, c_full + st, 0, n * p.c.P + jb * SG_COLS);
                if (++st == SG_B_STAGES) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // D = f32, A, B = bf16 K-major, N = 64, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(SG_COLS >> 3) << 17) | ((128u >> 4) << 24);
        mbar_wait(r_full, 0);
        int st = 0; uint32_t ph = 0;
        for (int jb = 0; jb < nblk; ++jb) {
            const int buf = jb & 1;
            mbar_wait(acc_empty + buf, (uint32_t)(((jb >> 1) & 1) ^ 1));
            mbar_wait(c_full + st, ph);
            tc_fence_after();
            if (lane == 0) {
                umma_f16_x4(tmem_base + (uint32_t)(buf * 64), make_kmajor_sw128_desc(smem_u32(r_tile)),
                            make_kmajor_sw128_desc(smem_u32(c_tiles + st * 8192)), idesc, 0u);
                umma_commit(c_empty + st);
                umma_commit(acc_full + buf);
            }
            __syncwarp();
            if (++st == SG_B_STAGES) { st = 0; ph ^= 1; }
        }
    } else {
        const int t = threadIdx.x - 64;                        // 0..255; threads 0..127 also stage the column blocks
        const int half = (warp - 2) >> 2;                      // which 32 columns of each 64-column block
        const int rl = (warp & 3) * 32 + lane;
        const int row = row0 + rl;
        const bool valid = row < p.c.P;
        float av[8], rm, rs, rd;
        sg_load_row(p.c, n, row, valid, av, rm, rs, rd);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        SgColRegs cr;
        if (t < 128) { sg_fetch_cols(p.c, n, 0, t, cr); sg_store_cols(cr, t, bs, cst); }
        named_bar_sync(1, 256);
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (int jb = 0; jb < nblk; ++jb) {
            const int buf = jb & 1;
            if (jb + 1 < nblk && t < 128) sg_fetch_cols(p.c, n, jb + 1, t, cr);
            mbar_wait(acc_full + buf, (uint32_t)((jb >> 1) & 1));
            tc_fence_after();
            const float* bb = bs + buf * SG_COLS * SG_DK + half * 32 * SG_DK;
            const float* cc = cst + buf * SG_COLS * 4 + half * 32 * 4;
            uint32_t v[32];
            tmem_ld32(t_lane + (uint32_t)(buf * 64 + half * 32), v);
            tc_fence_before();                                  // this warp's slice of the accumulator is in registers
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + buf);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float4 b0 = *reinterpret_cast<const float4*>(bb + c * SG_DK), b1 = *reinterpret_cast<const float4*>(bb + c * SG_DK + 4);
                const float4 cs = *reinterpret_cast<const float4*>(cc + c * 4);
                const float s = av[0] * b0.x + av[1] * b0.y + av[2] * b0.z + av[3] * b0.w + av[4] * b1.x + av[5] * b1.y + av[6] * b1.z + av[7] * b1.w;
                const float pw = exp2f(fminf(s - rm - cs.x, 0.f)) * (rs * cs.y);
                const float ds = pw * (__uint_as_float(v[c]) - rd - cs.z);
                acc[0] = fmaf(ds, b0.x, acc[0]); acc[1] = fmaf(ds, b0.y, acc[1]); acc[2] = fmaf(ds, b0.z, acc[2]); acc[3] = fmaf(ds, b0.w, acc[3]);
                acc[4] = fmaf(ds, b1.x, acc[4]); acc[5] = fmaf(ds, b1.y, acc[5]); acc[6] = fmaf(ds, b1.z, acc[6]); acc[7] = fmaf(ds, b1.w, acc[7]);
            }
            if (jb + 1 < nblk && t < 128) sg_store_cols(cr, t, bs + (buf ^ 1) * SG_COLS * SG_DK, cst + (buf ^ 1) * SG_COLS * 4);
            named_bar_sync(1, 256);
        }
        // the two column halves of a row meet in shared memory (the row tile is no longer needed)
        float* part = reinterpret_cast<float*>(r_tile);
        if (half == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) part[rl * 8 + j] = acc[j];
        }
        named_bar_sync(1, 256);
        if (half == 0 && valid) {
            float* o = p.out8 + ((long long)n * p.c.P + row) * SG_DK;
            reinterpret_cast<float4*>(o)[0] = make_float4(acc[0] + part[rl * 8], acc[1] + part[rl * 8 + 1], acc[2] + part[rl * 8 + 2], acc[3] + part[rl * 8 + 3]);
            reinterpret_cast<float4*>(o)[1] = make_float4(acc[4] + part[rl * 8 + 4], acc[5] + part[rl * 8 + 5], acc[6] + part[rl * 8 + 6], acc[7] + part[rl * 8 + 7]);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------
// backward prologue: dO = gamma * dy (bf16), D[r] = sum_ch dO O, dgamma += sum dy O.   warp per token.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgam_bwd_prep_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ o16, const float* __restrict__ gamma, long long rows,
                     __nv_bfloat16* __restrict__ do16, float* __restrict__ d_out, float* __restrict__ dgamma) {
    const float gm = gamma[0];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float gsum = 0.f;
    for (long long r = (long long)blockIdx.x * 8 + wib; r < rows; r += (long long)gridDim.x * 8) {
        const float2 g = *reinterpret_cast<const float2*>(dy + r * 64 + lane * 2);
        const __nv_bfloat162 o = *reinterpret_cast<const __nv_bfloat162*>(o16 + r * 64 + lane * 2);
        const float o0 = __low2float(o), o1 = __high2float(o);
        const float dot = g.x * o0 + g.y * o1;
        const __nv_bfloat162 d2 = __floats2bfloat162_rn(gm * g.x, gm * g.y);
        *reinterpret_cast<__nv_bfloat162*>(do16 + r * 64 + lane * 2) = d2;
        // D uses the ROUNDED dO so that sum_c P (dP - D) = 0 holds for the operands the tensor core sees
        float dd = __low2float(d2) * o0 + __high2float(d2) * o1;
        dd = warp_sum(dd);
        gsum += dot;
        if (lane == 0) d_out[r] = dd;
    }
    gsum = warp_sum(gsum);
    __shared__ float red[8];
    if (lane == 0) red[wib] = gsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        atomicAdd(dgamma, s);
    }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static size_t sg_pv_smem() { return 1024 + 2 * 16384 + SG_B_STAGES * 8192 + sizeof(float) * (2 * SG_COLS * SG_DK + 2 * SG_COLS * 4) + 256; }
static size_t sg_ds_smem() { return 1024 + 16384 + SG_B_STAGES * 8192 + sizeof(float) * (2 * SG_COLS * SG_DK + 2 * SG_COLS * 4) + 256; }

int sgam_stats(const void* a, const void* b, int ab_dtype, int N, int P, float* m, float* linv, cudaStream_t st) {
    dim3 grid((unsigned)cdiv(P, 128), (unsigned)N);
    if (ab_dtype == SR_F32) sgam_stats_kernel<float><<<grid, 128, 0, st>>>((const float*)a, (const float*)b, P, m, linv);
    else sgam_stats_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, P, m, linv);
    count_launch();
    return check_launch("sgam_stats_kernel");
}

int sgam_pv(const SgCommon& c, const void* vals16, __nv_bfloat16* o16, float* y32, const float* resid32, const float* gamma, cudaStream_t st) {
    int rc = load_driver_fns();
    if (rc != SR_OK) return rc;
    alignas(64) CUtensorMap map_v;
    rc = make_tiled2d_map(&map_v, vals16, (uint64_t)c.N * c.P, 64, SG_COLS);
    if (rc != SR_OK) return rc;
    SgPvParams p;
    p.c = c; p.o16 = o16; p.y32 = y32; p.resid32 = resid32; p.gamma = gamma;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(sgam_pv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_pv_smem()); attr = true; }
    sgam_pv_kernel<<<dim3((unsigned)cdiv(c.P, SG_ROWS), (unsigned)c.N), SG_THREADS, sg_pv_smem(), st>>>(map_v, p);
    count_launch();
    return check_launch("sgam_pv_kernel");
}

int sgam_ds(const SgCommon& c, const void* rowvals16, const void* colvals16, float* out8, cudaStream_t st) {
    int rc = load_driver_fns();
    if (rc != SR_OK) return rc;
    alignas(64) CUtensorMap map_r, map_c;
    rc = make_tiled2d_map(&map_r, rowvals16, (uint64_t)c.N * c.P, 64, SG_ROWS);
    if (rc != SR_OK) return rc;
    rc = make_tiled2d_map(&map_c, colvals16, (uint64_t)c.N * c.P, 64, SG_COLS);
    if (rc != SR_OK) return rc;
    SgDsParams p;
    p.c = c; p.out8 = out8;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(sgam_ds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_ds_smem()); attr = true; }
    sgam_ds_kernel<<<dim3((unsigned)cdiv(c.P, SG_ROWS), (unsigned)c.N), SG_THREADS, sg_ds_smem(), st>>>(map_r, map_c, p);
    count_launch();
    return check_launch("sgam_ds_kernel");
}

int sgam_bwd_prep(const float* dy, const void* o16, const float* gamma, long long rows, void* do16, float* d_out, float* dgamma, cudaStream_t st) {
    const int blocks = (int)(rows / 8 < 148 * 8 ? cdiv(rows, 8) : 148 * 8);
    sgam_bwd_prep_kernel<<<blocks, 256, 0, st>>>(dy, (const __nv_bfloat16*)o16, gamma, rows, (__nv_bfloat16*)do16, d_out, dgamma);
    count_launch();
    return check_launch("sgam_bwd_prep_kernel");
}

}  // namespace sr
