// Device helpers shared by the local-attention chain kernels (la_chain.cu: tile kernels, la_band.cu: band kernels).
#pragma once
#include "common.cuh"

namespace sr {

constexpr int LA_C = 64;
constexpr int LA_LD = LA_C + 8;

__device__ __forceinline__ uint32_t la_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(la_smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(la_smem_u32(p)));
}
// D(16x8, f32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* p, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&lo); v.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = v;
}

// v -> (hi, lo) bf16 pair with hi + lo = v to 16 mantissa bits: operands that are fp32 in the reference chain (the trunk
// gradient, the gate products, the 1x1 weights) enter the tensor-core products as hi*hi + lo*hi + hi*lo, so the chain keeps
// fp32-class accuracy (the dropped lo*lo term is 2^-18 relative) at three MMAs per product.
__device__ __forceinline__ void st_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, float a, float b, float c, float d) {
    const float ah = __bfloat162float(__float2bfloat16_rn(a)), bh = __bfloat162float(__float2bfloat16_rn(b));
    const float ch = __bfloat162float(__float2bfloat16_rn(c)), dh = __bfloat162float(__float2bfloat16_rn(d));
    st_bf16x4(hi, ah, bh, ch, dh);
    st_bf16x4(lo, a - ah, b - bh, c - ch, d - dh);
}
__device__ __forceinline__ void st_split1(__nv_bfloat16* hi, __nv_bfloat16* lo, float a) {
    const __nv_bfloat16 h = __float2bfloat16_rn(a);
    *hi = h;
    *lo = __float2bfloat16_rn(a - __bfloat162float(h));
}

// gate backward (tiny, per image): d(sigmoid) -> MLP backward; weight gradients via fp32 atomics.  Called by ALL threads of a
// block (block-uniform), the first LA_C of which do the work.  `ds` is read through L2 (it was produced by other blocks' atomics).
__device__ __forceinline__ void la_gate_bwd_body(int n, const float* __restrict__ ds, const float* __restrict__ s, const float* __restrict__ avg,
                                                 const float* __restrict__ mx, const float* __restrict__ fc1, const float* __restrict__ fc2, int Cr,
                                                 float* __restrict__ d_fc1, float* __restrict__ d_fc2, float* __restrict__ da, float* __restrict__ dmx) {
    const int c = threadIdx.x;
    __shared__ float a[LA_C], m[LA_C], dov[LA_C], pa[16], pm[16], dha[16], dhm[16];
    if (c < LA_C) {
        a[c] = avg[n * LA_C + c]; m[c] = mx[n * LA_C + c];
        const float sv = s[n * LA_C + c];
        dov[c] = __ldcg(ds + n * LA_C + c) * sv * (1.f - sv);
    }
    __syncthreads();
    if (c < Cr) {
        float u = 0.f, v = 0.f, d = 0.f;
        for (int k = 0; k < LA_C; ++k) { u += fc1[c * LA_C + k] * a[k]; v += fc1[c * LA_C + k] * m[k]; d += fc2[k * Cr + c] * dov[k]; }
        pa[c] = u; pm[c] = v;
        dha[c] = u > 0.f ? d : 0.f;
        dhm[c] = v > 0.f ? d : 0.f;
    }
    __syncthreads();
    if (c < LA_C) {
        float dav = 0.f, dmv = 0.f;
        for (int j = 0; j < Cr; ++j) {
            atomicAdd(d_fc2 + c * Cr + j, dov[c] * (fmaxf(pa[j], 0.f) + fmaxf(pm[j], 0.f)));
            atomicAdd(d_fc1 + j * LA_C + c, dha[j] * a[c] + dhm[j] * m[c]);
            dav += fc1[j * LA_C + c] * dha[j];
            dmv += fc1[j * LA_C + c] * dhm[j];
        }
        da[n * LA_C + c] = dav; dmx[n * LA_C + c] = dmv;
    }
}


// W [co][ci] (fp32, 16-byte aligned) -> hi / lo bf16 rows of pitch LA_LD in shared memory: four independent 16-byte loads per
// thread issued together (a scalar loop here costs 16 dependent L2 round trips at the start of every block)
__device__ __forceinline__ void la_stage_w_split(__nv_bfloat16* Ws, __nv_bfloat16* Wl, const float* __restrict__ Wm, int t) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const float4*>(Wm + (t + k * 256) * 4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = (t + k * 256) * 4, row = e >> 6, col = e & 63;
        st_split4(Ws + row * LA_LD + col, Wl + row * LA_LD + col, v[k].x, v[k].y, v[k].z, v[k].w);
    }
}

// parameter blocks of the band kernels (la_band.cu), filled by la_chain.cu
struct LaBandFwd {
    const __nv_bfloat16* x; const float* t; const float* acc_in;
    const float* fc1; const float* fc2; const float* w7; const float* Wm; const float* bias;
    const float* psum; const unsigned int* pkey; int T;          // pooling partials of x: [N][T][64]
    int N, H, W, P, Cr, R, bands;
    float* z32; __nv_bfloat16* z16; float* acc_out;
    float* s_out; float* m_out; float* avg_out; float* max_out; int* pstar; float* q; unsigned char* cstar;
    float* out_psum; unsigned int* out_pkey;                     // pooling partials of z16 for the next chain: [N][bands][64]
};

struct LaBandBwd {
    const float* g; const float* dm; const float* m; const float* q; const unsigned char* cstar;
    const __nv_bfloat16* x; const float* s; const float* avg; const float* mx; const float* fc1; const float* fc2; const float* w7;
    const float* wpart; int nparts;                    // [nparts][64*64 + 64] partial dW | db of the first kernel
    int N, H, W, P, Cr, R, bands;
    __nv_bfloat16* dx; float* d_w7; float* dW; float* db; float* d_fc1; float* d_fc2;
    float* dspart;                                     // [N][bands][64]
    float* w7part;                                     // [N][bands][98] per-band 7x7 weight-gradient partials
    float* ds; float* da; float* dmx;                  // [N][64] each
    int* tickets;                                      // [N], zero when the kernel starts; re-armed by the last band of each image
};

// orderable 16-bit key of a bf16 value (larger value <-> larger key) and its inverse: per-(image, channel) maxima with their
// first arg-max pixel travel as ONE 32-bit word  (key << 16) | (0xFFFF - pixel)  that plain unsigned max() combines
__device__ __forceinline__ unsigned int bf16_key(__nv_bfloat16 v) {
    const unsigned int b = __bfloat16_as_ushort(v);
    return (b & 0x8000u) ? (~b & 0xFFFFu) : (b | 0x8000u);
}
__device__ __forceinline__ float bf16_key_value(unsigned int k) {
    const unsigned int b = (k & 0x8000u) ? (k & 0x7FFFu) : (~k & 0xFFFFu);
    return __bfloat162float(__ushort_as_bfloat16((unsigned short)b));
}

}  // namespace sr
