// Band kernels of the fused local-attention chain (bf16 mode, C = 64):   z = Conv1x1( SLAM( CLAM(x) ) ) + t
// (reference model/sradsgan.py:101-151 CLAM / SLAM, :258-262 / :307-311 the 1x1 conv, :274 / :323 the residual, :459 the
// dense-sampling sum `out_all += y`).
//
// The tile path of la_chain.cu needs 4 launches forward and 5 (+ a memset) backward per chain because the chain has two
// global dependencies: the CLAM pooling over all pixels of an image, and the 7x7 SLAM stencil over per-pixel channel
// statistics.  Here
//   * the POOLING arrives as per-(image, channel) partial sums / packed maxima that the PRODUCER of x already emitted — the
//     epilogue of conv2's tensor-core kernel (conv_halo.cu) or of the previous chain (this file) — so nothing has to wait
//     for a reduction kernel (la_pool_pack_kernel is the stand-alone producer for callers without one);
//   * the STENCIL is crossed by giving every block a BAND of R image rows and letting it recompute the cheap per-pixel
//     statistics (channel mean / max of s*x) of the 3 halo rows above and below;
// so the whole forward chain is ONE kernel:   partials -> avg / max -> MLP -> s;  q (band + halo);  m = sigmoid(conv7(q));
// z = W (m s x) + b + t on the tensor cores (mma.sync, hi + lo bf16 operand split = fp32-class accuracy);  z32, z16, the
// dense-sampling accumulator and the pooling partials of z for the next chain are written in the epilogue.
// Backward: la_bwd_apply_mma_kernel (la_chain.cu; dz -> g, dm, per-block dW / db partials) -> la_bwd_band_kernel (7x7 input
// and weight gradients, du, dx, per-band ds partials, the dW partial reduction, and — in the LAST band of each image to
// finish — the CLAM gate backward) -> la_fix_kernel: 3 launches, no memset, no side stream.
// Algorithmic bytes per pixel: forward x 128 + t 256 + z32 256 + z16 128 (+ 512 for the accumulator) = 768 B (1280 B);
// backward gz32 256 + gz16 128 + x 128 + dx 128 + dz 256 = 896 B.
#include <algorithm>

#include "la_common.cuh"

namespace sr {

// per (image, pixel slice): channel sums and packed (max, first arg-max) keys of x — stand-alone producer of the partials
__global__ void __launch_bounds__(256)
la_pool_pack_kernel(const __nv_bfloat16* __restrict__ x, int P, int S, float* __restrict__ psum, unsigned int* __restrict__ pkey) {
    const int n = blockIdx.y, sl = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (P + S - 1) / S;
    const int p0 = sl * per, p1 = min(P, p0 + per);
    float s0 = 0.f, s1 = 0.f;
    unsigned int k0 = 0u, k1 = 0u;
    const __nv_bfloat16* base = x + (long long)n * P * LA_C + lane * 2;
    for (int p = p0 + warp; p < p1; p += 8) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(base + (long long)p * LA_C);
        s0 += __low2float(v); s1 += __high2float(v);
        k0 = max(k0, (bf16_key(__low2bfloat16(v)) << 16) | (unsigned int)(0xFFFF - p));
        k1 = max(k1, (bf16_key(__high2bfloat16(v)) << 16) | (unsigned int)(0xFFFF - p));
    }
    __shared__ float sh_s[8][LA_C];
    __shared__ unsigned int sh_k[8][LA_C];
    sh_s[warp][lane * 2] = s0; sh_s[warp][lane * 2 + 1] = s1;
    sh_k[warp][lane * 2] = k0; sh_k[warp][lane * 2 + 1] = k1;
    __syncthreads();
    if (threadIdx.x < LA_C) {
        const int c = threadIdx.x;
        float s = 0.f; unsigned int k = 0u;
        for (int w = 0; w < 8; ++w) { s += sh_s[w][c]; k = max(k, sh_k[w][c]); }
        const long long o = ((long long)n * S + sl) * LA_C + c;
        psum[o] = s; pkey[o] = k;
    }
}

__global__ void __launch_bounds__(256, 3)
la_fwd_band_kernel(const LaBandFwd p) {
    pdl_trigger();
    pdl_wait();                    // (common.cuh) launched with the programmatic-serialization attribute
    extern __shared__ __align__(16) unsigned char lb_smem[];
    __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(lb_smem);            // [co][ci] hi
    __nv_bfloat16* Wl = Ws + LA_C * LA_LD;                                    //          lo
    __nv_bfloat16* Vs = Wl + LA_C * LA_LD;                                    // [pixel][ci] = m*s*x hi
    __nv_bfloat16* Vl = Vs + LA_C * LA_LD;                                    //                     lo
    float* qs = reinterpret_cast<float*>(Vl + LA_C * LA_LD);                  // [(R+6)][W][2], zero outside the image
    float* ms = qs + (size_t)(p.R + 6) * p.W * 2;                             // [R*W] (rounded up to 4)
    float* s_s = ms + (((size_t)p.R * p.W + 3) & ~(size_t)3);                 // [64]
    float* bias_s = s_s + LA_C;                                               // [64]
    float* w7s = bias_s + LA_C;                                               // [100]
    float* avg_s = w7s + 100;                                                 // [64]
    float* max_s = avg_s + LA_C;                                              // [64]
    float* hid = max_s + LA_C;                                                // [32]: relu(fc1 avg), relu(fc1 max)
    float* red_s = hid + 32;                                                  // [4][64]
    unsigned int* red_k = reinterpret_cast<unsigned int*>(red_s + 4 * LA_C);  // [4][64]

    const int t = threadIdx.x;
    const int n = blockIdx.x / p.bands, band = blockIdx.x - n * p.bands;
    const int y0 = band * p.R, rows = min(p.R, p.H - y0);
    const int W = p.W, H = p.H, P = p.P;
    const int npx = rows * W;                                                  // pixels of this band (contiguous in memory)
    const long long pix0 = (long long)n * P + (long long)y0 * W;

    // ---- operands that do not depend on the data ----
    la_stage_w_split(Ws, Wl, p.Wm, t);
    if (t < LA_C) bias_s[t] = p.bias[t];
    if (t < 98) w7s[t] = p.w7[t];

    // ---- CLAM: pooled statistics from the producer's partials (fixed order), MLP, gate ----
    {
        const int c = t & 63, part = t >> 6;
        float sum = 0.f; unsigned int key = 0u;
        for (int r = part; r < p.T; r += 4) {
            const long long o = ((long long)n * p.T + r) * LA_C + c;
            sum += p.psum[o];
            key = max(key, p.pkey[o]);
        }
        red_s[part * LA_C + c] = sum; red_k[part * LA_C + c] = key;
        __syncthreads();
        if (part == 0) {
            sum = (red_s[c] + red_s[LA_C + c]) + (red_s[2 * LA_C + c] + red_s[3 * LA_C + c]);
            key = max(max(red_k[c], red_k[LA_C + c]), max(red_k[2 * LA_C + c], red_k[3 * LA_C + c]));
            const float a = sum / (float)P, m = bf16_key_value(key >> 16);
            avg_s[c] = a; max_s[c] = m;
            if (band == 0) {
                p.avg_out[n * LA_C + c] = a; p.max_out[n * LA_C + c] = m;
                p.pstar[n * LA_C + c] = 0xFFFF - (int)(key & 0xFFFFu);
            }
        }
        __syncthreads();
        if (t < p.Cr) {
            float u = 0.f, v = 0.f;
            for (int k = 0; k < LA_C; ++k) { const float w = p.fc1[t * LA_C + k]; u += w * avg_s[k]; v += w * max_s[k]; }
            hid[t] = fmaxf(u, 0.f) + fmaxf(v, 0.f);
        }
        __syncthreads();
        if (part == 0) {
            float o = 0.f;
            for (int j = 0; j < p.Cr; ++j) o += p.fc2[c * p.Cr + j] * hid[j];
            const float s = 1.f / (1.f + __expf(-o));
            s_s[c] = s;
            if (band == 0) p.s_out[n * LA_C + c] = s;
        }
        __syncthreads();
    }

    // ---- SLAM statistics q = [mean_c, max_c] of u = s*x on the band and its 3-row halo (4 threads per pixel, 16 channels each) ----
    {
        const int qn = (rows + 6) * W;
        const int sub = t & 3;
        float sv[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) sv[k] = s_s[sub * 16 + k];
        for (int base = 0; base < qn; base += 128) {
            // two 64-pixel passes per iteration, all four 16-byte loads issued before any of them is consumed
            int pi_[2], yy_[2], xx_[2]; bool inside_[2], valid_[2];
            uint4 a_[2], b_[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                pi_[u] = base + u * 64 + (t >> 2);
                inside_[u] = pi_[u] < qn;
                const int qr = inside_[u] ? pi_[u] / W : 0;
                xx_[u] = pi_[u] - qr * W;
                yy_[u] = y0 - 3 + qr;
                valid_[u] = inside_[u] && yy_[u] >= 0 && yy_[u] < H;
                a_[u] = make_uint4(0u, 0u, 0u, 0u); b_[u] = a_[u];
                if (valid_[u]) {
                    const __nv_bfloat16* xp = p.x + ((long long)n * P + (long long)yy_[u] * W + xx_[u]) * LA_C + sub * 16;
                    a_[u] = *reinterpret_cast<const uint4*>(xp); b_[u] = *reinterpret_cast<const uint4*>(xp + 8);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int pi = pi_[u], yy = yy_[u], xx = xx_[u];
                const bool inside = inside_[u], valid = valid_[u];
                float sum = 0.f, mv = -INFINITY; int mi = 0;
                if (valid) {
                    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a_[u]);
                    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b_[u]);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const __nv_bfloat162 v = k < 4 ? a2[k] : b2[k - 4];
                        const float u0 = __low2float(v) * sv[2 * k], u1 = __high2float(v) * sv[2 * k + 1];
                        sum += u0 + u1;
                        if (u0 > mv) { mv = u0; mi = sub * 16 + 2 * k; }
                        if (u1 > mv) { mv = u1; mi = sub * 16 + 2 * k + 1; }
                    }
                }
#pragma unroll
                for (int o = 1; o <= 2; o <<= 1) {
                    sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    const float ov = __shfl_xor_sync(0xffffffffu, mv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
                    if (ov > mv || (ov == mv && oi < mi)) { mv = ov; mi = oi; }
                }
                if (inside && sub == 0) {
                    const float mean = valid ? sum * (1.f / LA_C) : 0.f, mx = valid ? mv : 0.f;
                    qs[pi * 2] = mean; qs[pi * 2 + 1] = mx;
                    const int qr = pi / W;
                    if (valid && qr >= 3 && qr < 3 + rows) {                   // a pixel of the band itself: saved for the backward
                        const long long pix = (long long)n * P + (long long)yy * W + xx;
                        *reinterpret_cast<float2*>(p.q + pix * 2) = make_float2(mean, mx);
                        p.cstar[pix] = (unsigned char)mi;
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- SLAM gate m = sigmoid(conv7x7(q)) on the band ----
    for (int idx = t; idx < npx; idx += 256) {
        const int yy = idx / W, xx = idx - yy * W;
        float e = 0.f;
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
            const float* row = qs + (size_t)(yy + ky) * W * 2;
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const int x2 = xx + kx - 3;
                if (x2 < 0 || x2 >= W) continue;
                const float2 v = *reinterpret_cast<const float2*>(row + x2 * 2);
                e += w7s[ky * 7 + kx] * v.x + w7s[49 + ky * 7 + kx] * v.y;
            }
        }
        const float m = 1.f / (1.f + __expf(-e));
        ms[idx] = m;
        p.m_out[pix0 + idx] = m;
    }
    __syncthreads();

    // ---- z = W (m s x) + b + t, 64 pixels at a time (warp = 16-pixel row tile mt x 32-channel half nh) ----
    const int warp = t >> 5, lane = t & 31, mt = warp & 3, nh = warp >> 2, g = lane >> 2, tq = lane & 3;
    const int a_row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, a_col = ((lane >> 4) & 1) * 8;
    const int b_row = (lane & 7) + ((lane >> 4) & 1) * 8, b_col = ((lane >> 3) & 1) * 8;
    float csum[8];
    unsigned int ckey[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { csum[k] = 0.f; ckey[k] = 0u; }
    const int tiles = (npx + 63) >> 6;
    for (int tile = 0; tile < tiles; ++tile) {
        const int l0 = tile * 64;
        if (tile) __syncthreads();
        {
            const int pl = t >> 2, cb = (t & 3) * 4;
            const int li = l0 + pl;
            if (li < npx) {
                const float mp = ms[li];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int c = cb + 16 * jj;
                    float v[4];
                    load4<__nv_bfloat16>(p.x + (pix0 + li) * LA_C + c, v);
                    st_split4(Vs + pl * LA_LD + c, Vl + pl * LA_LD + c, v[0] * mp * s_s[c], v[1] * mp * s_s[c + 1], v[2] * mp * s_s[c + 2], v[3] * mp * s_s[c + 3]);
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) st_split4(Vs + pl * LA_LD + cb + 16 * jj, Vl + pl * LA_LD + cb + 16 * jj, 0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncthreads();
        // the epilogue's operands (residual trunk, dense-sampling accumulator) are requested BEFORE the tensor-core products
        float2 tres[2][4];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int li = l0 + mt * 16 + g + rr * 8;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int co = nh * 32 + nt * 8 + 2 * tq;
                tres[rr][nt] = make_float2(0.f, 0.f);
                if (li < npx) tres[rr][nt] = *reinterpret_cast<const float2*>(p.t + (pix0 + li) * LA_C + co);
            }
        }
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
                uint32_t b[4], bl[4];
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
            const int li = l0 + mt * 16 + g + rr * 8;
            if (li >= npx) continue;
            const long long pix = pix0 + li;
            const unsigned int ptag = (unsigned int)(0xFFFF - (y0 * W + li));          // pixel index inside the image
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int co = nh * 32 + nt * 8 + 2 * tq;
                const float o0 = acc[nt][rr * 2] + bias_s[co] + tres[rr][nt].x, o1 = acc[nt][rr * 2 + 1] + bias_s[co + 1] + tres[rr][nt].y;
                *reinterpret_cast<float2*>(p.z32 + pix * LA_C + co) = make_float2(o0, o1);
                const __nv_bfloat162 o16 = __floats2bfloat162_rn(o0, o1);
                *reinterpret_cast<__nv_bfloat162*>(p.z16 + pix * LA_C + co) = o16;
                if (p.acc_out) {
                    const float2 a = *reinterpret_cast<const float2*>(p.acc_in + pix * LA_C + co);
                    *reinterpret_cast<float2*>(p.acc_out + pix * LA_C + co) = make_float2(a.x + o0, a.y + o1);
                }
                csum[nt * 2] += __low2float(o16); csum[nt * 2 + 1] += __high2float(o16);
                ckey[nt * 2] = max(ckey[nt * 2], (bf16_key(__low2bfloat16(o16)) << 16) | ptag);
                ckey[nt * 2 + 1] = max(ckey[nt * 2 + 1], (bf16_key(__high2bfloat16(o16)) << 16) | ptag);
            }
        }
    }
    // ---- pooling partials of z16 for the chain that consumes it: one row per band ----
    if (p.out_psum) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int o = 4; o <= 16; o <<= 1) {
                csum[k] += __shfl_xor_sync(0xffffffffu, csum[k], o);
                ckey[k] = max(ckey[k], __shfl_xor_sync(0xffffffffu, ckey[k], o));
            }
        }
        __syncthreads();                                   // red_s / red_k are free again
        if (g == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int co = nh * 32 + (k >> 1) * 8 + 2 * tq + (k & 1);
                red_s[mt * LA_C + co] = csum[k]; red_k[mt * LA_C + co] = ckey[k];
            }
        }
        __syncthreads();
        if (t < LA_C) {
            const long long o = ((long long)n * p.bands + band) * LA_C + t;
            p.out_psum[o] = (red_s[t] + red_s[LA_C + t]) + (red_s[2 * LA_C + t] + red_s[3 * LA_C + t]);
            p.out_pkey[o] = max(max(red_k[t], red_k[LA_C + t]), max(red_k[2 * LA_C + t], red_k[3 * LA_C + t]));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward, second kernel (after la_bwd_apply_mma_kernel produced g = m * W^T dz, dm and the per-block dW / db partials)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
la_bwd_band_kernel(const LaBandBwd p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float lbb_smem[];
    float* qs = lbb_smem;                                       // [(R+6)][W][2]
    float* des = qs + (size_t)(p.R + 6) * p.W * 2;              // [(R+6)][W]: de = dm * m * (1 - m), zero outside the image
    float* dqs = des + (size_t)(p.R + 6) * p.W;                 // [R*W][2]
    float* w7s = dqs + (size_t)p.R * p.W * 2;                   // [100]
    float* red = w7s + 100;                                     // [8][64]
    __shared__ int is_last;

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int n = blockIdx.x / p.bands, band = blockIdx.x - n * p.bands;
    const int y0 = band * p.R, rows = min(p.R, p.H - y0);
    const int W = p.W, H = p.H, P = p.P;
    const int npx = rows * W;
    const long long pix0 = (long long)n * P + (long long)y0 * W;

    if (t < 98) w7s[t] = p.w7[t];
    for (int i = t; i < (rows + 6) * W; i += 256) {
        const int qr = i / W, xx = i - qr * W;
        const int yy = y0 - 3 + qr;
        float2 qv = make_float2(0.f, 0.f); float de = 0.f;
        if (yy >= 0 && yy < H) {
            const long long pix = (long long)n * P + (long long)yy * W + xx;
            qv = *reinterpret_cast<const float2*>(p.q + pix * 2);
            const float mv = p.m[pix];
            de = p.dm[pix] * mv * (1.f - mv);
        }
        qs[i * 2] = qv.x; qs[i * 2 + 1] = qv.y; des[i] = de;
    }
    __syncthreads();

    // ---- dq[p][ch] = sum_taps w7[ch][tap] * de[p - off]  (input gradient of the 7x7) ----
    for (int idx = t; idx < npx; idx += 256) {
        const int yy = idx / W, xx = idx - yy * W;
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
            const float* row = des + (size_t)(yy + 6 - ky) * W;             // image row y - (ky - 3)  <->  halo row (yy + 3) - (ky - 3)
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const int x2 = xx - (kx - 3);
                if (x2 < 0 || x2 >= W) continue;
                const float de = row[x2];
                a += w7s[ky * 7 + kx] * de;
                b += w7s[49 + ky * 7 + kx] * de;
            }
        }
        dqs[idx * 2] = a; dqs[idx * 2 + 1] = b;
    }
    // ---- dw7[ch][ky][kx] += sum over the band's pixels of de[p] * q[p + off][ch]  (weight gradient of the 7x7) ----
    if (t >= 128 && t < 128 + 98) {
        const int f = t - 128, ch = f / 49, ky = (f % 49) / 7, kx = f % 7;
        float acc = 0.f;
        for (int yy = 0; yy < rows; ++yy) {
            const float* drow = des + (size_t)(yy + 3) * W;
            const float* qrow = qs + ((size_t)(yy + ky) * W) * 2 + ch;
            const int xlo = max(0, 3 - kx), xhi = min(W, W + 3 - kx);
            for (int xx = xlo; xx < xhi; ++xx) acc = fmaf(drow[xx], qrow[(xx + kx - 3) * 2], acc);
        }
        p.w7part[((long long)n * p.bands + band) * 98 + f] = acc;
    }
    __syncthreads();

    // ---- du = g + dq_avg / C + dq_max [c == c*];  dx = s * du;  ds partial = sum_p du * x   (warp per pixel, lane = 2 channels) ----
    {
        const float s0 = p.s[n * LA_C + lane * 2], s1 = p.s[n * LA_C + lane * 2 + 1];
        float a0 = 0.f, a1 = 0.f;
        for (int l0 = warp; l0 < npx; l0 += 32) {
            float2 gv[4]; __nv_bfloat162 xv[4]; int cs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int li = l0 + 8 * u;
                gv[u] = make_float2(0.f, 0.f); xv[u] = __floats2bfloat162_rn(0.f, 0.f); cs[u] = 0;
                if (li < npx) {
                    const long long pix = pix0 + li;
                    gv[u] = *reinterpret_cast<const float2*>(p.g + pix * LA_C + lane * 2);
                    xv[u] = *reinterpret_cast<const __nv_bfloat162*>(p.x + pix * LA_C + lane * 2);
                    cs[u] = p.cstar[pix];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int li = l0 + 8 * u;
                if (li >= npx) continue;
                const float dqa = dqs[li * 2] * (1.f / LA_C), dqm = dqs[li * 2 + 1];
                const float du0 = gv[u].x + dqa + (cs[u] == lane * 2 ? dqm : 0.f);
                const float du1 = gv[u].y + dqa + (cs[u] == lane * 2 + 1 ? dqm : 0.f);
                a0 += du0 * __low2float(xv[u]); a1 += du1 * __high2float(xv[u]);
                *reinterpret_cast<__nv_bfloat162*>(p.dx + (pix0 + li) * LA_C + lane * 2) = __floats2bfloat162_rn(s0 * du0, s1 * du1);
            }
        }
        red[warp * LA_C + lane * 2] = a0; red[warp * LA_C + lane * 2 + 1] = a1;
    }
    __syncthreads();
    if (t < LA_C) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w * LA_C + t];
        p.dspart[((long long)n * p.bands + band) * LA_C + t] = v;
    }

    // ---- the 1x1 weight / bias gradient: this block adds its share of the first kernel's per-block partials (fixed order) ----
    {
        const int total = LA_C * LA_C + LA_C;
        const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
        const int e1 = min(total, ((int)blockIdx.x + 1) * per);
        __syncthreads();                                    // `red` (512 floats) is free again
        for (int eb = blockIdx.x * per; eb < e1; eb += 16) {
            const int ei = t & 15, kl = t >> 4, e = eb + ei;      // 16 entries x 16 lanes over the partial rows
            float v0 = 0.f, v1 = 0.f;
            if (e < e1) {
                int k = kl;
                for (; k + 16 < p.nparts; k += 32) { v0 += p.wpart[(long long)k * total + e]; v1 += p.wpart[(long long)(k + 16) * total + e]; }
                if (k < p.nparts) v0 += p.wpart[(long long)k * total + e];
            }
            red[kl * 16 + ei] = v0 + v1;
            __syncthreads();
            if (t < 16 && eb + t < e1) {
                float v = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) v += red[j * 16 + t];
                const int ee = eb + t;
                if (ee < LA_C * LA_C) p.dW[ee] += v; else p.db[ee - LA_C * LA_C] += v;
            }
            __syncthreads();
        }
    }

    // ---- the LAST band of image n to arrive finishes the image: ds = sum of the band partials, CLAM gate backward ----
    __threadfence();
    __syncthreads();
    if (t == 0) is_last = (atomicAdd(p.tickets + n, 1) == p.bands - 1) ? 1 : 0;
    __syncthreads();
    if (is_last) {
        __threadfence();
        if (t < LA_C) {
            float v = 0.f;
            for (int b = 0; b < p.bands; ++b) v += __ldcg(p.dspart + ((long long)n * p.bands + b) * LA_C + t);
            p.ds[n * LA_C + t] = v;
        }
        if (t >= 128 && t < 128 + 98) {                     // the image's 7x7 weight gradient: band partials in a fixed order, one atomic per image
            float v = 0.f;
            for (int b = 0; b < p.bands; ++b) v += __ldcg(p.w7part + ((long long)n * p.bands + b) * 98 + (t - 128));
            atomicAdd(p.d_w7 + (t - 128), v);
        }
        if (t == 0) p.tickets[n] = 0;
        __threadfence();
        __syncthreads();
        la_gate_bwd_body(n, p.ds, p.s, p.avg, p.mx, p.fc1, p.fc2, p.Cr, p.d_fc1, p.d_fc2, p.da, p.dmx);
    }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static int g_lb_sms = 0;
static int lb_sms() {
    if (!g_lb_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_lb_sms, cudaDevAttrMultiProcessorCount, dev); }
    return g_lb_sms;
}

// rows per band.  A block is a short chain of dependent phases (pooled gate -> q -> 7x7 -> GEMM tiles), so the SMs want several
// blocks resident (LB_OCC) rather than one tall band each: cost = waves of (N * bands) blocks over sms * LB_OCC slots, times the
// work of a block — (R + 6) rows of cheap per-pixel statistics + R rows of GEMM / stencil work (weighted 4x).  16 x 54^2 maps:
// R = 2 -> 27 bands per image, 432 blocks = one wave at 3 blocks per SM (80 registers x 256 threads).
constexpr int LB_OCC = 3;
int la_band_rows(int N, int H, int W) {
    const int sms = lb_sms();
    int best = 0; long long best_cost = 0;
    for (int R = 1; R <= 16; ++R) {
        if (R > H && best) break;
        const int r = R > H ? H : R;
        if ((long long)r * W > 2048) break;
        const long long blocks = (long long)N * cdiv(H, r);
        const long long cost = cdiv(blocks, (long long)sms * LB_OCC) * ((long long)(r + 6) * W + 4LL * r * W);
        if (!best || cost < best_cost) { best = r; best_cost = cost; }
    }
    return best ? best : 1;
}

int la_band_count(int N, int H, int W) { return (int)cdiv(H, la_band_rows(N, H, W)); }

bool la_band_supported(int N, int H, int W) {
    return (long long)H * W <= 65535 && W <= 512 && (long long)la_band_rows(N, H, W) * W <= 2048;
}

int la_pool_pack(const void* x, int N, int P, int S, float* psum, unsigned int* pkey, cudaStream_t st) {
    la_pool_pack_kernel<<<dim3(S, N), 256, 0, st>>>((const __nv_bfloat16*)x, P, S, psum, pkey);
    count_launch();
    return check_launch("la_pool_pack");
}

static size_t la_band_fwd_smem(int R, int W) {
    return (size_t)4 * LA_C * LA_LD * sizeof(__nv_bfloat16) +
           sizeof(float) * ((size_t)(R + 6) * W * 2 + (((size_t)R * W + 3) & ~(size_t)3) + 2 * LA_C + 100 + 2 * LA_C + 32 + 8 * LA_C);
}

int la_band_fwd(LaBandFwd p, cudaStream_t st) {
    p.R = la_band_rows(p.N, p.H, p.W);
    p.bands = (int)cdiv(p.H, p.R);
    p.P = p.H * p.W;
    const size_t smem = la_band_fwd_smem(p.R, p.W);
    static size_t attr = 0;
    if (smem > attr) { cudaFuncSetAttribute(la_fwd_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 96 * 1024)); attr = std::max<size_t>(smem, 96 * 1024); }
    launch_pdl(la_fwd_band_kernel, dim3(p.N * p.bands), dim3(256), smem, st, option("SR_PDL", 0) != 0, p);
    count_launch();
    return check_launch("la_fwd_band_kernel");
}

int la_band_bwd(LaBandBwd p, cudaStream_t st) {
    p.R = la_band_rows(p.N, p.H, p.W);
    p.bands = (int)cdiv(p.H, p.R);
    p.P = p.H * p.W;
    const size_t smem = sizeof(float) * ((size_t)(p.R + 6) * p.W * 3 + (size_t)p.R * p.W * 2 + 100 + 8 * LA_C);
    static size_t attr = 0;
    if (smem > attr) { cudaFuncSetAttribute(la_bwd_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 96 * 1024)); attr = std::max<size_t>(smem, 96 * 1024); }
    launch_pdl(la_bwd_band_kernel, dim3(p.N * p.bands), dim3(256), smem, st, option("SR_PDL", 0) != 0, p);
    count_launch();
    return check_launch("la_bwd_band_kernel");
}

}  // namespace sr
