/*
 * sradsgan_b200 — C ABI of the B200-native SRADSGAN hot path (libsradsgan_b200.so).
 *
 * The reference (Meng-333/SRADSGAN) is pure Python over torch ATen/cuDNN and has no FFI of its own
 * (SURVEY.md §2); this header therefore DEFINES the boundary a reference maintainer would bind with
 * ctypes.  Each entry point names the reference call sites whose arithmetic it
 * replaces (paths relative to SRADSGAN/ in the reference tree).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; the library never allocates, never
 *     synchronises and launches only on the caller's stream (`stream` = cudaStream_t as void*),
 *     so every call is CUDA-graph capturable;
 *   - activations are NHWC ("channels_last"), dtype SR_F32 or SR_BF16; parameters/gradients exchanged
 *     with the host framework stay in the reference's layout: OIHW fp32;
 *   - return value 0 = ok, <0 = error, text via sr_last_error() (thread local);
 *   - no CPU fallback: on a device that is not sm_100 every compute call fails with SR_ERR_ARCH.
 */
#ifndef SRADSGAN_B200_H
#define SRADSGAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SR_OK 0
#define SR_ERR_ARG (-1)
#define SR_ERR_CUDA (-2)
#define SR_ERR_ARCH (-3)
#define SR_ERR_UNSUPPORTED (-4)

enum { SR_F32 = 0, SR_BF16 = 1 };
enum { SR_ACT_NONE = 0, SR_ACT_LRELU = 1, SR_ACT_RELU = 2, SR_ACT_SIGMOID = 3 };
enum { SR_IMPL_AUTO = 0, SR_IMPL_SIMT = 1, SR_IMPL_TCGEN05 = 2 /* im2col-TMA kernel */, SR_IMPL_HALO = 3 /* halo-tile kernel */ };

/* Geometry of one convolution y = act(conv(x, w) + bias) [+ residual] [-> PixelShuffle(r)].
 * Replaces nn.Conv2d (+ nn.LeakyReLU / nn.ReLU / nn.PixelShuffle / `out += x`) at
 * model/sradsgan.py:222-223,251-253,274 (RAB), :332-336 (MSB), :375,:381-386 (GAB_UP), :427,:448,
 * :476-503 (Discriminator), :92-95 (VGG19[:12]). */
typedef struct sr_conv_desc {
    int32_t N, H, W, Cin;       /* input  NHWC */
    int32_t Ho, Wo, Cout;       /* output NHWC, before the optional pixel shuffle */
    int32_t kh, kw, stride, pad;
    int32_t in_dtype, out_dtype;/* SR_F32 / SR_BF16 (weights are packed in in_dtype) */
    int32_t act;                /* SR_ACT_* applied to conv+bias, before residual */
    float   slope;              /* LeakyReLU negative slope */
    int32_t shuffle_r;          /* 0/1: none; r>1: y is written as (N, Ho*r, Wo*r, Cout/r^2) */
    int32_t impl;               /* SR_IMPL_* */
    /* sr_conv2d_fwd only, nullable: the epilogue also emits the CLAM pooling partials of y (bf16) that the local-attention
     * chain behind conv2 of a RAB reads (model/sradsgan.py:253 -> :258, avg / max pooling of :117-121) —
     * [N][sr_conv_pool_rows(desc)][Cout] channel sums and packed keys (see sr_la_chain_args).  Needs a geometry for which
     * sr_conv_pool_rows() > 0. */
    float* pool_sum; uint32_t* pool_key;
} sr_conv_desc;

const char* sr_last_error(void);
int sr_version(void);
/* 0 when the current device is sm_100 (B200) and the tcgen05/TMA paths are usable. */
int sr_device_check(void);
/* number of kernels this library has launched since load (for bench.py's gpu_launches). */
int64_t sr_launch_count(void);

/* rows per image of the pooling partials sr_conv2d_fwd emits for this geometry when desc.pool_sum / pool_key are set
 * (0: not available — 3x3 / stride 1 / pad 1, bf16 in and out, Cin % 64 == 0, Cout == 64, no pixel shuffle, H*W <= 65535). */
int sr_conv_pool_rows(const sr_conv_desc* d);

/* 1 when sr_conv2d_fwd (kind 0) / sr_conv2d_dgrad (kind 1) / sr_conv2d_wgrad (kind 2) will run this
 * geometry on a tcgen05 kernel, 0 when it takes the SIMT kernel (bench.py attributes time per kernel). */
int sr_conv_uses_tcgen05(const sr_conv_desc* d, int kind);

/* OIHW fp32 master weights -> packed [kh*kw][Cout][Cin] (mode 0, forward/B-operand K-major) or
 * [kh*kw][Cin][Cout] (mode 1, dgrad) in `dtype`.  shuffle_r > 1 (mode 0, convs followed by
 * nn.PixelShuffle(r), model/sradsgan.py:381-386): rows are stored subpixel-major, row sub*(Cout/r^2)+c
 * = output channel c*r^2+sub, which is what sr_conv2d_fwd expects when desc.shuffle_r > 1. */
int sr_pack_weights(const float* w_oihw, void* packed, int Cout, int Cin, int kh, int kw,
                    int mode, int dtype, int shuffle_r, void* stream);

/* The same re-packing for MANY weights in one launch (all convolutions of a network after its optimiser step,
 * model/sradsgan.py:857-858,887).  table_dev: DEVICE array of n_entries records of 8 int64:
 * {w_oihw pointer, packed pointer, Cout, Cin, kh*kw, mode, shuffle_r, first block}, where entry e owns blocks
 * [first_block[e], first_block[e+1]) and needs ceil(Cout*Cin*kh*kw / 1024) of them; total_blocks = their sum. */
int sr_pack_weights_batched(const void* table_dev, int n_entries, int total_blocks, int dtype, void* stream);

/* y = act(conv(x,w)+bias) (+residual) ; w_packed from sr_pack_weights(mode 0); bias/residual may be NULL.
 * residual has y's layout and dtype. */
int sr_conv2d_fwd(const sr_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                  const void* residual, void* y, void* stream);
/* dx = conv_transpose(dy, w) for the forward conv described by d (act/shuffle ignored).
 * w_packed from sr_pack_weights(mode 1).  dy has dtype d->in_dtype, dx has d->out_dtype.
 * (autograd of every nn.Conv2d above; reference: torch.autograd, model/sradsgan.py:857,886,621) */
int sr_conv2d_dgrad(const sr_conv_desc* d, const void* dy, const void* w_packed_t, void* dx, void* stream);
/* dx = conv_transpose(dy, w) * act'(y_prev): the input gradient of the conv described by d, multiplied by the
 * derivative of the LeakyReLU / ReLU that PRODUCED this conv's input (y_prev = that activation's output, shaped like dx,
 * dtype d->out_dtype) — the mask is applied in the epilogue of the tensor-core kernel.  RAB backward: conv2.dgrad followed
 * by LeakyReLU(0.2).backward (model/sradsgan.py:251-253) in one kernel. */
int sr_conv2d_dgrad_act(const sr_conv_desc* d, const void* dy, const void* w_packed_t, const void* y_prev, int act, float slope,
                        void* dx, void* stream);
/* dw (OIHW fp32) (+)= sum_pixels dy (x) x ; dbias (fp32, may be NULL) (+)= sum_pixels dy.
 * accumulate=0 overwrites (the library zero-fills first), 1 adds (tied upsampler weights, GP double
 * backward).  x and dy have dtype d->in_dtype.
 * workspace (nullable, 16-byte aligned DEVICE scratch owned by the caller, >= sr_conv2d_wgrad_workspace_bytes(d), used by
 * this call only — one buffer per stream makes concurrent weight gradients re-entrant): with it the split-K partial tiles
 * are combined by plain stores + one reduce kernel instead of fp32 atomics (3x faster, and deterministic). */
size_t sr_conv2d_wgrad_workspace_bytes(const sr_conv_desc* d);
int sr_conv2d_wgrad(const sr_conv_desc* d, const void* x, const void* dy, float* dw_oihw, float* dbias,
                    int accumulate, void* workspace, uint64_t workspace_bytes, void* stream);

/* Tuning / experiment options by name (e.g. "SR_LA_BAND" 0: tile kernels for the local-attention chain; "SR_HALO_SPLITK",
 * "SR_WG_RMULT", ...; DESIGN.md §3 lists them).  The library never reads the environment: whoever wants a knob sets it here. */
int sr_set_option(const char* name, int value);

/* Up to 3 auxiliary streams with one fork event and one join event per stream, all created and owned by the CALLER (the
 * library creates no streams or events).  Calls whose work splits into independent launches — the four parity classes of a
 * stride-2 input gradient, the 7x7 weight gradient of the tile-path attention chain — fork them onto these streams and
 * re-join before returning, so every call stays stream-ordered for the caller and capturable.  n = 0 (default): everything
 * runs on the caller's stream.  The set is shared by all calls: issue such calls from one stream at a time. */
int sr_set_aux_streams(void* const* streams, void* fork_event, void* const* join_events, int n);

/* Fused local-attention tail of RAB / ResGroup (C = 64): z = Conv1x1(SLAM(CLAM(x))) + t, i.e.
 * nn modules CLAM (model/sradsgan.py:101-127), SLAM (:129-151), the 1x1 `conv` (:233/:297) and the
 * in-place residual `out += x` (:274/:323) in five kernels.  x: NHWC (x_dtype), t/z32: NHWC fp32,
 * z16 (nullable): copy of z in x_dtype for the next 3x3 conv.  fc1 [Cr][64], fc2 [64][Cr], w7 [2][7][7],
 * W [64][64] (OIHW), bias [64], all fp32.  Saved for backward: s, avg, max, pstar [N][64]; m [N][H*W];
 * q [N][H*W][2]; cstar [N][H*W] (u8).  workspace >= sr_la_chain_workspace_bytes(N,H,W). */
size_t sr_la_chain_workspace_bytes(int N, int H, int W);
int sr_la_chain_fwd(const void* x, int x_dtype, const float* t, const float* fc1, const float* fc2, const float* w7,
                    const float* W, const float* bias, int N, int H, int Wd, int C, int Cr, float* z32, void* z16,
                    float* s, float* m, float* avg, float* max, int32_t* pstar, float* q, uint8_t* cstar,
                    void* workspace, void* stream);
/* Backward of the chain for dz = gz32 + gz16 (either may be NULL).  dx in x_dtype; d_fc1/d_fc2/d_w7/dW/db
 * are ACCUMULATED into (fp32, reference layouts); dz_out (nullable, fp32 NHWC) receives dz, which is also
 * the gradient of the residual input t. */
int sr_la_chain_bwd(const float* gz32, const void* gz16, const void* x, int x_dtype, const float* s, const float* m,
                    const float* avg, const float* max, const int32_t* pstar, const float* q, const uint8_t* cstar,
                    const float* fc1, const float* fc2, const float* w7, const float* W, int N, int H, int Wd, int C, int Cr,
                    void* dx, float* d_fc1, float* d_fc2, float* d_w7, float* dW, float* db, float* dz_out,
                    void* workspace, void* stream);

/* The same chain through ONE parameter block, with the fusions of the band path (csrc/la_band.cu; bf16, H*W <= 65535):
 *   pool_sum / pool_key [N][pool_rows][64] (nullable): per-(image, channel) partial sums and packed maxima of x that the
 *       PRODUCER of x emitted in its epilogue (sr_conv2d_fwd with desc.pool_*, or a previous chain's out_pool_*): the CLAM
 *       pooling then costs no pass over x.  key = (orderable bf16 bits << 16) | (0xFFFF - pixel index in the image).
 *   acc_in / acc_out (nullable): acc_out = acc_in + z — the dense-sampling sum `out_all += y` (model/sradsgan.py:459).
 *   out_pool_sum / out_pool_key [N][sr_la_chain_pool_rows(N,H,W)][64] (nullable): the partials of z16 for the chain that
 *       consumes it (the ResGroup tail reads the last RAB's output, :303-311).
 * Forward = one kernel (+ one pooling kernel when no partials are given); backward = three (no memset, no side stream).
 * Shapes outside the band path fall back to sr_la_chain_fwd / _bwd; acc_* / out_pool_* / gacc then fail with
 * SR_ERR_UNSUPPORTED (query sr_la_chain_band_path first).  tickets: N int32, ZERO before the first use (every launch
 * re-arms them), one stream at a time; NULL selects the tile kernels. */
typedef struct sr_la_chain_args {
    int32_t N, H, W, C, Cr, x_dtype;
    const void* x; const float* t;
    const float* fc1; const float* fc2; const float* w7; const float* Wm; const float* bias;
    const float* pool_sum; const uint32_t* pool_key; int32_t pool_rows;
    const float* acc_in; float* acc_out;
    float* z32; void* z16;
    float* s; float* m; float* avg; float* max; int32_t* pstar; float* q; uint8_t* cstar;
    float* out_pool_sum; uint32_t* out_pool_key;
    void* workspace;
} sr_la_chain_args;
typedef struct sr_la_chain_grad_args {
    int32_t N, H, W, C, Cr, x_dtype;
    const float* gz32; const void* gz16; const float* gacc;      /* dz = gz32 + gz16 + gacc (any may be NULL, not all) */
    const void* x; const float* s; const float* m; const float* avg; const float* max; const int32_t* pstar; const float* q;
    const uint8_t* cstar;
    const float* fc1; const float* fc2; const float* w7; const float* Wm;
    void* dx; float* d_fc1; float* d_fc2; float* d_w7; float* dW; float* db; float* dz_out;
    int32_t* tickets;
    void* workspace;
} sr_la_chain_grad_args;
int sr_la_chain_band_path(int N, int H, int W, int x_dtype);     /* 1 when the band kernels take this shape */
int sr_la_chain_pool_rows(int N, int H, int W);                   /* rows per image of out_pool_sum / out_pool_key */
int sr_la_chain_forward(const sr_la_chain_args* a, void* stream);
int sr_la_chain_backward(const sr_la_chain_grad_args* a, void* stream);

/* Backward of the fused conv epilogue: out = PixelUnshuffle_r( gy * act'(y) ), act' recovered from the
 * sign of the stored output y (LeakyReLU / ReLU; nn.LeakyReLU / nn.PixelShuffle of model/sradsgan.py:242,
 * :381-386,:428).  gy,y: (N, Ho*r, Wo*r, C/r^2) NHWC; out: (N, Ho, Wo, C) NHWC with channel c*r^2+sub. */
int sr_act_bwd(const void* gy, int gy_dtype, const void* y, int y_dtype, int act, float slope, int shuffle_r,
               int N, int Ho, int Wo, int C, void* out, int out_dtype, void* stream);

/* nn.MaxPool2d(2, 2) of torchvision's VGG19 features[4] / [9] inside FeatureExtractor (model/sradsgan.py:92-94), NHWC,
 * C % 8 == 0.  x: (N, H, W, C) -> y: (N, H/2, W/2, C) (floor).  The backward takes the saved INPUT x and routes dy to the
 * first maximum of each window (row-major order, as torch's max_pool2d_with_indices does); dx: (N, H, W, C), fully written. */
int sr_maxpool2x2_fwd(const void* x, int dtype, int N, int H, int W, int C, void* y, void* stream);
int sr_maxpool2x2_bwd(const void* dy, const void* x, int dtype, int N, int H, int W, int C, void* dx, void* stream);

/* Train-mode BatchNorm2d + LeakyReLU(slope) over x viewed as [rows = N*H*W][C] (C % 4 == 0), first-order only
 * (nn.BatchNorm2d + nn.LeakyReLU(0.2) of the discriminator blocks, model/sradsgan.py:476-479).
 * Normalises with the batch mean / biased variance; running_mean / running_var (nullable) are updated in place
 * with `momentum` and the unbiased variance, as torch does.  save: [4][C] fp32 (mean, rstd, scale, shift) for
 * the backward; workspace: >= 2*C floats. */
int sr_bn_act_fwd(const void* x, int dtype, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                  float momentum, float slope, float* running_mean, float* running_var, void* y, float* save,
                  void* workspace, void* stream);
/* dx, dgamma, dbeta (overwritten) from gy and the forward input x. */
int sr_bn_act_bwd(const void* gy, const void* x, int dtype, int64_t rows, int C, const float* save, float slope,
                  void* dx, float* dgamma, float* dbeta, void* stream);

/* Second-order piece of the WGAN-GP penalty (torch.autograd.grad(..., create_graph=True) through D followed by
 * .backward(), model/sradsgan.py:621,639,886): with dx = sr_bn_act_bwd(gy, x) and a cotangent u = dL/d(dx), writes
 * d_gy = dL/d(gy), d_x = dL/d(x) (both in `dtype`) and d_gamma = dL/d(gamma) (fp32, overwritten).  `save`, `dgamma`,
 * `dbeta` are the outputs of sr_bn_act_fwd / sr_bn_act_bwd for the same (gy, x).  workspace: >= 3*C floats. */
int sr_bn_act_bwd_bwd(const void* u, const void* gy, const void* x, int dtype, int64_t rows, int C, const float* save,
                      const float* dgamma, const float* dbeta, float slope, void* d_gy, void* d_x, float* d_gamma,
                      void* workspace, void* stream);

/* Position attention SGAM (model/sradsgan.py:153-176) without the N x N energy / softmax tensors (N = P tokens per image,
 * q, k: [N][P][8] in qk_dtype, v / o / dO: [N][P][64] bf16, statistics fp32 [N][P]).
 *   sr_sgam_stats : m = row max of q.k^T, linv = 1 / sum exp(q.k^T - m)                      (softmax(dim=-1), :169)
 *   sr_sgam_pv    : acc[r] = sum_c exp(a_r.b_c - row_m[r] - col_m[c]) row_s[r] col_s[c] vals[c]   (tcgen05)
 *                   forward  (a=q, b=k, row_m=m, row_s=linv, vals=v): o16 = acc, y32 = gamma*acc + resid32   (:170-172)
 *                   backward (a=k, b=q, col_m=m, col_s=linv, vals=dO): o16 = dV
 *   sr_sgam_ds    : out8[r] = sum_c P(r,c) (rowvals[r].colvals[c] - row_d[r] - col_d[c]) b_c       (tcgen05 + SIMT)
 *                   dQ: a=q, b=k, rowvals=dO, colvals=v, row_m=m, row_s=linv, row_d=D;  dK: a=k, b=q, rowvals=v, colvals=dO,
 *                   col_m=m, col_s=linv, col_d=D
 *   sr_sgam_bwd_prep : dO = gamma*dy (bf16), D[r] = sum_ch dO o, dgamma += sum dy o.
 * Statistic pointers may be NULL (0 for m / d, 1 for s). */
int sr_sgam_stats(const void* q, const void* k, int qk_dtype, int N, int P, float* m, float* linv, void* stream);
int sr_sgam_pv(const void* a, const void* b, int ab_dtype, const void* vals16, const float* row_m, const float* row_s,
               const float* col_m, const float* col_s, int N, int P, void* o16, float* y32, const float* resid32,
               const float* gamma, void* stream);
int sr_sgam_ds(const void* a, const void* b, int ab_dtype, const void* rowvals16, const void* colvals16, const float* row_m,
               const float* row_s, const float* row_d, const float* col_m, const float* col_s, const float* col_d, int N, int P,
               float* out8, void* stream);
int sr_sgam_bwd_prep(const float* dy, const void* o16, const float* gamma, int64_t rows, void* do16, float* d_out, float* dgamma,
                     void* stream);

/* out[c] = sum over rows of x[rows][C] (fp32 accumulate); sq (may be NULL) = sum of squares.
 * BatchNorm2d batch statistics (model/sradsgan.py:478) and bias gradients. */
int sr_colsum(const void* x, int dtype, int64_t rows, int C, float* sum, float* sq, int accumulate,
              void* stream);

/* Fused Adam (torch.optim.Adam semantics, model/sradsgan.py:724-725,858,887) over a flat fp32 buffer,
 * optionally followed by the WGAN weight clamp (model/sradsgan.py:891-892) when clamp_hi > clamp_lo.
 * The bias-correction step count is `step` (host, >= 1) or, when step_dev != NULL, the int32 read from
 * device memory at execution time (so a captured CUDA graph can be replayed step after step). */
int sr_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                 float lr, float beta1, float beta2, float eps, int step, const int32_t* step_dev,
                 float grad_scale, float clamp_lo, float clamp_hi, void* stream);

/* CGAM, the channel global attention of GAB_UP (model/sradsgan.py:178-213, light=False; C = 64), fp32 throughout:
 * y = gamma * (softmax_j(max_j E_ij - E_ij) X) + x with E = X X^T the 64 x 64 gram over the P pixels of each image.
 * x, y32, dy, dx: [N][P][64] fp32 NHWC; y16 (nullable): the same values in y16_dtype for the next convolutions;
 * A [N][64][64]: the attention, written by _fwd and read by _bwd; gamma: DEVICE fp32 scalar; dgamma[0] (+)= its gradient
 * (accumulate != 0 adds).  workspace >= sr_cgam_workspace_bytes(N, P).  Replaces torch.bmm x2 / max / softmax (+ autograd). */
size_t sr_cgam_workspace_bytes(int N, int P);
int sr_cgam_fwd(const float* x, const float* gamma, int N, int P, float* y32, void* y16, int y16_dtype, float* A, void* workspace,
                void* stream);
int sr_cgam_bwd(const float* dy, const float* x, const float* A, const float* gamma, int N, int P, float* dx, float* dgamma,
                int accumulate, void* workspace, void* stream);

/* One separable pass of PIL's 8-bit resampling (`Image.resize(size, Image.BICUBIC)` in the reference's datasets,
 * data/dataset.py:403-438; Pillow libImaging/Resample.c): out = clip8((2^21 + sum_k in[lo_o + k] * coeffs[o][k]) >> 22), int32,
 * bit-exact.  in [planes][H][W] uint8; axis 0: along W -> out [planes][H][out_size]; axis 1: along H -> out [planes][out_size][W];
 * bounds [out_size][2] = (first input index, tap count <= ksize), coeffs [out_size][ksize] with 22 fractional bits — both built by
 * the caller exactly like precompute_coeffs / normalize_coeffs_8bpc (sradsgan_b200/data.py: pil_coeffs). */
int sr_resample_u8(const uint8_t* in, int planes, int H, int W, uint8_t* out, int out_size, int axis, const int32_t* bounds,
                   const int32_t* coeffs, int ksize, void* stream);

/* ---- Discriminator attention (CBAM after block 6: model/base_networks.py:366-457, model/sradsgan.py:476-499), csrc/cbam.cu ----
 * A closed family of memory-bound primitives: the derivative of each is again a member, so the host wires them as autograd
 * nodes that are differentiable to any order (the critic is differentiated twice by WGAN-GP, model/sradsgan.py:611-639).
 * full = [N][P][C] NHWC activations in `dtype` (C a multiple of 64, 16-byte aligned); chan = [N][C] fp32; pix = [N][P] fp32;
 * every optional operand may be NULL.  Deterministic (fixed summation order), nothing synchronises.
 *
 * sr_cbam_ew:  y = x*s*m + s2*(g0/C + g1*[c == cidx_p]) + (a + b*[p == idx_c]) + acc
 *     x full, s chan (NULL = 1), m pix (NULL = 1): the gate application m * (s * x) (ChannelAttention :380-383, SpatialAttention :455-457)
 *     s2 chan, g0 / g1 pix, cidx pix (int32): adjoint of the channel pooling torch.mean / torch.max(dim=1) (:447-450)
 *     a / b chan, idx chan (int32): adjoint of AdaptiveAvgPool2d / AdaptiveMaxPool2d (:371-372); acc full (acc_dtype): added
 * sr_cbam_red_c:  out_c = scale * sum_p a*b*m + sum_p a*g1*[c == cidx_p]      (a full, b full | NULL, m pix | NULL)
 * sr_cbam_pool_hw: avg_c = mean_p x, mx_c = max_p x, idx_c = its first pixel
 * sr_cbam_red_p:  out_p = scale * sum_c a*b*s                                  (a full, b full | NULL, s chan | NULL)
 * sr_cbam_cpool:  q [N][2][P]: plane 0 = mean_c s_c*x_pc, plane 1 = max_c s_c*x_pc; cidx_p = its first channel
 * sr_cbam_gather_hw: out_c = x[idx_c][c];   sr_cbam_gather_c: out_p = x[p][cidx_p] * s[cidx_p] (s NULL = 1)
 * sr_small_gemm_nt: C[M][N] = sum_k A[m*lda_m + k*lda_k] * B[n*ldb_n + k*ldb_k] (fp32; the 256 <-> 16 shared MLP :374-378 and
 *     its gradients as strided views; replaces cuBLAS gemv/gemm launches on [16][256] operands). */
int sr_cbam_ew(const void* x, const float* s, const float* m, const float* s2, const float* g0, const float* g1, const int32_t* cidx,
               const float* a, const float* b, const int32_t* idx, const void* acc, int acc_dtype, void* y, int dtype, int N, int P, int C,
               void* stream);
int sr_cbam_red_c(const void* a, int a_dtype, const void* b, int b_dtype, const float* m, const float* g1, const int32_t* cidx, float scale,
                  int N, int P, int C, float* out, void* stream);
int sr_cbam_pool_hw(const void* x, int dtype, int N, int P, int C, float* avg, float* mx, int32_t* idx, void* stream);
int sr_cbam_red_p(const void* a, int a_dtype, const void* b, int b_dtype, const float* s, float scale, int N, int P, int C, float* out,
                  void* stream);
int sr_cbam_cpool(const void* x, int dtype, const float* s, int N, int P, int C, float* q, int32_t* cidx, void* stream);
int sr_cbam_gather_hw(const void* x, int dtype, const int32_t* idx, int N, int P, int C, float* out, void* stream);
int sr_cbam_gather_c(const void* x, int dtype, const float* s, const int32_t* cidx, int N, int P, int C, float* out, void* stream);
int sr_small_gemm_nt(const float* A, int64_t lda_m, int64_t lda_k, const float* B, int64_t ldb_n, int64_t ldb_k, int M, int N, int K, float* C,
                     void* stream);

/* ---- loss reductions and elementwise glue of one iteration (csrc/losses.cu) --------------------------------------------
 * Reductions are deterministic (per-block partials, the last block adds them in a fixed order) and need a caller-owned
 * workspace of sr_reduce_workspace_bytes() bytes that was ZERO when first used (every launch re-arms it) and is used by one
 * stream at a time.  `out` / `g` are DEVICE fp32 scalars; nothing synchronises.
 *
 * sr_diff_mean_fwd: out[0] = mean |a - b|^p, p = 1 (nn.L1Loss) or 2 (nn.MSELoss) — criterion_content of
 *   model/sradsgan.py:685-688 applied to (gen_hr, imgs_hr) :834 and to the VGG19 features :838.  a: n elements, NHWC;
 *   b: the same layout, or (b_nchw_C > 0) an NCHW tensor with b_nchw_C channels and b_HW pixels per plane (the HR batch as the
 *   host framework holds it).  sr_diff_mean_bwd: da[i] = g[0] * scale * d(mean|a-b|^p)/da[i]  (sign(0) = 0). */
size_t sr_reduce_workspace_bytes(void);
int sr_diff_mean_fwd(const void* a, int a_dtype, const void* b, int b_dtype, int64_t n, int p, int b_nchw_C, int64_t b_HW,
                     float* out, void* workspace, void* stream);
int sr_diff_mean_bwd(const void* a, int a_dtype, const void* b, int b_dtype, int64_t n, int p, int b_nchw_C, int64_t b_HW,
                     const float* g, float scale, void* da, int da_dtype, void* stream);
/* out[0] = scale * mean(x): GANLoss('wgan-gp') (:46-52; scale = -1 for real targets);  dx[i] = g[0] * scale / n. */
int sr_mean_fwd(const void* x, int dtype, int64_t n, float scale, float* out, void* workspace, void* stream);
int sr_mean_bwd(const float* g, float scale, int64_t n, void* dx, int dtype, void* stream);
/* WGAN-GP penalty (:623-637): grad [pixels][C] (NHWC view of the B x C x H x W input gradient, C <= 4), per-pixel norm over
 * the C colour channels (norm 0 = L2, 1 = L1, 2 = Linf), penalty 0 = 'LS' (n-1)^2 or 1 = 'hinge' relu(n-1), mean over pixels.
 * _bwd: dgrad = g[0] * scale * d(penalty)/d(grad) — the cotangent that enters the double backward through D (:639,:886). */
int sr_gp_penalty_fwd(const void* grad, int dtype, int64_t pixels, int C, int norm, int penalty, float* out, void* workspace,
                      void* stream);
int sr_gp_penalty_bwd(const void* grad, int dtype, int64_t pixels, int C, int norm, int penalty, const float* g, float scale,
                      void* dgrad, int out_dtype, void* stream);
/* out (NHWC) = alpha[n] * real + (1 - alpha[n]) * fake: the WGAN-GP interpolates (:611).  n elements, per_image elements per
 * sample; real is NHWC, or NCHW with real_nchw_C channels / HW pixels per plane when real_nchw_C > 0; fake is NHWC. */
int sr_lerp_nhwc(const void* real, int real_dtype, int real_nchw_C, const void* fake, int fake_dtype, const float* alpha, int64_t n,
                 int64_t per_image, int64_t HW, void* out, int out_dtype, void* stream);
/* x: NCHW fp32 [N][C][HW] (C <= 4, the batch as the data loader delivers it, :821-823) -> out: NHWC [N][HW][C] in out_dtype. */
int sr_nchw_to_nhwc(const float* x, int64_t N, int C, int64_t HW, void* out, int out_dtype, void* stream);
/* out = a + b elementwise (b may be NULL: a cast), independent dtypes — sums of gradient branches of different precision. */
int sr_add_cast(const void* a, int a_dtype, const void* b, int b_dtype, int64_t n, void* out, int out_dtype, void* stream);

#ifdef SR_WITH_PROBES
/* Diagnostics, compiled only with -DSR_WITH_PROBES (csrc/debug_probe.cu; not part of the product build): D[128][64] (fp32) = A_view * B^T on tcgen05, where A is a
 * [rows_a][64] bf16 matrix staged by TMA with the 128-byte swizzle and A_view row r is smem row
 * shift_rows + (r/8)*(sbo_bytes/128) + r%8 — probes which shared-memory operand descriptors the tensor core
 * accepts (shifted start address, non-1024 B group stride, descriptor base_offset field). */
int sr_debug_umma_shift(const void* a, int rows_a, const void* b, int shift_rows, int sbo_bytes, int base_offset,
                        float* out, void* stream);

/* Diagnostics: issue-rate of tcgen05.mma (M=128, N=n, K=16 bf16, operands in shared memory).  Each of `grid` CTAs
 * issues iters x num_acc x k_steps instructions, k_steps consecutive ones into the same of num_acc TMEM accumulators;
 * cycles[cta] = SM clock ticks from first issue to completion. */
int sr_debug_umma_rate(int n, int num_acc, int iters, int k_steps, int grid, int64_t* cycles, void* stream);

/* Diagnostics: role timeline of the halo convolution kernel.  While `buf` (device, [CTAs][128] int64, zeroed by the caller) is
 * set, every conv_halo launch stamps the SM clock at its milestones (slot map: csrc/conv_halo.cu); NULL switches it off.
 * dbg: experiment bits for the epilogue (1: no global stores, 2: no accumulator reads either; results are then garbage). */
int sr_debug_halo_trace(int64_t* buf, int dbg);

/* Diagnostics: writes a [rows][row_bytes] buffer once with the store pattern of a convolution epilogue (0: lane = row, 16 B per
 * lane and instruction; 1: lane quads share a row; 2: lane octets; 3: fully contiguous) — L2 write bandwidth vs LSU line rate. */
int sr_debug_store_pattern(void* out, int rows, int row_bytes, int pattern, int grid, void* stream);
#endif

#ifdef __cplusplus
}
#endif
#endif /* SRADSGAN_B200_H */
