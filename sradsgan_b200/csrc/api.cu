// extern "C" surface of libsradsgan_b200.so (see include/sradsgan_b200.h).
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace sr {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }

struct Option { char name[32]; int value; };
static Option g_options[64];
static int g_num_options = 0;
int option(const char* name, int dflt) {
    for (int i = 0; i < g_num_options; ++i)
        if (strcmp(g_options[i].name, name) == 0) return g_options[i].value;
    return dflt;
}
static AuxStreams g_aux = {{nullptr, nullptr, nullptr}, nullptr, {nullptr, nullptr, nullptr}, 0};
const AuxStreams& aux_streams() { return g_aux; }
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return SR_ERR_CUDA;
    }
    return SR_OK;
}

// conv_simt.cu
int conv_fwd_simt(const sr_conv_desc*, const void*, const void*, const float*, const void*, void*, cudaStream_t);
int conv_dgrad_simt(const sr_conv_desc*, const void*, const void*, void*, cudaStream_t);
int conv_wgrad_simt(const sr_conv_desc*, const void*, const void*, float*, cudaStream_t);
// conv_tc.cu
bool conv_tc_supported(const sr_conv_desc*, bool dgrad);
int conv_tc_run(const sr_conv_desc*, bool dgrad, const void*, const void*, const float*, const void*, void*, cudaStream_t);
// conv_halo.cu
bool conv_halo_supported(const sr_conv_desc*, bool dgrad);
int conv_halo_pool_rows(const sr_conv_desc*);
int conv_halo_run(const sr_conv_desc*, bool dgrad, const void*, const void*, const float*, const void*, void*, cudaStream_t,
                  const void* mask = nullptr, float mask_slope = 0.f);
// conv_tc_wgrad.cu
bool conv_tc_wgrad_supported(const sr_conv_desc*);
int conv_tc_wgrad_run(const sr_conv_desc*, const void*, const void*, float*, float*, size_t, cudaStream_t);
size_t conv_tc_wgrad_workspace_bytes(const sr_conv_desc*);
bool conv_wgrad_halo_supported(const sr_conv_desc*);
int conv_wgrad_halo_run(const sr_conv_desc*, const void*, const void*, float*, float*, float*, size_t, cudaStream_t);
// conv_thin.cu
bool thin_fwd_supported(const sr_conv_desc*, bool dgrad);
bool thin_wgrad_supported(const sr_conv_desc*);
int thin_fwd_run(const sr_conv_desc*, bool dgrad, const void*, const void*, const float*, void*, cudaStream_t);
int thin_wgrad_run(const sr_conv_desc*, const void*, const void*, float*, cudaStream_t);
// la_chain.cu
size_t la_workspace_bytes(int, int, int);
int la_chain_fwd(const void*, int, const float*, const float*, const float*, const float*, const float*, const float*, int, int, int,
                 int, float*, void*, float*, float*, float*, float*, int*, float*, unsigned char*, float*, cudaStream_t);
int la_chain_bwd(const float*, const void*, const void*, int, const float*, const float*, const float*, const float*, const int*,
                 const float*, const unsigned char*, const float*, const float*, const float*, const float*, int, int, int, int,
                 void*, float*, float*, float*, float*, float*, float*, float*, cudaStream_t);
int act_bwd(const void*, int, const void*, int, int, float, int, int, int, int, int, void*, int, cudaStream_t);
bool la_chain_band_path(int, int, int, int);
int la_band_count(int, int, int);
int la_chain_forward(const sr_la_chain_args*, cudaStream_t);
int la_chain_backward(const sr_la_chain_grad_args*, cudaStream_t);
// bn.cu
int bn_act_fwd(const void*, int, long long, int, const float*, const float*, float, float, float, float*, float*, void*, float*, float*, cudaStream_t);
int bn_act_bwd(const void*, const void*, int, long long, int, const float*, float, void*, float*, float*, cudaStream_t);
int bn_act_bwd_bwd(const void*, const void*, const void*, int, long long, int, const float*, const float*, const float*, float, void*, void*,
                   float*, float*, cudaStream_t);
// elementwise.cu
int pack_weights(const float*, void*, int, int, int, int, int, int, int, cudaStream_t);
int pack_weights_batched(const long long*, int, int, int, cudaStream_t);
int maxpool2_fwd(const void*, int, int, int, int, int, void*, cudaStream_t);
int maxpool2_bwd(const void*, const void*, int, int, int, int, int, void*, cudaStream_t);
int colsum(const void*, int, long long, int, float*, float*, int, cudaStream_t);
int adam_step(float*, const float*, float*, float*, long long, float, float, float, float, int, const int*, float, float, float, cudaStream_t);

// cgam.cu
size_t cgam_workspace_bytes(int, int);
int cgam_fwd(const float*, const float*, int, int, float*, void*, int, float*, float*, cudaStream_t);
int cgam_bwd(const float*, const float*, const float*, const float*, int, int, float*, float*, int, float*, cudaStream_t);
// losses.cu
size_t reduce_workspace_bytes();
int diff_mean_fwd(const void*, int, const void*, int, long long, int, int, long long, float*, float*, cudaStream_t);
int diff_mean_bwd(const void*, int, const void*, int, long long, int, int, long long, const float*, float, void*, int, cudaStream_t);
int mean_fwd(const void*, int, long long, float, float*, float*, cudaStream_t);
int mean_bwd(const float*, float, long long, void*, int, cudaStream_t);
int gp_penalty_fwd(const void*, int, long long, int, int, int, float*, float*, cudaStream_t);
int gp_penalty_bwd(const void*, int, long long, int, int, int, const float*, float, void*, int, cudaStream_t);
int lerp_nhwc(const void*, int, int, const void*, int, const float*, long long, long long, long long, void*, int, cudaStream_t);
int nchw_to_nhwc(const float*, long long, int, long long, void*, int, cudaStream_t);
int add_cast(const void*, int, const void*, int, long long, void*, int, cudaStream_t);

int resample_u8(const unsigned char*, int, int, int, unsigned char*, int, int, const int*, const int*, int, cudaStream_t);

// cbam.cu
struct CbamEw {
    const void* x; const float* s; const float* m;
    const float* s2; const float* g0; const float* g1; const int* cidx;
    const float* a; const float* b; const int* idx;
    const void* acc;
    void* y;
    int N, P, C;
    float inv_c;
};
struct CbamRedC {
    const void* a; const void* b; const float* m; const float* g1; const int* cidx;
    float* out; float* out_max; int* out_idx;
    int N, P, C;
    float scale;
};
struct CbamRedP {
    const void* a; const void* b; const float* s;
    float* out; int* cidx;
    long long NP; int P, C;
    float scale;
};
int cbam_ew(const CbamEw&, int, int, cudaStream_t);
int cbam_red_c(const CbamRedC&, int, int, bool, cudaStream_t);
int cbam_red_p(const CbamRedP&, int, int, bool, cudaStream_t);
int cbam_gather_hw(const void*, int, const int*, int, int, int, float*, cudaStream_t);
int cbam_gather_c(const void*, int, const float*, const int*, int, int, int, float*, cudaStream_t);
int small_gemm_nt(const float*, long long, long long, const float*, long long, long long, int, int, int, float*, cudaStream_t);

// sgam.cu
struct SgCommon {
    const void* a; const void* b; int ab_f32;
    const float* row_m; const float* row_s; const float* row_d; const float* col_m; const float* col_s; const float* col_d;
    int N, P;
};
int sgam_stats(const void*, const void*, int, int, int, float*, float*, cudaStream_t);
int sgam_pv(const SgCommon&, const void*, __nv_bfloat16*, float*, const float*, const float*, cudaStream_t);
int sgam_ds(const SgCommon&, const void*, const void*, float*, cudaStream_t);
int sgam_bwd_prep(const float*, const void*, const float*, long long, void*, float*, float*, cudaStream_t);
#ifdef SR_WITH_PROBES
// debug_probe.cu (diagnostics: built only with -DSR_WITH_PROBES, see __graft_entry__.build(probes=True))
int debug_umma_shift(const void*, int, const void*, int, int, int, float*, cudaStream_t);
int debug_umma_rate(int, int, int, int, int, long long*, cudaStream_t);
int debug_store_pattern(void*, int, int, int, int, cudaStream_t);
extern int g_hl_dbg;
extern long long* g_hl_trace;          // conv_halo.cu: role timeline of the halo convolution ([grid][128] SM clock stamps)
#endif

static int g_arch_ok = -1;
static int arch_check() {
    if (g_arch_ok < 0) {
        int dev = 0, major = 0, minor = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device"); return SR_ERR_CUDA; }
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
        g_arch_ok = (major == 10) ? 1 : 0;
        if (!g_arch_ok) set_error("device is sm_%d%d; this library is built for sm_100a only", major, minor);
    }
    if (!g_arch_ok) {
        set_error("device is not sm_100 (B200); no fallback path exists");
        return SR_ERR_ARCH;
    }
    return SR_OK;
}

static int check_desc(const sr_conv_desc* d) {
    SR_REQUIRE(d != nullptr, "conv desc is NULL");
    SR_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "conv: non-positive dims");
    SR_REQUIRE(d->kh > 0 && d->kw > 0 && d->stride > 0 && d->pad >= 0, "conv: bad kernel/stride/pad");
    const int ho = (d->H + 2 * d->pad - d->kh) / d->stride + 1, wo = (d->W + 2 * d->pad - d->kw) / d->stride + 1;
    SR_REQUIRE(ho == d->Ho && wo == d->Wo, "conv: Ho/Wo (%d,%d) inconsistent with geometry (%d,%d)", d->Ho, d->Wo, ho, wo);
    SR_REQUIRE((d->in_dtype == SR_F32 || d->in_dtype == SR_BF16) && (d->out_dtype == SR_F32 || d->out_dtype == SR_BF16), "conv: bad dtype");
    if (d->shuffle_r > 1) SR_REQUIRE(d->Cout % (d->shuffle_r * d->shuffle_r) == 0, "conv: Cout %% r^2 != 0");
    return SR_OK;
}

}  // namespace sr

using namespace sr;

extern "C" {

const char* sr_last_error(void) { return g_err; }
int sr_version(void) { return 100; }
int sr_device_check(void) { return arch_check(); }
int64_t sr_launch_count(void) { return (int64_t)g_launches.load(); }

int sr_conv_pool_rows(const sr_conv_desc* d) {
    if (!d || d->impl == SR_IMPL_SIMT || d->impl == SR_IMPL_TCGEN05 || (d->Cout <= 4)) return 0;
    return conv_halo_pool_rows(d);
}

int sr_conv_uses_tcgen05(const sr_conv_desc* d, int kind) {
    if (!d || d->impl == SR_IMPL_SIMT) return 0;
    if (kind == 2) return conv_tc_wgrad_supported(d) ? 1 : 0;
    return (conv_tc_supported(d, kind == 1) || conv_halo_supported(d, kind == 1)) ? 1 : 0;
}

int sr_pack_weights(const float* w, void* packed, int Cout, int Cin, int kh, int kw, int mode, int dtype,
                    int shuffle_r, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(w && packed && Cout > 0 && Cin > 0 && kh > 0 && kw > 0, "pack_weights: bad arguments");
    SR_REQUIRE(mode == 0 || mode == 1, "pack_weights: mode must be 0 or 1");
    SR_REQUIRE(dtype == SR_F32 || dtype == SR_BF16, "pack_weights: bad dtype");
    SR_REQUIRE(shuffle_r <= 1 || (mode == 0 && Cout % (shuffle_r * shuffle_r) == 0), "pack_weights: bad shuffle_r");
    return pack_weights(w, packed, Cout, Cin, kh, kw, mode, dtype, shuffle_r, (cudaStream_t)stream);
}

int sr_pack_weights_batched(const void* table_dev, int n_entries, int total_blocks, int dtype, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(table_dev && n_entries > 0 && total_blocks > 0, "pack_weights_batched: bad arguments");
    SR_REQUIRE(dtype == SR_F32 || dtype == SR_BF16, "pack_weights_batched: bad dtype");
    return pack_weights_batched((const long long*)table_dev, n_entries, total_blocks, dtype, (cudaStream_t)stream);
}

int sr_conv2d_fwd(const sr_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual,
                  void* y, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    rc = check_desc(d);
    if (rc) return rc;
    SR_REQUIRE(x && w && y, "conv2d_fwd: NULL pointer");
    const bool tc_ok = conv_tc_supported(d, false);
    const bool halo_ok = conv_halo_supported(d, false);
    if (d->impl == SR_IMPL_HALO && !halo_ok) {
        set_error("conv2d_fwd: the halo-tile tcgen05 path needs 3x3 / stride 1 / pad 1, bf16, Cin %% 64 == 0 (Cin=%d Cout=%d k=%d s=%d)", d->Cin, d->Cout, d->kh, d->stride);
        return SR_ERR_UNSUPPORTED;
    }
    if (halo_ok && !(d->Cout <= 4 && residual) && (d->impl == SR_IMPL_AUTO || d->impl == SR_IMPL_HALO))
        return conv_halo_run(d, false, x, w, bias, residual, y, (cudaStream_t)stream);
    if (d->pool_sum || d->pool_key) { set_error("conv2d_fwd: pooling partials are emitted by the halo-tile kernel only (sr_conv_pool_rows() == 0 here)"); return SR_ERR_UNSUPPORTED; }
    if (d->impl == SR_IMPL_TCGEN05 && !tc_ok) {
        set_error("conv2d_fwd: tcgen05 path does not support this shape (Cin=%d Cout=%d k=%d s=%d dtype=%d)", d->Cin, d->Cout, d->kh, d->stride, d->in_dtype);
        return SR_ERR_UNSUPPORTED;
    }
    if (tc_ok && d->impl != SR_IMPL_SIMT) return conv_tc_run(d, false, x, w, bias, residual, y, (cudaStream_t)stream);
    if (d->impl == SR_IMPL_AUTO && !residual && thin_fwd_supported(d, false)) return thin_fwd_run(d, false, x, w, bias, y, (cudaStream_t)stream);
    return conv_fwd_simt(d, x, w, bias, residual, y, (cudaStream_t)stream);
}

int sr_conv2d_dgrad(const sr_conv_desc* d, const void* dy, const void* wt, void* dx, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    rc = check_desc(d);
    if (rc) return rc;
    SR_REQUIRE(dy && wt && dx, "conv2d_dgrad: NULL pointer");
    const bool tc_ok = conv_tc_supported(d, true);
    const bool halo_ok = conv_halo_supported(d, true);
    if (d->impl == SR_IMPL_HALO && !halo_ok) {
        set_error("conv2d_dgrad: the halo-tile tcgen05 path does not support this shape");
        return SR_ERR_UNSUPPORTED;
    }
    if (halo_ok && (d->impl == SR_IMPL_AUTO || d->impl == SR_IMPL_HALO)) return conv_halo_run(d, true, dy, wt, nullptr, nullptr, dx, (cudaStream_t)stream);
    if (d->impl == SR_IMPL_TCGEN05 && !tc_ok) {
        set_error("conv2d_dgrad: tcgen05 path does not support this shape");
        return SR_ERR_UNSUPPORTED;
    }
    if (tc_ok && d->impl != SR_IMPL_SIMT) return conv_tc_run(d, true, dy, wt, nullptr, nullptr, dx, (cudaStream_t)stream);
    if (d->impl == SR_IMPL_AUTO && thin_fwd_supported(d, true)) return thin_fwd_run(d, true, dy, wt, nullptr, dx, (cudaStream_t)stream);
    return conv_dgrad_simt(d, dy, wt, dx, (cudaStream_t)stream);
}

int sr_conv2d_dgrad_act(const sr_conv_desc* d, const void* dy, const void* wt, const void* y_prev, int act, float slope, void* dx,
                        void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    rc = check_desc(d);
    if (rc) return rc;
    SR_REQUIRE(dy && wt && dx && y_prev, "conv2d_dgrad_act: NULL pointer");
    SR_REQUIRE(act == SR_ACT_LRELU || act == SR_ACT_RELU, "conv2d_dgrad_act: LeakyReLU / ReLU only");
    const float ms = act == SR_ACT_RELU ? 0.f : slope;
    if (d->in_dtype == SR_BF16 && d->out_dtype == SR_BF16 && (d->impl == SR_IMPL_AUTO || d->impl == SR_IMPL_HALO) && conv_halo_supported(d, true))
        return conv_halo_run(d, true, dy, wt, nullptr, nullptr, dx, (cudaStream_t)stream, y_prev, ms);
    // unfused: input gradient, then the activation mask in place
    rc = sr_conv2d_dgrad(d, dy, wt, dx, stream);
    if (rc) return rc;
    return act_bwd(dx, d->out_dtype, y_prev, d->out_dtype, act, slope, 0, d->N, d->H, d->W, d->Cin, dx, d->out_dtype, (cudaStream_t)stream);
}

size_t sr_conv2d_wgrad_workspace_bytes(const sr_conv_desc* d) {
    if (!d || d->impl == SR_IMPL_SIMT || !conv_tc_wgrad_supported(d)) return 0;
    return conv_tc_wgrad_workspace_bytes(d);
}

int sr_conv2d_wgrad(const sr_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, int accumulate,
                    void* workspace, uint64_t workspace_bytes, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    rc = check_desc(d);
    if (rc) return rc;
    SR_REQUIRE(x && dy && dw, "conv2d_wgrad: NULL pointer");
    SR_REQUIRE((workspace == nullptr) == (workspace_bytes == 0) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
               "conv2d_wgrad: workspace pointer (16-byte aligned) and size must both be given or both be zero");
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cout * d->Cin * d->kh * d->kw, st);
    const bool tc_ok = conv_tc_wgrad_supported(d);
    if (d->impl == SR_IMPL_TCGEN05 && !tc_ok) {
        set_error("conv2d_wgrad: tcgen05 path does not support this shape");
        return SR_ERR_UNSUPPORTED;
    }
    if (tc_ok && d->impl != SR_IMPL_SIMT && conv_wgrad_halo_supported(d)) {
        // 3x3 / stride 1: shifted-view kernel, bias gradient accumulated inside it
        if (dbias && !accumulate) cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)d->Cout, st);
        return conv_wgrad_halo_run(d, x, dy, dw, dbias, (float*)workspace, (size_t)workspace_bytes, st);
    }
    if (tc_ok && d->impl != SR_IMPL_SIMT) rc = conv_tc_wgrad_run(d, x, dy, dw, (float*)workspace, (size_t)workspace_bytes, st);
    else if (d->impl == SR_IMPL_AUTO && thin_wgrad_supported(d)) rc = thin_wgrad_run(d, x, dy, dw, st);
    else rc = conv_wgrad_simt(d, x, dy, dw, st);
    if (rc) return rc;
    if (dbias) rc = colsum(dy, d->in_dtype, (long long)d->N * d->Ho * d->Wo, d->Cout, dbias, nullptr, accumulate, st);
    return rc;
}

int sr_set_option(const char* name, int value) {
    SR_REQUIRE(name && strlen(name) > 0 && strlen(name) < 32, "set_option: name must have 1..31 characters");
    for (int i = 0; i < g_num_options; ++i)
        if (strcmp(g_options[i].name, name) == 0) { g_options[i].value = value; return SR_OK; }
    SR_REQUIRE(g_num_options < 64, "set_option: table full");
    strcpy(g_options[g_num_options].name, name);
    g_options[g_num_options++].value = value;
    return SR_OK;
}

int sr_set_aux_streams(void* const* streams, void* fork_event, void* const* join_events, int n) {
    SR_REQUIRE(n >= 0 && n <= 3 && (n == 0 || (streams && fork_event && join_events)), "set_aux_streams: 0..3 streams with their events");
    for (int i = 0; i < 3; ++i) { g_aux.stream[i] = nullptr; g_aux.join[i] = nullptr; }
    for (int i = 0; i < n; ++i) {
        SR_REQUIRE(streams[i] && join_events[i], "set_aux_streams: NULL stream / event");
        g_aux.stream[i] = (cudaStream_t)streams[i]; g_aux.join[i] = (cudaEvent_t)join_events[i];
    }
    g_aux.fork = n ? (cudaEvent_t)fork_event : nullptr;
    g_aux.n = n;
    return SR_OK;
}

size_t sr_la_chain_workspace_bytes(int N, int H, int W) { return la_workspace_bytes(N, H, W); }

int sr_la_chain_fwd(const void* x, int x_dtype, const float* t, const float* fc1, const float* fc2, const float* w7, const float* W,
                    const float* bias, int N, int H, int Wd, int C, int Cr, float* z32, void* z16, float* s, float* m, float* avg,
                    float* mx, int32_t* pstar, float* q, uint8_t* cstar, void* workspace, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(C == 64 && Cr >= 1 && Cr <= 16, "la_chain: C must be 64 and 1 <= Cr <= 16 (got %d, %d)", C, Cr);
    SR_REQUIRE(x && t && fc1 && fc2 && w7 && W && bias && z32 && s && m && avg && mx && pstar && q && cstar && workspace, "la_chain_fwd: NULL pointer");
    SR_REQUIRE(x_dtype == SR_F32 || x_dtype == SR_BF16, "la_chain_fwd: bad dtype");
    return la_chain_fwd(x, x_dtype, t, fc1, fc2, w7, W, bias, N, H, Wd, Cr, z32, z16, s, m, avg, mx, pstar, q, cstar,
                        (float*)workspace, (cudaStream_t)stream);
}

int sr_la_chain_bwd(const float* gz32, const void* gz16, const void* x, int x_dtype, const float* s, const float* m, const float* avg,
                    const float* mx, const int32_t* pstar, const float* q, const uint8_t* cstar, const float* fc1, const float* fc2,
                    const float* w7, const float* W, int N, int H, int Wd, int C, int Cr, void* dx, float* d_fc1, float* d_fc2,
                    float* d_w7, float* dW, float* db, float* dz_out, void* workspace, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(C == 64 && Cr >= 1 && Cr <= 16, "la_chain: C must be 64 and 1 <= Cr <= 16 (got %d, %d)", C, Cr);
    SR_REQUIRE((gz32 || gz16) && x && s && m && avg && mx && pstar && q && cstar && fc1 && fc2 && w7 && W && dx && d_fc1 && d_fc2 && d_w7 && dW && db && workspace,
               "la_chain_bwd: NULL pointer");
    return la_chain_bwd(gz32, gz16, x, x_dtype, s, m, avg, mx, pstar, q, cstar, fc1, fc2, w7, W, N, H, Wd, Cr, dx, d_fc1, d_fc2,
                        d_w7, dW, db, dz_out, (float*)workspace, (cudaStream_t)stream);
}

int sr_la_chain_band_path(int N, int H, int W, int x_dtype) { return (N > 0 && H > 0 && W > 0 && la_chain_band_path(N, H, W, x_dtype)) ? 1 : 0; }

int sr_la_chain_pool_rows(int N, int H, int W) { return (N > 0 && H > 0 && W > 0) ? la_band_count(N, H, W) : 0; }

int sr_la_chain_forward(const sr_la_chain_args* a, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(a != nullptr, "la_chain_forward: NULL argument block");
    SR_REQUIRE(a->C == 64 && a->Cr >= 1 && a->Cr <= 16, "la_chain: C must be 64 and 1 <= Cr <= 16 (got %d, %d)", a->C, a->Cr);
    SR_REQUIRE(a->N > 0 && a->H > 0 && a->W > 0 && (a->x_dtype == SR_F32 || a->x_dtype == SR_BF16), "la_chain_forward: bad geometry / dtype");
    SR_REQUIRE(a->x && a->t && a->fc1 && a->fc2 && a->w7 && a->Wm && a->bias && a->z32 && a->s && a->m && a->avg && a->max && a->pstar && a->q && a->cstar && a->workspace,
               "la_chain_forward: NULL pointer");
    SR_REQUIRE((a->acc_in == nullptr) == (a->acc_out == nullptr), "la_chain_forward: acc_in and acc_out must both be given or both NULL");
    SR_REQUIRE((a->out_pool_sum == nullptr) == (a->out_pool_key == nullptr), "la_chain_forward: out_pool_sum / out_pool_key go together");
    return la_chain_forward(a, (cudaStream_t)stream);
}

int sr_la_chain_backward(const sr_la_chain_grad_args* a, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(a != nullptr, "la_chain_backward: NULL argument block");
    SR_REQUIRE(a->C == 64 && a->Cr >= 1 && a->Cr <= 16, "la_chain: C must be 64 and 1 <= Cr <= 16 (got %d, %d)", a->C, a->Cr);
    SR_REQUIRE((a->gz32 || a->gz16 || a->gacc) && a->x && a->s && a->m && a->avg && a->max && a->pstar && a->q && a->cstar && a->fc1 && a->fc2 && a->w7 && a->Wm &&
               a->dx && a->d_fc1 && a->d_fc2 && a->d_w7 && a->dW && a->db && a->workspace, "la_chain_backward: NULL pointer");
    return la_chain_backward(a, (cudaStream_t)stream);
}

int sr_act_bwd(const void* gy, int gy_dtype, const void* y, int y_dtype, int act, float slope, int shuffle_r, int N, int Ho, int Wo,
               int C, void* out, int out_dtype, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(gy && y && out && N > 0 && Ho > 0 && Wo > 0 && C > 0, "act_bwd: bad arguments");
    SR_REQUIRE(shuffle_r <= 1 || C % (shuffle_r * shuffle_r) == 0, "act_bwd: C %% r^2 != 0");
    return act_bwd(gy, gy_dtype, y, y_dtype, act, slope, shuffle_r, N, Ho, Wo, C, out, out_dtype, (cudaStream_t)stream);
}

int sr_maxpool2x2_fwd(const void* x, int dtype, int N, int H, int W, int C, void* y, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(x && y && N > 0 && H > 1 && W > 1 && C > 0 && C % 8 == 0, "maxpool2x2_fwd: bad arguments (C must be a multiple of 8)");
    SR_REQUIRE(dtype == SR_F32 || dtype == SR_BF16, "maxpool2x2_fwd: bad dtype");
    return maxpool2_fwd(x, dtype, N, H, W, C, y, (cudaStream_t)stream);
}

int sr_maxpool2x2_bwd(const void* dy, const void* x, int dtype, int N, int H, int W, int C, void* dx, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(dy && x && dx && N > 0 && H > 1 && W > 1 && C > 0 && C % 8 == 0, "maxpool2x2_bwd: bad arguments (C must be a multiple of 8)");
    SR_REQUIRE(dtype == SR_F32 || dtype == SR_BF16, "maxpool2x2_bwd: bad dtype");
    return maxpool2_bwd(dy, x, dtype, N, H, W, C, dx, (cudaStream_t)stream);
}

int sr_bn_act_fwd(const void* x, int dtype, int64_t rows, int C, const float* gamma, const float* beta, float eps, float momentum,
                  float slope, float* running_mean, float* running_var, void* y, float* save, void* workspace, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(x && gamma && beta && y && save && workspace && rows > 0 && C > 0 && C % 4 == 0, "bn_act_fwd: bad arguments");
    SR_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_act_fwd: running stats must both be given or both NULL");
    return bn_act_fwd(x, dtype, rows, C, gamma, beta, eps, momentum, slope, running_mean, running_var, y, save, (float*)workspace,
                      (cudaStream_t)stream);
}

int sr_bn_act_bwd(const void* gy, const void* x, int dtype, int64_t rows, int C, const float* save, float slope, void* dx,
                  float* dgamma, float* dbeta, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(gy && x && save && dx && dgamma && dbeta && rows > 0 && C > 0 && C % 4 == 0, "bn_act_bwd: bad arguments");
    return bn_act_bwd(gy, x, dtype, rows, C, save, slope, dx, dgamma, dbeta, (cudaStream_t)stream);
}

int sr_bn_act_bwd_bwd(const void* u, const void* gy, const void* x, int dtype, int64_t rows, int C, const float* save,
                      const float* dgamma, const float* dbeta, float slope, void* d_gy, void* d_x, float* d_gamma, void* workspace,
                      void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(u && gy && x && save && dgamma && dbeta && d_gy && d_x && d_gamma && workspace && rows > 0 && C > 0 && C % 4 == 0,
               "bn_act_bwd_bwd: bad arguments");
    return bn_act_bwd_bwd(u, gy, x, dtype, rows, C, save, dgamma, dbeta, slope, d_gy, d_x, d_gamma, (float*)workspace, (cudaStream_t)stream);
}

int sr_sgam_stats(const void* q, const void* k, int qk_dtype, int N, int P, float* m, float* linv, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(q && k && m && linv && N > 0 && P > 0, "sgam_stats: bad arguments");
    SR_REQUIRE(qk_dtype == SR_F32 || qk_dtype == SR_BF16, "sgam_stats: bad dtype");
    return sgam_stats(q, k, qk_dtype, N, P, m, linv, (cudaStream_t)stream);
}

int sr_sgam_pv(const void* a, const void* b, int ab_dtype, const void* vals16, const float* row_m, const float* row_s, const float* col_m,
               const float* col_s, int N, int P, void* o16, float* y32, const float* resid32, const float* gamma, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(a && b && vals16 && N > 0 && P > 0 && (o16 || y32), "sgam_pv: bad arguments");
    SR_REQUIRE(!y32 || resid32, "sgam_pv: y32 needs resid32");
    SR_REQUIRE((long long)N * P < (1ll << 31), "sgam_pv: too many tokens");
    SgCommon c{a, b, ab_dtype == SR_F32 ? 1 : 0, row_m, row_s, nullptr, col_m, col_s, nullptr, N, P};
    return sgam_pv(c, vals16, (__nv_bfloat16*)o16, y32, resid32, gamma, (cudaStream_t)stream);
}

int sr_sgam_ds(const void* a, const void* b, int ab_dtype, const void* rowvals16, const void* colvals16, const float* row_m,
               const float* row_s, const float* row_d, const float* col_m, const float* col_s, const float* col_d, int N, int P,
               float* out8, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(a && b && rowvals16 && colvals16 && out8 && N > 0 && P > 0, "sgam_ds: bad arguments");
    SR_REQUIRE((long long)N * P < (1ll << 31), "sgam_ds: too many tokens");
    SgCommon c{a, b, ab_dtype == SR_F32 ? 1 : 0, row_m, row_s, row_d, col_m, col_s, col_d, N, P};
    return sgam_ds(c, rowvals16, colvals16, out8, (cudaStream_t)stream);
}

int sr_sgam_bwd_prep(const float* dy, const void* o16, const float* gamma, int64_t rows, void* do16, float* d_out, float* dgamma,
                     void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(dy && o16 && gamma && do16 && d_out && dgamma && rows > 0, "sgam_bwd_prep: bad arguments");
    return sgam_bwd_prep(dy, o16, gamma, rows, do16, d_out, dgamma, (cudaStream_t)stream);
}

int sr_colsum(const void* x, int dtype, int64_t rows, int C, float* sum, float* sq, int accumulate, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(x && sum && C > 0 && rows >= 0, "colsum: bad arguments");
    return colsum(x, dtype, rows, C, sum, sq, accumulate, (cudaStream_t)stream);
}

int sr_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                 float beta2, float eps, int step, const int32_t* step_dev, float grad_scale, float clamp_lo, float clamp_hi,
                 void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && (step >= 1 || step_dev), "adam_step: bad arguments");
    return adam_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step < 1 ? 1 : step, step_dev, grad_scale,
                     clamp_lo, clamp_hi, (cudaStream_t)stream);
}

static bool dt_ok(int d) { return d == SR_F32 || d == SR_BF16; }

size_t sr_cgam_workspace_bytes(int N, int P) { return cgam_workspace_bytes(N, P); }

int sr_cgam_fwd(const float* x, const float* gamma, int N, int P, float* y32, void* y16, int y16_dtype, float* A, void* workspace,
                void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(x && gamma && y32 && A && workspace && N > 0 && P > 0 && (!y16 || dt_ok(y16_dtype)), "cgam_fwd: bad arguments");
    return cgam_fwd(x, gamma, N, P, y32, y16, y16_dtype, A, (float*)workspace, (cudaStream_t)stream);
}

int sr_cgam_bwd(const float* dy, const float* x, const float* A, const float* gamma, int N, int P, float* dx, float* dgamma, int accumulate,
                void* workspace, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(dy && x && A && gamma && dx && dgamma && workspace && N > 0 && P > 0, "cgam_bwd: bad arguments");
    return cgam_bwd(dy, x, A, gamma, N, P, dx, dgamma, accumulate, (float*)workspace, (cudaStream_t)stream);
}

size_t sr_reduce_workspace_bytes(void) { return reduce_workspace_bytes(); }

int sr_diff_mean_fwd(const void* a, int a_dtype, const void* b, int b_dtype, int64_t n, int p, int b_nchw_C, int64_t b_HW, float* out,
                     void* workspace, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(a && b && out && workspace && n > 0 && (p == 1 || p == 2) && dt_ok(a_dtype) && dt_ok(b_dtype), "diff_mean_fwd: bad arguments");
    SR_REQUIRE(b_nchw_C == 0 || (b_nchw_C > 0 && b_HW > 0 && n % ((int64_t)b_nchw_C * b_HW) == 0), "diff_mean_fwd: bad NCHW geometry");
    return diff_mean_fwd(a, a_dtype, b, b_dtype, n, p, b_nchw_C, b_HW, out, (float*)workspace, (cudaStream_t)stream);
}

int sr_diff_mean_bwd(const void* a, int a_dtype, const void* b, int b_dtype, int64_t n, int p, int b_nchw_C, int64_t b_HW, const float* g,
                     float scale, void* da, int da_dtype, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(a && b && g && da && n > 0 && (p == 1 || p == 2) && dt_ok(a_dtype) && dt_ok(b_dtype) && dt_ok(da_dtype), "diff_mean_bwd: bad arguments");
    SR_REQUIRE(b_nchw_C == 0 || (b_nchw_C > 0 && b_HW > 0 && n % ((int64_t)b_nchw_C * b_HW) == 0), "diff_mean_bwd: bad NCHW geometry");
    return diff_mean_bwd(a, a_dtype, b, b_dtype, n, p, b_nchw_C, b_HW, g, scale, da, da_dtype, (cudaStream_t)stream);
}

int sr_mean_fwd(const void* x, int dtype, int64_t n, float scale, float* out, void* workspace, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(x && out && workspace && n > 0 && dt_ok(dtype), "mean_fwd: bad arguments");
    return mean_fwd(x, dtype, n, scale, out, (float*)workspace, (cudaStream_t)stream);
}

int sr_mean_bwd(const float* g, float scale, int64_t n, void* dx, int dtype, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(g && dx && n > 0 && dt_ok(dtype), "mean_bwd: bad arguments");
    return mean_bwd(g, scale, n, dx, dtype, (cudaStream_t)stream);
}

int sr_gp_penalty_fwd(const void* grad, int dtype, int64_t pixels, int C, int norm, int penalty, float* out, void* workspace, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(grad && out && workspace && pixels > 0 && C >= 1 && C <= 4 && norm >= 0 && norm <= 2 && (penalty == 0 || penalty == 1) && dt_ok(dtype),
               "gp_penalty_fwd: bad arguments (C must be 1..4)");
    return gp_penalty_fwd(grad, dtype, pixels, C, norm, penalty, out, (float*)workspace, (cudaStream_t)stream);
}

int sr_gp_penalty_bwd(const void* grad, int dtype, int64_t pixels, int C, int norm, int penalty, const float* g, float scale, void* dgrad,
                      int out_dtype, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(grad && g && dgrad && pixels > 0 && C >= 1 && C <= 4 && norm >= 0 && norm <= 2 && (penalty == 0 || penalty == 1) && dt_ok(dtype) && dt_ok(out_dtype),
               "gp_penalty_bwd: bad arguments (C must be 1..4)");
    return gp_penalty_bwd(grad, dtype, pixels, C, norm, penalty, g, scale, dgrad, out_dtype, (cudaStream_t)stream);
}

int sr_lerp_nhwc(const void* real, int real_dtype, int real_nchw_C, const void* fake, int fake_dtype, const float* alpha, int64_t n,
                 int64_t per_image, int64_t HW, void* out, int out_dtype, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(real && fake && alpha && out && n > 0 && per_image > 0 && n % per_image == 0 && dt_ok(real_dtype) && dt_ok(fake_dtype) && dt_ok(out_dtype),
               "lerp_nhwc: bad arguments");
    SR_REQUIRE(real_nchw_C == 0 || (HW > 0 && per_image == (int64_t)real_nchw_C * HW), "lerp_nhwc: bad NCHW geometry");
    return lerp_nhwc(real, real_dtype, real_nchw_C, fake, fake_dtype, alpha, n, per_image, HW, out, out_dtype, (cudaStream_t)stream);
}

int sr_nchw_to_nhwc(const float* x, int64_t N, int C, int64_t HW, void* out, int out_dtype, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(x && out && N > 0 && C >= 1 && C <= 4 && HW > 0 && dt_ok(out_dtype), "nchw_to_nhwc: bad arguments (C must be 1..4)");
    return nchw_to_nhwc(x, N, C, HW, out, out_dtype, (cudaStream_t)stream);
}

int sr_add_cast(const void* a, int a_dtype, const void* b, int b_dtype, int64_t n, void* out, int out_dtype, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(a && out && n > 0 && dt_ok(a_dtype) && dt_ok(out_dtype) && (!b || dt_ok(b_dtype)), "add_cast: bad arguments");
    SR_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (!b || (reinterpret_cast<uintptr_t>(b) & 15) == 0),
               "add_cast: 16-byte aligned buffers required");
    return add_cast(a, a_dtype, b, b ? b_dtype : a_dtype, n, out, out_dtype, (cudaStream_t)stream);
}

int sr_resample_u8(const uint8_t* in, int planes, int H, int W, uint8_t* out, int out_size, int axis, const int32_t* bounds,
                   const int32_t* coeffs, int ksize, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(in && out && bounds && coeffs && planes > 0 && H > 0 && W > 0 && out_size > 0 && ksize > 0 && (axis == 0 || axis == 1),
               "resample_u8: bad arguments");
    return resample_u8(in, planes, H, W, out, out_size, axis, bounds, coeffs, ksize, (cudaStream_t)stream);
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
#define CBAM_DIMS_OK(N, P, C) ((N) > 0 && (P) > 0 && (C) >= 64 && (C) % 64 == 0 && (long long)(N) * (P) < (1ll << 31))

int sr_cbam_ew(const void* x, const float* s, const float* m, const float* s2, const float* g0, const float* g1, const int32_t* cidx,
               const float* a, const float* b, const int32_t* idx, const void* acc, int acc_dtype, void* y, int dtype, int N, int P, int C,
               void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(y && dt_ok(dtype) && CBAM_DIMS_OK(N, P, C), "cbam_ew: bad arguments (C must be a multiple of 64)");
    SR_REQUIRE(x || s2 || a || b || acc, "cbam_ew: no term given");
    SR_REQUIRE((!g1 || cidx) && (!b || idx) && (!acc || dt_ok(acc_dtype)) && ((!g0 && !g1) || s2), "cbam_ew: inconsistent optional operands");
    SR_REQUIRE(al16(x) && al16(s) && al16(s2) && al16(a) && al16(b) && al16(idx) && al16(acc) && al16(y), "cbam_ew: 16-byte aligned buffers required");
    CbamEw q{x, s, m, s2, g0, g1, cidx, a, b, idx, acc, y, N, P, C, 1.f / (float)C};
    return cbam_ew(q, dtype, acc ? acc_dtype : dtype, (cudaStream_t)stream);
}

int sr_cbam_red_c(const void* a, int a_dtype, const void* b, int b_dtype, const float* m, const float* g1, const int32_t* cidx, float scale,
                  int N, int P, int C, float* out, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(a && out && dt_ok(a_dtype) && (!b || dt_ok(b_dtype)) && (!g1 || cidx) && CBAM_DIMS_OK(N, P, C), "cbam_red_c: bad arguments");
    SR_REQUIRE(al16(a) && al16(b), "cbam_red_c: 16-byte aligned buffers required");
    CbamRedC q{a, b, m, g1, cidx, out, nullptr, nullptr, N, P, C, scale};
    return cbam_red_c(q, a_dtype, b ? b_dtype : a_dtype, false, (cudaStream_t)stream);
}

int sr_cbam_pool_hw(const void* x, int dtype, int N, int P, int C, float* avg, float* mx, int32_t* idx, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(x && avg && mx && idx && dt_ok(dtype) && CBAM_DIMS_OK(N, P, C) && al16(x), "cbam_pool_hw: bad arguments");
    CbamRedC q{x, nullptr, nullptr, nullptr, nullptr, avg, mx, idx, N, P, C, 1.f / (float)P};
    return cbam_red_c(q, dtype, dtype, true, (cudaStream_t)stream);
}

int sr_cbam_red_p(const void* a, int a_dtype, const void* b, int b_dtype, const float* s, float scale, int N, int P, int C, float* out,
                  void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(a && out && dt_ok(a_dtype) && (!b || dt_ok(b_dtype)) && CBAM_DIMS_OK(N, P, C), "cbam_red_p: bad arguments");
    SR_REQUIRE(al16(a) && al16(b) && al16(s), "cbam_red_p: 16-byte aligned buffers required");
    CbamRedP q{a, b, s, out, nullptr, (long long)N * P, P, C, scale};
    return cbam_red_p(q, a_dtype, b ? b_dtype : a_dtype, false, (cudaStream_t)stream);
}

int sr_cbam_cpool(const void* x, int dtype, const float* s, int N, int P, int C, float* q2, int32_t* cidx, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(x && q2 && cidx && dt_ok(dtype) && CBAM_DIMS_OK(N, P, C) && al16(x) && al16(s), "cbam_cpool: bad arguments");
    CbamRedP q{x, nullptr, s, q2, cidx, (long long)N * P, P, C, 1.f / (float)C};
    return cbam_red_p(q, dtype, dtype, true, (cudaStream_t)stream);
}

int sr_cbam_gather_hw(const void* x, int dtype, const int32_t* idx, int N, int P, int C, float* out, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(x && idx && out && dt_ok(dtype) && CBAM_DIMS_OK(N, P, C), "cbam_gather_hw: bad arguments");
    return cbam_gather_hw(x, dtype, idx, N, P, C, out, (cudaStream_t)stream);
}

int sr_cbam_gather_c(const void* x, int dtype, const float* s, const int32_t* cidx, int N, int P, int C, float* out, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(x && cidx && out && dt_ok(dtype) && CBAM_DIMS_OK(N, P, C), "cbam_gather_c: bad arguments");
    return cbam_gather_c(x, dtype, s, cidx, N, P, C, out, (cudaStream_t)stream);
}

int sr_small_gemm_nt(const float* A, int64_t lda_m, int64_t lda_k, const float* B, int64_t ldb_n, int64_t ldb_k, int M, int N, int K, float* C,
                     void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && (long long)M * N <= (1 << 22), "small_gemm_nt: bad arguments");
    return small_gemm_nt(A, lda_m, lda_k, B, ldb_n, ldb_k, M, N, K, C, (cudaStream_t)stream);
}

#ifdef SR_WITH_PROBES
int sr_debug_umma_shift(const void* a, int rows_a, const void* b, int shift_rows, int sbo_bytes, int base_offset, float* out,
                        void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(a && b && out && shift_rows >= 0 && sbo_bytes > 0, "debug_umma_shift: bad arguments");
    return debug_umma_shift(a, rows_a, b, shift_rows, sbo_bytes, base_offset, out, (cudaStream_t)stream);
}

int sr_debug_umma_rate(int n, int num_acc, int iters, int k_steps, int grid, int64_t* cycles, void* stream) {
    int rc = arch_check();
    if (rc) return rc;
    SR_REQUIRE(cycles != nullptr, "debug_umma_rate: NULL output");
    return debug_umma_rate(n, num_acc, iters, k_steps, grid, (long long*)cycles, (cudaStream_t)stream);
}

int sr_debug_store_pattern(void* out, int rows, int row_bytes, int pattern, int grid, void* stream) {
    return debug_store_pattern(out, rows, row_bytes, pattern, grid, (cudaStream_t)stream);
}

int sr_debug_halo_trace(int64_t* buf, int dbg) { g_hl_trace = (long long*)buf; g_hl_dbg = dbg; return SR_OK; }
#endif  // SR_WITH_PROBES

}  // extern "C"
