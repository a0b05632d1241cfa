"""ctypes binding of libsradsgan_b200.so (C ABI declared in include/sradsgan_b200.h).

The library is the ONLY compute path of this package: there is no CPU or eager-PyTorch fallback.  If the
shared object is missing, or the device is not sm_100, every op raises RuntimeError.
"""
import ctypes
import os
from collections import namedtuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SR_LIB_PATH: a DIAGNOSTICS build of the same library (python __graft_entry__.py --probes writes build/probes/...); the product
# always loads the in-tree file
LIB_PATH = os.environ.get("SR_LIB_PATH") or os.path.join(_HERE, "libsradsgan_b200.so")

SR_F32, SR_BF16 = 0, 1
ACT_NONE, ACT_LRELU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05, IMPL_HALO = 0, 1, 2, 3

EXPORTS = [
    "sr_last_error", "sr_version", "sr_device_check", "sr_launch_count", "sr_conv_uses_tcgen05", "sr_pack_weights",
    "sr_conv2d_fwd", "sr_conv2d_dgrad", "sr_conv2d_dgrad_act", "sr_conv2d_wgrad", "sr_colsum", "sr_adam_step",
    "sr_la_chain_workspace_bytes", "sr_la_chain_fwd", "sr_la_chain_bwd", "sr_act_bwd", "sr_bn_act_fwd", "sr_bn_act_bwd", "sr_bn_act_bwd_bwd", "sr_set_option", "sr_set_aux_streams", "sr_conv2d_wgrad_workspace_bytes",
    "sr_sgam_stats", "sr_sgam_pv", "sr_sgam_ds", "sr_sgam_bwd_prep", "sr_pack_weights_batched", "sr_maxpool2x2_fwd", "sr_maxpool2x2_bwd",
    "sr_reduce_workspace_bytes", "sr_diff_mean_fwd", "sr_diff_mean_bwd", "sr_mean_fwd", "sr_mean_bwd", "sr_gp_penalty_fwd", "sr_gp_penalty_bwd",
    "sr_lerp_nhwc", "sr_nchw_to_nhwc", "sr_add_cast", "sr_cgam_workspace_bytes", "sr_cgam_fwd", "sr_cgam_bwd",
    "sr_la_chain_band_path", "sr_la_chain_pool_rows", "sr_la_chain_forward", "sr_la_chain_backward", "sr_conv_pool_rows",
    "sr_cbam_ew", "sr_cbam_red_c", "sr_cbam_pool_hw", "sr_cbam_red_p", "sr_cbam_cpool", "sr_cbam_gather_hw", "sr_cbam_gather_c", "sr_small_gemm_nt",
    "sr_resample_u8",
]


class ConvDesc(ctypes.Structure):
    """struct sr_conv_desc"""
    _fields_ = [(n, ctypes.c_int32) for n in
                ("N", "H", "W", "Cin", "Ho", "Wo", "Cout", "kh", "kw", "stride", "pad", "in_dtype", "out_dtype", "act")]
    _fields_ += [("slope", ctypes.c_float), ("shuffle_r", ctypes.c_int32), ("impl", ctypes.c_int32),
                 ("pool_sum", ctypes.c_void_p), ("pool_key", ctypes.c_void_p)]


class LaChainArgs(ctypes.Structure):
    """struct sr_la_chain_args"""
    _fields_ = [(n, ctypes.c_int32) for n in ("N", "H", "W", "C", "Cr", "x_dtype")] + \
               [(n, ctypes.c_void_p) for n in ("x", "t", "fc1", "fc2", "w7", "Wm", "bias", "pool_sum", "pool_key")] + \
               [("pool_rows", ctypes.c_int32)] + \
               [(n, ctypes.c_void_p) for n in ("acc_in", "acc_out", "z32", "z16", "s", "m", "avg", "max", "pstar", "q", "cstar",
                                               "out_pool_sum", "out_pool_key", "workspace")]


class LaChainGradArgs(ctypes.Structure):
    """struct sr_la_chain_grad_args"""
    _fields_ = [(n, ctypes.c_int32) for n in ("N", "H", "W", "C", "Cr", "x_dtype")] + \
               [(n, ctypes.c_void_p) for n in ("gz32", "gz16", "gacc", "x", "s", "m", "avg", "max", "pstar", "q", "cstar", "fc1", "fc2", "w7",
                                               "Wm", "dx", "d_fc1", "d_fc2", "d_w7", "dW", "db", "dz_out", "tickets", "workspace")]


ConvGeom = namedtuple("ConvGeom", "N H W Cin Ho Wo Cout kh kw stride pad")


def conv_geom(x_shape, w_shape, stride, pad):
    n, cin, h, w = x_shape
    cout, cin_w, kh, kw = w_shape
    if cin != cin_w:
        raise ValueError("conv: input has %d channels, weight expects %d" % (cin, cin_w))
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (w + 2 * pad - kw) // stride + 1
    return ConvGeom(n, h, w, cin, ho, wo, cout, kh, kw, stride, pad)


_lib = None
_AUX = None          # the side streams / events registered with sr_set_aux_streams (kept alive for the life of the process)


def load():
    """Loads the shared object (once). Raises RuntimeError with the build hint when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "sradsgan_b200: %s not found — build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no fallback compute path" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    lib.sr_last_error.restype = ctypes.c_char_p
    lib.sr_last_error.argtypes = []
    lib.sr_version.restype = i32
    lib.sr_device_check.restype = i32
    lib.sr_launch_count.restype = i64
    lib.sr_conv_pool_rows.argtypes = [ctypes.POINTER(ConvDesc)]
    lib.sr_conv_pool_rows.restype = i32
    lib.sr_conv_uses_tcgen05.argtypes = [ctypes.POINTER(ConvDesc), i32]
    lib.sr_conv_uses_tcgen05.restype = i32
    lib.sr_pack_weights.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.sr_pack_weights_batched.argtypes = [vp, i32, i32, i32, vp]
    lib.sr_maxpool2x2_fwd.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
    lib.sr_maxpool2x2_bwd.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp]
    lib.sr_maxpool2x2_fwd.restype = i32
    lib.sr_maxpool2x2_bwd.restype = i32
    lib.sr_conv2d_fwd.argtypes = [ctypes.POINTER(ConvDesc), vp, vp, vp, vp, vp, vp]
    lib.sr_conv2d_dgrad.argtypes = [ctypes.POINTER(ConvDesc), vp, vp, vp, vp]
    lib.sr_conv2d_dgrad_act.argtypes = [ctypes.POINTER(ConvDesc), vp, vp, vp, i32, f32, vp, vp]
    lib.sr_conv2d_dgrad_act.restype = i32
    lib.sr_conv2d_wgrad.argtypes = [ctypes.POINTER(ConvDesc), vp, vp, vp, vp, i32, vp, ctypes.c_uint64, vp]
    lib.sr_conv2d_wgrad_workspace_bytes.argtypes = [ctypes.POINTER(ConvDesc)]
    lib.sr_conv2d_wgrad_workspace_bytes.restype = ctypes.c_size_t
    lib.sr_set_option.argtypes = [ctypes.c_char_p, i32]
    lib.sr_set_option.restype = i32
    lib.sr_set_aux_streams.argtypes = [ctypes.POINTER(vp), vp, ctypes.POINTER(vp), i32]
    lib.sr_set_aux_streams.restype = i32
    lib.sr_colsum.argtypes = [vp, i32, i64, i32, vp, vp, i32, vp]
    lib.sr_adam_step.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, vp, f32, f32, f32, vp]
    lib.sr_la_chain_workspace_bytes.argtypes = [i32, i32, i32]
    lib.sr_la_chain_workspace_bytes.restype = ctypes.c_size_t
    lib.sr_la_chain_fwd.argtypes = [vp, i32] + [vp] * 6 + [i32] * 5 + [vp] * 11
    lib.sr_la_chain_bwd.argtypes = [vp, vp, vp, i32] + [vp] * 11 + [i32] * 5 + [vp] * 9
    lib.sr_act_bwd.argtypes = [vp, i32, vp, i32, i32, f32, i32, i32, i32, i32, i32, vp, i32, vp]
    lib.sr_bn_act_fwd.argtypes = [vp, i32, i64, i32, vp, vp, f32, f32, f32, vp, vp, vp, vp, vp, vp]
    lib.sr_bn_act_bwd.argtypes = [vp, vp, i32, i64, i32, vp, f32, vp, vp, vp, vp]
    lib.sr_bn_act_bwd_bwd.argtypes = [vp, vp, vp, i32, i64, i32, vp, vp, vp, f32, vp, vp, vp, vp, vp]
    lib.sr_bn_act_bwd_bwd.restype = i32
    lib.sr_resample_u8.argtypes = [vp, i32, i32, i32, vp, i32, i32, vp, vp, i32, vp]
    lib.sr_resample_u8.restype = i32
    lib.sr_cbam_ew.argtypes = [vp] * 11 + [i32, vp, i32, i32, i32, i32, vp]
    lib.sr_cbam_red_c.argtypes = [vp, i32, vp, i32, vp, vp, vp, f32, i32, i32, i32, vp, vp]
    lib.sr_cbam_pool_hw.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.sr_cbam_red_p.argtypes = [vp, i32, vp, i32, vp, f32, i32, i32, i32, vp, vp]
    lib.sr_cbam_cpool.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp, vp]
    lib.sr_cbam_gather_hw.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp]
    lib.sr_cbam_gather_c.argtypes = [vp, i32, vp, vp, i32, i32, i32, vp, vp]
    lib.sr_small_gemm_nt.argtypes = [vp, i64, i64, vp, i64, i64, i32, i32, i32, vp, vp]
    for f in (lib.sr_cbam_ew, lib.sr_cbam_red_c, lib.sr_cbam_pool_hw, lib.sr_cbam_red_p, lib.sr_cbam_cpool, lib.sr_cbam_gather_hw,
              lib.sr_cbam_gather_c, lib.sr_small_gemm_nt):
        f.restype = i32
    if hasattr(lib, "sr_debug_umma_shift"):             # diagnostics build only (-DSR_WITH_PROBES)
        lib.sr_debug_umma_shift.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp]
        lib.sr_debug_umma_shift.restype = i32
        lib.sr_debug_umma_rate.argtypes = [i32, i32, i32, i32, i32, vp, vp]
        lib.sr_debug_umma_rate.restype = i32
    if hasattr(lib, "sr_debug_store_pattern"):
        lib.sr_debug_store_pattern.argtypes = [vp, i32, i32, i32, i32, vp]
        lib.sr_debug_store_pattern.restype = i32
    if hasattr(lib, "sr_debug_halo_trace"):
        lib.sr_debug_halo_trace.argtypes = [vp, i32]
        lib.sr_debug_halo_trace.restype = i32
    lib.sr_sgam_stats.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    lib.sr_sgam_pv.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]
    lib.sr_sgam_ds.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp]
    lib.sr_sgam_bwd_prep.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp]
    for name in ("sr_sgam_stats", "sr_sgam_pv", "sr_sgam_ds", "sr_sgam_bwd_prep"):
        getattr(lib, name).restype = i32
    lib.sr_reduce_workspace_bytes.restype = ctypes.c_size_t
    lib.sr_reduce_workspace_bytes.argtypes = []
    lib.sr_diff_mean_fwd.argtypes = [vp, i32, vp, i32, i64, i32, i32, i64, vp, vp, vp]
    lib.sr_diff_mean_bwd.argtypes = [vp, i32, vp, i32, i64, i32, i32, i64, vp, f32, vp, i32, vp]
    lib.sr_mean_fwd.argtypes = [vp, i32, i64, f32, vp, vp, vp]
    lib.sr_mean_bwd.argtypes = [vp, f32, i64, vp, i32, vp]
    lib.sr_gp_penalty_fwd.argtypes = [vp, i32, i64, i32, i32, i32, vp, vp, vp]
    lib.sr_gp_penalty_bwd.argtypes = [vp, i32, i64, i32, i32, i32, vp, f32, vp, i32, vp]
    lib.sr_lerp_nhwc.argtypes = [vp, i32, i32, vp, i32, vp, i64, i64, i64, vp, i32, vp]
    lib.sr_nchw_to_nhwc.argtypes = [vp, i64, i32, i64, vp, i32, vp]
    lib.sr_add_cast.argtypes = [vp, i32, vp, i32, i64, vp, i32, vp]
    lib.sr_cgam_workspace_bytes.restype = ctypes.c_size_t
    lib.sr_cgam_workspace_bytes.argtypes = [i32, i32]
    lib.sr_cgam_fwd.argtypes = [vp, vp, i32, i32, vp, vp, i32, vp, vp, vp]
    lib.sr_cgam_bwd.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, i32, vp, vp]
    for name in ("sr_diff_mean_fwd", "sr_diff_mean_bwd", "sr_mean_fwd", "sr_mean_bwd", "sr_gp_penalty_fwd", "sr_gp_penalty_bwd", "sr_lerp_nhwc",
                 "sr_nchw_to_nhwc", "sr_add_cast", "sr_cgam_fwd", "sr_cgam_bwd"):
        getattr(lib, name).restype = i32
    lib.sr_la_chain_band_path.argtypes = [i32, i32, i32, i32]
    lib.sr_la_chain_band_path.restype = i32
    lib.sr_la_chain_pool_rows.argtypes = [i32, i32, i32]
    lib.sr_la_chain_pool_rows.restype = i32
    lib.sr_la_chain_forward.argtypes = [ctypes.POINTER(LaChainArgs), vp]
    lib.sr_la_chain_forward.restype = i32
    lib.sr_la_chain_backward.argtypes = [ctypes.POINTER(LaChainGradArgs), vp]
    lib.sr_la_chain_backward.restype = i32
    # experiment knobs: the library never reads the environment; SR_* variables of THIS process are forwarded explicitly
    for k, v in os.environ.items():
        if k.startswith(("SR_LA_", "SR_HALO_", "SR_WG_", "SR_S2_", "SR_PDL")):
            try:
                lib.sr_set_option(k.encode(), int(v))
            except ValueError:
                pass
    for name in ("sr_la_chain_fwd", "sr_la_chain_bwd", "sr_act_bwd", "sr_bn_act_fwd", "sr_bn_act_bwd"):
        getattr(lib, name).restype = i32
    for name in ("sr_pack_weights", "sr_pack_weights_batched", "sr_conv2d_fwd", "sr_conv2d_dgrad", "sr_conv2d_dgrad_act", "sr_conv2d_wgrad", "sr_colsum", "sr_adam_step"):
        getattr(lib, name).restype = i32
    _lib = lib
    return lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("sradsgan_b200.%s failed (%d): %s" % (what, rc, _lib.sr_last_error().decode()))


def _dt(t):
    if t.dtype == torch.float32:
        return SR_F32
    if t.dtype == torch.bfloat16:
        return SR_BF16
    raise TypeError("sradsgan_b200: unsupported dtype %s (float32 / bfloat16 only)" % t.dtype)


def _tdtype(code):
    return torch.float32 if code == SR_F32 else torch.bfloat16


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _nhwc(t):
    """(N,C,H,W)-shaped tensor with NHWC memory (torch channels_last)."""
    if t.dim() != 4:
        raise ValueError("expected a 4-d tensor")
    return t.contiguous(memory_format=torch.channels_last)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sradsgan_b200: tensors must live on a CUDA device (sm_100); there is no CPU path")


class CudaBackend:
    """Thin tensor-level wrapper over the C ABI. Every method launches on torch's current stream."""

    name = "cuda"

    def __init__(self):
        self.lib = load()
        self.prof = None      # bench.py: list of (kernel class, flops, bytes, start event, end event)
        self._workspaces = {}

    def wgrad_workspace(self, d, device):
        """split-K scratch of the weight-gradient kernels: one persistent buffer PER STREAM (so weight gradients running on
        different streams never share one), handed to the library with every call"""
        need = int(self.lib.sr_conv2d_wgrad_workspace_bytes(ctypes.byref(d)))
        if need == 0:
            return None, 0
        key = (device, torch.cuda.current_stream().cuda_stream)
        ws = self._workspaces.get(key)
        if ws is None or ws.numel() < need:
            ws = self._workspaces[key] = torch.empty(need, dtype=torch.uint8, device=device)
        return ws, ws.numel()

    def ensure_aux_streams(self, device):
        """three side streams + fork / join events owned here and handed to the library (sr_set_aux_streams): the parity
        classes of stride-2 input gradients run next to each other on them"""
        global _AUX                                     # process-wide like the library's registration (a backend object may die)
        if _AUX is not None and _AUX["device"] == device:
            return
        streams = [torch.cuda.Stream(device=device) for _ in range(3)]
        events = [torch.cuda.Event() for _ in range(4)]
        cur = torch.cuda.current_stream(device)
        for e in events:
            e.record(cur)                               # materialises the CUDA event handle
        vp = ctypes.c_void_p
        sarr = (vp * 3)(*[s.cuda_stream for s in streams])
        jarr = (vp * 3)(*[e.cuda_event for e in events[1:]])
        _check(self.lib.sr_set_aux_streams(sarr, vp(events[0].cuda_event), jarr, 3), "set_aux_streams")
        _AUX = {"device": device, "streams": streams, "events": events}

    def _timed(self, kind, d, dgrad, call):
        """optional per-launch CUDA-event timing on the launching stream (bench.py roofline attribution)"""
        if self.prof is None or torch.cuda.is_current_stream_capturing():
            return call()
        tc = self.lib.sr_conv_uses_tcgen05(ctypes.byref(d), {"fwd": 0, "dgrad": 1, "wgrad": 2}[kind])
        m = d.N * d.Ho * d.Wo
        flops = 2.0 * m * d.Cout * d.Cin * d.kh * d.kw
        return self._timed_rec("conv_%s_%s" % (kind, "tcgen05" if tc else "simt"), flops, 0.0, call)

    def _timed_rec(self, kind, flops, nbytes, call):
        """bench.py attribution record: (kernel family, algorithmic FLOPs, algorithmic bytes, start event, end event)"""
        if self.prof is None or torch.cuda.is_current_stream_capturing():
            return call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = call()
        e1.record()
        self.prof.append((kind, flops, nbytes, e0, e1))
        return out

    def launch_count(self):
        return int(self.lib.sr_launch_count())

    def device_check(self):
        _check(self.lib.sr_device_check(), "device_check")

    # -- weights -------------------------------------------------------------------------------
    def pack_weights(self, w, mode, dtype, shuffle_r=0):
        _require_cuda(w)
        w = w.detach()
        if w.dtype != torch.float32 or not w.is_contiguous():
            w = w.float().contiguous()
        cout, cin, kh, kw = w.shape
        out = torch.empty((kh * kw, cout, cin) if mode == 0 else (kh * kw, cin, cout), dtype=dtype, device=w.device)
        _check(self.lib.sr_pack_weights(_ptr(w), _ptr(out), cout, cin, kh, kw, mode, _dt(out), int(shuffle_r), _stream()),
               "pack_weights")
        return out

    def pack_weights_batched(self, table, n_entries, total_blocks, dtype):
        """table: device int64 [n_entries, 8] (see include/sradsgan_b200.h); one launch re-packs every entry"""
        _require_cuda(table)
        _check(self.lib.sr_pack_weights_batched(_ptr(table), int(n_entries), int(total_blocks), (SR_F32 if dtype == torch.float32 else SR_BF16), _stream()),
               "pack_weights_batched")

    # -- convolution ---------------------------------------------------------------------------
    def _desc(self, g, in_dtype, out_dtype, act=ACT_NONE, slope=0.0, shuffle_r=0, impl=IMPL_AUTO):
        return ConvDesc(g.N, g.H, g.W, g.Cin, g.Ho, g.Wo, g.Cout, g.kh, g.kw, g.stride, g.pad, in_dtype, out_dtype,
                        act, float(slope), int(shuffle_r), int(impl))

    def conv_fwd(self, x, w_packed, bias, residual, g, act=ACT_NONE, slope=0.0, shuffle_r=0, out_dtype=None, impl=IMPL_AUTO, want_pool=False):
        """want_pool: also return the CLAM pooling partials of y emitted by the epilogue -> (y, (sum, key, rows) | None)"""
        _require_cuda(x, w_packed)
        x = _nhwc(x)
        out_dtype = x.dtype if out_dtype is None else out_dtype
        r = shuffle_r if shuffle_r and shuffle_r > 1 else 1
        y = torch.empty((g.N, g.Cout // (r * r), g.Ho * r, g.Wo * r), dtype=out_dtype, device=x.device,
                        memory_format=torch.channels_last)
        if residual is not None:
            residual = _nhwc(residual)
            if residual.dtype != out_dtype or residual.shape != y.shape:
                raise ValueError("conv_fwd: residual must match the output shape/dtype")
        if bias is not None:
            bias = bias.detach().float().contiguous()
        if w_packed.dtype != x.dtype:
            raise TypeError("conv_fwd: packed weights must have the activation dtype")
        d = self._desc(g, _dt(x), _dt(y), act, slope, shuffle_r, impl)
        pool = None
        if want_pool:
            rows = int(self.lib.sr_conv_pool_rows(ctypes.byref(d))) if residual is None else 0
            if rows > 0:
                pool = (torch.empty((g.N, rows, g.Cout), dtype=torch.float32, device=x.device),
                        torch.empty((g.N, rows, g.Cout), dtype=torch.int32, device=x.device), rows)
                d.pool_sum, d.pool_key = pool[0].data_ptr(), pool[1].data_ptr()
        self._timed("fwd", d, False, lambda: _check(
            self.lib.sr_conv2d_fwd(ctypes.byref(d), _ptr(x), _ptr(w_packed), _ptr(bias), _ptr(residual), _ptr(y), _stream()),
            "conv2d_fwd"))
        return (y, pool) if want_pool else y

    def conv_dgrad(self, dy, w_packed_t, g, out_dtype=None, impl=IMPL_AUTO):
        _require_cuda(dy, w_packed_t)
        dy = _nhwc(dy)
        out_dtype = dy.dtype if out_dtype is None else out_dtype
        dx = torch.empty((g.N, g.Cin, g.H, g.W), dtype=out_dtype, device=dy.device, memory_format=torch.channels_last)
        if w_packed_t.dtype != dy.dtype:
            raise TypeError("conv_dgrad: packed weights must have the gradient dtype")
        d = self._desc(g, _dt(dy), _dt(dx), impl=impl)
        if g.stride == 2:
            self.ensure_aux_streams(dy.device)
        self._timed("dgrad", d, True, lambda: _check(
            self.lib.sr_conv2d_dgrad(ctypes.byref(d), _ptr(dy), _ptr(w_packed_t), _ptr(dx), _stream()), "conv2d_dgrad"))
        return dx

    def conv_dgrad_act(self, dy, w_packed_t, g, y_prev, act, slope, impl=IMPL_AUTO):
        """dx = conv_dgrad(dy) * act'(y_prev) (y_prev: the activated tensor that was this conv's input)"""
        _require_cuda(dy, w_packed_t, y_prev)
        dy = _nhwc(dy)
        y_prev = _nhwc(y_prev)
        if y_prev.dtype != dy.dtype or tuple(y_prev.shape) != (g.N, g.Cin, g.H, g.W):
            raise ValueError("conv_dgrad_act: y_prev must have the input's shape and the gradient's dtype")
        dx = torch.empty((g.N, g.Cin, g.H, g.W), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        d = self._desc(g, _dt(dy), _dt(dx), impl=impl)
        self._timed("dgrad", d, True, lambda: _check(
            self.lib.sr_conv2d_dgrad_act(ctypes.byref(d), _ptr(dy), _ptr(w_packed_t), _ptr(y_prev), int(act), float(slope), _ptr(dx),
                                         _stream()), "conv2d_dgrad_act"))
        return dx

    def _same_dtype(self, x, dy):
        """weight-gradient operands must share a dtype: the SMALLER tensor is cast (the RGB-side fp32 input of a thin layer,
        not the 64-channel gradient next to it)"""
        if x.dtype == dy.dtype:
            return x, dy
        if x.numel() <= dy.numel():
            return self.add_cast(x, None, dy.dtype), dy
        return x, self.add_cast(dy, None, x.dtype)

    def conv_wgrad(self, x, dy, g, want_bias=True, impl=IMPL_AUTO):
        _require_cuda(x, dy)
        x, dy = self._same_dtype(_nhwc(x), _nhwc(dy))
        dw = torch.empty((g.Cout, g.Cin, g.kh, g.kw), dtype=torch.float32, device=x.device)
        db = torch.empty((g.Cout,), dtype=torch.float32, device=x.device) if want_bias else None
        d = self._desc(g, _dt(x), SR_F32, impl=impl)
        ws, nws = self.wgrad_workspace(d, x.device)
        self._timed("wgrad", d, False, lambda: _check(
            self.lib.sr_conv2d_wgrad(ctypes.byref(d), _ptr(x), _ptr(dy), _ptr(dw), _ptr(db), 0, _ptr(ws), nws, _stream()), "conv2d_wgrad"))
        return dw, db

    def conv_wgrad_into(self, x, dy, g, dw, db, impl=IMPL_AUTO):
        """dw (+)= wgrad, db (+)= colsum(dy): accumulate into existing fp32 buffers (FlatAdam's flat gradient views)"""
        _require_cuda(x, dy, dw)
        x, dy = self._same_dtype(_nhwc(x), _nhwc(dy))
        if dw.dtype != torch.float32 or not dw.is_contiguous() or (db is not None and (db.dtype != torch.float32 or not db.is_contiguous())):
            raise ValueError("conv_wgrad_into: contiguous fp32 gradient buffers required")
        d = self._desc(g, _dt(x), SR_F32, impl=impl)
        ws, nws = self.wgrad_workspace(d, x.device)
        self._timed("wgrad", d, False, lambda: _check(
            self.lib.sr_conv2d_wgrad(ctypes.byref(d), _ptr(x), _ptr(dy), _ptr(dw), _ptr(db), 1, _ptr(ws), nws, _stream()), "conv2d_wgrad"))

    # -- fused local-attention chain ---------------------------------------------------------------
    def la_band_path(self, x):
        n, c, h, w = x.shape
        return x.is_cuda and bool(self.lib.sr_la_chain_band_path(n, h, w, _dt(x)))

    def _la_tickets(self, device, n):
        t = getattr(self, "_la_ticket_buf", None)
        if t is None or t.device != device or t.numel() < n:
            t = self._la_ticket_buf = torch.zeros(max(n, 256), dtype=torch.int32, device=device)
        return t

    def la_chain_forward(self, x, t, fc1, fc2, w7, W, b, want_lowp=True, pool=None, acc=None, want_pool=False):
        """z = Conv1x1(SLAM(CLAM(x))) + t  ->  (z32, z16 | None, saved, acc + z | None, pooling partials of z16 | None).
        pool = (sum, key, rows): the producer's per-(image, channel) partials of x; acc: the dense-sampling accumulator."""
        _require_cuda(x, t)
        x = _nhwc(x)
        t = _nhwc(t.float())
        n, c, h, w = x.shape
        cr = fc1.shape[0]
        dev = x.device
        band = self.la_band_path(x) and want_lowp
        if not band and (acc is not None or want_pool):
            raise RuntimeError("la_chain_forward: accumulator / pooling outputs need the band path (bf16, H*W <= 65535)")
        f32 = dict(dtype=torch.float32, device=dev)
        z32 = torch.empty((n, c, h, w), memory_format=torch.channels_last, **f32)
        z16 = torch.empty((n, c, h, w), dtype=x.dtype, device=dev, memory_format=torch.channels_last) if want_lowp else None
        sv = {"s": torch.empty((n, c), **f32), "m": torch.empty((n, h * w), **f32), "avg": torch.empty((n, c), **f32),
              "max": torch.empty((n, c), **f32), "pstar": torch.empty((n, c), dtype=torch.int32, device=dev),
              "q": torch.empty((n, h * w, 2), **f32), "cstar": torch.empty((n, h * w), dtype=torch.uint8, device=dev)}
        ws = torch.empty(self.lib.sr_la_chain_workspace_bytes(n, h, w), dtype=torch.uint8, device=dev)
        wts = [a.detach().float().contiguous() for a in (fc1, fc2, w7, W, b)]
        acc_out = out_pool = None
        if acc is not None:
            acc = _nhwc(acc.float())
            acc_out = torch.empty_like(z32)
        if want_pool:
            rows = int(self.lib.sr_la_chain_pool_rows(n, h, w))
            out_pool = (torch.empty((n, rows, c), **f32), torch.empty((n, rows, c), dtype=torch.int32, device=dev), rows)
        if not band:
            pool = None
        a = LaChainArgs(n, h, w, c, cr, _dt(x), x.data_ptr(), t.data_ptr(), *[v.data_ptr() for v in wts],
                        pool[0].data_ptr() if pool else None, pool[1].data_ptr() if pool else None, int(pool[2]) if pool else 0,
                        acc.data_ptr() if acc is not None else None, acc_out.data_ptr() if acc_out is not None else None,
                        z32.data_ptr(), z16.data_ptr() if z16 is not None else None, sv["s"].data_ptr(), sv["m"].data_ptr(),
                        sv["avg"].data_ptr(), sv["max"].data_ptr(), sv["pstar"].data_ptr(), sv["q"].data_ptr(), sv["cstar"].data_ptr(),
                        out_pool[0].data_ptr() if out_pool else None, out_pool[1].data_ptr() if out_pool else None, ws.data_ptr())
        nbytes = float(x.numel()) * (x.element_size() + 4 + 4 + (x.element_size() if want_lowp else 0) + (8 if acc is not None else 0))
        self._timed_rec("la_chain_fwd", 0.0, nbytes, lambda: _check(self.lib.sr_la_chain_forward(ctypes.byref(a), _stream()), "la_chain_forward"))
        return z32, z16, sv, acc_out, out_pool

    def la_chain_backward(self, gz32, gz16, gacc, x, sv, fc1, fc2, w7, W, want_dz=True, into=None):
        """-> (dx, d_fc1, d_fc2, d_w7, dW, db, dz) for dz = gz32 + gz16 + gacc (any may be None);
        `into` = [d_fc1, d_fc2, d_w7, dW, db] fp32 buffers to accumulate into"""
        x = _nhwc(x)
        n, c, h, w = x.shape
        cr = fc1.shape[0]
        dev = x.device
        band = self.la_band_path(x)
        if gz32 is not None:
            gz32 = _nhwc(gz32.float())
        if gz16 is not None:
            gz16 = _nhwc(gz16.to(x.dtype))
        if gacc is not None:
            gacc = _nhwc(gacc.float())
            if not band:                               # tile kernels: fold the accumulator gradient into the fp32 one first
                gz32 = gacc if gz32 is None else self.add_cast(gz32, gacc, torch.float32)
                gacc = None
        f32 = dict(dtype=torch.float32, device=dev)
        dx = torch.empty_like(x)
        if into is not None:
            d_fc1, d_fc2, d_w7, dW, db = into
        else:
            d_fc1 = torch.zeros(fc1.shape, **f32); d_fc2 = torch.zeros(fc2.shape, **f32); d_w7 = torch.zeros(w7.shape, **f32)
            dW = torch.zeros(W.shape, **f32); db = torch.zeros((c,), **f32)
        present = [g for g in (gz32, gz16, gacc) if g is not None]
        if want_dz and len(present) == 1 and present[0].dtype == torch.float32:
            dz, dz_buf = present[0], None            # the residual gradient is the incoming gradient itself
        elif want_dz:
            dz = dz_buf = torch.empty((n, c, h, w), memory_format=torch.channels_last, **f32)
        else:
            dz, dz_buf = None, None
        ws = torch.empty(self.lib.sr_la_chain_workspace_bytes(n, h, w), dtype=torch.uint8, device=dev)
        wts = [a.detach().float().contiguous() for a in (fc1, fc2, w7, W)]
        tickets = self._la_tickets(dev, n) if band else None
        a = LaChainGradArgs(n, h, w, c, cr, _dt(x), gz32.data_ptr() if gz32 is not None else None, gz16.data_ptr() if gz16 is not None else None,
                            gacc.data_ptr() if gacc is not None else None, x.data_ptr(), sv["s"].data_ptr(), sv["m"].data_ptr(),
                            sv["avg"].data_ptr(), sv["max"].data_ptr(), sv["pstar"].data_ptr(), sv["q"].data_ptr(), sv["cstar"].data_ptr(),
                            *[v.data_ptr() for v in wts], dx.data_ptr(), d_fc1.data_ptr(), d_fc2.data_ptr(), d_w7.data_ptr(), dW.data_ptr(),
                            db.data_ptr(), dz_buf.data_ptr() if dz_buf is not None else None,
                            tickets.data_ptr() if tickets is not None else None, ws.data_ptr())
        es = x.element_size()
        nbytes = float(x.numel()) * (sum(g.element_size() for g in present) + es + es + (4 if dz_buf is not None else 0))
        self._timed_rec("la_chain_bwd", 0.0, nbytes, lambda: _check(self.lib.sr_la_chain_backward(ctypes.byref(a), _stream()), "la_chain_backward"))
        return dx, d_fc1, d_fc2, d_w7, dW, db, dz

    def la_chain_fwd(self, x, t, fc1, fc2, w7, W, b, want_lowp=True):
        """z = Conv1x1(SLAM(CLAM(x))) + t  ->  (z32, z16 | None, saved)"""
        return self.la_chain_forward(x, t, fc1, fc2, w7, W, b, want_lowp)[:3]

    def la_chain_bwd(self, gz32, gz16, x, sv, fc1, fc2, w7, W, want_dz=True, into=None):
        return self.la_chain_backward(gz32, gz16, None, x, sv, fc1, fc2, w7, W, want_dz, into)

    def act_bwd(self, gy, y, act, slope, shuffle_r, g, out_dtype):
        """gpre = PixelUnshuffle_r(gy * act'(y)) in out_dtype, shape (N, Cout, Ho, Wo)"""
        _require_cuda(gy, y)
        gy = _nhwc(gy)
        y = _nhwc(y)
        out = torch.empty((g.N, g.Cout, g.Ho, g.Wo), dtype=out_dtype, device=gy.device, memory_format=torch.channels_last)
        _check(self.lib.sr_act_bwd(_ptr(gy), _dt(gy), _ptr(y), _dt(y), int(act), float(slope), int(shuffle_r or 0), g.N, g.Ho, g.Wo,
                                   g.Cout, _ptr(out), _dt(out), _stream()), "act_bwd")
        return out

    # -- MaxPool2d(2, 2) (VGG19) ---------------------------------------------------------------------
    def maxpool2x2_fwd(self, x):
        _require_cuda(x)
        x = _nhwc(x)
        n, c, h, w = x.shape
        y = torch.empty((n, c, h // 2, w // 2), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        _check(self.lib.sr_maxpool2x2_fwd(_ptr(x), _dt(x), n, h, w, c, _ptr(y), _stream()), "maxpool2x2_fwd")
        return y

    def maxpool2x2_bwd(self, dy, x):
        _require_cuda(dy, x)
        x = _nhwc(x)
        dy = _nhwc(dy.to(x.dtype))
        n, c, h, w = x.shape
        dx = torch.empty_like(x)
        _check(self.lib.sr_maxpool2x2_bwd(_ptr(dy), _ptr(x), _dt(x), n, h, w, c, _ptr(dx), _stream()), "maxpool2x2_bwd")
        return dx

    # -- train-mode BatchNorm + LeakyReLU (first-order) ---------------------------------------------
    def bn_act_fwd(self, x, gamma, beta, running_mean, running_var, eps, momentum, slope):
        _require_cuda(x)
        x = _nhwc(x)
        n, c, h, w = x.shape
        y = torch.empty_like(x)
        save = torch.empty((4, c), dtype=torch.float32, device=x.device)
        ws = torch.empty((2 * c,), dtype=torch.float32, device=x.device)
        self._timed_rec("bn_act_fwd", 0.0, 3.0 * x.numel() * x.element_size(), lambda: _check(
            self.lib.sr_bn_act_fwd(_ptr(x), _dt(x), n * h * w, c, _ptr(gamma.detach().float().contiguous()),
                                   _ptr(beta.detach().float().contiguous()), float(eps), float(momentum), float(slope),
                                   _ptr(running_mean), _ptr(running_var), _ptr(y), _ptr(save), _ptr(ws), _stream()), "bn_act_fwd"))
        return y, save

    def bn_act_bwd(self, gy, x, save, slope):
        x = _nhwc(x)
        gy = _nhwc(gy.to(x.dtype))
        n, c, h, w = x.shape
        dx = torch.empty_like(x)
        dgamma = torch.empty((c,), dtype=torch.float32, device=x.device)
        dbeta = torch.empty((c,), dtype=torch.float32, device=x.device)
        self._timed_rec("bn_act_bwd", 0.0, 5.0 * x.numel() * x.element_size(), lambda: _check(
            self.lib.sr_bn_act_bwd(_ptr(gy), _ptr(x), _dt(x), n * h * w, c, _ptr(save), float(slope), _ptr(dx), _ptr(dgamma),
                                   _ptr(dbeta), _stream()), "bn_act_bwd"))
        return dx, dgamma, dbeta

    def bn_act_bwd_bwd(self, u, gy, x, save, dgamma, dbeta, slope):
        """cotangent u of dx -> (d_gy, d_x, d_gamma)"""
        x = _nhwc(x)
        gy = _nhwc(gy.to(x.dtype))
        u = _nhwc(u.to(x.dtype))
        n, c, h, w = x.shape
        d_gy = torch.empty_like(x)
        d_x = torch.empty_like(x)
        d_gamma = torch.empty((c,), dtype=torch.float32, device=x.device)
        ws = torch.empty((3 * c,), dtype=torch.float32, device=x.device)
        _check(self.lib.sr_bn_act_bwd_bwd(_ptr(u), _ptr(gy), _ptr(x), _dt(x), n * h * w, c, _ptr(save), _ptr(dgamma), _ptr(dbeta),
                                          float(slope), _ptr(d_gy), _ptr(d_x), _ptr(d_gamma), _ptr(ws), _stream()), "bn_act_bwd_bwd")
        return d_gy, d_x, d_gamma

    # -- SGAM flash attention (tokens-major views of NHWC tensors) ---------------------------------------
    def sgam_stats(self, q, k):
        """q, k: (N, 8, H, W) NHWC -> (m, linv): fp32 [N, H*W]"""
        q, k = _nhwc(q), _nhwc(k)
        n, d, h, w = q.shape
        m = torch.empty((n, h * w), dtype=torch.float32, device=q.device)
        linv = torch.empty_like(m)
        _check(self.lib.sr_sgam_stats(_ptr(q), _ptr(k), _dt(q), n, h * w, _ptr(m), _ptr(linv), _stream()), "sgam_stats")
        return m, linv

    def sgam_pv(self, a, b, vals, row_m=None, row_s=None, col_m=None, col_s=None, resid=None, gamma=None, want_o16=True):
        """-> (o16 | None, y32 | None); a, b: (N,8,H,W); vals: (N,64,H,W) bf16; resid: (N,64,H,W) fp32"""
        a, b, vals = _nhwc(a), _nhwc(b), _nhwc(vals)
        n, c, h, w = vals.shape
        o16 = torch.empty_like(vals) if want_o16 else None
        y32 = None
        if resid is not None:
            resid = _nhwc(resid.float())
            y32 = torch.empty_like(resid)
        _check(self.lib.sr_sgam_pv(_ptr(a), _ptr(b), _dt(a), _ptr(vals), _ptr(row_m), _ptr(row_s), _ptr(col_m), _ptr(col_s), n, h * w,
                                   _ptr(o16), _ptr(y32), _ptr(resid), _ptr(gamma), _stream()), "sgam_pv")
        return o16, y32

    def sgam_ds(self, a, b, rowvals, colvals, row_m=None, row_s=None, row_d=None, col_m=None, col_s=None, col_d=None):
        """-> out8 fp32 shaped (N, 8, H, W) NHWC"""
        a, b, rowvals, colvals = _nhwc(a), _nhwc(b), _nhwc(rowvals), _nhwc(colvals)
        n, c, h, w = rowvals.shape
        out = torch.empty((n, 8, h, w), dtype=torch.float32, device=a.device, memory_format=torch.channels_last)
        _check(self.lib.sr_sgam_ds(_ptr(a), _ptr(b), _dt(a), _ptr(rowvals), _ptr(colvals), _ptr(row_m), _ptr(row_s), _ptr(row_d),
                                   _ptr(col_m), _ptr(col_s), _ptr(col_d), n, h * w, _ptr(out), _stream()), "sgam_ds")
        return out

    def sgam_bwd_prep(self, dy, o16, gamma):
        """-> (dO bf16, D fp32 [N, H*W], dgamma fp32 [1])"""
        dy = _nhwc(dy.float())
        o16 = _nhwc(o16)
        n, c, h, w = dy.shape
        do16 = torch.empty_like(o16)
        d = torch.empty((n, h * w), dtype=torch.float32, device=dy.device)
        dgamma = torch.zeros((1,), dtype=torch.float32, device=dy.device)
        _check(self.lib.sr_sgam_bwd_prep(_ptr(dy), _ptr(o16), _ptr(gamma), n * h * w, _ptr(do16), _ptr(d), _ptr(dgamma), _stream()),
               "sgam_bwd_prep")
        return do16, d, dgamma

    # -- loss reductions / elementwise glue (csrc/losses.cu) ---------------------------------------------
    def reduce_ws(self, device):
        """persistent, initially zero workspace of the deterministic reductions (every launch re-arms its ticket)"""
        ws = getattr(self, "_reduce_ws", None)
        if ws is None or ws.device != device:
            ws = self._reduce_ws = torch.zeros(int(self.lib.sr_reduce_workspace_bytes()), dtype=torch.uint8, device=device)
        return ws

    @staticmethod
    def _dense(t):
        """(tensor, nchw_C, HW): a dense tensor as the kernels index it — NHWC / flat (nchw_C = 0) or NCHW (nchw_C = C)"""
        if t.dim() == 4 and t.shape[1] > 1 and t.shape[2] * t.shape[3] > 1:
            if t.is_contiguous(memory_format=torch.channels_last):
                return t, 0, t.shape[2] * t.shape[3]
            if t.is_contiguous():
                return t, t.shape[1], t.shape[2] * t.shape[3]
            return t.contiguous(memory_format=torch.channels_last), 0, t.shape[2] * t.shape[3]
        return t.contiguous(), 0, 1

    def _diff_args(self, a, b):
        _require_cuda(a, b)
        if a.shape != b.shape:
            raise ValueError("diff_mean: shapes differ")
        a, ca, hw = self._dense(a)
        b, cb, _ = self._dense(b)
        if ca:                                  # `a` is the tensor the gradient is written for: kept NHWC / flat
            if cb:
                ca = cb = 0                     # both NCHW: the same flat order
            else:
                a, ca = a.contiguous(memory_format=torch.channels_last), 0
        return a, b, cb, hw

    def diff_mean(self, a, b, p):
        """mean |a - b|^p -> fp32 0-dim tensor"""
        a, b, cb, hw = self._diff_args(a, b)
        out = torch.empty((), dtype=torch.float32, device=a.device)
        nbytes = float(a.numel()) * (a.element_size() + b.element_size())
        self._timed_rec("loss_reduce", 0.0, nbytes, lambda: _check(
            self.lib.sr_diff_mean_fwd(_ptr(a), _dt(a), _ptr(b), _dt(b), a.numel(), int(p), int(cb), int(hw), _ptr(out),
                                      _ptr(self.reduce_ws(a.device)), _stream()), "diff_mean_fwd"))
        return out

    def diff_mean_bwd(self, a, b, p, g, scale=1.0):
        """g * scale * d(mean|a-b|^p)/da, shaped / laid out / typed like a"""
        a2, b2, cb, hw = self._diff_args(a, b)
        da = torch.empty_like(a2)
        g = g.detach().float().contiguous()
        nbytes = float(a2.numel()) * (2 * a2.element_size() + b2.element_size())
        self._timed_rec("loss_reduce", 0.0, nbytes, lambda: _check(
            self.lib.sr_diff_mean_bwd(_ptr(a2), _dt(a2), _ptr(b2), _dt(b2), a2.numel(), int(p), int(cb), int(hw), _ptr(g), float(scale),
                                      _ptr(da), _dt(da), _stream()), "diff_mean_bwd"))
        return da

    def mean(self, x, scale=1.0):
        _require_cuda(x)
        x, _, _ = self._dense(x)
        out = torch.empty((), dtype=torch.float32, device=x.device)
        _check(self.lib.sr_mean_fwd(_ptr(x), _dt(x), x.numel(), float(scale), _ptr(out), _ptr(self.reduce_ws(x.device)), _stream()), "mean_fwd")
        return out

    def mean_bwd(self, g, scale, like):
        dx = torch.empty_like(like)
        if not (dx.is_contiguous() or dx.is_contiguous(memory_format=torch.channels_last)):
            dx = torch.empty(like.shape, dtype=like.dtype, device=like.device)
        g = g.detach().float().contiguous()
        _check(self.lib.sr_mean_bwd(_ptr(g), float(scale), dx.numel(), _ptr(dx), _dt(dx), _stream()), "mean_bwd")
        return dx

    def gp_penalty(self, grad, norm, penalty):
        """grad: (N, C<=4, H, W) -> mean over pixels of the penalty of the per-pixel channel norm (fp32 0-dim)"""
        _require_cuda(grad)
        grad = _nhwc(grad)
        n, c, h, w = grad.shape
        out = torch.empty((), dtype=torch.float32, device=grad.device)
        self._timed_rec("loss_reduce", 0.0, float(grad.numel()) * grad.element_size(), lambda: _check(
            self.lib.sr_gp_penalty_fwd(_ptr(grad), _dt(grad), n * h * w, c, int(norm), int(penalty), _ptr(out),
                                       _ptr(self.reduce_ws(grad.device)), _stream()), "gp_penalty_fwd"))
        return out

    def gp_penalty_bwd(self, grad, norm, penalty, g, scale=1.0):
        grad = _nhwc(grad)
        n, c, h, w = grad.shape
        d = torch.empty_like(grad)
        g = g.detach().float().contiguous()
        self._timed_rec("loss_reduce", 0.0, 2.0 * grad.numel() * grad.element_size(), lambda: _check(
            self.lib.sr_gp_penalty_bwd(_ptr(grad), _dt(grad), n * h * w, c, int(norm), int(penalty), _ptr(g), float(scale), _ptr(d), _dt(d),
                                       _stream()), "gp_penalty_bwd"))
        return d

    def lerp(self, real, fake, alpha, out_dtype):
        """alpha*real + (1-alpha)*fake per sample -> (N, C, H, W) NHWC in out_dtype; real may be NCHW or NHWC, fake NHWC"""
        _require_cuda(real, fake, alpha)
        real, cr, hw = self._dense(real)
        fake = _nhwc(fake)
        n = real.shape[0]
        alpha = alpha.detach().reshape(-1).float().contiguous()
        if alpha.numel() != n or real.shape != fake.shape:
            raise ValueError("lerp: alpha must hold one value per sample; real and fake must have the same shape")
        out = torch.empty(fake.shape, dtype=out_dtype, device=fake.device, memory_format=torch.channels_last)
        _check(self.lib.sr_lerp_nhwc(_ptr(real), _dt(real), int(cr), _ptr(fake), _dt(fake), _ptr(alpha), real.numel(), real.numel() // n,
                                     int(hw), _ptr(out), _dt(out), _stream()), "lerp_nhwc")
        return out

    def nchw_to_nhwc(self, x, out_dtype):
        """contiguous NCHW fp32 (C <= 4) -> the same logical tensor with NHWC memory in out_dtype"""
        _require_cuda(x)
        n, c, h, w = x.shape
        out = torch.empty((n, c, h, w), dtype=out_dtype, device=x.device, memory_format=torch.channels_last)
        _check(self.lib.sr_nchw_to_nhwc(_ptr(x), n, c, h * w, _ptr(out), _dt(out), _stream()), "nchw_to_nhwc")
        return out

    def add_cast(self, a, b, out_dtype):
        """a + b (b may be None) in out_dtype; a and b dense with the same memory layout"""
        _require_cuda(a, b)
        if b is not None and (a.shape != b.shape or a.stride() != b.stride()):
            b = b.contiguous(memory_format=torch.channels_last) if a.is_contiguous(memory_format=torch.channels_last) and a.dim() == 4 else b.contiguous()
            if a.stride() != b.stride():
                a = a.contiguous()
                b = b.contiguous()
        out = torch.empty_like(a, dtype=out_dtype)
        _check(self.lib.sr_add_cast(_ptr(a), _dt(a), _ptr(b), _dt(b) if b is not None else _dt(a), a.numel(), _ptr(out), _dt(out), _stream()),
               "add_cast")
        return out

    # -- CGAM (csrc/cgam.cu) ---------------------------------------------------------------------------
    def cgam_fwd(self, x, gamma, lowp_dtype=None):
        """x: (N, 64, H, W) fp32 NHWC -> (y32, y16 | None, A [N,64,64])"""
        _require_cuda(x, gamma)
        x = _nhwc(x.float())
        n, c, h, w = x.shape
        if c != 64:
            raise ValueError("cgam: 64 channels required")
        y32 = torch.empty_like(x)
        y16 = torch.empty_like(x, dtype=lowp_dtype) if lowp_dtype is not None and lowp_dtype != torch.float32 else None
        A = torch.empty((n, 64, 64), dtype=torch.float32, device=x.device)
        ws = torch.empty(int(self.lib.sr_cgam_workspace_bytes(n, h * w)), dtype=torch.uint8, device=x.device)
        g = gamma.detach().float().contiguous()
        self._timed_rec("cgam", 0.0, float(x.numel()) * (4 + 4 + 4 + (2 if y16 is not None else 0)), lambda: _check(
            self.lib.sr_cgam_fwd(_ptr(x), _ptr(g), n, h * w, _ptr(y32), _ptr(y16), _dt(y16) if y16 is not None else SR_F32, _ptr(A),
                                 _ptr(ws), _stream()), "cgam_fwd"))
        return y32, y16, A

    def cgam_bwd(self, dy, x, A, gamma, dgamma_into=None):
        """-> (dx fp32 NHWC, dgamma [1] fp32 or None when accumulated into dgamma_into)"""
        x = _nhwc(x.float())
        dy = _nhwc(dy.float())
        n, c, h, w = x.shape
        dx = torch.empty_like(x)
        dg = dgamma_into if dgamma_into is not None else torch.empty((1,), dtype=torch.float32, device=x.device)
        ws = torch.empty(int(self.lib.sr_cgam_workspace_bytes(n, h * w)), dtype=torch.uint8, device=x.device)
        g = gamma.detach().float().contiguous()
        self._timed_rec("cgam", 0.0, float(x.numel()) * 4 * 5, lambda: _check(
            self.lib.sr_cgam_bwd(_ptr(dy), _ptr(x), _ptr(A), _ptr(g), n, h * w, _ptr(dx), _ptr(dg), 1 if dgamma_into is not None else 0,
                                 _ptr(ws), _stream()), "cgam_bwd"))
        return dx, (None if dgamma_into is not None else dg)

    # -- reductions / optimiser ----------------------------------------------------------------
    # -- discriminator attention primitives (csrc/cbam.cu) ------------------------------------------------
    # full: (N, C, H, W) NHWC in the compute dtype; chan: [N, C] fp32; pix: [N, H*W] fp32 (any shape with N*H*W elements)
    @staticmethod
    def _f32c(t):
        return None if t is None else t.detach().float().contiguous()

    def cbam_ew(self, like, x=None, s=None, m=None, s2=None, g0=None, g1=None, cidx=None, a=None, b=None, idx=None, acc=None):
        n, c, h, w = like.shape
        x = _nhwc(x.detach()) if x is not None else None
        acc = _nhwc(acc.detach()) if acc is not None else None
        s, m, s2, g0, g1, a, b = (self._f32c(t) for t in (s, m, s2, g0, g1, a, b))
        y = torch.empty((n, c, h, w), dtype=like.dtype, device=like.device, memory_format=torch.channels_last)
        if x is not None and x.dtype != y.dtype:
            x = x.to(y.dtype)
        nbytes = float(y.numel() * y.element_size()) * (1 + (x is not None) + (acc is not None))
        self._timed_rec("d_attention", 0.0, nbytes, lambda: _check(self.lib.sr_cbam_ew(
            _ptr(x), _ptr(s), _ptr(m), _ptr(s2), _ptr(g0), _ptr(g1), _ptr(cidx), _ptr(a), _ptr(b), _ptr(idx), _ptr(acc),
            _dt(acc) if acc is not None else _dt(y), _ptr(y), _dt(y), n, h * w, c, _stream()), "cbam_ew"))
        return y

    def cbam_red_c(self, a, b=None, m=None, scale=1.0, g1=None, cidx=None):
        a = _nhwc(a.detach())
        b = _nhwc(b.detach()) if b is not None else None
        n, c, h, w = a.shape
        m, g1 = self._f32c(m), self._f32c(g1)
        out = torch.empty((n, c), dtype=torch.float32, device=a.device)
        self._timed_rec("d_attention", 0.0, float(a.numel() * a.element_size()) * (1 + (b is not None)), lambda: _check(self.lib.sr_cbam_red_c(
            _ptr(a), _dt(a), _ptr(b), _dt(b) if b is not None else _dt(a), _ptr(m), _ptr(g1), _ptr(cidx), float(scale), n, h * w, c, _ptr(out),
            _stream()), "cbam_red_c"))
        return out

    def cbam_pool_hw(self, x):
        """-> (pooled [2, N, C] fp32: avg then max, idx [N, C] int32: first pixel of the maximum)"""
        x = _nhwc(x.detach())
        n, c, h, w = x.shape
        pooled = torch.empty((2, n, c), dtype=torch.float32, device=x.device)
        idx = torch.empty((n, c), dtype=torch.int32, device=x.device)
        self._timed_rec("d_attention", 0.0, float(x.numel() * x.element_size()), lambda: _check(self.lib.sr_cbam_pool_hw(
            _ptr(x), _dt(x), n, h * w, c, _ptr(pooled[0]), _ptr(pooled[1]), _ptr(idx), _stream()), "cbam_pool_hw"))
        return pooled, idx

    def cbam_red_p(self, a, b=None, s=None, scale=1.0):
        a = _nhwc(a.detach())
        b = _nhwc(b.detach()) if b is not None else None
        n, c, h, w = a.shape
        s = self._f32c(s)
        out = torch.empty((n, h * w), dtype=torch.float32, device=a.device)
        self._timed_rec("d_attention", 0.0, float(a.numel() * a.element_size()) * (1 + (b is not None)), lambda: _check(self.lib.sr_cbam_red_p(
            _ptr(a), _dt(a), _ptr(b), _dt(b) if b is not None else _dt(a), _ptr(s), float(scale), n, h * w, c, _ptr(out), _stream()), "cbam_red_p"))
        return out

    def cbam_cpool(self, x, s):
        """-> (q (N, 2, H, W) fp32 NCHW: mean / max over the channels of s*x, cidx [N, H*W] int32)"""
        x = _nhwc(x.detach())
        n, c, h, w = x.shape
        s = self._f32c(s)
        q = torch.empty((n, 2, h, w), dtype=torch.float32, device=x.device)
        cidx = torch.empty((n, h * w), dtype=torch.int32, device=x.device)
        self._timed_rec("d_attention", 0.0, float(x.numel() * x.element_size()), lambda: _check(self.lib.sr_cbam_cpool(
            _ptr(x), _dt(x), _ptr(s), n, h * w, c, _ptr(q), _ptr(cidx), _stream()), "cbam_cpool"))
        return q, cidx

    def cbam_gather_hw(self, x, idx):
        x = _nhwc(x.detach())
        n, c, h, w = x.shape
        out = torch.empty((n, c), dtype=torch.float32, device=x.device)
        _check(self.lib.sr_cbam_gather_hw(_ptr(x), _dt(x), _ptr(idx), n, h * w, c, _ptr(out), _stream()), "cbam_gather_hw")
        return out

    def cbam_gather_c(self, x, s, cidx):
        x = _nhwc(x.detach())
        n, c, h, w = x.shape
        s = self._f32c(s)
        out = torch.empty((n, h * w), dtype=torch.float32, device=x.device)
        _check(self.lib.sr_cbam_gather_c(_ptr(x), _dt(x), _ptr(s), _ptr(cidx), n, h * w, c, _ptr(out), _stream()), "cbam_gather_c")
        return out

    def small_gemm_nt(self, a, b):
        """a [M, K] @ b [N, K]^T -> [M, N] fp32; a, b: fp32 2-d tensors with ANY strides (views are not copied)"""
        a, b = a.detach(), b.detach()
        if a.dtype != torch.float32:
            a = a.float()
        if b.dtype != torch.float32:
            b = b.float()
        (m, k), (n, k2) = a.shape, b.shape
        if k != k2:
            raise ValueError("small_gemm_nt: inner dimensions differ")
        out = torch.empty((m, n), dtype=torch.float32, device=a.device)
        _check(self.lib.sr_small_gemm_nt(_ptr(a), a.stride(0), a.stride(1), _ptr(b), b.stride(0), b.stride(1), m, n, k, _ptr(out), _stream()),
               "small_gemm_nt")
        return out

    def resample_u8(self, img, out_size, axis, bounds, coeffs):
        """one pass of PIL's 8-bit resampling: img (..., H, W) uint8 contiguous -> resampled along W (axis 0) or H (axis 1);
        bounds [out, 2] / coeffs [out, ksize] int32 on the device (data.pil_coeffs)"""
        _require_cuda(img, bounds, coeffs)
        img = img.contiguous()
        h, w = img.shape[-2:]
        planes = img.numel() // (h * w)
        shape = tuple(img.shape[:-2]) + ((out_size, w) if axis else (h, out_size))
        out = torch.empty(shape, dtype=torch.uint8, device=img.device)
        _check(self.lib.sr_resample_u8(_ptr(img), planes, h, w, _ptr(out), int(out_size), int(axis), _ptr(bounds), _ptr(coeffs),
                                       int(coeffs.shape[1]), _stream()), "resample_u8")
        return out

    def colsum(self, x2d, want_sq=False):
        _require_cuda(x2d)
        x2d = x2d.contiguous()
        rows, c = x2d.shape
        s = torch.empty((c,), dtype=torch.float32, device=x2d.device)
        q = torch.empty((c,), dtype=torch.float32, device=x2d.device) if want_sq else None
        _check(self.lib.sr_colsum(_ptr(x2d), _dt(x2d), rows, c, _ptr(s), _ptr(q), 0, _stream()), "colsum")
        return s, q

    def adam_step(self, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale=1.0, clamp=None,
                  step_tensor=None):
        _require_cuda(param, grad, exp_avg, exp_avg_sq)
        for t in (param, grad, exp_avg, exp_avg_sq):
            if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != param.numel():
                raise ValueError("adam_step: flat contiguous fp32 buffers of equal length required")
        lo, hi = (clamp if clamp is not None else (0.0, 0.0))
        _check(self.lib.sr_adam_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), lr, beta1,
                                     beta2, eps, int(step), _ptr(step_tensor), float(grad_scale), float(lo), float(hi),
                                     _stream()), "adam_step")


_backend = None


def backend():
    """The process-wide compute backend. Tests may install an emulation with set_backend()."""
    global _backend
    if _backend is None:
        _backend = CudaBackend()
    return _backend


def set_backend(b):
    """TEST HOOK: replace the backend (tests/ install oracle.ops_emu.EmuBackend to check the autograd
    wiring on CPU). The product never calls this."""
    global _backend
    prev = _backend
    _backend = b
    return prev
