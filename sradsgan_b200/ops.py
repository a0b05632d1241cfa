"""torch.autograd.Function wrappers over the C-ABI kernels.

The convolution trio (fwd / dgrad / wgrad) is closed under differentiation — each Function's backward is
written with the other two — so `torch.autograd.grad(..., create_graph=True)` through the discriminator
(the WGAN-GP penalty, reference model/sradsgan.py:621,639) works unchanged at the call site.
"""
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, IMPL_AUTO, conv_geom


class Config:
    """Process-wide numeric configuration of the hot path."""
    compute_dtype = torch.bfloat16     # activations / MMA operands ("bf16 mode"); torch.float32 = "fp32 mode"
    conv_impl = IMPL_AUTO              # SR_IMPL_* forced for every conv (tests)
    double_backward = False            # force the any-order differentiable (unfused) discriminator path
    input_grad_only = False            # backward passes skip parameter gradients (see input_grad_only())
    direct_grad = True                 # first-order backward adds weight gradients straight into FlatAdam's flat buffer
    fused_d_attention = True           # discriminator CBAM through the csrc/cbam.cu primitives (False: module-by-module ATen path)


config = Config()


class input_grad_only:
    """Context manager for `torch.autograd.grad(outputs, inputs=<activations>)` calls: custom Functions cannot see
    which of their inputs the engine actually needs (ctx.needs_input_grad is fixed at forward time), so without
    this the WGAN-GP input-gradient pass (reference model/sradsgan.py:621) would also compute — and discard — every
    discriminator weight gradient."""

    def __enter__(self):
        self.prev = config.input_grad_only
        config.input_grad_only = True

    def __exit__(self, *a):
        config.input_grad_only = self.prev


def _grad_target(p):
    """flat-buffer gradient view of parameter p when a first-order backward may accumulate into it directly"""
    if not config.direct_grad or torch.is_grad_enabled() or p is None or getattr(p, "_sr_shared", False):
        return None                    # weight-tied parameters (used more than once per forward) go through autograd's sum
    return getattr(p, "_sr_flat_grad", None)


# ----------------------------------------------------------------------------------------------
# weight gradients off the critical path
# ----------------------------------------------------------------------------------------------
class _WgradStream:
    """First-order weight gradients that accumulate straight into FlatAdam's flat buffer are needed by nothing before the
    optimiser step, while the input-gradient chain next to them is a sequence of short latency-bound kernels.  With
    SR_WGRAD_ASYNC=1 (default) they are launched on ONE side stream (one at a time: they share the split-K workspace) that
    waits for the producing kernels; `wgrad_join()` — called at the end of each phase and by FlatAdam.step / the gradient
    all-reduce — makes the main stream wait for them.  Works the same under CUDA-graph capture (a forked branch of the graph).
    The operands are kept referenced until the join, so the caching allocator cannot hand their memory to main-stream work."""
    enabled = os.environ.get("SR_WGRAD_ASYNC", "1") == "1"
    two_streams = os.environ.get("SR_WGRAD_STREAMS", "2") == "2"
    stream = None
    stream2 = None
    keep = []
    pending = []


def side_stream(device):
    ws = _WgradStream
    if ws.stream is None or ws.stream.device != device:
        ws.stream = torch.cuda.Stream(device=device)
    return ws.stream


def wgrad_join():
    ws = _WgradStream
    if ws.keep:
        cur = torch.cuda.current_stream()
        for st in ws.pending:          # only streams that received work since the last join (a stream that is not part of
            cur.wait_stream(st)        # the running graph capture must not be waited on)
        ws.pending.clear()
        ws.keep.clear()


def _wgrad_into(x, gy, g, tw, tb):
    be = _lib.backend()
    if not (_WgradStream.enabled and x.is_cuda):
        be.conv_wgrad_into(x, gy, g, tw, tb, impl=config.conv_impl)
        return
    ws = _WgradStream
    side_stream(x.device)
    st = ws.stream
    if (g.Cin <= 4 or g.Cout <= 4) and ws.two_streams:
        # thin layers (RGB input, 1-channel critic map) run direct / SIMT kernels that do not touch the split-K workspace:
        # they get their own stream and overlap with the tensor-core weight gradients of the other layers
        if ws.stream2 is None or ws.stream2.device != x.device:
            ws.stream2 = torch.cuda.Stream(device=x.device)
        st = ws.stream2
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        be.conv_wgrad_into(x, gy, g, tw, tb, impl=config.conv_impl)
    ws.keep.append((x, gy))
    if st not in ws.pending:
        ws.pending.append(st)


def _wgrad(x, gy, g, w, b, has_bias, need_w, need_b):
    """weight / bias gradient of one convolution inside a backward pass -> (gw, gb) to RETURN to autograd.
    First-order passes over FlatAdam-owned parameters accumulate in place (no temporary, no memset, no add
    kernel) and return None."""
    if config.input_grad_only or not (need_w or (has_bias and need_b)):
        return None, None
    tw = _grad_target(w)
    tb = _grad_target(b) if has_bias else None
    if tw is not None and (not has_bias or tb is not None):
        _wgrad_into(x, gy, g, tw, tb)
        return None, None              # autograd still runs the parameter's post-accumulate hooks (dp.BucketReducer)
    wgrad_join()                       # the synchronous weight-gradient kernels below use the same split-K workspace
    if torch.is_grad_enabled():
        gw, gb = ConvWgrad.apply(x, gy, g)
    else:
        gw, gb = _lib.backend().conv_wgrad(x, gy, g, want_bias=has_bias, impl=config.conv_impl)
    return gw, (gb if has_bias else None)


def set_precision(mode):
    """'bf16' (tcgen05 path, <=1e-2 per-layer) or 'fp32' (SIMT path, <=1e-4 per-layer)."""
    config.compute_dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[mode]


# ----------------------------------------------------------------------------------------------
# packed-weight cache: OIHW fp32 master -> packed compute-dtype operand, rebuilt when the master changes
# ----------------------------------------------------------------------------------------------
_generation = 0


def bump_weight_generation():
    """Called by the fused optimiser (it updates parameters through raw pointers)."""
    global _generation
    _generation += 1


def next_generation():
    """fresh, never repeated stamp for a network's parameters (FlatAdam.step)"""
    global _generation_counter
    _generation_counter += 1
    return _generation_counter


_generation_counter = 0


def _capturing(t):
    return t.is_cuda and torch.cuda.is_current_stream_capturing()


def _ver(w):
    ref = getattr(w, "_sr_genref", None)        # one mutable stamp per optimiser (FlatAdam.gen), bumped in O(1)
    return (_generation, ref[0] if ref is not None else getattr(w, "_sr_gen", 0), w._version, w.data_ptr())


def packed(w, mode, dtype, shuffle_r=0):
    """Packed operand for weight `w`. Cached ON the nn.Parameter object (never for temporaries such as the
    double-backward cotangents, whose storage may be recycled), keyed by layout and validated by
    (global generation, owning optimiser's generation, tensor version, storage address).
    While a CUDA graph is being captured a cached operand is only trusted when something inside the SAME graph
    keeps it fresh on replay (a PackPlan entry, re-packed at the start of every step) or the weight is frozen
    (VGG19); otherwise the pack kernel is captured with the use."""
    if not isinstance(w, torch.nn.Parameter):
        return _lib.backend().pack_weights(w, mode, dtype, shuffle_r)
    cache = w.__dict__.setdefault("_sr_pack", {})
    key = (mode, dtype, shuffle_r)
    hit = cache.get(key)
    capturing = _capturing(w)
    if hit is not None and hit[0] == _ver(w) and (not capturing or hit[2] or getattr(w, "_sr_frozen", False)):
        return hit[1]
    p = _lib.backend().pack_weights(w, mode, dtype, shuffle_r)
    if not capturing:
        cache[key] = (_ver(w), p, False)
    return p


class PackPlan:
    """Every packed operand the parameters `params` have been used with so far (their `_sr_pack` caches), re-packed
    by ONE kernel launch (`sr_pack_weights_batched`) into the same persistent tensors.  The trainer calls `repack()`
    at the start of each step, inside the captured graph, so the ~200 per-use pack launches of a step collapse
    into one per network and the operands always follow the master weights."""

    def __init__(self, params):
        self.entries = []          # (param, key, packed tensor)
        groups = {}                # operand dtype -> [rows, blocks]   (bf16 operands; fp32 ones for the RGB-side thin layers)
        self.ptrs = []
        for w in params:
            if w.dim() != 4 or w.dtype != torch.float32 or not w.is_contiguous():
                continue
            for key, hit in w.__dict__.get("_sr_pack", {}).items():
                mode, dtype, shuffle_r = key
                grp = groups.setdefault(dtype, [[], 0])
                cout, cin, kh, kw = w.shape
                grp[0].append([w.data_ptr(), hit[1].data_ptr(), cout, cin, kh * kw, mode, int(shuffle_r), grp[1]])
                grp[1] += (w.numel() + 1023) // 1024
                self.entries.append((w, key, hit[1]))
                self.ptrs.append(w.data_ptr())
        dev = self.entries[0][0].device if self.entries else None
        # one launch per operand dtype: (dtype, device table, entries, blocks)
        self.tables = [(dt, torch.tensor(rows, dtype=torch.int64, device=dev), len(rows), blocks) for dt, (rows, blocks) in groups.items()]
        self.table = self.tables[0][1] if self.tables else None
        self.dtype = self.tables[0][0] if self.tables else None
        self.blocks = sum(t[3] for t in self.tables)

    def valid(self):
        """master weights still where the table points (FlatAdam views are stable; load_state_dict copies in place)"""
        return all(w.data_ptr() == ptr for (w, _, _), ptr in zip(self.entries, self.ptrs))

    def repack(self):
        if not self.tables:
            return
        for dt, table, n, blocks in self.tables:
            _lib.backend().pack_weights_batched(table, n, blocks, dt)
        for w, key, t in self.entries:
            w.__dict__["_sr_pack"][key] = (_ver(w), t, True)


def to_compute(x):
    """NCHW-shaped tensor in the compute dtype with NHWC memory.
    RGB-side tensors (<= 4 channels: the LR / HR batches, the generator output, the WGAN-GP interpolates) are NOT rounded to
    bf16: the layers that read them are direct FMA kernels (csrc/conv_thin.cu) that take fp32 inputs and fp32-packed weights at
    no cost, and rounding an 8-bit-mantissa image in front of a train-mode BatchNorm stack costs ~2x in the discriminator's
    per-layer error (profiles/r02_parity_fullsize.txt)."""
    lp = getattr(x, "_sr_lowp", None)
    if lp is not None and lp.dtype == config.compute_dtype:
        return lp                      # twin written by the producing kernel's epilogue: no cast pass
    if x.dim() == 4 and x.shape[1] <= 4 and x.dtype == torch.float32:
        return to_input(x)
    if x.dtype != config.compute_dtype:
        x = x.to(config.compute_dtype)
    return x.contiguous(memory_format=torch.channels_last)


def to_input(x):
    """fp32 (N, C <= 4, H, W) tensor with NHWC memory: one layout kernel for the loader's NCHW batches, a no-op for
    tensors the library produced"""
    if x.is_contiguous(memory_format=torch.channels_last):
        return x
    if x.is_cuda and x.is_contiguous() and not x.requires_grad:
        return _lib.backend().nchw_to_nhwc(x, torch.float32)
    return x.contiguous(memory_format=torch.channels_last)


def _out_dtype(x, out_dtype=None):
    """dtype a convolution writes: the caller's choice, else the compute dtype (fp32 RGB-side inputs feed bf16 feature maps)"""
    return out_dtype if out_dtype is not None else (config.compute_dtype if x.dtype == torch.float32 else x.dtype)


# ----------------------------------------------------------------------------------------------
# differentiable-to-any-order convolution primitives
# ----------------------------------------------------------------------------------------------
class ConvFwd(Function):
    """y = conv2d(x, w) + b   (x: NHWC compute dtype; w: OIHW fp32 master; b: fp32 or None)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad):
        g = conv_geom(x.shape, w.shape, stride, pad)
        ctx.g = g
        ctx.has_bias = b is not None
        ctx.bias = b                   # only its identity is used (gradient target lookup)
        ctx.save_for_backward(x, w)
        return _lib.backend().conv_fwd(x, packed(w, 0, x.dtype), b, None, g, out_dtype=_out_dtype(x), impl=config.conv_impl)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        g = ctx.g
        gy = gy.to(_out_dtype(x))
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = ConvDgrad.apply(gy, w, g, x.dtype)
        gw, gb = _wgrad(x, gy, g, w, ctx.bias, ctx.has_bias, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return gx, gw, gb, None, None


class ConvDgrad(Function):
    """dx = conv_transpose2d(dy, w) for the forward geometry g (linear in dy and in w)."""

    @staticmethod
    def forward(ctx, gy, w, g, x_dtype=None):
        ctx.g = g
        ctx.save_for_backward(gy, w)
        return _lib.backend().conv_dgrad(gy, packed(w, 1, gy.dtype), g, out_dtype=x_dtype, impl=config.conv_impl)

    @staticmethod
    def backward(ctx, ggx):
        gy, w = ctx.saved_tensors
        g = ctx.g
        if not (g.Cin <= 4 and ggx.dtype == torch.float32):      # RGB-side cotangents stay fp32 (thin kernels)
            ggx = ggx.to(gy.dtype)
        d_gy = d_w = None
        if ctx.needs_input_grad[0]:
            d_gy = ConvFwd.apply(ggx, w, None, g.stride, g.pad).to(gy.dtype)
        if ctx.needs_input_grad[1]:
            d_w, _ = _wgrad(ggx, gy, g, w, None, False, True, False)
        return d_gy, d_w, None, None


class ConvWgrad(Function):
    """(dw, db) = (sum_pix dy (x) x, sum_pix dy): OIHW fp32 (linear in x and in dy)."""

    @staticmethod
    def forward(ctx, x, gy, g):
        ctx.g = g
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, gy)
        wgrad_join()
        dw, db = _lib.backend().conv_wgrad(x, gy, g, want_bias=True, impl=config.conv_impl)
        return dw, db

    @staticmethod
    def backward(ctx, ggw, ggb):
        x, gy = ctx.saved_tensors
        g = ctx.g
        d_x = d_gy = None
        if ctx.needs_input_grad[0] and ggw is not None:
            d_x = ConvDgrad.apply(gy, ggw, g, x.dtype)
        if ctx.needs_input_grad[1]:
            if ggw is not None:
                d_gy = ConvFwd.apply(x, ggw, ggb, g.stride, g.pad).to(gy.dtype)
            elif ggb is not None:
                d_gy = ggb.to(gy.dtype).view(1, -1, 1, 1).expand_as(gy)
        return d_x, d_gy, None


def conv2d(x, w, b=None, stride=1, pad=0):
    """Plain convolution, differentiable to any order."""
    return ConvFwd.apply(to_compute(x), w, b, stride, pad)


# ----------------------------------------------------------------------------------------------
# fused first-order convolution: act(conv + bias) (+ residual) (-> PixelShuffle)
# ----------------------------------------------------------------------------------------------
class ConvFused(Function):
    """y = PixelShuffle_r( act(conv(x,w)+b) ) + residual — one kernel forward; backward = act' mask,
    unshuffle, dgrad, wgrad.  First-order only (generator / VGG path)."""

    @staticmethod
    def forward(ctx, x, w, b, residual, stride, pad, act, slope, shuffle_r, out_dtype):
        g = conv_geom(x.shape, w.shape, stride, pad)
        ctx.g, ctx.act, ctx.slope, ctx.r = g, act, slope, shuffle_r
        ctx.has_bias, ctx.has_res = b is not None, residual is not None
        ctx.bias = b
        if act != ACT_NONE and residual is not None:
            raise NotImplementedError("ConvFused: activation together with a residual is not used by this model")
        y = _lib.backend().conv_fwd(x, packed(w, 0, x.dtype, shuffle_r), b, residual, g, act, slope, shuffle_r,
                                    out_dtype=_out_dtype(x, out_dtype), impl=config.conv_impl)
        # the activation derivative is recovered from the sign of the output (slope > 0), which is only
        # possible when no residual was added on top
        ctx.save_for_backward(x, w, y if act != ACT_NONE else None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        g = ctx.g
        g_res = gy if (ctx.has_res and ctx.needs_input_grad[3]) else None
        cd = _out_dtype(x)               # dtype of the gradient operands (the compute dtype; x itself may be an fp32 RGB tensor)
        if ctx.act != ACT_NONE or (ctx.r and ctx.r > 1):
            gpre = _lib.backend().act_bwd(gy, y if y is not None else gy, ctx.act, ctx.slope, ctx.r, g, cd)
        elif gy.dtype != cd:
            gpre = _lib.backend().add_cast(gy.contiguous(memory_format=torch.channels_last), None, cd) if gy.is_cuda else gy.to(cd).contiguous(memory_format=torch.channels_last)
        else:
            gpre = gy.contiguous(memory_format=torch.channels_last)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _lib.backend().conv_dgrad(gpre, packed(w, 1, gpre.dtype), g, out_dtype=x.dtype, impl=config.conv_impl)
        gw, gb = _wgrad(x, gpre, g, w, ctx.bias, ctx.has_bias, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return gx, gw, gb, g_res, None, None, None, None, None, None


def conv2d_fused(x, w, b=None, residual=None, stride=1, pad=0, act=ACT_NONE, slope=0.0, shuffle_r=0, out_dtype=None):
    x = to_compute(x)
    if residual is not None:
        od = out_dtype or x.dtype
        residual = residual.to(od).contiguous(memory_format=torch.channels_last)
    return ConvFused.apply(x, w, b, residual, stride, pad, act, slope, shuffle_r, out_dtype)


class ConvActConv(Function):
    """y2 = conv2(act(conv1(x) + b1)) + b2 [+ residual] — the two 3x3 convolutions of a RAB (reference model/sradsgan.py:251-253)
    or of an EDSR ResnetBlock (model/base_networks.py:283-297, with its `torch.add(out, residual)` as the epilogue of conv2) as
    ONE autograd node, so that the backward can multiply conv2's input gradient by act'(y1) inside the epilogue of the
    tensor-core dgrad kernel (sr_conv2d_dgrad_act) instead of a separate pass over the wide tensor."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, act, slope, residual=None, out_dtype=None, want_pool=False):
        g1 = conv_geom(x.shape, w1.shape, 1, 1)
        be = _lib.backend()
        y1 = be.conv_fwd(x, packed(w1, 0, x.dtype), b1, None, g1, act, slope, impl=config.conv_impl)
        g2 = conv_geom(y1.shape, w2.shape, 1, 1)
        pool = None
        if want_pool:                  # conv2's epilogue also emits the CLAM pooling partials of y2 for the chain behind it
            y2, pool = be.conv_fwd(y1, packed(w2, 0, x.dtype), b2, residual, g2, out_dtype=out_dtype, impl=config.conv_impl, want_pool=True)
        else:
            y2 = be.conv_fwd(y1, packed(w2, 0, x.dtype), b2, residual, g2, out_dtype=out_dtype, impl=config.conv_impl)
        ctx.g1, ctx.g2, ctx.act, ctx.slope = g1, g2, act, slope
        ctx.params = (w1, b1, w2, b2)
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, y1, w1, w2)
        if pool is not None:
            ctx.mark_non_differentiable(pool[0], pool[1])
            ctx.pool_rows = pool[2]
            return y2, pool[0], pool[1]
        return y2

    @staticmethod
    @once_differentiable
    def backward(ctx, gy2_in, *_unused):
        x, y1, w1, w2 = ctx.saved_tensors
        _, b1, _, b2 = ctx.params
        be = _lib.backend()
        g_res = gy2_in if (ctx.has_res and ctx.needs_input_grad[7]) else None
        gy2 = gy2_in.to(x.dtype).contiguous(memory_format=torch.channels_last)
        g1 = be.conv_dgrad_act(gy2, packed(w2, 1, x.dtype), ctx.g2, y1, ctx.act, ctx.slope, impl=config.conv_impl)
        gw2, gb2 = _wgrad(y1, gy2, ctx.g2, w2, b2, True, ctx.needs_input_grad[3], ctx.needs_input_grad[4])
        gx = be.conv_dgrad(g1, packed(w1, 1, x.dtype), ctx.g1, impl=config.conv_impl) if ctx.needs_input_grad[0] else None
        gw1, gb1 = _wgrad(x, g1, ctx.g1, w1, b1, True, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return gx, gw1, gb1, gw2, gb2, None, None, g_res, None, None


class MaxPool2x2(Function):
    """nn.MaxPool2d(2, 2) of VGG19 (first-order): one vectorised kernel each way instead of ATen's indices kernels."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return _lib.backend().maxpool2x2_fwd(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        return _lib.backend().maxpool2x2_bwd(gy, x)


def maxpool2x2(x):
    return MaxPool2x2.apply(to_compute(x))


def conv_act_conv(x, conv1, conv2, act, slope, residual=None, out_dtype=None, want_pool=False):
    """conv1 -> activation -> conv2 (+ residual) for two 3x3 / stride 1 / pad 1 Conv2d modules with biases.
    want_pool: conv2's epilogue also emits the CLAM pooling partials of its output (attached as `._sr_pool`)."""
    if residual is not None:
        od = out_dtype or config.compute_dtype
        residual = residual.to(od).contiguous(memory_format=torch.channels_last)
    x = to_compute(x)
    want_pool = bool(want_pool and residual is None and x.is_cuda and x.dtype == torch.bfloat16)
    out = ConvActConv.apply(x, conv1.weight, conv1.bias, conv2.weight, conv2.bias, act, slope, residual, out_dtype, want_pool)
    if isinstance(out, tuple):
        y2, ps, pk = out
        y2._sr_pool = (ps, pk, ps.shape[1])
        return y2
    return out


# ----------------------------------------------------------------------------------------------
# fused local-attention chain (CLAM -> SLAM -> 1x1 conv -> + residual), C = 64
# ----------------------------------------------------------------------------------------------
class LocalAttnChain(Function):
    """(z32, z16[, acc + z][, pooling partials of z16]) = Conv1x1(SLAM(CLAM(x))) + t.  z32 continues the fp32 residual trunk,
    z16 (same values in the compute dtype) feeds the next 3x3 convolution; their gradients — and the gradient of the
    dense-sampling accumulator, when the chain adds its output to one (reference model/sradsgan.py:459) — are summed inside the
    backward kernel.  In fp32 mode only z32 is produced."""

    @staticmethod
    def forward(ctx, x, t, acc, fc1, fc2, w7, W, b, lowp, pool, want_pool):
        z32, z16, sv, acc_out, out_pool = _lib.backend().la_chain_forward(x, t, fc1, fc2, w7, W, b, want_lowp=lowp, pool=pool, acc=acc,
                                                                            want_pool=want_pool)
        ctx.sv = sv
        ctx.layout = (lowp, acc is not None, out_pool is not None)
        ctx.params = (fc1, fc2, w7, W, b)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, fc1, fc2, w7, W)
        outs = [z32]
        if lowp:
            outs.append(z16)
        if acc is not None:
            outs.append(acc_out)
        if out_pool is not None:
            ctx.mark_non_differentiable(out_pool[0], out_pool[1])
            outs += [out_pool[0], out_pool[1]]
        return tuple(outs) if len(outs) > 1 else z32

    @staticmethod
    @once_differentiable
    def backward(ctx, gz32, *rest):
        x, fc1, fc2, w7, W = ctx.saved_tensors
        lowp, has_acc, _ = ctx.layout
        rest = list(rest)
        gz16 = rest.pop(0) if lowp else None
        gacc = rest.pop(0) if has_acc else None
        if gz32 is None and gz16 is None and gacc is None:
            return (None,) * 11
        targets = [_grad_target(p) for p in ctx.params]
        into = targets if all(t is not None for t in targets) else None
        dx, d_fc1, d_fc2, d_w7, dW, db, dz = _lib.backend().la_chain_backward(gz32, gz16, gacc, x, ctx.sv, fc1, fc2, w7, W,
                                                                              want_dz=ctx.needs_input_grad[1], into=into)
        dacc = gacc if ctx.needs_input_grad[2] else None       # d(acc + z)/d(acc) = identity: the gradient passes through untouched
        if into is not None:
            return dx, dz, dacc, None, None, None, None, None, None, None, None
        return dx, dz, dacc, d_fc1, d_fc2, d_w7, dW, db, None, None, None


def local_attn_chain(x, t, ca, sa, conv, acc=None, want_pool=False):
    """x: conv output (compute dtype; its `._sr_pool` partials are used when present), t: residual (fp32 trunk); ca/sa/conv:
    CLAM, SLAM, 1x1 Conv2d modules.  Returns the fp32 trunk tensor — its compute-dtype twin rides along as `._sr_lowp` (picked up
    by to_compute), the pooling partials of the twin as `._sr_pool` when want_pool — or (trunk, acc + trunk) when the
    dense-sampling accumulator `acc` is given."""
    pool = getattr(x, "_sr_pool", None)
    x = to_compute(x)
    t = t.float().contiguous(memory_format=torch.channels_last)
    lowp = config.compute_dtype != torch.float32
    band = lowp and _lib.backend().la_band_path(x)
    if acc is not None and not band:                       # tile kernels (fp32 mode, very large maps): the sum stays an elementwise add
        z = local_attn_chain(x, t, ca, sa, conv)
        return z, acc + z
    out = LocalAttnChain.apply(x, t, acc, ca.fc1.weight, ca.fc2.weight, sa.conv1.weight, conv.weight, conv.bias, lowp,
                               pool if band else None, bool(want_pool and band))
    if not isinstance(out, tuple):
        return out
    out = list(out)
    z32 = out.pop(0)
    if lowp:
        z32._sr_lowp = out.pop(0)
    new_acc = out.pop(0) if acc is not None else None
    if out:
        z32._sr_pool = (out[0], out[1], out[0].shape[1])
        if lowp:
            z32._sr_lowp._sr_pool = z32._sr_pool
    return (z32, new_acc) if acc is not None else z32


# ----------------------------------------------------------------------------------------------
# SGAM position attention without the N x N tensors (bf16 mode; fp32 mode keeps the explicit softmax path)
# ----------------------------------------------------------------------------------------------
class SGAMAttention(Function):
    """y = gamma * (V softmax(Q^T K)^T) + x  (reference model/sradsgan.py:164-176) through the flash-style kernels of
    csrc/sgam.cu.  q, k: (N, 8, H, W), v: (N, 64, H, W) in the compute dtype (bf16); x, y: fp32 trunk tensors."""

    @staticmethod
    def forward(ctx, q, k, v, x, gamma):
        be = _lib.backend()
        m, linv = be.sgam_stats(q, k)
        g32 = gamma.detach().float().contiguous()
        o16, y = be.sgam_pv(q, k, v, row_m=m, row_s=linv, resid=x, gamma=g32)
        ctx.save_for_backward(q, k, v, g32, m, linv, o16)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        q, k, v, g32, m, linv, o16 = ctx.saved_tensors
        be = _lib.backend()
        do16, d, dgamma = be.sgam_bwd_prep(dy, o16, g32)
        dv, _ = be.sgam_pv(k, q, do16, col_m=m, col_s=linv)
        dq = be.sgam_ds(q, k, do16, v, row_m=m, row_s=linv, row_d=d)
        dk = be.sgam_ds(k, q, v, do16, col_m=m, col_s=linv, col_d=d)
        return dq.to(q.dtype), dk.to(k.dtype), dv, dy, dgamma


def sgam_attention(q, k, v, x, gamma):
    return SGAMAttention.apply(to_compute(q), to_compute(k), to_compute(v), x.float().contiguous(memory_format=torch.channels_last), gamma)


# ----------------------------------------------------------------------------------------------
# any-order differentiable fused blocks of the discriminator (first-order passes AND the WGAN-GP double backward
# run on the same kernels): conv+bias+LeakyReLU, and train-mode BatchNorm2d+LeakyReLU
# ----------------------------------------------------------------------------------------------
class ActBwd(Function):
    """gpre = gy * act'(y).  Linear in gy (y only selects the slope), so its own derivative is ActBwd again."""

    @staticmethod
    def forward(ctx, gy, y, act, slope, g):
        ctx.act, ctx.slope, ctx.g = act, slope, g
        ctx.save_for_backward(y)
        return _lib.backend().act_bwd(gy, y, act, slope, 0, g, y.dtype)

    @staticmethod
    def backward(ctx, u):
        (y,) = ctx.saved_tensors
        return ActBwd.apply(u.to(y.dtype), y, ctx.act, ctx.slope, ctx.g), None, None, None, None


class ConvActFwd(Function):
    """y = act(conv(x, w) + b) in ONE kernel, differentiable to any order (backward = ActBwd -> ConvDgrad / ConvWgrad)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, act, slope):
        g = conv_geom(x.shape, w.shape, stride, pad)
        ctx.g, ctx.act, ctx.slope, ctx.has_bias = g, act, slope, b is not None
        ctx.bias = b
        y = _lib.backend().conv_fwd(x, packed(w, 0, x.dtype), b, None, g, act, slope, out_dtype=_out_dtype(x), impl=config.conv_impl)
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        g = ctx.g
        gpre = ActBwd.apply(gy.to(y.dtype), y, ctx.act, ctx.slope, g)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = ConvDgrad.apply(gpre, w, g, x.dtype)
        gw, gb = _wgrad(x, gpre, g, w, ctx.bias, ctx.has_bias, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return gx, gw, gb, None, None, None, None


def conv2d_act(x, w, b, stride, pad, act, slope):
    return ConvActFwd.apply(to_compute(x), w, b, stride, pad, act, slope)


class BNActBwd(Function):
    """(dx, dgamma, dbeta) of y = lrelu(batchnorm_train(x)); its backward is the fused double-backward kernel
    (cotangent of dx only — the penalty never differentiates the parameter gradients)."""

    @staticmethod
    def forward(ctx, gy, x, gamma, save, slope):
        dx, dgamma, dbeta = _lib.backend().bn_act_bwd(gy, x, save, slope)
        ctx.slope = slope
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(gy, x, save, dgamma, dbeta)
        return dx, dgamma, dbeta

    @staticmethod
    @once_differentiable
    def backward(ctx, u, u_dgamma, u_dbeta):
        if u_dgamma is not None or u_dbeta is not None:
            raise NotImplementedError("BNActBwd: second derivatives through the BatchNorm parameter gradients are not used by WGAN-GP")
        if u is None:
            return None, None, None, None, None
        gy, x, save, dgamma, dbeta = ctx.saved_tensors
        d_gy, d_x, d_gamma = _lib.backend().bn_act_bwd_bwd(u, gy, x, save, dgamma, dbeta, ctx.slope)
        return d_gy, d_x, d_gamma, None, None


class BatchNormLeakyReLU(Function):
    """y = lrelu(batchnorm_train(x; gamma, beta)) with running-stat update; differentiable twice."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, slope):
        y, save = _lib.backend().bn_act_fwd(x, gamma, beta, running_mean, running_var, eps, momentum, slope)
        ctx.slope = slope
        ctx.save_for_backward(x, gamma, save)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gamma, save = ctx.saved_tensors
        dx, dgamma, dbeta = BNActBwd.apply(gy.to(x.dtype), x, gamma, save, ctx.slope)
        return dx, dgamma, dbeta, None, None, None, None, None


def bn_leaky_relu(x, bn, slope, bump=True):
    """bn: sradsgan_b200.nn.BatchNorm2d in training mode.  bump=False: the caller advances `num_batches_tracked` itself (the critic does it
    for all its BatchNorm layers with ONE multi-tensor launch per pass instead of one tiny kernel per layer)"""
    y = BatchNormLeakyReLU.apply(to_compute(x), bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum, slope)
    if bump:
        with torch.no_grad():
            bn.num_batches_tracked += 1
    return y


# ----------------------------------------------------------------------------------------------
# loss reductions (csrc/losses.cu): scalar losses as fp32 0-dim tensors on the device, first-order
# ----------------------------------------------------------------------------------------------
class DiffMeanLoss(Function):
    """mean |a - b|^p (p = 1: nn.L1Loss, 2: nn.MSELoss; reference model/sradsgan.py:685-688, :834, :838), gradient to `a` only
    (the target is the HR batch / the detached VGG features)."""

    @staticmethod
    def forward(ctx, a, b, p):
        ctx.p = p
        ctx.save_for_backward(a, b)
        return _lib.backend().diff_mean(a, b, p)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return _lib.backend().diff_mean_bwd(a, b, ctx.p, g), None, None


def diff_mean_loss(a, b, p=1):
    return DiffMeanLoss.apply(a, b.detach(), p)


class MeanLoss(Function):
    """scale * mean(x): GANLoss('wgan-gp') (reference model/sradsgan.py:46-52)"""

    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        ctx.save_for_backward(x)
        return _lib.backend().mean(x, scale)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return _lib.backend().mean_bwd(g, ctx.scale, x), None


def mean_loss(x, scale=1.0):
    return MeanLoss.apply(x, float(scale))


GP_NORMS = {"L2": 0, "L1": 1, "Linf": 2}
GP_PENALTIES = {"LS": 0, "hinge": 1}


class GPPenalty(Function):
    """mean over pixels of (||grad||_p over the colour channels - 1)^2 (or its hinge): reference model/sradsgan.py:623-637.
    Its backward produces the cotangent that enters the double backward through the discriminator."""

    @staticmethod
    def forward(ctx, grad, norm, penalty):
        ctx.norm, ctx.penalty = norm, penalty
        ctx.save_for_backward(grad)
        return _lib.backend().gp_penalty(grad, norm, penalty)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return _lib.backend().gp_penalty_bwd(grad, ctx.norm, ctx.penalty, g), None, None


def gp_penalty(grad, norm="L2", penalty="LS"):
    return GPPenalty.apply(grad, GP_NORMS[norm], GP_PENALTIES[penalty])


def gp_interpolates(real, fake, alpha):
    """alpha * real + (1 - alpha) * fake of detached samples (reference :611) as an fp32 NHWC leaf"""
    return _lib.backend().lerp(real.detach(), fake.detach(), alpha, torch.float32)


# ----------------------------------------------------------------------------------------------
# CGAM channel attention (csrc/cgam.cu)
# ----------------------------------------------------------------------------------------------
class CGAMAttention(Function):
    """y = gamma * (softmax(rowmax(X X^T) - X X^T) X) + x (reference model/sradsgan.py:202-213), fp32; optionally also the
    compute-dtype twin of y for the convolutions that follow."""

    @staticmethod
    def forward(ctx, x, gamma, lowp):
        y32, y16, A = _lib.backend().cgam_fwd(x, gamma, config.compute_dtype if lowp else None)
        ctx.gamma = gamma
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, A, gamma)
        return (y32, y16) if lowp else y32

    @staticmethod
    @once_differentiable
    def backward(ctx, gy32, gy16=None):
        x, A, gamma = ctx.saved_tensors
        if gy32 is None and gy16 is None:
            return None, None, None
        be = _lib.backend()
        if gy32 is None:
            dy = be.add_cast(gy16.contiguous(memory_format=torch.channels_last), None, torch.float32)
        elif gy16 is None:
            dy = gy32
        else:
            dy = be.add_cast(gy32.contiguous(memory_format=torch.channels_last), gy16.contiguous(memory_format=torch.channels_last), torch.float32)
        tg = _grad_target(ctx.gamma)
        dx, dgamma = be.cgam_bwd(dy, x, A, gamma, dgamma_into=tg)
        return dx, dgamma, None


def cgam_attention(x, gamma):
    x = x.float().contiguous(memory_format=torch.channels_last)
    lowp = config.compute_dtype != torch.float32
    out = CGAMAttention.apply(x, gamma, lowp)
    if lowp:
        y32, y16 = out
        y32._sr_lowp = y16
        return y32
    return out


# ----------------------------------------------------------------------------------------------
# Discriminator attention (CBAM after block 6; csrc/cbam.cu): a closed family of primitives whose backward passes are
# each other, so the critic stays differentiable to any order (WGAN-GP differentiates it twice) while every pass over
# the [N, C, H, W] activation is one custom kernel.   full = (N, C, H, W) NHWC; chan = [N, C] fp32; pix = [N, H*W] fp32
# ----------------------------------------------------------------------------------------------
def _need(ctx, i):
    return ctx.needs_input_grad[i]


class Gate3(Function):
    """y = m_p * s_c * x_pc  (x full, s chan, m pix)"""

    @staticmethod
    def forward(ctx, x, s, m):
        ctx.save_for_backward(x, s, m)
        return _lib.backend().cbam_ew(x, x=x, s=s, m=m)

    @staticmethod
    def backward(ctx, g):
        x, s, m = ctx.saved_tensors
        return (Gate3.apply(g, s, m) if _need(ctx, 0) else None,
                RedC.apply(g, x, m) if _need(ctx, 1) else None,
                RedP.apply(g, x, s) if _need(ctx, 2) else None)


class RedC(Function):
    """out_c = sum_p a_pc * b_pc * m_p  (a, b full, m pix) -> chan"""

    @staticmethod
    def forward(ctx, a, b, m):
        ctx.save_for_backward(a, b, m)
        return _lib.backend().cbam_red_c(a, b, m)

    @staticmethod
    def backward(ctx, g):
        a, b, m = ctx.saved_tensors
        return (Gate3.apply(b, g, m) if _need(ctx, 0) else None,
                Gate3.apply(a, g, m) if _need(ctx, 1) else None,
                RedP.apply(a, b, g) if _need(ctx, 2) else None)


class RedP(Function):
    """out_p = sum_c a_pc * b_pc * s_c  (a, b full, s chan) -> pix"""

    @staticmethod
    def forward(ctx, a, b, s):
        ctx.save_for_backward(a, b, s)
        return _lib.backend().cbam_red_p(a, b, s)

    @staticmethod
    def backward(ctx, g):
        a, b, s = ctx.saved_tensors
        g = g.reshape(g.shape[0], -1)
        return (Gate3.apply(b, s, g) if _need(ctx, 0) else None,
                Gate3.apply(a, s, g) if _need(ctx, 1) else None,
                RedC.apply(a, b, g) if _need(ctx, 2) else None)


class PoolHW(Function):
    """AdaptiveAvgPool2d(1) and AdaptiveMaxPool2d(1) of x in one pass -> [2, N, C] fp32 (avg, max); the arg-max pixels are saved
    (reference model/base_networks.py:371-372)"""

    @staticmethod
    def forward(ctx, x):
        pooled, idx = _lib.backend().cbam_pool_hw(x)
        ctx.idx = idx
        ctx.like = x
        ctx.hw = x.shape[2] * x.shape[3]
        return pooled

    @staticmethod
    def backward(ctx, g):
        return BcastScatterHW.apply(g[0] / ctx.hw, g[1], ctx.idx, ctx.like)


class PoolLinHW(Function):
    """[sum_p x_pc, x[idx_c][c]] -> [2, N, C]: PoolHW with the arg-max pixels given (linear in x)"""

    @staticmethod
    def forward(ctx, x, idx):
        be = _lib.backend()
        ctx.idx = idx
        ctx.like = x
        return torch.stack([be.cbam_red_c(x), be.cbam_gather_hw(x, idx)])

    @staticmethod
    def backward(ctx, g):
        return BcastScatterHW.apply(g[0], g[1], ctx.idx, ctx.like), None


class BcastScatterHW(Function):
    """y_pc = a_c + b_c * [p == idx_c] -> full (the adjoint of PoolLinHW)"""

    @staticmethod
    def forward(ctx, a, b, idx, like):
        ctx.idx = idx
        return _lib.backend().cbam_ew(like, a=a, b=b, idx=idx)

    @staticmethod
    def backward(ctx, g):
        r = PoolLinHW.apply(g, ctx.idx)
        return r[0], r[1], None, None


class CPool(Function):
    """q = [mean_c s_c x_pc, max_c s_c x_pc] -> (N, 2, H, W) fp32 NCHW, the input of SpatialAttention's 7x7 convolution
    (reference model/base_networks.py:447-450 applied to ChannelAttention's output s*x without materialising it)"""

    @staticmethod
    def forward(ctx, x, s):
        q, cidx = _lib.backend().cbam_cpool(x, s)
        ctx.cidx = cidx
        ctx.save_for_backward(x, s)
        return q

    @staticmethod
    def backward(ctx, g):
        x, s = ctx.saved_tensors
        n = g.shape[0]
        g0, g1 = g[:, 0].reshape(n, -1), g[:, 1].reshape(n, -1)
        return (CPoolBx.apply(g0, g1, s, ctx.cidx, x) if _need(ctx, 0) else None,
                CPoolBs.apply(g0, g1, x, ctx.cidx) if _need(ctx, 1) else None)


class CPoolLin(Function):
    """CPool with the arg-max channels given (bilinear in x and s)"""

    @staticmethod
    def forward(ctx, x, s, cidx):
        be = _lib.backend()
        ctx.cidx = cidx
        ctx.save_for_backward(x, s)
        n, c, h, w = x.shape
        return torch.stack([be.cbam_red_p(x, None, s, 1.0 / c), be.cbam_gather_c(x, s, cidx)], dim=1).view(n, 2, h, w)

    @staticmethod
    def backward(ctx, g):
        x, s = ctx.saved_tensors
        n = g.shape[0]
        g0, g1 = g[:, 0].reshape(n, -1), g[:, 1].reshape(n, -1)
        return (CPoolBx.apply(g0, g1, s, ctx.cidx, x) if _need(ctx, 0) else None,
                CPoolBs.apply(g0, g1, x, ctx.cidx) if _need(ctx, 1) else None, None)


class CPoolBx(Function):
    """y_pc = s_c * (g0_p / C + g1_p * [c == cidx_p]) -> full (the x-adjoint of CPoolLin)"""

    @staticmethod
    def forward(ctx, g0, g1, s, cidx, like):
        ctx.cidx = cidx
        ctx.save_for_backward(g0, g1, s)
        return _lib.backend().cbam_ew(like, s2=s, g0=g0, g1=g1, cidx=cidx)

    @staticmethod
    def backward(ctx, G):
        g0, g1, s = ctx.saved_tensors
        n = G.shape[0]
        gg = CPoolLin.apply(G, s, ctx.cidx) if (_need(ctx, 0) or _need(ctx, 1)) else None
        return (gg[:, 0].reshape(n, -1) if gg is not None else None, gg[:, 1].reshape(n, -1) if gg is not None else None,
                CPoolBs.apply(g0, g1, G, ctx.cidx) if _need(ctx, 2) else None, None, None)


class CPoolBs(Function):
    """out_c = sum_p x_pc * (g0_p / C + g1_p * [c == cidx_p]) -> chan (the s-adjoint of CPoolLin)"""

    @staticmethod
    def forward(ctx, g0, g1, x, cidx):
        ctx.cidx = cidx
        ctx.save_for_backward(g0, g1, x)
        return _lib.backend().cbam_red_c(x, None, g0, 1.0 / x.shape[1], g1, cidx)

    @staticmethod
    def backward(ctx, G):
        g0, g1, x = ctx.saved_tensors
        n = G.shape[0]
        gg = CPoolLin.apply(x, G, ctx.cidx) if (_need(ctx, 0) or _need(ctx, 1)) else None
        return (gg[:, 0].reshape(n, -1) if gg is not None else None, gg[:, 1].reshape(n, -1) if gg is not None else None,
                CPoolBx.apply(g0, g1, G, ctx.cidx, x) if _need(ctx, 2) else None, None)


class SmallMatmulNT(Function):
    """a [M, K] @ b [N, K]^T in fp32 (the 256 <-> 16 shared MLP of ChannelAttention; strided views, no cuBLAS launch)"""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return _lib.backend().small_gemm_nt(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return (SmallMatmulNT.apply(g, b.t()) if _need(ctx, 0) else None,
                SmallMatmulNT.apply(g.t(), a.t()) if _need(ctx, 1) else None)


def cbam_attention(x, fc1_w, fc2_w, conv7):
    """ChannelAttention(C) followed by SpatialAttention (reference model/base_networks.py:366-383, :441-457) on the compute-dtype
    activation x (N, C, H, W): three passes over x (pooling, channel pooling of s*x, gate application), everything else on
    [N, C] / [N, 2, H, W] tensors."""
    n, c, h, w = x.shape
    cr = fc1_w.shape[0]
    pooled = PoolHW.apply(x)                                                   # [2, N, C]: avg, max
    hid = torch.relu(SmallMatmulNT.apply(pooled.view(2 * n, c), fc1_w.view(cr, c)))
    o = SmallMatmulNT.apply(hid, fc2_w.view(c, cr)).view(2, n, c)
    s = torch.sigmoid(o[0] + o[1])                                             # [N, C]
    q = CPool.apply(x, s)                                                      # (N, 2, H, W) fp32
    m = torch.sigmoid(conv7(q).float()).view(n, h * w)
    return Gate3.apply(x, s, m)
