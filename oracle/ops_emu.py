"""TEST INFRASTRUCTURE ONLY — torch-CPU emulation of each C-ABI entry point (include/sradsgan_b200.h).

Two uses, both in tests/:
  * on the GPU box: per-kernel checker (same packed operands in, compare outputs);
  * in the CPU container: installed with sradsgan_b200._lib.set_backend() so that the host-side autograd
    wiring (sradsgan_b200/ops.py, model/, trainer) can be checked against the oracle without a GPU.
It follows the documented semantics of the ABI (packed weight layouts, subpixel-major PixelShuffle
packing, epilogue order act -> residual), not the kernels' code.
"""
import torch
import torch.nn.functional as F

ACT_NONE, ACT_LRELU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3


def unpack_weights(packed, mode, g_or_shape, shuffle_r=0):
    """inverse of sr_pack_weights -> OIHW fp32"""
    cout, cin, kh, kw = g_or_shape
    p = packed.float()
    if mode == 0:
        w = p.reshape(kh, kw, cout, cin).permute(2, 3, 0, 1)
        if shuffle_r and shuffle_r > 1:
            r2 = shuffle_r * shuffle_r
            cq = cout // r2
            rows = torch.arange(cout)
            orig = (rows % cq) * r2 + rows // cq          # packed row n' holds original channel orig[n']
            w_full = torch.empty_like(w)
            w_full[orig] = w
            w = w_full
    else:
        w = p.reshape(kh, kw, cin, cout).permute(3, 2, 0, 1)
    return w.contiguous()


class EmuBackend:
    name = "emu"

    def __init__(self, band=False):
        self.launches = 0
        self.band = band            # emulate the band path's extras (dense-sampling accumulator, pooling partials) of the chain

    def launch_count(self):
        return self.launches

    def device_check(self):
        return None

    def pack_weights(self, w, mode, dtype, shuffle_r=0):
        self.launches += 1
        w = w.detach().float()
        cout, cin, kh, kw = w.shape
        if mode == 0:
            if shuffle_r and shuffle_r > 1:
                r2 = shuffle_r * shuffle_r
                cq = cout // r2
                rows = torch.arange(cout)
                w = w[(rows % cq) * r2 + rows // cq]
            out = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
        else:
            out = w.permute(2, 3, 1, 0).reshape(kh * kw, cin, cout)
        return out.contiguous().to(dtype)

    def conv_fwd(self, x, w_packed, bias, residual, g, act=ACT_NONE, slope=0.0, shuffle_r=0, out_dtype=None, impl=0, want_pool=False):
        if want_pool:               # the partials only save the chain a pooling pass: the emulation recomputes the pooling from x
            return self.conv_fwd(x, w_packed, bias, residual, g, act, slope, shuffle_r, out_dtype, impl), None
        self.launches += 1
        out_dtype = x.dtype if out_dtype is None else out_dtype
        w = unpack_weights(w_packed, 0, (g.Cout, g.Cin, g.kh, g.kw), shuffle_r)
        y = F.conv2d(x.float(), w, None if bias is None else bias.detach().float(), stride=g.stride, padding=g.pad)
        if act == ACT_LRELU:
            y = F.leaky_relu(y, slope)
        elif act == ACT_RELU:
            y = F.relu(y)
        elif act == ACT_SIGMOID:
            y = torch.sigmoid(y)
        if shuffle_r and shuffle_r > 1:
            y = F.pixel_shuffle(y, shuffle_r)
        if residual is not None:
            y = y + residual.float()
        return y.to(out_dtype).contiguous(memory_format=torch.channels_last)

    def conv_dgrad(self, dy, w_packed_t, g, out_dtype=None, impl=0):
        self.launches += 1
        out_dtype = dy.dtype if out_dtype is None else out_dtype
        w = unpack_weights(w_packed_t, 1, (g.Cout, g.Cin, g.kh, g.kw))
        dx = torch.nn.grad.conv2d_input((g.N, g.Cin, g.H, g.W), w, dy.float(), stride=g.stride, padding=g.pad)
        return dx.to(out_dtype).contiguous(memory_format=torch.channels_last)

    def conv_dgrad_act(self, dy, w_packed_t, g, y_prev, act, slope, impl=0):
        dx = self.conv_dgrad(dy, w_packed_t, g).float()
        yp = y_prev.float()
        dx = torch.where(yp > 0, dx, dx * (slope if act == ACT_LRELU else 0.0))
        return dx.to(dy.dtype).contiguous(memory_format=torch.channels_last)

    def conv_wgrad(self, x, dy, g, want_bias=True, impl=0):
        self.launches += 1
        dw = torch.nn.grad.conv2d_weight(x.float(), (g.Cout, g.Cin, g.kh, g.kw), dy.float(), stride=g.stride, padding=g.pad)
        db = dy.float().sum(dim=(0, 2, 3)) if want_bias else None
        return dw.contiguous(), db

    def conv_wgrad_into(self, x, dy, g, dw, db, impl=0):
        gw, gb = self.conv_wgrad(x, dy, g, want_bias=db is not None)
        dw.add_(gw)
        if db is not None:
            db.add_(gb)

    # fused local-attention chain: emulated with the oracle's own building blocks + torch autograd
    @staticmethod
    def _la_math(x, t, fc1, fc2, w7, W, b):
        xf = x.float()
        avg = F.adaptive_avg_pool2d(xf, 1)
        mx = F.adaptive_max_pool2d(xf, 1)
        gate = torch.sigmoid(F.conv2d(F.relu(F.conv2d(avg, fc1)), fc2) + F.conv2d(F.relu(F.conv2d(mx, fc1)), fc2))
        u = gate * xf
        q = torch.cat([u.mean(1, keepdim=True), u.max(1, keepdim=True)[0]], 1)
        m = torch.sigmoid(F.conv2d(q, w7, padding=3))
        return F.conv2d(m * u, W, b) + t.float()

    def la_chain_fwd(self, x, t, fc1, fc2, w7, W, b, want_lowp=True):
        self.launches += 5
        with torch.no_grad():
            z = self._la_math(x, t, fc1.detach().float(), fc2.detach().float(), w7.detach().float(), W.detach().float(), b.detach().float())
        z32 = z.contiguous(memory_format=torch.channels_last)
        z16 = z32.to(x.dtype) if want_lowp else None
        return z32, z16, {"t_shape": t.shape, "b": b.detach().float()}

    def la_band_path(self, x):
        return self.band

    def la_chain_forward(self, x, t, fc1, fc2, w7, W, b, want_lowp=True, pool=None, acc=None, want_pool=False):
        z32, z16, sv = self.la_chain_fwd(x, t, fc1, fc2, w7, W, b, want_lowp)
        acc_out = (acc.float() + z32).contiguous(memory_format=torch.channels_last) if acc is not None else None
        n, c = x.shape[0], x.shape[1]
        out_pool = (torch.zeros(n, 1, c), torch.zeros(n, 1, c, dtype=torch.int32), 1) if want_pool else None
        return z32, z16, sv, acc_out, out_pool

    def la_chain_backward(self, gz32, gz16, gacc, x, sv, fc1, fc2, w7, W, want_dz=True, into=None):
        if gacc is not None:
            gz32 = gacc.float() if gz32 is None else gz32.float() + gacc.float()
        return self.la_chain_bwd(gz32, gz16, x, sv, fc1, fc2, w7, W, want_dz, into)

    def la_chain_bwd(self, gz32, gz16, x, sv, fc1, fc2, w7, W, want_dz=True, into=None):
        self.launches += 6
        dz = (gz32.float() if gz32 is not None else 0) + (gz16.float() if gz16 is not None else 0)
        with torch.enable_grad():
            xs = x.detach().float().requires_grad_(True)
            ps = [p.detach().float().clone().requires_grad_(True) for p in (fc1, fc2, w7, W)]
            bs = sv["b"].clone().requires_grad_(True)
            z = self._la_math(xs, torch.zeros(sv["t_shape"]), ps[0], ps[1], ps[2], ps[3], bs)
            grads = torch.autograd.grad(z, [xs] + ps + [bs], dz)
        if into is not None:
            for t, gr in zip(into, grads[1:]):
                t.add_(gr)
        return (grads[0].to(x.dtype).contiguous(memory_format=torch.channels_last), grads[1], grads[2], grads[3], grads[4], grads[5],
                dz.contiguous(memory_format=torch.channels_last) if want_dz else None)

    def act_bwd(self, gy, y, act, slope, shuffle_r, g, out_dtype):
        self.launches += 1
        gp = gy.float()
        if act == ACT_LRELU:
            gp = torch.where(y.float() > 0, gp, gp * slope)
        elif act == ACT_RELU:
            gp = torch.where(y.float() > 0, gp, torch.zeros_like(gp))
        if shuffle_r and shuffle_r > 1:
            gp = F.pixel_unshuffle(gp, shuffle_r)
        return gp.to(out_dtype).contiguous(memory_format=torch.channels_last)

    def maxpool2x2_fwd(self, x):
        self.launches += 1
        return F.max_pool2d(x, 2, 2).contiguous(memory_format=torch.channels_last)

    def maxpool2x2_bwd(self, dy, x):
        self.launches += 1
        xr = x.detach().float().requires_grad_(True)
        with torch.enable_grad():
            y = F.max_pool2d(xr, 2, 2)
        (dx,) = torch.autograd.grad(y, xr, dy.float())
        return dx.to(x.dtype).contiguous(memory_format=torch.channels_last)

    def bn_act_fwd(self, x, gamma, beta, running_mean, running_var, eps, momentum, slope):
        self.launches += 3
        xf = x.float()
        mean = xf.mean((0, 2, 3)); var = xf.var((0, 2, 3), unbiased=False)
        n = xf.numel() / xf.shape[1]
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(mean, alpha=momentum)
            running_var.mul_(1 - momentum).add_(var * (n / max(n - 1, 1)), alpha=momentum)
        rstd = torch.rsqrt(var + eps)
        scale = gamma.detach().float() * rstd
        shift = beta.detach().float() - mean * scale
        y = F.leaky_relu(xf * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), slope)
        return y.to(x.dtype).contiguous(memory_format=torch.channels_last), torch.stack([mean, rstd, scale, shift])

    def bn_act_bwd(self, gy, x, save, slope):
        self.launches += 2
        mean, rstd, scale, shift = save[0], save[1], save[2], save[3]
        xf, g = x.float(), gy.float()
        v = lambda t: t.view(1, -1, 1, 1)
        z = xf * v(scale) + v(shift)
        g = torch.where(z > 0, g, g * slope)
        xh = (xf - v(mean)) * v(rstd)
        m = xf.numel() / xf.shape[1]
        dbeta = g.sum((0, 2, 3)); dgamma = (g * xh).sum((0, 2, 3))
        dx = v(scale) * (g - v(dbeta) / m - xh * v(dgamma) / m)
        return dx.to(x.dtype).contiguous(memory_format=torch.channels_last), dgamma, dbeta

    def bn_act_bwd_bwd(self, u, gy, x, save, dgamma, dbeta, slope):
        """autograd through a differentiable restatement of bn_act_bwd (independent of the kernel's closed form)"""
        self.launches += 2
        with torch.enable_grad():
            xs = x.detach().float().requires_grad_(True)
            gs = gy.detach().float().requires_grad_(True)
            rstd0, scale0, shift0 = save[1], save[2], save[3]
            gam = (scale0 / rstd0).detach().requires_grad_(True)
            eps_var = (1.0 / (rstd0 * rstd0)).view(1, -1, 1, 1)               # var + eps of the forward pass
            v = lambda t: t.view(1, -1, 1, 1)
            mean = xs.mean((0, 2, 3)); var = xs.var((0, 2, 3), unbiased=False)
            rstd = torch.rsqrt(v(var) + (eps_var - v(var.detach())))          # same eps, differentiable in x
            xh = (xs - v(mean)) * rstd
            mask = torch.where(x.float() * v(scale0) + v(shift0) > 0, 1.0, slope)
            gz = gs * mask
            dx = v(gam) * rstd * (gz - gz.mean((0, 2, 3), keepdim=True) - xh * (gz * xh).mean((0, 2, 3), keepdim=True))
            d_gy, d_x, d_gamma = torch.autograd.grad((dx * u.float()).sum(), [gs, xs, gam])
        cl = lambda t: t.to(x.dtype).contiguous(memory_format=torch.channels_last)
        return cl(d_gy), cl(d_x), d_gamma

    # -- loss reductions / glue (csrc/losses.cu) ------------------------------------------------------
    def diff_mean(self, a, b, p):
        self.launches += 1
        d = a.float() - b.float()
        return (d.abs() if p == 1 else d * d).mean()

    def diff_mean_bwd(self, a, b, p, g, scale=1.0):
        self.launches += 1
        d = a.float() - b.float()
        k = g.float() * scale / a.numel()
        r = torch.sign(d) * k if p == 1 else 2.0 * d * k
        return r.to(a.dtype)

    def mean(self, x, scale=1.0):
        self.launches += 1
        return x.float().mean() * scale

    def mean_bwd(self, g, scale, like):
        self.launches += 1
        return torch.full_like(like, 1.0) * (g.float() * scale / like.numel()).to(like.dtype)

    @staticmethod
    def _gp_math(grad, norm, penalty):
        g = grad.float()
        n = g.norm(2, 1) if norm == 0 else (g.norm(1, 1) if norm == 1 else g.abs().max(1)[0])
        return ((n - 1) ** 2 if penalty == 0 else torch.relu(n - 1)).mean()

    def gp_penalty(self, grad, norm, penalty):
        self.launches += 1
        with torch.no_grad():
            return self._gp_math(grad, norm, penalty)

    def gp_penalty_bwd(self, grad, norm, penalty, g, scale=1.0):
        self.launches += 1
        with torch.enable_grad():
            gr = grad.detach().float().requires_grad_(True)
            (d,) = torch.autograd.grad(self._gp_math(gr, norm, penalty), gr)
        return (d * (g.float() * scale)).to(grad.dtype).contiguous(memory_format=torch.channels_last)

    def lerp(self, real, fake, alpha, out_dtype):
        self.launches += 1
        a = alpha.float().view(-1, 1, 1, 1)
        return (a * real.float() + ((1 - a) * fake.float())).to(out_dtype).contiguous(memory_format=torch.channels_last)

    def nchw_to_nhwc(self, x, out_dtype):
        self.launches += 1
        return x.to(out_dtype).contiguous(memory_format=torch.channels_last)

    def add_cast(self, a, b, out_dtype):
        self.launches += 1
        r = a.float() if b is None else a.float() + b.float()
        return r.to(out_dtype)

    # -- CGAM (csrc/cgam.cu): the oracle's formula + torch autograd ----------------------------------------
    @staticmethod
    def _cgam_math(x, gamma):
        b, c, h, w = x.shape
        q = x.reshape(b, c, -1)
        e = torch.bmm(q, q.permute(0, 2, 1))
        att = torch.softmax(torch.max(e, -1, keepdim=True)[0].expand_as(e) - e, dim=-1)
        return gamma * torch.bmm(att, q).reshape(b, c, h, w) + x, att

    def cgam_fwd(self, x, gamma, lowp_dtype=None):
        self.launches += 3
        with torch.no_grad():
            y, att = self._cgam_math(x.float().contiguous(), gamma.detach().float())
        y32 = y.contiguous(memory_format=torch.channels_last)
        y16 = y32.to(lowp_dtype) if lowp_dtype is not None and lowp_dtype != torch.float32 else None
        return y32, y16, att

    def cgam_bwd(self, dy, x, A, gamma, dgamma_into=None):
        self.launches += 3
        with torch.enable_grad():
            xs = x.detach().float().contiguous().requires_grad_(True)
            gs = gamma.detach().float().clone().requires_grad_(True)
            y, _ = self._cgam_math(xs, gs)
            dx, dg = torch.autograd.grad(y, [xs, gs], dy.float())
        dx = dx.contiguous(memory_format=torch.channels_last)
        if dgamma_into is not None:
            dgamma_into.add_(dg.reshape(dgamma_into.shape))
            return dx, None
        return dx, dg.reshape(1)

    # -- discriminator attention primitives (csrc/cbam.cu) ------------------------------------------------
    @staticmethod
    def _pc(t):
        """(N, C, H, W) -> [N, P, C] fp32"""
        n, c, h, w = t.shape
        return t.detach().float().permute(0, 2, 3, 1).reshape(n, h * w, c)

    def cbam_ew(self, like, x=None, s=None, m=None, s2=None, g0=None, g1=None, cidx=None, a=None, b=None, idx=None, acc=None):
        self.launches += 1
        n, c, h, w = like.shape
        P = h * w
        y = torch.zeros(n, P, c)
        f = lambda t, shape: None if t is None else t.detach().float().reshape(shape)
        s, s2, a, b = (f(t, (n, 1, c)) for t in (s, s2, a, b))
        m, g0, g1 = (f(t, (n, P, 1)) for t in (m, g0, g1))
        if x is not None:
            y = y + self._pc(x) * (s if s is not None else 1.0) * (m if m is not None else 1.0)
        if s2 is not None:
            wgt = torch.zeros(n, P, c)
            if g0 is not None:
                wgt = wgt + g0 / c
            if g1 is not None:
                wgt = wgt + g1 * torch.nn.functional.one_hot(cidx.long().reshape(n, P), c).float()
            y = y + s2 * wgt
        if a is not None:
            y = y + a
        if b is not None:
            y = y + b * torch.nn.functional.one_hot(idx.long().reshape(n, c), P).float().permute(0, 2, 1)
        if acc is not None:
            y = y + self._pc(acc)
        return y.reshape(n, h, w, c).permute(0, 3, 1, 2).to(like.dtype).contiguous(memory_format=torch.channels_last)

    def cbam_red_c(self, a, b=None, m=None, scale=1.0, g1=None, cidx=None):
        self.launches += 1
        n, c, h, w = a.shape
        P = h * w
        v = self._pc(a)
        if b is not None:
            v = v * self._pc(b)
        wgt = torch.full((n, P, 1), float(scale))
        if m is not None:
            wgt = wgt * m.detach().float().reshape(n, P, 1)
        out = (v * wgt).sum(1)
        if g1 is not None:
            out = out + (v * g1.detach().float().reshape(n, P, 1) * torch.nn.functional.one_hot(cidx.long().reshape(n, P), c).float()).sum(1)
        return out

    def cbam_pool_hw(self, x):
        self.launches += 1
        v = self._pc(x)
        mx, idx = v.max(dim=1)
        # first maximum (torch.max over a dim does not promise it)
        first = (v == mx.unsqueeze(1)).float().argmax(dim=1)
        return torch.stack([v.mean(1), mx]), first.to(torch.int32)

    def cbam_red_p(self, a, b=None, s=None, scale=1.0):
        self.launches += 1
        n, c, h, w = a.shape
        v = self._pc(a)
        if b is not None:
            v = v * self._pc(b)
        if s is not None:
            v = v * s.detach().float().reshape(n, 1, c)
        return v.sum(2) * scale

    def cbam_cpool(self, x, s):
        self.launches += 1
        n, c, h, w = x.shape
        v = self._pc(x) * s.detach().float().reshape(n, 1, c)
        mx = v.max(dim=2)[0]
        first = (v == mx.unsqueeze(2)).float().argmax(dim=2)
        return torch.stack([v.mean(2), mx], dim=1).reshape(n, 2, h, w), first.to(torch.int32)

    def cbam_gather_hw(self, x, idx):
        self.launches += 1
        return torch.gather(self._pc(x), 1, idx.long().unsqueeze(1)).squeeze(1)

    def cbam_gather_c(self, x, s, cidx):
        self.launches += 1
        n, c, h, w = x.shape
        v = self._pc(x)
        if s is not None:
            v = v * s.detach().float().reshape(n, 1, c)
        return torch.gather(v, 2, cidx.long().reshape(n, -1, 1)).squeeze(2)

    def small_gemm_nt(self, a, b):
        self.launches += 1
        return a.detach().float() @ b.detach().float().t()

    def resample_u8(self, img, out_size, axis, bounds, coeffs):
        """integer restatement of Pillow's ImagingResampleHorizontal_8bpc / Vertical_8bpc (libImaging/Resample.c)"""
        self.launches += 1
        x = img.to(torch.int32)
        if axis:
            x = x.transpose(-1, -2)
        ks = coeffs.shape[1]
        idx = (bounds[:, :1].long() + torch.arange(ks)[None]).clamp_(max=x.shape[-1] - 1)          # taps past the count have coefficient 0
        acc = (x[..., idx] * coeffs.to(torch.int32)).sum(-1, dtype=torch.int32) + (1 << 21)
        out = (acc >> 22).clamp_(0, 255).to(torch.uint8)
        return (out.transpose(-1, -2) if axis else out).contiguous()

    def colsum(self, x2d, want_sq=False):
        self.launches += 1
        f = x2d.float()
        return f.sum(0), (f * f).sum(0) if want_sq else None

    def adam_step(self, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale=1.0, clamp=None,
                  step_tensor=None):
        self.launches += 1
        if step_tensor is not None:
            step = int(step_tensor.item())
        g = grad * grad_scale
        exp_avg.mul_(beta1).add_(g, alpha=1 - beta1)
        exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1 = 1 - beta1 ** step
        bc2 = 1 - beta2 ** step
        denom = exp_avg_sq.sqrt() / (bc2 ** 0.5) + eps
        param.addcdiv_(exp_avg, denom, value=-lr / bc1)
        if clamp is not None and clamp[1] > clamp[0]:
            param.clamp_(clamp[0], clamp[1])
