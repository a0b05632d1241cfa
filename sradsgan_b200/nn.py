"""Layer classes with torch.nn-compatible constructors / state_dict keys whose arithmetic runs through
the sradsgan_b200 C-ABI kernels.  Class names deliberately contain 'Conv2d' / 'BatchNorm' so that the
reference's `weights_init_normal` (utils/utils.py:97-114, applied with Module.apply at
model/sradsgan.py:713-714) initialises them exactly like torch's own layers."""
import math

import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU


class Conv2d(nn.Module):
    """Drop-in for torch.nn.Conv2d(in, out, k, stride, padding, bias=...) (square kernels, no groups/dilation)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, dilation=1):
        super().__init__()
        if isinstance(kernel_size, (tuple, list)):
            if kernel_size[0] != kernel_size[1]:
                raise NotImplementedError("square kernels only")
            kernel_size = kernel_size[0]
        if dilation != 1:
            raise NotImplementedError("dilation != 1 is not used by SRADSGAN")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):   # same default as torch.nn.Conv2d
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_channels * self.kernel_size * self.kernel_size)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        """plain conv (+bias), differentiable to any order"""
        return ops.conv2d(x, self.weight, self.bias, self.stride, self.padding)

    def fused(self, x, act=ACT_NONE, slope=0.0, residual=None, shuffle_r=0, out_dtype=None):
        """act(conv(x)+b) [+ residual] [-> PixelShuffle(r)] in one kernel (first-order autograd only)"""
        return ops.conv2d_fused(x, self.weight, self.bias, residual, self.stride, self.padding, act, slope,
                                shuffle_r, out_dtype)

    def extra_repr(self):
        return "%d, %d, kernel_size=%d, stride=%d, padding=%d, bias=%s" % (
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding, self.bias is not None)


class LeakyReLU(nn.Module):
    """Placeholder keeping the reference's Sequential indices; fused into the producing conv where possible."""

    def __init__(self, negative_slope=0.01, inplace=False):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, x):
        return torch.nn.functional.leaky_relu(x, self.negative_slope)


class PixelShuffle(nn.Module):
    def __init__(self, upscale_factor):
        super().__init__()
        self.upscale_factor = upscale_factor

    def forward(self, x):
        return torch.nn.functional.pixel_shuffle(x, self.upscale_factor)


class BatchNorm2d(nn.Module):
    """torch.nn.BatchNorm2d(train-mode) semantics: batch statistics (biased var) for normalisation,
    running stats updated with momentum 0.1 and the unbiased variance; fp32 statistics.
    Written with differentiable primitives so the WGAN-GP double backward passes through it."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def forward(self, x):
        xf = x.float()
        if self.training:
            mean = xf.mean(dim=(0, 2, 3))
            var = xf.var(dim=(0, 2, 3), unbiased=False)
            with torch.no_grad():
                n = x.numel() / x.shape[1]
                self.running_mean.mul_(1 - self.momentum).add_(mean, alpha=self.momentum)
                self.running_var.mul_(1 - self.momentum).add_(var * (n / max(n - 1, 1)), alpha=self.momentum)
                self.num_batches_tracked += 1
        else:
            mean, var = self.running_mean, self.running_var
        scale = self.weight * torch.rsqrt(var + self.eps)
        shift = self.bias - mean * scale
        y = xf * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        return y.to(x.dtype)


class ReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x):
        return torch.relu(x)


class MaxPool2d(nn.Module):
    def __init__(self, kernel_size=2, stride=2):
        super().__init__()
        self.kernel_size, self.stride = kernel_size, stride

    def forward(self, x):
        if self.kernel_size == 2 and self.stride == 2 and x.dim() == 4 and x.shape[1] % 8 == 0:
            return ops.maxpool2x2(x)        # VGG19 features[4] / [9]
        return torch.nn.functional.max_pool2d(x, self.kernel_size, self.stride)
