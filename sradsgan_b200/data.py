"""Data sources for the trainer. The tuple contract (lr, hr, bicubic, path) with float [0,1] tensors
follows reference data/dataset.py:386-441; the folder pipeline itself is out of the hot path."""
import os

import torch
import torch.nn.functional as F
from torch.utils.data import Dataset


class SyntheticSRDataset(Dataset):
    """hr ~ U[0,1), lr = bicubic(hr) clamped — generated per index from a seed."""

    def __init__(self, length, crop_size=216, scale=4, seed=1234):
        self.length, self.crop, self.scale, self.seed = length, crop_size, scale, seed

    def __len__(self):
        return self.length

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        hr = torch.rand(3, self.crop, self.crop, generator=g)
        lr = F.interpolate(hr[None], size=self.crop // self.scale, mode="bicubic", align_corners=False).clamp(0, 1)[0]
        bc = F.interpolate(lr[None], size=self.crop, mode="bicubic", align_corners=False).clamp(0, 1)[0]
        return lr, hr, bc, "synthetic_%d" % i


class FolderSRDataset(Dataset):
    """<data_dir>/<dataset>/**.{png,jpg,tif}: center crop -> HR, PIL bicubic -> LR (reference data/dataset.py:403-438)."""

    EXT = (".png", ".jpg", ".jpeg", ".tif", ".tiff", ".bmp")

    def __init__(self, data_dir, names, crop_size, scale, max_samples=None):
        self.files = []
        for n in names:
            root = os.path.join(data_dir, n)
            if not os.path.isdir(root):
                raise FileNotFoundError(root)
            for d, _, fs in sorted(os.walk(root)):
                self.files += [os.path.join(d, f) for f in sorted(fs) if f.lower().endswith(self.EXT)]
        if max_samples:
            self.files = self.files[:max_samples]
        self.crop, self.scale = crop_size, scale

    def __len__(self):
        return len(self.files)

    def __getitem__(self, i):
        from PIL import Image
        import torchvision.transforms as T
        img = Image.open(self.files[i]).convert("RGB")
        hr_img = T.CenterCrop(self.crop)(img)
        lr_img = hr_img.resize((self.crop // self.scale,) * 2, Image.BICUBIC)
        bc_img = lr_img.resize((self.crop,) * 2, Image.BICUBIC)
        tt = T.ToTensor()
        return tt(lr_img), tt(hr_img), tt(bc_img), self.files[i]


# ----------------------------------------------------------------------------------------------
# device-side input pipeline
# ----------------------------------------------------------------------------------------------
_PIL_PRECISION_BITS = 32 - 8 - 2
_coeff_cache = {}


def pil_coeffs(in_size, out_size):
    """Fixed-point BICUBIC coefficient table of one resampling pass, built like Pillow's precompute_coeffs +
    normalize_coeffs_8bpc (src/libImaging/Resample.c; Keys cubic a = -0.5, support 2 scaled by the shrink factor; IEEE double
    arithmetic in the same operation order, truncating casts) -> (bounds int32 [out, 2] = (first tap, tap count),
    coeffs int32 [out, ksize] with 22 fractional bits)."""
    key = (in_size, out_size)
    if key in _coeff_cache:
        return _coeff_cache[key]
    import math
    import numpy as np
    a = -0.5

    def cubic(x):
        x = abs(x)
        if x < 1.0:
            return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
        if x < 2.0:
            return (((x - 5) * x + 8) * x - 4) * a
        return 0.0

    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [cubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << _PIL_PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PIL_PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    out = (torch.from_numpy(bounds), torch.from_numpy(kk))
    _coeff_cache[key] = out
    return out


_dev_coeffs = {}


def _resample_pass(img, out_size, axis):
    """one pass of PIL's 8-bit resampling along W (axis 0) or H (axis 1) of a (..., H, W) uint8 tensor — on a CUDA tensor the
    library kernel (sr_resample_u8), on a host tensor the same integer arithmetic in torch (the reference resamples on the host too)"""
    in_size = img.shape[-2] if axis else img.shape[-1]
    bounds, kk = pil_coeffs(in_size, out_size)
    if img.is_cuda:
        from . import _lib
        key = (in_size, out_size, img.device)
        if key not in _dev_coeffs:
            _dev_coeffs[key] = (bounds.to(img.device), kk.to(img.device))
        return _lib.backend().resample_u8(img, out_size, axis, *_dev_coeffs[key])
    x = img.to(torch.int32)
    if axis:
        x = x.transpose(-1, -2)
    idx = (bounds[:, :1].long() + torch.arange(kk.shape[1])[None]).clamp_(max=x.shape[-1] - 1)
    acc = (x[..., idx] * kk).sum(-1, dtype=torch.int32) + (1 << (_PIL_PRECISION_BITS - 1))
    out = (acc >> _PIL_PRECISION_BITS).clamp_(0, 255).to(torch.uint8)
    return (out.transpose(-1, -2) if axis else out).contiguous()


def pil_bicubic(img_u8, out_h, out_w):
    """`PIL.Image.resize((out_w, out_h), Image.BICUBIC)` of 8-bit images, BIT-EXACT, on the tensor's device.
    img_u8: (N, C, H, W) uint8 tensor, or a float tensor holding the integer values 0..255 (the result has the input's dtype).
    PIL resamples 8-bit images in two separable passes (horizontal, then vertical) with fixed-point coefficients (22 fractional
    bits), rounding and clipping to uint8 after EACH pass (libImaging/Resample.c) — reproduced with the same integer arithmetic
    (`pil_coeffs`, `sr_resample_u8`).  Parity: tests/test_input_pipeline_cpu.py (every pixel equal to PIL's)."""
    h, w = img_u8.shape[-2:]
    y = img_u8 if img_u8.dtype == torch.uint8 else img_u8.to(torch.uint8)
    if w != out_w:
        y = _resample_pass(y, out_w, 0)
    if h != out_h:
        y = _resample_pass(y, out_h, 1)
    return y if img_u8.dtype == torch.uint8 else y.to(img_u8.dtype)


def synthesize_lr_bc(hr_u8, scale):
    """(lr, hr, bc) in [0, 1] float, exactly what the reference's dataset returns per image (data/dataset.py:403-438):
    LR = PIL-bicubic(HR), BC = PIL-bicubic(LR back to the HR size), all `to_tensor`-scaled (/255).  hr_u8: (N, 3, H, W) uint8."""
    h, w = hr_u8.shape[-2:]
    lr = pil_bicubic(hr_u8, h // scale, w // scale)
    bc = pil_bicubic(lr, h, w)
    d = torch.full((1,), 255.0, device=hr_u8.device)      # tensor / tensor: IEEE division like `to_tensor` on the host (a Python-scalar
    return lr.float() / d, hr_u8.float() / d, bc.float() / d   # divisor is turned into a multiplication by the reciprocal on CUDA: 1 ulp off)


class FolderHRDataset(FolderSRDataset):
    """Only the centre-cropped HR image as a uint8 CHW tensor: LR / bicubic images are synthesised on the device by
    `DevicePrefetcher`, so the loader workers do no resampling and the host->device copy carries 1 byte per sample."""

    def __getitem__(self, i):
        from PIL import Image
        import numpy as np
        import torchvision.transforms as T
        img = T.CenterCrop(self.crop)(Image.open(self.files[i]).convert("RGB"))
        return torch.from_numpy(np.asarray(img).copy()).permute(2, 0, 1).contiguous(), self.files[i]


class DevicePrefetcher:
    """Double-buffered input pipeline (the idea of the reference's unused `DataPrefetcher`, data/dataset.py:55-86): batch k+1 is
    copied from pinned host memory on a side stream and turned into (lr, hr, bc) on the device while step k trains.
    `loader` yields (hr_uint8 (N,3,H,W), paths) — e.g. a DataLoader over FolderHRDataset with pin_memory=True.
    Iterating yields (lr, hr, bc, paths) like the reference's loaders."""

    def __init__(self, loader, device, scale):
        self.loader, self.device, self.scale = loader, torch.device(device), scale
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def _stage(self, batch):
        hr_u8, paths = batch
        if self.stream is None:
            return synthesize_lr_bc(hr_u8.to(self.device), self.scale) + (paths,)
        with torch.cuda.stream(self.stream):
            dev = hr_u8.to(self.device, non_blocking=True)
            out = synthesize_lr_bc(dev, self.scale)
        return out + (paths,)

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur = nxt
            if self.stream is not None:
                torch.cuda.current_stream(self.device).wait_stream(self.stream)
                for t in cur[:3]:
                    t.record_stream(torch.cuda.current_stream(self.device))
            try:
                nxt = self._stage(next(it))          # overlaps with the consumer's work on `cur`
            except StopIteration:
                nxt = None
            yield cur
