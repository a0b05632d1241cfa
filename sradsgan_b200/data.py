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
def pil_bicubic(img_u8, out_h, out_w):
    """`PIL.Image.resize((out_w, out_h), Image.BICUBIC)` of uint8-valued images, on the tensor's device.
    img_u8: (N, C, H, W) float tensor holding integer values 0..255.  PIL resamples 8-bit images in two separable
    antialiased passes (horizontal, then vertical; Keys cubic a = -0.5 with its support scaled by the shrink factor) and
    rounds + clips to uint8 after EACH pass — reproduced here with two 1-D `interpolate(..., antialias=True)` calls.
    Parity (tests/test_input_pipeline_cpu.py): identical on >99.8 % of the pixels (bit-exact on most images); the rest differ
    by 1 grey level (2 after a down + up round trip) — PIL evaluates the filter with fixed-point coefficients."""
    h, w = img_u8.shape[-2:]
    y = img_u8.float()
    if w != out_w:
        y = F.interpolate(y, size=(h, out_w), mode="bicubic", antialias=True, align_corners=False)
        y = torch.floor(y + 0.5).clamp_(0, 255)
    if h != out_h:
        y = F.interpolate(y, size=(out_h, out_w), mode="bicubic", antialias=True, align_corners=False)
        y = torch.floor(y + 0.5).clamp_(0, 255)
    return y


def synthesize_lr_bc(hr_u8, scale):
    """(lr, hr, bc) in [0, 1] float, exactly what the reference's dataset returns per image (data/dataset.py:403-438):
    LR = PIL-bicubic(HR), BC = PIL-bicubic(LR back to the HR size), all `to_tensor`-scaled (/255).  hr_u8: (N, 3, H, W) uint8."""
    hr = hr_u8.float()
    h, w = hr.shape[-2:]
    lr = pil_bicubic(hr, h // scale, w // scale)
    bc = pil_bicubic(lr, h, w)
    d = hr.new_full((1,), 255.0)      # tensor / tensor: IEEE division like `to_tensor` on the host (a Python-scalar divisor is turned
    return lr / d, hr / d, bc / d      # into a multiplication by the reciprocal on CUDA: 1 ulp off)


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
