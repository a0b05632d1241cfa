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
