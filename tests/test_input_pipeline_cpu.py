"""CPU: device-side input pipeline — `data.pil_bicubic` against PIL's own BICUBIC resampling (what the
reference's datasets call, data/dataset.py:403-438), and the double-buffered prefetcher against the per-sample PIL path."""
import numpy as np
import pytest
import torch
from PIL import Image

from sradsgan_b200.data import DevicePrefetcher, FolderHRDataset, FolderSRDataset, pil_bicubic, synthesize_lr_bc


def _images(n=3, size=216, seed=0):
    rs = np.random.RandomState(seed)
    noise = (rs.rand(n, size, size, 3) * 255).astype(np.uint8)                       # worst case: overshoot clipped in both passes
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    smooth = np.stack([127 + 120 * np.sin(xx / (7 + c)) * np.cos(yy / (11 + 2 * c)) for c in range(3)], -1)
    return np.concatenate([noise, smooth[None].clip(0, 255).astype(np.uint8)], 0)


@pytest.mark.parametrize("scale", [2, 3, 4, 8, 9])
def test_pil_bicubic_matches_pil(scale):
    imgs = _images()
    x = torch.from_numpy(imgs).permute(0, 3, 1, 2).float()
    lo = 216 // scale
    got = pil_bicubic(x, lo, lo)
    up = pil_bicubic(got, 216, 216)
    for i in range(imgs.shape[0]):
        want = np.asarray(Image.fromarray(imgs[i]).resize((lo, lo), Image.BICUBIC)).astype(np.float32)
        d = np.abs(got[i].permute(1, 2, 0).numpy() - want)
        assert d.max() <= 1 and (d > 0).mean() < 2e-3, (scale, i, d.max(), (d > 0).mean())      # mostly bit-exact
        lr_pil = Image.fromarray(got[i].permute(1, 2, 0).numpy().astype(np.uint8))
        want_up = np.asarray(lr_pil.resize((216, 216), Image.BICUBIC)).astype(np.float32)
        du = np.abs(up[i].permute(1, 2, 0).numpy() - want_up)
        assert du.max() <= 2 and (du > 0).mean() < 2e-3, (scale, i, du.max(), (du > 0).mean())


def test_prefetcher_matches_per_sample_pil_pipeline(tmp_path):
    root = tmp_path / "set"
    root.mkdir()
    imgs = _images(n=4, size=80, seed=3)
    for i, im in enumerate(imgs):
        Image.fromarray(im).save(root / ("img_%02d.png" % i))
    ref = FolderSRDataset(str(tmp_path), ["set"], crop_size=72, scale=4)
    hr_ds = FolderHRDataset(str(tmp_path), ["set"], crop_size=72, scale=4)
    loader = torch.utils.data.DataLoader(hr_ds, batch_size=2, shuffle=False, drop_last=True)
    batches = list(DevicePrefetcher(loader, "cpu", 4))
    assert len(batches) == len(ref) // 2
    k = 0
    for lr, hr, bc, paths in batches:
        assert lr.shape == (2, 3, 18, 18) and hr.shape == (2, 3, 72, 72) and bc.shape == hr.shape
        for j in range(2):
            w_lr, w_hr, w_bc, w_path = ref[k]
            assert paths[j] == w_path
            assert torch.equal(hr[j], w_hr)
            assert (lr[j] - w_lr).abs().max() <= 1.01 / 255 and ((lr[j] != w_lr).float().mean() < 2e-3)
            assert (bc[j] - w_bc).abs().max() <= 2.01 / 255 and ((bc[j] != w_bc).float().mean() < 4e-3)
            k += 1
    lr2, hr2, bc2 = synthesize_lr_bc(torch.stack([hr_ds[0][0], hr_ds[1][0]]), 4)
    assert torch.equal(lr2, batches[0][0]) and torch.equal(bc2, batches[0][2])
