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
        assert np.array_equal(got[i].permute(1, 2, 0).numpy(), want), (scale, i)                 # byte work: bit-exact
        lr_pil = Image.fromarray(got[i].permute(1, 2, 0).numpy().astype(np.uint8))
        want_up = np.asarray(lr_pil.resize((216, 216), Image.BICUBIC)).astype(np.float32)
        assert np.array_equal(up[i].permute(1, 2, 0).numpy(), want_up), (scale, i)


@pytest.mark.parametrize("size,out", [((37, 53), (11, 90)), ((64, 64), (64, 17)), ((9, 200), (31, 7)), ((216, 216), (217, 215))])
def test_pil_bicubic_odd_sizes_uint8(size, out):
    """non-integer ratios, one-axis passes, up and down at once — uint8 in, uint8 out, every pixel equal to PIL's"""
    rs = np.random.RandomState(size[0] + out[1])
    img = (rs.rand(size[0], size[1], 3) * 255).astype(np.uint8)
    got = pil_bicubic(torch.from_numpy(img).permute(2, 0, 1)[None].contiguous(), out[0], out[1])
    want = np.asarray(Image.fromarray(img).resize((out[1], out[0]), Image.BICUBIC))
    assert got.dtype == torch.uint8 and np.array_equal(got[0].permute(1, 2, 0).numpy(), want)


def test_resample_emulation_matches_host_path():
    """oracle/ops_emu.py's statement of sr_resample_u8 (the checker of the CUDA kernel) == the host path == PIL"""
    from oracle import ops_emu
    from sradsgan_b200.data import pil_coeffs
    rs = np.random.RandomState(1)
    img = torch.from_numpy((rs.rand(2, 3, 40, 56) * 255).astype(np.uint8))
    emu = ops_emu.EmuBackend()
    h = emu.resample_u8(img, 14, 0, *pil_coeffs(56, 14))
    v = emu.resample_u8(h, 10, 1, *pil_coeffs(40, 10))
    assert torch.equal(v, pil_bicubic(img, 10, 14))


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
            assert torch.equal(lr[j], w_lr) and torch.equal(bc[j], w_bc)                      # bit-exact vs the per-sample PIL path
            k += 1
    lr2, hr2, bc2 = synthesize_lr_bc(torch.stack([hr_ds[0][0], hr_ds[1][0]]), 4)
    assert torch.equal(lr2, batches[0][0]) and torch.equal(bc2, batches[0][2])
