"""GPU: the double-buffered device-side input pipeline on a CUDA stream == the same synthesis on the host."""
import numpy as np
import pytest
import torch
from PIL import Image

from sradsgan_b200.data import DevicePrefetcher, FolderHRDataset, synthesize_lr_bc

pytestmark = pytest.mark.gpu


def test_prefetcher_on_cuda_matches_host(tmp_path):
    root = tmp_path / "set"
    root.mkdir()
    rs = np.random.RandomState(5)
    for i in range(6):
        Image.fromarray((rs.rand(80, 80, 3) * 255).astype(np.uint8)).save(root / ("img_%02d.png" % i))
    ds = FolderHRDataset(str(tmp_path), ["set"], crop_size=72, scale=3)
    loader = torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False, drop_last=True, pin_memory=True)
    got = [(lr.cpu(), hr.cpu(), bc.cpu(), paths) for lr, hr, bc, paths in DevicePrefetcher(loader, "cuda", 3)]
    assert len(got) == 3
    for k, (lr, hr, bc, paths) in enumerate(got):
        hr_u8 = torch.stack([ds[2 * k][0], ds[2 * k + 1][0]])
        w_lr, w_hr, w_bc = synthesize_lr_bc(hr_u8, 3)
        assert list(paths) == [ds.files[2 * k], ds.files[2 * k + 1]]
        assert torch.equal(hr, w_hr)
        assert torch.equal(lr, w_lr) and torch.equal(bc, w_bc)            # integer resampling: bit-exact on both devices


@pytest.mark.parametrize("size,out", [((216, 216), (54, 54)), ((216, 216), (24, 24)), ((37, 53), (11, 90)), ((9, 200), (31, 7))])
def test_resample_kernel_matches_pil(size, out):
    """sr_resample_u8 (two passes) == PIL.Image.resize(..., BICUBIC), every byte"""
    from sradsgan_b200.data import pil_bicubic
    rs = np.random.RandomState(size[1] + out[0])
    imgs = (rs.rand(3, size[0], size[1], 3) * 255).astype(np.uint8)
    x = torch.from_numpy(imgs).permute(0, 3, 1, 2).contiguous().cuda()
    got = pil_bicubic(x, out[0], out[1]).cpu()
    for i in range(3):
        want = np.asarray(Image.fromarray(imgs[i]).resize((out[1], out[0]), Image.BICUBIC))
        assert np.array_equal(got[i].permute(1, 2, 0).numpy(), want)
