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
        # the CUDA and CPU resampling kernels may round a tie differently: one grey level on a handful of pixels at most
        assert (lr - w_lr).abs().max() <= 1.01 / 255 and (lr != w_lr).float().mean() < 2e-3
        assert (bc - w_bc).abs().max() <= 2.01 / 255 and (bc != w_bc).float().mean() < 4e-3
