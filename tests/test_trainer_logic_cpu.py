"""CPU: host logic of the trainer around the hot path — the early-stopping bookkeeping of the reference's train()
(model/sradsgan.py:986-1036), the uint8 validation metrics (:1111-1114), per-class validation (:1393-1601),
`mfe_test_single` (:1603-1640) and the data-parallel sharding of folder datasets.  The C-ABI kernels are replaced by their
emulation (oracle/ops_emu.py); the kernels themselves are checked by the `-m gpu` tests."""
import os
import socket
import sys
import types

import numpy as np
import pytest
import torch

from oracle import ops_emu
from oracle import sradsgan_oracle as O
from sradsgan_b200 import _lib, ops, utils as U
from sradsgan_b200.model.sradsgan import GeneratorResNet, ResGroup, SRADSGAN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def emu():
    prev = _lib.set_backend(ops_emu.EmuBackend())
    prev_dtype = ops.config.compute_dtype
    ops.set_precision("fp32")
    yield
    ops.config.compute_dtype = prev_dtype
    _lib.set_backend(prev)


def _args(**kw):
    base = dict(model_name="SRADSGAN", train_dataset=[], test_dataset=[], crop_size=32, test_crop_size=32, hr_height=32,
                hr_width=32, num_threads=0, num_channels=3, scale_factor=4, epoch=0, num_epochs=1, save_epochs=1,
                batch_size=2, test_batch_size=1, lr=2e-4, b1=0.9, b2=0.999, data_dir="", root_dir="", save_dir="/tmp/sr_logic",
                gpu_mode=True, n_cpu=0, sample_interval=1000, clip_value=0.01, lambda_gp=10, gp=True, penalty_type="LS",
                grad_penalty_Lp_norm="L2", relativeGan=False, loss_Lp_norm="L1", weight_gan=1e-3, weight_content=1e-2,
                max_train_samples=10, precision="fp32", seed=0)
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_update_best_follows_the_reference_elif_chain():
    """reference :986-1003: PSNR, else SSIM, else ERGAS, else LPIPS improvement resets the counter; a validation that
    produced nothing (None) must not count as 'no improvement' (rollback loop on synthetic runs)."""
    best = {"psnr": 0.0, "ssim": 0.0, "ergas": 10000.0, "lpips": 10000.0, "step": 0, "no_improve": 0}
    f = SRADSGAN._update_best
    f(best, None, 0)
    assert best["no_improve"] == 0
    f(best, (20.0, 0.5, 3.0, None), 0)                 # psnr improves (ssim / ergas are NOT updated: elif chain)
    assert (best["psnr"], best["ssim"], best["ergas"], best["step"], best["no_improve"]) == (20.0, 0.0, 10000.0, 0, 0)
    f(best, (19.0, 0.6, 3.5, None), 1)                 # psnr worse, ssim better -> reset
    assert (best["ssim"], best["step"], best["no_improve"]) == (0.6, 1, 0)
    f(best, (19.0, 0.6, 2.5, None), 2)                 # only ergas better (lower)
    assert (best["ergas"], best["step"], best["no_improve"]) == (2.5, 2, 0)
    for e in range(3, 8):
        f(best, (19.0, 0.6, 2.5, None), e)             # nothing improves
    assert best["no_improve"] == 5 and best["step"] == 2


def _skimage_ssim_u8(x, y):
    """restatement of skimage.measure.compare_ssim(x, y, multichannel=True) for uint8 HWC arrays (skimage 0.16 source:
    uniform_filter size 7, K1 .01, K2 .03, data_range 255, sample covariance, crop pad = 3, channel mean)"""
    from scipy.ndimage import uniform_filter
    vals = []
    for c in range(x.shape[2]):
        a, b = x[:, :, c].astype(np.float64), y[:, :, c].astype(np.float64)
        win, ndim = 7, 2
        npix = win ** ndim
        cov_norm = npix / (npix - 1.0)
        ux, uy = uniform_filter(a, size=win), uniform_filter(b, size=win)
        uxx, uyy, uxy = uniform_filter(a * a, size=win), uniform_filter(b * b, size=win), uniform_filter(a * b, size=win)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
        s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
        pad = (win - 1) // 2
        vals.append(s[pad:-pad, pad:-pad].mean())
    return float(np.mean(vals))


def test_eval_metrics_u8_match_the_host_definitions():
    g = torch.Generator().manual_seed(3)
    gt = torch.rand(3, 3, 24, 20, generator=g)
    pred = (gt + 0.08 * torch.randn(3, 3, 24, 20, generator=g))       # also out-of-range values
    m = U.eval_metrics_u8(pred, gt, scale=4)
    for i in range(3):
        a, b = O.quantize_u8(gt[i]), O.quantize_u8(pred[i])            # reference save_img1 / ToPILImage truncation
        mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
        assert abs(m["mse"][i].item() - mse) < 1e-9
        assert abs(m["psnr"][i].item() - 10 * np.log10(255.0 ** 2 / mse)) < 1e-9          # skimage compare_psnr, data_range 255
        erg = 100.0 * np.sqrt(mse / np.mean(a, dtype=np.float64) ** 2 / 3) / 4             # utils/utils.py:954-962
        assert abs(m["ergas"][i].item() - erg) < 1e-9
        assert abs(m["ssim"][i].item() - _skimage_ssim_u8(b, a)) < 1e-9


def _write_pngs(root, names, size=40, seed=0):
    from PIL import Image
    rs = np.random.RandomState(seed)
    os.makedirs(root, exist_ok=True)
    for n in names:
        Image.fromarray(rs.randint(0, 256, (size, size, 3), dtype=np.uint8)).save(os.path.join(root, n))


def test_mfe_test_single_uint8_output_is_the_oracles(emu, tmp_path):
    """reference :1603-1640 + save_img1 (utils/utils.py:169-187): the saved SR image == quantize_u8(oracle generator run on
    the centre crop), bit for bit in fp32 mode; the bicubic baseline image is written next to it."""
    from PIL import Image
    img_dir = str(tmp_path / "in")
    _write_pngs(img_dir, ["a.png"], size=20, seed=5)
    net = SRADSGAN(_args(test_crop_size=12, save_dir=str(tmp_path / "out"), scale_factor=3))
    ng, nb = 1, 1
    sd = O.tie_upsampling(O.make_state(O.generator_spec(3, ng, nb), seed=8, init="fan"))
    torch.save(sd, str(tmp_path / "g.pkl"))
    net.new_generator = lambda: GeneratorResNet(ResGroup, n_residual_blocks=ng, n_basic_blocks=nb, upscale_factor=3)
    out = net.mfe_test_single(os.path.join(img_dir, "a.png"), modelpath=str(tmp_path / "g.pkl"))
    assert tuple(out.shape) == (3, 36, 36)
    import torchvision.transforms as T
    x = T.Compose([T.CenterCrop(12), T.ToTensor()])(Image.open(os.path.join(img_dir, "a.png")))
    with torch.no_grad():
        want = O.quantize_u8(O.generator_forward(sd, x[None], 3, ng, nb)[0])
    got = np.asarray(Image.open(str(tmp_path / "out" / "SR_SRADSGAN_a.png")))
    assert got.shape == want.shape
    mism = int((got != want).sum())
    assert mism <= got.size // 500, "uint8 mismatches: %d of %d" % (mism, got.size)     # emulation == oracle up to fp32 summation order
    assert int(np.abs(got.astype(int) - want.astype(int)).max()) <= 1
    bc = np.asarray(Image.open(str(tmp_path / "out" / "SR_Bicubic_a.png")))
    assert bc.shape == want.shape


def test_validate_by_class_reports_every_class_and_the_total(emu, tmp_path):
    data = str(tmp_path / "data")
    for ci, cname in enumerate(["airplane", "beach", "river"]):
        _write_pngs(os.path.join(data, "UC", cname), ["%s%02d.png" % (cname, i) for i in range(2 + ci)], size=24, seed=ci)
    net = SRADSGAN(_args(data_dir=data, test_dataset=["UC"], crop_size=16, test_crop_size=16, hr_height=16, hr_width=16,
                         save_dir=str(tmp_path / "out"), scale_factor=2))
    net.new_generator = lambda: GeneratorResNet(ResGroup, n_residual_blocks=1, n_basic_blocks=1, upscale_factor=2)
    res = net.mfeNew_validateByClass(epoch=7, save_img=True)
    assert list(res.keys()) == ["airplane", "beach", "river", "Total"]
    assert [res[c]["n"] for c in ("airplane", "beach", "river")] == [2, 3, 4] and res["Total"]["n"] == 9
    tot = sum(res[c]["sr"]["psnr"] * res[c]["n"] for c in ("airplane", "beach", "river")) / 9
    assert abs(res["Total"]["sr"]["psnr"] - tot) < 1e-9
    assert set(res["beach"]["sr"]) == {"mse", "psnr", "ssim", "ergas"} and "bicubic" in res["beach"]
    assert os.path.exists(str(tmp_path / "out" / "validate" / "river" / "river03_x2_7.png"))
    log = open(str(tmp_path / "out" / "logs" / "val_log.txt")).read()
    assert "model: beach" in log and "model: Total" in log and "bicubic_psnr" in log and "sradsgan_ssim" in log


def test_validate_returns_none_without_a_test_set(emu):
    net = SRADSGAN(_args(synthetic_steps=1))
    net.new_generator = lambda: GeneratorResNet(ResGroup, n_residual_blocks=1, n_basic_blocks=1, upscale_factor=4)
    net.build()
    assert net.validate() is None


def test_train_runs_epochs_past_five_without_rollback_on_synthetic_data(emu, tmp_path):
    """with no validation data the rollback heuristic used to fire every 5 epochs and rewind to epoch 1."""
    net = SRADSGAN(_args(synthetic_steps=1, num_epochs=7, crop_size=16, hr_height=16, hr_width=16, save_dir=str(tmp_path / "run"),
                         log_interval=1000))
    net.new_generator = lambda: GeneratorResNet(ResGroup, n_residual_blocks=1, n_basic_blocks=1, upscale_factor=4)
    avg_G, avg_D = net.train()
    assert len(avg_G) == 7 and len(avg_D) == 7
    assert net.lr == 2e-4 and net.optimizer_G.param_groups[0]["lr"] == 2e-4
    assert os.path.exists(str(tmp_path / "run" / "model" / "generator_param_epoch_7.pkl"))


def _shard_worker(rank, world, port, data, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ops_emu as E
        from sradsgan_b200 import _lib as L, ops as P
        from sradsgan_b200.model.sradsgan import SRADSGAN as S
        L.set_backend(E.EmuBackend())
        P.set_precision("fp32")
        net = S(_args(data_dir=data, train_dataset=["T"], crop_size=16, hr_height=16, hr_width=16, batch_size=2, max_train_samples=0))
        loader = net.load_dataset('train', max_samples=0)
        seen = []
        for epoch in range(2):
            net._train_sampler.set_epoch(epoch)
            seen.append(sorted(os.path.basename(p) for b in loader for p in b[3]))
        q.put((rank, seen))
    finally:
        dist.destroy_process_group()


def test_folder_dataset_is_sharded_across_ranks(tmp_path):
    """every rank used to iterate the same shuffled file list — the all-reduce averaged N copies of the
    same gradient.  With the DistributedSampler the ranks' batches of an epoch are disjoint and change with the epoch."""
    import torch.multiprocessing as mp
    data = str(tmp_path / "data")
    _write_pngs(os.path.join(data, "T"), ["im%02d.png" % i for i in range(12)], size=20)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, data, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for epoch in range(2):
        a, b = set(got[0][epoch]), set(got[1][epoch])
        assert len(a) == 6 and len(b) == 6 and not (a & b)
    assert got[0][0] != got[0][1]                      # set_epoch reshuffles
