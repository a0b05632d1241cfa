"""Small host-side helpers mirroring the reference's utils/utils.py where the hot path needs them."""
import os
import time
from math import log10

import torch


def weights_init_normal(m, mean=0.0, std=0.02):
    """reference utils/utils.py:97-114 (the variant `train()` applies, model/sradsgan.py:713-714)."""
    classname = m.__class__.__name__
    if classname.find('Linear') != -1 or classname.find('Conv2d') != -1 or classname.find('ConvTranspose2d') != -1:
        m.weight.data.normal_(mean, std)
        if m.bias is not None:
            m.bias.data.zero_()
    elif classname.find('BatchNorm') != -1:
        m.weight.data.normal_(1.0, 0.02)
        if m.bias is not None:
            m.bias.data.zero_()


def psnr(pred, gt):
    """reference utils/utils.py:700-709 (float [0,1] PSNR after clamping)"""
    diff = pred.clamp(0, 1).double() - gt.clamp(0, 1).double()
    mse = float((diff * diff).mean())
    return 100 if mse == 0 else 10 * log10(1.0 / mse)


def psnr_batch(pred, gt):
    """per-image `psnr` of a batch, computed where the tensors live (validation keeps the images on the
    GPU and reads back ONE small tensor at the end instead of round-tripping every image through the host).
    Returns a float64 tensor of shape (B,); identical values to `psnr` image by image (100 where the MSE is zero)."""
    diff = pred.clamp(0, 1).double() - gt.clamp(0, 1).double()
    mse = (diff * diff).flatten(1).mean(dim=1)
    return torch.where(mse == 0, torch.full_like(mse, 100.0), 10.0 * torch.log10(1.0 / mse.clamp_min(1e-300)))


def ergas_batch(pred, gt, scale=4):
    """reference utils/utils.py:954-962 (`compare_ergas2`, the variant the trainer logs) on [0,255]-quantised images, per
    image, on the device: 100 * sqrt(mse / mean(img1)^2 / channels) / scale with img1 = ground truth."""
    a = (gt * 255.0).clamp(0, 255).floor().double()
    b = (pred * 255.0).clamp(0, 255).floor().double()
    mse = ((a - b) ** 2).flatten(1).mean(dim=1)
    mean2 = a.flatten(1).mean(dim=1) ** 2
    return 100.0 * torch.sqrt(mse / mean2.clamp_min(1e-300) / pred.shape[1]) / scale


def to_u8_levels(img):
    """grey levels 0..255 as float64 on the tensor's device: `*255 -> clamp -> truncate`, what `ToPILImage` / `save_img1`
    make of a float image (reference model/sradsgan.py:1103-1110, utils/utils.py:169-175; out-of-range generator outputs are
    clamped where the reference's `.byte()` would wrap)."""
    return (img.detach().float() * 255.0).clamp(0, 255).floor().double()


def ssim_u8_batch(a, b):
    """skimage.measure.compare_ssim(img1, img2, multichannel=True) of uint8 images (reference model/sradsgan.py:1113, :1483),
    per image, on the device.  a, b: (B, C, H, W) grey levels 0..255 (float64).  skimage defaults: 7x7 uniform window,
    K1 = 0.01, K2 = 0.03, data_range 255, sample covariance (x 49/48), mean over the pixels whose window lies inside the image
    (its 3-pixel border crop makes the filter's boundary mode irrelevant) and over the channels."""
    import torch.nn.functional as F
    win, c1, c2 = 7, (0.01 * 255.0) ** 2, (0.03 * 255.0) ** 2
    norm = win * win / (win * win - 1.0)
    mean = lambda t: F.avg_pool2d(t, win, stride=1)
    ux, uy = mean(a), mean(b)
    vx = norm * (mean(a * a) - ux * ux)
    vy = norm * (mean(b * b) - uy * uy)
    vxy = norm * (mean(a * b) - ux * uy)
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    return s.flatten(1).mean(dim=1)


def eval_metrics_u8(pred, gt, scale=4):
    """the four per-image numbers the reference's validation loops log (model/sradsgan.py:1111-1114, :1481-1484), computed on
    the device from the uint8-quantised images: skimage compare_mse / compare_psnr (data_range 255) / compare_ssim and
    utils.compare_ergas2.  pred, gt: (B, C, H, W) float [0, 1].  Returns a dict of float64 tensors of shape (B,)."""
    a, b = to_u8_levels(gt), to_u8_levels(pred)
    mse = ((a - b) ** 2).flatten(1).mean(dim=1)
    ps = torch.where(mse == 0, torch.full_like(mse, float("inf")), 10.0 * torch.log10(255.0 ** 2 / mse.clamp_min(1e-300)))
    mean2 = a.flatten(1).mean(dim=1) ** 2
    ergas = 100.0 * torch.sqrt(mse / mean2.clamp_min(1e-300) / pred.shape[1]) / scale
    return {"mse": mse, "psnr": ps, "ssim": ssim_u8_batch(a, b), "ergas": ergas}


def quantize_u8(img):
    """reference utils/utils.py:169-175: img*255, clamp to [0,255], astype(uint8) (truncation), CHW -> HWC.
    The quantisation runs where the tensor lives; only the uint8 image crosses to the host."""
    return (img.detach() * 255.0).clamp(0, 255).to(torch.uint8).permute(1, 2, 0).contiguous().cpu().numpy()


def save_img1(img, save_dir, img_path, cuda=True):
    os.makedirs(save_dir, exist_ok=True)
    arr = quantize_u8(img)
    from PIL import Image
    Image.fromarray(arr).save(img_path)


def mkdir_and_rename(path):
    """reference utils/utils.py:830-838: an existing experiment directory is renamed with a timestamp"""
    if os.path.exists(path):
        new_name = path + '_archived_' + time.strftime('%y%m%d-%H%M%S')
        print('Path already exists. Rename it to [{:s}]'.format(new_name))
        os.rename(path, new_name)
    os.makedirs(path)


class CsvLogger:
    """Replaces the reference's TF1 summary writer (utils/logger.py) with plain text files of the same names."""

    def __init__(self, log_dir):
        os.makedirs(log_dir, exist_ok=True)
        self.loss_path = os.path.join(log_dir, 'loss_log.txt')
        self.val_path = os.path.join(log_dir, 'val_log.txt')

    def scalar_summary(self, tag, value, step):
        with open(self.loss_path, 'a') as f:
            f.write('%d,%s,%.8g\n' % (step, tag, value))

    def print_format_results(self, mode, rlt):
        msg = ' '.join('%s: %s' % (k, ('%.4e' % v) if isinstance(v, float) else v) for k, v in rlt.items())
        print(msg)
        with open(self.loss_path if mode == 'train' else self.val_path, 'a') as f:
            f.write(msg + '\n')
