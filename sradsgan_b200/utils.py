"""Small host-side helpers mirroring the reference's utils/utils.py where the hot path needs them."""
import os
import time
from math import log10

import numpy as np
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


def quantize_u8(img):
    """reference utils/utils.py:169-175: img*255, clamp to [0,255], astype(uint8) (truncation), CHW -> HWC"""
    return (img * 255.0).clamp(0, 255).detach().cpu().numpy().transpose(1, 2, 0).astype(np.uint8)


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
