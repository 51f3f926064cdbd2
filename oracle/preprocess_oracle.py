"""CPU oracle for the image-preprocessing row.  TEST INFRASTRUCTURE ONLY (see p2p_oracle.py).

`load_im_flexible` (utils/datasets/preprocess.py:32-60) = PIL decode -> cal_rescale_size (:83-91) ->
transforms.functional.resize(img, (ht, wt), Image.BICUBIC) (:51) -> ToTensor + Normalize (:93-97, ImageNet mean/std).

The resize lives in a third-party dependency that is not under /root/reference: Pillow (unpinned in the reference's
environment.yml; 12.2.0 installed here), src/libImaging/Resample.c.  Its 8-bit path is integer arithmetic; this module
restates the published algorithm in numpy (precompute_coeffs, normalize_coeffs_8bpc with PRECISION_BITS = 22,
ImagingResampleHorizontal_8bpc, ImagingResampleVertical_8bpc) and is PINNED against Pillow itself in
tests/test_cpu_host.py (bit-exact on up-, down- and mixed scaling).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
MEAN = np.array([0.485, 0.456, 0.406], np.float32)       # preprocess.py:93
STD = np.array([0.229, 0.224, 0.225], np.float32)


def cal_rescale_size(image_size, w, h, k_size=2, scale_factor=1 / 16):
    """utils/datasets/preprocess.py:83-91."""
    wt = int(np.floor(w / (max(w, h) / image_size) * scale_factor / k_size) / scale_factor * k_size)
    ht = int(np.floor(h / (max(w, h) / image_size) * scale_factor / k_size) / scale_factor * k_size)
    return wt, ht


def target_size(wo, ho, k_size=2, upsample=16, imsize=None):
    """The size logic of load_im_flexible (preprocess.py:41-48): never up-sample beyond the original."""
    if not (imsize and imsize > 0):
        imsize = max(wo, ho)
    elif imsize > max(wo, ho):
        imsize = max(wo, ho)
    return cal_rescale_size(imsize, wo, ho, k_size, 1.0 / upsample)


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _coefs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc (box = whole image)."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 2.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds, kk = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / fscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        k = [w / ww if ww != 0.0 else w for w in k] + [0.0] * (ksize - xmax)
        kk.append([int(-0.5 + w * (1 << PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << PRECISION_BITS)) for w in k])
        bounds.append((xmin, xmax))
    return bounds, kk


def resize_bicubic_u8(img, wt, ht):
    """PIL `Image.resize((wt, ht), Image.BICUBIC)` of an RGB uint8 [H,W,3] array: horizontal pass, then vertical."""
    ho, wo, _ = img.shape
    cur = img.astype(np.int64)
    if wt != wo:
        bounds, kk = _coefs(wo, wt)
        out = np.zeros((ho, wt, 3), np.int64)
        for xx, (xmin, xmax) in enumerate(bounds):
            acc = np.full((ho, 3), 1 << (PRECISION_BITS - 1), np.int64)
            for x in range(xmax):
                acc += cur[:, xmin + x, :] * kk[xx][x]
            out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
        cur = out
    if ht != ho:
        bounds, kk = _coefs(ho, ht)
        out = np.zeros((ht, cur.shape[1], 3), np.int64)
        for yy, (ymin, ymax) in enumerate(bounds):
            acc = np.full((cur.shape[1], 3), 1 << (PRECISION_BITS - 1), np.int64)
            for y in range(ymax):
                acc += cur[ymin + y] * kk[yy][y]
            out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255)
        cur = out
    return cur.astype(np.uint8)


def to_tensor_normalize(img_u8):
    """ToTensor (float32(u8) / 255) + Normalize ((x - mean) / std in float32) -> [3,H,W] float32."""
    x = img_u8.astype(np.float32) / np.float32(255.0)
    x = (x - MEAN) / STD
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def load_im_flexible_array(img_u8, k_size=2, upsample=16, imsize=None):
    """load_im_flexible on an already decoded RGB array -> ([3,ht,wt] float32, (sx, sy))."""
    ho, wo, _ = img_u8.shape
    wt, ht = target_size(wo, ho, k_size, upsample, imsize)
    return to_tensor_normalize(resize_bicubic_u8(img_u8, wt, ht)), (wo / wt, ho / ht)
