"""Image-quality metrics of the reference's evaluation scripts, restated for the parity gates of north_star
("PSNR/SSIM/SUV metrics identical to 0.05 dB / 1e-3").  TEST INFRASTRUCTURE ONLY (tests/, bench parity): the product
never imports this file.

* ``psnr``            inference/test_metrics.py:378-399 -> skimage.metrics.peak_signal_noise_ratio
                      = 10 log10(R^2 / MSE), R = max - min of the (cropped) ground truth, MSE in float64.
* ``scale12bit``      inference/test_metrics.py:21-26.
* ``ms_ssim``         inference/test_metrics.py:249-274 -> torchmetrics multiscale_structural_similarity_index_measure
                      (defaults: 11x11 gaussian window sigma 1.5, k1 0.01, k2 0.03, betas (0.0448, 0.2856, 0.3001, 0.2363,
                      0.1333), normalize='relu', data_range = the larger of the two images' ranges), per slice on
                      scale12bit volumes, averaged.  skimage / torchmetrics are absent from the image and unpinned by
                      the reference: this is a restatement of their published definitions, checked against closed forms
                      (tests/test_host_logic.py) — "parity unpinned" for the library arithmetic itself.
* ``inverse_normalize``, ``pet_to_suv``  PET_SUV_calculate/pred_suv_denorm_from_gold_minmax.py:83-87, 126-141.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch
import torch.nn.functional as F


def psnr(true_img: np.ndarray, pred_img: np.ndarray) -> float:
    t, p = true_img.astype(np.float64), pred_img.astype(np.float64)
    data_range = t.max() - t.min()
    mse = np.mean((t - p) ** 2, dtype=np.float64)
    return float(10 * np.log10((data_range ** 2) / mse))


def scale12bit(img: np.ndarray) -> np.ndarray:
    return np.clip(((img - np.mean(img)) / (np.std(img) / 400.)) + 2048., 1e-10, 4095)


def _gauss_window(size: int, sigma: float) -> torch.Tensor:
    d = torch.arange((1 - size) / 2, (1 + size) / 2, 1, dtype=torch.float64)
    g = torch.exp(-(d / sigma) ** 2 / 2)
    g = g / g.sum()
    return (g[:, None] @ g[None, :])


def _ssim_and_cs(p: torch.Tensor, t: torch.Tensor, data_range: float, sigma=1.5, k1=0.01, k2=0.03):
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    size = int(3.5 * sigma + 0.5) * 2 + 1
    pad = (size - 1) // 2
    p = F.pad(p, (pad, pad, pad, pad), mode="reflect")
    t = F.pad(t, (pad, pad, pad, pad), mode="reflect")
    k = _gauss_window(size, sigma).to(p.dtype)[None, None]
    mu_p, mu_t = F.conv2d(p, k), F.conv2d(t, k)
    s_pp = F.conv2d(p * p, k) - mu_p ** 2
    s_tt = F.conv2d(t * t, k) - mu_t ** 2
    s_pt = F.conv2d(p * t, k) - mu_p * mu_t
    upper, lower = 2 * s_pt + c2, s_pp + s_tt + c2
    full = ((2 * mu_p * mu_t + c1) * upper) / ((mu_p ** 2 + mu_t ** 2 + c1) * lower)
    crop = (Ellipsis, slice(pad, -pad), slice(pad, -pad))
    return full[crop].mean().item(), (upper / lower)[crop].mean().item()


def ms_ssim_slice(pred: np.ndarray, true: np.ndarray,
                  betas: Sequence[float] = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)) -> float:
    """One 2-D slice (H, W > 160)."""
    p = torch.from_numpy(np.ascontiguousarray(pred)).double()[None, None]
    t = torch.from_numpy(np.ascontiguousarray(true)).double()[None, None]
    data_range = float(max(p.max() - p.min(), t.max() - t.min()))
    vals = []
    for i in range(len(betas)):
        sim, cs = _ssim_and_cs(p, t, data_range)
        vals.append(cs if i < len(betas) - 1 else sim)
        p, t = F.avg_pool2d(p, (2, 2)), F.avg_pool2d(t, (2, 2))
    out = 1.0
    for v, b in zip(vals, betas):
        out *= max(v, 0.0) ** b
    return out


def ms_ssim(true_vol: np.ndarray, pred_vol: np.ndarray) -> float:
    """ssim_torch on [slices, H, W] volumes: scale12bit each volume, MS-SSIM per slice, mean."""
    t, p = scale12bit(true_vol.astype(np.float64)), scale12bit(pred_vol.astype(np.float64))
    return float(np.mean([ms_ssim_slice(p[i], t[i]) for i in range(t.shape[0])]))


def inverse_normalize(pred_arr: np.ndarray, min_v: float, max_v: float, clip_norm: bool = True) -> np.ndarray:
    pred = pred_arr.astype(np.float32)
    if clip_norm:
        pred = np.clip(pred, -1.0, 1.0)
    return (pred + 1.0) * 0.5 * (max_v - min_v) + min_v


def pet_to_suv(pet: np.ndarray, weight_kg: float, total_dose_bq: float, half_life_s: float, decay_time_s: float,
               slope: float = 1.0, intercept: float = 0.0, use_rescale: bool = True) -> np.ndarray:
    decay_dose = float(total_dose_bq) * pow(2.0, -float(decay_time_s) / float(half_life_s))
    factor = (1000.0 * float(weight_kg)) / decay_dose
    pet_f = pet.astype(np.float32)
    if use_rescale:
        pet_f = pet_f * float(slope) + float(intercept)
    return (pet_f * factor).astype(np.float32)


def suv_mean(arr: np.ndarray) -> float:
    pos = arr[arr > 0]
    return float(np.mean(pos)) if pos.size else 0.0
