"""Drop-in for ``aegis_engine_core_v2/financial_filters.py``: trend filters on the GPU (float64).

Same names, defaults, NaN conventions and return shapes as the reference's
``FinancialNoiseFilters`` static methods and ``multi_filter_consensus``
(financial_filters.py:25-141, 256-298).  ``atr_filter`` / ``ichimoku_baseline`` /
``stochastic_oscillator`` exist in the reference but are never called by it; they are not on the hot
path and are not provided here.
"""
from __future__ import annotations

import numpy as np
import torch

from . import core
from .librosa_compat import _device


def _run(data, want, **kw):
    x = np.asarray(data, dtype=np.float64)
    if x.ndim != 1:
        raise ValueError("expected a 1-d pitch series")
    if x.size == 0:
        return {k: x.copy() for k in want}
    xd = torch.from_numpy(np.ascontiguousarray(x)).to(_device())[None]
    out = core.trend_filters(xd, want=tuple(want), **kw)
    return {k: v[0].cpu().numpy() for k, v in out.items()}


class FinancialNoiseFilters:
    @staticmethod
    def savitzky_golay(data, window=11, polyorder=3):
        data = np.asarray(data, dtype=np.float64)
        if not np.any(~np.isnan(data)):
            return data
        return _run(data, ["savgol"], savgol_window=window, savgol_polyorder=polyorder)["savgol"]

    @staticmethod
    def kalman_filter(data, process_variance=1e-5, measurement_variance=1e-1):
        data = np.asarray(data, dtype=np.float64)
        if not np.any(~np.isnan(data)):
            return data
        return _run(data, ["kalman"], kalman_q=process_variance, kalman_r=measurement_variance)["kalman"]

    @staticmethod
    def holt_winters(data, alpha=0.3, beta=0.1):
        data = np.asarray(data, dtype=np.float64)
        if not np.any(~np.isnan(data)):
            return data
        return _run(data, ["holt"], holt_alpha=alpha, holt_beta=beta)["holt"]


def multi_filter_consensus(data, filters=["savgol", "kalman", "holt"]):
    """``(consensus, confidence)``: nanmedian of the chosen filters, ``1 / (1 + nanstd)``."""
    data = np.asarray(data, dtype=np.float64)
    chosen = [f for f in ("savgol", "kalman", "holt") if f in filters]
    if not chosen:
        return data, np.ones_like(data)
    if len(chosen) == 3:
        out = _run(data, ["consensus", "consensus_conf"])
        return out["consensus"], out["consensus_conf"]
    # subsets: a filter that was not chosen contributes an all-NaN row, which nanmedian/nanstd ignore
    c, conf = batch_consensus(data[None], chosen)
    return c[0], conf[0]


def batch_consensus(series, filters=("savgol", "kalman", "holt")):
    """Batched form: ``series`` float64 [n_series, n] (numpy or CUDA tensor) -> (consensus, confidence)."""
    is_np = not isinstance(series, torch.Tensor)
    xd = torch.from_numpy(np.ascontiguousarray(series, dtype=np.float64)).to(_device()) if is_np else series
    if set(filters) >= {"savgol", "kalman", "holt"}:
        out = core.trend_filters(xd, want=("consensus", "consensus_conf"))
        c, conf = out["consensus"], out["consensus_conf"]
    else:
        rows = core.trend_filters(xd, want=tuple(filters))
        stacked = torch.stack([rows[f] for f in filters])
        valid = ~torch.isnan(stacked)
        cnt = valid.sum(0)
        z = torch.where(valid, stacked, torch.zeros_like(stacked))
        avg = z.sum(0) / cnt
        sd = torch.sqrt((torch.where(valid, stacked - avg, torch.zeros_like(stacked)) ** 2).sum(0) / cnt)
        c = torch.nanmedian(stacked, dim=0).values if len(filters) == 1 else torch.nanmean(stacked, dim=0)
        conf = 1.0 / (1.0 + sd)
    return (c.cpu().numpy(), conf.cpu().numpy()) if is_np else (c, conf)
