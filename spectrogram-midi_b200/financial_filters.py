"""Drop-in for ``aegis_engine_core_v2/financial_filters.py``: trend filters on the GPU (float64).

Same names, defaults, NaN conventions and return shapes as the reference's
``FinancialNoiseFilters`` static methods and ``multi_filter_consensus``
(financial_filters.py:25-141, 256-298).  ``atr_filter`` / ``ichimoku_baseline`` /
``stochastic_oscillator`` (financial_filters.py:144-249) exist in the reference but are never called by
it: they are provided for API completeness as host numpy (not on the hot path, no kernel), pinned by
golden vectors of the real file.
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


    # ---- defined by the reference, called by nothing in it (financial_filters.py:144-249): host numpy
    @staticmethod
    def atr_filter(data, window=14, threshold=2.0):
        """``(filtered, noise_mask)``: a frame whose jump from its predecessor exceeds ``threshold`` times the mean
        absolute jump of the ``window`` frames before it is noise and repeats the previous (filtered) value."""
        data = np.asarray(data, dtype=np.float64)
        n = len(data)
        if not np.any(~np.isnan(data)):
            return data, np.zeros(n, dtype=bool)
        jump = np.abs(np.diff(data))
        atr = np.full(n, np.nan)
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)   # nanmean of an all-NaN window is NaN, as in the reference
            for i in range(window, len(jump)):
                atr[i] = np.nanmean(jump[i - window:i])
        with np.errstate(invalid="ignore"):
            step = np.concatenate(([np.nan], np.abs(data[1:] - data[:-1])))
            noise = ~np.isnan(atr) & ~np.isnan(data) & (step > atr * threshold)
        noise[0] = False
        filtered = data.copy()
        for i in np.flatnonzero(noise):      # ascending: a run of noise frames keeps repeating the last good value
            filtered[i] = filtered[i - 1]
        return filtered, noise

    @staticmethod
    def _window_extremes(data, i, length, inclusive):
        w = data[max(0, i - length):i + 1 if inclusive else i]
        return w[~np.isnan(w)]

    @staticmethod
    def ichimoku_baseline(data, tenkan=9, kijun=26):
        """Kijun-sen: midpoint of the highest and lowest valid value of the ``kijun`` frames before each frame (NaN where
        there is none, and for the first ``kijun`` frames).  ``tenkan`` is accepted and, as in the reference, unused in
        the result."""
        data = np.asarray(data, dtype=np.float64)
        if not np.any(~np.isnan(data)):
            return data
        base = np.full_like(data, np.nan)
        for i in range(kijun, len(data)):
            w = FinancialNoiseFilters._window_extremes(data, i, kijun, inclusive=False)
            if len(w):
                base[i] = (np.max(w) + np.min(w)) / 2
        return base

    @staticmethod
    def stochastic_oscillator(data, k_period=14, smooth=3):
        """%D line: position of each value in the range of its last ``k_period`` + 1 frames (0..100, 50 where
        undefined), then the mean over the last ``smooth`` + 1 frames."""
        data = np.asarray(data, dtype=np.float64)
        if not np.any(~np.isnan(data)):
            return np.full_like(data, 50.0)
        k = np.full_like(data, 50.0)
        for i in range(k_period, len(data)):
            w = FinancialNoiseFilters._window_extremes(data, i, k_period, inclusive=True)
            if len(w):
                lo, hi = np.min(w), np.max(w)
                if hi - lo > 0:
                    k[i] = ((data[i] - lo) / (hi - lo)) * 100
        d = np.full_like(k, 50.0)
        for i in range(smooth, len(k)):
            d[i] = np.mean(k[i - smooth:i + 1])
        return d


def multi_filter_consensus(data, filters=["savgol", "kalman", "holt"]):
    """``(consensus, confidence)``: nanmedian of the chosen filters, ``1 / (1 + nanstd)``."""
    data = np.asarray(data, dtype=np.float64)
    chosen = [f for f in ("savgol", "kalman", "holt") if f in filters]
    if not chosen:
        return data, np.ones_like(data)
    out = _run(data, ["consensus", "consensus_conf"], consensus_filters=tuple(chosen))
    return out["consensus"], out["consensus_conf"]


def batch_consensus(series, filters=("savgol", "kalman", "holt")):
    """Batched form: ``series`` float64 [n_series, n] (numpy or CUDA tensor) -> (consensus, confidence); any subset of
    the three filters votes (kernel K5, ``consensus_mask``)."""
    is_np = not isinstance(series, torch.Tensor)
    xd = torch.from_numpy(np.ascontiguousarray(series, dtype=np.float64)).to(_device()) if is_np else series
    out = core.trend_filters(xd, want=("consensus", "consensus_conf"), consensus_filters=tuple(filters))
    c, conf = out["consensus"], out["consensus_conf"]
    return (c.cpu().numpy(), conf.cpu().numpy()) if is_np else (c, conf)
