"""Drop-in for the numeric series operators of ``aegis_engine_core_v2/financial_analysis.py``.

``FinancialPitchAnalyzer`` keeps the reference's constructor and method signatures
(financial_analysis.py:36-226, 368-423) for the numeric parts: SMA, EMA, Bollinger bands, MACD and
the trend / confidence arrays of ``analyze_pitch_financial``.  The string-label loops
(``detect_articulation_bollinger``, ``detect_slides_macd``) and the RSI ghost-note filter operate
on labels / note events and remain with the consumer.
"""
from __future__ import annotations

import numpy as np

from .financial_filters import _run, multi_filter_consensus


class FinancialPitchAnalyzer:
    def __init__(self, sr=22050, hop_length=512):
        self.sr = sr
        self.hop_length = hop_length
        self.ms_per_frame = (hop_length / sr) * 1000

    @staticmethod
    def _check_window(data, window):
        if len(data) < window:  # np.convolve(..., 'same') returns max(n, window) samples: the reference raises here
            raise IndexError(f"boolean index did not match indexed array: series of {len(data)} < window {window}")

    def simple_moving_average(self, data, window=5):
        self._check_window(data, window)
        return _run(data, ["sma"], sma_window=window)["sma"]

    def exponential_moving_average(self, data, span=5):
        return _run(data, ["ema"], ema_span=span)["ema"]

    def bollinger_bands(self, data, window=20, num_std=2):
        self._check_window(data, window)
        out = _run(data, ["boll_ma", "boll_upper", "boll_lower"], boll_window=window, boll_num_std=float(num_std))
        return out["boll_ma"], out["boll_upper"], out["boll_lower"]

    def macd(self, data, fast=12, slow=26, signal=9):
        out = _run(data, ["macd_line", "macd_sig", "macd_hist"], macd_fast=fast, macd_slow=slow, macd_signal=signal)
        return out["macd_line"], out["macd_sig"], out["macd_hist"]

    def bollinger_confidence(self, f0, window=10):
        """The ``confidence`` array of analyze_pitch_financial (financial_analysis.py:404-416)."""
        f0 = np.asarray(f0, dtype=np.float64)
        _, upper, lower = self.bollinger_bands(f0, window)
        bw = upper - lower
        conf = np.zeros_like(f0)
        ok = ~np.isnan(f0) & ~np.isnan(bw)
        conf[ok] = np.where(bw[ok] > 0, 1.0 / (1.0 + bw[ok]), 1.0)
        return conf

    def analyze_pitch_numeric(self, f0, use_advanced_filters=True):
        """``trend`` and ``confidence`` of analyze_pitch_financial (:386-416) without the label loops."""
        f0 = np.asarray(f0, dtype=np.float64)
        if use_advanced_filters:
            trend, _ = multi_filter_consensus(f0, filters=["savgol", "kalman", "holt"])
        else:
            trend = self.exponential_moving_average(f0, span=5)
        return {"trend": trend, "confidence": self.bollinger_confidence(f0, 10)}
