"""Drop-in for the numeric series operators of ``aegis_engine_core_v2/financial_analysis.py``.

``FinancialPitchAnalyzer`` keeps the reference's constructor and method signatures
(financial_analysis.py:36-226, 368-423) for the numeric parts: SMA, EMA, Bollinger bands, MACD and
the trend / confidence arrays of ``analyze_pitch_financial``.  On the hot path the string-label loops and the RSI
ghost-note filter run inside kernel K8; ``detect_articulation_bollinger`` / ``detect_slides_macd`` are also provided here
as host label loops over the GPU-computed bands / MACD series (financial_analysis.py:148-196, 228-271), which is what the
``use_financial=False`` branch of ``get_midi_events_financial`` calls per note event.
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

    def detect_articulation_bollinger(self, f0, window=10, sensitivity=2.0):
        """Per-frame labels from the band position of ``f0`` (financial_analysis.py:148-196): 'bend' above the upper band,
        'noise' below the lower one, 'vibrato' once the side has changed twice in a row, else 'normal'; None at NaN."""
        f0 = np.asarray(f0, dtype=np.float64)
        _, upper, lower = self.bollinger_bands(f0, window, sensitivity)
        labels, prev, flips = [], 0, 0          # side: +1 above, -1 below, 0 inside
        for x, up, lo in zip(f0, upper, lower):
            if np.isnan(x):
                labels.append(None)
                continue
            side = 1 if x > up else (-1 if x < lo else 0)
            flips = flips + 1 if (prev != side and prev != 0) else 0
            labels.append("vibrato" if flips >= 2 else ("bend" if side > 0 else ("noise" if side < 0 else "normal")))
            prev = side
        return labels

    def detect_slides_macd(self, f0, threshold=0.5):
        """Per-frame slide labels from the MACD (5 / 20 / 9) of the MIDI-pitch series (financial_analysis.py:228-271)."""
        from .librosa_compat import hz_to_midi

        f0 = np.asarray(f0, dtype=np.float64)
        semis = np.full_like(f0, np.nan)
        ok = ~np.isnan(f0)
        if np.any(ok):
            semis[ok] = hz_to_midi(f0[ok])
        line, _, hist = self.macd(semis, fast=5, slow=20, signal=9)
        out = []
        for m, h in zip(line, hist):
            if np.isnan(m):
                out.append(None)
            elif m > threshold and h > 0:
                out.append("slide_up")
            elif m < -threshold and h < 0:
                out.append("slide_down")
            else:
                out.append("normal")
        return out

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
