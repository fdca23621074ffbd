"""Drop-in for ``aegis_engine_core_v2/guitar_specific.py``: electric-guitar filters on the GPU (kernel K6).

Same names, arguments, defaults and return values as the reference (``GuitarSpecificFilters`` static methods
:24-233 and ``apply_guitar_filters`` :240-277, called from aegis_engine_financial.py:132-147): numpy in, numpy
out.  ``detect_hammer_on_pull_off`` (:153-207) is defined by the reference but never called; it is an event-list
builder (consumer side) and is not part of this path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import core
from .librosa_compat import _device, midi_to_hz


def _img(S_dB):
    S = np.asarray(S_dB)
    if S.ndim != 2:
        raise ValueError("not enough values to unpack (expected 2)")  # `n_mels, time_steps = S_dB.shape`
    return torch.from_numpy(np.ascontiguousarray(S, dtype=np.float32)).to(_device())[None]


class GuitarSpecificFilters:
    GUITAR_E2_HZ = midi_to_hz(40)
    GUITAR_E6_HZ = midi_to_hz(88)

    @staticmethod
    def filter_subharmonic_noise(f0, voiced_flag, fmin_hz=82.4):
        f0 = np.asarray(f0, dtype=np.float64)
        if f0.size == 0:
            return f0.copy(), np.asarray(voiced_flag, dtype=bool).copy()
        dev = _device()
        res = core.guitar_filters(None, sr=22050, f0=torch.from_numpy(np.ascontiguousarray(f0)).to(dev)[None],
                                  voiced_flag=torch.from_numpy(np.ascontiguousarray(voiced_flag, dtype=np.uint8)).to(dev)[None],
                                  fmin_hz=fmin_hz)
        return res["f0"][0].cpu().numpy(), res["voiced"][0].cpu().numpy().astype(bool)

    @staticmethod
    def detect_palm_mute(S_dB, hop_length, sr, duration_ms=50):
        Sd = _img(S_dB)
        if Sd.shape[2] == 0:
            return np.zeros(0, dtype=bool)
        res = core.guitar_filters(Sd, sr=sr, hop_length=hop_length, duration_ms=duration_ms, want_rake=False,
                                  want_distortion=False)
        return res["mute_mask"][0].cpu().numpy().astype(bool)

    @staticmethod
    def detect_rake_enhanced(S_dB, hop_length, sr, rake_mask_basic):
        Sd = _img(S_dB)
        if Sd.shape[2] == 0:
            return np.asarray(rake_mask_basic, dtype=bool).copy()
        rk = torch.from_numpy(np.ascontiguousarray(rake_mask_basic, dtype=np.uint8)).to(Sd.device)[None]
        res = core.guitar_filters(Sd, sr=sr, hop_length=hop_length, rake_mask=rk, want_mute=False, want_distortion=False)
        return res["rake_mask"][0].cpu().numpy().astype(bool)

    @staticmethod
    def classify_distortion_level(S_dB):
        res = core.guitar_filters(_img(S_dB), sr=22050, want_rake=False, want_mute=False)
        return core.DISTORTION_LABELS[int(res["distortion"][0])]


def apply_guitar_filters(f0, voiced_flag, S_dB, hop_length, sr, rake_mask):
    """One launch for all four filters; returns the reference's dict (guitar_specific.py:240-277)."""
    Sd = _img(S_dB)
    dev = Sd.device
    res = core.guitar_filters(
        Sd, sr=sr, hop_length=hop_length,
        f0=torch.from_numpy(np.ascontiguousarray(f0, dtype=np.float64)).to(dev)[None],
        voiced_flag=torch.from_numpy(np.ascontiguousarray(voiced_flag, dtype=np.uint8)).to(dev)[None],
        rake_mask=torch.from_numpy(np.ascontiguousarray(rake_mask, dtype=np.uint8)).to(dev)[None])
    return {
        "f0": res["f0"][0].cpu().numpy(),
        "voiced": res["voiced"][0].cpu().numpy().astype(bool),
        "rake_mask": res["rake_mask"][0].cpu().numpy().astype(bool),
        "mute_mask": res["mute_mask"][0].cpu().numpy().astype(bool),
        "distortion": core.DISTORTION_LABELS[int(res["distortion"][0])],
    }
