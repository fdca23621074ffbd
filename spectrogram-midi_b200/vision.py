"""Drop-in for ``aegis_engine_core/vision.py``: spectrogram rake-noise mask on the GPU."""
from __future__ import annotations

import numpy as np
import torch

from . import core
from .librosa_compat import _device


def detect_rake_patterns(S_dB, hop_length, sr, broadband_threshold_ratio):
    """Same contract as ``detect_rake_patterns`` (aegis_engine_core/vision.py:3-38).

    ``S_dB`` float [n_mels, T] dB image -> bool [T]: a column is broadband when its maximum is
    >= -60 dB and more than ``ratio`` of its bins lie within 20 dB of that maximum; only closed runs
    of ``int(10/ms_per_frame) .. int(30/ms_per_frame)`` such columns are kept.
    """
    S = np.asarray(S_dB)
    if S.ndim != 2:
        raise ValueError("not enough values to unpack (expected 2)")  # what `n_mels, time_steps = S_dB.shape` raises
    n_mels, T = S.shape
    if T == 0:
        return np.zeros(0, dtype=bool)
    Sd = torch.from_numpy(np.ascontiguousarray(S, dtype=np.float32)).to(_device())[None]
    out = core.mel_post(Sd, None, sr=sr, hop_length=hop_length, rake_ratio=broadband_threshold_ratio,
                        want_sdb=False, want_rake=True, input_is_db=True)
    return out["rake_mask"][0].cpu().numpy().astype(bool)
