"""Drop-in for ``aegis_engine_core/midi_logic.py``: the frame -> note-event logic filter on the GPU (kernel K7).

``get_midi_events`` has the reference's signature and returns the same list of dicts (midi_logic.py:32-148);
``detect_articulations`` (:6-30) is folded into the kernel.  The f0 smoothing step of the reference always
falls back to the raw f0 (its ``librosa.util.softmask(..., margin=0.5)`` call raises TypeError, :43-49), which is
what this path implements.
"""
from __future__ import annotations

import numpy as np
import torch

from . import core, tables
from .librosa_compat import _device


def _note_lut(f0: np.ndarray):
    """MIDI note of every distinct positive f0 value with the reference's own expression
    ``int(round(librosa.hz_to_midi(freq)))`` (midi_logic.py:69), and every frame's index into that table."""
    f = np.asarray(f0, dtype=np.float64)
    pos = f > 0
    vals, inv = np.unique(f[pos], return_inverse=True)
    if len(vals) == 0 or len(vals) >= 65535:
        return None, None
    lut = np.array([int(round(float(tables.hz_to_midi(v)))) for v in vals], dtype=np.int16)
    idx = np.full(f.shape, 65535, dtype=np.uint16)
    idx[pos] = inv.astype(np.uint16)
    return idx, lut


def get_midi_events(rake_mask, f0, voiced_flag, active_probs, rms, sr, hop_length, confidence_threshold, **kwargs):
    f0 = np.asarray(f0, dtype=np.float64)
    n = len(f0)
    if n == 0:
        return []
    dev = _device()
    idx, lut = _note_lut(f0)

    def up(a, dt):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a)[:n], dtype=dt)).to(dev)[None]

    res = core.note_events(
        up(rake_mask, np.uint8), up(f0, np.float64), up(voiced_flag, np.uint8), up(active_probs, np.float64), up(rms, np.float32),
        sr=sr, hop_length=hop_length, confidence_threshold=confidence_threshold,
        noise_gate_db=kwargs.get("noise_gate_db", -40), sustain_ms=kwargs.get("sustain_ms", 50),
        min_note_duration_ms=kwargs.get("min_note_duration_ms", 50),
        pitch_index=None if idx is None else torch.from_numpy(idx.view(np.int16)).to(dev)[None],
        note_lut=None if lut is None else torch.from_numpy(lut).to(dev))
    return core.note_events_to_list(res["events"], res["n_events"], 0)
