"""Drop-in for ``aegis_engine_core/worker.py`` (Turbo Mode worker)."""
from __future__ import annotations

from . import librosa_compat as librosa


def _pyin_worker(args):
    """``(chunk, sr, hop_length) -> (f0, voiced_flag, voiced_prob)`` (worker.py:3-15), E2..C6."""
    chunk, sr, hop_length = args
    return librosa.pyin(chunk, fmin=librosa.note_to_hz("E2"), fmax=librosa.note_to_hz("C6"), sr=sr, hop_length=hop_length)
