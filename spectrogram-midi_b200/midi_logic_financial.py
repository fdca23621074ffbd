"""Drop-in for ``aegis_engine_core_v2/midi_logic_financial.py``: the v2 logic filter on the GPU (kernels K5 + K8).

``get_midi_events_financial`` has the reference's signature and returns the same list of dicts
(midi_logic_financial.py:117-388) for ``use_financial=True`` -- the only mode the v2 engine uses
(aegis_engine_financial.py:160-171 passes its own default, True).  The Bollinger / MACD frame labels, the adaptive
threshold, the RSI ghost-note filter and the key / scale / chord pass all run on the device; the reference's progress
prints are not reproduced.  ``use_financial=False`` (the v1-style fallback inside the reference function) is not part
of this path and raises ``NotImplementedError``: use ``midi_logic.get_midi_events`` for the v1 filter.
"""
from __future__ import annotations

import numpy as np
import torch

from . import core
from .librosa_compat import _device


def get_midi_events_financial(rake_mask, f0, voiced_flag, active_probs, rms, sr, hop_length,
                              confidence_threshold=None, **kwargs):
    if not kwargs.get("use_financial", True):
        raise NotImplementedError("use_financial=False is the reference's v1-style fallback; call midi_logic.get_midi_events")
    f0 = np.asarray(f0, dtype=np.float64)
    n = len(f0)
    if n == 0:
        return []
    dev = _device()

    def up(a, dt):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a)[:n], dtype=dt)).to(dev)[None]

    res = core.note_events_financial(
        up(rake_mask, np.uint8), up(f0, np.float64), up(voiced_flag, np.uint8), up(active_probs, np.float64), up(rms, np.float32),
        sr=sr, hop_length=hop_length, confidence_threshold=confidence_threshold,
        noise_gate_db=kwargs.get("noise_gate_db", -40), sustain_ms=kwargs.get("sustain_ms", 50),
        min_note_duration_ms=kwargs.get("min_note_duration_ms", 50),
        use_harmonic_filter=kwargs.get("use_harmonic_filter", True), harmonic_tolerance=kwargs.get("harmonic_tolerance", 1))
    return core.fin_events_to_list(res, 0)
