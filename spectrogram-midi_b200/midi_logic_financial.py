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
        return _events_without_financial(rake_mask, f0, voiced_flag, active_probs, rms, sr, hop_length, confidence_threshold, **kwargs)
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


def detect_articulations_financial(f0, start, end, analyzer):
    """Articulation of one note from its pitch slice (midi_logic_financial.py:16-74): the label ('bend', 'vibrato', 'noise'
    from the Bollinger bands, window min(5, n), 1.5 sigma; 'slide' when the MACD flags two or more frames) that occurs
    most often, if it covers at least 30 % of the slice; else None."""
    if end <= start:
        return None
    sl = np.asarray(f0[start:end + 1], dtype=np.float64)
    sl = sl[~np.isnan(sl)]
    if len(sl) < 3:
        return None
    labels = analyzer.detect_articulation_bollinger(sl, window=min(5, len(sl)), sensitivity=1.5)
    counts = {}
    for a in labels:
        if a and a != "normal":
            counts[a] = counts.get(a, 0) + 1
    n_slide = sum(1 for x in analyzer.detect_slides_macd(sl, threshold=0.3) if x and x != "normal")
    if n_slide >= 2:
        counts["slide"] = n_slide
    if not counts:
        return None
    name, cnt = max(counts.items(), key=lambda kv: kv[1])     # first of the most frequent, in insertion order
    return name if cnt / len(labels) >= 0.3 else None


def _events_without_financial(rake_mask, f0, voiced_flag, active_probs, rms, sr, hop_length, confidence_threshold=None, **kwargs):
    """``get_midi_events_financial(use_financial=False)`` (midi_logic_financial.py:178-323)."""
    from . import librosa_compat as librosa
    from .financial_analysis import FinancialPitchAnalyzer

    noise_gate_db = kwargs.get("noise_gate_db", -40)
    sustain_ms = kwargs.get("sustain_ms", 50)
    min_note_duration_ms = kwargs.get("min_note_duration_ms", 50)
    f0 = np.asarray(f0, dtype=np.float64)
    voiced_flag = np.asarray(voiced_flag)
    analyzer = FinancialPitchAnalyzer(sr=sr, hop_length=hop_length)
    try:   # librosa.util.softmask has no `margin` keyword: the reference always lands in its except branch (raw f0)
        import scipy.signal

        f0_smooth = librosa.util.softmask(f0, voiced_flag.astype(np.float64), margin=0.5)
        f0_smooth = scipy.signal.medfilt(f0_smooth, kernel_size=3)
    except Exception:
        f0_smooth = f0
    if confidence_threshold is None:
        confidence_threshold = 0.7
    if len(f0_smooth) == 0:
        return []
    rms_db = librosa.amplitude_to_db(np.asarray(rms), ref=np.max)
    min_frames = int((min_note_duration_ms / 1000.0) * sr / hop_length)
    sustain_frames = int((sustain_ms / 1000.0) * sr / hop_length)
    events, cur = [], None

    def close(ev):
        ev["technique"] = detect_articulations_financial(f0_smooth, ev["start"], ev["end"], analyzer)
        events.append(ev)

    for t in range(len(f0_smooth)):
        freq = f0_smooth[t]
        energy = rms_db[t]
        ok = bool(voiced_flag[t]) and not np.isnan(freq) and not (energy < noise_gate_db) and freq > 0 and not rake_mask[t]
        if not ok:
            if cur is not None:
                close(cur)
                cur = None
            continue
        note = int(round(librosa.hz_to_midi(freq)))
        if cur is not None and cur["note"] == note:
            cur["end"] = t
            continue
        if cur is not None:
            close(cur)
        conf = active_probs[t]
        cur = {"note": note, "start": t, "end": t, "confidence": conf, "velocity": int(np.clip((energy + 80) * 1.5, 0, 127)),
               "track": "main" if conf >= confidence_threshold else "safe", "financial_artic": None, "financial_slide": None}
    if cur is not None:
        events.append(cur)            # the reference labels the last open note only in financial mode: no 'technique' key
    events = [e for e in events if (e["end"] - e["start"]) >= min_frames]
    if len(events) > 1:
        merged, c = [], events[0]
        for nxt in events[1:]:
            if nxt["note"] == c["note"] and (nxt["start"] - c["end"]) <= sustain_frames and not c.get("technique"):
                c["end"] = nxt["end"]
            else:
                merged.append(c)
                c = nxt
        merged.append(c)
        events = merged
    return events
