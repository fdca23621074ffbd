"""CPU restatement of the v2 ("financial") logic filter.  TEST INFRASTRUCTURE ONLY.

Follows, with ``use_financial=True`` (the engine's default, aegis_engine_financial.py:160-171):

* ``get_midi_events_financial``            aegis_engine_core_v2/midi_logic_financial.py:117-388
* ``adaptive_confidence_threshold``        aegis_engine_core_v2/midi_logic_financial.py:77-114
* ``detect_articulation_bollinger``        aegis_engine_core_v2/financial_analysis.py:148-196
* ``detect_slides_macd``                   aegis_engine_core_v2/financial_analysis.py:228-271
* ``rsi`` / ``filter_ghost_notes_rsi``     aegis_engine_core_v2/financial_analysis.py:277-364
* ``HarmonicAnalyzer.detect_key`` / ``filter_out_of_scale_notes`` / ``analyze_chord_progression`` /
  ``adaptive_filter_by_context``           aegis_engine_core_v2/harmonic_analysis.py:46-283

Pinned: ``tests/golden/make_golden_financial_events.py`` runs the real files (behind the ``librosa`` / ``mido``
shims of ``make_golden.py``) on seeded inputs and ``tests/test_oracle.py`` compares.  The per-frame labels are small
integer codes here (the reference uses strings); the note events are dicts with the reference's keys.
"""
from __future__ import annotations

import numpy as np

from . import librosa_ref as L
from . import reference_files as R

ARTIC = (None, "normal", "bend", "vibrato", "noise")          # code -> label of detect_articulation_bollinger
SLIDE = (None, "normal", "slide_up", "slide_down")            # code -> label of detect_slides_macd
CHROMATIC = ("C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B")
MODES = ("major", "minor", "blues")
SCALES = ((0, 2, 4, 5, 7, 9, 11), (0, 2, 3, 5, 7, 8, 10), (0, 3, 5, 6, 7, 10))


def articulation_codes(f0, window=10, sensitivity=2.0):
    """Band-position labels with the boundary-crossing ("vibrato") counter (financial_analysis.py:158-196)."""
    f0 = np.asarray(f0, dtype=np.float64)
    _, upper, lower = R.bollinger_bands(f0, window, sensitivity)
    side = np.where(f0 > upper, 1, np.where(f0 < lower, -1, 0))   # comparisons with NaN bands are False -> inside
    codes = np.zeros(len(f0), dtype=np.int8)
    prev, crossings = 0, 0
    for i in np.flatnonzero(~np.isnan(f0)):
        crossings = crossings + 1 if (side[i] != prev and prev != 0) else 0
        codes[i] = 3 if crossings >= 2 else (2 if side[i] > 0 else (4 if side[i] < 0 else 1))
        prev = side[i]
    return codes


def slide_codes(f0, threshold=0.3):
    """MACD(5, 20, 9) of the semitone series against +-threshold (financial_analysis.py:244-271)."""
    line, _, hist = R.semitone_macd(f0)
    codes = np.ones(len(line), dtype=np.int8)
    codes[(line > threshold) & (hist > 0)] = 2
    codes[(line < -threshold) & (hist < 0)] = 3
    codes[np.isnan(line)] = 0
    return codes


def adaptive_confidence_threshold(confidence_values):
    """mean - std of the positive confidences, clipped to [0.3, 0.8] (midi_logic_financial.py:91-105)."""
    valid = confidence_values[confidence_values > 0]
    if len(valid) == 0:
        return 0.5
    return float(np.clip(np.mean(valid) - np.std(valid), 0.3, 0.8))


def rsi(data, period=14):
    """Wilder-smoothed relative strength index, 50 before the first full period (financial_analysis.py:294-323)."""
    data = np.asarray(data, dtype=np.float64)
    n = len(data)
    d = np.diff(data)
    gain, loss = np.where(d > 0, d, 0), np.where(d < 0, -d, 0)
    out = np.full(n, 50.0)
    if len(d) < period:
        return out
    g, l = np.mean(gain[:period]), np.mean(loss[:period])
    for i in range(period, n):
        if i > period:
            g = (g * (period - 1) + gain[i - 1]) / period
            l = (l * (period - 1) + loss[i - 1]) / period
        out[i] = 100 if l == 0 else 100 - (100 / (1 + g / l))
    return out


def filter_ghost_notes_rsi(events, rsi_threshold=70):
    """Drop events that start where the RSI of the note-density series is >= threshold
    (financial_analysis.py:337-364; 'start' / 'end' are frame numbers and the bins are tenths of a frame)."""
    if not events:
        return events
    n_bins = int(max(e["end"] for e in events) * 10)
    density = np.zeros(n_bins)
    for e in events:
        a, b = int(e["start"] * 10), int(e["end"] * 10)
        if a < n_bins:
            density[a : min(b, n_bins)] += 1
    r = rsi(density, 14)
    return [e for e in events if not (int(e["start"] * 10) < len(r)) or r[int(e["start"] * 10)] < rsi_threshold]


def detect_key(notes):
    """(root, mode index, score): the scale with the largest share of the pitch-class histogram, first best wins in
    the order root 0..11 x (major, minor, blues) (harmonic_analysis.py:46-123)."""
    if len(notes) == 0:
        return 0, 0, 0.0
    hist = np.zeros(12)
    for n in notes:
        hist[int(n) % 12] += 1.0
    hist = hist / (np.sum(hist) + 1e-6)
    best = (0, 0, 0.0)
    for root in range(12):
        for mode, scale in enumerate(SCALES):
            s = 0.0
            for iv in scale:
                s += hist[(root + iv) % 12]
            if s > best[2]:
                best = (root, mode, s)
    return best


def scale_classes(root, mode):
    return [(root + iv) % 12 for iv in SCALES[mode]]


def out_of_scale(notes, root, mode, tolerance=1):
    """harmonic_analysis.py:145-181: circular pitch-class distance to the nearest scale note > tolerance."""
    sc = scale_classes(root, mode)
    out = np.zeros(len(notes), dtype=bool)
    for i, n in enumerate(notes):
        pc = int(n) % 12
        out[i] = min(min(abs(pc - s), 12 - abs(pc - s)) for s in sc) > tolerance
    return out


def chord_windows(notes, times, window=2000):
    """[(t0, root pc, quality)], quality 1 major / 2 minor / 0 unknown, one per non-empty 2-s window
    (harmonic_analysis.py:183-230).  The root is the most frequent pitch class, first seen wins ties."""
    if len(notes) == 0:
        return []
    out = []
    for t0 in range(0, int(np.max(times)), window):
        pcs = [int(n) % 12 for n, t in zip(notes, times) if t0 <= t < t0 + window]
        if not pcs:
            continue
        counts = {}
        for pc in pcs:
            counts[pc] = counts.get(pc, 0) + 1
        root = max(counts, key=counts.get)                      # dicts keep insertion order: first maximal
        quality = 1 if (root + 4) % 12 in pcs else (2 if (root + 3) % 12 in pcs else 0)
        out.append((t0, root, quality))
    return out


def context_confidences(notes, times, conf, root, mode):
    """harmonic_analysis.py:232-283: x0.8 for a scale note outside the window's triad, x0.5 for a non-scale note."""
    chords = chord_windows(notes, times)
    if not chords:
        return conf
    adj = np.array(conf, dtype=np.float64)
    sc = scale_classes(root, mode)
    for i, (n, t) in enumerate(zip(notes, times)):
        hit = next((c for c in chords if c[0] <= t < c[0] + 2000), None)
        if hit is None or hit[2] == 0:
            continue
        _, r, q = hit
        triad = (r, (r + (4 if q == 1 else 3)) % 12, (r + 7) % 12)
        pc = int(n) % 12
        if pc not in triad:
            adj[i] *= 0.8 if pc in sc else 0.5
    return adj


def frame_series(f0, voiced_flag, active_probs):
    """Phase 1 (midi_logic_financial.py:155-176): trend, articulation / slide codes, combined confidence."""
    f0_clean = np.where(voiced_flag, f0, np.nan)
    trend, _ = R.multi_filter_consensus(f0_clean)
    conf = R.bollinger_confidence(f0_clean, window=10)
    return {
        "trend": np.asarray(trend, dtype=np.float64),
        "artic": articulation_codes(f0_clean, window=10),
        "slide": slide_codes(f0_clean, threshold=0.3),
        "combined": active_probs * 0.5 + conf * 0.5,
    }


def get_midi_events_financial(rake_mask, f0, voiced_flag, active_probs, rms, sr, hop_length, confidence_threshold=None,
                              noise_gate_db=-40, sustain_ms=50, min_note_duration_ms=50, use_harmonic_filter=True,
                              harmonic_tolerance=1):
    f0 = np.asarray(f0, dtype=np.float64)
    voiced_flag = np.asarray(voiced_flag, dtype=bool)
    S = frame_series(f0, voiced_flag, np.asarray(active_probs, dtype=np.float64))
    trend, combined = S["trend"], S["combined"]
    if confidence_threshold is None:
        confidence_threshold = adaptive_confidence_threshold(combined)

    rms_db = L.amplitude_to_db(rms, ref=np.max)
    min_frames = int((min_note_duration_ms / 1000.0) * sr / hop_length)
    sustain_frames = int((sustain_ms / 1000.0) * sr / hop_length)
    n = len(trend)
    with np.errstate(invalid="ignore"):
        active = voiced_flag[:n] & ~np.isnan(trend) & ~(rms_db[:n] < noise_gate_db) & (trend > 0) & ~np.asarray(rake_mask, dtype=bool)[:n]
    note = np.full(n, -1, dtype=np.int64)
    for t in np.flatnonzero(active):
        note[t] = int(round(L.hz_to_midi(trend[t])))

    # Phase 2 (:204-290): maximal runs of one note; the run's label is its last frame label other than None / normal,
    # or the start frame's label when there is none
    events = []
    t = 0
    while t < n:
        if note[t] < 0:
            t += 1
            continue
        s = t
        while t + 1 < n and note[t + 1] == note[s]:
            t += 1
        later = [c for c in S["artic"][s + 1 : t + 1] if c > 1]
        art = ARTIC[later[-1] if later else S["artic"][s]]
        energy = rms_db[s]
        events.append({
            "note": int(note[s]), "start": s, "end": t, "confidence": combined[s],
            "velocity": int(np.clip((energy + 80) * 1.5, 0, 127)),
            "track": "main" if combined[s] >= confidence_threshold else "safe",
            "financial_artic": art, "financial_slide": SLIDE[S["slide"][s]], "technique": art,
        })
        t += 1
    if not events:
        return []

    # Phase 3 (:299-328)
    events = [e for e in events if (e["end"] - e["start"]) >= min_frames]
    if len(events) > 1:
        merged, cur = [], events[0]
        for nxt in events[1:]:
            if nxt["note"] == cur["note"] and (nxt["start"] - cur["end"]) <= sustain_frames and not cur.get("technique"):
                cur["end"] = nxt["end"]
            else:
                merged.append(cur)
                cur = nxt
        merged.append(cur)
        events = merged
    if len(events) > 10:
        events = filter_ghost_notes_rsi(events, 70)

    # Phase 4 (:334-384)
    if use_harmonic_filter and len(events) > 5:
        notes = np.array([e["note"] for e in events])
        root, mode, score = detect_key(notes)
        bad = out_of_scale(notes, root, mode, harmonic_tolerance)
        if bad.sum() > 0:
            for e, b in zip(events, bad):
                e["harmonic_valid"] = not b
            kept = [e for e, b in zip(events, bad) if not b]
            if kept:
                adj = context_confidences(
                    np.array([e["note"] for e in kept]),
                    np.array([e["start"] * (hop_length / sr) * 1000 for e in kept]),
                    np.array([e["confidence"] for e in kept]), root, mode)
                for e, c in zip(kept, adj):
                    e["confidence"] = c
                    e["track"] = "main" if c >= confidence_threshold else "safe"
            events = kept
            events[0]["key_info"] = {"key": CHROMATIC[root], "mode": MODES[mode], "confidence": score}   # IndexError when nothing is kept, as in the reference
    return events


def events_rows(events):
    """Comparison rows: int [note, start, end, velocity, main, artic code, slide code, harmonic_valid (-1 unset)]
    and float [confidence]; plus (root, mode, score) of the first event's key_info or None."""
    ints = np.array([[e["note"], e["start"], e["end"], e["velocity"], int(e["track"] == "main"),
                      ARTIC.index(e.get("technique")), SLIDE.index(e.get("financial_slide")),
                      -1 if "harmonic_valid" not in e else int(e["harmonic_valid"])] for e in events],
                    dtype=np.int64).reshape(-1, 8)
    conf = np.array([e["confidence"] for e in events], dtype=np.float64)
    key = None
    if events and "key_info" in events[0]:
        k = events[0]["key_info"]
        key = (CHROMATIC.index(k["key"]), MODES.index(k["mode"]), float(k["confidence"]))
    return ints, conf, key


def golden_case(g, name):
    """(frame arrays, sr, keyword arguments) of one case of tests/golden/fin_events_golden.npz"""
    k = f"fin/{name}"
    a = g[f"{k}/args"]
    frames = (g[f"{k}/rake_mask"], g[f"{k}/f0"], g[f"{k}/voiced_flag"], g[f"{k}/voiced_prob"], g[f"{k}/rms"])
    kw = dict(confidence_threshold=None if np.isnan(a[1]) else float(a[1]), noise_gate_db=float(a[2]), sustain_ms=float(a[3]),
              min_note_duration_ms=float(a[4]), use_harmonic_filter=bool(a[5]), harmonic_tolerance=int(a[6]))
    return frames, int(a[0]), kw
