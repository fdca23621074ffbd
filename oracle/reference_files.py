"""CPU restatement of the reference's OWN numpy code on the hot path.  TEST INFRASTRUCTURE ONLY.

Pinned: ``tests/golden/make_golden.py`` runs the real files under ``/root/reference`` on seeded
inputs (in the build container) and ``tests/test_oracle.py`` checks these restatements against the
committed outputs.  Written vectorised / run-length style rather than as per-frame loops; each
function cites the reference lines it follows.
"""
from __future__ import annotations

import numpy as np
import scipy.signal

from . import librosa_ref as L


# --------------------------------------------------------------------------------------------
# aegis_engine_core/vision.py:3-38
# --------------------------------------------------------------------------------------------
def rake_columns(S_dB, broadband_threshold_ratio):
    """Per-column broadband test (vision.py:11-21)."""
    S_dB = np.asarray(S_dB)
    n_mels = S_dB.shape[0]
    col_max = S_dB.max(axis=0)
    active = np.sum(S_dB > (col_max - 20)[None, :], axis=0)
    return (~(col_max < -60)) & ((active / n_mels) > broadband_threshold_ratio)


def rake_frame_limits(hop_length, sr):
    """(min_frames, max_frames) of vision.py:23-25."""
    ms_per_frame = (hop_length / sr) * 1000
    return int(10 / ms_per_frame), int(30 / ms_per_frame)


def run_length_gate(flags, min_frames, max_frames):
    """Keep closed runs of True whose length is within [min, max] (vision.py:27-36).

    A run that is still open at the last element is never emitted.
    """
    flags = np.asarray(flags, dtype=bool)
    out = np.zeros_like(flags)
    padded = np.concatenate([[False], flags, [False]]).astype(np.int8)
    edges = np.diff(padded)
    starts = np.flatnonzero(edges == 1)
    ends = np.flatnonzero(edges == -1)  # exclusive
    for s, e in zip(starts, ends):
        if e >= len(flags):  # still open at the end of the clip
            continue
        if min_frames <= (e - s) <= max_frames:
            out[s:e] = True
    return out


def detect_rake_patterns(S_dB, hop_length, sr, broadband_threshold_ratio):
    lo, hi = rake_frame_limits(hop_length, sr)
    return run_length_gate(rake_columns(S_dB, broadband_threshold_ratio), lo, hi)


# --------------------------------------------------------------------------------------------
# aegis_engine_core_v2/financial_filters.py
# --------------------------------------------------------------------------------------------
def savitzky_golay(data, window=11, polyorder=3):
    """financial_filters.py:25-59 — filter the NaN-compacted series, scatter back."""
    data = np.asarray(data, dtype=np.float64)
    valid = ~np.isnan(data)
    if not valid.any():
        return data
    out = np.full_like(data, np.nan)
    n_valid = int(valid.sum())
    if n_valid > window:
        wl = min(window, n_valid if n_valid % 2 == 1 else n_valid - 1)
        out[valid] = scipy.signal.savgol_filter(data[valid], window_length=wl, polyorder=polyorder, mode="nearest")
    return out


def kalman_filter(data, process_variance=1e-5, measurement_variance=1e-1):
    """financial_filters.py:62-99 — scalar random-walk Kalman, NaNs skipped, state carried."""
    data = np.asarray(data, dtype=np.float64)
    valid = np.flatnonzero(~np.isnan(data))
    if valid.size == 0:
        return data
    out = np.full_like(data, np.nan)
    x = data[valid[0]]
    p = 1.0
    for i in valid:
        p_pred = p + process_variance
        k = p_pred / (p_pred + measurement_variance)
        x = x + k * (data[i] - x)
        p = (1 - k) * p_pred
        out[i] = x
    return out


def holt_winters(data, alpha=0.3, beta=0.1):
    """financial_filters.py:102-141 — Holt level+trend, NaNs skipped, state carried."""
    data = np.asarray(data, dtype=np.float64)
    valid = np.flatnonzero(~np.isnan(data))
    if valid.size == 0:
        return data
    if valid.size < 2:
        return data
    out = np.full_like(data, np.nan)
    level = data[valid[0]]
    trend = data[valid[1]] - data[valid[0]]
    for i in valid:
        forecast = level + trend
        level_new = alpha * data[i] + (1 - alpha) * forecast
        trend = beta * (level_new - level) + (1 - beta) * trend
        level = level_new
        out[i] = level
    return out


def multi_filter_consensus(data, filters=("savgol", "kalman", "holt")):
    """financial_filters.py:256-298 — nanmedian of the chosen filters; conf = 1/(1+nanstd)."""
    import warnings

    data = np.asarray(data, dtype=np.float64)
    rows = []
    if "savgol" in filters:
        rows.append(savitzky_golay(data))
    if "kalman" in filters:
        rows.append(kalman_filter(data))
    if "holt" in filters:
        rows.append(holt_winters(data))
    if not rows:
        return data, np.ones_like(data)
    stacked = np.array(rows)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        consensus = np.nanmedian(stacked, axis=0)
        std = np.nanstd(stacked, axis=0)
    return consensus, 1.0 / (1.0 + std)


def atr_filter(data, window=14, threshold=2.0):
    """financial_filters.py:144-180 (defined by the reference, never called)."""
    import warnings

    data = np.asarray(data, dtype=np.float64)
    if not (~np.isnan(data)).any():
        return data, np.zeros_like(data, dtype=bool)
    tr = np.abs(np.diff(data))
    atr = np.full(len(data), np.nan)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for i in range(window, len(tr)):
            atr[i] = np.nanmean(tr[max(0, i - window) : i])
    noise = np.zeros(len(data), dtype=bool)
    for i in range(1, len(data)):
        if not np.isnan(atr[i]) and not np.isnan(data[i]) and np.abs(data[i] - data[i - 1]) > atr[i] * threshold:
            noise[i] = True
    out = data.copy()
    for i in np.flatnonzero(noise):
        out[i] = out[i - 1] if i > 0 else data[i]
    return out, noise


def _rolling_midrange(data, period):
    out = np.full_like(data, np.nan)
    for i in range(period, len(data)):
        w = data[max(0, i - period) : i]
        w = w[~np.isnan(w)]
        if w.size:
            out[i] = (w.max() + w.min()) / 2
    return out


def ichimoku_baseline(data, tenkan=9, kijun=26):
    """financial_filters.py:183-213 (never called by the reference)."""
    data = np.asarray(data, dtype=np.float64)
    if not (~np.isnan(data)).any():
        return data
    return _rolling_midrange(data, kijun)


def stochastic_oscillator(data, k_period=14, smooth=3):
    """financial_filters.py:216-249 (never called by the reference)."""
    data = np.asarray(data, dtype=np.float64)
    if not (~np.isnan(data)).any():
        return np.full_like(data, 50.0)
    k = np.full_like(data, 50.0)
    for i in range(k_period, len(data)):
        w = data[max(0, i - k_period) : i + 1]
        w = w[~np.isnan(w)]
        if w.size and (w.max() - w.min()) > 0:
            k[i] = ((data[i] - w.min()) / (w.max() - w.min())) * 100
    d = np.full_like(k, 50.0)
    for i in range(smooth, len(k)):
        d[i] = np.mean(k[max(0, i - smooth) : i + 1])
    return d


# --------------------------------------------------------------------------------------------
# aegis_engine_core_v2/financial_analysis.py (numeric series ops)
# --------------------------------------------------------------------------------------------
def simple_moving_average(data, window=5):
    """financial_analysis.py:45-69 — NaN->0, np.convolve(..., 'same'), NaNs restored."""
    data = np.asarray(data, dtype=np.float64)
    nan = np.isnan(data)
    sm = np.convolve(np.where(nan, 0, data), np.ones(window) / window, mode="same")
    sm[nan] = np.nan
    return sm


def exponential_moving_average(data, span=5):
    """financial_analysis.py:71-107 — alpha = 2/(span+1); restart after a NaN gap."""
    data = np.asarray(data, dtype=np.float64)
    alpha = 2 / (span + 1)
    ema = np.full_like(data, np.nan)
    valid = np.flatnonzero(~np.isnan(data))
    if valid.size == 0:
        return ema
    first = valid[0]
    ema[first] = data[first]
    for i in valid[1:]:
        ema[i] = data[i] if np.isnan(ema[i - 1]) else alpha * data[i] + (1 - alpha) * ema[i - 1]
    return ema


def rolling_std(data, window):
    """Population sigma over the trailing window's valid points, needs >= 2 (financial_analysis.py:134-141)."""
    data = np.asarray(data, dtype=np.float64)
    std = np.full_like(data, np.nan)
    for i in range(len(data)):
        w = data[max(0, i - window + 1) : i + 1]
        w = w[~np.isnan(w)]
        if w.size > 1:
            std[i] = np.std(w)
    return std


def bollinger_bands(data, window=20, num_std=2):
    """financial_analysis.py:113-146."""
    ma = simple_moving_average(data, window)
    std = rolling_std(data, window)
    return ma, ma + (num_std * std), ma - (num_std * std)


def macd(data, fast=12, slow=26, signal=9):
    """financial_analysis.py:203-226."""
    line = exponential_moving_average(data, fast) - exponential_moving_average(data, slow)
    sig = exponential_moving_average(line, signal)
    return line, sig, line - sig


def bollinger_confidence(f0, window=10):
    """financial_analysis.py:404-416."""
    f0 = np.asarray(f0, dtype=np.float64)
    _, upper, lower = bollinger_bands(f0, window)
    bw = upper - lower
    conf = np.zeros_like(f0)
    ok = ~np.isnan(f0) & ~np.isnan(bw)
    conf[ok] = np.where(bw[ok] > 0, 1.0 / (1.0 + bw[ok]), 1.0)
    return conf


def semitone_macd(f0):
    """The MACD series inside detect_slides_macd (financial_analysis.py:244-252): 5/20/9 on MIDI numbers."""
    f0 = np.asarray(f0, dtype=np.float64)
    st = np.full_like(f0, np.nan)
    ok = ~np.isnan(f0)
    if ok.any():
        st[ok] = L.hz_to_midi(f0[ok])
    return macd(st, fast=5, slow=20, signal=9)


# --------------------------------------------------------------------------------------------
# aegis_engine_core/midi_logic.py (the unchanged consumer; used to compare note events)
# --------------------------------------------------------------------------------------------
def detect_articulations(f0, start, end):
    """midi_logic.py:6-30."""
    if end <= start:
        return (None, 0.0)
    seg = f0[start : end + 1]
    seg = seg[seg > 0]
    if len(seg) < 3:
        return (None, 0.0)
    notes = L.hz_to_midi(seg)
    x = np.arange(len(notes))
    coef = np.polyfit(x, notes, 1)
    slope = coef[0]
    resid = notes - np.polyval(coef, x)
    if np.max(resid) - np.min(resid) > 0.3:
        return ("vibrato", slope)
    if slope > 0.05:
        return ("bend", slope)
    if abs(slope) > 0.02:
        return ("slide", slope)
    return (None, 0.0)


def get_midi_events(rake_mask, f0, voiced_flag, active_probs, rms, sr, hop_length, confidence_threshold,
                    noise_gate_db=-40, sustain_ms=50, min_note_duration_ms=50):
    """midi_logic.py:32-148 with the softmask TypeError branch taken (f0_smooth = raw f0, :47-49)."""
    f0 = np.asarray(f0)
    rms_db = L.amplitude_to_db(rms, ref=np.max)
    min_frames = int((min_note_duration_ms / 1000.0) * sr / hop_length)
    sustain_frames = int((sustain_ms / 1000.0) * sr / hop_length)

    n = len(f0)
    active = np.asarray(voiced_flag, dtype=bool)[:n] & ~(rms_db[:n] < noise_gate_db) & (f0 > 0) & ~np.asarray(rake_mask, dtype=bool)[:n]
    note = np.full(n, -1, dtype=np.int64)
    for t in np.flatnonzero(active):
        note[t] = int(round(L.hz_to_midi(f0[t])))

    events = []
    t = 0
    while t < n:
        if note[t] < 0:
            t += 1
            continue
        s = t
        while t + 1 < n and note[t + 1] == note[s]:
            t += 1
        energy = rms_db[s]
        conf = active_probs[s]
        ev = {
            "note": int(note[s]), "start": s, "end": t, "confidence": conf,
            "velocity": int(np.clip((energy + 80) * 1.5, 0, 127)),
            "track": "main" if conf >= confidence_threshold else "safe",
            "rms_energy": energy,
        }
        ev["technique"], ev["slope"] = detect_articulations(f0, s, t)
        events.append(ev)
        t += 1

    if not events:
        return []
    events = [e for e in events if (e["end"] - e["start"]) >= min_frames]

    if len(events) > 1:
        merged = []
        cur = events[0]
        for nxt in events[1:]:
            if nxt["note"] == cur["note"] and (nxt["start"] - cur["end"]) <= sustain_frames and not cur.get("technique"):
                cur["end"] = nxt["end"]
            else:
                merged.append(cur)
                cur = nxt
        merged.append(cur)
        events = merged

    for cur, nxt in zip(events[:-1], events[1:]):
        gap_ms = (nxt["start"] - cur["end"]) * (hop_length / sr) * 1000
        if gap_ms < 30:
            dp = nxt["note"] - cur["note"]
            v_ratio = nxt["velocity"] / max(cur["velocity"], 1)
            e_ratio = nxt.get("rms_energy", 0) / max(cur.get("rms_energy", 1), -80)
            weak = v_ratio < 0.7 or e_ratio < 0.8
            if 0 < dp <= 2 and weak:
                nxt["technique"], nxt["slope"] = "hammer_on", 0.0
            elif -2 <= dp < 0 and weak:
                nxt["technique"], nxt["slope"] = "pull_off", 0.0
    return events


def events_key(events):
    """Integer-exact comparison key of a note-event list (note, start, end, velocity, track, technique)."""
    return [(e["note"], int(e["start"]), int(e["end"]), e["velocity"], e["track"], e.get("technique")) for e in events]


# --------------------------------------------------------------------------------------------
# aegis_engine_core_v2/guitar_specific.py:24-277 (caller: aegis_engine_financial.py:132-147)
# Pinned by tests/golden/guitar_golden.npz (tests/golden/make_golden_guitar.py runs the real file).
# --------------------------------------------------------------------------------------------
def filter_subharmonic_noise(f0, voiced_flag, fmin_hz=82.4):
    """guitar_specific.py:24-61: below fmin -> NaN / unvoiced, unless the octave above lands in [fmin, 4 fmin)."""
    f0 = np.asarray(f0, dtype=np.float64)
    out_f0 = f0.copy()
    out_v = np.asarray(voiced_flag, dtype=bool).copy()
    with np.errstate(invalid="ignore"):
        sub = f0 < fmin_hz                      # NaN compares False: unvoiced frames are left alone
        corrected = f0 * 2
        fix = sub & (fmin_hz <= corrected) & (corrected < fmin_hz * 4)
    out_f0[sub] = np.nan
    out_v[sub] = False
    out_f0[fix] = corrected[fix]
    out_v[fix] = True
    return out_f0, out_v


def palm_mute_columns(S_dB):
    """guitar_specific.py:84-94: mean dB of the low half > 2 x mean dB of the high half (float32 arithmetic)."""
    S_dB = np.asarray(S_dB)
    mid = S_dB.shape[0] // 2
    low = np.mean(S_dB[:mid, :], axis=0)
    high = np.mean(S_dB[mid:, :], axis=0)
    return (low / (high + 1e-6)) > 2.0


def detect_palm_mute(S_dB, hop_length, sr, duration_ms=50):
    """guitar_specific.py:63-110: closed runs of mute columns of at most int(duration_ms / ms_per_frame) frames."""
    ms_per_frame = (hop_length / sr) * 1000
    return run_length_gate(palm_mute_columns(S_dB), 0, int(duration_ms / ms_per_frame))


def detect_rake_enhanced(S_dB, hop_length, sr, rake_mask_basic):
    """guitar_specific.py:112-151: a > 10 dB jump of the mean dB whose next `threshold_frames` differences average
    below zero marks those frames."""
    S_dB = np.asarray(S_dB)
    out = np.asarray(rake_mask_basic, dtype=bool).copy()
    total = np.mean(S_dB, axis=0)
    diff = np.diff(total, prepend=total[0]) if total.size else total
    n = int(30 / ((hop_length / sr) * 1000))
    if n <= 0:
        return out      # np.mean of an empty slice is NaN: the reference never fires
    for i in np.flatnonzero(diff > 10):
        if i >= 1 and i + n < len(diff) and np.mean(diff[i : i + n]) < 0:
            out[i : i + n] = True
    return out


def classify_distortion_level(S_dB):
    """guitar_specific.py:209-233."""
    S_dB = np.asarray(S_dB)
    high = np.mean(S_dB[int(S_dB.shape[0] * 0.7):, :])
    ratio = high / (np.mean(S_dB) + 1e-6)
    return "heavy" if ratio > 0.4 else ("light" if ratio > 0.25 else "clean")


def apply_guitar_filters(f0, voiced_flag, S_dB, hop_length, sr, rake_mask):
    """guitar_specific.py:240-277."""
    f0f, vf = filter_subharmonic_noise(f0, voiced_flag, fmin_hz=82.4)
    return {"f0": f0f, "voiced": vf, "rake_mask": detect_rake_enhanced(S_dB, hop_length, sr, rake_mask),
            "mute_mask": detect_palm_mute(S_dB, hop_length, sr), "distortion": classify_distortion_level(S_dB)}
