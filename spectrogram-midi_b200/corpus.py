"""Seeded synthetic guitar-like corpora (host side).

A seeded restatement of the reference's fixture generator ``generate_test_signal.py:5-97``
(Karplus-Strong plucked string + 25 ms noise rakes, peak-normalised to 0.9) and of the clip built
by ``benchmark_aegis.py:23-47`` (sine C-major scale + noise burst + hiss).  The reference's
generators are unseeded, so no golden bytes exist upstream; these define the corpus the parity
tests and the benchmark run on.
"""
from __future__ import annotations

import numpy as np
import scipy.signal

from .tables import midi_to_hz


def karplus_strong(frequency: float, duration: float, sr: int, rng: np.random.Generator,
                   decay_factor: float = 0.996) -> np.ndarray:
    """Plucked string of ``generate_test_signal.py:5-40``.

    The reference's ring buffer reads ``buf[ptr-1]`` *after* it was overwritten, so the loop filter
    is the one-pole recursion  w[n] = 0.5*decay*(w[n-N] + w[n-1])  (not the textbook two-tap
    average).  Evaluated here one period at a time with ``lfilter`` instead of per sample.
    """
    period = int(sr / frequency)
    n_samples = int(sr * duration)
    noise = rng.uniform(-1, 1, period)
    out = np.empty(max(n_samples, period) + period, dtype=np.float64)
    out[:period] = noise
    c = 0.5 * decay_factor
    prev_block = noise
    carry = noise[-1]
    pos = period
    while pos < n_samples:
        blk, _ = scipy.signal.lfilter([c], [1.0, -c], prev_block, zi=[c * carry])
        out[pos : pos + period] = blk
        carry = blk[-1]
        prev_block = blk
        pos += period
    return out[:n_samples]


def noise_rake(duration: float, sr: int, rng: np.random.Generator) -> np.ndarray:
    """Broadband burst of ``generate_test_signal.py:44-53``."""
    n = int(sr * duration)
    return rng.normal(0, 0.8, n) * np.linspace(1, 0, n) ** 2


def _normalise(track: np.ndarray) -> np.ndarray:
    peak = np.max(np.abs(track))
    if peak > 0:
        track = track / peak * 0.9
    return track.astype(np.float32)


def test_track(sr: int = 44100, seed: int = 0, duration: float | None = None) -> np.ndarray:
    """The E2 / rake / A2 / rake / D3 track of ``generate_test_signal.py:55-97`` (seeded).

    With ``duration`` the note pattern is continued (E2, A2, D3, G3, B3, E4 cycling, a rake before
    every second note) until the clip is that long — BASELINE cfg1 uses 10 s at 22 050 Hz.
    """
    rng = np.random.default_rng(seed)
    silence = np.zeros(int(0.2 * sr))
    parts = [
        silence, karplus_strong(82.41, 1.0, sr, rng), silence, noise_rake(0.025, sr, rng),
        silence[:1000], karplus_strong(110.00, 1.0, sr, rng), silence, noise_rake(0.025, sr, rng),
        karplus_strong(146.83, 1.5, sr, rng),
    ]
    if duration is not None:
        target = int(duration * sr)
        extra = [196.00, 246.94, 329.63, 82.41, 110.00, 146.83]
        i = 0
        while sum(len(p) for p in parts) < target:
            parts.append(silence)
            if i % 2 == 0:
                parts.append(noise_rake(0.025, sr, rng))
            parts.append(karplus_strong(extra[i % len(extra)], 1.0, sr, rng))
            i += 1
        track = np.concatenate(parts)[:target]
    else:
        track = np.concatenate(parts)
    return _normalise(track)


def random_clip(seed: int, duration: float = 30.0, sr: int = 22050) -> np.ndarray:
    """BASELINE cfg2/cfg3 clip: random KS plucks E2..E5 (0.2-1.5 s), 0-4 rakes, 0.2 s silences."""
    rng = np.random.default_rng(seed)
    n_total = int(duration * sr)
    out = np.zeros(n_total, dtype=np.float64)
    n_rakes = int(rng.integers(0, 5))
    rake_slots = set(rng.integers(2, 30, n_rakes).tolist())
    pos = int(0.2 * sr)
    k = 0
    while pos < n_total:
        if k in rake_slots:
            r = noise_rake(0.025, sr, rng)
            seg = r[: n_total - pos]
            out[pos : pos + len(seg)] = seg
            pos += len(r) + int(rng.integers(200, 2000))
        midi = int(rng.integers(40, 77))  # E2 .. E5
        dur = float(rng.uniform(0.2, 1.5))
        amp = float(rng.uniform(0.3, 1.0))
        note = karplus_strong(float(midi_to_hz(midi)), dur, sr, rng) * amp
        seg = note[: max(0, n_total - pos)]
        out[pos : pos + len(seg)] = seg
        pos += len(note)
        if rng.random() < 0.35:
            pos += int(0.2 * sr)
        k += 1
    return _normalise(out)


def benchmark_signal(sr: int = 22050, seed: int = 0) -> np.ndarray:
    """The 4 s clip of ``benchmark_aegis.py:23-47``: C-major sine scale, 50 ms burst at 1.0 s, hiss."""
    rng = np.random.default_rng(seed)
    chunks = []
    for n in [60, 62, 64, 65, 67, 69, 71, 72]:
        t = np.linspace(0, 0.5, int(sr * 0.5))
        chunks.append(0.5 * np.sin(2 * np.pi * float(midi_to_hz(n)) * t))
    y = np.concatenate(chunks)
    s = int(sr * 1.0)
    d = int(sr * 0.05)
    y[s : s + d] += rng.normal(0, 0.8, d)
    y += rng.normal(0, 0.02, len(y))
    return y.astype(np.float32)


def clip_batch(n_clips: int, duration: float, sr: int = 22050, first_seed: int = 0, workers: int | None = None) -> np.ndarray:
    """float32 [n_clips, int(duration*sr)]; clip i uses seed first_seed+i (BASELINE: seed = clip index)."""
    seeds = range(first_seed, first_seed + n_clips)
    if workers is None or workers <= 1 or n_clips < 4:
        clips = [random_clip(s, duration, sr) for s in seeds]
    else:
        import multiprocessing as mp
        from functools import partial

        with mp.get_context("fork").Pool(workers) as pool:
            clips = pool.map(partial(random_clip, duration=duration, sr=sr), seeds)
    return np.stack(clips)


def plan_events(n_clips: int, duration: float, sr: int = 22050, first_seed: int = 0) -> dict:
    """Event list for the device-side synthesiser (``core.synth_events``): the same clip recipe as
    ``random_clip`` (KS plucks E2..E5 of 0.2-1.5 s, 0-4 rakes of 25 ms, 0.2 s silences; seed = clip
    index), planned on the host, rendered by one GPU thread per event."""
    n_total = int(duration * sr)
    clip, start, length, period, amp, seed = [], [], [], [], [], []
    rake_len = int(0.025 * sr)
    for c in range(n_clips):
        rng = np.random.default_rng(first_seed + c)
        n_rakes = int(rng.integers(0, 5))
        rake_slots = set(rng.integers(2, 30, n_rakes).tolist())
        pos = int(0.2 * sr)
        k = 0
        while pos < n_total:
            if k in rake_slots:
                ln = min(rake_len, n_total - pos)
                clip.append(c); start.append(pos); length.append(ln); period.append(0)
                amp.append(1.0); seed.append(int(rng.integers(0, 2**32 - 1)))
                pos += rake_len + int(rng.integers(200, 2000))
                if pos >= n_total:
                    break
            midi = int(rng.integers(40, 77))
            dur = float(rng.uniform(0.2, 1.5))
            n_note = int(sr * dur)
            ln = min(n_note, n_total - pos)
            clip.append(c); start.append(pos); length.append(ln)
            period.append(int(sr / float(midi_to_hz(midi))))
            amp.append(float(rng.uniform(0.3, 1.0))); seed.append(int(rng.integers(0, 2**32 - 1)))
            pos += n_note
            if rng.random() < 0.35:
                pos += int(0.2 * sr)
            k += 1
    return dict(clip=np.asarray(clip, np.int32), start=np.asarray(start, np.int32), length=np.asarray(length, np.int32),
                period=np.asarray(period, np.int32), amp=np.asarray(amp, np.float32), seed=np.asarray(seed, np.uint32))
