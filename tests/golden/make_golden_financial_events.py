"""Generate tests/golden/fin_events_golden.npz from the REAL reference files of the v2 logic filter:
``aegis_engine_core_v2/midi_logic_financial.py`` (with ``financial_analysis.py``, ``financial_filters.py`` and
``harmonic_analysis.py`` next to it; SURVEY.md §8f rank 2, caller aegis_engine_financial.py:160-171).

Run in the build container only (``/root/reference`` must exist):

    python tests/golden/make_golden_financial_events.py

The modules are imported by path, unmodified, as a stub package behind the ``librosa`` / ``mido`` shims of
``make_golden.py``.  Inputs are seeded (oracle perception outputs of corpus clips, and synthetic frame series that
exercise band crossings, slides, the RSI ghost-note filter and the out-of-scale filter at tolerance 0); inputs and
the reference's outputs are stored side by side.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import make_golden as MG  # noqa: E402
from oracle import librosa_ref as L  # noqa: E402
from oracle import financial_events as FE  # noqa: E402  (only for the label <-> code tables and events_rows)
from spectrogram_midi_b200 import corpus  # noqa: E402


def synthetic_frames(seed, n, sr, quarter_tones=False):
    """Frame series of a made-up performance: notes of 5..60 frames with jitter, some with vibrato, bends, slides and
    outliers; unvoiced gaps; a few rake frames; an energy envelope that dips under the noise gate now and then."""
    r = np.random.default_rng(seed)
    f0 = np.full(n, np.nan)
    voiced = np.zeros(n, bool)
    rms = np.full(n, 1e-4, dtype=np.float32)
    t = int(r.integers(0, 6))
    scale = np.array([40, 43, 45, 46, 47, 50, 52, 55, 57, 58, 59, 62, 64, 66, 69])   # mostly E blues, some strangers
    while t < n:
        dur = int(r.integers(3, 60))
        m = float(r.choice(scale)) + (0.37 if not quarter_tones else 0.0) * float(r.random() < 0.15)
        k = np.arange(dur)
        cents = r.normal(0, 4.0, dur)
        kind = r.integers(0, 6)
        if kind == 1:
            cents += 45.0 * np.sin(2 * np.pi * k / r.uniform(5, 9))
        elif kind == 2:
            cents += np.minimum(k * r.uniform(4, 14), 190.0)
        elif kind == 3:
            cents += k * r.uniform(-25, 25)
        elif kind == 4:
            cents[r.integers(0, dur, size=max(1, dur // 8))] += r.choice([-1, 1]) * r.uniform(60, 300)
        e = min(n, t + dur)
        f0[t:e] = 440.0 * 2.0 ** ((m - 69.0 + cents[: e - t] / 100.0) / 12.0)
        voiced[t:e] = True
        level = r.uniform(0.004, 0.4)
        rms[t:e] = (level * np.exp(-k[: e - t] / r.uniform(15, 80))).astype(np.float32)
        t = e + int(r.choice([0, 0, 1, 2, 3, 6, 12]))
    drop = r.random(n) < 0.03
    voiced &= ~drop
    probs = np.where(voiced, r.uniform(0.25, 1.0, n), r.uniform(0.0, 0.3, n))
    f0_raw = np.where(voiced, f0, np.nan)
    rake = r.random(n) < 0.02
    return rake, f0_raw, voiced, probs, rms


def main():
    MG._install_shims()
    pkg = types.ModuleType("refv2")
    pkg.__path__ = [f"{MG.REF}/aegis_engine_core_v2"]
    sys.modules["refv2"] = pkg
    MG._load("refv2.financial_filters", f"{MG.REF}/aegis_engine_core_v2/financial_filters.py", "refv2")
    MG._load("refv2.financial_analysis", f"{MG.REF}/aegis_engine_core_v2/financial_analysis.py", "refv2")
    MG._load("refv2.harmonic_analysis", f"{MG.REF}/aegis_engine_core_v2/harmonic_analysis.py", "refv2")
    logic = MG._load("refv2.midi_logic_financial", f"{MG.REF}/aegis_engine_core_v2/midi_logic_financial.py", "refv2")
    vision = MG._load("ref_vision", f"{MG.REF}/aegis_engine_core/vision.py")

    cases = []   # (name, frames tuple, sr, kwargs)
    for name, y, sr in [("track22050", corpus.test_track(22050, 0, 10.0), 22050),
                        ("clip11", corpus.random_clip(11, 20.0, 22050), 22050)]:
        S_dB = L.load_audio_features(y, sr)
        mask = vision.detect_rake_patterns(S_dB, 512, sr, 0.6)
        f0, vf, vp = L.pyin(y, fmin=L.note_to_hz("E2"), fmax=L.note_to_hz("C6"), sr=sr, hop_length=512)
        rms = L.rms(y, hop_length=512)[0]
        cases.append((name, (mask, f0, vf, vp, rms), sr, {}))
        cases.append((name + "_tol0", (mask, f0, vf, vp, rms), sr, {"harmonic_tolerance": 0}))
    for seed, n, sr in [(1, 431, 22050), (2, 1292, 22050), (3, 1292, 44100), (4, 700, 22050)]:
        fr = synthetic_frames(seed, n, sr)
        cases.append((f"synth{seed}", fr, sr, {}))
        cases.append((f"synth{seed}_tol0", fr, sr, {"harmonic_tolerance": 0}))
    fr = synthetic_frames(5, 900, 22050)
    cases.append(("synth5_fixed_thr", fr, 22050, {"confidence_threshold": 0.62}))
    cases.append(("synth5_no_harmonic", fr, 22050, {"use_harmonic_filter": False, "min_note_duration_ms": 0, "sustain_ms": 120}))
    cases.append(("synth5_gate", fr, 22050, {"noise_gate_db": -25, "harmonic_tolerance": 0}))
    fr = synthetic_frames(6, 60, 22050)
    cases.append(("synth6_short", fr, 22050, {"harmonic_tolerance": 0}))
    silent = (np.zeros(50, bool), np.full(50, np.nan), np.zeros(50, bool), np.zeros(50), np.full(50, 1e-3, np.float32))
    cases.append(("silent", silent, 22050, {}))

    out = {}
    names = []
    for name, (rake, f0, vf, vp, rms), sr, kw in cases:
        kw = dict(kw)
        thr = kw.pop("confidence_threshold", None)
        with contextlib.redirect_stdout(io.StringIO()):
            ev = logic.get_midi_events_financial(rake_mask=rake.copy(), f0=f0.copy(), voiced_flag=vf.copy(), active_probs=vp.copy(),
                                                 rms=rms.copy(), sr=sr, hop_length=512, confidence_threshold=thr,
                                                 use_financial=True, **kw)
        ints, conf, key = FE.events_rows(ev)
        k = f"fin/{name}"
        names.append(name)
        out[f"{k}/rake_mask"], out[f"{k}/f0"], out[f"{k}/voiced_flag"] = rake, f0, vf
        out[f"{k}/voiced_prob"], out[f"{k}/rms"] = vp, rms
        out[f"{k}/args"] = np.array([sr, np.nan if thr is None else thr, kw.get("noise_gate_db", -40), kw.get("sustain_ms", 50),
                                     kw.get("min_note_duration_ms", 50), float(kw.get("use_harmonic_filter", True)),
                                     kw.get("harmonic_tolerance", 1)], dtype=np.float64)
        out[f"{k}/events"], out[f"{k}/confidence"] = ints, conf
        out[f"{k}/key"] = np.array([-1.0, -1.0, 0.0] if key is None else key, dtype=np.float64)
        print(f"{name:22s} events {len(ev):3d}  key {key}  techniques {sorted(set(ints[:, 5].tolist()))}")
    out["fin/names"] = np.array(names)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fin_events_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
