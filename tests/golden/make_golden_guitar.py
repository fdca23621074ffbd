"""Generate tests/golden/guitar_golden.npz from the REAL reference file
``aegis_engine_core_v2/guitar_specific.py`` (caller aegis_engine_financial.py:132-147).

Run in the build container only (``/root/reference`` must exist):

    python tests/golden/make_golden_guitar.py

The reference module is imported by path, unmodified, behind the same three-function ``librosa`` shim
as ``make_golden.py`` (it only needs ``midi_to_hz`` at import time).  Inputs are seeded; inputs and the
reference's outputs are stored side by side.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import make_golden as MG  # noqa: E402
from oracle import librosa_ref as L  # noqa: E402
from spectrogram_midi_b200 import corpus  # noqa: E402

DISTORTION = {"clean": 0, "light": 1, "heavy": 2}


def synthetic_db(rng, T, sr):
    """dB image with stretches that fire the reference's palm-mute test (mean dB of the low half > 2 x mean dB of the
    high half: with negative dB values that is a much QUIETER low half) in runs of several lengths, one still open at
    the end, and broadband bursts that rise by > 10 dB and fall again (enhanced-rake triggers)."""
    img = rng.uniform(-62.0, -48.0, size=(128, T)).astype(np.float32)
    img[64:] -= 6.0
    for s, n in [(7, 1), (15, 2), (25, 3), (40, 4), (55, 5), (70, 9), (T - 3, 3)]:
        img[:64, s : s + n] = rng.uniform(-75.0, -65.0, size=(64, min(n, T - s))).astype(np.float32)
        img[64:, s : s + n] = rng.uniform(-32.0, -28.0, size=(64, min(n, T - s))).astype(np.float32)
    for s, n in [(90, 1), (100, 2), (110, 3), (T - 2, 1)]:
        img[:, s : s + n] = rng.uniform(-14.0, -2.0, size=(128, min(n, T - s))).astype(np.float32)
    return img


def main():
    MG._install_shims()
    guitar = MG._load("ref_guitar_specific", f"{MG.REF}/aegis_engine_core_v2/guitar_specific.py")
    rng = np.random.default_rng(4321)
    out = {}
    cases = []
    for sr in (22050, 44100):
        y = corpus.test_track(sr, seed=0)
        S_dB = L.load_audio_features(y, sr)
        f0, vf, _ = L.pyin(y, fmin=L.note_to_hz("E2"), fmax=L.note_to_hz("C6"), sr=sr, hop_length=512)
        cases.append((f"track{sr}", S_dB, f0, vf, 512, sr))
    for sr, hop in ((22050, 512), (44100, 512), (44100, 128)):
        T = 140
        S_dB = synthetic_db(rng, T, sr)
        f0 = rng.uniform(30.0, 400.0, T)
        f0[rng.random(T) < 0.2] = np.nan
        f0[:6] = [41.2, 41.203, 82.4, 82.39999, 164.79, 20.0]   # around the E2 gate and its octave window
        vf = ~np.isnan(f0) & (rng.random(T) < 0.9)
        cases.append((f"synthetic{sr}_{hop}", S_dB, f0, vf, hop, sr))
    # a hot, bright image: distortion 'heavy' / 'light' branches
    hot = rng.uniform(-30.0, -5.0, size=(128, 50)).astype(np.float32)
    hot[90:] *= 0.2
    cases.append(("bright", hot, np.full(50, 220.0), np.ones(50, bool), 512, 22050))
    mid = rng.uniform(-40.0, -20.0, size=(128, 50)).astype(np.float32)
    mid[90:] *= 0.33
    cases.append(("lightdist", mid, np.full(50, 110.0), np.ones(50, bool), 512, 22050))
    cln = rng.uniform(-40.0, -30.0, size=(128, 50)).astype(np.float32)
    cln[90:] = rng.uniform(-6.0, -2.0, size=(38, 50)).astype(np.float32)
    cases.append(("cleandist", cln, np.full(50, 110.0), np.ones(50, bool), 512, 22050))
    for name, S_dB, f0, vf, hop, sr in cases:
        import oracle.reference_files as R
        rake = R.detect_rake_patterns(S_dB, hop, sr, 0.6)
        res = guitar.apply_guitar_filters(f0.copy(), vf.copy(), S_dB.copy(), hop, sr, rake.copy())
        k = f"guitar/{name}"
        out[f"{k}/S_dB"] = S_dB.astype(np.float32)
        out[f"{k}/f0"] = np.asarray(f0, np.float64)
        out[f"{k}/voiced"] = np.asarray(vf, bool)
        out[f"{k}/rake_in"] = np.asarray(rake, bool)
        out[f"{k}/args"] = np.array([hop, sr], dtype=np.int64)
        out[f"{k}/out_f0"] = np.asarray(res["f0"], np.float64)
        out[f"{k}/out_voiced"] = np.asarray(res["voiced"], bool)
        out[f"{k}/out_rake"] = np.asarray(res["rake_mask"], bool)
        out[f"{k}/out_mute"] = np.asarray(res["mute_mask"], bool)
        out[f"{k}/out_distortion"] = np.array([DISTORTION[res["distortion"]]], dtype=np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "guitar_golden.npz")
    np.savez_compressed(path, **out)
    summary = {n: (int(out[f"guitar/{n}/out_mute"].sum()), int((out[f"guitar/{n}/out_rake"] ^ out[f"guitar/{n}/rake_in"]).sum()),
                   int(out[f"guitar/{n}/out_distortion"][0])) for n, *_ in cases}
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB; (mute frames, rake frames added, distortion): {summary}")


if __name__ == "__main__":
    main()
