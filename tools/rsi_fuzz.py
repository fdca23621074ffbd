"""One-off fuzz of K8's closed-form RSI verdicts against its own stepwise walk (AEGIS_FIN_EXACT_RSI=1) on many random clips
and thresholds: the event records must be byte-identical.  usage: rsi_fuzz.py [clips] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import spectrogram_midi_b200 as P
from test_gpu_parity import _fin_frames
n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
rng = np.random.default_rng(seed)
bad = 0
for T in (700, 1292, 2600):
    clips = []
    for i in range(n_clips):
        rake, f0, vf, vp, rms = _fin_frames(seed * 100000 + T * 1000 + i, T, steady_grid=bool(i % 3 == 1))
        if i % 5 == 0:   # silences of random length
            lo = int(rng.integers(0, T // 2)); hi = min(T, lo + int(rng.integers(50, T)))
            f0[lo:hi], vf[lo:hi], vp[lo:hi] = np.nan, False, 0.05
        clips.append((rake, f0, vf, vp, rms))
    stack = [torch.from_numpy(np.stack([c[j] for c in clips])).to(dev) for j in range(5)]
    for thr in (70, 62.5, 50.0, 30, 99.9999, 100.0):
        os.environ.pop("AEGIS_FIN_EXACT_RSI", None)
        fast = P.core.note_events_financial(*stack, sr=22050, hop_length=512, rsi_threshold=thr)
        os.environ["AEGIS_FIN_EXACT_RSI"] = "1"
        walk = P.core.note_events_financial(*stack, sr=22050, hop_length=512, rsi_threshold=thr)
        os.environ.pop("AEGIS_FIN_EXACT_RSI")
        same = torch.equal(fast["n_events"], walk["n_events"]) and torch.equal(fast["events"], walk["events"])
        bad += not same
        print(f"T {T} thr {thr}: events {int(fast['n_events'].sum())} {'ok' if same else 'MISMATCH'}")
print("FUZZ", "PASSED" if bad == 0 else f"FAILED ({bad})")
sys.exit(1 if bad else 0)
