"""Time K2 (YIN candidates) and K3 (Viterbi forward / back-trace) separately on cfg3-style clips."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import core, tables, batch
n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sr, dur = 22050, 30.0
dev = torch.device("cuda", 0)
plan = P.corpus.plan_events(n_clips, dur, sr)
y = core.synth_events(n_clips, int(dur * sr), plan, dev)
cfg = tables.pyin_config(float(sr), 512, batch.E2, batch.C6)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(reps + 1):
    a = ev(); obs = core.yin_candidates(y, cfg); b = ev(); dec = core.viterbi_decode(obs, cfg, n_clips); c = ev()
    torch.cuda.synchronize()
    if it:
        print(f"clips {n_clips}: yin {a.elapsed_time(b):.2f} ms  viterbi {b.elapsed_time(c):.2f} ms  total {a.elapsed_time(c):.2f} ms  "
              f"-> {n_clips * dur / (a.elapsed_time(c) / 1e3):.0f} audio-s/s; mean candidates/frame {float(obs['cand_count'].float().mean()):.2f} "
              f"voiced {float(dec['voiced_flag'].float().mean()):.3f}")
