"""A/B of K3 builds: decode the bench corpus with the library named by AEGIS_B200_LIB (default: the in-tree one), print the
time of K2 / K3 and a hash of states + back-traced f0 so that two builds can be compared.  usage: k3_ab.py [n_clips] [reps]"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import core, tables, batch
n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sr, dur = 22050, 30.0
dev = torch.device("cuda", 0)
plan = P.corpus.plan_events(n_clips, dur, sr)
y = core.synth_events(n_clips, int(dur * sr), plan, dev)
cfg = tables.pyin_config(float(sr), 512, batch.E2, batch.C6)
obs = core.yin_candidates(y, cfg)
ts = []
for it in range(reps + 1):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); dec = core.viterbi_decode(obs, cfg, n_clips); b.record(); torch.cuda.synchronize()
    if it: ts.append(a.elapsed_time(b))
h = hashlib.sha256(dec["states"].cpu().numpy().tobytes() + dec["voiced_flag"].cpu().numpy().tobytes()).hexdigest()[:16]
print(f"lib {os.path.basename(P._native.LIB_PATH)} clips {n_clips}: viterbi {min(ts):.2f} ms (min of {reps}), states sha {h}, voiced {float(dec['voiced_flag'].float().mean()):.4f}")
