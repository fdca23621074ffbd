"""Distribution of pYIN frame types on the bench corpus: what the Viterbi decoder's three regimes see."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import core, tables, batch
n_clips, sr, dur = 256, 22050, 30.0
dev = torch.device("cuda", 0)
y = core.synth_events(n_clips, int(dur * sr), P.corpus.plan_events(n_clips, dur, sr), dev)
cfg = tables.pyin_config(float(sr), 512, batch.E2, batch.C6)
obs = core.yin_candidates(y, cfg)
vp = obs["voiced_prob"].flatten()
cc = obs["cand_count"].flatten()
n = vp.numel()
print(f"frames {n}: voiced_prob == 1: {float((vp == 1).float().mean()):.3f}  == 0: {float((vp == 0).float().mean()):.3f}  "
      f"in (0, 1): {float(((vp > 0) & (vp < 1)).float().mean()):.3f}   >= 0.999999: {float((vp >= 0.999999).float().mean()):.3f}  >= 0.99: {float((vp >= 0.99).float().mean()):.3f}")
print("candidates per frame histogram:", torch.bincount(cc.clamp(max=12)).tolist())
mid = (vp > 0) & (vp < 1)
print("voiced_prob quantiles of the in-between frames:", [round(float(q), 4) for q in torch.quantile(vp[mid][:4_000_000], torch.tensor([0.05, 0.25, 0.5, 0.75, 0.95], device=dev, dtype=vp.dtype))])
