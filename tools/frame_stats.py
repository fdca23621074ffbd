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
# ---- sequences the decoder sees (proxy: a frame "collapses" when voiced_prob == 1 with 1..32 candidates; the bound itself holds in ~90 % of those)
T = obs["voiced_prob"].shape[1]
d = ((obs["voiced_prob"] == 1) & (obs["cand_count"].view(n_clips, T) >= 1) & (obs["cand_count"].view(n_clips, T) <= 32))
prev = torch.zeros_like(d); prev[:, 1:] = d[:, :-1]
tot = d.numel()
print(f"run frames (collapse after collapse) {float((d & prev).sum()) / tot:.3f}  first collapse {float((d & ~prev).sum()) / tot:.3f}  "
      f"general after collapse (sparse sources) {float((~d & prev).sum()) / tot:.3f}  general after general (dense) {float((~d & ~prev).sum()) / tot:.3f}")
# lengths of the stretches of non-collapsing frames
dn = (~d).to(torch.int8).cpu()
import numpy as np
lens = []
for c in range(n_clips):
    x = np.concatenate([[0], dn[c].numpy(), [0]])
    e = np.flatnonzero(np.diff(x))
    lens.extend((e[1::2] - e[0::2]).tolist())
lens = np.array(lens)
print(f"stretches of non-collapsing frames: {len(lens)} ({len(lens) / n_clips:.1f} per clip), length 1: {np.mean(lens == 1):.2f}  2: {np.mean(lens == 2):.2f}  3-5: {np.mean((lens >= 3) & (lens <= 5)):.2f}  "
      f"6-20: {np.mean((lens >= 6) & (lens <= 20)):.2f}  > 20: {np.mean(lens > 20):.2f};  frames in stretches > 20: {lens[lens > 20].sum() / max(1, lens.sum()):.2f} of the non-collapsing frames")
vpn = obs["voiced_prob"][~d]
print("voiced_prob of the non-collapsing frames: == 0:", round(float((vpn == 0).float().mean()), 3), " <= 0.02:", round(float((vpn <= 0.02).float().mean()), 3),
      " 0.02..0.99:", round(float(((vpn > 0.02) & (vpn < 0.99)).float().mean()), 3), " >= 0.99:", round(float((vpn >= 0.99).float().mean()), 3))
