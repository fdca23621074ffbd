import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import corpus, tables
from oracle import librosa_ref as L
dev = torch.device("cuda", 0)
E2, C6 = L.note_to_hz("E2"), L.note_to_hz("C6")
seed, dur = int(sys.argv[1]), float(sys.argv[2])
y = corpus.random_clip(seed, dur, 22050)
cfg = tables.pyin_config(22050.0, 512, E2, C6)
n = cfg.n_pitch_bins
r32 = L.pyin(y, fmin=E2, fmax=C6, sr=22050, hop_length=512, return_intermediates=True)
obs = P.core.yin_candidates(torch.from_numpy(y).to(dev)[None], cfg)
T = obs["n_frames"]
cb = obs["cand_bin"].cpu().numpy().astype(np.int64); cp = obs["cand_prob"].cpu().numpy(); cc = obs["cand_count"].cpu().numpy()
dense = np.zeros((n, T))
for t in range(T): dense[cb[t, :cc[t]], t] = cp[t, :cc[t]]
ref = r32[3]["observation_probs"][:n]
diff = np.abs(dense - ref).max(axis=0)
bad = np.flatnonzero(diff > 2e-3)
print("frames", T, "bad", len(bad), bad[:40])
fr = L.frame_signal(y)
yin = r32[3]["yin_frames"]
for t in bad[:8]:
    e = float((fr[:, t].astype(np.float64) ** 2).sum())
    print(f"frame {t}: energy {e:.3e} maxabs {np.abs(fr[:, t]).max():.3e} nonzero samples {(fr[:, t] != 0).sum()}")
    print("   gpu :", [(int(b), round(float(p), 4)) for b, p in zip(cb[t, :cc[t]], cp[t, :cc[t]])][:8])
    nz = np.flatnonzero(ref[:, t])
    print("   ref :", [(int(b), round(float(ref[b, t]), 4)) for b in nz][:8])
    tr = np.flatnonzero(L.localmin_rows(yin[:, t:t+1])[:, 0])
    print("   ref troughs (lag idx, height):", [(int(k), round(float(yin[k, t]), 4)) for k in tr if yin[k, t] < 1.0][:8])
dec = P.core.viterbi_decode(obs, cfg, 1)
st = dec["states"][0].cpu().numpy().astype(np.uint16)
print("state diffs at", np.flatnonzero(st != r32[3]["states"])[:40])
o2 = P.core.yin_candidates(torch.from_numpy(y).to(dev)[None], cfg, want_cmnd=True)
g = o2["cmnd"][0].cpu().numpy().T   # [lags, T]
print("cmnd shape", g.shape, yin.shape, "max abs diff overall", np.abs(g - yin).max())
err = np.abs(g - yin).max(axis=0)
print("frames with cmnd err > 1e-4:", np.flatnonzero(err > 1e-4)[:30])
for t in bad[:3]:
    k = np.argmax(np.abs(g[:, t] - yin[:, t]))
    print(f"frame {t}: max cmnd err {err[t]:.3e} at lag idx {k}; gpu {g[k-1:k+2, t]} ref {yin[k-1:k+2, t]}")
    print("    first 6 gpu", g[:6, t], "ref", yin[:6, t])
    tr = np.flatnonzero(L.localmin_rows(yin[:, t:t+1])[:, 0]); trg = np.flatnonzero(L.localmin_rows(g[:, t:t+1])[:, 0])
    print("    ref argmin trough", tr[np.argmin(yin[tr, t])], yin[tr, t].min(), " gpu argmin trough", trg[np.argmin(g[trg, t])], g[trg, t].min())
