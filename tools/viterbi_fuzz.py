"""One-off fuzz of K3 (forward + cone back-trace) against the dense float64 decoder of the oracle: random sparse observations
with exact voiced_prob == 1 frames (collapse rule), far jumps (out-of-band transitions, out-of-cone back-pointers), empty
frames, edge bins and runs of slowly drifting candidates (the regime of real notes).  usage: viterbi_fuzz.py [trials] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import tables
from oracle import librosa_ref as L
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
E2, C6 = 82.4068892282175, 1046.5022612023945
rng = np.random.default_rng(seed)


def sparse(obs, n, max_cand):
    T = obs.shape[1]
    cb = np.zeros((T, max_cand), np.int16); cp = np.zeros((T, max_cand)); cc = np.zeros(T, np.int32)
    for t in range(T):
        nz = np.flatnonzero(obs[:n, t])
        cc[t] = len(nz); cb[t, : len(nz)] = nz; cp[t, : len(nz)] = obs[nz, t]
    vp = np.clip(obs[:n].sum(axis=0), 0, 1)
    return dict(cand_bin=torch.from_numpy(cb).to(dev), cand_prob=torch.from_numpy(cp).to(dev), cand_count=torch.from_numpy(cc).to(dev),
                voiced_prob=torch.from_numpy(vp[None]).to(dev), n_frames=T, max_cand=max_cand)


bad = 0
for trial in range(trials):
    sr = 22050 if trial % 3 else 44100
    cfg = tables.pyin_config(float(sr), 512, E2, C6)
    n = cfg.n_pitch_bins
    trans, _ = L.pyin_transition(n, 10, sr, 512)
    p_init = np.zeros(2 * n); p_init[n:] = 1 / n
    T = int(rng.integers(60, 260))
    obs = np.zeros((2 * n, T))
    centre = int(rng.integers(0, n))
    for t in range(T):
        mode = rng.random()
        if mode < 0.12:
            bins = np.array([], int)                                     # empty frame
        elif mode < 0.7:                                                 # a note: candidates near a drifting centre (+ octave)
            centre = int(np.clip(centre + rng.integers(-6, 7), 0, n - 1))
            extra = [b for b in (centre - 120, centre + 120, centre + int(rng.integers(-40, 41))) if 0 <= b < n and rng.random() < 0.4]
            bins = np.unique(np.array([centre] + extra))
        else:                                                            # jump: anything anywhere
            centre = int(rng.integers(0, n))
            bins = np.unique(np.concatenate([[centre], rng.choice(n, size=int(rng.integers(0, 5)), replace=False)]))
        pr = rng.random(len(bins)) + 1e-3
        if len(bins) and rng.random() < 0.55:
            pr = pr / pr.sum()
            pr[0] += 1.0 - pr.sum()                                      # push the sum to 1 (clipped below)
        elif len(bins):
            pr = pr * rng.random() / len(bins)
        obs[bins, t] = np.maximum(pr, 1e-12)
    vp = np.clip(obs[:n].sum(axis=0, keepdims=True), 0, 1)
    obs[n:, :] = (1 - vp) / n
    ref = L.viterbi(obs, trans, p_init)
    dec = P.core.viterbi_decode(sparse(obs, n, cfg.max_troughs), cfg, 1)
    got = dec["states"][0].cpu().numpy().astype(np.uint16)
    jumps = int((np.abs(np.diff(ref.astype(int) % n)) > cfg.n_pitch_bins // 9).sum())
    ok = np.array_equal(got, ref)
    bad += not ok
    print(f"trial {trial}: sr {sr} T {T} vp==1 frames {int((vp == 1).sum())} path jumps {jumps} {'ok' if ok else 'MISMATCH at ' + str(np.flatnonzero(got != ref)[:5])}")
print("FUZZ", "PASSED" if bad == 0 else f"FAILED ({bad})")
sys.exit(1 if bad else 0)
