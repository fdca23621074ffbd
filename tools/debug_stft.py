import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import corpus
from oracle import librosa_ref as L
dev = torch.device("cuda", 0)
y = corpus.test_track(22050, 0)
yd = torch.from_numpy(y).to(dev)[None]
a = P.core.stft_features(yd, want_mag=True)["mag"][0].cpu().numpy()
b = P.core.stft_features(yd, want_mag=True)["mag"][0].cpu().numpy()
print("deterministic:", np.array_equal(a, b))
c = P.core.stft_features(yd * 2, want_mag=True)["mag"][0].cpu().numpy()
d = np.abs(2 * a - c)
print("homogeneity max dev", d.max(), "at", np.unravel_index(d.argmax(), d.shape), "value", a.flat[d.argmax()], "n nonzero dev", (d > 0).sum())
ref = L.stft_magnitude(y)
tol = 1e-4 * np.abs(ref) + 1e-5 * ref.max(axis=0, keepdims=True) + 1e-30
bad = np.abs(a - ref) > tol
print("bad bins", bad.sum(), "frames with bad:", np.flatnonzero(bad.any(axis=0)))
for t in np.flatnonzero(bad.any(axis=0))[:12]:
    k = np.flatnonzero(bad[:, t])
    print(f" frame {t}: nbad {len(k)} framemax {ref[:, t].max():.3e} gpu max {a[:, t].max():.3e} worst abs {np.abs(a[:, t]-ref[:, t]).max():.3e} first bins {k[:5]} ref {ref[k[:3], t]} gpu {a[k[:3], t]}")
# distribution of relative-to-framemax error
fm = ref.max(axis=0, keepdims=True) + 1e-30
rel = np.abs(a - ref) / fm
print("percentiles of err/framemax:", np.percentile(rel, [50, 90, 99, 99.9, 100]))
