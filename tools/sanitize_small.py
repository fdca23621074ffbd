"""Small end-to-end pass: every kernel once on tiny inputs (compute-sanitizer is closed on this pool; kept as a quick manual check)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import corpus
dev = torch.device("cuda", 0)
y = torch.from_numpy(np.stack([corpus.random_clip(s, 1.5, 22050) for s in range(3)])).to(dev)
for hop in (512, 256):
    f = P.core.stft_features(y, sr=22050, hop_length=hop, want_mag=True, want_mel=True, want_rms=True)
f = P.core.stft_features(y[:, :4099].contiguous(), sr=22050, want_mag=True, want_mel=True, want_rms=True)   # LSU store path (pitch 9 frames... unaligned)
res = P.batch.analyze_batch(y, sr=22050, with_trend=True, with_guitar=True, with_onsets=True, with_sdb=True)
ev = P.batch.note_events_batch(P.batch.analyze_batch(y, sr=22050), sr=22050)
torch.cuda.synchronize()
print("ok", int(ev["n_events"].sum()), float(f["mag"].abs().max()))
