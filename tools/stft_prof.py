"""ncu target: a few launches of the fused STFT kernel at BASELINE cfg2 shape (clips x 30 s @ 22.05 kHz).
Usage: python tools/stft_prof.py [n_clips] [launches]   (prints CUDA-event ms per launch when run plain)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import corpus

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 5
pitch_to = int(sys.argv[3]) if len(sys.argv) > 3 else 1  # round the |X| row pitch up to a multiple of this many frames
dev = torch.device("cuda", 0)
y = corpus.synth_batch_device(n_clips, 30.0, 22050, device=dev) if hasattr(corpus, "synth_batch_device") else None
if y is None:
    base = torch.from_numpy(corpus.clip_batch(8, 30.0, 22050, first_seed=0)).to(dev)
    y = base.repeat((n_clips + 7) // 8, 1)[:n_clips].contiguous()
T = 1 + y.shape[1] // 512
Tp = (T + pitch_to - 1) // pitch_to * pitch_to
mag = torch.empty((n_clips, 1025, Tp), device=dev)[:, :, :T]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(launches + 1)]
for i in range(launches):
    ev[i].record()
    P.core.stft_features(y, sr=22050, want_mag=True, want_mel=True, want_rms=True, mag_out=mag)
ev[launches].record()
torch.cuda.synchronize()
print(f"pitch {Tp}: ms per launch:", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(launches)])
