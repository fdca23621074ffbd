"""Time K9 (aegis_resample_poly) on a cfg2-sized batch: python tools/resample_prof.py [n_clips] [seconds]"""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import spectrogram_midi_b200 as P

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
dev = torch.device("cuda:0")
for orig, target, dtype, ch in [(44100, 22050, torch.float32, 1), (88200, 22050, torch.int16, 1), (44100, 22050, torch.int16, 1), (44100, 22050, torch.int16, 2),
                                (48000, 22050, torch.float32, 1), (22050, 22050, torch.int16, 1)]:
    n_in = int(secs * orig)
    if dtype == torch.int16:
        x = torch.randint(-30000, 30000, (n_clips, n_in * ch), dtype=torch.int16, device=dev)
    else:
        x = torch.rand((n_clips, n_in * ch), device=dev) * 2 - 1
    for _ in range(2):
        y = P.core.resample_poly(x, orig, target, n_channels=ch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        y = P.core.resample_poly(x, orig, target, n_channels=ch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = (x.numel() * x.element_size() + y.numel() * 4) / 1e9
    print(f"{orig}->{target} {str(dtype).split('.')[-1]} x{ch}ch: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s algorithmic ({gb:.2f} GB)  "
          f"{n_clips * secs / ms * 1e3 / 1e6:.1f} M audio-s/s")
    del x, y
