"""ncu / timing target: the post-STFT kernels of the bench step (mel_post, onset peaks) at cfg2 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import corpus, core
n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
base = torch.from_numpy(corpus.clip_batch(8, 30.0, 22050, first_seed=0)).to(dev)
y = base.repeat((n_clips + 7) // 8, 1)[:n_clips].contiguous()
feat = core.stft_features(y, sr=22050, want_mag=False, want_mel=True, want_rms=True)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(4):
    a = ev()
    post = core.mel_post(feat["mel"], feat["mel_max"], sr=22050, want_sdb=False, want_rake=False, want_onset=True)
    b = ev()
    pk = core.onset_peaks(post["onset_env"], post["env_minmax"], sr=22050)
    c = ev()
    torch.cuda.synchronize()
    print(f"mel_post {a.elapsed_time(b):.3f} ms  onset_peaks {b.elapsed_time(c):.3f} ms")
