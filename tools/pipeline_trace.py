"""Timeline of batch.TranscribePipeline on BASELINE-sized input (1024 x 30 s PCM clips): where an end-to-end step spends its time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import core, batch
n_clips, sr, dur = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 22050, 30.0
dev = torch.device("cuda", 0)
y = core.synth_events(n_clips, int(dur * sr), P.corpus.plan_events(n_clips, dur, sr), dev)
pcm = torch.empty((n_clips, int(dur * sr)), dtype=torch.int16, pin_memory=True)
pcm.copy_((y * 32767.0).round().to(torch.int16))
kw = {}
if len(sys.argv) > 2: kw["chunk_clips"] = int(sys.argv[2])
if len(sys.argv) > 3: kw["group_clips"] = int(sys.argv[3])
pipe = batch.TranscribePipeline(n_clips, int(dur * sr), sr=sr, device=dev, **kw)
for _ in range(2): pipe.run(pcm)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): pipe.run(pcm)
torch.cuda.synchronize(); print(f"chunk {pipe.chunk} group {pipe.group}: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms per run", getattr(pipe, "group_decision", None))
pipe.trace = []
pipe.run(pcm)
for label, ms in pipe.trace: print(f"{ms:8.2f} ms  {label}")
