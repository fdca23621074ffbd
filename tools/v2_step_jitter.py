"""Per-step times of the v2 pipeline (analyze_batch with trend + guitar filters, financial logic filter) on the bench corpus,
with and without an `nvidia-smi -lms 200` poller running beside it (what bench.py's clock sampler does)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import core, batch
n_clips, sr, dur = 1024, 22050, 30.0
dev = torch.device("cuda", 0)
y = core.synth_events(n_clips, int(dur * sr), P.corpus.plan_events(n_clips, dur, sr), dev)


def step():
    r2 = batch.analyze_batch(y, sr=sr, hop_length=512, with_trend=True, with_guitar=True, nan_to_num=False)
    return batch.note_events_financial_batch(r2, sr=sr, hop_length=512)


def run(n, label):
    for _ in range(3):
        step()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(label, " ".join(f"{t:.1f}" for t in ts), " reserved GB", torch.cuda.memory_reserved() / 2**30)


run(20, "no poller :")
p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "200"],
                     stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
run(20, "with poller:")
p.terminate()
run(10, "no poller :")
