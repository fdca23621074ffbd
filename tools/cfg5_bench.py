"""BASELINE cfg5 on one GPU: 100 000 synthetic 10 s clips @ 22 050 Hz, full perception incl. rake mask, pYIN, RMS,
financial trend filter, guitar filters, both note-event logic filters (v1, v2) and the MIDI files of the v2 events, streamed in sub-batches rendered on the
device (88 GB of audio in total never exist at once).  Usage: python tools/cfg5_bench.py [n_clips] [sub_batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpy as np
import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import batch, core, corpus, midi_writer

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
sub = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sr, dur = 22050, 10.0
dev = torch.device("cuda", 0)
n_samples = int(sr * dur)
t_plan = t_synth = 0.0
ev_gpu = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
gpu_ms = 0.0
n_events = n_frames_voiced = n_rake = n_mute = n_fin = midi_bytes = 0
fin_ms = t_midi = 0.0
ev_fin = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
t0 = time.time()
for first in range(0, n_total, sub):
    n = min(sub, n_total - first)
    a = time.time()
    plan = corpus.plan_events(n, dur, sr, first_seed=first)
    b = time.time()
    y = core.synth_events(n, n_samples, plan, dev)
    torch.cuda.synchronize()
    c = time.time()
    t_plan += b - a
    t_synth += c - b
    ev_gpu[0].record()
    res = batch.analyze_batch(y, sr=sr, with_trend=True, with_guitar=True, nan_to_num=False)
    v1 = dict(res)
    ev = batch.note_events_batch(v1, sr=sr, confidence_threshold=0.7)
    ev_gpu[1].record()
    ev_fin[0].record()
    fin = batch.note_events_financial_batch(res, sr=sr)
    ev_fin[1].record()
    torch.cuda.synchronize()
    gpu_ms += ev_gpu[0].elapsed_time(ev_gpu[1])
    fin_ms += ev_fin[0].elapsed_time(ev_fin[1])
    # v2 MIDI files for every clip of the sub-batch (native writer on the event records, one D2H copy)
    m0 = time.time()
    rec = fin["events"].cpu().numpy().view(core.FIN_EVENT_DTYPE)[..., 0]
    cnt = fin["n_events"].cpu().numpy()
    for c in range(n):
        if cnt[c]:
            midi_bytes += len(midi_writer.smf_bytes_financial(rec[c, : cnt[c]], sr, 512))
    t_midi += time.time() - m0
    n_fin += int(cnt.sum())
    n_events += int(ev["n_events"].sum())
    n_frames_voiced += int(res["voiced_flag"].sum())
    n_rake += int(res["rake_mask"].sum())
    n_mute += int(res["mute_mask"].sum())
wall = time.time() - t0
audio_s = n_total * dur
print(f"cfg5: {n_total} clips x {dur:.0f} s in sub-batches of {sub}: perception + trend + guitar filters + note events "
      f"{gpu_ms / 1e3:.2f} s on the device = {audio_s / (gpu_ms / 1e3):.0f} audio-s/s; wall {wall:.1f} s "
      f"(host event planning {t_plan:.1f} s, device synthesis {t_synth:.1f} s); "
      f"{n_events} note events, {n_frames_voiced} voiced frames, {n_rake} rake frames, {n_mute} palm-mute frames")
print(f"      v2 logic filter (K5 x2 + K8) {fin_ms / 1e3:.2f} s on the device, {n_fin} events; their {n_total} MIDI files "
      f"({midi_bytes / 1e6:.1f} MB) serialised on the host in {t_midi:.1f} s")
