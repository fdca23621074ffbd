"""Does K2 (FP32-bound block sums + issue-bound frame stage) overlap with K3 (latency / issue-bound decoder) when the batch
is decoded piece by piece on separate streams?  Device-resident clips; prints the time of the sequential whole-batch
calls and of the rolling schedule for several piece sizes / decoder streams, and checks the decoded states are equal.
usage: overlap_probe.py [n_clips]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import core, tables, batch

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sr, dur = 22050, 30.0
dev = torch.device("cuda", 0)
plan = P.corpus.plan_events(n_clips, dur, sr)
y = core.synth_events(n_clips, int(dur * sr), plan, dev)
cfg = tables.pyin_config(float(sr), 512, batch.E2, batch.C6)
T = core.frame_count(y.shape[1], 512)
mc = cfg.max_troughs
obs = {"cand_bin": torch.empty((n_clips * T, mc), dtype=torch.int16, device=dev),
       "cand_prob": torch.empty((n_clips * T, mc), dtype=torch.float64, device=dev),
       "cand_count": torch.empty((n_clips * T,), dtype=torch.int32, device=dev),
       "voiced_prob": torch.empty((n_clips, T), dtype=torch.float64, device=dev)}
main = torch.cuda.current_stream(dev)


def k2(c0, c1):
    sl = slice(c0 * T, c1 * T)
    core.yin_candidates(y[c0:c1], cfg, out={"cand_bin": obs["cand_bin"][sl], "cand_prob": obs["cand_prob"][sl],
                                            "cand_count": obs["cand_count"][sl], "voiced_prob": obs["voiced_prob"][c0:c1]})


def k3(c0, c1):
    sl = slice(c0 * T, c1 * T)
    o = {"cand_bin": obs["cand_bin"][sl], "cand_prob": obs["cand_prob"][sl], "cand_count": obs["cand_count"][sl],
         "voiced_prob": obs["voiced_prob"][c0:c1], "n_frames": T, "max_cand": mc}
    return core.viterbi_decode(o, cfg, c1 - c0)


def timed(fn, reps=3):
    best = 1e9
    out = None
    for _ in range(reps + 1):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(main); out = fn(); b.record(main); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


def sequential():
    k2(0, n_clips)
    return [k3(0, n_clips)]


def rolling(piece, n_dec):
    decs = [torch.cuda.Stream(device=dev) for _ in range(n_dec)]

    def run():
        outs, done = [], []
        for i, c0 in enumerate(range(0, n_clips, piece)):
            c1 = min(n_clips, c0 + piece)
            k2(c0, c1)
            e = torch.cuda.Event(); e.record(main)
            sd = decs[i % n_dec]
            with torch.cuda.stream(sd):
                sd.wait_event(e)
                outs.append(k3(c0, c1))
                d = torch.cuda.Event(); d.record(sd); done.append(d)
        for d in done:
            main.wait_event(d)
        return outs
    return run


t_seq, ref = timed(sequential)
ref_states = ref[0]["states"].clone()
print(f"sequential K2 + K3, {n_clips} clips: {t_seq:.2f} ms")
for piece, n_dec in ((148, 2), (148, 3), (148, 4), (296, 2), (296, 1), (74, 4), (74, 6), (148, 7)):
    t, outs = timed(rolling(piece, n_dec))
    st = torch.cat([o["states"] for o in outs])
    print(f"rolling piece {piece:4d}, {n_dec} decoder streams: {t:.2f} ms  ({t_seq / t:.2f}x)  states equal: {bool(torch.equal(st, ref_states))}")
