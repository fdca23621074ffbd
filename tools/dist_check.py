"""torchrun check of the multi-GPU paths on real GPUs (run with gpurun --gpus 2):
   torchrun --nproc-per-node 2 tools/dist_check.py
 1. one long clip analysed by all ranks (exact + windowed) == the single-GPU full-clip result
 2. by-clip sharding: every rank analyses its share, counts gathered."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import spectrogram_midi_b200 as P
from spectrogram_midi_b200 import distributed as D, batch, corpus

rank, world, local = D.init_from_env()
dev = torch.device("cuda", local)
sr = 22050
y = np.concatenate([corpus.random_clip(100 + i, 10.0, sr) for i in range(6)])  # 60 s
ref = P.AegisEngine(sample_rate=sr).audio_to_midi(y, None)
f0_ref = ref["f0"].copy(); f0_ref[~ref["voiced_flag"]] = np.nan
ok = True
for mode in ("exact", "windowed"):
    for wpr in (1, 3):
        res = D.analyze_long_clip(y, sr=sr, mode=mode, windows_per_rank=wpr)
        same_v = (res["voiced_flag"] == ref["voiced_flag"]).mean()
        same_f = np.array_equal(np.nan_to_num(res["f0"]), np.nan_to_num(f0_ref))
        same_r = np.array_equal(res["rake_mask"], ref["rake_mask"]) and np.allclose(res["rms"], ref["rms"], rtol=1e-6)
        same_p = np.array_equal(res["voiced_probs"], ref["voiced_probs"])
        if rank == 0:
            print(f"[world {world}] long clip mode={mode} windows/rank={wpr}: voiced agree {same_v:.4f} f0 identical {same_f} rake+rms {same_r} probs identical {same_p}", flush=True)
        if mode == "exact":
            ok &= same_f and same_v == 1.0 and same_r and same_p
        else:
            ok &= same_v > 0.99 and same_r and same_p
n_clips = 10
clips = corpus.clip_batch(n_clips, 4.0, sr, first_seed=7)
res = D.analyze_clips_sharded(lambda idx: torch.from_numpy(clips[idx]).to(dev), n_clips, sr=sr)
full = batch.analyze_batch(torch.from_numpy(clips).to(dev), sr=sr)
mine = res["clip_indices"]
ok &= torch.equal(res["states"], full["states"][mine]) and sum(res["clips_per_rank"]) == n_clips
if rank == 0:
    print(f"[world {world}] by-clip sharding: clips per rank {res['clips_per_rank']}, states identical to the unsharded batch: {torch.equal(res['states'], full['states'][mine])}", flush=True)
# fewer clips than ranks: the rank without clips still joins the count gather, on the device NCCL needs
one = D.analyze_clips_sharded(lambda idx: torch.from_numpy(clips[idx]).to(dev), 1, sr=sr)
ok &= one["clips_per_rank"] == [1] + [0] * (world - 1) and (len(one["clip_indices"]) == (1 if rank == 0 else 0))
if rank == 0:
    print(f"[world {world}] one clip over {world} ranks: clips per rank {one['clips_per_rank']}", flush=True)
# note events of the long clip: every rank contributes the events that start in its frames, gathered over NCCL
ev = D.analyze_long_clip(y, sr=sr, mode="exact", return_events=True)
want = P.AegisEngine(sample_rate=sr).note_events(ref)
same_ev = [(int(r["note"]), int(r["start"]), int(r["end"])) for r in ev["events"]] == [(e["note"], e["start"], e["end"]) for e in want]
counts = D.gather_counts(int(ev["events_local"]))
ok &= same_ev and sum(counts) == len(want)
if rank == 0:
    print(f"[world {world}] event gather: {len(want)} events, per rank {counts}, identical to the single-GPU list: {same_ev}", flush=True)
t = torch.tensor([int(ok)], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
if rank == 0:
    print("DIST CHECK", "PASSED" if int(t) else "FAILED", flush=True)
sys.exit(0 if int(t) else 1)
