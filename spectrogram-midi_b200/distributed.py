"""Multi-GPU execution of the hot path: one process per GPU, ``torch.distributed`` (NCCL) plumbing.

Two modes:

* **by clip** (BASELINE cfg2/3/5): clips are independent -- every ``ref=np.max`` is per clip -- so each
  rank analyses ``batch.shard_range(n_clips, rank, world)`` and nothing crosses NVLink on the data
  path.  ``gather_counts`` is an optional reporting collective.
* **one long clip in overlapping time windows** (BASELINE cfg4).  The reference's Turbo Mode
  (aegis_engine.py:183-216) cuts the clip into independent chunks and concatenates, which differs
  from the serial result at every seam; the target here is the *serial* result.  Frame-local stages
  (STFT, mel, RMS, YIN candidates) are computed for the rank's own frame range from a window of audio
  with a +-1024-sample halo; the couplings that remain are
    1. ``power_to_db(ref=np.max)``: the global mel-power maximum  -> ``all_reduce(MAX)`` of one float;
    2. the HMM is one chain over all frames.  ``mode="exact"`` all-gathers the sparse observations
       (a few hundred bytes per frame) and decodes the single chain (identical to one GPU);
       ``mode="windowed"`` decodes each rank's frames plus a burn-in margin on both sides and keeps the
       interior (faster; equal to the serial path wherever the paths have coalesced, which the tests
       measure);
    3. results are all-gathered as ragged per-frame arrays / note-event records.
  The collectives are latency-sized (bytes to a few MB); NCCL over NVSwitch is used as is.

The per-window analysis is injected (``analyze``) so the host-side planning / stitching / collectives are
testable on CPU with the gloo backend; the default is the CUDA path and there is no CPU fallback.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .batch import shard_range

N_FFT = 2048
FRAME_ALIGN = 8  # STFT tiles are 8 frames, YIN packs frame pairs: windows start on a tile boundary


# ------------------------------------------------------------------------------------------------
# process-group helpers
# ------------------------------------------------------------------------------------------------
def init_from_env(backend: Optional[str] = None) -> tuple:
    """(rank, world, local_rank); initialises the default group when launched by torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, world, local_rank


def _world(group=None) -> tuple:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def collective_device(group=None):
    """Where tensors handed to a collective must live: the rank's current CUDA device under NCCL (a CPU tensor makes
    NCCL fail on that rank while the others block), the host under gloo or without a group."""
    if dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return None


def all_reduce_max(t: torch.Tensor, group=None) -> torch.Tensor:
    if _world(group)[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def all_gather_ragged(t: torch.Tensor, group=None) -> List[torch.Tensor]:
    """All-gather tensors whose first dimension differs per rank (padded to the maximum, then trimmed)."""
    rank, world = _world(group)
    if world == 1:
        return [t]
    if t.dtype in (torch.int16, torch.bool) or t.dtype == getattr(torch, "uint16", None):
        # neither NCCL nor gloo has a 16-bit integer type: ship the bytes
        shape, dtype = tuple(t.shape[1:]), t.dtype
        row_bytes = int(np.prod(shape, dtype=np.int64)) * t.element_size()
        raw = t.contiguous().view(torch.uint8).reshape(t.shape[0], row_bytes)
        return [p.contiguous().view(dtype).reshape((p.shape[0],) + shape) for p in all_gather_ragged(raw, group)]
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in sizes]
    m = max(sizes)
    padded = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    padded[: t.shape[0]] = t
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return [o[:s] for o, s in zip(out, sizes)]


def gather_counts(local_count: int, device=None, group=None) -> List[int]:
    """Per-rank counts (e.g. clips analysed, events found) for reporting.  ``device=None`` picks the device the
    group's backend needs (a rank that owns zero clips has no tensor to borrow one from)."""
    t = torch.tensor([local_count], dtype=torch.int64, device=device if device is not None else collective_device(group))
    return [int(x[0]) for x in all_gather_ragged(t, group)]


# compact transport record (22 bytes, packed) for reporting; the product path gathers the kernels' own 40-byte records
EVENT_DTYPE = np.dtype([("note", "<i2"), ("start", "<i4"), ("end", "<i4"), ("velocity", "u1"), ("track", "u1"),
                        ("technique", "u1"), ("_pad", "u1"), ("confidence", "<f4"), ("slope", "<f4")])


def gather_note_events(events: np.ndarray, device=None, group=None, dtype=EVENT_DTYPE) -> np.ndarray:
    """All-gather note-event records from every rank, in rank order: north_star's only cross-GPU exchange of the
    long-clip path.  ``dtype`` is the record layout (``EVENT_DTYPE``, 22 B, or ``core.NOTE_EVENT_DTYPE``, the 40-byte
    records kernel K7 writes); the bytes travel as uint8 rows, padded to the largest rank and trimmed."""
    dtype = np.dtype(dtype)
    events = np.ascontiguousarray(events, dtype=dtype)
    raw = torch.from_numpy(events.view(np.uint8).reshape(len(events), dtype.itemsize).copy())
    if device is None:
        device = collective_device(group)
    if device is not None:
        raw = raw.to(device)
    parts = all_gather_ragged(raw, group)
    flat = torch.cat(parts).cpu().numpy()
    return flat.reshape(-1).view(dtype) if flat.size else np.zeros(0, dtype)


# ------------------------------------------------------------------------------------------------
# window planning for one long clip
# ------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Window:
    own_lo: int      # frames this rank contributes: [own_lo, own_hi)
    own_hi: int
    win_lo: int      # frames it analyses (own range + burn-in margin): [win_lo, win_hi)
    win_hi: int
    s0: int          # audio slice [s0, s1) handed to the kernels
    s1: int
    pad: int         # virtual zeros before s0 (non-zero only at the clip start)


def plan_windows(n_samples: int, hop_length: int, world: int, burn_frames: int) -> List[Window]:
    n_frames = 1 + n_samples // hop_length
    out = []
    for r in range(world):
        own = shard_range(n_frames, r, world)
        lo, hi = max(0, own.start - burn_frames), min(n_frames, own.stop + burn_frames)
        lo -= lo % FRAME_ALIGN  # keep the kernels' frame pairing / tiling identical to the full-clip run (bit-equal frames)
        first = lo * hop_length - N_FFT // 2
        s0 = max(0, first)
        s1 = min(n_samples, (hi - 1) * hop_length + N_FFT // 2) if hi > lo else s0
        out.append(Window(own.start, own.stop, lo, hi, s0, max(s0, s1), s0 - first))
    return out


# ------------------------------------------------------------------------------------------------
# per-window analysis back end (CUDA by default; injectable so the host logic is testable with gloo)
# ------------------------------------------------------------------------------------------------
class CudaBackend:
    """The product path: every call goes to the sm_100a kernels.  No CPU fallback."""

    def __init__(self, *, sr, hop_length, fmin, fmax, rake_sensitivity, device=None):
        from . import tables

        self.sr, self.hop, self.ratio = sr, hop_length, rake_sensitivity
        self.cfg = tables.pyin_config(float(sr), int(hop_length), float(fmin), float(fmax))
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())

    def features(self, y_win: np.ndarray, w: "Window") -> dict:
        """mel power / clip-local mel max / RMS / sparse pYIN observations for frames [win_lo, win_hi)."""
        from . import core

        n = w.win_hi - w.win_lo
        yd = torch.from_numpy(np.ascontiguousarray(y_win, dtype=np.float32)).to(self.device)[None]
        feat = core.stft_features(yd, sr=self.sr, hop_length=self.hop, want_mag=False, want_mel=True, want_rms=True,
                                  pad=w.pad, n_frames=n)
        obs = core.yin_candidates(yd, self.cfg, pad=w.pad, n_frames=n)
        mc = obs["max_cand"]
        return {"mel": feat["mel"], "mel_max": feat["mel_max"], "rms": feat["rms"][0],
                "cand_bin": obs["cand_bin"].view(n, mc), "cand_prob": obs["cand_prob"].view(n, mc),
                "cand_count": obs["cand_count"], "voiced_prob": obs["voiced_prob"][0]}

    def rake_mask(self, feat: dict, mel_max: torch.Tensor) -> torch.Tensor:
        from . import core

        post = core.mel_post(feat["mel"], mel_max, sr=self.sr, hop_length=self.hop, rake_ratio=self.ratio,
                             want_sdb=False, want_rake=True)
        return post["rake_mask"][0]

    def decode(self, cand_bin, cand_prob, cand_count, voiced_prob) -> dict:
        """Viterbi over one chain of sparse observations -> f0 (NaN unvoiced), voiced_flag."""
        from . import core

        T, mc = cand_bin.shape
        obs = dict(cand_bin=cand_bin.contiguous(), cand_prob=cand_prob.contiguous(), cand_count=cand_count.contiguous(),
                   voiced_prob=voiced_prob[None].contiguous(), n_frames=T, max_cand=mc)
        dec = core.viterbi_decode(obs, self.cfg, 1)
        return {"f0": dec["f0"][0], "voiced_flag": dec["voiced_flag"][0], "states": dec["states"][0]}

    def note_events(self, res: dict, **kwargs) -> np.ndarray:
        """``get_midi_events`` (midi_logic.py:32-148) on full-length host arrays with kernel K7 -> records of
        ``core.NOTE_EVENT_DTYPE`` in time order.  MIDI notes come from the Viterbi states through the host-made table."""
        from . import core, tables

        def dev(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.device)[None]

        lut = np.array([int(round(float(tables.hz_to_midi(f)))) for f in self.cfg.freqs], dtype=np.int16)
        states = res.get("states")
        out = core.note_events(dev(res["rake_mask"], np.uint8), dev(np.nan_to_num(res["f0"]), np.float64),
                               dev(res["voiced_flag"], np.uint8), dev(res["voiced_probs"], np.float64), dev(res["rms"], np.float32),
                               sr=self.sr, hop_length=self.hop,
                               pitch_index=None if states is None else dev(states, np.int16),
                               note_lut=None if states is None else torch.from_numpy(lut).to(self.device), **kwargs)
        n = int(out["n_events"][0])
        if n > out["max_events"]:
            raise RuntimeError(f"{n} note events exceed the buffer of {out['max_events']}")
        return out["events"][0, :n].cpu().numpy().reshape(-1).view(core.NOTE_EVENT_DTYPE).copy()


MIN_MARGIN_FRAMES = 32  # the rake run-length gate looks at up to 30 neighbouring columns


def analyze_long_clip(y, *, sr: float, hop_length: int = 512, fmin: Optional[float] = None, fmax: Optional[float] = None,
                      rake_sensitivity: float = 0.6, mode: str = "exact", burn_seconds: float = 2.0, group=None,
                      backend=None, windows_per_rank: int = 1, return_events: bool = False,
                      event_kwargs: Optional[dict] = None, solo: bool = False) -> dict:
    """Perception arrays of ONE long clip computed by all ranks of ``group`` (each rank can read the clip,
    or at least its own windows, from host memory).  Every rank returns the full-length result:
    ``rake_mask, f0 (NaN unvoiced), voiced_flag, voiced_probs, rms`` as numpy arrays (aegis_engine.py:72-75).

    The clip is cut into ``world * windows_per_rank`` overlapping windows; rank r owns windows
    ``r*windows_per_rank ...`` (``windows_per_rank > 1`` bounds device memory for hour-long clips and lets a
    single GPU exercise the same stitching).

    ``return_events=True`` adds ``events``: the v1 note events of the whole clip (``get_midi_events``,
    midi_logic.py:32-148; records of ``core.NOTE_EVENT_DTYPE``, keyword arguments in ``event_kwargs``), assembled with
    the path's one event collective: the note state machine is a single sequential chain over the clip (and its
    dB scale needs the clip-wide RMS maximum), so every rank runs it on the gathered frame arrays -- microseconds per
    second of audio -- keeps the events that START inside its own frames, and ``gather_note_events`` puts the list
    together (NCCL over NVLink on a GPU box).  The list equals the single-GPU one by construction.
    ``solo=True`` ignores the process group: this rank analyses the whole clip alone (the serial answer a multi-rank
    run is compared with).
    """
    if mode not in ("exact", "windowed"):
        raise ValueError("mode must be 'exact' or 'windowed'")
    rank, world = (0, 1) if solo else _world(group)
    reduce_max = (lambda t: t) if solo else (lambda t: all_reduce_max(t, group))
    gather = (lambda t: [t]) if solo else (lambda t: all_gather_ragged(t, group))
    if backend is None:
        from .batch import C6, E2

        backend = CudaBackend(sr=sr, hop_length=hop_length, fmin=E2 if fmin is None else fmin,
                              fmax=C6 if fmax is None else fmax, rake_sensitivity=rake_sensitivity)
    n_samples = int(y.shape[-1])
    burn = int(round(burn_seconds * sr / hop_length)) if mode == "windowed" else 0
    plan = plan_windows(n_samples, hop_length, world * windows_per_rank, max(burn, MIN_MARGIN_FRAMES))
    mine = plan[rank * windows_per_rank : (rank + 1) * windows_per_rank]
    feats = [backend.features(np.asarray(y[w.s0 : w.s1], dtype=np.float32), w) for w in mine]

    # coupling 1: power_to_db(ref=np.max) needs the clip-global maximum of the mel power
    local_max = feats[0]["mel_max"].clone()
    for f in feats[1:]:
        local_max = torch.maximum(local_max, f["mel_max"])
    mel_max = reduce_max(local_max)

    own_parts = []
    for w, feat in zip(mine, feats):
        rake = backend.rake_mask(feat, mel_max)
        a, b = w.own_lo - w.win_lo, w.own_hi - w.win_lo
        own = {"rms": feat["rms"][a:b], "rake_mask": rake[a:b], "voiced_probs": feat["voiced_prob"][a:b]}
        if mode == "windowed":
            # coupling 2, approximate: decode own frames + burn-in margin, keep the interior
            dec = backend.decode(feat["cand_bin"], feat["cand_prob"], feat["cand_count"], feat["voiced_prob"])
            own["f0"], own["voiced_flag"] = dec["f0"][a:b], dec["voiced_flag"][a:b]
            if "states" in dec:
                own["states"] = dec["states"][a:b]
        else:
            for k in ("cand_bin", "cand_prob", "cand_count"):
                own[k] = feat[k][a:b]
        own_parts.append(own)
    own = {k: torch.cat([o[k] for o in own_parts]) for k in own_parts[0]}
    full = {k: torch.cat(gather(v.contiguous())) for k, v in own.items()}
    if mode == "exact":
        # coupling 2, exact: the gathered sparse observations are decoded as the single chain they are
        dec = backend.decode(full.pop("cand_bin"), full.pop("cand_prob"), full.pop("cand_count"), full["voiced_probs"])
        full["f0"], full["voiced_flag"] = dec["f0"], dec["voiced_flag"]
        if "states" in dec:
            full["states"] = dec["states"]
    res = {
        "rake_mask": full["rake_mask"].cpu().numpy().astype(bool), "f0": full["f0"].cpu().numpy(),
        "voiced_flag": full["voiced_flag"].cpu().numpy().astype(bool), "voiced_probs": full["voiced_probs"].cpu().numpy(),
        "rms": full["rms"].cpu().numpy(),
    }
    if "states" in full:
        res["states"] = full["states"].cpu().numpy()
    if return_events:
        ev = backend.note_events(res, **(event_kwargs or {}))
        lo, hi = (mine[0].own_lo, mine[-1].own_hi) if mine else (0, 0)
        keep = (ev["start"] >= lo) & (ev["start"] < hi)
        res["events_local"] = int(keep.sum())
        res["events"] = ev[keep] if solo else gather_note_events(ev[keep], group=group, dtype=ev.dtype)
    return res


# ------------------------------------------------------------------------------------------------
# by-clip sharding
# ------------------------------------------------------------------------------------------------
def analyze_clips_sharded(load_clips: Callable[[Sequence[int]], torch.Tensor], n_clips: int, *, sr: float, group=None,
                          **analyze_kwargs) -> dict:
    """Each rank analyses its contiguous share of ``n_clips`` clips (``load_clips(indices)`` returns the CUDA
    batch for those clip indices).  No collective touches the data path; the per-rank clip counts are gathered
    for reporting only."""
    from . import batch

    rank, world = _world(group)
    mine = shard_range(n_clips, rank, world)
    y = load_clips(list(mine))
    res = batch.analyze_batch(y, sr=sr, **analyze_kwargs) if len(mine) else {}
    res["clip_indices"] = list(mine)
    res["clips_per_rank"] = gather_counts(len(mine), group=group)   # the device the backend needs, also on a rank without clips
    return res
