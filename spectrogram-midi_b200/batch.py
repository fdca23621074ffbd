"""Batched perception on one GPU: the whole Aegis analysis front end for many clips at once.

This is the data-parallel form of ``AegisEngine.audio_to_midi`` (aegis_engine.py:41-75): every
clip is independent (all ``ref=np.max`` are per clip), so a batch is just a leading dimension and
multi-GPU work is a partition of clip indices (``shard_range``) with no data-path collective.
Inputs and outputs are CUDA tensors; ``to_host`` converts a result to the reference's dict of
numpy arrays.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import core, tables

E2 = tables.note_to_hz("E2")
C6 = tables.note_to_hz("C6")


def shard_range(n_items: int, rank: int, world_size: int) -> range:
    """Contiguous, balanced partition of ``n_items`` clip indices (first ``n % world`` ranks get one more)."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def spectral_features(y: torch.Tensor, *, sr: float, hop_length: int = 512, rake_sensitivity: float = 0.6,
                      with_mag: bool = False, with_sdb: bool = False, with_rake: bool = True,
                      with_onsets: bool = True, with_rms: bool = True, mag_out: Optional[torch.Tensor] = None) -> dict:
    """STFT |X| / mel dB / rake mask / RMS / onset envelope + onset frames for a batch (BASELINE cfg2)."""
    feat = core.stft_features(y, sr=sr, hop_length=hop_length, want_mag=with_mag, want_mel=True,
                              want_rms=with_rms, mag_out=mag_out)
    post = core.mel_post(feat["mel"], feat["mel_max"], sr=sr, hop_length=hop_length, rake_ratio=rake_sensitivity,
                         want_sdb=with_sdb, want_rake=with_rake, want_onset=with_onsets)
    out = {"n_frames": feat["n_frames"]}
    if with_mag:
        out["mag"] = feat["mag"]
    if with_rms:
        out["rms"] = feat["rms"]
    if with_sdb:
        out["S_dB"] = post["S_dB"]
    if with_rake:
        out["rake_mask"] = post["rake_mask"]
    if with_onsets:
        out["onset_env"] = post["onset_env"]
        pk = core.onset_peaks(post["onset_env"], post["env_minmax"], sr=sr, hop_length=hop_length)
        out["onset_peaks"], out["n_onsets"] = pk["peaks"], pk["n_peaks"]
    return out


def analyze_batch(y: torch.Tensor, *, sr: float, hop_length: int = 512, rake_sensitivity: float = 0.6,
                  fmin: float = E2, fmax: float = C6, with_sdb: bool = False, with_onsets: bool = False,
                  with_trend: bool = False, nan_to_num: bool = True, clips_per_launch: Optional[int] = None,
                  with_guitar: bool = False) -> dict:
    """Full perception phase for a batch: keys of aegis_engine.py:72-75 (+ optional extras).

    ``f0`` follows v1 (`np.nan_to_num`, aegis_engine.py:69) unless ``nan_to_num=False`` (v2 keeps NaN,
    aegis_engine_financial.py:122).  ``with_trend`` adds ``multi_filter_consensus`` on the NaN-masked f0
    (midi_logic_financial.py:158-161).  ``with_guitar`` applies ``apply_guitar_filters`` the way
    ``audio_to_midi_financial`` does (aegis_engine_financial.py:132-147): octave-corrected f0 / voiced flags,
    enhanced rake mask, and the palm-mute frames removed from the voiced flags; adds ``mute_mask`` and
    ``distortion``.
    """
    spec = spectral_features(y, sr=sr, hop_length=hop_length, rake_sensitivity=rake_sensitivity,
                             with_sdb=with_sdb or with_guitar, with_onsets=with_onsets)
    pit = core.pyin_batch(y, sr=sr, fmin=fmin, fmax=fmax, hop_length=hop_length, clips_per_launch=clips_per_launch)
    f0_nan = pit["f0"]
    extra = {}
    if with_guitar:
        gf = core.guitar_filters(spec["S_dB"], sr=sr, hop_length=hop_length, f0=f0_nan, voiced_flag=pit["voiced_flag"],
                                 rake_mask=spec["rake_mask"])
        f0_nan = gf["f0"]
        spec["rake_mask"] = gf["rake_mask"]
        pit["voiced_flag"] = gf["voiced"] & (gf["mute_mask"] ^ 1)   # voiced_flag & ~mute_mask (:147)
        extra = {"mute_mask": gf["mute_mask"], "distortion": gf["distortion"]}
        if not with_sdb:
            del spec["S_dB"]
    out = {
        **extra,
        "rake_mask": spec["rake_mask"], "voiced_flag": pit["voiced_flag"], "voiced_probs": pit["voiced_prob"],
        "rms": spec["rms"], "f0": torch.nan_to_num(f0_nan, nan=0.0) if nan_to_num else f0_nan,
        "states": pit["states"], "n_frames": spec["n_frames"],
    }
    for k in ("S_dB", "onset_env", "onset_peaks", "n_onsets"):
        if k in spec:
            out[k] = spec[k]
    if with_trend:
        tr = core.trend_filters(f0_nan, want=("consensus", "consensus_conf"))
        out["trend"], out["trend_conf"] = tr["consensus"], tr["consensus_conf"]
    return out


def note_events_batch(result: dict, *, sr: float, hop_length: int = 512, fmin: float = E2, fmax: float = C6,
                      confidence_threshold: float = 0.7, **kwargs) -> dict:
    """Logic-filter phase for a whole batch on the device: ``get_midi_events`` (midi_logic.py:32-148) applied to an
    ``analyze_batch`` result (v1 arrays: f0 with zeros).  The note numbers come from the Viterbi states through a
    table computed with the reference's expression ``int(round(hz_to_midi(f)))`` on the pYIN frequency grid."""
    cfg = tables.pyin_config(float(sr), hop_length, fmin, fmax)
    dev = result["f0"].device
    lut = core._dev_tensor(("note_lut",) + cfg.cache_key, dev,     # uploaded once per configuration and device
                           lambda: np.array([int(round(float(tables.hz_to_midi(f)))) for f in cfg.freqs], dtype=np.int16))
    f0 = torch.nan_to_num(result["f0"], nan=0.0)
    return core.note_events(result["rake_mask"], f0, result["voiced_flag"], result["voiced_probs"], result["rms"],
                            sr=sr, hop_length=hop_length, confidence_threshold=confidence_threshold,
                            pitch_index=result["states"], note_lut=lut, **kwargs)


def note_events_financial_batch(result: dict, *, sr: float, hop_length: int = 512, confidence_threshold=None, **kwargs) -> dict:
    """v2 logic-filter phase for a whole batch on the device: ``get_midi_events_financial``
    (midi_logic_financial.py:117-388) applied to an ``analyze_batch(nan_to_num=False)`` result; palm-muted frames
    (``with_guitar=True``) are removed from the voiced flags first, as aegis_engine_financial.py:147 does."""
    voiced = result["voiced_flag"].to(torch.uint8)
    if "mute_mask" in result:
        voiced = voiced & (result["mute_mask"].to(torch.uint8) ^ 1)
    return core.note_events_financial(result["rake_mask"], result["f0"], voiced, result["voiced_probs"], result["rms"],
                                      sr=sr, hop_length=hop_length, confidence_threshold=confidence_threshold, **kwargs)


def to_host(result: dict, clip: int, y: Optional[np.ndarray] = None) -> dict:
    """One clip of a batch result as the reference's perception dict (numpy, aegis_engine.py:72-75)."""
    host = {
        "rake_mask": result["rake_mask"][clip].cpu().numpy().astype(bool),
        "f0": result["f0"][clip].cpu().numpy(),
        "voiced_flag": result["voiced_flag"][clip].cpu().numpy().astype(bool),
        "voiced_probs": result["voiced_probs"][clip].cpu().numpy(),
        "rms": result["rms"][clip].cpu().numpy(),
    }
    if y is not None:
        host["y"] = y
    for k in ("S_dB", "onset_env", "trend", "trend_conf"):
        if k in result:
            host[k] = result[k][clip].cpu().numpy()
    if "mute_mask" in result:
        host["mute_mask"] = result["mute_mask"][clip].cpu().numpy().astype(bool)
        host["distortion"] = core.DISTORTION_LABELS[int(result["distortion"][clip])]
    if "onset_peaks" in result:
        host["onset_frames"] = np.flatnonzero(result["onset_peaks"][clip].cpu().numpy())
    return host


class HostPipeline:
    """Host-buffer plugin call for batches: pinned host audio in, host results out.

    The batch is cut into chunks of ``chunk_clips`` clips; chunk i+1 is copied host->device on a copy
    stream while chunk i runs through the kernels (double-buffered device input), and the small
    per-frame results (RMS, onset envelope, onset flags) are copied back as each chunk finishes.
    |X| is produced in HBM per chunk (it is the input of the onset path) and not copied to the host.

    With ``pcm_rate`` the host batch is 16-bit PCM as it sits in a WAV file (int16 ``[n_clips, n_source_frames *
    pcm_channels]``, channels interleaved, at ``pcm_rate`` Hz): half the PCIe bytes of float32, and the ingest kernel
    K9 does what ``librosa.load`` does after the file read (scale, mix down, resample to ``sr``) on the device.
    ``n_samples`` stays the clip length at the engine rate.
    """

    def __init__(self, n_clips: int, n_samples: int, *, sr: float, hop_length: int = 512, device=None, chunk_clips: int = 128,
                 pcm_rate: Optional[int] = None, pcm_channels: int = 1, copy_streams: int = 2):
        self.sr, self.hop = sr, hop_length
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.n_clips, self.n_samples = n_clips, n_samples
        self.chunk = min(chunk_clips, n_clips)
        self.T = core.frame_count(n_samples, hop_length)
        self.pcm_rate, self.pcm_channels = pcm_rate, pcm_channels
        if pcm_rate is None:
            self.in_shape, in_dtype = (n_clips, n_samples), torch.float32
        else:
            g = int(np.gcd(int(pcm_rate), int(sr)))
            up, down = int(sr) // g, int(pcm_rate) // g
            n_src = -(-n_samples * down // up)                 # fewest source frames whose resampled length reaches n_samples
            while -(-n_src * up // down) > n_samples:
                n_src -= 1
            if -(-n_src * up // down) != n_samples:
                raise ValueError(f"no source length at {pcm_rate} Hz resamples to exactly {n_samples} samples at {sr} Hz")
            self.in_shape, in_dtype = (n_clips, n_src * pcm_channels), torch.int16
        self.nbuf = max(2, copy_streams + 1)      # one chunk in the kernels, one per copy stream in flight
        self.inbuf = [torch.empty((self.chunk, self.in_shape[1]), dtype=in_dtype, device=self.dev) for _ in range(self.nbuf)]
        self.mag = core.alloc_frames((self.chunk, core.N_BINS), self.T, self.dev)
        self.copy_streams = [torch.cuda.Stream(device=self.dev) for _ in range(max(1, copy_streams))]
        self.in_ready = [torch.cuda.Event() for _ in range(self.nbuf)]
        self.in_free = [torch.cuda.Event() for _ in range(self.nbuf)]
        self.rms = torch.empty((n_clips, self.T), dtype=torch.float32, pin_memory=True)
        self.env = torch.empty((n_clips, self.T), dtype=torch.float32, pin_memory=True)
        self.peaks = torch.empty((n_clips, self.T), dtype=torch.uint8, pin_memory=True)
        self.h2d_bytes = n_clips * self.in_shape[1] * (4 if pcm_rate is None else 2)
        self.d2h_bytes = n_clips * self.T * (4 + 4 + 1)

    def run(self, y_host: torch.Tensor) -> dict:
        if y_host.is_cuda or tuple(y_host.shape) != self.in_shape or y_host.dtype != self.inbuf[0].dtype:
            raise ValueError(f"expected a host {self.inbuf[0].dtype} tensor {self.in_shape}")
        main = torch.cuda.current_stream(self.dev)
        # chunks of `chunk` clips; the last one is cut into a half and two quarters: what remains to be done after the last
        # byte has landed is one small chunk's kernels, not a full one's (0.85 -> 0.3 ms of a 25 ms step)
        pieces = [(c0, min(self.chunk, self.n_clips - c0)) for c0 in range(0, self.n_clips, self.chunk)]
        if len(pieces) >= 4 and pieces[-1][1] >= 16:
            c0, n = pieces.pop()
            h, q = n // 2, n // 4
            pieces += [(c0, h), (c0 + h, q), (c0 + h + q, n - h - q)]
        starts = [c for c, _ in pieces]
        sizes = [n for _, n in pieces]
        for b in range(self.nbuf):
            self.in_free[b].record(main)

        def issue_copy(i):   # one cudaMemcpyAsync per chunk, on the chunk's copy stream
            b, c0 = i % self.nbuf, starts[i]
            n = sizes[i]
            cs = self.copy_streams[i % len(self.copy_streams)]
            with torch.cuda.stream(cs):
                cs.wait_event(self.in_free[b])
                self.inbuf[b][:n].copy_(y_host[c0 : c0 + n], non_blocking=True)
                self.in_ready[b].record(cs)

        ahead = self.nbuf - 1
        for i in range(min(ahead, len(starts))):
            issue_copy(i)
        for i, c0 in enumerate(starts):
            b = i % self.nbuf
            n = sizes[i]
            main.wait_event(self.in_ready[b])
            yb = self.inbuf[b][:n]
            if self.pcm_rate is not None:
                yb = core.resample_poly(yb, self.pcm_rate, int(self.sr), n_channels=self.pcm_channels)
            feat = core.stft_features(yb, sr=self.sr, hop_length=self.hop, want_mag=True, want_mel=True, want_rms=True,
                                      mag_out=self.mag[:n])
            post = core.mel_post(feat["mel"], feat["mel_max"], sr=self.sr, hop_length=self.hop, want_sdb=False,
                                 want_rake=False, want_onset=True)
            pk = core.onset_peaks(post["onset_env"], post["env_minmax"], sr=self.sr, hop_length=self.hop)
            self.in_free[b].record(main)
            if i + ahead < len(starts):
                issue_copy(i + ahead)
            self.rms[c0 : c0 + n].copy_(feat["rms"], non_blocking=True)
            self.env[c0 : c0 + n].copy_(post["onset_env"], non_blocking=True)
            self.peaks[c0 : c0 + n].copy_(pk["peaks"], non_blocking=True)
        main.synchronize()
        return {"rms": self.rms, "onset_env": self.env, "onset_peaks": self.peaks}


class TranscribePipeline:
    """Host-buffer plugin call for full transcription of a batch: 16-bit PCM (or float32) host audio in, the reference's
    perception arrays and the note-event records out.

    The host batch travels in PIECES of ``chunk_clips`` clips (default: one clip per SM; one cudaMemcpyAsync each on ONE
    copy stream: with two streams the copy engine drained one stream's pieces before the other's, and piece 1 landed after
    piece 6 -- `tools/pipeline_trace.py`) into a device buffer for the whole batch.  The frame-parallel kernels
    run on every piece as soon as it has landed, while later pieces are still on the bus: ingest (K9, when the host batch
    is PCM), K1 + K4 (mel dB, rake mask, RMS) and K2 (pYIN observations, written into batch-wide buffers).  The decoder
    is sequential over frames with one CTA per clip, four resident per SM, so K3 (Viterbi) and K7 (note events) run on
    GROUPS of ``group_clips`` clips (default: everything in one launch, or wave by wave when the first run shows the step is
    copy-bound), then the group's
    ``rake_mask, f0, voiced_flag, voiced_probs, rms`` (the dict of aegis_engine.py:72-75) and its event records go to
    pinned host arrays.  Results equal ``analyze_batch`` + ``note_events_batch`` on the whole batch bit for bit.
    """

    def __init__(self, n_clips: int, n_samples: int, *, sr: float, hop_length: int = 512, device=None, chunk_clips: Optional[int] = None,
                 pcm: bool = True, copy_streams: int = 1, confidence_threshold: float = 0.7, group_clips: Optional[int] = None,
                 fmin: float = E2, fmax: float = C6, rake_sensitivity: float = 0.6):
        from . import _native as nat

        self.sr, self.hop, self.thr, self.ratio = sr, hop_length, confidence_threshold, rake_sensitivity
        self.fmin, self.fmax = fmin, fmax
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.n_clips, self.n_samples, self.pcm = n_clips, n_samples, pcm
        with torch.cuda.device(self.dev):
            n_sm = int(nat.load().aegis_device_sm_count())
        self.chunk = max(1, min(n_sm if chunk_clips is None else chunk_clips, n_clips))
        # a group = the clips of one decoder launch.  Full waves of Viterbi chains (4 per SM) use the SMs best, and one launch
        # for everything (two waves for 1024 clips) is fastest when the step is COMPUTE-bound: 25.9 ms against 14.9 + 13.4 ms
        # for 592 + 432 clips (tools/pipeline_trace.py).  When the step is COPY-bound (several ranks behind one host link: a
        # piece lands every 7 ms and keeps the SMs busy for 3) the first wave's decode fits into the idle time under the
        # remaining copies and only the last group's decode is left after the last piece.  group_clips=None decides from
        # the first run's event times (copy per piece against kernels per piece) and keeps the choice.
        self.n_sm = n_sm
        self.auto_group = group_clips is None
        self.group = max(self.chunk, min(8 * n_sm if group_clips is None else group_clips, n_clips))
        self.T = T = core.frame_count(n_samples, hop_length)
        self.cfg = tables.pyin_config(float(sr), int(hop_length), float(fmin), float(fmax))
        mc = self.cfg.max_troughs
        in_dtype = torch.int16 if pcm else torch.float32
        self.in_shape = (n_clips, n_samples)
        dev = self.dev
        self.inbuf = torch.empty(self.in_shape, dtype=in_dtype, device=dev)
        self.copy_streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, copy_streams))]
        self.pieces = list(range(0, n_clips, self.chunk))
        self.landed = [torch.cuda.Event(enable_timing=True) for _ in self.pieces]
        # batch-wide device buffers filled piece by piece
        self.obs = {"cand_bin": torch.empty((n_clips * T, mc), dtype=torch.int16, device=dev),
                    "cand_prob": torch.empty((n_clips * T, mc), dtype=torch.float64, device=dev),
                    "cand_count": torch.empty((n_clips * T,), dtype=torch.int32, device=dev),
                    "voiced_prob": torch.empty((n_clips, T), dtype=torch.float64, device=dev)}
        self.rake = torch.empty((n_clips, T), dtype=torch.uint8, device=dev)
        self.rms = torch.empty((n_clips, T), dtype=torch.float32, device=dev)
        min_frames, _ = core.note_frame_limits(sr, hop_length)
        self.max_events = T // (min_frames + 1) + 1
        pin = dict(pin_memory=True)
        self.out = {
            "rake_mask": torch.empty((n_clips, T), dtype=torch.uint8, **pin),
            "f0": torch.empty((n_clips, T), dtype=torch.float64, **pin),
            "voiced_flag": torch.empty((n_clips, T), dtype=torch.uint8, **pin),
            "voiced_probs": torch.empty((n_clips, T), dtype=torch.float64, **pin),
            "rms": torch.empty((n_clips, T), dtype=torch.float32, **pin),
            "events": torch.empty((n_clips, self.max_events, core.NOTE_EVENT_DTYPE.itemsize), dtype=torch.uint8, **pin),
            "n_events": torch.empty((n_clips,), dtype=torch.int32, **pin),
        }
        self.h2d_bytes = n_clips * n_samples * (2 if pcm else 4)
        self.d2h_bytes = sum(int(t.numel()) * t.element_size() for t in self.out.values())

    def _piece(self, c0: int, c1: int):
        """frame-parallel kernels of clips [c0, c1): K9, K1, K4, K2"""
        T = self.T
        yb = self.inbuf[c0:c1]
        if self.pcm:
            yb = core.resample_poly(yb, int(self.sr), int(self.sr))   # int16 -> float32 / 32768 on the device (K9)
        spec = spectral_features(yb, sr=self.sr, hop_length=self.hop, rake_sensitivity=self.ratio, with_onsets=False)
        self.rake[c0:c1].copy_(spec["rake_mask"])
        self.rms[c0:c1].copy_(spec["rms"])
        sl = slice(c0 * T, c1 * T)
        core.yin_candidates(yb, self.cfg, out={"cand_bin": self.obs["cand_bin"][sl], "cand_prob": self.obs["cand_prob"][sl],
                                               "cand_count": self.obs["cand_count"][sl], "voiced_prob": self.obs["voiced_prob"][c0:c1]})

    def _group(self, g0: int, g1: int):
        """decoder + logic filter of clips [g0, g1): K3, K7, then the results to the host"""
        T = self.T
        sl = slice(g0 * T, g1 * T)
        obs = {"cand_bin": self.obs["cand_bin"][sl], "cand_prob": self.obs["cand_prob"][sl], "cand_count": self.obs["cand_count"][sl],
               "voiced_prob": self.obs["voiced_prob"][g0:g1], "n_frames": T, "max_cand": self.cfg.max_troughs}
        dec = core.viterbi_decode(obs, self.cfg, g1 - g0)
        res = {"rake_mask": self.rake[g0:g1], "voiced_flag": dec["voiced_flag"], "voiced_probs": self.obs["voiced_prob"][g0:g1],
               "rms": self.rms[g0:g1], "f0": torch.nan_to_num(dec["f0"], nan=0.0), "states": dec["states"]}
        ev = note_events_batch(res, sr=self.sr, hop_length=self.hop, fmin=self.fmin, fmax=self.fmax, confidence_threshold=self.thr,
                               max_events=self.max_events)
        for k in ("rake_mask", "f0", "voiced_flag", "voiced_probs", "rms"):
            self.out[k][g0:g1].copy_(res[k], non_blocking=True)
        self.out["events"][g0:g1].copy_(ev["events"], non_blocking=True)
        self.out["n_events"][g0:g1].copy_(ev["n_events"], non_blocking=True)

    def run(self, y_host: torch.Tensor) -> dict:
        if y_host.is_cuda or tuple(y_host.shape) != self.in_shape or y_host.dtype != self.inbuf.dtype:
            raise ValueError(f"expected a host {self.inbuf.dtype} tensor {self.in_shape}")
        main = torch.cuda.current_stream(self.dev)
        trace = getattr(self, "trace", None)     # debugging aid: set pipe.trace = [] to get (label, ms since start) marks
        start = torch.cuda.Event(enable_timing=trace is not None)
        start.record(main)           # the previous run's kernels are done with the input buffer
        marks = []

        def mark(label):
            if trace is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(main)
                marks.append((label, e))

        for i, c0 in enumerate(self.pieces):
            cs = self.copy_streams[i % len(self.copy_streams)]
            with torch.cuda.stream(cs):
                if i < len(self.copy_streams):
                    cs.wait_event(start)
                self.inbuf[c0 : c0 + self.chunk].copy_(y_host[c0 : c0 + self.chunk], non_blocking=True)
                self.landed[i].record(cs)
        g0 = 0
        probe = None
        for i, c0 in enumerate(self.pieces):
            c1 = min(self.n_clips, c0 + self.chunk)
            main.wait_event(self.landed[i])
            mark(f"piece {i} landed")
            if self.auto_group and i == len(self.pieces) // 2:   # time one piece's kernels (first run only)
                probe = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                probe[0].record(main)
            self._piece(c0, c1)
            if probe is not None and i == len(self.pieces) // 2:
                probe[1].record(main)
            mark(f"piece {i} K9+K1+K4+K2 done")
            if c1 - g0 >= self.group or c1 == self.n_clips:
                self._group(g0, c1)
                mark(f"group [{g0},{c1}) K3+K7+D2H issued/done")
                g0 = c1
        main.synchronize()
        if self.auto_group:
            self.auto_group = False
            if probe is not None and len(self.pieces) >= 4 and self.n_clips > 4 * self.n_sm:
                k = len(self.pieces) // 2
                copy_ms = self.landed[k].elapsed_time(self.landed[k + 1])       # one piece on the bus
                kern_ms = probe[0].elapsed_time(probe[1])                       # one piece through K9 + K1 + K4 + K2
                if copy_ms > 1.5 * kern_ms:
                    self.group = max(self.chunk, (4 * self.n_sm // self.chunk) * self.chunk)   # copy-bound: decode wave by wave
                self.group_decision = {"copy_ms_per_piece": copy_ms, "kernel_ms_per_piece": kern_ms, "group_clips": self.group}
        if trace is not None:
            trace[:] = [(label, start.elapsed_time(e)) for label, e in marks]
        return self.out
