#!/usr/bin/env python
"""Benchmark of the Aegis hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at every N (weak scaling, sharded by clip, no data-path collective): BASELINE.json
configs[1] -- a batch of 1,024 synthetic 30 s clips at 22 050 Hz per GPU; one step = one pass of
STFT |X| + mel dB onset strength + onset peak picking + frame RMS over the batch.
`value`   : audio-seconds analysed per second with the batch resident in HBM (CUDA events).
`e2e`     : the same through the host-buffer plugin call: pinned 16-bit PCM host buffers (the payload of a WAV file) ->
            H2D -> K9 ingest -> kernels -> D2H of RMS / onset envelope / onset flags, copies inside the timed region;
            beside it the copy-only time of the same bytes (the ceiling of any host-fed path) and `e2e_f32` (float32 buffers).
`roofline`: the STFT kernel (dominant), algorithmic bytes 4*N + 4*1025*T per clip over its own
            CUDA-event time, against MEASURED_PEAKS.json's HBM copy bandwidth; `traffic` from the committed ncu capture,
            null when the kernel sources changed since it was taken.
`transcribe`: BASELINE cfg3 + the v1 logic filter on the same clips -- mel dB + rake mask + pYIN + RMS + note events:
            device-resident value, per-kernel times, end to end from PCM host buffers, its own CPU baseline, and the
            issue-slot utilisation of the Viterbi kernel (latency bound: no HBM roofline applies).
`cpu_baseline` / `--impl reference`: the CPU oracle (port of the reference's librosa path; librosa
            itself is not installable here) on the box's host cores: one single-threaded worker process per core, by
            clip, on a bounded sample of the same clips.
`aux.long_clip` (N > 1): BASELINE cfg4, one hour at 44.1 kHz over the N ranks with the NCCL event gather, compared by hash
            with one rank alone.
Prints exactly ONE JSON line on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 22050
CLIP_SECONDS = 30.0
N_CLIPS = int(os.environ.get("AEGIS_BENCH_CLIPS", "1024"))
HOP = 512
METRIC = "audio-sec transcribed/sec (realtime factor)"
UNIT = "audio-s/s"
WORKLOAD = "cfg2: 1024 x 30 s clips @22050 Hz per GPU, STFT |X| + onset strength/peaks + RMS (n_fft 2048, hop 512)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference path, one clip per task, all host cores
# ------------------------------------------------------------------------------------------------
def _cpu_spectral_one(y):
    """STFT |X| + onset envelope/peaks + RMS of one clip with the CPU oracle (same outputs as one GPU step)."""
    import numpy as np

    from oracle import librosa_ref as L

    S = L.stft_magnitude(y)
    mel = np.einsum("ft,mf->mt", S ** 2, L.mel_filterbank(SR), optimize=True)
    env = L.onset_strength(S=L.power_to_db(mel), sr=SR)
    peaks = L.onset_detect(onset_envelope=env, sr=SR)
    r = L.rms(y)
    return float(S[3, 3]) + float(env.sum()) + len(peaks) + float(r.sum())


def _cpu_transcribe_one(y):
    """The reference's v1 pipeline for one clip with the CPU oracle: mel dB + rake mask + pYIN E2..C6 + RMS +
    get_midi_events (aegis_engine.py:41-96)."""
    import warnings

    import numpy as np

    from oracle import librosa_ref as L
    from oracle import reference_files as R

    S_dB = L.load_audio_features(y, SR)
    mask = R.detect_rake_patterns(S_dB, HOP, SR, 0.6)
    f0, vf, vp = L.pyin(y, fmin=L.note_to_hz("E2"), fmax=L.note_to_hz("C6"), sr=SR, hop_length=HOP)
    rms = L.rms(y)[0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ev = R.get_midi_events(mask, np.nan_to_num(f0), vf, vp, rms, SR, HOP, 0.70)
    return len(ev)


_THREAD_VARS = ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS")


def _pool_init():
    """Worker start-up.  The workers are SPAWNED with the thread variables already set to 1 in the environment they
    inherit (`single_threaded_env`), so their BLAS / OpenMP runtimes come up single-threaded: one process per core,
    one thread per process, as the reference's multiprocessing pool runs (aegis_engine.py:207-210).  (Round 1 set the
    variables here, after a fork: the parent's BLAS was already loaded with all its threads and the pool thrashed.)
    threadpool_limits is the second line of defence for a runtime that ignores the environment."""
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=1)
    except Exception:  # pragma: no cover
        pass


class single_threaded_env:
    """Set the BLAS / OpenMP thread variables to 1 in os.environ while worker processes are started."""

    def __enter__(self):
        self.saved = {k: os.environ.get(k) for k in _THREAD_VARS}
        for k in _THREAD_VARS:
            os.environ[k] = "1"

    def __exit__(self, *exc):
        for k, v in self.saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


class CpuPool:
    """multiprocessing.Pool(cores) of single-threaded workers, started with `spawn` (no fork of a process that holds a
    CUDA context or a many-threaded BLAS)."""

    def __init__(self, cores, fn=None):
        import multiprocessing as mp

        self.cores, self.fn = cores, fn or _cpu_spectral_one
        with single_threaded_env():
            self.pool = mp.get_context("spawn").Pool(cores, initializer=_pool_init)
            self.pool.map(_worker_threads, range(cores))    # every worker is up (and its BLAS loaded) before the env is restored

    def throughput(self, clips, repeats=1, warm=True):
        """(audio-s/s, wall seconds) of the CPU oracle over `clips` (numpy [n, N]), one clip per task."""
        if warm:
            self.pool.map(self.fn, [clips[i] for i in range(min(len(clips), self.cores))])  # caches / imports / FFT plans
        t0 = time.perf_counter()
        for _ in range(repeats):
            self.pool.map(self.fn, [clips[i] for i in range(len(clips))], chunksize=1)
        dt = time.perf_counter() - t0
        return repeats * clips.shape[0] * clips.shape[1] / SR / dt, dt

    def worker_threads(self):
        return max(self.pool.map(_worker_threads, range(self.cores)))

    def close(self):
        self.pool.close()
        self.pool.join()


def _worker_threads(_):
    """largest BLAS / OpenMP thread count of this worker (1 when the pool is set up correctly)"""
    import numpy  # noqa: F401  (loads the BLAS)
    try:
        from threadpoolctl import threadpool_info

        return max([int(d.get("num_threads", 1)) for d in threadpool_info()] or [1])
    except Exception:  # pragma: no cover
        return -1


def serial_throughput(clips, fn=None):
    """audio-s/s of the CPU oracle on ONE core, in this process (SURVEY.md §8d asks for serial next to the pool)."""
    from threadpoolctl import threadpool_limits

    fn = fn or _cpu_spectral_one
    with threadpool_limits(limits=1):   # BLAS is already loaded in this process: the environment variables come too late
        fn(clips[0])                    # warm caches / FFT plans
        t0 = time.perf_counter()
        for c in clips:
            fn(c)
        return clips.shape[0] * clips.shape[1] / SR / (time.perf_counter() - t0)


def library_versions():
    import platform

    import numpy
    import scipy

    return {"python": platform.python_version(), "numpy": numpy.__version__, "scipy": scipy.__version__, "librosa": None}


def pooled_rate(pool, clips):
    """clips per second of the CPU oracle with the full pool busy (sizes the bounded sample)."""
    k = min(len(clips), 2 * pool.cores)
    _, dt = pool.throughput(clips[:k])
    return k / dt


def cpu_baseline_record(pool, sample, what, serial_clips=3):
    """The `cpu_baseline` object: pooled throughput on `sample`, the one-core figure beside it, and a self-check that the
    pool really used its cores (round 1's pool ran many-threaded BLAS in every worker and was ~10x slow)."""
    v, dt = pool.throughput(sample)
    serial = serial_throughput(sample[:serial_clips], pool.fn)
    rec = {"value": v, "unit": UNIT, "cores": pool.cores, "kind": "port",
           "sample": f"{what}; multiprocessing.Pool({pool.cores}) of single-threaded spawned workers, one clip per task, {dt:.1f} s wall; "
                     "numpy/scipy oracle port of the librosa path (librosa is not installable here)",
           "serial_value": serial, "serial_sample": f"{serial_clips} clips on one core, same process",
           "worker_blas_threads": pool.worker_threads(), "versions": library_versions()}
    if v < 0.5 * pool.cores * serial:
        # expected on SMT hosts (os.cpu_count() counts hyper-threads) and when the sample is shorter than a few tasks per
        # worker; a pool that thrashes shows up as a ratio far below this
        rec["cpu_baseline_suspect"] = True
    rec["pool_efficiency"] = v / (pool.cores * serial)
    return rec


def host_sample_clips(n, seed0=0):
    import spectrogram_midi_b200 as P

    return P.corpus.clip_batch(n, CLIP_SECONDS, SR, first_seed=seed0, workers=min(n, os.cpu_count() or 1))


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path.  librosa cannot be installed
    here (no network, not in the wheelhouse), so this times the oracle port with every host core: one single-threaded
    worker process per core, sharded by clip (the generous reading of Turbo Mode, aegis_engine.py:183-216)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    import numpy as np  # noqa: F401

    t_probe0 = time.perf_counter()
    pool = CpuPool(cores)
    probe = host_sample_clips(2 * max(cores, 4))
    rate = pooled_rate(pool, probe)
    # bounded sample: ~8 s of wall per step on all cores and the whole run within ~2.5 minutes
    per_step_s = min(8.0, 150.0 / max(1, args.steps))
    n = int(min(1024, max(cores, round(rate * per_step_s))))
    clips = host_sample_clips(n) if n > len(probe) else probe[:n]
    log(f"[reference] cores={cores} worker BLAS threads={pool.worker_threads()} pooled rate={rate:.1f} clips/s "
        f"sample={n} clips per step (setup {time.perf_counter() - t_probe0:.1f}s)")
    for _ in range(args.warmup):
        pool.throughput(clips[: 2 * cores], warm=False)
    t_total = 0.0
    for _ in range(args.steps):
        _, dt = pool.throughput(clips, warm=False)
        t_total += dt
    value = args.steps * n * CLIP_SECONDS / t_total
    serial = serial_throughput(clips[:3])
    pool.close()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD, "sample_clips_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} clips x 30 s per step, multiprocessing.Pool({cores}) of single-threaded workers by clip, "
                                   "numpy/scipy oracle port",
                         "serial_value": serial, "pool_efficiency": value / (cores * serial)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if value < 0.5 * cores * serial:
        line["cpu_baseline"]["cpu_baseline_suspect"] = True
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(dev_index):
    """Pin this rank to the CPUs next to its GPU (NVML's ideal affinity) BEFORE any pinned host buffer is allocated, so
    the buffers are first-touched on the GPU's NUMA node: with eight ranks copying 2.7 GB per step each, buffers that
    all sit on one socket make the end-to-end number a cross-socket memory benchmark.  Returns the CPU list or None."""
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(dev_index)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        before = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = os.sched_getaffinity(0)
        if not after:
            os.sched_setaffinity(0, before)
            return None
        return sorted(after)
    except Exception:  # pragma: no cover - NVML missing, cpuset without the GPU's CPUs, ...
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _hash_sources(names):
    import hashlib

    h = hashlib.sha256()
    for n in names:
        with open(os.path.join(ROOT, "spectrogram-midi_b200", "csrc", n), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def committed_capture(path, sources):
    """A number that can only come from a profiler (DRAM bytes, issue utilisation) is read from the committed summary of
    the ncu capture -- but only when that capture was taken from the kernel sources as they are now (the summary
    records their hash); otherwise None and the reason."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", path)))
    except (OSError, ValueError):
        return None, f"profiles/{path} missing"
    if rec.get("source_sha16") != _hash_sources(sources):
        return None, f"profiles/{path} is stale: kernel sources changed since the capture"
    return rec, None


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import spectrogram_midi_b200 as P
    from spectrogram_midi_b200 import _native, batch, core

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the Aegis B200 path has no CPU fallback; use --impl reference for the CPU arm)")
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    _native.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_cpus = bind_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_samples = int(CLIP_SECONDS * SR)
    T = 1 + n_samples // HOP
    # ---- synthetic corpus, rendered on the device (seed = global clip index)
    t0 = time.perf_counter()
    plan = P.corpus.plan_events(N_CLIPS, CLIP_SECONDS, SR, first_seed=rank * N_CLIPS)
    y = core.synth_events(N_CLIPS, n_samples, plan, dev)
    torch.cuda.synchronize()
    log(f"[rank {rank}] corpus: {N_CLIPS} clips x {CLIP_SECONDS:.0f} s rendered in {time.perf_counter() - t0:.1f} s")

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def wall(fn, reps):
        barrier()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        barrier()
        return (time.perf_counter() - t) / reps

    audio_s_per_step = world * N_CLIPS * CLIP_SECONDS

    # ================================================================================================================
    # headline (BASELINE cfg2): STFT |X| + onset strength / peaks + RMS, batch resident in HBM
    # ================================================================================================================
    mag = core.alloc_frames((N_CLIPS, 1025), T, dev)  # rows on 32-byte sector boundaries
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stft_ms = []

    def step(timed=False):
        # the same calls batch.spectral_features makes, with CUDA events around the STFT launch
        if timed:
            ev_a.record()
        feat = core.stft_features(y, sr=SR, hop_length=HOP, want_mag=True, want_mel=True, want_rms=True, mag_out=mag)
        if timed:
            ev_b.record()
        post = core.mel_post(feat["mel"], feat["mel_max"], sr=SR, hop_length=HOP, want_sdb=False, want_rake=False, want_onset=True)
        pk = core.onset_peaks(post["onset_env"], post["env_minmax"], sr=SR, hop_length=HOP)
        return feat, post, pk

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # Python's cyclic collector stays off while anything is timed (as timeit does): with the corpus plan's ~10^5 small objects
    # alive, a full collection landing inside a step cost tens of milliseconds (seen as one v2 step in five at 2x)
    import gc

    gc.collect()
    gc.disable()
    sampler = ClockSampler(local_rank)
    sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    start.record()
    for _ in range(args.steps):
        step(timed=True)
        launches += 4  # stft_fused, mel_post (onset_flux4), peak_candidates, peak_select
        ev_b.synchronize()
        stft_ms.append(ev_a.elapsed_time(ev_b))
    stop.record()
    barrier()
    elapsed_ms = start.elapsed_time(stop)

    # ================================================================================================================
    # end to end through the host-buffer plugin call (cfg2).  PRIMARY: 16-bit PCM host buffers -- what a WAV file
    # holds; the ingest kernel K9 scales to float on the device.  SECONDARY: float32 host buffers.
    # ================================================================================================================
    e2e_steps = max(2, min(args.steps, 4))
    pcm_host = torch.empty((N_CLIPS, n_samples), dtype=torch.int16, pin_memory=True)
    pcm_host.copy_((y * 32767.0).round().to(torch.int16))
    pipe16 = batch.HostPipeline(N_CLIPS, n_samples, sr=SR, hop_length=HOP, device=dev, chunk_clips=128, pcm_rate=SR)
    for _ in range(2):
        pipe16.run(pcm_host)
    pcm_s = wall(lambda: pipe16.run(pcm_host), e2e_steps)
    pcm_bytes = (int(pipe16.h2d_bytes), int(pipe16.d2h_bytes))
    # the ceiling of that path: the same chunks copied host -> device and nothing else (one cudaMemcpyAsync per chunk)
    sink = torch.empty((128, n_samples), dtype=torch.int16, device=dev)

    def copy_only(src, buf):
        for c0 in range(0, N_CLIPS, 128):
            buf[: min(128, N_CLIPS - c0)].copy_(src[c0 : c0 + 128], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    copy_only(pcm_host, sink)
    h2d16_s = wall(lambda: copy_only(pcm_host, sink), e2e_steps)
    del pipe16, sink

    y_host = torch.empty((N_CLIPS, n_samples), dtype=torch.float32, pin_memory=True)
    y_host.copy_(y)
    pipe = batch.HostPipeline(N_CLIPS, n_samples, sr=SR, hop_length=HOP, device=dev, chunk_clips=128)
    for _ in range(2):
        pipe.run(y_host)
    e2e_s = wall(lambda: pipe.run(y_host), e2e_steps)
    f32_bytes = (int(pipe.h2d_bytes), int(pipe.d2h_bytes))
    sink32 = torch.empty((128, n_samples), dtype=torch.float32, device=dev)
    copy_only(y_host, sink32)
    h2d32_s = wall(lambda: copy_only(y_host, sink32), e2e_steps)
    del pipe, y_host, sink32

    # ================================================================================================================
    # transcription (BASELINE cfg3 + the v1 logic filter): spectral + rake mask + pYIN E2..C6 + RMS + note events
    # on the SAME 1024 clips per GPU: device-resident, per-kernel times, and end to end from PCM host buffers
    # ================================================================================================================
    tr = {}
    if not args.no_pyin:
        cfg = P.tables.pyin_config(float(SR), HOP, batch.E2, batch.C6)

        def transcribe():
            res = batch.analyze_batch(y, sr=SR, hop_length=HOP)
            evs = batch.note_events_batch(res, sr=SR, hop_length=HOP)
            return res, evs

        for _ in range(2):
            transcribe()
        tr_steps = max(2, min(args.steps, 3))
        barrier()
        a = ev()
        for _ in range(tr_steps):
            res, evs = transcribe()
        b = ev()
        barrier()
        tr_ms = a.elapsed_time(b) / tr_steps
        n_events_mean = float(evs["n_events"].float().mean())
        del res, evs
        # per-kernel times of one step (CUDA events around each call; same calls analyze_batch makes)
        marks = [ev()]
        feat = core.stft_features(y, sr=SR, hop_length=HOP, want_mag=False, want_mel=True, want_rms=True); marks.append(ev())
        post = core.mel_post(feat["mel"], feat["mel_max"], sr=SR, hop_length=HOP, want_sdb=False, want_rake=True, want_onset=False); marks.append(ev())
        obs = core.yin_candidates(y, cfg); marks.append(ev())
        dec = core.viterbi_decode(obs, cfg, N_CLIPS); marks.append(ev())
        res = {"rake_mask": post["rake_mask"], "f0": dec["f0"], "voiced_flag": dec["voiced_flag"], "voiced_probs": obs["voiced_prob"],
               "rms": feat["rms"], "states": dec["states"]}
        evs = batch.note_events_batch(res, sr=SR, hop_length=HOP); marks.append(ev())
        torch.cuda.synchronize()
        names = ["K1 stft_fused (mel + rms, no |X| store)", "K4 mel_post (dB + rake mask)", "K2 yin", "K3 viterbi (forward + backtrace)", "K7 note events"]
        stage_ms = {n: marks[i].elapsed_time(marks[i + 1]) for i, n in enumerate(names)}
        del feat, post, obs, dec, res, evs
        # the v2 ("financial") pipeline on the same clips, BASELINE cfg5's shape of work (aegis_engine_financial.py:73-171):
        # perception + dB image + guitar filters (K6) + consensus trend (K5) + the financial logic filter (K5 x2 + K8)

        def transcribe_v2():
            r2 = batch.analyze_batch(y, sr=SR, hop_length=HOP, with_trend=True, with_guitar=True, nan_to_num=False)
            return batch.note_events_financial_batch(r2, sr=SR, hop_length=HOP)

        v2_ms = None
        try:
            for _ in range(3):
                transcribe_v2()
            barrier()
            # every step is timed on its own and the MEDIAN is reported, with the worst step beside it (a full Python garbage
            # collection inside one step used to double it; the collector is now off while anything is timed)
            per_step = []
            for _ in range(tr_steps + 2):
                a = ev()
                e2 = transcribe_v2()
                b = ev()
                torch.cuda.synchronize()
                per_step.append(a.elapsed_time(b))
            barrier()
            per_step.sort()
            v2_ms = per_step[len(per_step) // 2]
            v2_ms_worst = per_step[-1]
            v2_events = float(e2["n_events"].float().mean())
            del e2
        except Exception as e:  # pragma: no cover
            log("transcribe_v2 failed:", e)
        # end to end: PCM host buffers in, perception arrays + event records out
        tp = batch.TranscribePipeline(N_CLIPS, n_samples, sr=SR, hop_length=HOP, device=dev, chunk_clips=128, pcm=True)
        for _ in range(2):
            tp.run(pcm_host)
        tr_e2e_s = wall(lambda: tp.run(pcm_host), tr_steps)
        tr_bytes = (int(tp.h2d_bytes), int(tp.d2h_bytes))
        del tp
        vit, why = committed_capture("r2_viterbi_issue.json", ["viterbi.cu"])
        tr = {"ms": tr_ms, "e2e_s": tr_e2e_s, "stage_ms": stage_ms, "bytes": tr_bytes, "n_events_mean": n_events_mean,
              "issue": vit, "issue_note": why, "v2_ms": v2_ms, "v2_events": v2_events if v2_ms is not None else None,
              "v2_ms_worst": v2_ms_worst if v2_ms is not None else None}
    del pcm_host
    clocks = sampler.stop()   # sampled every 200 ms from the headline loop to the end of the transcription section
    gc.enable()

    # ---- long clip (BASELINE cfg4) at N > 1: one hour at 44.1 kHz over the ranks, exact and windowed, against one rank alone
    long_clip = None
    if world > 1 and not args.no_long_clip:
        long_clip = run_long_clip(args, P, core, dev, rank, world)

    vals = [elapsed_ms, e2e_s, pcm_s, h2d16_s, h2d32_s, tr.get("ms", 0.0), tr.get("e2e_s", 0.0), tr.get("v2_ms") or 0.0]
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_s, pcm_s, h2d16_s, h2d32_s, tr_ms_max, tr_e2e_max, v2_ms_max = [float(v) for v in t]
    value = audio_s_per_step * args.steps / (elapsed_ms / 1e3)

    # ---- roofline of the STFT kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = N_CLIPS * (4 * n_samples + 4 * 1025 * T)
    stft_avg_ms = float(np.mean(stft_ms))
    achieved = alg_bytes / (stft_avg_ms / 1e3) / 1e9
    # DRAM traffic of that kernel per launch: dram__bytes_read.sum + dram__bytes_write.sum of the committed `ncu --set full`
    # capture of the same launch shape -- valid only for the kernel sources it was taken from (hash checked)
    traffic, traffic_note = None, None
    cap, why = committed_capture("r2_stft_traffic.json", ["stft_fused.cu", "rfft2048x2.cuh"])
    if cap is not None and N_CLIPS == 1024:
        traffic = float(cap["dram_bytes_read"]) + float(cap["dram_bytes_write"])
    else:
        traffic_note = why or "capture is for 1024 clips per launch"

    cpu = cpu_tr = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        pool = CpuPool(cores)                        # spawned, single-threaded workers (no fork of this CUDA process)
        rate = pooled_rate(pool, y[: 2 * cores].cpu().numpy())
        n = int(min(N_CLIPS, max(cores, round(12.0 * rate))))   # ~12 s of wall with every core busy
        cpu = cpu_baseline_record(pool, y[:n].cpu().numpy(), f"first {n} clips of the batch ({n * CLIP_SECONDS:.0f} audio-s)")
        pool.close()
        if tr:
            pool = CpuPool(cores, _cpu_transcribe_one)
            probe = y[:cores].cpu().numpy()
            _, dt = pool.throughput(probe)           # warm (numba JIT in every worker), then one clip per worker
            n = int(min(N_CLIPS, max(cores, round(15.0 * cores / dt))))
            cpu_tr = cpu_baseline_record(pool, y[:n].cpu().numpy(), f"first {n} clips of the batch ({n * CLIP_SECONDS:.0f} audio-s), "
                                         "spectral + rake mask + pYIN + RMS + get_midi_events", serial_clips=1)
            pool.close()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "clips_per_gpu": N_CLIPS, "clip_seconds": CLIP_SECONDS, "sr": SR, "frames_per_clip": T, "sharding": "by clip",
                       "host_affinity": None if host_cpus is None else f"rank 0 pinned to the {len(host_cpus)} CPUs next to its GPU (NVML)",
                       "l2": "inputs (2.7 GB) and outputs (5.4 GB) per step exceed the 126 MB L2; no flush needed"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_note": traffic_note, "kernel": "stft_fused_kernel", "algorithmic_bytes_per_launch": alg_bytes,
                         "ms_per_launch": stft_avg_ms, "peak_source": peak_src, "share_of_step": stft_avg_ms / (elapsed_ms / args.steps)},
            "cpu_baseline": cpu,
            "e2e": {"value": audio_s_per_step / pcm_s, "unit": UNIT, "h2d_bytes_per_step": pcm_bytes[0], "d2h_bytes_per_step": pcm_bytes[1],
                    "ms_per_step": pcm_s * 1e3,
                    "note": "PRIMARY end-to-end: the clips as 16-bit PCM host buffers (the payload of a WAV file), pinned; chunked H2D on "
                            "two copy streams overlapped with the kernels; K9 scales to float on the device; D2H of rms + onset envelope "
                            "+ onset flags (|X| stays in HBM)",
                    "h2d_copy_only_ms": h2d16_s * 1e3, "h2d_copy_only_gbs": world * pcm_bytes[0] / h2d16_s / 1e9,
                    "fraction_of_copy_ceiling": h2d16_s / pcm_s},
            "e2e_f32": {"value": audio_s_per_step / e2e_s, "unit": UNIT, "h2d_bytes_per_step": f32_bytes[0], "d2h_bytes_per_step": f32_bytes[1],
                        "ms_per_step": e2e_s * 1e3, "note": "secondary: the same call with float32 host buffers (twice the PCIe bytes)",
                        "h2d_copy_only_ms": h2d32_s * 1e3, "h2d_copy_only_gbs": world * f32_bytes[0] / h2d32_s / 1e9,
                        "fraction_of_copy_ceiling": h2d32_s / e2e_s},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if tr:
            st = tr["stage_ms"]
            tot = sum(st.values())
            issue = tr["issue"]
            line["transcribe"] = {
                "workload": "cfg3 + v1 logic filter on the same 1024 x 30 s clips per GPU: mel dB + rake mask + pYIN E2..C6 "
                            "(441 bins / 882 states) + RMS + get_midi_events (aegis_engine.py:41-96)",
                "value": audio_s_per_step / (tr_ms_max / 1e3), "unit": UNIT, "ms_per_step": tr_ms_max,
                "kernel_ms": st, "kernel_share": {k: v / tot for k, v in st.items()},
                "note_events_per_clip": tr["n_events_mean"],
                "v2_financial": None if not v2_ms_max else {
                    "workload": "the v2 engine's pipeline on the same clips (BASELINE cfg5's shape of work): perception + dB image + "
                                "guitar filters (K6) + consensus trend (K5) + financial logic filter (K5 x2 + K8)",
                    "value": audio_s_per_step / (v2_ms_max / 1e3), "unit": UNIT, "ms_per_step": v2_ms_max,
                    "timing": "median of individually timed steps", "ms_per_step_worst": tr.get("v2_ms_worst"),
                    "note_events_per_clip": tr["v2_events"]},
                "e2e": {"value": audio_s_per_step / tr_e2e_max, "unit": UNIT, "ms_per_step": tr_e2e_max * 1e3,
                        "h2d_bytes_per_step": tr["bytes"][0], "d2h_bytes_per_step": tr["bytes"][1],
                        "note": "16-bit PCM host buffers in; rake_mask, f0, voiced_flag, voiced_probs, rms and the note-event records out"},
                "roofline": {"bound": "issue", "kernel": "viterbi_forward_kernel",
                             "achieved": None if issue is None else issue["smsp_issue_active_pct"], "peak": 100.0, "unit": "% issue slots",
                             "frac": None if issue is None else issue["smsp_issue_active_pct"] / 100.0,
                             "source": "committed ncu capture profiles/r2_viterbi_issue.json (a profiler metric cannot be read inside a "
                                       "timed run)" if issue is not None else tr["issue_note"],
                             "ms_per_launch": st["K3 viterbi (forward + backtrace)"]},
                "cpu_baseline": cpu_tr,
            }
        if long_clip is not None:
            line["aux"] = {"long_clip": long_clip}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_long_clip(args, P, core, dev, rank, world):
    """BASELINE cfg4: ONE 3600 s recording at 44.1 kHz analysed by all ranks in overlapping time windows
    (distributed.analyze_long_clip): frame-local kernels per window, all_reduce(MAX) of the mel maximum, all-gather of the
    sparse observations (exact) or of the decoded frames (windowed), and the note-event gather -- all NCCL.  Rank 0 then
    analyses the clip alone and the results are compared by hash."""
    import hashlib

    import numpy as np
    import torch
    import torch.distributed as dist

    from spectrogram_midi_b200 import distributed as D

    sr, seg_s = 44100, 30.0
    n_seg = int(os.environ.get("AEGIS_BENCH_LONG_SEGMENTS", "120"))     # 120 x 30 s = one hour
    plan = P.corpus.plan_events(n_seg, seg_s, sr, first_seed=7000)      # every rank renders the same recording
    y = core.synth_events(n_seg, int(seg_s * sr), plan, dev).reshape(-1).cpu().numpy()
    torch.cuda.synchronize()

    def digest(res):
        h = hashlib.sha256()
        for k in ("rake_mask", "f0", "voiced_flag", "voiced_probs", "rms"):
            h.update(np.ascontiguousarray(res[k]).tobytes())
        if "events" in res:
            h.update(np.ascontiguousarray(res["events"]).tobytes())
        return h.hexdigest()[:16]

    out = {"clip_seconds": n_seg * seg_s, "sr": sr, "n_samples": int(y.shape[0]), "ranks": world}
    for mode in ("exact", "windowed"):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = D.analyze_long_clip(y, sr=sr, mode=mode, burn_seconds=2.0, return_events=True)
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        counts = D.gather_counts(int(res["events_local"]))
        out[mode] = {"wall_s": float(tt[0]), "audio_s_per_s": n_seg * seg_s / float(tt[0]), "sha16": digest(res),
                     "events": int(len(res["events"])), "events_per_rank": counts, "frames": int(len(res["f0"]))}
        if mode == "exact":
            exact_res = res
    # one rank alone (no collectives): the serial answer
    if rank == 0:
        t0 = time.perf_counter()
        solo = D.analyze_long_clip(y, sr=sr, mode="exact", windows_per_rank=world, return_events=True, solo=True)
        torch.cuda.synchronize()
        out["one_rank"] = {"wall_s": time.perf_counter() - t0, "sha16": digest(solo), "events": int(len(solo["events"]))}
        out["exact"]["equals_one_rank"] = out["exact"]["sha16"] == out["one_rank"]["sha16"]
        both = exact_res["voiced_flag"] & solo["voiced_flag"]
        out["exact"]["voiced_agreement"] = float((exact_res["voiced_flag"] == solo["voiced_flag"]).mean())
        out["speedup_vs_one_rank"] = out["one_rank"]["wall_s"] / out["exact"]["wall_s"]
        del both
    dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pyin", action="store_true", help="skip the transcription (pYIN) record")
    ap.add_argument("--no-long-clip", action="store_true", help="skip the one-hour clip record of N > 1 runs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
