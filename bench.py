#!/usr/bin/env python
"""Benchmark of the Aegis hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at every N (weak scaling, sharded by clip, no data-path collective): BASELINE.json
configs[1] -- a batch of 1,024 synthetic 30 s clips at 22 050 Hz per GPU; one step = one pass of
STFT |X| + mel dB onset strength + onset peak picking + frame RMS over the batch.
`value`  : audio-seconds analysed per second with the batch resident in HBM (CUDA events).
`e2e`    : the same through the host-buffer plugin call (pinned host audio -> H2D -> kernels -> D2H of
           RMS / onset envelope / onset flags), copies inside the timed region.
`roofline`: the STFT kernel (dominant), algorithmic bytes 4*N + 4*1025*T per clip over its own
           CUDA-event time, against MEASURED_PEAKS.json's HBM copy bandwidth.
`cpu_baseline` / `--impl reference`: the CPU oracle (port of the reference's librosa path; librosa
           itself is not installable here) on the box's host cores, on a bounded sample of the same clips.
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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    start.record()
    for _ in range(args.steps):
        step(timed=True)
        launches += 4  # stft_fused, mel_post, peak_candidates, peak_select
        ev_b.synchronize()
        stft_ms.append(ev_a.elapsed_time(ev_b))
    stop.record()
    barrier()
    elapsed_ms = start.elapsed_time(stop)

    # ---- end to end through the host-buffer plugin call
    pipe = batch.HostPipeline(N_CLIPS, n_samples, sr=SR, hop_length=HOP, device=dev, chunk_clips=128)
    y_host = torch.empty((N_CLIPS, n_samples), dtype=torch.float32, pin_memory=True)
    y_host.copy_(y)
    for _ in range(2):
        pipe.run(y_host)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        res = pipe.run(y_host)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    clocks = sampler.stop()

    # ---- the same call fed with 16-bit PCM (what a WAV file holds): half the PCIe bytes, K9 converts on the device.
    # A secondary number: the samples are the float clips rounded to int16, so results differ by that quantisation.
    h2d_bytes, d2h_bytes = int(pipe.h2d_bytes), int(pipe.d2h_bytes)
    del pipe, y_host
    pipe16 = batch.HostPipeline(N_CLIPS, n_samples, sr=SR, hop_length=HOP, device=dev, chunk_clips=128, pcm_rate=SR)
    pcm_host = torch.empty((N_CLIPS, n_samples), dtype=torch.int16, pin_memory=True)
    pcm_host.copy_((y * 32767.0).round().to(torch.int16))
    for _ in range(2):
        pipe16.run(pcm_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe16.run(pcm_host)
    barrier()
    pcm_s = (time.perf_counter() - t0) / e2e_steps

    t = torch.tensor([elapsed_ms, e2e_s, pcm_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_s, pcm_s = float(t[0]), float(t[1]), float(t[2])
    audio_s_per_step = world * N_CLIPS * CLIP_SECONDS
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

    # DRAM traffic of that kernel per launch: dram__bytes_read.sum + dram__bytes_write.sum from the committed
    # `ncu --set full` capture of the same launch shape (not re-measured here: no profiler inside a timed run)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_stft_traffic.json")))
        if N_CLIPS == 1024:
            traffic = float(tr["dram_bytes_read"]) + float(tr["dram_bytes_write"])
    except (OSError, KeyError, ValueError):
        pass

    aux = {}
    if rank == 0 and not args.no_pyin:  # cfg3 on the same clips (not the headline; reported for context)
        try:
            sub = y[: min(N_CLIPS, 256)]
            for _ in range(2):
                core.pyin_batch(sub, sr=SR, fmin=batch.E2, fmax=batch.C6)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            core.pyin_batch(sub, sr=SR, fmin=batch.E2, fmax=batch.C6)
            b.record()
            b.synchronize()
            aux["pyin_E2_C6_audio_s_per_s"] = sub.shape[0] * CLIP_SECONDS / (a.elapsed_time(b) / 1e3)
            aux["pyin_clips"] = int(sub.shape[0])
            # the whole perception phase of audio_to_midi (spectral + rake + pYIN + RMS + trend) on the same clips
            for _ in range(2):
                batch.analyze_batch(sub, sr=SR, with_onsets=True, with_trend=True)
            torch.cuda.synchronize()
            a.record()
            batch.analyze_batch(sub, sr=SR, with_onsets=True, with_trend=True)
            b.record()
            b.synchronize()
            aux["full_perception_audio_s_per_s"] = sub.shape[0] * CLIP_SECONDS / (a.elapsed_time(b) / 1e3)
        except Exception as e:  # pragma: no cover
            aux["pyin_error"] = str(e)[:200]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        pool = CpuPool(cores)                        # spawned, single-threaded workers (no fork of this CUDA process)
        rate = pooled_rate(pool, y[: 2 * cores].cpu().numpy())
        n = int(min(N_CLIPS, max(cores, round(15.0 * rate))))   # ~15 s of wall with every core busy
        cpu = cpu_baseline_record(pool, y[:n].cpu().numpy(), f"first {n} clips of the batch ({n * CLIP_SECONDS:.0f} audio-s)")
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
                         "kernel": "stft_fused_kernel", "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": stft_avg_ms,
                         "peak_source": peak_src, "share_of_step": stft_avg_ms / (elapsed_ms / args.steps)},
            "cpu_baseline": cpu,
            "e2e": {"value": audio_s_per_step / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_s * 1e3, "d2h": "rms + onset envelope + onset flags (|X| stays in HBM)"},
            "e2e_pcm16": {"value": audio_s_per_step / pcm_s, "unit": UNIT, "h2d_bytes_per_step": int(pipe16.h2d_bytes),
                          "d2h_bytes_per_step": int(pipe16.d2h_bytes), "ms_per_step": pcm_s * 1e3,
                          "note": "secondary: same call with the clips as 16-bit PCM host buffers (a WAV payload); scaled to float on the GPU (K9)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "aux": aux,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pyin", action="store_true", help="skip the auxiliary pYIN timing")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
