"""CPU tests of the multi-GPU host logic (gloo, world_size 2): sharding, window planning, ragged
gathers, and the long-clip stitching of ``distributed.analyze_long_clip`` with the per-window
analysis replaced by an oracle-backed back end (test seam; the product default is CUDA-only)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import spectrogram_midi_b200 as P  # noqa: E402
from spectrogram_midi_b200 import batch, corpus, tables  # noqa: E402
from spectrogram_midi_b200 import distributed as D  # noqa: E402

from oracle import librosa_ref as L  # noqa: E402
from oracle import reference_files as R  # noqa: E402

SR, HOP = 22050, 512
E2, C6 = L.note_to_hz("E2"), L.note_to_hz("C6")


def test_shard_range_is_a_balanced_partition():
    for n in (0, 1, 7, 8, 1024, 100000):
        for world in (1, 2, 3, 4, 8):
            parts = [batch.shard_range(n, r, world) for r in range(world)]
            assert sum(len(p) for p in parts) == n
            assert [i for p in parts for i in p] == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


@pytest.mark.parametrize("n_samples,world,burn", [(220500, 2, 86), (220500, 8, 32), (5000, 4, 32), (158_760_000, 8, 172)])
def test_plan_windows_cover_the_clip(n_samples, world, burn):
    T = 1 + n_samples // HOP
    ws = D.plan_windows(n_samples, HOP, world, burn)
    assert [w.own_lo for w in ws][0] == 0 and ws[-1].own_hi == T
    for a, b in zip(ws[:-1], ws[1:]):
        assert a.own_hi == b.own_lo
    for w in ws:
        assert w.win_lo <= w.own_lo <= w.own_hi <= w.win_hi
        lo = max(0, w.own_lo - burn)
        assert w.win_lo == lo - lo % 8 and w.win_hi == min(T, w.own_hi + burn)
        assert 0 <= w.s0 <= w.s1 <= n_samples and w.pad % 4 == 0 and 0 <= w.pad <= 1024
        # first frame of the window starts at s0 - pad in clip coordinates
        assert w.s0 - w.pad == w.win_lo * HOP - 1024
        if w.win_hi > w.win_lo:   # the slice reaches the end of the last frame or the end of the clip
            assert w.s1 == min(n_samples, (w.win_hi - 1) * HOP + 1024)


class OracleBackend:
    """Per-window analysis with the CPU oracle, in the tensor format distributed.CudaBackend produces."""

    def __init__(self, ratio=0.6):
        self.cfg = tables.pyin_config(float(SR), HOP, E2, C6)
        self.ratio = ratio
        self.trans, _ = L.pyin_transition(self.cfg.n_pitch_bins, 10, SR, HOP)

    def _frames(self, y_win, w):
        n = w.win_hi - w.win_lo
        buf = np.zeros((n - 1) * HOP + 2048, np.float32)
        seg = y_win[: len(buf) - w.pad]
        buf[w.pad : w.pad + len(seg)] = seg
        idx = np.arange(2048)[:, None] + HOP * np.arange(n)[None, :]
        return buf[idx]

    def features(self, y_win, w):
        import scipy.fft

        cfg = self.cfg
        fr = self._frames(y_win, w)
        S = np.abs(scipy.fft.rfft(L.hann_window()[:, None] * fr, axis=0).astype(np.complex64))
        mel = np.einsum("ft,mf->mt", S ** 2, L.mel_filterbank(SR), optimize=True).astype(np.float32)
        rms = np.sqrt(np.mean(np.abs(fr) ** 2, axis=0))
        yin = L.cmnd(fr, 2048, 1024, cfg.min_period, cfg.max_period)
        th, beta = L.beta_threshold_probs()
        obs, vp = L.pyin_observations(yin, L.parabolic_interpolation(yin), SR, th, 2, beta, 0.01, cfg.min_period, E2,
                                      cfg.n_pitch_bins, 10)
        n, nb, mc = fr.shape[1], cfg.n_pitch_bins, cfg.max_troughs
        cb, cp, cc = np.zeros((n, mc), np.int16), np.zeros((n, mc)), np.zeros(n, np.int32)
        for t in range(n):
            nz = np.flatnonzero(obs[:nb, t])
            cc[t] = len(nz)
            cb[t, : len(nz)] = nz
            cp[t, : len(nz)] = obs[nz, t]
        return {"mel": torch.from_numpy(mel)[None], "mel_max": torch.tensor([mel.max()]), "rms": torch.from_numpy(rms),
                "cand_bin": torch.from_numpy(cb), "cand_prob": torch.from_numpy(cp), "cand_count": torch.from_numpy(cc),
                "voiced_prob": torch.from_numpy(vp[0])}

    def rake_mask(self, feat, mel_max):
        mel = feat["mel"][0].numpy()
        db = 10.0 * np.log10(np.maximum(1e-10, mel)) - 10.0 * np.log10(np.maximum(1e-10, np.float32(mel_max[0])))
        db = np.maximum(db, np.float32(-80.0))
        return torch.from_numpy(R.detect_rake_patterns(db, HOP, SR, self.ratio).astype(np.uint8))

    def decode(self, cand_bin, cand_prob, cand_count, voiced_prob):
        nb = self.cfg.n_pitch_bins
        T = cand_bin.shape[0]
        obs = np.zeros((2 * nb, T))
        cb, cp, cc = cand_bin.numpy(), cand_prob.numpy(), cand_count.numpy()
        for t in range(T):
            obs[cb[t, : cc[t]], t] = cp[t, : cc[t]]
        obs[nb:, :] = (1 - voiced_prob.numpy()[None]) / nb
        p_init = np.zeros(2 * nb)
        p_init[nb:] = 1 / nb
        st = L.viterbi(obs, self.trans, p_init)
        f0 = self.cfg.freqs[st % nb].copy()
        voiced = st < nb
        f0[~voiced] = np.nan
        return {"f0": torch.from_numpy(f0), "voiced_flag": torch.from_numpy(voiced.astype(np.uint8)),
                "states": torch.from_numpy(st.astype(np.int16))}

    def note_events(self, res, **kwargs):
        """the reference's v1 logic filter (oracle restatement) -> the 40-byte records kernel K7 writes"""
        return _event_records(R.get_midi_events(res["rake_mask"], np.nan_to_num(res["f0"]), res["voiced_flag"], res["voiced_probs"],
                                                res["rms"], SR, HOP, kwargs.get("confidence_threshold", 0.7)))


def _event_records(ev):
    import warnings  # noqa: F401

    rec = np.zeros(len(ev), P.core.NOTE_EVENT_DTYPE)
    for i, e in enumerate(ev):
        rec[i]["note"], rec[i]["start"], rec[i]["end"], rec[i]["velocity"] = e["note"], e["start"], e["end"], e["velocity"]
        rec[i]["rms_energy"], rec[i]["track"] = e["rms_energy"], e["track"] == "main"
        rec[i]["technique"] = P.core.TECHNIQUES.index(e.get("technique"))
        rec[i]["confidence"], rec[i]["slope"] = e["confidence"], e.get("slope", 0.0)
    return rec


def _worker(rank, world, port, y, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ragged gather + max-reduce + event gather
        parts = D.all_gather_ragged(torch.arange(3 + 2 * rank, dtype=torch.float64)[:, None] + rank)
        assert [p.shape[0] for p in parts] == [3 + 2 * r for r in range(world)]
        assert float(D.all_reduce_max(torch.tensor([float(rank)]))[0]) == world - 1
        ev = np.zeros(rank + 1, D.EVENT_DTYPE)
        ev["note"], ev["start"] = 40 + rank, np.arange(rank + 1)
        allev = D.gather_note_events(ev)
        assert len(allev) == sum(r + 1 for r in range(world)) and list(allev["note"][:1]) == [40]
        assert D.gather_counts(10 + rank) == [10 + r for r in range(world)]
        assert D.collective_device() is None   # gloo: host tensors
        # by-clip sharding with fewer clips than ranks: a rank without clips still takes part in the count gather
        empty = D.analyze_clips_sharded(lambda idx: torch.zeros((0, 8)), 0, sr=SR)
        assert empty["clips_per_rank"] == [0] * world and empty["clip_indices"] == []
        res = {}
        for mode in ("exact", "windowed"):
            res[mode] = D.analyze_long_clip(y, sr=SR, hop_length=HOP, mode=mode, burn_seconds=1.5, backend=OracleBackend(),
                                            return_events=(mode == "exact"))
        ev = res["exact"].pop("events")
        res["exact"]["events_bytes"] = ev.view(np.uint8)
        res["exact"]["events_local"] = np.array([res["exact"]["events_local"]])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{f"{m}/{k}": v for m, r in res.items() for k, v in r.items()})
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_long_clip_on_two_ranks_equals_the_serial_result(tmp_path):
    y = corpus.test_track(SR, seed=2)  # ~4.2 s with rakes and silences
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), y, str(tmp_path)), nprocs=world, join=True, start_method="fork")
    f0, vf, vp = L.pyin(y, fmin=E2, fmax=C6, sr=SR, hop_length=HOP)
    rake = R.detect_rake_patterns(L.load_audio_features(y, SR), HOP, SR, 0.6)
    rms = L.rms(y)[0]
    local_counts = []
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        # exact mode: identical to the serial decode, on every rank
        np.testing.assert_array_equal(z["exact/voiced_flag"], vf)
        np.testing.assert_array_equal(z["exact/f0"], f0)
        np.testing.assert_array_equal(z["exact/voiced_probs"], vp)
        np.testing.assert_array_equal(z["exact/rake_mask"], rake)
        np.testing.assert_allclose(z["exact/rms"], rms, rtol=1e-6)
        # the gathered note events (each rank contributed the events that start in its own frames) are the serial list
        serial = {k.split("/")[1]: z[k] for k in z.files if k.startswith("exact/")}
        want = _event_records(R.get_midi_events(serial["rake_mask"], np.nan_to_num(serial["f0"]), serial["voiced_flag"],
                                                serial["voiced_probs"], serial["rms"], SR, HOP, 0.7))
        got = z["exact/events_bytes"].view(P.core.NOTE_EVENT_DTYPE)
        assert len(got) == len(want) > 0 and got.tobytes() == want.tobytes()
        local_counts.append(int(z["exact/events_local"][0]))
        # windowed mode: frame-local outputs identical, decode equal once the paths have coalesced
        np.testing.assert_array_equal(z["windowed/rake_mask"], rake)
        np.testing.assert_array_equal(z["windowed/voiced_probs"], vp)
        assert (z["windowed/voiced_flag"] == vf).mean() >= 0.99
        assert len(z["windowed/f0"]) == len(f0)
    assert sum(local_counts) == len(got) and min(local_counts) >= 0   # a partition of the list, contributed by both ranks


def test_single_process_path_needs_no_group():
    y = corpus.test_track(SR, seed=2)[: SR * 2]
    res = D.analyze_long_clip(y, sr=SR, hop_length=HOP, backend=OracleBackend())
    f0, vf, vp = L.pyin(y, fmin=E2, fmax=C6, sr=SR, hop_length=HOP)
    np.testing.assert_array_equal(res["voiced_flag"], vf)
    np.testing.assert_array_equal(res["f0"], f0)
