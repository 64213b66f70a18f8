#!/usr/bin/env python
"""bench.py — MFCC frames/sec on B200 (BASELINE.json `metric`), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload A|B|C]

A "step" is one pass of the hot path (mfcc_compute_batch) over one resident
batch of synthetic PCM.  At N=1 the workload is BASELINE.json configs[1]:
1,024 x 10 s of 16 kHz int16, 25 ms / 10 ms, 512-pt FFT, 26 mel, 13 MFCC
(1,021,952 frames, 327.7 MB in — larger than the 126 MB L2, so every step
re-reads its input from HBM).  At N>1 every rank holds its own such batch
(weak scaling, configs[4]: "10k hours sharded", iterated over a resident buffer)
and there is no data-path collective.

value     device-timed frames/s, inputs resident in HBM, CUDA events, max over ranks
e2e       same metric through mfcc_compute_host: pinned HOST buffers, H2D + kernels + D2H timed
roofline  FFT FLOPs (2.5 N log2 N per frame) against the FP32 FMA peak measured live by
          tools/microbench; `roofline_hbm` is the stream-bytes view (hop*2 + n_cep*4 B per frame)
cpu_baseline / --impl reference
          the in-repo C baseline (oracle/mfcc_cpu_fast.c: the oracle's spec with a real-input FFT, -O3, AVX2
          clones; checked against the plain oracle) on all host cores — the nominal reference,
          simotin13/mfcc, is a C compiler with no MFCC path to time
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mfcc_b200 import CONFIGS  # noqa: E402
from mfcc_b200.synth import fast_fixed_batch, ragged_batch  # noqa: E402

def _fixed(n_utts, n_samp):
    return lambda seed: fast_fixed_batch(n_utts, n_samp, seed=seed)


def _ragged(n_utts, lo, hi):
    return lambda seed: ragged_batch(n_utts, lo, hi, seed=seed)


def _tiled(n_utts, n_samp, distinct):
    """Long streams: `distinct` independent noise streams repeated to n_utts (host RAM and start-up time)."""
    def make(seed):
        base, _ = fast_fixed_batch(distinct, n_samp, seed=seed)
        pcm = np.tile(base, n_utts // distinct)
        return pcm, np.arange(n_utts + 1, dtype=np.int64) * n_samp
    return make


WORKLOADS = {
    # name: (config, maker(seed) -> (pcm, offsets), description, maker of ONE --impl reference step: a bounded
    #        sample of the same workload, about 1 M frames (0.2 M for the 2048-point config))
    "A": ("A", _fixed(1024, 160000),
          "configs[1]: 1024 x 10 s 16 kHz int16; frame 400 / hop 160 / nfft 512 / 26 mel / 13 cep", _fixed(1024, 160000)),
    "B": ("B", _fixed(16384, 16000),
          "configs[2] fixed-2.0 s variant: 16384 x 2 s 8 kHz int16; 200/80/256/20/13", _fixed(5120, 16000)),
    "B3": ("B", _ragged(16384, 4000, 24000),
           "configs[2]: 16384 utterances of 0.5-3.0 s (uniform, seed 3) 8 kHz int16, one launch; 200/80/256/20/13", _ragged(5120, 4000, 24000)),
    "C": ("C", _tiled(64, 28800000, 8),
          "configs[3]: 64 x 10 min 48 kHz int16 (8 distinct noise streams repeated); 1200/480/2048/80/40", _fixed(64, 1440000)),
    "C8": ("C", _fixed(16, 28800000 // 4),
           "configs[3] reduced: 16 x 150 s 48 kHz int16; 1200/480/2048/80/40", _fixed(64, 1440000)),
}


def fft_flops(nfft: int) -> float:
    return 2.5 * nfft * math.log2(nfft)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows), "reasons": reasons}


def measured_fp32_peak(sm_mhz=None):
    """FP32 FMA peak in TFLOP/s measured live by tools/fp32_peak (>= 5 ms per launch, 20 launches back to back between
    two events); returns (TFLOP/s, how, details).  `sm_mhz` is the SM clock bench.py sampled under load: the tool then
    also states measured / (SMs x 128 x 2 x that clock)."""
    exe = os.path.join(ROOT, "tools", "fp32_peak")
    derived = 148 * 128 * 2 * 1.965e9 / 1e12
    try:
        cmd = [exe] + (["--sm-mhz", str(sm_mhz)] if sm_mhz else [])
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=120).stdout
        d = json.loads(out.strip().splitlines()[-1])
        return float(d["tflops"]), "measured live: tools/fp32_peak (scalar FFMA, 32 warps/SM, 20 launches of >= 5 ms)", d
    except Exception:
        return derived, "derived 148 SM x 128 lanes x 2 x 1.965 GHz (tools/fp32_peak unavailable)", None


def cpu_leg(p, pcm, offsets, n_utts_sample, threads, repeats=1):
    """Time the oracle on a bounded sample (the first n_utts_sample utterances)."""
    import oracle
    off = offsets[: n_utts_sample + 1]
    x = pcm[: int(off[-1])]
    oracle.mfcc_batch(p, x[: int(off[1])], off[:2], nthreads=1, fast=True)  # build + page in
    best = float("inf")
    frames = 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        out, fo = oracle.mfcc_batch(p, x, off, nthreads=threads, fast=True)
        best = min(best, time.perf_counter() - t0)
        frames = int(fo[-1])
    return frames / best, frames, best


def run_reference(args, p, cfg_name, desc, ref_maker):
    """--impl reference: the CPU implementation of the path on the host cores.  The nominal
    reference has none, so this is the in-repo scalar C oracle ("port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    pcm, off = ref_maker(1000)
    n_utts = len(off) - 1
    import oracle
    oracle.mfcc_batch(p, pcm[: int(off[1])], off[:2], nthreads=1, fast=True)
    for _ in range(args.warmup):
        oracle.mfcc_batch(p, pcm, off, nthreads=cores, fast=True)
    t0 = time.perf_counter()
    frames = 0
    for _ in range(args.steps):
        _, fo = oracle.mfcc_batch(p, pcm, off, nthreads=cores, fast=True)
        frames += int(fo[-1])
    dt = time.perf_counter() - t0
    v = frames / dt
    sample = (f"{n_utts} utterances, {int(off[-1])} samples per step ({frames // args.steps} frames), "
              f"{cores} pthreads over utterances")
    line = {
        "impl": "reference", "metric": "mfcc_frames_per_sec", "value": v, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "audio_seconds_per_s": v * p.hop_len / p.sample_rate,
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "in-repo C baseline (oracle/mfcc_cpu_fast.c: real-input FFT through a half-size complex FFT, -O3, "
                                 "AVX2 + FMA clones of the hot loops, pthreads over utterances; checked against the plain parity "
                                 "oracle); simotin13/mfcc has no MFCC path to time"},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    init when NCCL_DEBUG is set in the environment), so fd 1 is pointed at stderr for the whole run and the result
    line goes to the saved descriptor."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


def workload_config(name, world):
    """The `config` object of the JSON line: a function of the workload definition only, so that the two arms
    (ours / --impl reference) print the same one."""
    cfg_name, _, desc, _ = WORKLOADS[name]
    return {"workload": desc, "params": cfg_name, "workload_key": name,
            "l2": "input per step exceeds the 126 MB L2: every timed step re-reads it from HBM (no flush needed)",
            "parallelism": f"utterance-sharded x{world}, no data-path collective"}


class Ctx:
    pass


def measure(ctx, name, steps, warmup, e2e_steps, with_cpu, peak=None):
    """Device-timed steps, the e2e leg, rooflines and (N = 1) the CPU baseline of ONE workload."""
    import torch
    import torch.distributed as dist
    from mfcc_b200 import api
    world, rank, local = ctx.world, ctx.rank, ctx.local
    cfg_name, maker, desc, ref_maker = WORKLOADS[name]
    p = CONFIGS[cfg_name]()
    plan = api.Plan(p, device=local, kernel=ctx.kernel)

    # Synthetic batch, generated on the host, resident in HBM before timing.
    pcm, off = maker(1000 + rank)
    n_utts = len(off) - 1
    batch = plan.batch(off)
    frames = batch.total_frames
    in_bytes, out_bytes = pcm.nbytes, frames * plan.out_dim * 4
    if in_bytes <= 126 * 1024 * 1024:
        raise SystemExit("bench.py: the workload's input must exceed the 126 MB L2 (timed steps re-read it from HBM)")
    d_pcm = torch.from_numpy(pcm).cuda()
    d_out = torch.empty((frames, plan.out_dim), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        plan.compute_batch(batch, d_pcm, d_out, stream)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    # keep the GPU busy long enough for nvidia-smi to see load: untimed pre-roll, then the timed K steps
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 0.05:
        plan.compute_batch(batch, d_pcm, d_out, stream)
        torch.cuda.synchronize()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = api.launch_count()
    t_host0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(steps):
        plan.compute_batch(batch, d_pcm, d_out, stream)
    ev1.record(stream)
    barrier()
    launches = api.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    # hold the load a little longer so the sampler has rows inside the region even for short runs
    t_hold = time.perf_counter()
    while time.perf_counter() - t_hold < 0.6:
        plan.compute_batch(batch, d_pcm, d_out, stream)
        torch.cuda.synchronize()
    clocks = sampler.stop(t_host0 - 0.3, time.perf_counter()) if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        fr = torch.tensor([frames], dtype=torch.float64, device="cuda")
        dist.all_reduce(fr, op=dist.ReduceOp.SUM)
        total_frames = int(fr.item())
    else:
        total_frames = frames
    value = total_frames * steps / (ms * 1e-3)

    # ---- e2e: host buffers through mfcc_compute_host (H2D + kernels + D2H inside the timed region) ----
    h_in = api.PinnedBuffer((pcm.size,), np.int16)
    h_in.array[:] = pcm
    h_out = api.PinnedBuffer((frames, plan.out_dim), np.float32)
    plan.compute_host(h_in.array, off, h_out.array)  # warm: allocates the plan's device buffers
    plan.compute_host(h_in.array, off, h_out.array)
    barrier()
    e2e_step_ms = []
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        plan.compute_host(h_in.array, off, h_out.array)   # returns with the features in h_out
        e2e_step_ms.append(round((time.perf_counter() - ts) * 1e3, 3))
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_value = total_frames * e2e_steps / e2e_dt
    e2e_ok = bool(np.array_equal(h_out.array, d_out.cpu().numpy()))
    # telephony workloads: the same batch shape as G.711 mu-law codes (1 byte per sample over PCIe), through
    # mfcc_compute_host_g711 — the codes are expanded inside the kernel's staging
    e2e_g711 = None
    if cfg_name == "B":
        h_codes = api.PinnedBuffer((pcm.size,), np.uint8)
        h_codes.array[:] = (pcm >> 8).astype(np.uint8)            # any byte stream is a valid code stream
        plan.compute_host(h_codes.array, off, h_out.array, alaw=False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.compute_host(h_codes.array, off, h_out.array, alaw=False)
        torch.cuda.synchronize()
        dt_g = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt_g], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_g = float(t.item())
        e2e_g711 = {"value": total_frames * e2e_steps / dt_g, "unit": "frames/s", "h2d_bytes_per_step": int(pcm.size),
                    "d2h_bytes_per_step": out_bytes, "api": "mfcc_compute_host_g711 (mu-law codes in pinned host memory)"}
        h_codes.close()
    # the same pipeline with the fused post-processing between kernel and read-back (mfcc_compute_host_post): PCM in,
    # per-utterance CMVN + delta + delta-delta rows (3 x out_dim columns) out — three times the bytes on the way back
    e2e_post = None
    if name == "A":
        h_out3 = api.PinnedBuffer((frames, 3 * plan.out_dim), np.float32)
        plan.compute_host(h_in.array, off, h_out3.array, post=(2, 2, 2))
        plan.compute_host(h_in.array, off, h_out3.array, post=(2, 2, 2))
        barrier()
        l_p0 = api.launch_count()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.compute_host(h_in.array, off, h_out3.array, post=(2, 2, 2))
        torch.cuda.synchronize()
        dt_p = time.perf_counter() - t0
        l_p1 = api.launch_count()
        if world > 1:
            t = torch.tensor([dt_p], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_p = float(t.item())
        e2e_post = {"value": total_frames * e2e_steps / dt_p, "unit": "frames/s", "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": 3 * out_bytes, "gpu_launches_per_step": (l_p1 - l_p0) // max(e2e_steps, 1),
                    "api": "mfcc_compute_host_post (CMVN mean + variance, regression window 2, order 2: 39 columns back)"}
        h_out3.close()
    kernel_name = plan.kernel_name
    out_dim = plan.out_dim
    h_in.close()
    h_out.close()
    del d_pcm, d_out, batch
    plan.close()
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel (the only kernel in the step) ----
    per_launch_s = ms * 1e-3 / max(launches, 1)
    frames_per_launch = frames  # one launch covers the rank's whole batch
    if peak is None:
        peak = measured_fp32_peak(clocks.get("sm_mhz") if clocks else None)
    fp32_peak, fp32_how, fp32_detail = peak
    flops = fft_flops(p.nfft) * frames_per_launch
    ach_tf = flops / per_launch_s / 1e12
    bytes_alg = (p.hop_len * 2 + out_dim * 4) * frames_per_launch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = {"bound": "fp32", "achieved": ach_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": ach_tf / fp32_peak, "traffic": None, "kernel": kernel_name,
                "algorithmic": f"2.5*N*log2(N) = {fft_flops(p.nfft):.0f} FLOP/frame x {frames_per_launch} frames/launch",
                "peak_source": fp32_how, "peak_detail": fp32_detail,
                "note": "the path is FP32 CUDA-core bound (FFT FLOPs / stream bytes = 24-50 FLOP/B against a ridge of about 11), "
                        "so the binding roofline is the measured FP32 FMA peak; the HBM view is in roofline_hbm"}
    roofline_hbm = {"bound": "hbm", "achieved": bytes_alg / per_launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": bytes_alg / per_launch_s / 1e9 / hbm_peak, "traffic": None,
                    "algorithmic": f"{p.hop_len * 2 + out_dim * 4} B/frame x {frames_per_launch} frames/launch",
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            t = json.load(open(prof)).get(kernel_name)
            if t:   # measured DRAM bytes per frame (one ncu --set full capture) x the frames of one launch
                roofline["traffic"] = roofline_hbm["traffic"] = int(round(t["bytes_per_frame"] * frames_per_launch))
                roofline["traffic_source"] = t["capture"]
        except Exception:
            pass

    cpu = None
    if with_cpu and world == 1:   # the CPU baseline is reported at N = 1 only
        os.sched_setaffinity(0, ctx.full_affinity)   # this leg uses every host core
        cores = os.cpu_count() or 1
        # bounded sample: configs[1] is small enough to run whole (~5 CPU-seconds of scalar C per pass, best of 3);
        # the other workloads use the --impl reference step's sample of the same workload
        pcm_r, off_r = (pcm, off) if name == "A" else ref_maker(1000)
        n_s = len(off_r) - 1
        v, fr_s, dt = cpu_leg(p, pcm_r, off_r, n_s, cores, repeats=3)
        n_1 = max(1, n_s // 4)
        v1, fr_1, dt1 = cpu_leg(p, pcm_r, off_r, n_1, 1)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{n_s} utterances of the workload ({fr_s} frames), {cores} pthreads over utterances, best of 3 passes, {dt:.2f} s; single thread: first {n_1} utterances, {dt1:.2f} s",
               "single_thread_value": v1,
               "note": "in-repo C baseline (oracle/mfcc_cpu_fast.c: real-input FFT through a half-size complex FFT, -O3, AVX2 + FMA "
                       "clones of the hot loops; 2.1 x the plain parity oracle per thread); simotin13/mfcc has no MFCC path to time"}

    return {
        "value": value, "unit": "frames/s", "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "config": workload_config(name, world),
        "batch": {"frames_per_step_per_gpu": frames, "utterances": n_utts, "input_mb_per_step": round(in_bytes / 1e6, 1)},
        "kernel": kernel_name,
        "audio_seconds_per_s": value * p.hop_len / p.sample_rate,
        # BASELINE.json configs[4] (10,000 hours of 16 kHz PCM over the N GPUs, iterated over the resident batch): what this rate
        # means for that corpus — device-timed, and end to end from host buffers
        "configs4_10k_hours": {"device_timed_s": 3.6e7 / (value * p.hop_len / p.sample_rate),
                               "from_host_buffers_s": 3.6e7 / (e2e_value * p.hop_len / p.sample_rate)} if name == "A" else None,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                "steps": e2e_steps, "step_ms": e2e_step_ms,
                "api": "mfcc_compute_host (pinned host buffers, 4-stream chunk pipeline)",
                "host_cores_bound": len(ctx.numa_cores) if ctx.numa_cores else None,
                "matches_device_path": e2e_ok},
        "e2e_g711": e2e_g711,
        "e2e_post": e2e_post,
        "gpu_launches": launches,
        "roofline": roofline, "roofline_hbm": roofline_hbm,
        "cpu_baseline": cpu,
        "_peak": peak,
    }


def measure_stream(ctx, n_streams=1024, chunk_ms=20, calls=100, warm=10):
    """Serving shape: n_streams live 8 kHz telephony streams, every call delivers chunk_ms of new
    audio per stream through mfcc_stream_feed_many — one pinned staging copy, ONE kernel launch and one read-back per
    call whatever the stream count.  Host buffers in, host buffers out: the whole call is timed on the host clock."""
    from mfcc_b200 import api
    p = CONFIGS["B"]()
    plan = api.Plan(p, device=ctx.local, kernel=ctx.kernel)
    n = p.sample_rate * chunk_ms // 1000
    rng = np.random.default_rng(5)
    audio = np.clip(np.rint(rng.standard_normal((n_streams, (calls + warm + 2) * n), dtype=np.float32) * 3000.0), -32768, 32767).astype(np.int16)
    grp = api.StreamGroup(plan, n_streams, max_frames_per_feed=8)
    l0 = api.launch_count()
    frames, lat = 0, []
    for k in range(calls + warm):
        chunk = audio[:, k * n:(k + 1) * n]
        t0 = time.perf_counter()
        rows, counts = grp.feed(chunk)
        dt = time.perf_counter() - t0
        if k >= warm:
            lat.append(dt * 1e3)
            frames += int(counts.sum())
        elif k == warm - 1:
            l0 = api.launch_count()
    launches = api.launch_count() - l0
    # spot check against the offline rows of stream 0 (bit-identical by construction; tests/test_gpu_parity.py checks all)
    lat.sort()
    total_s = sum(lat) * 1e-3
    out = {"value": frames / total_s, "unit": "frames/s", "streams": n_streams, "chunk_ms": chunk_ms, "calls": calls,
           "frames_per_call": frames // calls, "ms_per_call": {"mean": sum(lat) / len(lat), "p50": lat[len(lat) // 2], "p99": lat[int(len(lat) * 0.99)]},
           "realtime_factor": (frames / total_s) * p.hop_len / p.sample_rate / n_streams,
           "gpu_launches": launches, "kernel": plan.kernel_name,
           "api": "mfcc_stream_feed_many (host buffers; one H2D, one launch, one D2H per call)",
           "config": {"workload": f"{n_streams} concurrent 8 kHz streams x {chunk_ms} ms chunks; 200/80/256/20/13", "params": "B"}}
    grp.close()
    plan.close()
    return out


def measure_post(ctx, steps=20, warmup=3, n_utts=4096, frames_per_utt=998):
    """Post-processing: per-utterance CMVN (mean and variance) + delta + delta-delta, fused and stacked
    (mfcc_post_batch), on a feature matrix of 4 x configs[1] (4,096 utterances x 998 frames x 13 cepstra = 212 MB, larger
    than the 126 MB L2, so every timed step re-reads it from HBM).  HBM-bound: the roofline is bytes over time."""
    import torch
    from mfcc_b200 import api
    p = CONFIGS["A"]()
    plan = api.Plan(p, device=ctx.local, kernel=ctx.kernel)
    L, H = p.frame_len, p.hop_len
    off = np.arange(n_utts + 1, dtype=np.int64) * (L + (frames_per_utt - 1) * H)
    batch = plan.batch(off)
    frames, dim = batch.total_frames, plan.out_dim
    assert frames == n_utts * frames_per_utt
    # synthetic cepstra of MFCC magnitude, made on the host (no library kernel runs in this process): c0 ~ 60 +- 4, the rest +- 4
    h = np.random.default_rng(11).standard_normal((frames, dim), dtype=np.float32) * np.float32(4.0)
    h[:, 0] += np.float32(60.0)
    feat = torch.from_numpy(h).cuda()
    del h
    out = torch.empty((frames, 3 * dim), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()
    res = {}
    l0 = l1 = 0
    for key, cmvn in (("cmvn_delta2", 2), ("delta2_only", 0)):
        for _ in range(warmup):
            plan.post(batch, feat, cmvn, 2, 2, out, stream)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = api.launch_count()
        ev0.record(stream)
        for _ in range(steps):
            plan.post(batch, feat, cmvn, 2, 2, out, stream)
        ev1.record(stream)
        torch.cuda.synchronize()
        l1 = api.launch_count()
        res[key] = ev0.elapsed_time(ev1) / steps
    ms = res["cmvn_delta2"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    b_full = (2 * dim + 3 * dim) * 4      # statistics pass read + apply pass read + stacked row written
    b_apply = (dim + 3 * dim) * 4
    line = {"value": frames / (ms * 1e-3), "unit": "frames/s", "steps": steps, "warmup": warmup, "ms_per_step": ms,
            "gpu_launches": (l1 - l0), "kernel": "post_stats_kernel + post_finalize_kernel + post_apply_kernel<2>",
            "api": "mfcc_post_batch (device feature matrix in, stacked static | delta | delta-delta matrix out)",
            "config": {"workload": f"{n_utts} utterances x {frames_per_utt} frames x {dim} cepstra (4 x configs[1]) -> {3 * dim} columns; "
                                   "CMVN mean + variance, regression window 2, order 2", "params": "A",
                       "l2": "input 212 MB and output 638 MB per step both exceed the 126 MB L2"},
            "roofline": {"bound": "hbm", "achieved": b_full * frames / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": b_full * frames / (ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                         "algorithmic": f"{b_full} B/frame (2 reads of {dim * 4} B: statistics pass and apply pass; {3 * dim * 4} B written) x {frames} frames per step (three launches)",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
            "delta2_only": {"ms_per_step": res["delta2_only"], "value": frames / (res["delta2_only"] * 1e-3),
                            "roofline_frac": b_apply * frames / (res["delta2_only"] * 1e-3) / 1e9 / hbm_peak,
                            "algorithmic": f"{b_apply} B/frame, one launch (no statistics pass)"}}
    try:   # measured DRAM bytes per frame of the two kernels together (one ncu --set full capture, profiles/r2_post.md)
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("post_cmvn_delta2")
        if t:
            line["roofline"]["traffic"] = int(round(t["bytes_per_frame"] * frames))
            line["roofline"]["traffic_source"] = t["capture"]
    except Exception:
        pass
    del feat, out, batch
    plan.close()
    torch.cuda.empty_cache()
    return line


def run_shard(ctx, args):
    """--shard: ONE ragged configs[2] batch (the same on every rank, as if read from shared storage), partitioned by
    cumulative frame count (sharding.partition); every rank runs mfcc_compute_host on ITS slice (host buffers in, host
    buffers out), then the optional feature gather over NCCL.  Strong scaling: the total work is fixed as N grows.
    Reported: frames/s of the sharded compute (max over ranks), the gather time on a side stream, and the time of both
    when the gather of chunk i overlaps nothing / runs after the compute (the collective is 52 B per frame)."""
    import torch
    import torch.distributed as dist
    from mfcc_b200 import api
    from mfcc_b200.sharding import partition, local_slice, gather_features, frame_counts
    world, rank, local = ctx.world, ctx.rank, ctx.local
    cfg_name, maker, desc, _ = WORKLOADS["B3"]
    p = CONFIGS[cfg_name]()
    plan = api.Plan(p, device=local, kernel=ctx.kernel)
    pcm, off = maker(1000)                       # identical on every rank
    parts = partition(p, off, world)
    u0, u1 = parts[rank]
    s0, s1, loc = local_slice(off, u0, u1)
    rows = int(frame_counts(p, off)[u0:u1].sum())
    total_rows = int(frame_counts(p, off).sum())
    h_in = api.PinnedBuffer((max(s1 - s0, 1),), np.int16)
    h_in.array[: s1 - s0] = pcm[s0:s1]
    h_out = api.PinnedBuffer((max(rows, 1), plan.out_dim), np.float32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    d_feat = torch.empty((max(rows, 1), plan.out_dim), dtype=torch.float32, device="cuda")
    side = torch.cuda.Stream()
    for _ in range(max(args.warmup, 2)):
        plan.compute_host(h_in.array[: s1 - s0], loc, h_out.array)
    # compute only
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        plan.compute_host(h_in.array[: s1 - s0], loc, h_out.array)
    torch.cuda.synchronize()
    dt_compute = max_over_ranks(time.perf_counter() - t0) / args.e2e_steps
    # gather only (device-resident rows, NCCL all_gather on a side stream), device-timed
    gather_ms, full_ok = None, None
    if world > 1:
        d_feat[:rows].copy_(torch.from_numpy(h_out.array[:rows]))
        for _ in range(2):
            with torch.cuda.stream(side):
                full, counts = gather_features(d_feat, rows, plan.out_dim)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            e0.record(side)
            for _ in range(args.e2e_steps):
                full, counts = gather_features(d_feat, rows, plan.out_dim)
            e1.record(side)
        barrier()
        gather_ms = max_over_ranks(e0.elapsed_time(e1) / args.e2e_steps)
        full_ok = bool(full.shape[0] == total_rows and sum(counts) == total_rows)
        # compute + gather: the gather of step i runs on the side stream while step i + 1 computes
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            plan.compute_host(h_in.array[: s1 - s0], loc, h_out.array)
            d_feat[:rows].copy_(torch.from_numpy(h_out.array[:rows]), non_blocking=True)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                full, counts = gather_features(d_feat, rows, plan.out_dim)
        torch.cuda.synchronize()
        dt_both = max_over_ranks(time.perf_counter() - t0) / args.e2e_steps
    else:
        dt_both = dt_compute
    if rank == 0:
        emit({"metric": "mfcc_frames_per_sec", "mode": "shard", "value": total_rows / dt_compute, "unit": "frames/s",
              "n_gpus": world, "steps": args.e2e_steps, "warmup": max(args.warmup, 2), "ms_per_step": dt_compute * 1e3,
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
              "config": workload_config("B3", world),
              "sharding": {"by": "cumulative frames, contiguous utterance ranges", "utterances_per_rank": [b - a for a, b in parts],
                           "frames_total": total_rows, "frames_rank0": rows},
              "collective": {"op": "NCCL all_gather of [rows_r, 13] f32 (padded to the largest block)", "ms": gather_ms,
                             "bytes_total": total_rows * plan.out_dim * 4, "result_complete": full_ok,
                             "compute_plus_gather_ms": dt_both * 1e3,
                             "overlap": "gather of step i on a side stream under the compute of step i + 1"},
              "value_with_gather": total_rows / dt_both,
              "api": "mfcc_compute_host per rank on its slice (host buffers), sharding.partition / gather_features"})
    h_in.close()
    h_out.close()
    plan.close()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="A", choices=sorted(WORKLOADS))
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "fused"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--shard", action="store_true",
                    help="strong-scaling mode: ONE ragged configs[2] batch partitioned by frames over the ranks, "
                         "mfcc_compute_host per rank, NCCL gather of the features timed beside it")
    ap.add_argument("--extra", default=None,
                    help="comma-separated workloads measured after the headline one and nested under `workloads` "
                         "(default at N = 1 with the default workload: B3,C — BASELINE.json configs[2] and configs[3]; 'none' to skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    cfg_name, maker, desc, ref_maker = WORKLOADS[args.workload]
    p = CONFIGS[cfg_name]()
    if args.impl == "reference":
        run_reference(args, p, cfg_name, desc, ref_maker)
        return

    import torch
    import torch.distributed as dist
    from mfcc_b200 import KERNEL_AUTO, KERNEL_GENERIC, KERNEL_FUSED

    ctx = Ctx()
    ctx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = rank = int(os.environ.get("RANK", "0"))
    ctx.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the MFCC path has no CPU fallback)")
    torch.cuda.set_device(local)
    from mfcc_b200.sharding import bind_near_gpu
    ctx.full_affinity = os.sched_getaffinity(0)
    ctx.numa_cores = bind_near_gpu(local)   # pinned e2e buffers are first-touched on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx.kernel = {"auto": KERNEL_AUTO, "generic": KERNEL_GENERIC, "fused": KERNEL_FUSED}[args.kernel]

    if args.shard:
        run_shard(ctx, args)
        if world > 1:
            dist.destroy_process_group()
        return
    res = measure(ctx, args.workload, args.steps, args.warmup, args.e2e_steps, not args.no_cpu)
    extra = args.extra
    if extra is None:
        extra = "B3,C" if (world == 1 and args.workload == "A") else "none"
    nested = {}
    if extra != "none" and world == 1 and res is not None:
        for name in extra.split(","):
            # shorter step counts: these lines ride along with the headline one
            r = measure(ctx, name, max(3, args.steps // 2), args.warmup, max(2, args.e2e_steps // 3), not args.no_cpu,
                        peak=res["_peak"])
            r.pop("_peak", None)
            nested[name] = r
        nested["stream"] = measure_stream(ctx)
        nested["post"] = measure_post(ctx)
    if rank == 0:
        res.pop("_peak", None)
        line = {"metric": "mfcc_frames_per_sec", "value": res.pop("value"), "unit": res.pop("unit"), "n_gpus": world,
                "steps": res.pop("steps"), "warmup": res.pop("warmup"), "ms_per_step": res.pop("ms_per_step"),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        line.update(res)
        if nested:
            line["workloads"] = nested
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
