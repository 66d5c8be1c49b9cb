#!/usr/bin/env python
"""bench.py -- feature frames/sec of the Meyda frame path on B200.

Workload (BASELINE.json configs[3], the one `metric` is quoted on): synthetic
20,000 clips x 30 s @ 44.1 kHz mono float32, bufferSize 2048, hop 512, all 18
features; clips are sharded over the ranks (strong scaling, no collective on
the data path).  A "step" is one pass over the rank's resident clips, run in
waves through a reused output ring because the full-set output (1.7 TB) does
not fit any memory.  If the clips do not fit the GPU's free memory the count
is reduced and `config.clips_total` says so.

  value     frames/s with the audio already resident in HBM (CUDA events, max over ranks)
  e2e       frames/s through the public host-memory API (pinned host buffers,
            H2D + kernel + D2H inside the timed region), same shapes, bounded batch
  roofline  algorithmic bytes per launch / average launch time vs measured HBM peak
  cpu_baseline  the oracle's C restatement on the host cores, bounded sample

`--impl reference` times the reference algorithm's CPU restatement (the oracle
port -- Node.js is absent, so the JS itself cannot run) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 44100.0
N, HOP = 2048, 512
CLIP_SECONDS = 30
CLIP_LEN = int(SR * CLIP_SECONDS)  # 1,323,000
CLIPS_TOTAL = 20000
METRIC = "feature frames/sec (full set, N=2048)"
UNIT = "frames/s"


def frames_per_clip():
    return (CLIP_LEN - N) // HOP + 1  # 2,580


def algorithmic_bytes_per_frame(features):
    """4*hop in + 4 * requested output floats."""
    per = {"buffer": N, "complexSpectrum": 2 * N, "amplitudeSpectrum": N // 2, "powerSpectrum": N // 2,
           "loudness": 25, "mfcc": 13}
    return 4 * HOP + 4 * sum(per.get(f, 1) for f in features)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled for the whole run (one nvidia-smi process: it needs a moment to
    start, and the short configurations last under a second); window(t0, t1) summarises the samples of one region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []  # (arrival time, fields)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def close(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def window(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)  # (let the samples of the region's tail arrive)
        sm, mx, reasons = [], [], set()
        for t, r in list(self.rows):
            if t < t0 or t > t1 + 0.15:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def gen_clips(torch, n_clips: int, seed: int, device, clip_len: int = CLIP_LEN):
    """Deterministic synthetic audio on the device: white noise (amp 0.25) + three
    sines (amp 0.2) at log-uniform 55..15000 Hz per clip; |x| < 0.85."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.empty(n_clips, clip_len, dtype=torch.float32, device=device)
    t = torch.arange(clip_len, dtype=torch.float32, device=device) / SR
    chunk = max(1, min(64, (64 * 1323000) // clip_len))
    for c0 in range(0, n_clips, chunk):
        c1 = min(n_clips, c0 + chunk)
        m = c1 - c0
        blk = (torch.rand(m, clip_len, device=device, generator=g) - 0.5) * 0.5
        f = torch.exp(torch.rand(m, 3, device=device, generator=g) * (np.log(15000.0) - np.log(55.0)) + np.log(55.0))
        ph = torch.rand(m, 3, device=device, generator=g) * (2 * np.pi)
        for p in range(3):
            blk += 0.2 * torch.sin(2 * np.pi * f[:, p:p + 1] * t[None, :] + ph[:, p:p + 1])
        x[c0:c1] = blk
        del blk
    return x


def run_cpu_oracle(sample_clips: np.ndarray, threads: int):
    """Oracle port on the host cores; returns (frames, seconds)."""
    from oracle import c_oracle
    n_clips = sample_clips.shape[0]
    t0 = time.perf_counter()
    r = c_oracle.extract(sample_clips.reshape(-1), N, HOP, SR, "hanning", arrays=True, threads=threads,
                         n_clips=n_clips, ring_per_thread=256)
    dt = time.perf_counter() - t0
    return r["frames_processed"], dt


def host_sample_clips(n_clips: int, length: int = CLIP_LEN) -> np.ndarray:
    from oracle import meyda_oracle as mo
    return np.stack([mo.synth_clip(i, length) for i in range(n_clips)])


def cpu_calibrated_sample(threads: int, target_s: float):
    """Pick a clip length/count so the oracle runs for about target_s seconds."""
    probe = host_sample_clips(threads, N + HOP * 63)  # 64 frames per thread
    f, dt = run_cpu_oracle(probe, threads)
    rate = f / max(dt, 1e-6)
    want_frames = max(64 * threads, int(rate * target_s))
    per_thread = max(64, want_frames // threads)
    clips_per_thread = max(1, -(-per_thread // frames_per_clip()))
    fpc = min(frames_per_clip(), -(-per_thread // clips_per_thread))
    length = N + HOP * (fpc - 1)
    base = host_sample_clips(threads, length)
    return np.concatenate([base] * clips_per_thread, axis=0), fpc


def reference_arm(args):
    """`--impl reference`: the reference algorithm on the CPU (oracle port, all
    host threads), bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    clips, fpc = cpu_calibrated_sample(threads, target_s=6.0)
    for _ in range(args.warmup):
        run_cpu_oracle(clips[:, : N + HOP * 31], threads)
    times, frames = [], 0
    for _ in range(args.steps):
        f, dt = run_cpu_oracle(clips, threads)
        times.append(dt)
        frames = f
    total_t = sum(times)
    value = frames * args.steps / total_t
    sample = "%d clips x %d frames (bufferSize 2048, hop 512, all 18 features) per step" % (clips.shape[0], fpc)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(CLIPS_TOTAL, args.gpus, None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Node.js is absent from this image, so the reference's JavaScript cannot run; this is the "
                "oracle's C restatement of the same algorithm (f64 arithmetic, f32 stores) on pthreads",
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(clips_total, n_gpus, wave_clips):
    return {"workload": "BASELINE configs[3]: synthetic clips x 30 s @44.1 kHz mono f32, bufferSize=2048 hop=512, "
                        "all 18 features, clip-sharded",
            "clips_total": clips_total, "clip_seconds": CLIP_SECONDS, "sample_rate": SR, "bufferSize": N,
            "hop": HOP, "features": "all 18", "window": "hanning", "frames_per_clip": frames_per_clip(),
            "parallelism": "clip-shard x%d, no collective" % n_gpus, "wave_clips": wave_clips,
            "l2_policy": "inputs (>=13 GB per rank) and outputs far larger than the 126 MB L2; no flush needed"}


def bind_near_gpu(index: int):
    """Multi-GPU runs: keep this rank's host threads (and therefore its pinned staging buffers, first-touch) on the
    CPUs NVML reports as local to its GPU, so that the end-to-end leg's PCIe copies do not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d cpus local to gpu %d" % (len(cpus), index)
    except Exception as e:  # best effort: the measurement does not depend on it
        return "unbound (%s)" % (str(e)[:60],)
    return "unbound"


def alloc_outputs(torch, plan, nf, dev, pin=False):
    return {k: torch.empty((nf,) + tuple(s[1:]), dtype=torch.int32 if d == np.int32 else torch.float32,
                           **({"pin_memory": True} if pin else {"device": dev}))
            for k, (s, d) in plan.output_shapes(nf).items()}


def flops_per_frame(n, feats):
    """real FFT 2.5 N log2 N + window N + magnitude 3 N/2 + ~2 flop per bin and running sum of the
    requested epilogues (+ the 13 x 26 DCT)."""
    m = n // 2
    f = 2.5 * n * np.log2(n) + n + 3 * m
    sums = 0
    if set(feats) & {"spectralCentroid", "spectralSpread", "spectralSkewness", "spectralKurtosis"}:
        sums += 4
    if "spectralFlatness" in feats:
        sums += 1
    if "spectralSlope" in feats:
        sums += 2
    if "spectralRolloff" in feats:
        sums += 1
    if set(feats) & {"loudness", "perceptualSpread", "perceptualSharpness"}:
        sums += 1
    if "mfcc" in feats:
        sums += 2
    f += 2 * m * sums
    if set(feats) & {"powerSpectrum", "mfcc"}:
        f += m
    if "mfcc" in feats:
        f += 2 * 13 * 26
    return float(f)


def run_resident(torch, dist, mb, dev, world, n, hop, feats, clips_rank, clip_len, seed, steps, warmup, flags=0,
                 ring_bytes_max=24 << 30, sampler=None):
    """Device-resident throughput of one configuration on this rank: `clips_rank` synthetic clips of `clip_len`
    samples stay in HBM, a step is one pass over them in waves through a reused output ring; CUDA events on the
    launching stream.  Returns a dict (this rank's numbers; the caller reduces over ranks)."""
    plan = mb.Plan(n, hop, SR, "hanning", feats, device=dev.index, flags=flags)
    fpc = (clip_len - n) // hop + 1
    _, lay = plan.query(np.array([clip_len], np.int64))
    out_bpf = max(4, int(lay.bytes_per_frame))
    free_b, _ = torch.cuda.mem_get_info(dev)
    budget = int(free_b * 0.92)
    ring_budget = min(ring_bytes_max, budget // 5)
    wave = max(1, min(clips_rank, ring_budget // (fpc * out_bpf)))
    max_clips = (budget - wave * fpc * out_bpf) // (clip_len * 4)
    reduced = clips_rank > max_clips
    clips_rank = int(min(clips_rank, max_clips))
    if world > 1:  # every rank processes the same number of clips
        t = torch.tensor([clips_rank], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        reduced = reduced or int(t.item()) < clips_rank
        clips_rank = int(t.item())
    wave = min(wave, clips_rank)
    x = gen_clips(torch, clips_rank, seed=seed, device=dev, clip_len=clip_len)
    outs = alloc_outputs(torch, plan, wave * fpc, dev)
    ptrs = {k: v.data_ptr() for k, v in outs.items()}
    stream = torch.cuda.current_stream(dev)
    assert stream.cuda_stream != 0
    plan.set_stream(stream.cuda_stream)
    tabs = [(np.arange(w0, min(clips_rank, w0 + wave), dtype=np.int64) * clip_len,
             np.full(min(clips_rank, w0 + wave) - w0, clip_len, np.int64)) for w0 in range(0, clips_rank, wave)]

    def step():
        for off, ln in tabs:
            plan.extract_device(x.data_ptr(), x.numel(), off, ln, ptrs, sync=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_load0 = time.time()  # (the clocks window covers warm-up and timed steps: both run the same kernel under load)
    for _ in range(warmup):
        step()
    barrier()
    l0 = plan.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.window(t_load0, time.time()) if sampler else None
    res = {"plan": plan, "x": x, "ms": ms, "launches": plan.launch_count - l0, "main_launches": len(tabs) * steps,
           "frames_rank": clips_rank * fpc, "clips_rank": clips_rank, "wave": wave, "fpc": fpc, "reduced": reduced,
           "refined_last_wave": plan.refined_frames, "clocks": clocks, "kernel": plan.kernel_name}
    del outs
    return res


def reduce_max(torch, dist, dev, world, v):
    if world == 1:
        return v
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(torch, dist, dev, world, v):
    if world == 1:
        return v
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def parity_spot_check(torch, plan, x, clip_len, n, hop, dev, nfc=40):
    """This rank's first clips and its LAST clip (SURVEY.md 8d generator rule) against the oracle, flat tolerances,
    no noise band: returns (note, banded count)."""
    from oracle import c_oracle
    from tests import parity
    L = n + hop * (nfc - 1)
    picks = sorted(set([0, 1, 2, 3, x.shape[0] - 1]) & set(range(x.shape[0])))
    small = alloc_outputs(torch, plan, nfc, dev)
    banded = 0
    for c in picks:
        plan.extract_device(x.data_ptr(), x.numel(), np.array([c * clip_len], np.int64), np.array([L], np.int64),
                            {k: v.data_ptr() for k, v in small.items()})
        ref = c_oracle.extract(x[c, :L].cpu().numpy(), n, hop, SR)
        b = parity.compare_all({k: v.cpu().numpy() for k, v in small.items()}, ref, n, noise_band=None)
        banded += sum(b.values())
    return "ok (%d frames of clips %s vs the oracle, flat tolerances, no noise band)" % (nfc, picks), banded


def pcie_ceiling(torch, dev, h2d_bytes, d2h_bytes, reps=3):
    """Raw pinned-memory copies of the same byte counts, both directions at once on two streams: what the host link
    gives this rank while every other rank does the same (the caller puts a barrier in front)."""
    hin = torch.empty(max(1, h2d_bytes), dtype=torch.uint8, pin_memory=True)
    hout = torch.empty(max(1, d2h_bytes), dtype=torch.uint8, pin_memory=True)
    din = torch.empty(max(1, h2d_bytes), dtype=torch.uint8, device=dev)
    dout = torch.empty(max(1, d2h_bytes), dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    best = None
    for r in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if r > 0:
            best = dt if best is None else min(best, dt)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=CLIPS_TOTAL, help="total clips over all ranks (default: the named config)")
    ap.add_argument("--features", default="all")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip BASELINE configs[2] and [4]")
    ap.add_argument("--generic", action="store_true", help="force the generic kernel")
    ap.add_argument("--no-refine", action="store_true", help="float32 FFT only: no adaptive exact second pass (A/B)")
    args = ap.parse_args()
    if os.environ.get("MEYDA_B200_HOST_THREADS"):  # (tuning runs: tools/gpu.sh; the default is the library's own choice)
        import meyda_b200 as _mb
        _mb.set_host_threads(int(os.environ["MEYDA_B200_HOST_THREADS"]))
    if os.environ.get("MEYDA_B200_HOST_ROWS"):
        import meyda_b200 as _mb
        _mb.set_host_rows(int(os.environ["MEYDA_B200_HOST_ROWS"]))
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    import meyda_b200 as mb
    from meyda_b200 import _capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the Meyda B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_note = bind_near_gpu(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    # A real (non-default) stream: the legacy default stream's handle is 0, which mb_plan_set_stream
    # reads as "use the plan's own stream" -- the events must sit on the launching stream.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    feats = mb.FEATURES if args.features == "all" else args.features.split(",")
    from meyda_b200.sharding import shard_range
    c0, c1 = shard_range(args.clips, world, rank)
    flags = (_capi.MB_FLAG_GENERIC_KERNEL if args.generic else 0) | (_capi.MB_FLAG_NO_REFINE if args.no_refine else 0)

    # ---- headline: BASELINE configs[3]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    r = run_resident(torch, dist, mb, dev, world, N, HOP, feats, c1 - c0, CLIP_LEN, 0x4D455944 + rank, args.steps,
                     args.warmup, flags=flags, sampler=sampler)
    plan, x, fpc = r["plan"], r["x"], r["fpc"]
    ms_max = reduce_max(torch, dist, dev, world, r["ms"])
    launches_all = int(reduce_sum(torch, dist, dev, world, r["launches"]))
    frames_all = r["frames_rank"] * world
    value = frames_all * args.steps / (ms_max * 1e-3)
    clocks = r["clocks"]

    # ---- roofline of the dominant kernel on this rank (one launch per wave; an adaptive plan adds one, normally
    # empty, exact-FFT launch per wave, which is inside the same timed region)
    peak, peak_src = measured_peaks()
    alg_bpf = algorithmic_bytes_per_frame(feats)
    avg_launch_s = (r["ms"] * 1e-3) / max(1, r["main_launches"])
    frames_per_launch = r["frames_rank"] * args.steps / max(1, r["main_launches"])
    achieved = alg_bpf * frames_per_launch / avg_launch_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            # measured by ncu on a full launch of this kernel with the full feature set; DRAM traffic is
            # proportional to the frames of a launch, so it is restated for this run's (average) launch size
            if tj.get("kernel") == plan.kernel_name and feats == mb.FEATURES:
                traffic = tj["dram_bytes_per_launch"] * frames_per_launch / tj["frames_per_launch"]
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": "ncu --set full capture (profiles/traffic.json), scaled to this run's frames per launch",
                "peak_source": peak_src, "kernel": plan.kernel_name,
                "algorithmic_bytes_per_frame": alg_bpf, "frames_per_launch": frames_per_launch,
                "avg_launch_ms": avg_launch_s * 1e3}

    # ---- parity spot check against the oracle (outside the timed region)
    parity_note, parity_banded = None, None
    if rank == 0:
        try:
            parity_note, parity_banded = parity_spot_check(torch, plan, x, CLIP_LEN, N, HOP, dev)
        except AssertionError as e:  # report, never hide
            parity_note = "FAILED: %s" % (str(e)[:200],)

    # ---- e2e: public host-memory API, H2D + kernel + D2H timed; pinned (the facade's default) and pageable buffers
    e2e = None
    if not args.no_e2e:
        e2e_clips = 12
        host_x = torch.empty(e2e_clips, CLIP_LEN, dtype=torch.float32, pin_memory=True)
        host_x.copy_(x[:e2e_clips])
        nf = e2e_clips * fpc
        hx = host_x.numpy().reshape(-1)
        ho = plan.alloc_host_outputs(nf, pinned=True)  # mb_host_alloc, what Plan.extract_host and js/addon.cc hand out
        off = np.arange(e2e_clips, dtype=np.int64) * CLIP_LEN
        ln = np.full(e2e_clips, CLIP_LEN, np.int64)
        plan.set_stream(None)

        def timed(fn, reps):
            for _ in range(2):
                fn()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return reduce_max(torch, dist, dev, world, (time.perf_counter() - t0) / reps)

        dt = timed(lambda: plan.extract_host(hx, off, ln, out=ho), args.steps)
        rows_mode = mb.get_host_rows()  # 0 / 1 / 2 (mb_set_host_rows; the default picks by cores per visible device)
        host_made = ({"buffer"} | ({"power_spectrum"} if "amplitude_spectrum" in ho else set())) if rows_mode >= 1 else set()  # rows the host fills itself
        half = {"complex_real", "complex_imag"} if rows_mode >= 2 else set()  # bins 0 .. N/2 copied, the mirrored half made on the host
        d2h = int(sum((v.nbytes // N * (N // 2 + 1) if k in half else v.nbytes) for k, v in ho.items() if k not in host_made))
        e2e = {"value": nf * world / dt, "unit": UNIT, "h2d_bytes_per_step": int(hx.nbytes), "d2h_bytes_per_step": d2h,
               "host_rows_mode": rows_mode,
               "batch": "%d clips x 30 s per rank per step, mb_extract(MB_MEM_HOST) into mb_host_alloc (pinned) arrays; host-rows mode %d: "
                        "%s produced on the host while the device works and not copied back" % (
                            e2e_clips, rows_mode, {0: "nothing is", 1: "the `buffer` rows (the caller's own samples) and the powerSpectrum rows "
                            "(amplitude squared) are", 2: "the `buffer` rows, the powerSpectrum rows and the mirrored half of complexSpectrum are"}[rows_mode])}
        # the host link's own ceiling for these byte counts, every rank at once
        if world > 1:
            dist.barrier()
        ceil_dt = reduce_max(torch, dist, dev, world, pcie_ceiling(torch, dev, int(hx.nbytes), d2h))
        e2e["pcie_ceiling"] = {"value": nf * world / ceil_dt, "unit": UNIT,
                               "h2d_GBps_per_rank": hx.nbytes / ceil_dt / 1e9, "d2h_GBps_per_rank": d2h / ceil_dt / 1e9,
                               "how": "raw pinned cudaMemcpyAsync of the same H2D and D2H byte counts on two streams, all ranks at once"}
        e2e["frac_of_pcie_ceiling"] = e2e["value"] / e2e["pcie_ceiling"]["value"]
        ho_page = plan.alloc_host_outputs(nf, pinned=False)
        hx_page = np.array(hx)
        dtp = timed(lambda: plan.extract_host(hx_page, off, ln, out=ho_page), max(1, min(args.steps, 3)))
        e2e["pageable"] = {"value": nf * world / dtp, "unit": UNIT, "note": "the same call with pageable numpy arrays on both sides"}
        if numa_note:
            e2e["host_binding"] = numa_note
        # one process driving every visible GPU through mb_extract_multi (what the N-API caller does)
        if world == 1 and torch.cuda.device_count() > 1:
            ndev = torch.cuda.device_count()
            plans = [plan] + [mb.Plan(N, HOP, SR, "hanning", feats, device=d, flags=flags) for d in range(1, ndev)]
            mc = e2e_clips * ndev
            mx = torch.empty(mc, CLIP_LEN, dtype=torch.float32, pin_memory=True)
            for i in range(ndev):
                mx[i * e2e_clips:(i + 1) * e2e_clips].copy_(x[:e2e_clips])
            mo_ = plan.alloc_host_outputs(mc * fpc, pinned=True)
            moff = np.arange(mc, dtype=np.int64) * CLIP_LEN
            mln = np.full(mc, CLIP_LEN, np.int64)
            mxn = mx.numpy().reshape(-1)
            dtm = timed(lambda: mb.meyda.extract_multi(plans, mxn, moff, mln, out=mo_), max(1, min(args.steps, 3)))
            e2e["multi_device_one_process"] = {"value": mc * fpc / dtm, "unit": UNIT, "devices": ndev,
                                               "call": "mb_extract_multi, %d clips per device per step" % e2e_clips}
            for p_ in plans[1:]:
                p_.close()
            del mx, mo_
        del host_x, ho, ho_page
        plan.set_stream(stream.cuda_stream)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        clips, cfpc = cpu_calibrated_sample(threads, target_s=12.0)
        f, dt = run_cpu_oracle(clips, threads)
        cpu_baseline = {"value": f / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "%d clips x %d frames, all 18 features, oracle C restatement on %d pthreads (%.1f s)"
                                  % (clips.shape[0], cfpc, threads, dt)}
    kernel_name, wave_clips, reduced, clips_used = plan.kernel_name, r["wave"], r["reduced"], r["clips_rank"] * world
    refined_last = r["refined_last_wave"]
    plan.close()
    del x, r, plan
    torch.cuda.empty_cache()

    # ---- BASELINE configs[2] and [4] (compute-bound: reported against the FP32 FFMA peak measured in this process)
    secondary = []
    if not args.no_secondary:
        import ctypes
        f32 = ctypes.c_double(0)
        _capi.check(_capi.lib().mb_measure_peaks(local_rank, ctypes.byref(f32), None))
        ffma = float(f32.value)
        C3 = ["mfcc", "spectralCentroid", "spectralSpread", "spectralSkewness", "spectralKurtosis"]
        C5 = ["amplitudeSpectrum", "spectralRolloff", "spectralFlatness", "spectralSlope"]
        for name, n2, hop2, feats2, total, clen in (
                ("BASELINE configs[2]: synthetic 10,000 clips x 10 s @44.1 kHz mono f32, bufferSize=2048 hop=512, mfcc + spectral moments",
                 2048, 512, C3, max(world, 10000 * args.clips // CLIPS_TOTAL), 441000),
                ("BASELINE configs[4]: bufferSize=32768 hop=8192, amplitudeSpectrum + rolloff/flatness/slope over a 1,024-channel x 60 s synthetic array",
                 32768, 8192, C5, max(world, 1024 * args.clips // CLIPS_TOTAL), 2646000)):
            a0, a1 = shard_range(total, world, rank)
            r2 = run_resident(torch, dist, mb, dev, world, n2, hop2, feats2, a1 - a0, clen, 0x4D455944 + 1000 + rank,
                              args.steps, args.warmup, flags=flags, sampler=sampler)
            ms2 = reduce_max(torch, dist, dev, world, r2["ms"])
            fps = r2["frames_rank"] * world * args.steps / (ms2 * 1e-3)
            fl = flops_per_frame(n2, feats2)
            bpf2 = 4 * hop2 + 4 * sum({"amplitudeSpectrum": n2 // 2, "mfcc": 13}.get(f_, 1) for f_ in feats2)
            note2 = None
            if rank == 0:
                try:
                    note2, _ = parity_spot_check(torch, r2["plan"], r2["x"], clen, n2, hop2, dev, nfc=6 if n2 > 4096 else 40)
                except AssertionError as e:
                    note2 = "FAILED: %s" % (str(e)[:200],)
            per_gpu = fps / world
            secondary.append({
                "config": {"workload": name, "clips_total": r2["clips_rank"] * world, "bufferSize": n2, "hop": hop2,
                           "features": feats2, "frames_per_clip": r2["fpc"], "wave_clips": r2["wave"],
                           "parallelism": "clip-shard x%d, no collective" % world},
                "metric": "feature frames/sec", "value": fps, "unit": UNIT, "ms_per_step": ms2 / args.steps,
                "kernel": r2["kernel"], "gpu_launches": int(reduce_sum(torch, dist, dev, world, r2["launches"])),
                "roofline": {"bound": "fp32", "achieved": per_gpu * fl / 1e12, "peak": ffma, "unit": "TFLOP/s",
                             "frac": per_gpu * fl / 1e12 / ffma, "flops_per_frame": fl,
                             "peak_source": "FFMA micro-kernel run in this process (mb_measure_peaks)",
                             "hbm_frac": per_gpu * bpf2 / 1e9 / peak, "algorithmic_bytes_per_frame": bpf2,
                             # what the committed ncu capture of this kernel shows to be busiest (not measured in this run)
                             "limiter_ncu": ("L1 / shared-memory data pipe 78 % busy (810 wavefronts per frame), issue slots 58 %: profiles/r02_ncu_c3.txt"
                                             if n2 == 2048 else
                                             "one CTA of 16 warps per SM, ~14 block barriers per frame; issue slots 43 %, no pipe above 45 %: profiles/r02_ncu_big32768.txt")},
                "clocks": r2["clocks"], "parity": note2})
            r2["plan"].close()
            del r2
            torch.cuda.empty_cache()

    if sampler:
        sampler.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(clips_used, world, wave_clips),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches_all,
            "clocks": clocks, "parity": parity_note, "parity_banded": parity_banded, "kernel": kernel_name,
            "frames_per_step": frames_all, "refined_frames_last_wave": refined_last,
            "secondary": secondary,
        }
        if reduced:
            line["config"]["note"] = "clip count reduced from %d to fit GPU memory" % args.clips
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
