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
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
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
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
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


def gen_clips(torch, n_clips: int, seed: int, device):
    """Deterministic synthetic audio on the device: white noise (amp 0.25) + three
    sines (amp 0.2) at log-uniform 55..15000 Hz per clip; |x| < 0.85."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.empty(n_clips, CLIP_LEN, dtype=torch.float32, device=device)
    t = torch.arange(CLIP_LEN, dtype=torch.float32, device=device) / SR
    chunk = 64
    for c0 in range(0, n_clips, chunk):
        c1 = min(n_clips, c0 + chunk)
        m = c1 - c0
        blk = (torch.rand(m, CLIP_LEN, device=device, generator=g) - 0.5) * 0.5
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
    ap.add_argument("--generic", action="store_true", help="force the generic kernel")
    ap.add_argument("--no-refine", action="store_true", help="float32 FFT only: no adaptive exact second pass (A/B)")
    args = ap.parse_args()
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

    feats = mb.FEATURES if args.features == "all" else args.features.split(",")
    from meyda_b200.sharding import shard_range
    c0, c1 = shard_range(args.clips, world, rank)
    my_clips = c1 - c0
    fpc = frames_per_clip()
    plan = mb.Plan(N, HOP, SR, "hanning", feats, device=local_rank,
                   flags=(_capi.MB_FLAG_GENERIC_KERNEL if args.generic else 0) | (_capi.MB_FLAG_NO_REFINE if args.no_refine else 0))
    _, lay = plan.query(np.array([CLIP_LEN], np.int64))
    out_bpf = int(lay.bytes_per_frame)

    # ---- memory plan: resident clips + output ring for one wave
    free_b, _total_b = torch.cuda.mem_get_info(dev)
    budget = int(free_b * 0.92)
    ring_budget = min(24 << 30, budget // 5)
    wave_clips = max(1, min(my_clips, ring_budget // (fpc * out_bpf)))
    ring_bytes = wave_clips * fpc * out_bpf
    max_clips = (budget - ring_bytes) // (CLIP_LEN * 4)
    reduced = False
    if my_clips > max_clips:
        my_clips = int(max_clips)
        reduced = True
    if world > 1:  # every rank processes the same number of clips
        t = torch.tensor([my_clips], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) < my_clips:
            my_clips, reduced = int(t.item()), True
    wave_clips = min(wave_clips, my_clips)
    x = gen_clips(torch, my_clips, seed=0x4D455944 + rank, device=dev)
    shapes = plan.output_shapes(wave_clips * fpc)
    outs = {k: torch.empty(s, dtype=torch.int32 if d == np.int32 else torch.float32, device=dev)
            for k, (s, d) in shapes.items()}
    out_ptrs = {k: v.data_ptr() for k, v in outs.items()}
    # A real (non-default) stream: the legacy default stream's handle is 0, which mb_plan_set_stream
    # reads as "use the plan's own stream" -- the events below must sit on the launching stream.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    plan.set_stream(stream.cuda_stream)
    waves = [(w0, min(my_clips, w0 + wave_clips)) for w0 in range(0, my_clips, wave_clips)]
    wave_tabs = [(np.arange(w0, w1, dtype=np.int64) * CLIP_LEN, np.full(w1 - w0, CLIP_LEN, np.int64))
                 for w0, w1 in waves]

    def step():
        for off, ln in wave_tabs:
            plan.extract_device(x.data_ptr(), x.numel(), off, ln, out_ptrs, sync=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = plan.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = plan.launch_count - launches0
    refined_last_wave = plan.refined_frames  # frames of the last wave redone with the exact FFT (adaptive plans)
    # the dominant kernel runs once per wave; an adaptive plan adds one (normally empty) exact-FFT launch per wave
    main_launches = len(wave_tabs) * args.steps
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches_all = int(lt.item())
    else:
        ms_max, launches_all = ms, launches
    frames_rank = my_clips * fpc
    frames_all = frames_rank * world
    value = frames_all * args.steps / (ms_max * 1e-3)

    # ---- roofline of the dominant (only) kernel on this rank
    peak, peak_src = measured_peaks()
    alg_bpf = algorithmic_bytes_per_frame(feats)
    avg_launch_s = (ms * 1e-3) / max(1, main_launches)
    frames_per_launch = frames_rank * args.steps / max(1, main_launches)
    achieved = alg_bpf * frames_per_launch / avg_launch_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            # measured by ncu on a full 302-clip launch of this kernel with the full feature set; DRAM traffic is
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

    # ---- parity spot check against the oracle on this rank's first clip (outside the timed region)
    parity_note = None
    if rank == 0:
        try:
            from oracle import c_oracle
            from tests import parity
            nfc = 40
            L = N + HOP * (nfc - 1)
            small = {k: torch.empty((nfc,) + tuple(s[1:]), dtype=torch.int32 if d == np.int32 else torch.float32,
                                    device=dev) for k, (s, d) in plan.output_shapes(nfc).items()}
            plan.extract_device(x.data_ptr(), x.numel(), np.array([0], np.int64), np.array([L], np.int64),
                                {k: v.data_ptr() for k, v in small.items()})
            ref = c_oracle.extract(x[0, :L].cpu().numpy(), N, HOP, SR)
            parity.compare_all({k: v.cpu().numpy() for k, v in small.items()}, ref, N)
            parity_note = "ok (%d frames vs oracle)" % nfc
        except AssertionError as e:  # report, never hide
            parity_note = "FAILED: %s" % (str(e)[:200],)

    # ---- e2e: public host-memory API, pinned buffers, H2D + kernel + D2H timed
    e2e = None
    if not args.no_e2e:
        e2e_clips = 12
        host_x = torch.empty(e2e_clips, CLIP_LEN, dtype=torch.float32, pin_memory=True)
        host_x.copy_(x[:e2e_clips])
        nf = e2e_clips * fpc
        host_out = {k: torch.empty((nf,) + tuple(s[1:]), dtype=torch.int32 if d == np.int32 else torch.float32,
                                   pin_memory=True) for k, (s, d) in plan.output_shapes(nf).items()}
        hx = host_x.numpy().reshape(-1)
        ho = {k: v.numpy() for k, v in host_out.items()}
        off = np.arange(e2e_clips, dtype=np.int64) * CLIP_LEN
        ln = np.full(e2e_clips, CLIP_LEN, np.int64)
        plan.set_stream(None)
        for _ in range(2):
            plan.extract_host(hx, off, ln, out=ho)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            plan.extract_host(hx, off, ln, out=ho)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": nf * world * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(hx.nbytes),
               "d2h_bytes_per_step": int(sum(v.nbytes for v in ho.values())),
               "batch": "%d clips x 30 s per rank per step, pinned host memory, mb_extract(MB_MEM_HOST)" % e2e_clips}
        if numa_note:
            e2e["host_binding"] = numa_note
        del host_x, host_out

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        clips, cfpc = cpu_calibrated_sample(threads, target_s=12.0)
        f, dt = run_cpu_oracle(clips, threads)
        cpu_baseline = {"value": f / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "%d clips x %d frames, all 18 features, oracle C restatement on %d pthreads (%.1f s)"
                                  % (clips.shape[0], cfpc, threads, dt)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(my_clips * world, world, wave_clips),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches_all,
            "clocks": clocks, "parity": parity_note, "kernel": plan.kernel_name,
            "frames_per_step": frames_all, "refined_frames_last_wave": refined_last_wave,
        }
        if reduced:
            line["config"]["note"] = "clip count reduced from %d to fit GPU memory" % args.clips
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
