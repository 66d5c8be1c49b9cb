"""Turns gpurun_out/{launches.csv, prof_*.ncu-rep} into the tracked summaries under profiles/."""
import csv, collections, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
name = sys.argv[2] if len(sys.argv) > 2 else "warp2048"
out_dir = os.path.join(ROOT, "profiles"); os.makedirs(out_dir, exist_ok=True)
go = os.path.join(ROOT, "gpurun_out")

# ---- launch list: per-kernel totals and shares
if name == "warp2048":  # (the launch list is taken on the headline bench only)
    rows = [r for r in csv.reader(open(os.path.join(go, "launches.csv"))) if r and not r[0].startswith("==")]
    hdr = rows[0]; col = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr) or r[col["Metric Name"]] != "gpu__time_duration.sum": continue
        k = r[col["Kernel Name"]]; v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v_us
    tot = sum(v for _, v in agg.values())
    mine = sum(v for k, (_, v) in agg.items() if "mb_" in k)
    with open(os.path.join(out_dir, f"{tag}_launches_{name}.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# command: python bench.py --clips 600 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary (includes torch's input generation)\n")
        f.write("%-90s %6s %12s %7s\n" % ("kernel", "count", "total_us", "share"))
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-90s %6d %12.1f %6.2f%%\n" % (k[:90], n, v, 100 * v / tot))
        f.write("\nmeyda_b200 kernels: %.1f us of %.1f us profiled (%.1f%%); the rest is torch generating the synthetic clips\n" % (mine, tot, 100 * mine / tot))
        steps = [(k, n, v) for k, (n, v) in agg.items() if "mb_" in k]
        f.write("share of the feature kernel inside the bench step (only meyda_b200 kernels run in the timed region): %s\n" % ", ".join("%s %.1f%%" % (k.split("(")[0][-40:], 100 * v / mine) for k, n, v in steps))


# ---- full capture: headline metrics
rep = os.path.join(go, f"prof_{name}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines())); h = r[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg"]
vals = {}
with open(os.path.join(out_dir, f"{tag}_ncu_{name}.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:... -c 1 (one launch, ~40 replays; tools/gpu.sh profile %s)\n" % name)
    f.write("# kernel: %s\n" % r[2][h.index("Kernel Name")])
    for i, m in enumerate(h):
        if m in want:
            f.write("%-70s %-14s %s\n" % (m, r[1][i], r[2][i])); vals[m] = r[2][i]
    frames = float(sys.argv[3]) if len(sys.argv) > 3 else None
    if frames:
        unit = {m: r[1][i] for i, m in enumerate(h)}
        gb = lambda m: float(vals[m]) * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "Tbyte": 1e3}.get(unit[m], 1.0)
        dr, dw = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
        bpf = int(sys.argv[4]) if len(sys.argv) > 4 else 35016
        f.write("\nframes in this launch: %d; algorithmic bytes %d x %d = %.3f GB; DRAM traffic %.3f GB (%.3fx)\n" % (
            frames, frames, bpf, frames * bpf / 1e9, dr + dw, (dr + dw) / (frames * bpf / 1e9)))
        f.write("warp-instructions per frame: %.0f\n" % (float(vals["smsp__inst_executed.sum"]) / frames))
        if name == "warp2048":  # bench.py reads this one for the headline kernel's roofline.traffic
            json.dump({"dram_bytes_per_launch": (dr + dw) * 1e9, "frames_per_launch": frames, "kernel": name, "round": tag},
                      open(os.path.join(out_dir, "traffic.json"), "w"))
# ---- stall / opcode breakdown from the source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, x in enumerate(rows) if x and x[0] == "Address"); hd = rows[hi]; c = {x: i for i, x in enumerate(hd)}
def F(x):
    try: return float(x)
    except ValueError: return 0.0
st = collections.Counter(); ops = collections.Counter()
for x in rows[hi + 1:]:
    for s in hd:
        if s.startswith("stall_") and "Not Issued" not in s: st[s] += F(x[c[s]])
    t = x[c["Source"]].strip().split()
    if t: ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += F(x[c["Instructions Executed"]])
with open(os.path.join(out_dir, f"{tag}_ncu_{name}.txt"), "a") as f:
    S = sum(st.values()); T = sum(ops.values())
    f.write("\nwarp stall samples: " + ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100 * v / S) for k, v in st.most_common(10)) + "\n")
    f.write("executed SASS by opcode (warp-instructions per frame): " + ", ".join("%s %.0f" % (k, v / (frames or 1)) for k, v in ops.most_common(24)) + "\n")
    sass = subprocess.run("cuobjdump -sass %s | grep -cE 'UBLKCP'" % os.path.join(ROOT, "meyda_b200/_lib/libmeyda_b200.so"), shell=True, capture_output=True, text=True).stdout.strip()
    f.write("TMA evidence: %s UBLKCP (cp.async.bulk) instructions in the library's SASS\n" % sass)
print(open(os.path.join(out_dir, f"{tag}_ncu_{name}.txt")).read())
