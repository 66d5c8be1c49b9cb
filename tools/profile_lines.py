"""ncu source page of gpurun_out/prof_<name>.ncu-rep, aggregated by CUDA source line, into
profiles/<tag>_ncu_<name>_source.txt: share of executed warp-instructions, share of warp stall samples, shared-memory
wavefronts and the line's dominant stall reasons.  usage: profile_lines.py <tag> <name> [frames in the launch] [title]"""
import collections, csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, name = sys.argv[1], sys.argv[2]
frames = float(sys.argv[3]) if len(sys.argv) > 3 and float(sys.argv[3]) > 0 else None
title = sys.argv[4] if len(sys.argv) > 4 else ""
rep = os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % name)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None; cur = None; fname = None
lines = collections.OrderedDict()
S = N = 0
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if len(r) > 10 and r[0] == "Line No": hdr = r; col = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != "":
        try: cur = (fname, int(r[0]), r[1].strip())
        except ValueError: pass
        continue
    if r[2] in ("-", "..."): continue
    num = lambda k: int(r[col[k]]) if k in col and r[col[k]].isdigit() else 0
    a = lines.setdefault(cur, {"samp": 0, "inst": 0, "sass": 0, "wav": 0, "exc": 0, "st": collections.Counter()})
    a["samp"] += num("# Samples"); a["inst"] += num("Instructions Executed"); a["sass"] += 1
    a["wav"] += num("L1 Wavefronts Shared"); a["exc"] += num("L1 Wavefronts Shared Excessive")
    for h, i in col.items():
        if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit(): a["st"][h[6:]] += int(r[i])
    S += num("# Samples"); N += num("Instructions Executed")
path = os.path.join(ROOT, "profiles", "%s_ncu_%s_source.txt" % (tag, name))
with open(path, "w") as f:
    f.write("# ncu source page of profiles/%s_ncu_%s.txt by CUDA source line%s\n" % (tag, name, (" (" + title + ")") if title else ""))
    f.write("total warp-instructions %.4g, stall samples %.4g%s\n" % (N, S, (", warp-instructions per frame %.0f" % (N / frames)) if frames else ""))
    f.write("# (per-instruction counts of the source page: they run 10-30 % above smsp__inst_executed.sum of the raw page; read the shares)\n")
    f.write("%-24s %6s %7s %7s %5s %9s %8s  %-44s %s\n" % ("file:line", "inst%", "samp%", "/frame", "sass", "smem wav", "excess", "top stalls", "source"))
    for k, a in sorted(lines.items(), key=lambda kv: -kv[1]["samp"])[:70]:
        tot = sum(a["st"].values()) or 1
        top = ",".join("%s:%d%%" % (n, round(100 * v / tot)) for n, v in a["st"].most_common(3))
        f.write("%-24s %5.2f%% %6.2f%% %7s %5d %9.3g %8.3g  %-44s %s\n" % (
            "%s:%d" % (k[0], k[1]), 100 * a["inst"] / N, 100 * a["samp"] / S, ("%.1f" % (a["inst"] / frames)) if frames else "-",
            a["sass"], a["wav"], a["exc"], top, k[2][:100]))
print(path)
