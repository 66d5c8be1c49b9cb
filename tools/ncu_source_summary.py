"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump by CUDA
source line: instructions executed, stall samples, shared-memory excess."""
import csv, sys, collections
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
# duplicate 'Source' header: first is CUDA source text, second SASS
src_cols = [i for i, h in enumerate(hdr) if h == "Source"]
agg = collections.OrderedDict()
def F(x):
    try: return float(x)
    except ValueError: return 0.0
tot_inst = tot_samp = 0
stall_names = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    try: line = int(r[col["Line No"]])
    except ValueError: continue
    inst = F(r[col["Instructions Executed"]])
    samp = F(r[col["# Samples"]])
    exc = F(r[col["L1 Wavefronts Shared Excessive"]])
    wav = F(r[col["L1 Wavefronts Shared"]])
    a = agg.setdefault(line, {"src": r[src_cols[0]].strip()[:90], "inst": 0, "samp": 0, "exc": 0, "wav": 0, "stalls": collections.Counter(), "n": 0})
    a["inst"] += inst; a["samp"] += samp; a["exc"] += exc; a["wav"] += wav; a["n"] += 1
    for s in stall_names:
        v = F(r[col[s]])
        if v: a["stalls"][s] += v
    tot_inst += inst; tot_samp += samp
frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
print("total warp-instructions %.4g, samples %.4g%s" % (tot_inst, tot_samp, ", per frame %.0f" % (tot_inst / frames) if frames else ""))
print("%5s %7s %7s %6s %8s %8s  %-40s %s" % ("line", "inst%", "samp%", "sass", "smemwav", "excess", "top stalls", "source"))
for line, a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]:
    st = ",".join("%s:%.0f%%" % (k.replace("stall_", ""), 100 * v / max(a["samp"], 1)) for k, v in a["stalls"].most_common(3))
    print("%5d %6.2f%% %6.2f%% %6d %8.3g %8.3g  %-40s %s" % (line, 100 * a["inst"] / tot_inst, 100 * a["samp"] / tot_samp, a["n"], a["wav"], a["exc"], st, a["src"]))
