import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        if "REFINE" in d["config"]: continue
        print("   %-46s %8.1f M  refined %s" % (d["config"][:46], d["frames_per_s"] / 1e6, d.get("refined")))
