"""Aggregate the LAST backbone step of an ncu launch list (tools/run_step.py marks it with a
fill launch of 1234.0) into per-kernel time shares.
Usage: python tools/launch_summary.py launches.csv out_summary.csv"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
H = rows[0]
ik, im, iv = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
seq = [(r[ik], float(r[iv].replace(",", ""))) for r in rows[1:] if r[im] == "gpu__time_duration.sum"]
unit = next(r[H.index("Metric Unit")] for r in rows[1:] if r[im] == "gpu__time_duration.sum")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
# the marker is the last elementwise fill before the final step
fps = [i for i, (k, _) in enumerate(seq) if "fps_kernel" in k or "fps_bucket_kernel" in k]
first_fps = fps[-4]  # four FPS launches per forward (SA1..SA4)
mark = max(i for i, (k, _) in enumerate(seq[:first_fps]) if "FillFunctor" in k)
last = seq[mark + 1:]
agg = {}
for k, v in last:
    name = k.split("(")[0].replace("void ", "").replace("pn2::<unnamed>::", "").replace("unnamed>::", "")[:80]
    a = agg.setdefault(name, [0.0, 0])
    a[0] += v * scale
    a[1] += 1
tot = sum(a[0] for a in agg.values())
with open(sys.argv[2], "w") as f:
    f.write("# ncu launch list, round 1 -- last of 4 backbone forward steps (8 scenes x 40k pts, C_in=1)\n")
    f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv python tools/run_step.py 4\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's roofline_ops, not absolutes\n")
    f.write("# %d launches, %.1f us total\n" % (len(last), tot))
    f.write("us,share_pct,launches,kernel\n")
    for k, (v, c) in sorted(agg.items(), key=lambda t: -t[1][0]):
        f.write("%.1f,%.1f,%d,%s\n" % (v, 100 * v / tot, c, k))
print(open(sys.argv[2]).read())
