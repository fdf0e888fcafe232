"""Per-source-line instruction and stall-sample totals of one kernel in an .ncu-rep
(ncu --set full --import-source on, built with -lineinfo).
Usage: python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fn, path, hdr, agg, seen = None, None, None, {}, set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1]
    elif r[0] == "Function Name":
        fn = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and fn and want in fn and r[0].isdigit():
        if (fn, path, r[0]) in seen and False:
            continue
        i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        key = (fn[:60], path.split("/")[-1], int(r[0]), r[1].strip()[:100])
        a = agg.setdefault(key, [0, 0])
        a[0] += int(r[i_inst]) if r[i_inst].isdigit() else 0
        a[1] += int(r[i_samp]) if r[i_samp].isdigit() else 0
fns = sorted({k[0] for k in agg})
for f in fns[:1] if want else fns:
    items = [(k, v) for k, v in agg.items() if k[0] == f]
    ti, ts = sum(v[0] for _, v in items) or 1, sum(v[1] for _, v in items) or 1
    print("## %s  inst=%d samples=%d" % (f, ti, ts))
    for k, v in sorted(items, key=lambda t: -t[1][0])[:top]:
        print("%5.1f%% inst %5.1f%% stall  %s:%d  %s" % (100.0 * v[0] / ti, 100.0 * v[1] / ts, k[1], k[2], k[3]))
