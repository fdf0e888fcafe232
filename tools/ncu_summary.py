"""Turn an .ncu-rep (ncu --set full --import-source on) into the compact text summary kept under
profiles/: per-kernel headline metrics + the SASS lines that hold the most stall samples.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_r01.txt ["note"]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = page(rep, "raw")
    H, U = raw[0], raw[1]
    lines = ["# ncu summary of %s" % rep, "# " + note,
             "# captured with: ncu --set full --clock-control none --import-source on (see profiles/README.md)", ""]
    ik = H.index("Kernel Name")
    for n, r in enumerate(raw[2:]):
        lines.append("## launch %d: %s" % (n, r[ik][:110]))
        for k in KEYS:
            if k in H:
                i = H.index(k)
                lines.append("  %-68s %s %s" % (k, r[i], U[i]))
        lines.append("")
    src = page(rep, "source")
    # the source page concatenates kernels: header rows start with "Kernel Name"
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1] if len(r) > 1 else "?", "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    done = set()
    for b in blocks:
        rows = b["rows"]
        if not rows or b["name"] in done or len(done) >= 8:
            continue
        done.add(b["name"])  # one stall table per kernel (repeat launches of one kernel look alike)
        hdr = rows[0]
        try:
            isrc, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        except ValueError:
            continue
        data = []
        for r in rows[1:]:
            if len(r) > max(isrc, isamp, iex):
                try:
                    data.append((int(r[isamp] or 0), int(r[iex] or 0), r[isrc]))
                except ValueError:
                    pass
        tot = sum(d[0] for d in data) or 1
        lines.append("## top stall-sample SASS lines: %s (total samples %d)" % (b["name"][:90], tot))
        for s, e, t in sorted(data, key=lambda d: -d[0])[:14]:
            lines.append("  %6d %5.1f%%  exec=%-9d %s" % (s, 100.0 * s / tot, e, t.strip()[:90]))
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        agg = {}
        for r in rows[1:]:
            for i, h in stall_cols:
                if len(r) > i and r[i].isdigit():
                    agg[h] = agg.get(h, 0) + int(r[i])
        lines.append("  stall reasons: " + ", ".join("%s=%d" % kv for kv in sorted(agg.items(), key=lambda t: -t[1])[:8]))
        lines.append("")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst, len(lines), "lines")


if __name__ == "__main__":
    main()
