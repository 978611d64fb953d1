"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/name.txt ["note"]"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none summary of {rep}", f"# {note}"]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"kernel: {d.get('Kernel Name', '?')}  (launch id {d.get('ID', '?')})")
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:86s} {d[k]:>16s} {units[hdr.index(k)]}")
        try:
            rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
            u = units[hdr.index("dram__bytes_read.sum")]
            lines.append(f"  traffic (dram read+write) = {rd + wr:.3f} {u}")
        except Exception:
            pass
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        hdr = rows[1]
        ci = {h: i for i, h in enumerate(hdr)}
        data = []
        for r in rows[2:]:
            try:
                data.append((int(r[ci["# Samples"]]), r))
            except Exception:
                pass
        tot = sum(s for s, _ in data) or 1
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(int(r[ci[h]] or 0) for _, r in data) for h in stalls}
        lines.append(f"stall samples (all instructions, {tot} samples, {len(data)} SASS instructions):")
        for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
            lines.append(f"  {h:28s} {100.0 * v / tot:5.1f} %")
        lines.append("hottest SASS instructions:")
        for s, r in sorted(data, key=lambda x: -x[0])[:12]:
            lines.append(f"  {100.0 * s / tot:5.1f} %  {r[ci['Source']].strip()[:90]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
