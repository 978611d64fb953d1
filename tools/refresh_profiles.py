"""Turns the files tools/capture_r02.sh brought back in gpurun_out/ into the committed evidence under profiles/:
ncu summaries, launch lists + per-kernel summaries, and profiles/ncu_traffic.json (stamped with the hash of the sources).
Runs here (no GPU): python tools/refresh_profiles.py"""
import collections, csv, io, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
INCR = len(sys.argv) > 1 and sys.argv[1] == "b"      # only what tools/capture_r02b.sh brought back (conv_ts kernel + tf32 launch list)
for n in (("ts_c128_tf32",) if INCR else ("pair_c64", "pair2_c128", "rs2", "unet_c32", "gru", "voc_post", "rs3_tf32", "pair3_tf32")):
    subprocess.run([sys.executable, "tools/ncu_summary.py", f"gpurun_out/r02_{n}.ncu-rep", f"profiles/r02_{n}.txt",
                    f"round 2, captured from the sources of commit {head}, B=32 (tools/capture_r02.sh)"], capture_output=True)
for P in (("tf32",) if INCR else ("bf16", "tf32")):
    rows = list(csv.reader(open(f"gpurun_out/r02_launches_{P}.csv")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 2:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    half = data[len(data) // 2:]
    agg, tot = collections.OrderedDict(), 0.0
    for r in half:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("unnamed>::", "")
        ms = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-3) * float(r[iv].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
    out = io.StringIO()
    out.write(f"# one restore() step, B = 32 x 10 s, precision {P}: ncu --metrics gpu__time_duration.sum --clock-control none (second of two steps)\n")
    out.write(f"# launches {len(half)}, summed kernel time {tot:.2f} ms (serialised, cold-cache: compare SHARES with bench.py's breakdown_ms, not absolutes)\n")
    out.write("kernel,launches,total_ms,share\n")
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"{k},{c},{ms:.3f},{ms / tot:.4f}\n")
    open(f"profiles/r02_launch_list_summary_{P}.csv", "w").write(out.getvalue())
    shutil.copy(f"gpurun_out/r02_launches_{P}.csv", f"profiles/r02_launch_list_{P}.csv")
if INCR:
    sys.exit(0)
if os.path.exists("profiles/ncu_traffic.json"):
    os.remove("profiles/ncu_traffic.json")
subprocess.run([sys.executable, "tools/update_traffic.py", "gpurun_out/r02_pair_c64.ncu-rep", "bf16:voc.rs3.pair:B32", "1",
                "profiles/r02_pair_c64.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per fused pair launch)",
                "resstack_pair_tc.cu,tc_ptx.cuh,vfx_common.cuh"], check=True)
subprocess.run([sys.executable, "tools/update_traffic.py", "gpurun_out/r02_pair3_tf32.ncu-rep", "tf32:voc.rs3.pair:B32", "1",
                "profiles/r02_pair3_tf32.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per fused tf32 pair launch)",
                "resstack_pair3_tc.cu,tc_ptx.cuh,vfx_common.cuh"], check=True)
subprocess.run([sys.executable, "tools/update_traffic.py", "gpurun_out/r02_pair2_c128.ncu-rep", "bf16:voc.rs2.pair:B32", "1",
                "profiles/r02_pair2_c128.txt (ncu --set full): one launch of the two-CTA pipelined C = 128 pair",
                "resstack_pair2_tc.cu,tc_ptx.cuh,vfx_common.cuh"], check=True)
print(open("profiles/ncu_traffic.json").read()[:600])
