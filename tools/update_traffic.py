"""Records the DRAM traffic of a ResStack-pair capture in profiles/ncu_traffic.json, stamped with the hash of the CUDA sources
it was taken from (bench.py quotes `roofline.traffic` only when that hash matches the sources it runs).
usage: python tools/update_traffic.py <rep.ncu-rep> <key e.g. bf16:voc.rs3.pair:B32> <launches per pair: 1 fused / 2 unfused> <source note> <csrc files,comma separated>"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import csrc_sha

rep, key, per_pair, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
files = sorted(sys.argv[5].split(","))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
vals = [float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]] for r in rows[2:]]
pick = [int(i) for i in sys.argv[6].split(",")] if len(sys.argv) > 6 else None      # explicit launch indices making up ONE pair
launches = [vals[i] for i in pick] if pick else (vals[:per_pair] if per_pair > 1 else vals)
per = sum(launches) if (pick or per_pair > 1) else sum(launches) / len(launches)
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d[key] = {"dram_bytes_per_pair": per, "launches": [round(v) for v in vals], "files": files, "csrc_sha": csrc_sha(files), "source": note}
json.dump(d, open(path, "w"), indent=1)
print(key, per / 1e9, "GB per pair")
