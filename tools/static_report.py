#!/usr/bin/env python
"""Static evidence from the built library (no GPU): per-kernel registers / spills / static smem (cuobjdump
--dump-resource-usage) and counts of the SASS mnemonics that show which hardware paths a kernel uses
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, SYNCS = mbarrier, UCGABAR / CGA = cluster ...).

    python tools/static_report.py > profiles/r01_static_sass_report.txt
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "voicefixer_b200", "libvfx_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMACCTL", "UTMACMDFLUSH", "SYNCS", "UCGABAR",
         "ELECT", "R2UR", "VOTEU", "HMMA", "FFMA", "MUFU", "LDG", "STG", "LDS", "STS", "STAS", "LDL", "STL", "BAR", "ERRBAR", "MEMBAR"]


def demangle(name):
    m = re.search(r"_cu_[0-9a-f]{8}\d+([a-z0-9_]+_kernel)(I[^E]*E)?", name)
    if not m:
        return name
    kern = m.group(1)
    targ = m.group(2) or ""
    targ = targ.replace("13__nv_bfloat16", "bf16").replace("I", "<", 1).replace("E", ">") if targ else ""
    targ = re.sub(r"L[ij](\d+)", r"\1", targ).replace("<f>", "<f32>")
    return kern + targ


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    for fn, line in re.findall(r"Function (\S+):\s*\n?\s*(REG:[^\n]*)", res):
        usage[fn] = dict(kv.split(":") for kv in re.findall(r"([A-Z]+(?:\[\d\])?:\d+)", line))
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, total, cur = {}, {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur], total[cur] = collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?PT?\d*\s+)?([A-Z0-9_]+)", line)
        if cur and m:
            total[cur] += 1
            op = m.group(1)
            for w in WATCH:
                if op == w or op.startswith(w):
                    counts[cur][w] += 1
                    break
    print("# static report of voicefixer_b200/libvfx_b200.so (sm_100a); tools/static_report.py")
    print("# kernel | regs | stack(B) | local(B) | static smem(B) | SASS instructions | watched mnemonics")
    for fn in sorted(usage, key=demangle):
        u = usage[fn]
        c = counts.get(fn, {})
        watched = " ".join(f"{k}={c[k]}" for k in WATCH if c.get(k))
        print(f"{demangle(fn):42s} | {u.get('REG', '?'):>3s} | {u.get('STACK', '0'):>3s} | {u.get('LOCAL', '0'):>2s} | "
              f"{u.get('SHARED', '0'):>6s} | {total.get(fn, 0):5d} | {watched}")


if __name__ == "__main__":
    sys.exit(main())
