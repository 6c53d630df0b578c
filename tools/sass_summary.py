#!/usr/bin/env python
"""ISA evidence for the built library: per kernel, the counts of the Blackwell-only instructions in its SASS
(`cuobjdump -sass`) and its resource usage (`cuobjdump -res-usage`).  Runs without a GPU.

  python tools/sass_summary.py > profiles/sass_summary.json

UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA load; .MULTICAST =
multicast::cluster), UTMASTG = TMA store, SYNCS = mbarrier ops, UCGABAR = cluster barrier, HMMA / IMMA would be legacy mma.sync.
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "scene_graph_commonsense_b200", "libhiercom_b200.so")
WATCH = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "UCGABAR", "HMMA", "IMMA", "HADD2", "VHMNMX", "HMNMX2",
         "LDG.E.128", "STG.E.128", "LDS.128", "SHFL", "REDUX", "ATOM", "RED")


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["instructions"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                cur[w] += 1
        if op.startswith("UTMALDG"):
            cur[".".join(op.split(".")[:3]) if "MULTICAST" in op else ".".join(op.split(".")[:2])] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            cur["UTCHMMA.2CTA"] += 1
    usage = {}
    name = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        if name and "REG:" in line:
            usage[name] = {k.lower(): int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL|CONSTANT\[0\]):(\d+)", line)}
            name = None
    dm = demangle(list(kernels))
    out = {"library": os.path.relpath(LIB, ROOT), "arch": sorted(set(re.findall(r"arch = (sm_\w+)", sass))), "kernels": []}
    total = collections.Counter()
    for k, c in kernels.items():
        short = re.sub(r"\(.*", "", dm.get(k, k)).replace("void ", "")
        row = {"kernel": short, "instructions": c["instructions"]}
        row.update({w: c[w] for w in sorted(c) if w != "instructions" and c[w]})
        row.update(usage.get(k, {}))
        out["kernels"].append(row)
        total.update({w: v for w, v in c.items() if w != "instructions"})
    out["totals"] = {w: total[w] for w in sorted(total) if total[w]}
    out["kernel_count"] = len(kernels)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
