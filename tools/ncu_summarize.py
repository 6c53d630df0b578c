#!/usr/bin/env python
"""Turn `ncu --set full` captures into the small JSON summary committed under profiles/.

  python tools/ncu_summarize.py gpurun_out/prof_dense_TAG.ncu-rep [more.ncu-rep ...] > profiles/ncu_summary_TAG.json

Runs `ncu -i <rep> --page raw --csv` (works without a GPU) and keeps, per captured launch: kernel name, grid/block, registers,
duration, DRAM bytes read/written (their sum is `roofline.traffic` in bench.py), DRAM and L2 figures, issue activity and the
tensor-pipe activity as a fraction of ELAPSED cycles.  ncu replays every kernel cold-cache and serialised: compare shares.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "time_ms",
    "dram__bytes_read.sum": "dram_read_GB",
    "dram__bytes_write.sum": "dram_write_GB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__cycles_active.avg": "cycles_active",
    "sm__cycles_elapsed.max": "cycles_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct_of_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "lts__t_bytes.sum": "l2_bytes_GB",
}
SCALE = {"time_ms": {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3},
         "GB": {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    start = out.find('"ID"')
    rd = list(csv.reader(io.StringIO(out[start:])))
    header, units, body = rd[0], rd[1], rd[2:]
    return header, units, body


def summarize(rep):
    header, units, body = rows_of(rep)
    col = {h: i for i, h in enumerate(header)}
    res = []
    for r in body:
        if len(r) < len(header):
            continue
        d = {"kernel": r[col["Kernel Name"]][:96]}
        for metric, name in KEEP.items():
            if metric not in col:
                continue
            raw = r[col[metric]].replace(",", "")
            try:
                v = float(raw)
            except ValueError:
                continue
            u = units[col[metric]]
            if name == "time_ms":
                v *= SCALE["time_ms"].get(u, 1.0)
            elif name.endswith("_GB"):
                v *= SCALE["GB"].get(u, 1e-9)
            d[name] = v
        if "tensor_pipe_active_pct_of_active" in d and d.get("cycles_elapsed"):
            d["tensor_pipe_active_pct_of_elapsed"] = d["tensor_pipe_active_pct_of_active"] * d["cycles_active"] / d["cycles_elapsed"]
        if "dram_read_GB" in d and "dram_write_GB" in d:
            d["dram_bytes_per_launch"] = (d["dram_read_GB"] + d["dram_write_GB"]) * 1e9
        res.append(d)
    return res


def main():
    out = {"source": "ncu --set full --clock-control none; summarised by tools/ncu_summarize.py (cold-cache, serialised launches: compare shares)"}
    for rep in sys.argv[1:]:
        key = rep.split("/")[-1].replace(".ncu-rep", "")
        try:
            out[key] = summarize(rep)
        except Exception as e:                       # keep going: one unreadable capture must not lose the others
            out[key] = {"error": repr(e)}
    json.dump(out, sys.stdout, indent=1)
    sys.stdout.write("\n")


if __name__ == "__main__":
    main()
