#!/bin/bash
# ncu --set full of the HBM-bound kernels of the shared-footprint path (pooling with cover, zero fill, bookkeeping), summarised on the box
TAG=${1:-r01L}
OUT=gpurun_out; mkdir -p $OUT
timeout 400 ncu --set full --clock-control none -k regex:"pair_relu_pool_tiled|cells_zero|pair_cover_masks|pair_cell_keys|tile_cell_masks|conv3_blocks_kernel" -s 2 -c 8 -o /tmp/prof_s_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_s_$TAG.log 2>&1
echo "capture exit $?"
ncu -i /tmp/prof_s_$TAG.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/ncu_raw_s_$TAG.csv.gz
python tools/ncu_summarize.py /tmp/prof_s_$TAG.ncu-rep > $OUT/ncu_summary_s_$TAG.json 2>/dev/null
python - <<PY
import json
d = json.load(open("$OUT/ncu_summary_s_$TAG.json"))
for k, v in d.items():
    if isinstance(v, list):
        for r in v:
            print({kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in r.items() if kk in ("kernel", "time_ms", "dram_read_GB", "dram_write_GB", "dram_pct", "l2_hit_pct", "issue_active_pct", "grid")})
PY
