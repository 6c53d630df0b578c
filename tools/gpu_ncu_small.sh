#!/bin/bash
# ncu --set full of the tensor-core launches of ONE pass of the default path, exported as gzipped raw CSV (small enough to come back):
# launch order of tc_gemm_kernel in a fresh process: conv1, conv2 x2, background (conv2 x2, conv3), conv3_box, fc1_box, conv3 x chunks, fc1, fc2
TAG=${1:-r01y}
OUT=gpurun_out; mkdir -p $OUT
ARGS=${BENCH_ARGS:---steps 1 --warmup 1 --no-cpu-baseline}
timeout 600 ncu --set full --clock-control none -k regex:tc_gemm_kernel -s 6 -c 4 -o /tmp/prof_a_$TAG -f python bench.py $ARGS > $OUT/ncu_a_$TAG.log 2>&1
echo "capture a exit $?"
timeout 600 ncu --set full --clock-control none -k regex:tc_gemm_kernel -s ${FC1_SKIP:-16} -c 2 -o /tmp/prof_b_$TAG -f python bench.py $ARGS > $OUT/ncu_b_$TAG.log 2>&1
echo "capture b exit $?"
for x in a b; do
  ncu -i /tmp/prof_${x}_$TAG.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/ncu_raw_${x}_$TAG.csv.gz
  python tools/ncu_summarize.py /tmp/prof_${x}_$TAG.ncu-rep > $OUT/ncu_summary_${x}_$TAG.json 2>/dev/null
  sz=$(stat -c %s /tmp/prof_${x}_$TAG.ncu-rep 2>/dev/null || echo 0)
  echo "rep $x: $sz bytes"
  if [ "$sz" -gt 0 ] && [ "$sz" -lt 15000000 ]; then cp /tmp/prof_${x}_$TAG.ncu-rep $OUT/; fi
done
ls -la $OUT | tail -12
python - <<PY
import json
for x in "ab":
    try:
        d = json.load(open("$OUT/ncu_summary_%s_$TAG.json" % x))
        for k, v in d.items():
            if isinstance(v, list):
                for r in v:
                    print(x, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in r.items() if kk in ("kernel", "time_ms", "dram_read_GB", "dram_write_GB", "l2_hit_pct", "tensor_pipe_active_pct_of_elapsed", "l2_bytes_GB", "grid")})
    except Exception as e:
        print("summary", x, "failed:", e)
PY
