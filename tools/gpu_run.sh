#!/bin/bash
# One entry point for every GPU-side run (replaces the per-round one-off scripts).  Usage on the GPU box:
#   bash tools/gpu_run.sh TAG stage [stage ...]
# stages:
#   tests[:pytest-args]   all (or the given) GPU parity tests -> gpurun_out/pytest_gpu_TAG.log   (stops the script on failure)
#   smoke                 __graft_entry__.smoke()
#   bench[:args]          default bench line -> gpurun_out/bench_TAG.json (args appended, e.g. bench:--workload=cfg3)
#   ref                   bench.py --impl reference
#   ab                    A/B bench lines for each "NAME|ENV|ARGS" entry of $RUNS (';'-separated)
#   launches              ncu launch list (gpu__time_duration) of a 1-step bench -> launches_TAG.csv
#   ncu[:suffix]          ncu --set full of $NCU_K (kernel regex; $NCU_EXTRA e.g. "--kernel-name-base demangled" to match template
#                         arguments) skipping $NCU_S launches, $NCU_C captures, of `python $NCU_CMD`
#                         -> raw CSV (gz) + summary json (+ the .ncu-rep when small)
#   kernels               tools/bench_kernels.py
#   sgb                   tools/bench_sgb.py
#   scale:N               torchrun N ranks, cfg2 then cfg3
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print({k: d.get(k) for k in ("value", "ms_per_step", "conv3_blocks_per_step", "fc1_cells_per_step")}, "e2e", (d.get("e2e") or {}).get("value"),
          d.get("clocks"), {k: round(v["ms_per_step"], 2) for k, v in (d.get("kernel_breakdown") or {}).items()},
          (r.get("kernel") or "")[:40], r.get("achieved"), r.get("frac"), d.get("recall"), d.get("parity_sample"), d.get("cpu_baseline"))
except Exception as e:
    print("no line:", e)
PY
}
for stage in "$@"; do
  name=${stage%%:*}; arg=""; [[ "$stage" == *:* ]] && arg=${stage#*:}
  case $name in
    tests)
      echo "== pytest gpu $arg"; timeout 2400 python -m pytest ${arg:+${arg//,/ }} $([ -z "$arg" ] && echo tests) -m gpu -q --timeout=900 > $OUT/pytest_gpu_$TAG.log 2>&1; rc=$?
      echo "pytest exit $rc"; grep -E "^PARITY|passed|failed|Error|error" $OUT/pytest_gpu_$TAG.log | tail -40
      if [ $rc -ne 0 ]; then tail -${TAIL:-80} $OUT/pytest_gpu_$TAG.log; [ -z "$CONTINUE" ] && exit $rc; fi ;;
    smoke) echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
    bench)
      f=$OUT/bench_${TAG}${arg:+_$(echo $arg | tr -c 'a-zA-Z0-9\n' '_')}.json
      echo "== bench $arg"; timeout 1200 python bench.py ${arg//,/ } > $f 2> ${f%.json}.err; echo "exit $?"; summ $f; tail -3 ${f%.json}.err ;;
    ref) echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "exit $?"; head -c 900 $OUT/bench_ref_$TAG.json ;;
    ab)
      IFS=';' read -ra LIST <<< "$RUNS"
      for item in "${LIST[@]}"; do
        IFS='|' read -r n envs args <<< "$item"
        echo "== bench $n ($envs $args)"
        env $envs timeout 600 python bench.py --steps ${AB_STEPS:-6} --warmup 3 --no-cpu-baseline $args > $OUT/bench_${n}_$TAG.json 2> $OUT/bench_${n}_$TAG.err; echo "exit $?"
        summ $OUT/bench_${n}_$TAG.json; tail -2 $OUT/bench_${n}_$TAG.err
      done ;;
    launches)
      echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_$TAG.csv \
          python bench.py --steps 1 --warmup 1 --no-cpu-baseline ${arg//,/ } > $OUT/ncu_bench_$TAG.log 2>&1; echo "exit $?"; wc -l $OUT/launches_$TAG.csv ;;
    ncu)
      echo "== ncu --set full -k ${NCU_K:-regex:tc_gemm_kernel} -s ${NCU_S:-0} -c ${NCU_C:-4} : python ${NCU_CMD:-bench.py --steps 1 --warmup 1 --no-cpu-baseline}"
      timeout ${NCU_TIMEOUT:-900} ncu --set full --clock-control none --import-source on ${NCU_EXTRA} -k "${NCU_K:-regex:tc_gemm_kernel}" -s ${NCU_S:-0} -c ${NCU_C:-4} -o /tmp/prof_${TAG}${arg} -f \
          python ${NCU_CMD:-bench.py --steps 1 --warmup 1 --no-cpu-baseline} > $OUT/ncu_${TAG}${arg}.log 2>&1; echo "capture exit $?"
      ncu -i /tmp/prof_${TAG}${arg}.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/ncu_raw_${TAG}${arg}.csv.gz
      python tools/ncu_summarize.py /tmp/prof_${TAG}${arg}.ncu-rep > $OUT/ncu_summary_${TAG}${arg}.json 2>/dev/null
      sz=$(stat -c %s /tmp/prof_${TAG}${arg}.ncu-rep 2>/dev/null || echo 0); echo "rep: $sz bytes"
      if [ "$sz" -gt 0 ] && [ "$sz" -lt 20000000 ]; then cp /tmp/prof_${TAG}${arg}.ncu-rep $OUT/; fi
      python - <<PY
import json
try:
    d = json.load(open("$OUT/ncu_summary_${TAG}${arg}.json"))
    for k, v in d.items():
        if isinstance(v, list):
            for r in v:
                print({kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in r.items()})
except Exception as e:
    print("summary failed:", e)
PY
      ;;
    kernels) echo "== per-kernel bench"; timeout 900 python tools/bench_kernels.py > $OUT/kernels_$TAG.json 2> $OUT/kernels_$TAG.err; echo "exit $?"; tail -2 $OUT/kernels_$TAG.err ;;
    sgb) echo "== sgb cfg5"; timeout 600 python tools/bench_sgb.py > $OUT/sgb_cfg5_$TAG.json 2> $OUT/sgb_cfg5_$TAG.err; echo "exit $?"; cat $OUT/sgb_cfg5_$TAG.json; tail -2 $OUT/sgb_cfg5_$TAG.err ;;
    scale)
      N=${arg:-2}; nvidia-smi -L > $OUT/gpus_$TAG.txt
      for WL in ${SCALE_WL:-cfg2 cfg3}; do
        f=$OUT/scale_${TAG}_${WL}_n$N.json
        timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $WL > $f 2> ${f%.json}.err
        echo "$WL n=$N exit $?"; summ $f; tail -2 ${f%.json}.err
      done ;;
    *) echo "unknown stage $stage"; exit 64 ;;
  esac
done
