#!/bin/bash
# Round-end evidence session on one B200: sanitizers, the GPU parity suite, smoke, the bench lines of every BASELINE config
# (with the CPU reference leg), launch lists and one --set full capture of the dominant kernel.
TAG=${1:-r02}
mkdir -p gpurun_out
for T in racecheck memcheck; do
  timeout 330 compute-sanitizer --tool $T --print-limit 5 python tools/sanitize.py > gpurun_out/${TAG}_sanitizer_$T.log 2>&1
  echo "== $T: $(grep -c 'Race reported' gpurun_out/${TAG}_sanitizer_$T.log) race reports; $(grep 'SUMMARY' gpurun_out/${TAG}_sanitizer_$T.log | tail -1)"
done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/${TAG}_smoke.log
for SPEC in c2:100 c3:20 c5:8 c4:40; do
  WL=${SPEC%%:*}; ST=${SPEC##*:}
  timeout 600 python bench.py --workload $WL --steps $ST --warmup 5 2> gpurun_out/${TAG}_bench_${WL}.err > gpurun_out/${TAG}_bench_${WL}.json
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${WL}.json").read())
    cb = d.get("cpu_baseline") or {}
    print("[$WL] value", round(d["value"], 1), d["unit"], "| fps", round(d["fps"], 2), "| serial fps", round(d["serial"]["fps"], 2), "kernel alone ms", round(d["serial"]["roofline"]["kernel_ms"], 4),
          "frac", round(d["serial"]["roofline"]["frac"], 3), "| e2e fps", round(d["e2e"]["fps"], 2), "| cpu ref fps", cb.get("fps"), "on", cb.get("cores"), "cores", "| clk", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
except Exception as e:
    print("[$WL] failed:", e); print(open("gpurun_out/${TAG}_bench_${WL}.err").read()[-1500:])
PY
done
for WL in c2 c3 c4 c5; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_${WL}.csv \
      python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rt_pool_kernel -s 6 -c 1 -f -o gpurun_out/${TAG}_prof_c2 \
    python bench.py --workload c2 --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_ncu_c2.err
ls -la gpurun_out/${TAG}_prof_c2.ncu-rep
