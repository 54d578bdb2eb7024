#!/bin/bash
# Round 2, last GPU session (5.8 GPU-minutes left): the parity suite, smoke, the default bench line, ncu evidence of the default
# (4-warp CTA) kernel; then, as far as the budget goes, the other BASELINE configs.
TAG=r03
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/${TAG}_smoke.log
run_bench() {
  WL=$1; ST=$2; shift; shift
  timeout 300 python bench.py --workload $WL --steps $ST --warmup 5 "$@" 2> gpurun_out/${TAG}_bench_${WL}.err > gpurun_out/${TAG}_bench_${WL}.json
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
}
run_bench c2 100
timeout 120 ncu --set full --clock-control none --import-source on -k regex:rt_pool_kernel -s 12 -c 1 -f -o gpurun_out/${TAG}_prof_c2 \
    python bench.py --workload c2 --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_ncu_c2.err
ls -la gpurun_out/${TAG}_prof_c2.ncu-rep
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_c2.csv \
    python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
run_bench c4 40 --no-cpu-baseline
run_bench c3 20 --no-cpu-baseline
run_bench c5 8 --no-cpu-baseline
