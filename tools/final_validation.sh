#!/bin/bash
# Round-end validation on one B200 box: GPU tests, smoke, bench line, reference arm, launch list of the bench command.
#   gpurun --timeout 1500 -- 'bash tools/final_validation.sh r2_final'
tag=${1:-final}
o=gpurun_out
python -m pytest tests -m gpu -q > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $o/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $o/${tag}_smoke.log
python bench.py > $o/${tag}_bench_line.json 2> $o/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_reference_line.json 2> $o/${tag}_reference.err; echo "reference rc=$?"; cut -c1-200 $o/${tag}_reference_line.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > $o/${tag}_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_bench_launches_coldcache.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > $o/${tag}_ncu_bench.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json
r=json.loads(open("$o/${tag}_bench_line.json").read().strip().splitlines()[-1])
print("value",round(r["value"],1),"ms",round(r["ms_per_step"],4),"e2e",round(r["e2e"]["value"],1),"roofline",round(r["roofline"]["frac"],3),"whole",round(r["roofline"]["whole_path"]["frac"],3),"sum",round(r["roofline"]["whole_path_sum_of_kernels"]["frac"],3),"limiter",r["roofline"]["limiter"]["family"],"cpu",r.get("cpu_baseline",{}).get("value"),"launches",r.get("gpu_launches"),r.get("clocks"))
PY
