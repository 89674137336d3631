#!/bin/bash
# round-2 evidence run (one GPU): smoke, fp64 probe, config rows with clocks, full GPU suite, both bench arms
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.txt 2>&1; tail -3 gpurun_out/smoke_r2.txt
benchmarks/build/fp64_probe > gpurun_out/fp64_probe_r2.jsonl 2>&1; cat gpurun_out/fp64_probe_r2.jsonl
python benchmarks/bench_configs.py > gpurun_out/configs_r2.jsonl 2> gpurun_out/configs_r2.err
tail -5 gpurun_out/configs_r2.err
python - <<'PY'
import json
for l in open("gpurun_out/configs_r2.jsonl"):
    try: d = json.loads(l)
    except Exception: print("BAD", l[:200]); continue
    print(round(d.get("ms", 0), 3), d.get("frac_of_measured_peak"), d.get("frac_of_measured_fp64_rate"), (d.get("clocks") or {}).get("sm_mhz"), d["config"][:90])
PY
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_r2.txt; cat gpurun_out/pytest_gpu_r2.txt
python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -3 gpurun_out/bench_r2_n1.err; cat gpurun_out/bench_r2_n1.json
python bench.py --impl reference > gpurun_out/bench_r2_n1_reference.json 2>> gpurun_out/bench_r2_n1.err; cat gpurun_out/bench_r2_n1_reference.json
