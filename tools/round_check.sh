#!/bin/bash
# What the driver runs at round end on one B200: GPU tests, smoke, the default bench line and the reference arm.
# usage (on the GPU box): bash tools/round_check.sh <tag>
tag=${1:-r02z}
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
( time python bench.py ) > gpurun_out/${tag}_bench_headline.json 2> gpurun_out/${tag}_bench_headline.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_headline.json").read().strip().splitlines()[-1])
print(round(d["value"]), "Mrays/s", round(d["ms_per_step"], 2), "ms frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"], "|", d["roofline"]["traffic_source"],
      "| e2e", round(d["e2e"]["value"]), "pcie_frac", round(d["e2e"]["pcie_frac"], 3), "| census flips", d["parity_census"]["flips"], "| cpu", round(d["cpu_baseline"]["value"], 2), "launches", d["gpu_launches"])
PY
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
head -c 200 gpurun_out/${tag}_bench_reference.json; echo
