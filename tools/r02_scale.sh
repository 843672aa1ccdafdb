#!/bin/bash
# round 2: the default bench line at N GPUs of one box, exactly as the driver launches it (strong split + NVLink gather);
# at N = 2 also the two-GPU gather tests.   usage: bash tools/r02_scale.sh <N> <tag>
N=$1; tag=${2:-r02m}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  ( time timeout 600 python -m pytest tests/test_gpu_jobs.py -x -q -m gpu -k "gather" ) > gpurun_out/${tag}_pytest_gather.log 2>&1
  grep -E "passed|failed" gpurun_out/${tag}_pytest_gather.log | tail -1
fi
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 ) \
    > gpurun_out/${tag}_bench_headline_${N}gpu.json 2> gpurun_out/${tag}_bench_headline_${N}gpu.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 2 --warmup 1 ) \
    > gpurun_out/${tag}_bench_reference_${N}gpu.json 2> gpurun_out/${tag}_bench_reference_${N}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_headline_${N}gpu.json").read().strip().splitlines()[-1])
g = d.get("gather") or {}
print("N=$N", round(d["value"]), "Mrays/s", round(d["ms_per_step"], 2), "ms | gather:", {k: (round(v["value"]), v.get("checksum_ok")) for k, v in (g.get("pipelined") or {}).items()},
      "seq", round((g.get("sequential") or {}).get("value", 0)), "errors", g.get("errors"), "| e2e", round(d["e2e"]["value"]), "pcie_frac", round(d["e2e"].get("pcie_frac", 0), 3))
r = open("gpurun_out/${tag}_bench_reference_${N}gpu.json").read().strip().splitlines()
print("reference arm lines:", len(r), json.loads(r[-1])["value"] if r else None)
PY
tail -2 gpurun_out/${tag}_bench_headline_${N}gpu.err
