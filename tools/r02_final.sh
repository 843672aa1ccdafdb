#!/bin/bash
# round 2, final single-GPU pass: what the driver runs at round end (GPU tests, smoke, default bench, reference arm) and the
# evidence behind profiles/r02_* (tools/r02_evidence.sh)
tag=${1:-r02k}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${tag}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
bash tools/r02_evidence.sh ${tag}
