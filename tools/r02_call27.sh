#!/bin/bash
# round 2, GPU call 27: planar host output (25 B/ray) -- lossless test, e2e of both host formats in the headline bench
tag=r02aa
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "planar or host_buffer or two_host" 2>&1 | tail -3 > gpurun_out/${tag}_ab.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --census-rays 0 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}.err
python - <<'PY' >> gpurun_out/r02aa_ab.txt
import json
d=json.loads(open('gpurun_out/r02aa_bench.json').read().strip().splitlines()[-1])
e=d['e2e']; print('value', round(d['value']), 'e2e planar', round(e['value'],1), 'pcie_frac', round(e.get('pcie_frac',0),3), 'records', round(e['records_32B']['value'],1), 'pcie_frac', round(e['records_32B'].get('pcie_frac',0),3), e.get('pcie'))
PY
tail -3 gpurun_out/${tag}.err
cat gpurun_out/${tag}_ab.txt
