#!/bin/bash
# round 2, GPU call 32: compute-sanitizer (memcheck, racecheck) on the thin-lens retry kernel as shipped (byte-wide tables,
# 10-byte row tables in shared memory, 16-bit path for the wide image) and on the planar pack kernel
tag=r02ai
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest -x -q -m gpu tests/test_gpu_parity.py -k "guarded_thin_lens_is_bit_exact or thin_lens_hex_bokeh" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Error|hazard" | head -6 > gpurun_out/${tag}_sanitizer_thin_$tool.txt
  echo "== $tool"; cat gpurun_out/${tag}_sanitizer_thin_$tool.txt
done
