#!/bin/bash
# compute-sanitizer over a trimmed GPU test selection (memcheck, racecheck, synccheck, initcheck).
# usage (on the GPU box): bash tools/sanitize.sh > gpurun_out/sanitizer.txt
SEL="tests/test_gpu_parity.py::test_gpu_reproduces_reference_golden_vectors tests/test_gpu_parity.py::test_guarded_kolb_bokeh_image_sizes tests/test_gpu_parity.py::test_guarded_thin_lens_is_bit_exact tests/test_gpu_parity.py::test_edge_samples_follow_the_rulings"
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== compute-sanitizer --tool $tool"
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $SEL -x -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Race|hazard|Uninitialized" | head -8
done
