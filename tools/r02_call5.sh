#!/bin/bash
# round 2, GPU call 5 (EIGHT GPUs): the strong-split headline job with the NVLink gather, config 4 and the config 5 sweep.
tag=r02e
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
run() {  # name, extra args...
  name=$1; shift
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 "$@" ) \
      > gpurun_out/${tag}_bench_${name}_8gpu.json 2> gpurun_out/${tag}_bench_${name}_8gpu.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_${name}_8gpu.json").read().strip().splitlines()[-1])
    g = d.get("gather") or {}
    print("${name}", round(d["value"]), "Mrays/s", round(d["ms_per_step"], 2), "ms | gather:",
          {k: round(v["value"]) for k, v in (g.get("pipelined") or {}).items()}, "seq", round((g.get("sequential") or {}).get("value", 0)),
          "errors", g.get("errors"), "| e2e", round((d.get("e2e") or {}).get("value", 0)))
except Exception as e:
    print("${name} FAILED", e)
PY
  tail -2 gpurun_out/${tag}_bench_${name}_8gpu.err
}
run headline --steps 5 --warmup 3
run config4 --workload config4 --steps 2 --warmup 1 --no-e2e --transports fused,push
for lens in double_gauss_f2.0.dat fisheye_muller_f4.0.dat mori_f2.8.dat petzval_f1.25.dat petzval_f1.6.dat telephoto_f5.0.dat tessar_f2.8.dat triplet_f2.5.dat; do
  run config5_${lens%.dat} --workload config5:$lens --steps 2 --warmup 1 --no-e2e --transports fused
done
