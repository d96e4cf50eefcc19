#!/bin/bash
# Usage (under gpurun --gpus N): bash tools/multi_gpu.sh N tag [config ...]
# One bench line per config at N GPUs (torchrun, one rank per GPU), kept under gpurun_out/.
N=$1; tag=$2; shift 2
cfgs=${@:-c4 c2}
mkdir -p gpurun_out
for c in $cfgs; do
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --config $c --steps 3 --warmup 3 --cpu-sample-mib 4 > gpurun_out/bench_${tag}_${c}_n$N.json 2> gpurun_out/bench_${tag}_${c}_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $N --config $c --steps 3 --warmup 3 --cpu-sample-mib 4 > gpurun_out/bench_${tag}_${c}_n$N.json 2> gpurun_out/bench_${tag}_${c}_n$N.err
  fi
  python - gpurun_out/bench_${tag}_${c}_n$N.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["name"], "N", d["n_gpus"], "MiB/s %.0f e2e %.0f ms %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
    for r in d.get("per_rank", []):
        print("   rank", r["rank"], "ms %.1f" % r["ms_per_step"], {k: round(v, 1) for k, v in r.get("stage_ms", {}).items() if v > 1.0}, r.get("chunks", ""))
except Exception as e:
    print("FAILED", sys.argv[1], e)
    print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
