#!/bin/bash
# Usage (under gpurun): bash tools/gpu_quick.sh <tag> [ncu-kernel-regex]
# Fast inner-loop check: the core parity tests, a 256 MiB bench, optionally a full ncu capture at 64 MiB.
tag=${1:-q}; kre=${2:-}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; rc=$?
tail -5 gpurun_out/pytest_$tag.log
if [ $rc -ne 0 ]; then echo "TESTS FAILED"; exit 1; fi
timeout 600 python bench.py --size-mib 256 --steps 3 --warmup 3 --cpu-sample-mib 8 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { tail -20 gpurun_out/bench_$tag.err; exit 2; }
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("MiB/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "stage_ms", {k: round(v,2) for k,v in d["stage_ms"].items()})
PY
if [ -n "$kre" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$kre" -s 6 -c 2 -o gpurun_out/prof_$tag \
      python bench.py --size-mib 64 --steps 1 --warmup 3 --verify none --cpu-sample-mib 8 > gpurun_out/ncu_f_$tag.log 2>&1
fi
echo DONE
