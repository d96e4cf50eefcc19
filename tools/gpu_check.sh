#!/bin/bash
# Usage (under gpurun): bash tools/gpu_check.sh <tag> [tests|notests] [ncu-kernel-regex]
# Runs the GPU parity tests, a 1 GiB bench, an ncu launch list and (optionally) a full-set capture.
tag=${1:-run}; tests=${2:-tests}; kre=${3:-}
mkdir -p gpurun_out
if [ "$tests" = "tests" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; rc=$?
  tail -5 gpurun_out/pytest_$tag.log
  if [ $rc -ne 0 ]; then echo "TESTS FAILED"; exit 1; fi
fi
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { tail -20 gpurun_out/bench_$tag.err; exit 2; }
cat gpurun_out/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --size-mib 256 --steps 1 --warmup 3 --verify none --cpu-sample-mib 8 > gpurun_out/ncu_l_$tag.log 2>&1
if [ -n "$kre" ]; then
  ncu --set full --clock-control none --import-source on -k regex:"$kre" -s 6 -c 2 -o gpurun_out/prof_$tag \
      python bench.py --size-mib 64 --steps 1 --warmup 3 --verify none --cpu-sample-mib 8 > gpurun_out/ncu_f_$tag.log 2>&1
fi
echo DONE
