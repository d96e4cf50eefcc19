#!/bin/bash
# Usage (under gpurun): bash tools/gpu_iter.sh <tag> [config ...]
# Inner-loop check of a kernel change: the parity tests that exercise the LZ77 stage most, then a 1 GiB bench per
# config (default c2), once for the in-tree library and once per variant under build/variants/ (tools/tune_variants.sh build).
tag=${1:-it}; shift
cfgs=${@:-c2}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "oneshot_bit_exact or afl or pinned_sizes or custom_options or match_stage or every_preset or pieces_concatenate or writer_pieces_without or lz77_stage" > gpurun_out/pytest_$tag.log 2>&1; rc=$?
tail -3 gpurun_out/pytest_$tag.log
if [ $rc -ne 0 ]; then echo "TESTS FAILED"; exit 1; fi
for lib in intree build/variants/*.so; do
  [ "$lib" = intree ] || [ -f "$lib" ] || continue
  for c in $cfgs; do
    name=$(basename $lib .so)
    if [ "$lib" = intree ]; then unset DFL_LIB_PATH; else export DFL_LIB_PATH=$PWD/$lib; fi
    timeout 200 python bench.py --config $c --steps 3 --warmup 2 --cpu-sample-mib 1 > gpurun_out/bench_${tag}_${name}_$c.json 2> gpurun_out/bench_${tag}_${name}_$c.err || { echo "$name $c FAILED"; tail -3 gpurun_out/bench_${tag}_${name}_$c.err; continue; }
    python - "$name" "$c" gpurun_out/bench_${tag}_${name}_$c.json <<'PY'
import json,sys
d=json.load(open(sys.argv[3]))
print(sys.argv[1], sys.argv[2], "MiB/s %.0f e2e %.0f ms %.1f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(v, 2) for k, v in d["stage_ms"].items() if v > 0.4})
PY
  done
done
echo DONE
