#!/bin/bash
# Build libdeflate_b200 variants with different compile-time knobs (here, no GPU needed) ...
#   bash tools/tune_variants.sh build "base:" "ns:-DDFL_PARSE_STRIDED=0" "lc256:-DDFL_LC_BYTES=256 -DDFL_PARSE_CTAS=5"
# ... and time them on a GPU box (TUNE_MIB = input size, TUNE_CONFIG = bench config):
#   gpurun -- 'bash tools/tune_variants.sh run'
set -e
cd "$(dirname "$0")/.."
mode=$1; shift
mkdir -p build/variants gpurun_out
if [ "$mode" = build ]; then
  rm -f build/variants/*.so
  for v in "$@"; do
    tag=${v%%:*}; flags=${v#*:}
    (cd deflate-rs_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $flags \
        -o ../../build/variants/lib_$tag.so dfl_kernels.cu dfl_api.cu dfl_comm.cu -ldl) &
  done
  wait
  ls -la build/variants
else
  for lib in build/variants/*.so; do
    DFL_LIB_PATH=$PWD/$lib timeout 300 python bench.py --config ${TUNE_CONFIG:-c2} --size-mib ${TUNE_MIB:-1024} --steps 3 --warmup 2 --cpu-sample-mib 1 --verify prefix > gpurun_out/tune.json 2> gpurun_out/tune.err || { echo "$lib FAILED"; tail -3 gpurun_out/tune.err; continue; }
    python - "$lib" <<'PY'
import json,sys
d=json.load(open("gpurun_out/tune.json"))
print(sys.argv[1], "MiB/s %.0f e2e %.0f"%(d["value"], d["e2e"]["value"]), {k: round(v, 2) for k, v in d["stage_ms"].items() if v > 0.4})
PY
  done
fi
