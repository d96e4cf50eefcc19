#!/bin/bash
# Build libdeflate_b200 variants with different chain-kernel constants (here, no GPU needed) ...
#   bash tools/tune_variants.sh build "W=24,S=4,P=8,C=1024" "W=24,S=2,P=8,C=1024" ...
# ... and time them on a GPU box:
#   gpurun -- 'bash tools/tune_variants.sh run'
set -e
cd "$(dirname "$0")/.."
mode=$1; shift
mkdir -p build/variants gpurun_out
if [ "$mode" = build ]; then
  for v in "$@"; do
    W=24; S=4; P=8; C=1024; M=3; PATHSEL=walk; PS=4096; PW=512; U=2; X=""; TAG=base
    eval "$(echo "$v" | tr ',' ';')"
    out=build/variants/lib_${PATHSEL}_W${W}_S${S}_P${P}_C${C}_M${M}_PS${PS}_PW${PW}_U${U}_${TAG}.so
    (cd deflate-rs_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
        -DDFL_CHAIN_WARPS=$W -DDFL_CHAIN_DEEP_STEPS=$S -DDFL_CHAIN_PARK_MIN=$P -DDFL_CHAIN_CHUNK=$C -DDFL_MATCH_CTAS=$M -DDFL_PARSE_SEG=$PS -DDFL_PARSE_WARM=$PW -DDFL_WALK_UNROLL=$U $X \
        -o ../../$out dfl_kernels.cu dfl_api.cu) &
  done
  wait
  ls -la build/variants
else
  for lib in build/variants/*.so; do
    case "$lib" in *lib_walk*) export DFL_MATCH_PATH=walk;; *) export DFL_MATCH_PATH=chains;; esac
    DFL_LIB_PATH=$PWD/$lib timeout 300 python bench.py --size-mib ${TUNE_MIB:-256} --steps 3 --warmup 2 --cpu-sample-mib 1 --verify prefix > gpurun_out/tune.json 2> gpurun_out/tune.err || { echo "$lib FAILED"; tail -3 gpurun_out/tune.err; continue; }
    python - "$lib" <<'PY'
import json,sys
d=json.load(open("gpurun_out/tune.json"))
print(sys.argv[1], "match_ms %.2f sort_ms %.2f parse_ms %.2f total MiB/s %.0f e2e %.0f"%(d["stage_ms"]["match"], d["stage_ms"]["window_sort"], d["stage_ms"]["parse"], d["value"], d["e2e"]["value"]))
PY
  done
fi
