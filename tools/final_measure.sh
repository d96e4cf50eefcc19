#!/bin/bash
# Usage (under gpurun): bash tools/final_measure.sh <tag>
# The round's evidence in one call: a bench line per single-GPU config, the ncu launch list of one c2 step, the
# per-stage DRAM traffic passes, and full-set captures of the two dominant kernels.  Everything lands in gpurun_out/.
tag=${1:-final}
mkdir -p gpurun_out
for c in c2 c3 c5 c2b x_random x_zeros x_issue44; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_${tag}_$c.json 2> gpurun_out/bench_${tag}_$c.err || { echo "$c FAILED"; tail -3 gpurun_out/bench_${tag}_$c.err; continue; }
  python - $c gpurun_out/bench_${tag}_$c.json <<'PY'
import json,sys
d=json.load(open(sys.argv[2]))
print(sys.argv[1], "MiB/s %.0f e2e %.0f ms %.1f ratio %.4f"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["ratio"]), {k: round(v, 2) for k, v in d["stage_ms"].items() if v > 0.25}, "cpu", round(d["cpu_baseline"]["value"],1), d.get("bounded_mode"))
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}_c2_256MiB.csv \
    python bench.py --size-mib 256 --steps 1 --warmup 1 --verify none --cpu-sample-mib 1 > gpurun_out/ncu_l_$tag.log 2>&1
for c in c2 c3; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/traffic_${tag}_$c.csv python bench.py --config $c --size-mib 256 --steps 1 --warmup 1 --verify none --cpu-sample-mib 1 > gpurun_out/traffic_${tag}_$c.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_match|k_parse_spec|k_window_sort" -s 3 -c 3 -o gpurun_out/prof_${tag}_c2 \
    python bench.py --size-mib 512 --steps 1 --warmup 1 --verify none --cpu-sample-mib 1 > gpurun_out/ncu_f_$tag.log 2>&1
echo DONE
