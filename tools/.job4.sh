timeout 1300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2w.log 2>&1; tail -4 gpurun_out/pytest_r2w.log
for c in c4 c2 c3 c5 x_issue44 x_zeros; do
  timeout 300 python bench.py --config $c --steps 3 --warmup 3 --cpu-sample-mib 4 > gpurun_out/bench_r2w_$c.json 2> gpurun_out/bench_r2w_$c.err || { echo "$c FAILED"; tail -3 gpurun_out/bench_r2w_$c.err; continue; }
  python - $c gpurun_out/bench_r2w_$c.json <<'PY'
import json,sys
d=json.load(open(sys.argv[2]))
print(sys.argv[1], "MiB/s %.0f e2e %.0f ms %.1f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(v, 2) for k, v in d.get("stage_ms",{}).items() if v > 0.4})
PY
done
python tools/latency.py 2>&1 | tail -6 | cut -c1-110
