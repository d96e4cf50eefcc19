#!/usr/bin/env python3
"""Turns an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` log of one
bench.py step into the per-stage table bench.py's `roofline.traffic` reads (profiles/rNN_dram_traffic.json).

  gpurun -- 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
             --csv --log-file gpurun_out/traffic.csv python bench.py --size-mib 256 --steps 1 --warmup 1 --verify none'
  python tools/ncu_traffic.py gpurun_out/traffic.csv 256 c2 profiles/r02_dram_traffic.json

Only the launches of the LAST pipeline run in the log are counted (the warm-up steps launch the same kernels).
"""
import csv
import json
import re
import sys

# kernel name -> pipeline stage (dfl_api.cu issue_pipeline marks)
STAGE_OF = [
    (r"k_window_sort", "window_sort"), (r"k_match", "match"), (r"k_span_scatter", "match"),
    (r"k_parse|k_reset_bad|k_chain_predict|k_lz77_seq", "parse"), (r"k_seg_scan|k_compact", "token_layout"),
    (r"k_block_stats|k_set_tokens", "block_stats"), (r"k_block_codes", "block_codes"), (r"k_block_scan", "block_scan"),
    (r"k_pack", "pack"), (r"k_adler32", "adler32"), (r"k_crc32", "crc32"), (r"k_finalize", "finalize"),
]


def stage_of(kernel):
    for pat, st in STAGE_OF:
        if re.search(pat, kernel):
            return st
    return None


def main():
    log, size_mib, cfg, out = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    rows = [r for r in csv.reader(open(log, errors="replace")) if len(r) >= 15 and r[0].isdigit()]
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r[0]), {"kernel": r[4]})
        d[r[12]] = float(r[14].replace(",", ""))
        d["unit_" + r[12]] = r[13]
    ids = sorted(launches)
    # the last pipeline run starts at the last k_window_sort (or k_lz77_seq) launch that follows a k_finalize
    start = ids[0]
    prev_final = False
    for i in ids:
        k = launches[i]["kernel"]
        if prev_final and stage_of(k) in ("window_sort", "parse"):
            start = i
        prev_final = "k_finalize" in k
    n = size_mib << 20
    stages = {}
    for i in ids:
        if i < start:
            continue
        L = launches[i]
        st = stage_of(L["kernel"])
        if st is None:
            continue
        s = stages.setdefault(st, {"launches": 0, "ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        s["launches"] += 1
        s["ms"] += L.get("gpu__time_duration.sum", 0.0) / 1e6
        s["dram_read_bytes"] += L.get("dram__bytes_read.sum", 0.0) * scale.get(L.get("unit_dram__bytes_read.sum", "byte"), 1)
        s["dram_write_bytes"] += L.get("dram__bytes_write.sum", 0.0) * scale.get(L.get("unit_dram__bytes_write.sum", "byte"), 1)
    tot = 0.0
    for st, s in stages.items():
        s["dram_bytes_per_input_byte"] = (s["dram_read_bytes"] + s["dram_write_bytes"]) / n
        tot += s["dram_bytes_per_input_byte"]
    doc = {"config": cfg, "size_mib": size_mib, "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
           "--clock-control none, one bench.py step (times are cold-cache and serialised: shares, not absolutes)",
           "stages": stages, "total_dram_bytes_per_input_byte": tot}
    json.dump(doc, open(out, "w"), indent=1)
    print(json.dumps({k: round(v["dram_bytes_per_input_byte"], 2) for k, v in stages.items()}), "total", round(tot, 2))


if __name__ == "__main__":
    main()
