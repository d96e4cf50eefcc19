#!/usr/bin/env python3
"""Regenerates tests/golden/*.json.  Run in the build container only (reads /root/reference).

1. ref_tables.json  -- the constant tables of the reference (huffman_table.rs), parsed from the
   Rust source text.  tests/test_oracle_kat.py checks the oracle's generated tables against them.
2. oracle_pins.json -- size + sha256 of the oracle's output for every reference fixture and preset,
   so that later edits to oracle/ cannot silently change its behaviour.
"""
import hashlib
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src"
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rust_array(src, name):
    m = re.search(r"(?:static|const)\s+" + name + r"\s*:\s*\[[^\]]*\]\s*=\s*\[(.*?)\];", src, re.S)
    body = re.sub(r"//.*", "", m.group(1))
    return [int(x) for x in re.findall(r"\d+", body)]


def main():
    src = open(os.path.join(REF, "huffman_table.rs")).read()
    tables = {name: rust_array(src, name) for name in
              ["FIXED_CODE_LENGTHS", "LENGTH_EXTRA_BITS_LENGTH", "LENGTH_CODE", "BASE_LENGTH",
               "DISTANCE_CODES", "DISTANCE_EXTRA_BITS", "DISTANCE_BASE"]}
    hl = open(os.path.join(REF, "huffman_lengths.rs")).read()
    tables["HUFFMAN_LENGTH_ORDER"] = rust_array(hl, "HUFFMAN_LENGTH_ORDER")
    tables["_source"] = "parsed from /root/reference/src/huffman_table.rs:32-111, huffman_lengths.rs:27-29"
    json.dump(tables, open(os.path.join(HERE, "ref_tables.json"), "w"))

    import zlib
    import oracle_lib as o
    fx = os.path.join(ROOT, "tests", "fixtures")
    files = ["pg11.txt", "short.bin", "issue_18_201911.bin", "dump.bin"] + \
        sorted("afl/" + f for f in os.listdir(os.path.join(fx, "afl")))
    pins = {}
    for f in files:
        data = open(os.path.join(fx, f), "rb").read()
        for pname, pf in o.PRESETS.items():
            c = o.compress(data, pf(), o.RAW)
            assert zlib.decompress(c, -15) == data
            pins[f + ":" + pname] = [len(c), hashlib.sha256(c).hexdigest()]
    data = zlib.decompress(open(os.path.join(fx, "issue_44.zlib"), "rb").read())
    for pname in ("default", "fast"):
        c = o.compress(data, o.PRESETS[pname](), o.RAW)
        assert zlib.decompress(c, -15) == data
        pins["issue_44:" + pname] = [len(c), hashlib.sha256(c).hexdigest()]
    json.dump(pins, open(os.path.join(HERE, "oracle_pins.json"), "w"), indent=0, sort_keys=True)
    print("wrote", len(pins), "pins")


if __name__ == "__main__":
    main()
