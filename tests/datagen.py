"""Seeded synthetic inputs for the BASELINE.json configs.  numpy only.

All generators are deterministic functions of (seed, size).  They build one 1 MiB-segment "mix"
from a vocabulary taken from tests/fixtures/pg11.txt (public-domain text shipped with the repo).
"""
import os

import numpy as np

_FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures")
_SEG = 1 << 20


def _vocab():
    text = open(os.path.join(_FIX, "pg11.txt"), "rb").read()
    words = text.split()
    uniq, counts = np.unique(np.array(words, dtype=object), return_counts=True)
    order = np.argsort(-counts, kind="stable")
    return [uniq[i] for i in order]


_VOCAB = None


def _zipf_text(rng, size):
    global _VOCAB
    if _VOCAB is None:
        _VOCAB = _vocab()
    v = _VOCAB
    ranks = np.arange(1, len(v) + 1, dtype=np.float64)
    p = ranks ** -1.1
    p /= p.sum()
    n_words = size // 4 + 16
    idx = rng.choice(len(v), size=n_words, p=p)
    seps = rng.choice([b" ", b" ", b" ", b" ", b", ", b". ", b"\r\n", b"; "], size=n_words,
                      p=[0.55, 0.1, 0.1, 0.05, 0.07, 0.06, 0.05, 0.02])
    parts = []
    total = 0
    # phrase reuse: with probability 0.10 a recent 3..8 word phrase is repeated, as prose does
    reuse = rng.random(n_words) < 0.10
    back = rng.integers(8, 4000, size=n_words)
    plen = rng.integers(3, 9, size=n_words)
    i = 0
    while total < size and i < n_words:
        if reuse[i] and len(parts) > back[i]:
            start = len(parts) - int(back[i])
            ph = parts[start:start + int(plen[i])]
            parts.extend(ph)
            total += sum(len(x) for x in ph)
        else:
            w = v[idx[i]] + seps[i]
            parts.append(w)
            total += len(w)
        i += 1
    return (b"".join(parts) + b" " * size)[:size]


def _xml_records(rng, size):
    tags = [b"page", b"title", b"id", b"revision", b"timestamp", b"contributor", b"username", b"text", b"comment"]
    out = []
    total = 0
    k = int(rng.integers(1000, 100000))
    while total < size:
        t = tags[int(rng.integers(0, len(tags)))]
        val = str(k).encode() if rng.random() < 0.5 else ("%0.3f" % rng.random()).encode()
        rec = b"  <" + t + b' key="' + str(int(rng.integers(0, 50))).encode() + b'">' + val + b"</" + t + b">\n"
        out.append(rec)
        total += len(rec)
        k += int(rng.integers(1, 4))
    return b"".join(out)[:size]


def _binary_records(rng, size):
    n = size // 32 + 1
    base = rng.integers(0, 256, size=32, dtype=np.int64)
    deltas = rng.integers(-2, 3, size=(n, 32), dtype=np.int64)
    deltas *= (rng.random((n, 32)) < 0.30)
    deltas[:, :8] = 0
    rec = (base[None, :] + np.cumsum(deltas, axis=0)) & 0xFF
    rec[:, 0:4] = (np.arange(n)[:, None] >> (8 * np.arange(4))[None, :]) & 0xFF
    return rec.astype(np.uint8).tobytes()[:size]


def _sparse(rng, size):
    a = np.zeros(size, dtype=np.uint8)
    k = size // 12
    pos = rng.integers(0, size, size=k)
    a[pos] = rng.integers(1, 256, size=k, dtype=np.uint8)
    return a.tobytes()


def _random(rng, size):
    return rng.integers(0, 256, size=size, dtype=np.uint8).tobytes()


def silesia_mix(size: int, seed: int = 0x51DE51A) -> bytes:
    """C2: 55 % Zipf word text, 15 % XML-like, 15 % binary records, 10 % sparse, 5 % random, in 1 MiB
    segments drawn i.i.d..  Segments are generated once per kind and re-cut at random
    offsets so that 1 GiB can be produced in seconds; matches never reach across segments further
    than the 32 KiB window anyway."""
    rng = np.random.default_rng(seed)
    kinds = [_zipf_text, _xml_records, _binary_records, _sparse, _random]
    probs = [0.55, 0.15, 0.15, 0.10, 0.05]
    pool_mib = 24
    pools = [np.frombuffer(k(np.random.default_rng(seed + 1 + i), pool_mib * _SEG), dtype=np.uint8)
             for i, k in enumerate(kinds)]
    n_seg = (size + _SEG - 1) // _SEG
    which = rng.choice(len(kinds), size=n_seg, p=probs)
    offs = rng.integers(0, (pool_mib - 1) * _SEG, size=n_seg)
    out = np.empty(n_seg * _SEG, dtype=np.uint8)
    for s in range(n_seg):
        out[s * _SEG:(s + 1) * _SEG] = pools[which[s]][offs[s]:offs[s] + _SEG]
    return out[:size].tobytes()


def enwik_like(size: int, seed: int = 0xE2010C) -> bytes:
    """C3: word text interleaved with wiki/XML markup tokens and ascending integers."""
    rng = np.random.default_rng(seed)
    pool_mib = 16
    text = np.frombuffer(_zipf_text(np.random.default_rng(seed + 1), pool_mib * _SEG), dtype=np.uint8)
    xml = np.frombuffer(_xml_records(np.random.default_rng(seed + 2), pool_mib * _SEG), dtype=np.uint8)
    n_seg = (size + 65535) // 65536
    out = np.empty(n_seg * 65536, dtype=np.uint8)
    which = rng.random(n_seg) < 0.25
    offs = rng.integers(0, pool_mib * _SEG - 65536, size=n_seg)
    for s in range(n_seg):
        src = xml if which[s] else text
        out[s * 65536:(s + 1) * 65536] = src[offs[s]:offs[s] + 65536]
    return out[:size].tobytes()


def png_idat_like(size: int, seed: int = 0x1DA7) -> bytes:
    """C4: filtered RGBA scanlines, 1024 px -> 4097 bytes/row: filter byte then small residuals."""
    rng = np.random.default_rng(seed)
    rows = size // 4097 + 1
    res = rng.geometric(0.35, size=(rows, 4096)) - 1
    sign = rng.integers(0, 2, size=(rows, 4096)) * 2 - 1
    body = ((res * sign) & 0xFF).astype(np.uint8)
    filt = rng.integers(0, 5, size=(rows, 1), dtype=np.uint8)
    return np.concatenate([filt, body], axis=1).tobytes()[:size]


def binary_like(size: int, seed: int = 0xB1A2) -> bytes:
    """C5: executable-like mix (opcode n-grams, pointer tables, strings, zero padding)."""
    rng = np.random.default_rng(seed)
    parts = []
    total = 0
    ops = rng.integers(0, 256, size=(64, 6), dtype=np.uint8)
    while total < size:
        kind = rng.random()
        if kind < 0.5:
            idx = rng.choice(64, size=4096, p=None)
            blk = ops[idx].reshape(-1)[: int(rng.integers(2048, 16384))].tobytes()
        elif kind < 0.7:
            n = int(rng.integers(256, 2048))
            ptr = (0x00400000 + np.cumsum(rng.integers(4, 64, size=n))).astype("<u4")
            blk = ptr.tobytes()
        elif kind < 0.85:
            blk = _zipf_text(rng, int(rng.integers(512, 4096))).replace(b" ", b"\x00")
        else:
            blk = bytes(int(rng.integers(64, 4096)))
        parts.append(blk)
        total += len(blk)
    return b"".join(parts)[:size]


def random_bytes(size: int, seed: int = 7) -> bytes:
    """Incompressible input: every block ends up stored (stored_block.rs:13-40)."""
    return _random(np.random.default_rng(seed), size)


def zero_bytes(size: int, seed: int = 0) -> bytes:
    """One repeated byte: 258-byte matches at distance 1 end to end; speculative parses never resynchronise."""
    return bytes(size)


def issue44_like(size: int, seed: int = 0) -> bytes:
    """tests/fixtures/issue_44.zlib inflated (25 MiB over three byte values, long runs) and repeated."""
    import zlib
    raw = zlib.decompress(open(os.path.join(_FIX, "issue_44.zlib"), "rb").read())
    reps = -(-size // len(raw))
    return (raw * reps)[:size]
