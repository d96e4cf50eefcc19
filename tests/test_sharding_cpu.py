"""SURVEY.md 8(e) on CPU: the N > 1 host logic (piece planning, the size all-gather + grouped
send/recv that brings the compressed pieces to rank 0, checksum combination) under world_size 2
with the gloo backend.  The pieces themselves are produced by the oracle here (it plays the role
the CUDA encode plays on a GPU box: tests/test_gpu_parity.py checks that encode_piece_device
produces exactly these bytes)."""
import os
import socket
import zlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as o
from conftest import fixture_bytes
from deflate_rs_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_pieces(data, bounds):
    """Output of the oracle's writer split at its flush points: piece g = bytes produced by write(piece g)
    + flush() (Finish for the last)."""
    s = o.Stream(o.opts_default(), o.RAW)
    pieces, seen = [], 0
    for g, (lo, hi) in enumerate(bounds):
        s.write(data[lo:hi])
        if g + 1 == len(bounds):
            out = s.finish()
        else:
            s.flush()
            out = s.output()
        pieces.append(out[seen:])
        seen = len(out)
    return pieces


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = fixture_bytes("pg11.txt")
        bounds = sharding.piece_bounds(len(data), world, align=4096)
        mine = _oracle_pieces(data, bounds)[rank]           # what this rank's GPU would have produced
        local = torch.frombuffer(bytearray(mine + b"\0" * 64), dtype=torch.uint8)   # device buffers carry slack
        buf, offs = sharding.gather_streams(local, len(mine), dst=0)
        # checksums travel the same way (8 bytes per rank)
        lo, hi = bounds[rank]
        ad = torch.tensor([zlib.adler32(data[lo:hi]), hi - lo], dtype=torch.int64)
        ads = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(ads, ad)
        if rank == 0:
            stream = bytes(buf[:offs[-1]].numpy())
            q.put((stream, offs, [(int(a[0]), int(a[1])) for a in ads]))
        else:
            assert buf is None
        # independent units: every unit lands on exactly one rank
        units = sharding.assign_units(7, world, rank)
        got = [None] * world
        dist.all_gather_object(got, units)
        assert sorted(sum(got, [])) == list(range(7))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gather_of_one_stream():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    stream, offs, ads = q.get(timeout=90)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = fixture_bytes("pg11.txt")
    bounds = sharding.piece_bounds(len(data), world, align=4096)
    # the gathered pieces are, in order, exactly the reference writer's output with flush() at the boundary
    s = o.Stream(o.opts_default(), o.RAW)
    s.write(data[:bounds[0][1]]); s.flush(); s.write(data[bounds[1][0]:])
    assert stream == s.finish()
    assert zlib.decompress(stream, -15) == data
    assert offs[0] == 0 and len(offs) == world + 1 and offs[-1] == len(stream)
    # every piece but the last ends with the sync marker, so the seams are byte aligned
    assert stream[offs[1] - 4:offs[1]] == b"\x00\x00\xff\xff"
    assert sharding.combine_adler32(ads) == zlib.adler32(data)


def test_piece_bounds_and_adler_combine():
    for n in (0, 1, 65535, 65536, 1 << 20, (1 << 20) + 17):
        for world in (1, 2, 3, 8):
            b = sharding.piece_bounds(n, world)
            assert b[0][0] == 0 and b[-1][1] == n and len(b) == world
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert all(lo % 65536 == 0 or lo == n for lo, _ in b)
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, 300000, dtype=np.uint8).tobytes()
    cuts = [0, 1, 70000, 70000, 299999, 300000]
    parts = [(zlib.adler32(data[a:b]), b - a) for a, b in zip(cuts, cuts[1:])]
    assert sharding.combine_adler32(parts) == zlib.adler32(data)
