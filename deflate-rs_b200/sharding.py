"""Sharding of the encode path over the GPUs of one node.

The path shards without any data-path collective: DEFLATE blocks only refer to the 32 KiB of
*plaintext* in front of them, which every rank can read from its own copy of the input.  Two
layouts are supported:

  independent streams   unit i -> rank i mod G; every unit is a complete stream
                        (`assign_units`)
  one stream, G pieces  contiguous pieces; rank g encodes piece g with the 32 KiB in front of it
                        as dictionary and ends it with the reference's sync marker
                        (compress.rs:258-261), the last rank sets BFINAL; the concatenation is
                        the stream the reference's writer produces with flush() called at the
                        piece boundaries (`piece_bounds`, `encode_piece_device`,
                        `combine_adler32`)

The one exchange step is bringing the compressed pieces to rank 0: an all-gather of the sizes, then
grouped point-to-point transfers (NCCL has no gather-v).  On GPUs that is the library's own
`dfl_gather_device` over its own NCCL communicator (`Comm`: ncclAllGather + grouped
ncclSend/ncclRecv on the caller's stream, csrc/dfl_comm.cu); `gather_streams` is the same exchange
written with torch.distributed, which also runs over `gloo` on CPU tensors -- that is how
tests/test_sharding_cpu.py covers the host-side logic without a GPU.
"""
import ctypes

WINDOW = 32768
ADLER_MOD = 65521


def assign_units(n_units: int, world: int, rank: int):
    """Independent units (e.g. PNG IDAT chunks): round robin."""
    return list(range(rank, n_units, world))


def piece_bounds(n: int, world: int, align: int = 1 << 16):
    """Contiguous pieces of one stream, boundaries aligned to `align` bytes (except the end).
    Returns [(lo, hi)] * world; trailing pieces may be empty for tiny inputs."""
    per = -(-n // world)
    per = -(-per // align) * align
    out = []
    for g in range(world):
        lo = min(n, g * per)
        hi = min(n, lo + per)
        out.append((lo, hi))
    return out


def combine_adler32(parts):
    """parts: [(adler32 of piece, length)] in stream order -> Adler-32 of the concatenation
    (RFC 1950; the same arithmetic as dfl_core.h adler32_combine)."""
    a, b = 1, 0
    for ad, ln in parts:
        a2, b2 = ad & 0xFFFF, ad >> 16
        b = (b + b2 + (ln % ADLER_MOD) * ((a + ADLER_MOD - 1) % ADLER_MOD)) % ADLER_MOD
        a = (a + a2 + ADLER_MOD - 1) % ADLER_MOD
    return (b << 16) | a


def encode_piece_device(src, lo: int, hi: int, options, last: bool, out=None, stream=None):
    """Encode src[lo:hi) (a CUDA uint8 tensor holding at least the bytes from max(0, lo - 32768) on
    -- here: the whole input) as one piece of a longer raw DEFLATE stream.  Returns (out, n_bytes)."""
    import torch

    from . import CompressionOptions, _native

    L = _native.lib()
    d0 = max(0, lo - WINDOW)
    d0 &= ~15
    view = src[d0:hi]
    if out is None:
        out = torch.empty(L.dfl_bound(hi - lo, _native.RAW) + 64, dtype=torch.uint8, device=src.device)
    opts = CompressionOptions.from_(options)._c()
    sz = ctypes.c_size_t()
    st = ctypes.c_void_p(stream if stream is not None else torch.cuda.current_stream(src.device).cuda_stream)
    with torch.cuda.device(src.device):
        rc = L.dfl_compress_device_piece(ctypes.c_void_p(view.data_ptr()), hi - d0, lo - d0, ctypes.byref(opts),
                                         _native.FLUSH_FINISH if last else _native.FLUSH_SYNC,
                                         ctypes.c_void_p(out.data_ptr()), out.numel(), ctypes.byref(sz), st)
    _native.check(rc, "dfl_compress_device_piece")
    return out, sz.value


def gather_streams(local, n_bytes: int, dst: int = 0, group=None, recv_buf=None):
    """Bring every rank's first n_bytes of `local` (uint8 tensor, CUDA under nccl / CPU under gloo) to
    rank `dst`.  Returns (buffer, offsets) on dst -- rank r's bytes are buffer[offsets[r]:offsets[r+1]] --
    and (None, offsets) elsewhere.  One size all-gather, then grouped send/recv."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes_t = torch.zeros(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(sizes_t, torch.tensor([n_bytes], dtype=torch.int64, device=local.device), group=group)
    sizes = [int(x) for x in sizes_t.tolist()]
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    if rank == dst:
        if recv_buf is None or recv_buf.numel() < offs[-1]:
            recv_buf = torch.empty(offs[-1] + offs[-1] // 8 + 4096, dtype=torch.uint8, device=local.device)
        ops = [dist.P2POp(dist.irecv, recv_buf[offs[r]:offs[r + 1]], r, group) for r in range(world)
               if r != dst and sizes[r] > 0]
        recv_buf[offs[dst]:offs[dst + 1]].copy_(local[:n_bytes])
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return recv_buf, offs
    if n_bytes > 0:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, local[:n_bytes], dst, group)]):
            w.wait()
    return None, offs


class Comm:
    """The library's NCCL communicator (dfl_comm_*), bootstrapped through an existing torch.distributed group:
    rank 0 creates the NCCL unique id, a broadcast hands it to the others, every rank then joins.  One process
    per GPU; the communicator is bound to the current CUDA device."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist

        from . import _native

        self._L = _native.lib()
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        idbuf = (ctypes.c_uint8 * _native.COMM_ID_BYTES)()
        if self.rank == 0:
            _native.check(self._L.dfl_comm_unique_id(idbuf), "dfl_comm_unique_id")
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.tensor(list(idbuf), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0, group=group)
        idbuf = (ctypes.c_uint8 * _native.COMM_ID_BYTES)(*t.cpu().tolist())
        h = ctypes.c_void_p()
        rc = self._L.dfl_comm_init(ctypes.byref(h), self.world, self.rank, idbuf)
        if rc != 0:
            raise RuntimeError(f"dfl_comm_init: status {rc}: {self._L.dfl_comm_last_error().decode()}")
        self._h = h

    def gather(self, local, n_bytes: int, recv_buf=None, root: int = 0, stream=None):
        """dfl_gather_device: rank r's first n_bytes of `local` arrive at recv_buf[offs[r]:offs[r+1]] on `root`.
        Returns (sizes, recv_buf).  The payload is only queued on `stream` (default: torch's current stream)."""
        import torch

        sizes = (ctypes.c_size_t * self.world)()
        st = ctypes.c_void_p(stream if stream is not None else torch.cuda.current_stream(local.device).cuda_stream)
        dst = ctypes.c_void_p(recv_buf.data_ptr()) if recv_buf is not None else None
        cap = recv_buf.numel() if recv_buf is not None else 0
        rc = self._L.dfl_gather_device(self._h, ctypes.c_void_p(local.data_ptr()), n_bytes, dst, cap, sizes, root, st)
        if rc != 0:
            raise RuntimeError(f"dfl_gather_device: status {rc}: {self._L.dfl_comm_last_error().decode()}")
        return [int(x) for x in sizes], recv_buf

    def close(self):
        if self._h:
            self._L.dfl_comm_free(self._h)
            self._h = None
