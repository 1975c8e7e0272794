"""Block-parallel bzip2 encode over the GPUs of one box.

bzip2 blocks are independent once the RLE1 stage has cut them, so every rank encodes a contiguous
range of blocks on its own GPU with no data-path collective.  The only exchange is the final
bitstream gather (north_star: "NCCL over NVLink only for the final bitstream gather"):

  1. every rank cuts the blocks (b2_bzip2_plan -- a cheap scan) and encodes blocks
     [first, first+count) into a fragment that starts at bit 0 of a private buffer
     (b2_bzip2_encode_range_dev)
  2. all_gather of (fragment bits, block count)   -> every rank knows its global bit offset
  3. the fragment is shifted to (global offset mod 8) so that only whole bytes move
  4. gather of the byte fragments to rank 0 (NCCL), neighbouring fragments share at most one
     byte, which is OR-ed; rank 0 adds "BZh"+level and the trailer (stream CRC folded over the
     per-block CRCs in order, lib/Bzip2.js:917,925-927)

The resulting stream is byte-identical to Bzip2.compressFile on one GPU (and to the oracle).
The shifting / merging below is plain torch tensor code so that the same logic runs on CPU tensors
with the gloo backend in the unit tests (tests/test_sharded_host.py), where the per-range encoder
is injected.
"""
import ctypes as C
import time

import torch
import torch.distributed as dist

PHASES = {}   # wall-clock milliseconds of the last compress_file_sharded call per phase (rank local)


def _tick(name, t0):
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    PHASES[name] = PHASES.get(name, 0.0) + (t1 - t0) * 1e3
    return t1

SQRTPI = 0x177245385090


def block_range(nblocks, rank, world):
    first = rank * nblocks // world
    last = (rank + 1) * nblocks // world
    return first, last - first


def shift_right_bits(frag, nbits, phase):
    """Returns a uint8 tensor holding `frag`'s first nbits starting at bit `phase` (MSB first)."""
    nbytes = (phase + nbits + 7) // 8
    if frag.is_cuda and nbits:
        from . import _native
        out = torch.empty(((phase + nbits + 31) // 32) * 4, dtype=torch.uint8, device=frag.device)
        torch.cuda.current_stream().synchronize()
        rc = _native.lib().b2_bitshift_dev(frag.data_ptr(), nbits, phase, out.data_ptr())
        if rc:
            raise RuntimeError("b2_bitshift_dev: " + _native.last_error())
        return out[:nbytes]
    src = frag[: (nbits + 7) // 8]
    out = torch.zeros(nbytes, dtype=torch.uint8, device=frag.device)
    if nbits == 0:
        return out
    if phase == 0:
        out[: src.numel()] = src
    else:
        s16 = src.to(torch.int16)
        out[: src.numel()] |= (s16 >> phase).to(torch.uint8)
        spill = ((s16 << (8 - phase)) & 0xFF).to(torch.uint8)
        k = min(nbytes - 1, src.numel())
        out[1: 1 + k] |= spill[:k]
    # clear the bits behind the fragment
    tail = (phase + nbits) % 8
    if tail:
        out[-1] &= (0xFF << (8 - tail)) & 0xFF
    return out


def fold_stream_crc(crcs):
    s = 0
    for c in crcs:
        s = (((s << 1) | (s >> 31)) ^ int(c)) & 0xFFFFFFFF  # lib/Bzip2.js:917
    return s


def trailer_bytes(bitpos, stream_crc):
    """The 80 trailer bits placed at absolute bit `bitpos`: (first byte index, bytes)."""
    val = (SQRTPI << 32) | stream_crc
    phase = bitpos % 8
    total = phase + 80
    nbytes = (total + 7) // 8
    val <<= nbytes * 8 - total
    return bitpos // 8, val.to_bytes(nbytes, "big")


def assemble(level, frags, bits, crcs_per_rank, device):
    """Rank-0 side: frags[r] = byte tensor already shifted to its phase; bits[r] = fragment bits."""
    offs, o = [], 32
    for b in bits:
        offs.append(o)
        o += int(b)
    total_bits = o + 80
    nbytes_out = (total_bits + 7) // 8
    out = torch.empty(nbytes_out, dtype=torch.uint8, device=device)
    out[max(0, nbytes_out - 12):] = 0   # trailer region (OR-ed below)
    out[:4] = torch.tensor(list(b"BZh" + bytes([0x30 + level])), dtype=torch.uint8, device=device)
    edge = {}                   # bytes shared by two neighbours: byte index -> OR of the contributions
    for r, f in enumerate(frags):
        if bits[r] == 0:
            continue
        b0 = offs[r] // 8
        nb = (offs[r] % 8 + int(bits[r]) + 7) // 8
        if nb > 2:
            out[b0 + 1: b0 + nb - 1] = f[1: nb - 1].to(device)     # interior bytes belong to this fragment alone
        for bi in {0, nb - 1}:
            edge[b0 + bi] = edge.get(b0 + bi, 0) | int(f[bi])
    for k, v in edge.items():
        if k >= 4:
            out[k] = v
    crcs = [c for rc in crcs_per_rank for c in rc]
    b0, tb = trailer_bytes(o, fold_stream_crc(crcs))
    # the first trailer byte may share its byte with the last fragment (already written via `edge`)
    out[b0: b0 + len(tb)] |= torch.tensor(list(tb), dtype=torch.uint8, device=device)
    return out


def compress_sharded(encode_range, nblocks, level, device, group=None):
    """encode_range(first, count) -> (uint8 tensor fragment starting at bit 0, nbits, [block crcs]).
    Returns the complete .bz2 stream as a uint8 tensor on rank 0 (None elsewhere)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    first, count = block_range(nblocks, rank, world)
    t0 = time.perf_counter()
    frag, nbits, crcs = encode_range(first, count)
    t0 = _tick("encode_range", t0)
    if world == 1:
        return assemble(level, [frag], [nbits], [crcs], device)
    # 2. everybody learns every fragment's size
    mine = torch.tensor([nbits, count], dtype=torch.int64, device=device)
    allv = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    bits = [int(v[0]) for v in allv]
    counts = [int(v[1]) for v in allv]
    off = 32 + sum(bits[:rank])
    phase = off % 8
    t0 = _tick("allgather_sizes", t0)
    shifted = shift_right_bits(frag, nbits, phase)
    t0 = _tick("shift", t0)
    # 3. gather the (padded) byte fragments and the block CRCs on rank 0
    maxlen = max((32 + sum(bits[:r])) % 8 + bits[r] + 7 for r in range(world)) // 8 + 1
    maxcnt = max(counts + [1])
    pad = torch.zeros(maxlen, dtype=torch.uint8, device=device)
    pad[: shifted.numel()] = shifted
    crct = torch.zeros(maxcnt, dtype=torch.int64, device=device)
    if count:
        crct[:count] = torch.tensor([int(c) for c in crcs], dtype=torch.int64, device=device)
    if rank == 0:
        glist = [torch.empty(maxlen, dtype=torch.uint8, device=device) for _ in range(world)]
        clist = [torch.zeros(maxcnt, dtype=torch.int64, device=device) for _ in range(world)]
        t0 = _tick("pad_alloc", t0)
        dist.gather(pad, glist, dst=0, group=group)
        dist.gather(crct, clist, dst=0, group=group)
        t0 = _tick("nccl_gather", t0)
        crcs_per_rank = [clist[r][: counts[r]].tolist() for r in range(world)]
        out = assemble(level, glist, bits, crcs_per_rank, device)
        _tick("assemble", t0)
        return out
    t0 = _tick("pad_alloc", t0)
    dist.gather(pad, None, dst=0, group=group)
    dist.gather(crct, None, dst=0, group=group)
    _tick("nccl_gather", t0)
    return None


def _range_encoder(L, d_in, n, level):
    from . import _native

    def encode_range(first, count):
        if count == 0:
            return torch.zeros(8, dtype=torch.uint8, device=d_in.device), 0, []
        cap = count * 1400000 + 4096
        out = torch.empty(cap, dtype=torch.uint8, device=d_in.device)
        bits = C.c_uint64()
        crcs = (C.c_uint32 * count)()
        rc = L.b2_bzip2_encode_range_dev(d_in.data_ptr(), n, level, first, count, 0, out.data_ptr(), cap, C.byref(bits), crcs)
        if rc:
            raise RuntimeError("b2_bzip2_encode_range_dev: " + _native.last_error())
        return out, int(bits.value), list(crcs)

    return encode_range


def gpu_encode_range_fn(d_in, level):
    """encode_range callable backed by libb2bz.so for a uint8 CUDA tensor holding the whole input
    (exact plan: every block boundary of the file is cut on this GPU)."""
    from . import _native
    L = _native.lib()
    n = d_in.numel()
    total = C.c_size_t()
    rc = L.b2_bzip2_plan(d_in.data_ptr(), n, level, C.byref(total))
    if rc:
        raise RuntimeError("b2_bzip2_plan: " + _native.last_error())
    return _range_encoder(L, d_in, n, level), int(total.value)


def spec_plan_ok(infos, n):
    """infos[r] = (raw_start, raw_end, first, planned, cut, total) of every rank's speculative plan."""
    world = len(infos)
    total = infos[0][5]
    pos = 0
    nxt = 0
    for r in range(world):
        s, e, first, planned, cut, tot = infos[r]
        if tot != total or first != nxt or cut != planned:
            return False
        if planned:
            if s != pos:
                return False
            pos = e
        nxt = first + planned
    return nxt == total and pos == n


def compress_file_sharded(d_in, level=9, group=None):
    """Whole-file bzip2 encode of a CUDA uint8 tensor present on every rank; stream on rank 0.

    Block cutting: every rank cuts only ITS share of the blocks from a speculative start boundary
    (b2_bzip2_plan_spec); the ranks then check that the pieces chain exactly (end(r) == start(r+1), ...).
    If a run-phase slip makes the speculation fail anywhere, all ranks fall back to the exact plan."""
    from . import _native
    L = _native.lib()
    n = d_in.numel()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        enc, nblocks = gpu_encode_range_fn(d_in, level)
        return compress_sharded(enc, nblocks, level, d_in.device, group)
    PHASES.clear()
    t0 = time.perf_counter()
    info = (C.c_uint64 * 6)()
    rc = L.b2_bzip2_plan_spec(d_in.data_ptr(), n, level, rank, world, info)
    if rc:
        raise RuntimeError("b2_bzip2_plan_spec: " + _native.last_error())
    mine = torch.tensor([int(v) for v in info], dtype=torch.int64, device=d_in.device)
    allv = [torch.zeros(6, dtype=torch.int64, device=d_in.device) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    infos = [tuple(int(x) for x in v.tolist()) for v in allv]
    _tick("plan_spec+verify", t0)
    if spec_plan_ok(infos, n):
        enc = _range_encoder(L, d_in, n, level)
        first, count = infos[rank][2], infos[rank][3]
        return compress_sharded(lambda f, c: enc(first, count), infos[0][5], level, d_in.device, group)
    enc, nblocks = gpu_encode_range_fn(d_in, level)
    return compress_sharded(enc, nblocks, level, d_in.device, group)


# ---- sharded decode -------------------------------------------------------------------------------
def decode_shard_rows(L, d_in, rank, world):
    """Stage 1 on one rank: returns (info, rows) -- rows = int64 tensor [own candidates, 6] on the CPU."""
    from . import _native
    info = (C.c_uint64 * 3)()
    rc = L.b2_dec_shard_open(d_in.data_ptr(), d_in.numel(), rank, world, info)
    if rc:
        raise RuntimeError("b2_dec_shard_open: %s (code %d)" % (_native.last_error(), rc))
    total, lo, hi = int(info[0]), int(info[1]), int(info[2])
    buf = (C.c_uint64 * (6 * max(hi - lo, 1)))()
    rc = L.b2_dec_shard_export(buf)
    if rc:
        raise RuntimeError("b2_dec_shard_export: " + _native.last_error())
    rows = torch.tensor([int(v) if int(v) < 2 ** 63 else int(v) - 2 ** 64 for v in buf[: 6 * (hi - lo)]], dtype=torch.int64).reshape(-1, 6)
    return (total, lo, hi), rows


def decode_shard_finish(L, all_rows, multistream, device):
    """Stage 2 on one rank: all_rows = [total candidates, 6] int64 (CPU).  Returns (own output tensor, res)."""
    from . import _native
    flat = [int(v) & (2 ** 64 - 1) for v in all_rows.reshape(-1).tolist()]
    arr = (C.c_uint64 * max(len(flat), 1))(*flat)
    res = (C.c_uint64 * 5)()
    need = 0
    for r in all_rows.tolist():
        need += r[4] if r[0] == 0 else 0
    out = torch.empty(max(need, 1), dtype=torch.uint8, device=device)   # upper bound: everything decodable
    rc = L.b2_dec_shard_finish(arr, int(bool(multistream)), out.data_ptr(), out.numel(), res)
    vals = [int(v) for v in res]
    err_idx = vals[3] if vals[3] < 2 ** 63 else vals[3] - 2 ** 64
    err_code = vals[4] if vals[4] < 2 ** 63 else vals[4] - 2 ** 64
    msg = _native.last_error() if rc else ""
    return out[: vals[1]] if rc == 0 else None, dict(off=vals[0], len=vals[1], total=vals[2], err_idx=err_idx if rc else -1,
                                                      err_code=err_code if rc else 0, msg=msg)


def decompress_file_sharded(d_in, multistream=False, group=None):
    """Decode a .bz2 stream held on every rank; the decoded bytes are gathered on rank 0 (uint8 tensor)."""
    from . import _native
    from .bzip2 import Bzip2Error
    L = _native.lib()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    device = d_in.device
    (total, lo, hi), rows = decode_shard_rows(L, d_in, rank, world)
    if world > 1:
        per = max((r + 1) * total // world - r * total // world for r in range(world))
        pad = torch.zeros((max(per, 1), 6), dtype=torch.int64, device=device)
        if hi > lo:
            pad[: hi - lo] = rows.to(device)
        allp = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(allp, pad, group=group)
        parts = []
        for r in range(world):
            cnt = (r + 1) * total // world - r * total // world
            parts.append(allp[r][:cnt].cpu())
        all_rows = torch.cat(parts) if parts else rows
    else:
        all_rows = rows
    out, res = decode_shard_finish(L, all_rows, multistream, device)
    # earliest failing event over all ranks wins (every rank sees the same event list)
    mine = torch.tensor([res["err_idx"] if res["err_idx"] >= 0 else 2 ** 62, res["err_code"], res["off"], res["len"], res["total"]],
                        dtype=torch.int64, device=device)
    if world > 1:
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine, group=group)
    else:
        allv = [mine]
    errs = [(int(v[0]), int(v[1]), r) for r, v in enumerate(allv) if int(v[0]) < 2 ** 62]
    if errs:
        idx, code, who = min(errs)
        raise Bzip2Error(code, res["msg"] if who == rank else "Data error")
    if world == 1:
        return out
    # gather the shards on rank 0
    maxlen = max(int(v[3]) for v in allv)
    padb = torch.empty(max(maxlen, 1), dtype=torch.uint8, device=device)
    padb[: out.numel()] = out
    if rank == 0:
        glist = [torch.empty_like(padb) for _ in range(world)]
        dist.gather(padb, glist, dst=0, group=group)
        full = torch.empty(int(allv[0][4]), dtype=torch.uint8, device=device)
        for r in range(world):
            o, ln = int(allv[r][2]), int(allv[r][3])
            full[o: o + ln] = glist[r][:ln]
        return full
    dist.gather(padb, None, dst=0, group=group)
    return None
