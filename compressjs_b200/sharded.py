"""Block-parallel bzip2 encode over the GPUs of one box.

bzip2 blocks are independent once the RLE1 stage has cut them, so every rank encodes a contiguous
range of blocks on its own GPU with no data-path collective.  The only exchange is the final
bitstream gather (north_star: "NCCL over NVLink only for the final bitstream gather"):

  1. every rank cuts the blocks (b2_bzip2_plan -- a cheap scan) and encodes blocks
     [first, first+count) into a fragment that starts at bit 0 of a private buffer
     (b2_bzip2_encode_range_dev)
  2. all_gather of (fragment bits, block count)   -> every rank knows its global bit offset
  3. the fragment is shifted to (global offset mod 8) so that only whole bytes move
  4. the byte fragments travel to rank 0 straight into their final place in the output buffer
     (NCCL send/recv into offset views; no padded staging, no second copy); neighbouring fragments
     share at most one byte, which is OR-ed from the two edge bytes that ride along with the sizes;
     rank 0 adds "BZh"+level and the trailer (stream CRC combined from the per-rank folds of the
     block CRCs, lib/Bzip2.js:917,925-927)

With a sharded INPUT (compress_shares) a rank holds only its share of the bytes plus a halo: the ranks
exchange tiny share summaries (RLE1 run state, leading run, RLE1 output) so that every rank knows the
run state and the RLE1 output in front of its share, cuts its blocks speculatively and checks that the
pieces chain up (b2_bzip2_share_summary / b2_bzip2_plan_share).

The resulting stream is byte-identical to Bzip2.compressFile on one GPU (and to the oracle).
The shifting / merging below is plain torch tensor code so that the same logic runs on CPU tensors
with the gloo backend in the unit tests (tests/test_sharded_host.py), where the per-range encoder
is injected.
"""
import ctypes as C
import time

import numpy as np

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


def rotl32(v, k):
    k %= 32
    return ((v << k) | (v >> (32 - k))) & 0xFFFFFFFF if k else v & 0xFFFFFFFF


class SharedHostBuffer:
    """A page-locked host buffer that all ranks of one box map (POSIX shared memory, registered with CUDA in every
    process): every GPU downloads its fragment over its own PCIe link straight to its final place in the stream."""

    def __init__(self, nbytes, group=None):
        from multiprocessing import shared_memory, resource_tracker
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        names = [None]
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=max(int(nbytes), 4096))
            names = [self.shm.name]
        if dist.is_initialized():
            dist.broadcast_object_list(names, src=0, group=group)
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=names[0])
            try:  # only the creator unlinks the segment
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.owner = rank == 0
        self.tensor = torch.frombuffer(self.shm.buf, dtype=torch.uint8)
        self.registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), self.tensor.numel(), 0)
            self.registered = int(rc) == 0

    def close(self):
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self.registered = False
        self.tensor = None
        try:
            self.shm.close()
            if self.owner:
                self.shm.unlink()
        except Exception:
            pass


def place_fragments(frag, nbits, count, crcs, level, device, group=None, host_out=None, keep_sharded=False):
    """Every rank contributes a fragment (uint8 tensor starting at bit 0, nbits long) with `count` blocks and their
    CRCs; returns the complete .bz2 stream on rank 0 (None elsewhere).  The interior bytes of every fragment are
    received directly at their final byte offset of the output.  With host_out (a SharedHostBuffer's tensor, the same
    memory on every rank) the stream is assembled in host memory instead: every rank downloads its own fragment into
    place and rank 0 gets the stream's length back.  With keep_sharded nothing moves: every rank gets a ShardedStream
    (its piece at its final bit position; .gather() finishes the job when one GPU wants the whole stream)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    t0 = time.perf_counter()
    if world == 1:
        return assemble(level, [frag], [nbits], [crcs], device)
    # 1. sizes, block counts and the fold of the own block CRCs (lib/Bzip2.js:917 is linear: folds combine by rotation)
    mine = torch.tensor([nbits, count, fold_stream_crc(crcs)], dtype=torch.int64, device=device)
    allv = [torch.zeros(3, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    meta = [[int(x) for x in v.tolist()] for v in allv]
    bits = [m[0] for m in meta]
    offs, o = [], 32
    for b in bits:
        offs.append(o)
        o += b
    total_bits = o + 80
    phase = offs[rank] % 8
    t0 = _tick("allgather_sizes", t0)
    shifted = shift_right_bits(frag, nbits, phase)
    nb = (phase + nbits + 7) // 8 if nbits else 0
    t0 = _tick("shift", t0)
    # 2. the two edge bytes of every fragment (they may share their byte with a neighbour)
    eb = torch.zeros(2, dtype=torch.uint8, device=device)
    if nb:
        eb[0] = shifted[0]
        eb[1] = shifted[nb - 1]
    alle = [torch.zeros(2, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(alle, eb, group=group)
    nbs = [((offs[r] % 8) + bits[r] + 7) // 8 if bits[r] else 0 for r in range(world)]
    if host_out is not None:
        # 3'. interiors: every GPU writes its own piece of the host buffer (parallel PCIe links), then a barrier
        if nb > 2:
            b0 = offs[rank] // 8
            host_out[b0 + 1: b0 + nb - 1].copy_(shifted[1: nb - 1], non_blocking=True)
            if shifted.is_cuda:
                torch.cuda.current_stream().synchronize()
        dist.barrier(group=group)
        t0 = _tick("d2h_place", t0)
        if rank != 0:
            return None
        hb = host_out.numpy()
        for r in range(world):
            if nbs[r]:
                b0 = offs[r] // 8
                e0, e1 = int(alle[r][0]), int(alle[r][1])
                if r == 0 or (offs[r] % 8) == 0:
                    hb[b0] = e0
                else:
                    hb[b0] |= e0
                if nbs[r] > 1:
                    hb[b0 + nbs[r] - 1] = e1
        scrc = 0
        for m in meta:
            scrc = rotl32(scrc, m[1]) ^ m[2]
        b0t, tb = trailer_bytes(o, scrc)
        first_shared = (o % 8) != 0
        for k, v in enumerate(tb):
            if k == 0 and first_shared:
                hb[b0t] |= v
            else:
                hb[b0t + k] = v
        hb[:4] = np.frombuffer(b"BZh" + bytes([0x30 + level]), dtype=np.uint8)
        _tick("edges_trailer", t0)
        return (total_bits + 7) // 8
    st = ShardedStream(shifted, nb, offs, nbs, alle, meta, o, total_bits, level, rank, world, group, device)
    if keep_sharded:
        return st
    return st.gather()


class ShardedStream:
    """The finished stream, left where it was produced: every rank holds its fragment already shifted to its final bit
    position (`piece`, whose byte 0 is byte `offset` of the stream; neighbours share at most their edge bytes, which
    combine by OR) and everything needed to finish it (edge bytes of all fragments, block counts and CRC folds).
    gather() moves the pieces to rank 0 and returns the complete .bz2 there."""

    def __init__(self, shifted, nb, offs, nbs, alle, meta, o, total_bits, level, rank, world, group, device):
        self.piece, self.nb, self.offs, self.nbs, self.alle, self.meta = shifted, nb, offs, nbs, alle, meta
        self.end_bit, self.total_bits, self.level = o, total_bits, level
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.offset = offs[rank] // 8
        self.total_bytes = (total_bits + 7) // 8

    def _fixups(self):
        """Header, shared edge bytes and trailer: {byte offset: value} (the same on every rank)."""
        edge = {}
        for r in range(self.world):
            if self.nbs[r]:
                b0 = self.offs[r] // 8
                e0, e1 = int(self.alle[r][0]), int(self.alle[r][1])
                edge[b0] = edge.get(b0, 0) | e0
                edge[b0 + self.nbs[r] - 1] = edge.get(b0 + self.nbs[r] - 1, 0) | e1
        scrc = 0
        for m in self.meta:
            scrc = rotl32(scrc, m[1]) ^ m[2]
        b0t, tb = trailer_bytes(self.end_bit, scrc)
        fix = {b0t + k: v for k, v in enumerate(tb)}
        for k, v in edge.items():
            if k >= 4:
                fix[k] = fix.get(k, 0) | v
        for k, v in enumerate(b"BZh" + bytes([0x30 + self.level])):
            fix[k] = v
        return fix

    def write_file(self, path):
        """Every rank writes the interior bytes of its piece at their offset of `path` (one file, visible to all ranks);
        rank 0 sizes the file first and adds the header, the bytes neighbours share and the trailer.  Returns the
        stream length.  The consumer-side counterpart of leaving the stream sharded: no rank ever holds the whole file."""
        import os
        if self.rank == 0:
            with open(path, "wb") as f:
                f.truncate(self.total_bytes)
        if self.world > 1:
            dist.barrier(group=self.group)
        fd = os.open(path, os.O_WRONLY)
        try:
            if self.nb > 2:
                os.pwrite(fd, self.piece[1: self.nb - 1].cpu().numpy().tobytes(), self.offset + 1)
            if self.rank == 0:
                for k, v in sorted(self._fixups().items()):
                    os.pwrite(fd, bytes([v]), k)
        finally:
            os.close(fd)
        if self.world > 1:
            dist.barrier(group=self.group)
        return self.total_bytes

    def gather(self):
        rank, world, group, device = self.rank, self.world, self.group, self.device
        shifted, nb, offs, nbs, alle, meta, o = self.piece, self.nb, self.offs, self.nbs, self.alle, self.meta, self.end_bit
        t0 = time.perf_counter()
        # 3. interiors: point to point into the output
        ops, out = [], None
        if rank == 0:
            out = torch.empty(self.total_bytes, dtype=torch.uint8, device=device)
            for r in range(1, world):
                if nbs[r] > 2:
                    b0 = offs[r] // 8
                    ops.append(dist.P2POp(dist.irecv, out[b0 + 1: b0 + nbs[r] - 1], r, group))
            if nbs[0] > 2:
                b0 = offs[0] // 8
                out[b0 + 1: b0 + nbs[0] - 1] = shifted[1: nbs[0] - 1]
        elif nb > 2:
            ops.append(dist.P2POp(dist.isend, shifted[1: nb - 1].contiguous(), 0, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        t0 = _tick("p2p_place", t0)
        if rank != 0:
            return None
        # 4. header, edge bytes, trailer
        fix = self._fixups()
        idx = torch.tensor(list(fix.keys()), dtype=torch.int64, device=device)
        val = torch.tensor(list(fix.values()), dtype=torch.uint8, device=device)
        out[idx] = val
        _tick("edges_trailer", t0)
        return out


def compress_sharded(encode_range, nblocks, level, device, group=None):
    """encode_range(first, count) -> (uint8 tensor fragment starting at bit 0, nbits, [block crcs]).
    Returns the complete .bz2 stream as a uint8 tensor on rank 0 (None elsewhere)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    first, count = block_range(nblocks, rank, world)
    t0 = time.perf_counter()
    frag, nbits, crcs = encode_range(first, count)
    _tick("encode_range", t0)
    return place_fragments(frag, nbits, count, crcs, level, device, group)


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


# ---- sharded input ---------------------------------------------------------------------------------
# Host mirror of the RLE1 scan state of csrc/rle1.cu (rs_make / rs_combine): bit 63 non-empty, bit 62 "one run",
# first byte << 24, last byte << 16, trailing run length mod 255 << 8, length mod 255.
def outfresh(c):
    """RLE1 bytes produced by c bytes of one run consumed from a fresh state (lib/Bzip2.js:636-667)."""
    q, r = divmod(c, 255)
    return 5 * q + (r if r <= 3 else 5)


def rs_combine(A, B):
    if not (A >> 63):
        return B
    if not (B >> 63):
        return A
    a_all, b_all = (A >> 62) & 1, (B >> 62) & 1
    a_fc, a_lc, a_tr, a_len = (A >> 24) & 255, (A >> 16) & 255, (A >> 8) & 255, A & 255
    b_fc, b_lc, b_tr, b_len = (B >> 24) & 255, (B >> 16) & 255, (B >> 8) & 255, B & 255
    join = a_lc == b_fc
    trail = (a_tr + b_len) % 255 if (b_all and join) else b_tr
    return (1 << 63) | ((a_all & b_all & int(join)) << 62) | (a_fc << 24) | (b_lc << 16) | (trail << 8) | ((a_len + b_len) % 255)


def share_plan_inputs(summaries, level):
    """summaries[r] = (state, lead, w_fresh, n) of rank r's share (b2_bzip2_share_summary), in rank order.
    Returns ([(state_in, w_in, first, count, raw_offset)] per rank, total blocks, total RLE1 bytes).
    A block belongs to the rank in whose share it starts: block k starts behind the byte that completes k * blockSize
    RLE1 bytes (if no run-phase slip happened before it -- the ranks verify that afterwards)."""
    BS = level * 100000 - 19
    st, W, g = 0, 0, 0
    ins = []
    for (state, lead, w_fresh, n) in summaries:
        ins.append((st, W, g))
        if n:
            c = ((st >> 8) & 255) if (st >> 63) and ((st >> 16) & 255) == ((state >> 24) & 255) else 0
            W += outfresh(c + lead) - outfresh(c) + (w_fresh - outfresh(lead))
            st = rs_combine(st, state)
            g += n
    total = (W + BS - 1) // BS

    def before(w):      # blocks that start at or before the raw position whose RLE1 prefix is w
        return 0 if w == 0 else min(total, w // BS + 1)
    firsts = [before(w) for (_, w, _) in ins] + [total]
    res = []
    for r, (st_in, w_in, g0) in enumerate(ins):
        n = summaries[r][3]
        first = firsts[r]
        count = (firsts[r + 1] - first) if n else 0
        res.append((st_in, w_in, first, count, g0))
    return res, total, W


def _i64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _u64(v):
    return v + (1 << 64) if v < 0 else v


def compress_shares(d_buf, share_len, level=9, group=None, host_out=None, keep_sharded=False):
    """Whole-file bzip2 encode when every rank holds only ITS share of the input: d_buf = CUDA uint8 tensor with the
    share (share_len bytes) followed by a halo (the first bytes of the next shares; empty on the last rank).  The shares
    are contiguous in rank order.  Returns the stream on rank 0 (None elsewhere); with keep_sharded a ShardedStream on
    every rank (the output stays sharded like the input; .gather() assembles it on rank 0)."""
    from . import _native
    L = _native.lib()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = d_buf.device
    PHASES.clear()
    if d_buf.is_cuda:
        torch.cuda.current_stream().synchronize()   # the library works on its own stream: the input must have landed
    t0 = time.perf_counter()
    summ = (C.c_uint64 * 4)()
    rc = L.b2_bzip2_share_summary(d_buf.data_ptr(), share_len, summ)
    if rc:
        raise RuntimeError("b2_bzip2_share_summary: " + _native.last_error())
    mine = torch.tensor([_i64(int(v)) for v in summ], dtype=torch.int64, device=dev)
    if world > 1:
        allv = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(allv, mine, group=group)
    else:
        allv = [mine]
    summaries = [tuple(_u64(int(x)) for x in v.tolist()) for v in allv]
    plan, total, _ = share_plan_inputs(summaries, level)
    st_in, w_in, first, count, g0 = plan[rank]
    n_total = sum(sm[3] for sm in summaries)
    t0 = _tick("share_summaries", t0)
    info = (C.c_uint64 * 6)()
    rc = L.b2_bzip2_plan_share(d_buf.data_ptr(), d_buf.numel(), level, st_in, w_in, first, count, info)
    if rc:
        raise RuntimeError("b2_bzip2_plan_share: " + _native.last_error())
    row = [int(v) for v in info]
    mine = torch.tensor([row[0] + g0, row[1] + g0, row[2], row[3], row[4], total], dtype=torch.int64, device=dev)
    if world > 1:
        alli = [torch.zeros(6, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(alli, mine, group=group)
    else:
        alli = [mine]
    infos = [tuple(int(x) for x in v.tolist()) for v in alli]
    _tick("plan_share+verify", t0)
    if not spec_plan_ok(infos, n_total):
        # a run-phase slip or a block longer than the halo: every rank gets the whole input and the exact plan decides
        PHASES["fallback_full_input"] = 1.0
        lens = [sm[3] for sm in summaries]
        parts = [torch.empty(max(ln, 1), dtype=torch.uint8, device=dev) for ln in lens]
        maxlen = max(lens + [1])
        pad = torch.zeros(maxlen, dtype=torch.uint8, device=dev)
        pad[:share_len] = d_buf[:share_len]
        if world > 1:
            gl = [torch.empty(maxlen, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(gl, pad, group=group)
            full = torch.cat([gl[r][: lens[r]] for r in range(world)])
        else:
            full = d_buf[:share_len]
        del parts
        return compress_file_sharded(full, level, group)
    t0 = time.perf_counter()
    enc = _range_encoder(L, d_buf, d_buf.numel(), level)
    frag, nbits, crcs = enc(first, count)
    _tick("encode_range", t0)
    return place_fragments(frag, nbits, count, crcs, level, dev, group, host_out, keep_sharded)


def share_bounds(n, rank, world, halo):
    """Equal contiguous shares of n bytes: (first byte, share length, bytes to hold = share + halo)."""
    g0 = rank * n // world
    g1 = (rank + 1) * n // world
    return g0, g1 - g0, min(n, g1 + halo) - g0


# ---- sharded decode -------------------------------------------------------------------------------
def decode_shard_rows(L, d_in, rank, world):
    """Stage 1 on one rank: returns (info, rows) -- rows = int64 tensor [own candidates, 6] on the CPU."""
    from . import _native
    if d_in.is_cuda:
        torch.cuda.current_stream().synchronize()   # the library works on its own stream: the input must have landed
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
        # every rank scanned the same stream: the candidate counts must agree (a cheap guard against mismatched collectives)
        chk = torch.tensor([total, d_in.numel()], dtype=torch.int64, device=device)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk, group=group)
        if any(int(v[0]) != total or int(v[1]) != d_in.numel() for v in allc):
            raise RuntimeError("decompress_file_sharded: the ranks do not hold the same stream")
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
    # the shards travel to rank 0 straight into their place in the decoded stream (send/recv into offset views)
    ops, full = [], None
    if rank == 0:
        full = torch.empty(int(allv[0][4]), dtype=torch.uint8, device=device)
        o0, l0 = int(allv[0][2]), int(allv[0][3])
        full[o0: o0 + l0] = out[:l0]
        for r in range(1, world):
            o, ln = int(allv[r][2]), int(allv[r][3])
            if ln:
                ops.append(dist.P2POp(dist.irecv, full[o: o + ln], r, group))
    elif out.numel():
        ops.append(dist.P2POp(dist.isend, out.contiguous(), 0, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return full
