"""CPU tests (gloo, world_size 2 and 3) of the multi-GPU host logic in compressjs_b200/sharded.py:
block-range partition, fragment bit lengths exchange, shift to the global bit phase, byte gather with
OR-ed boundary bytes, header/trailer and stream CRC.  The per-range encoder is injected (the oracle
cut into block ranges), so no GPU is needed; the assembled stream must equal the single-process one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from tests import util as T
from compressjs_b200 import sharded as S


def _bits_slice(stream, start, nbits):
    a = np.unpackbits(np.frombuffer(stream, dtype=np.uint8))[start:start + nbits]
    pad = (-len(a)) % 8
    return np.packbits(np.concatenate([a, np.zeros(pad, dtype=np.uint8)])).tobytes()


def _oracle_encoder(data, level):
    z, tr = O.bzip2_compress(data, level, trace=True)

    def encode_range(first, count):
        if count == 0:
            return torch.zeros(8, dtype=torch.uint8), 0, []
        b0 = tr[first].bit_start
        b1 = tr[first + count - 1].bit_start + tr[first + count - 1].bit_len
        frag = _bits_slice(z, b0, b1 - b0)
        return torch.frombuffer(bytearray(frag) + bytearray(8), dtype=torch.uint8), b1 - b0, [tr[k].crc for k in range(first, first + count)]

    return encode_range, len(tr), z


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, level, n, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = T.texty(n, seed) if n else b""
    enc, nblocks, z = _oracle_encoder(data, level)
    out = S.compress_sharded(enc, nblocks, level, torch.device("cpu"))
    if rank == 0:
        q.put(bytes(out.numpy().tobytes()) == z)
    dist.destroy_process_group()


def _worker_host(rank, world, port, level, n, seed, q):
    """Same, but the stream is assembled in a host buffer shared by the ranks (every rank writes its own piece)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = T.texty(n, seed)
    enc, nblocks, z = _oracle_encoder(data, level)
    shm = S.SharedHostBuffer(len(z) + 64)
    shm.tensor[: len(z) + 64] = 0xA5   # stale bytes of an earlier step must not leak into the stream
    dist.barrier()
    first, count = S.block_range(nblocks, rank, world)
    frag, nbits, crcs = enc(first, count)
    ln = S.place_fragments(frag, nbits, count, crcs, level, torch.device("cpu"), None, shm.tensor)
    if rank == 0:
        q.put(ln == len(z) and bytes(shm.tensor[:ln].numpy().tobytes()) == z)
    dist.barrier()
    shm.close()
    dist.destroy_process_group()


def _worker_kept(rank, world, port, level, n, seed, q, path):
    """The stream is left sharded (ShardedStream): every rank's piece already sits at its final bit position; gather()
    then assembles the same bytes on rank 0."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = T.texty(n, seed)
    enc, nblocks, z = _oracle_encoder(data, level)
    first, count = S.block_range(nblocks, rank, world)
    frag, nbits, crcs = enc(first, count)
    ss = S.place_fragments(frag, nbits, count, crcs, level, torch.device("cpu"), None, None, True)
    ok = ss.total_bytes == len(z)
    piece = ss.piece[: ss.nb].numpy()
    ref = np.frombuffer(z, dtype=np.uint8)[ss.offset: ss.offset + ss.nb]
    if ss.nb > 2:
        ok = ok and bool((piece[1:-1] == ref[1:-1]).all())       # interior bytes are final
    if ss.nb:
        ok = ok and not (piece[0] & ~ref[0]) and not (piece[-1] & ~ref[-1])   # edge bytes: only bits of the stream
    out = ss.gather()
    if rank == 0:
        ok = ok and bytes(out.numpy().tobytes()) == z
    # ... or every rank writes its own byte range of one file
    ln = ss.write_file(path)
    if rank == 0:
        with open(path, "rb") as f:
            ok = ok and ln == len(z) and f.read() == z
    q.put(bool(ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 450000), (3, 520000)])
def test_stream_left_sharded_then_gathered(world, n, tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    path = str(tmp_path / "sharded.bz2")
    procs = [ctx.Process(target=_worker_kept, args=(r, world, port, 1, n, 11, q, path)) for r in range(world)]
    for p in procs:
        p.start()
    oks = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(oks)


@pytest.mark.parametrize("world,n", [(2, 450000), (3, 520000)])
def test_sharded_assembly_in_a_shared_host_buffer(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_host, args=(r, world, port, 1, n, 9, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert ok


@pytest.mark.parametrize("world,n", [(2, 450000), (3, 520000), (2, 99981), (2, 0)])
def test_sharded_assembly_matches_single_stream(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1, n, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert ok


def test_shift_and_trailer_helpers():
    g = T.rng(3)
    for _ in range(50):
        nbits = int(g.integers(1, 300))
        phase = int(g.integers(0, 8))
        bits = g.integers(0, 2, size=nbits, dtype=np.uint8)
        frag = torch.frombuffer(bytearray(np.packbits(bits).tobytes()) + bytearray(4), dtype=torch.uint8)
        out = S.shift_right_bits(frag, nbits, phase).numpy()
        exp = np.packbits(np.concatenate([np.zeros(phase, dtype=np.uint8), bits]))
        assert out.tobytes() == exp.tobytes()
    assert S.fold_stream_crc([0xEA29357D]) == 0xEA29357D
    b0, tb = S.trailer_bytes(0, 0x12345678)
    assert (b0, tb.hex()) == (0, "17724538509012345678")


def test_single_process_path():
    data = T.texty(250000, 9)
    enc, nblocks, z = _oracle_encoder(data, 1)
    out = S.compress_sharded(enc, nblocks, 1, torch.device("cpu"))
    assert out.numpy().tobytes() == z


def test_spec_plan_verification():
    ok = [(0, 100, 0, 2, 2, 5), (100, 250, 2, 3, 3, 5)]
    assert S.spec_plan_ok(ok, 250)
    assert not S.spec_plan_ok([(0, 100, 0, 2, 2, 5), (101, 250, 2, 3, 3, 5)], 250)   # gap between ranks
    assert not S.spec_plan_ok([(0, 100, 0, 2, 2, 5), (100, 240, 2, 3, 3, 5)], 250)   # does not reach the end
    assert not S.spec_plan_ok([(0, 100, 0, 2, 2, 5), (100, 250, 2, 3, 2, 5)], 250)   # a rank cut fewer blocks
    assert not S.spec_plan_ok([(0, 100, 0, 2, 2, 5), (100, 250, 2, 3, 3, 6)], 250)   # totals disagree
    assert S.spec_plan_ok([(0, 250, 0, 1, 1, 1), (0, 0, 1, 0, 0, 1)], 250)           # rank without blocks


# ---- sharded input: share summaries -> run state / RLE1 output in front of every share -> block ownership ----------
def _py_share_summary(b):
    """Plain-Python restatement of b2_bzip2_share_summary for a share `b` (bytes): (state, lead, w_fresh, n)."""
    n = len(b)
    if n == 0:
        return (0, 0, 0, 0)
    lead = 1
    while lead < n and b[lead] == b[0]:
        lead += 1
    trail = 1
    while trail < n and b[n - 1 - trail] == b[n - 1]:
        trail += 1
    allsame = int(lead == n)
    w, i = 0, 0
    while i < n:
        j = i
        while j < n and b[j] == b[i]:
            j += 1
        w += S.outfresh(j - i)
        i = j
    state = (1 << 63) | (allsame << 62) | (b[0] << 24) | (b[n - 1] << 16) | ((trail % 255) << 8) | (n % 255)
    return (state, lead, w, n)


@pytest.mark.parametrize("kind", ["ascii", "runs", "text"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_share_plan_inputs_match_the_sequential_block_cut(kind, world):
    """The blocks a rank is told to cut (first, count) must be exactly the blocks of the reference walk
    (lib/Bzip2.js:636-667, oracle rle1_split) that start inside its share -- whenever the W-space speculation holds,
    i.e. when block k really starts behind the byte completing k * blockSize RLE1 bytes."""
    level = 1
    n = 7 * 99981 + 4321
    data = {"ascii": T.ascii_random, "runs": T.runs, "text": T.texty}[kind](n, 77)
    starts, lens, _, _ = O.rle1_split(data, level)
    bounds = [(r * n // world, (r + 1) * n // world) for r in range(world)]
    summaries = [_py_share_summary(data[a:b]) for a, b in bounds]
    plan, total, W = S.share_plan_inputs(summaries, level)
    # W must be the RLE1 output of the whole input under maximal-run phases
    assert W == _py_share_summary(data)[2]
    BS = level * 100000 - 19
    assert total == (W + BS - 1) // BS
    spec_ok = total == len(starts)
    nxt = 0
    for r, (st_in, w_in, first, count, g0) in enumerate(plan):
        assert g0 == bounds[r][0] and first == nxt
        assert w_in == _py_share_summary(data[:g0])[2]
        nxt = first + count
        if spec_ok and kind != "runs":
            mine = [k for k in range(len(starts)) if (bounds[r][0] < int(starts[k]) <= bounds[r][1]) or (k == 0 and r == 0)]
            assert mine == list(range(first, first + count)), (r, mine, first, count)
    assert nxt == total
