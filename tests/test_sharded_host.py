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
