"""GPU: block-range encode entry point (b2_bzip2_encode_range_dev) and the sharded assembly --
several simulated ranks in one process must reproduce the single-call stream bit for bit."""
import pytest
import torch

from oracle import oracle as O
from tests import util as T

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_range_fragments_assemble_to_the_reference_stream(world):
    from compressjs_b200 import sharded as S
    data = T.texty(5 * 99981 + 777, 21)
    level = 1
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    enc, nblocks = S.gpu_encode_range_fn(d_in, level)
    exp = O.bzip2_compress(data, level)
    frags, bits, crcs = [], [], []
    for r in range(world):
        first, count = S.block_range(nblocks, r, world)
        f, nb, cr = enc(first, count)
        off = 32 + sum(bits)
        frags.append(S.shift_right_bits(f, nb, off % 8))
        bits.append(nb)
        crcs.append(cr)
    out = S.assemble(level, frags, bits, crcs, d_in.device)
    assert bytes(out.cpu().numpy().tobytes()) == exp


def test_compress_file_sharded_single_rank():
    from compressjs_b200 import sharded as S
    data = T.ascii_random(3 * 899981 // 2, 4)
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    out = S.compress_file_sharded(d_in, 9)
    assert bytes(out.cpu().numpy().tobytes()) == O.bzip2_compress(data, 9)


@pytest.mark.parametrize("kind", ["ascii", "runs"])
def test_speculative_range_plan_chains_or_is_rejected(kind):
    """b2_bzip2_plan_spec for simulated ranks: when the pieces chain (spec_plan_ok) the assembled stream must be
    the reference stream; run-heavy data must be detected as a failed speculation, never silently mis-cut."""
    import ctypes as C
    from compressjs_b200 import sharded as S, _native
    L = _native.lib()
    data = T.ascii_random(7 * 99981 + 1234, 31) if kind == "ascii" else T.runs(6 * 99981, 32)
    level, world = 1, 4
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    exp = O.bzip2_compress(data, level)
    infos, frags, bits, crcs = [], [], [], []
    for r in range(world):
        info = (C.c_uint64 * 6)()
        assert L.b2_bzip2_plan_spec(d_in.data_ptr(), len(data), level, r, world, info) == 0, _native.last_error()
        infos.append(tuple(int(v) for v in info))
        f, nb, cr = S._range_encoder(L, d_in, len(data), level)(infos[-1][2], infos[-1][4])
        frags.append(f); bits.append(nb); crcs.append(cr)
    ok = S.spec_plan_ok(infos, len(data))
    if kind == "ascii":
        assert ok
    if ok:
        sh, acc = [], 0
        for r in range(world):
            sh.append(S.shift_right_bits(frags[r], bits[r], (32 + acc) % 8))
            acc += bits[r]
        out = S.assemble(level, sh, bits, crcs, d_in.device)
        assert bytes(out.cpu().numpy().tobytes()) == exp


@pytest.mark.parametrize("world", [1, 3])
def test_sharded_decode_stages(world):
    """b2_dec_shard_open/export/finish with simulated ranks: the per-rank outputs concatenate to the input."""
    import bz2
    from compressjs_b200 import sharded as S, _native
    L = _native.lib()
    data = T.texty(7 * 99981 + 321, 41)
    z = bz2.compress(data, 1)
    d_in = torch.frombuffer(bytearray(z), dtype=torch.uint8).cuda()
    # every simulated rank needs its own session: stage 1 rows first (sessions are one at a time, so redo stage 1 per rank)
    rows_all = []
    for r in range(world):
        (total, lo, hi), rows = S.decode_shard_rows(L, d_in, r, world)
        rows_all.append(rows)
    all_rows = torch.cat(rows_all)
    assert all_rows.shape[0] == total
    out = bytearray(len(data))
    for r in range(world):
        S.decode_shard_rows(L, d_in, r, world)  # reopen this rank's session
        o, res = S.decode_shard_finish(L, all_rows, False, d_in.device)
        assert o is not None, res
        assert res["total"] == len(data)
        out[res["off"]: res["off"] + res["len"]] = bytes(o.cpu().numpy().tobytes())
    assert bytes(out) == data


def test_sharded_decode_single_process_api():
    import bz2
    from compressjs_b200 import sharded as S
    data = T.ascii_random(3 * 99981 + 77, 42)
    d_in = torch.frombuffer(bytearray(bz2.compress(data, 1)), dtype=torch.uint8).cuda()
    assert bytes(S.decompress_file_sharded(d_in).cpu().numpy().tobytes()) == data


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("kind", ["ascii", "text", "runs"])
def test_sharded_input_shares_reproduce_the_stream(kind, world):
    """Every simulated rank holds only its share of the input + a halo (b2_bzip2_share_summary / b2_bzip2_plan_share):
    when the pieces chain up, the assembled stream must be the reference stream; run-heavy data must be rejected by the
    chain check, never silently mis-cut."""
    import ctypes as C
    from compressjs_b200 import sharded as S, _native
    L = _native.lib()
    level = 1
    n = 9 * 99981 + 777
    data = {"ascii": T.ascii_random, "text": T.texty, "runs": T.runs}[kind](n, 123 + world)
    exp = O.bzip2_compress(data, level)
    halo = 150000
    bufs, summaries = [], []
    for r in range(world):
        g0, ln, hold = S.share_bounds(n, r, world, halo)
        d = torch.frombuffer(bytearray(data[g0:g0 + hold]), dtype=torch.uint8).cuda()
        sm = (C.c_uint64 * 4)()
        assert L.b2_bzip2_share_summary(d.data_ptr(), ln, sm) == 0, _native.last_error()
        bufs.append((d, ln, g0))
        summaries.append(tuple(int(v) for v in sm))
    plan, total, _ = S.share_plan_inputs(summaries, level)
    infos, frags, bits, crcs = [], [], [], []
    for r in range(world):
        d, ln, g0 = bufs[r]
        st_in, w_in, first, count, off = plan[r]
        assert off == g0
        info = (C.c_uint64 * 6)()
        assert L.b2_bzip2_plan_share(d.data_ptr(), d.numel(), level, st_in, w_in, first, count, info) == 0, _native.last_error()
        row = [int(v) for v in info]
        infos.append((row[0] + g0, row[1] + g0, row[2], row[3], row[4], total))
        f, nb, cr = S._range_encoder(L, d, d.numel(), level)(first, row[4])
        frags.append(f); bits.append(nb); crcs.append(cr)
    ok = S.spec_plan_ok(infos, n)
    if kind != "runs":
        assert ok
    if ok:
        sh, o = [], 32
        for f, nb in zip(frags, bits):
            sh.append(S.shift_right_bits(f, nb, o % 8))
            o += nb
        out = S.assemble(level, sh, bits, crcs, frags[0].device)
        assert bytes(out.cpu().numpy().tobytes()) == exp
