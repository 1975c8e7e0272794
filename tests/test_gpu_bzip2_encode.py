"""GPU parity: Bzip2.compressFile (lib/Bzip2.js:879-929) -- CUDA stream bytes == oracle, bit exact,
and every stream decodes with libbz2 and with the oracle's decoder."""
import bz2
import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from tests import util as T

pytestmark = pytest.mark.gpu


def _check(data, level, libbz2=True):
    from compressjs_b200 import Bzip2, _native
    got = Bzip2.compressFile(data, None, level)
    exp, tr = O.bzip2_compress(data, level, trace=True)
    if got != exp:
        gt = _native.last_trace()
        msg = ["stream mismatch: got %d bytes, expected %d; blocks got %d expected %d" % (len(got), len(exp), len(gt), len(tr))]
        for k, (a, b) in enumerate(zip(gt, tr)):
            fa = dict((f, getattr(a, f)) for f, _ in a._fields_)
            fb = dict((f, getattr(b, f)) for f, _ in b._fields_)
            diff = {f: (fa[f], fb[f]) for f in fa if fa[f] != fb[f] and f != "pad"}
            if diff:
                msg.append("block %d differs: %r" % (k, diff))
                break
        first = next((i for i in range(min(len(got), len(exp))) if got[i] != exp[i]), None)
        msg.append("first differing byte: %r" % first)
        raise AssertionError("\n".join(msg))
    if libbz2:
        assert bz2.decompress(got) == bytes(data)
    return got


def test_sample0_config1():
    # BASELINE config 1: "This is a test\n", level 1 -> the 57-byte stream of SURVEY.md Appendix C
    got = _check(b"This is a test\n", 1)
    assert got.hex() == ("425a6831314159265359ea29357d000002538000104000040022600c00200021aa8f6f4a9ef5"
                         "0806059702fb58a70bb9229c284875149abe80")


def test_empty_and_tiny():
    from compressjs_b200 import Bzip2
    assert Bzip2.compressFile(b"", None, 9) == O.bzip2_compress(b"", 9)
    assert len(Bzip2.compressFile(b"", None, 9)) == 14
    for d in (b"a", b"ab", b"aaaa", b"aaaaa", b"\x00" * 300, bytes(range(256))):
        _check(d, 9)


@pytest.mark.parametrize("level", [1, 5, 9])
@pytest.mark.parametrize("kind,n", [("ascii", 250000), ("text", 1200000), ("runs", 400000)])
def test_synthetic(kind, n, level):
    data = {"ascii": T.ascii_random, "text": T.texty, "runs": T.runs}[kind](n, seed=n + level)
    _check(data, level)


def test_rle1_block_boundary_quirks():
    # block fills on the 4th byte of a run: no count byte is written (libbz2 rejects this stream,
    # the reference's own decoder accepts it -- SURVEY.md section 7)
    base = T.ascii_random(99977, 3).replace(b"aaaa", b"abab")
    data = base + b"a" * 20 + T.ascii_random(1000, 4)
    _check(data, 1, libbz2=False)
    # block ends right after a zero count byte
    data = T.ascii_random(99976, 5) + b"b" * 300 + T.ascii_random(500, 6)
    _check(data, 1, libbz2=False)
    # long runs straddling block boundaries, runs of 255/256/600
    data = (b"x" * 70000 + b"y" * 255 + b"z" * 256 + T.ascii_random(29000, 7) + b"w" * 200000 + b"v" * 4 + b"u" * 5) * 2
    _check(data, 1, libbz2=False)
    _check(b"q" * 5000000, 1, libbz2=False)


@pytest.mark.parametrize("name", ["sample0", "sample1", "sample2", "sample3", "sample4", "sample5"])
@pytest.mark.parametrize("level", [1, 9])
def test_reference_samples(name, level):
    data = T.fixture(name + ".ref")
    got = _check(data, level)
    gold = T.golden().get("bzip2_%s_-%d" % (name, level))
    if gold:
        assert (len(got), hashlib.sha256(got).hexdigest()) == (gold["size"], gold["sha256"])


def test_multi_batch_and_trace():
    from compressjs_b200 import Bzip2, _native
    data = T.texty(3 * 99981 + 5000, 11) + T.ascii_random(2 * 99981, 12)
    _check(data, 1)
    st = _native.stats()
    assert st["blocks"] >= 5 and st["kernel_launches"] > 20


@pytest.mark.parametrize("skew", [1, 7, 255])
def test_device_entry_with_unaligned_input_pointer(skew):
    """b2_bzip2_compress_dev / b2_crc32 on a device pointer that is not 16-byte aligned (the CRC pieces are cut
    on address-aligned windows; lib/CRC32.js:72-103 has no such notion, the result must not depend on it)."""
    import ctypes as C
    import torch
    from compressjs_b200 import _native
    L = _native.lib()
    data = T.texty(250000, 5) + T.runs(70000, 6) + T.ascii_random(123457, 7)
    buf = torch.zeros(len(data) + 512, dtype=torch.uint8, device="cuda")
    buf[skew:skew + len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    cap = L.b2_bzip2_bound(len(data))
    out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    out_n = C.c_size_t()
    rc = L.b2_bzip2_compress_dev(buf.data_ptr() + skew, len(data), 1, out.data_ptr(), cap, C.byref(out_n))
    assert rc == 0, _native.last_error()
    got = out[:out_n.value].cpu().numpy().tobytes()
    assert got == O.bzip2_compress(data, 1)
    # decode into an unaligned output pointer as well (per-block CRC check of the decoded ranges)
    dec = torch.zeros(len(data) + 512, dtype=torch.uint8, device="cuda")
    dn = C.c_size_t()
    rc = L.b2_bzip2_decompress_dev(out.data_ptr(), out_n.value, 0, dec.data_ptr() + skew, len(data), C.byref(dn))
    assert rc == 0 and dn.value == len(data), _native.last_error()
    assert dec[skew:skew + len(data)].cpu().numpy().tobytes() == data


@pytest.mark.parametrize("chunk", [4096, 150000, 1 << 20])
def test_host_entry_pipelined_upload(chunk, monkeypatch):
    """b2_bzip2_compress with a PINNED input uploads in chunks and encodes the blocks that are final in the
    prefix that has arrived (block cuts only depend on earlier bytes, lib/Bzip2.js:636-667); the stream must be
    the one-shot stream whatever the chunking."""
    import ctypes as C
    import torch
    from compressjs_b200 import _native
    L = _native.lib()
    monkeypatch.setenv("B2_H2D_CHUNK", str(chunk))
    data = T.texty(330000, 21) + T.runs(250000, 22) + T.ascii_random(420001, 23) + b"z" * 70000
    pinned = torch.empty(len(data), dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = np.frombuffer(data, dtype=np.uint8)
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    rc = L.b2_bzip2_compress(pinned.data_ptr(), len(data), 1, C.byref(out), C.byref(n))
    assert rc == 0, _native.last_error()
    got = bytes(np.ctypeslib.as_array(out, (n.value,)))
    L.b2_free(out)
    assert got == O.bzip2_compress(data, 1)
    tr = _native.last_trace()
    assert sum(t.raw_len for t in tr) == len(data) and tr[0].raw_start == 0
    assert all(a.raw_start + a.raw_len == b.raw_start for a, b in zip(tr, tr[1:]))


@pytest.mark.parametrize("kind", ["ascii", "text", "slips"])
def test_many_blocks_parallel_block_walk(kind):
    """Files of >= 16 blocks cut their blocks with several CTAs from speculated boundaries and accept the cut only
    if the segments chain up (else one CTA walks again): the stream must equal the sequential reference walk
    (lib/Bzip2.js:636-667) with and without run-phase slips at block ends."""
    if kind == "ascii":
        data = T.ascii_random(2600000, 31)
    elif kind == "text":
        data = T.texty(3100000, 32)
    else:
        # runs that straddle many block boundaries (blocks filling on a run's 4th byte shift every later boundary)
        parts = []
        for i in range(40):
            parts.append(T.ascii_random(99000 + 37 * i, 100 + i).replace(b"aaaa", b"abab"))
            parts.append(bytes([65 + i % 26]) * (900 + 13 * i))
        data = b"".join(parts)
    _check(data, 1, libbz2=(kind != "slips"))
