"""GPU parity at the scales and seams the benchmark runs at ("parity at benchmark scale"):

* a seeded fuzz of small inputs (0..6000 bytes, alphabets of 1..256 symbols) through b2_bzip2_compress: table-count
  thresholds (lib/Bzip2.js:826-830), one- and two-selector blocks, tiny alphabets;
* many 900k blocks at -9 with a small BWT batch, ascii -> text -> ascii, so that batch seams, the MSD path, the
  8-byte ("wide") mode and the hand-over of the mode between batches (bwt.cu) all run against the oracle;
* streams with thousands of tiny blocks / members (candidate buffer of the magic scan).
"""
import bz2
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
from tests import util as T

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fuzz_case(g, k):
    n = int(g.integers(0, 6001)) if k % 7 else int(g.choice([0, 1, 2, 3, 4, 5, 49, 50, 51, 199, 200, 599, 600, 1199, 1200, 2399, 2400, 2401]))
    a = int(g.choice([1, 2, 3, 4, 8, 16, 64, 95, 200, 256]))
    syms = g.permutation(256)[:a].astype(np.uint8)
    kind = k % 4
    if kind == 0:      # uniform over the alphabet
        d = syms[g.integers(0, a, size=n)]
    elif kind == 1:    # skewed (geometric) -> long MTF zero runs
        d = syms[np.minimum(g.geometric(0.45, size=n) - 1, a - 1)]
    elif kind == 2:    # runs (RLE1) of random lengths
        out = bytearray()
        while len(out) < n:
            out += bytes([int(syms[int(g.integers(0, a))])]) * int(g.choice([1, 1, 2, 3, 4, 5, 6, 255, 256, 300]))
        d = np.frombuffer(bytes(out[:n]), dtype=np.uint8)
    else:              # periodic with a random period
        p = max(1, int(g.integers(1, 40)))
        d = np.resize(syms[g.integers(0, a, size=p)], n)
    return d.tobytes()


def test_small_input_fuzz_vs_oracle():
    from compressjs_b200 import Bzip2
    g = T.rng(777)
    for k in range(300):
        d = _fuzz_case(g, k)
        level = int(g.integers(1, 10))
        got = Bzip2.compressFile(d, None, level)
        exp = O.bzip2_compress(d, level)
        assert got == exp, "case %d: n=%d level=%d" % (k, len(d), level)
        if k % 10 == 0:
            assert Bzip2.decompressFile(got) == d


_SEAM_SCRIPT = r"""
import sys, hashlib
sys.path.insert(0, %(root)r)
from compressjs_b200 import Bzip2, _native
from tests import util as T
bs = 899981
data = T.ascii_random(8 * bs + 1000, 41) + T.texty(8 * bs, 42) + T.ascii_random(4 * bs - 5000, 43)
z = Bzip2.compressFile(data, None, 9)
st = _native.stats()
print("RESULT", len(z), hashlib.sha256(z).hexdigest(), st["blocks"], st["msd_launches"], st["radix_launches"])
"""


def test_batch_seams_and_mode_handover_at_level_9():
    """20 blocks of 900k with B2_BWT_BATCH=8 (three batches; the middle one is text)."""
    import hashlib
    env = dict(os.environ, B2_BWT_BATCH="8")
    r = subprocess.run([sys.executable, "-c", _SEAM_SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][-1].split()
    bs = 899981
    data = T.ascii_random(8 * bs + 1000, 41) + T.texty(8 * bs, 42) + T.ascii_random(4 * bs - 5000, 43)
    exp = O.bzip2_compress(data, 9, threads=min(os.cpu_count() or 1, 20))
    assert (int(line[1]), line[2]) == (len(exp), hashlib.sha256(exp).hexdigest())
    assert int(line[3]) >= 20
    assert int(line[4]) >= 1 and int(line[5]) >= 4   # both the MSD path and the LSD passes ran


def test_thousands_of_tiny_blocks_and_members():
    from compressjs_b200 import Bzip2
    # periodic input at level 1: ~36 compressed bytes per 100k block, far more magics than compressed_size / 8000
    data = b"abc" * (1100 * 99981 // 3)
    z = Bzip2.compressFile(data, None, 1)
    assert len(z) < 1100 * 200
    assert Bzip2.decompressFile(z) == data
    assert bz2.decompress(z) == data
    # multistream: 1500 tiny members
    member = bz2.compress(b"hello, world\n", 9)
    cat = member * 1500
    assert Bzip2.decompressFile(cat, None, True) == b"hello, world\n" * 1500


def test_decode_batch_seams_and_reclassified_blocks():
    """Streams of more than one decode batch, and of more blocks than the decoder keeps count-byte classes for (production:
    2048 / 16384 blocks; $B2_DEC_BATCH / $B2_DEC_KEEP_CLS shrink both): the batch seams and the second classification pass
    must not change a byte.  Runs in a child process because the library reads the hooks per call but the tests share it."""
    code = r"""
import bz2, sys
sys.path.insert(0, %r)
from tests import util as T
from compressjs_b200 import Bzip2
data = T.runs(1300000, 61) + T.texty(900000, 62) + b"z" * 300000 + T.ascii_random(700000, 63) + bytes(range(256)) * 1200
z1 = bz2.compress(data, 1)                       # ~35 blocks of 100k
assert Bzip2.decompressFile(z1) == data
z2 = Bzip2.compressFile(data, None, 2)
assert Bzip2.decompressFile(z2) == data
cat = z1 + bz2.compress(b"tail member " * 5000, 3)
assert Bzip2.decompressFile(cat, None, True) == data + b"tail member " * 5000
rows = []
Bzip2.table(z1, lambda pos, size: rows.append((pos, size)))
assert sum(sz for _, sz in rows) == len(data), (sum(sz for _, sz in rows), len(data))
assert len(rows) >= 15, len(rows)
print("ok", len(rows))
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B2_DEC_BATCH="7", B2_DEC_KEEP_CLS="5")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


@pytest.mark.parametrize("level,window,pinned", [(1, 4 << 20, True), (9, 4 << 20, False), (5, 3 << 20, True)])
def test_streaming_windows_reproduce_the_one_shot_stream(level, window, pinned, monkeypatch):
    """b2_bzip2_compress with an input larger than its streaming window (B2_STREAM_WINDOW; production default 8 GiB): the
    input passes through the device window by window, the output window is drained and rebased in between -- the stream
    must be the reference stream (lib/Bzip2.js:879-929 reads its input strictly forward, block cuts do not care)."""
    import ctypes as C
    import torch
    from compressjs_b200 import _native
    L = _native.lib()
    monkeypatch.setenv("B2_STREAM_WINDOW", str(window))
    monkeypatch.setenv("B2_H2D_CHUNK", str(1 << 20))
    data = T.ascii_random(9 << 20, 51) + T.runs(3 << 20, 52) + T.texty(7 << 20, 53) + b"q" * 700000 + T.ascii_random(2500000, 54)
    if pinned:
        buf = torch.empty(len(data), dtype=torch.uint8, pin_memory=True)
        buf.numpy()[:] = np.frombuffer(data, dtype=np.uint8)
        ptr = buf.data_ptr()
    else:
        arr = np.frombuffer(data, dtype=np.uint8)
        ptr = arr.ctypes.data
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    rc = L.b2_bzip2_compress(ptr, len(data), level, C.byref(out), C.byref(n))
    assert rc == 0, _native.last_error()
    got = bytes(np.ctypeslib.as_array(out, (n.value,)))
    L.b2_free(out)
    exp = O.bzip2_compress(data, level, threads=min(os.cpu_count() or 1, 16))
    assert got == exp
    tr = _native.last_trace()
    assert sum(t.raw_len for t in tr) == len(data) and tr[0].raw_start == 0 and tr[0].bit_start == 32
    assert all(a.bit_start + a.bit_len == b.bit_start for a, b in zip(tr, tr[1:]))
