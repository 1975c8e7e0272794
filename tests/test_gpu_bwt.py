"""GPU parity: forward cyclic BWT (lib/BWT.js:372-417 bwtransform2) -- CUDA vs oracle, bit exact."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import util as T

pytestmark = pytest.mark.gpu

KATS = [  # test/bwtest.js:39-79
    ("bcababa", "cbbaaab", 5),
    ("ABCDEFGHIJKLMNOPQRSTUVWXYZ", "ZABCDEFGHIJKLMNOPQRSTUVWXY", 0),
    ("ZYXWVUTSRQPONMLKJIHGFEDCBA", "BCDEFGHIJKLMNOPQRSTUVWXYZA", 25),
    ("SIX.MIXED.PIXIES.SIFT.SIXTY.PIXIE.DUST.BOXES", "TEXYDST.E.IXIXIXXSSMPPS.B..E.S.EUSFXDIIOIIIT", 29),
]


def test_bwt_kats():
    for i, o, p in KATS:
        assert T.native_bwt(i.encode()) == (o.encode(), p)
    mary = ("Mary had a little lamb, its fleece was white as snow" * 8 + "Nary had a little lamb, its fleece was white as snow").encode()
    u, p = T.native_bwt(mary)
    assert p == 99 and (u, p) == O.bwt_cyclic(mary)


@pytest.mark.parametrize("data", [b"ab", b"aa", b"aaaa", b"abababab", b"abcabcabc", b"a" * 1000, b"ab" * 777, b"\x00" * 300 + b"\xff" * 300,
                                  bytes(range(256)) * 3, b"abc" * 341 + b"abd"])
def test_bwt_degenerate(data):
    assert T.native_bwt(data) == O.bwt_cyclic(data)


@pytest.mark.parametrize("n,kind", [(1000, "ascii"), (4096, "ascii"), (4097, "text"), (70000, "text"), (100000, "runs"), (300000, "ascii")])
def test_bwt_random(n, kind):
    data = {"ascii": T.ascii_random, "text": T.texty, "runs": T.runs}[kind](n, seed=n)
    assert T.native_bwt(data) == O.bwt_cyclic(data)


def test_bwt_full_block_and_batch():
    blocks = [T.ascii_random(899981, 5), T.texty(899981, 6), T.runs(50000, 7), b"x" * 70000, T.texty(12345, 8), b"q",
              (T.texty(3000, 9) * 40)[:100000]]
    got = T.native_bwt_batch(blocks)
    for b, g in zip(blocks, got):
        assert g == O.bwt_cyclic(b)
    st = T.native().stats()
    assert st["radix_launches"] > 0 and st["kernel_launches"] > 0


@pytest.mark.parametrize("n,alpha,tilt", [(300000, 200, 0.02), (899981, 256, 0.01), (450000, 128, 0.03), (250000, 95, 0.0)])
def test_bwt_sparse_ties_on_mildly_skewed_alphabets(n, alpha, tilt):
    """The MSD + shared-memory bucket sort path (bwt_msd.cu) spreads keys over interpolation cells assuming a uniform
    alphabet; a tilted symbol distribution still takes that path (few 5-byte ties) but fills some cells with many records
    (the warp-per-cell ordering) and plants a few repeated 40-byte phrases (tie groups for the resolver)."""
    g = T.rng(n + alpha)
    p = 1.0 / (1.0 + tilt * np.arange(alpha))
    p /= p.sum()
    a = g.choice(alpha, size=n, p=p).astype(np.uint8)
    for _ in range(30):
        src, dst = int(g.integers(0, n - 100)), int(g.integers(0, n - 100))
        a[dst:dst + 40] = a[src:src + 40]
    data = a.tobytes()
    assert T.native_bwt(data) == O.bwt_cyclic(data)
    assert T.native().stats()["msd_launches"] >= 1


def test_bwt_fixture_sample3():
    data = T.fixture("sample3.ref")  # highly repetitive: many doubling rounds
    assert T.native_bwt(data) == O.bwt_cyclic(data)


# ---- the sentinel family: BWT.suffixsort / bwtransform / unbwtransform (lib/BWT.js:305-363) -------------------
SENT_CASES = [b"ab", b"aa", b"ba", b"aaa", b"aaaa", b"abab", b"abcabcabc", b"mississippi", b"\x00", b"\x00\x00", b"\x00\x00\x00\x00\x00",
              b"a\x00", b"\x00a\x00\x00", b"ab\x00\x00ab\x00", b"a" * 1000, b"ab" * 777, b"\x00" * 300 + b"\xff" * 300,
              bytes(range(256)) * 3, b"abc" * 341 + b"abd", b"\x00\x01" * 50 + b"\x00"]


def _sentinel_all(data):
    from compressjs_b200 import BWT
    n = len(data)
    sa = np.zeros(n, dtype=np.int32)
    assert BWT.suffixsort(data, sa, n) == 0
    u = np.zeros(n, dtype=np.uint8)
    p1 = BWT.bwtransform(data, u, None, n)
    back = np.zeros(n, dtype=np.uint8)
    BWT.unbwtransform(u, back, None, n, p1)
    return sa, bytes(u), p1, bytes(back)


@pytest.mark.parametrize("data", SENT_CASES)
def test_sentinel_family_small(data):
    sa, u, p1, back = _sentinel_all(data)
    assert list(sa) == list(O.suffixsort(data))
    assert (u, p1) == O.bwt_sentinel(data)
    assert back == data
    assert O.unbwt_sentinel(u, p1) == data


def test_sentinel_bwtest_kats():
    # test/bwtest.js:39-79 runs its vectors through bwtransform/unbwtransform as well (round trip) -- same here
    for i, _, _ in KATS:
        d = i.encode()
        sa, u, p1, back = _sentinel_all(d)
        assert back == d and (u, p1) == O.bwt_sentinel(d)


@pytest.mark.parametrize("n,kind", [(1, "ascii"), (1000, "ascii"), (4097, "text"), (70000, "text"), (100000, "runs"), (300000, "ascii"),
                                    (900000, "text"), (1048574, "ascii")])
def test_sentinel_family_random(n, kind):
    data = {"ascii": T.ascii_random, "text": T.texty, "runs": T.runs}[kind](n, seed=n + 17)
    sa, u, p1, back = _sentinel_all(data)
    assert back == data
    assert (u, p1) == O.bwt_sentinel(data)
    assert np.array_equal(sa, np.asarray(O.suffixsort(data), dtype=np.int32))


def test_sentinel_limits():
    from compressjs_b200 import BWT
    big = np.zeros(1048575, dtype=np.uint8)
    with pytest.raises(RuntimeError):
        BWT.suffixsort(big, np.zeros(big.size, dtype=np.int32), big.size)
    u = np.zeros(0, dtype=np.uint8)
    assert BWT.bwtransform(b"", u, None, 0) == 0 and BWT.suffixsort(b"", np.zeros(0, dtype=np.int32), 0) == 0
