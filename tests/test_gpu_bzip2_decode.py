"""GPU parity: Bzip2.decompressFile / decompressBlock / table (lib/Bzip2.js:454-548) -- CUDA vs the
reference's own decode fixtures (test/bzip2-basic.js, bzip2-block.js, bzip2-table.js), vs libbz2
streams and vs the oracle's error behaviour."""
import bz2

import numpy as np
import pytest

from oracle import oracle as O
from tests import util as T

pytestmark = pytest.mark.gpu


def B():
    from compressjs_b200 import Bzip2
    return Bzip2


@pytest.mark.parametrize("k", range(5))
def test_decode_reference_fixtures(k):  # test/bzip2-basic.js:5-20
    ref = T.fixture("sample%d.ref" % k)
    z = T.fixture("sample%d.bz2" % k)
    assert B().decompressFile(z) == ref
    assert B().decompressFile(z, len(ref)) == ref  # fixed-size output argument (:16)
    with pytest.raises(TypeError):
        B().decompressFile(z, len(ref) + 1)


def test_decode_block_fixtures():  # test/bzip2-block.js:5-27
    for f, pos, blk in [("sample0", 32, "sample0.ref"), ("sample2", 544888, "sample2.544888"), ("sample4", 32, "sample4.32"),
                        ("sample4", 1596228, "sample4.1596228"), ("sample4", 2342106, "sample4.2342106")]:
        assert B().decompressBlock(T.fixture(f + ".bz2"), pos) == T.fixture(blk)


@pytest.mark.parametrize("k", range(5))
def test_table_fixtures(k):  # test/bzip2-table.js:5-17
    exp = [tuple(map(int, l.split("\t"))) for l in T.fixture("sample%d.bzt" % k).decode().strip().split("\n")]
    rows = []
    B().table(T.fixture("sample%d.bz2" % k), lambda pos, size: rows.append((pos, size)))
    assert rows == exp


@pytest.mark.parametrize("level", [1, 9])
@pytest.mark.parametrize("kind,n", [("ascii", 300000), ("text", 1500000), ("runs", 700000)])
def test_roundtrip_own_and_libbz2_streams(kind, n, level):  # test/file.js:5-46
    data = {"ascii": T.ascii_random, "text": T.texty, "runs": T.runs}[kind](n, seed=n + level)
    z = B().compressFile(data, None, level)
    assert B().decompressFile(z) == data
    zl = bz2.compress(data, level)  # independent producer (libbz2 table search, <= 17 bit codes)
    assert B().decompressFile(zl) == data


def test_quirk_streams_decode():
    # reference-encoder streams that libbz2 rejects: block ending in an uncounted run of four
    data = T.ascii_random(99977, 3).replace(b"aaaa", b"abab") + b"a" * 20 + T.ascii_random(1000, 4)
    z = O.bzip2_compress(data, 1)
    assert B().decompressFile(z) == data
    data = (b"x" * 70000 + b"y" * 255 + b"z" * 256 + T.ascii_random(29000, 7) + b"w" * 200000 + b"v" * 4 + b"u" * 5) * 2
    assert B().decompressFile(O.bzip2_compress(data, 1)) == data
    assert B().decompressFile(O.bzip2_compress(b"q" * 3000000, 2)) == b"q" * 3000000
    assert B().decompressFile(O.bzip2_compress(b"", 9)) == b""
    assert B().decompressFile(O.bzip2_compress(b"ab" * 50000, 1)) == b"ab" * 50000  # periodic block (many BWT cycles)


def test_multistream():
    a, b = T.texty(150000, 1), T.ascii_random(120000, 2)
    z = bz2.compress(a, 9) + bz2.compress(b, 9)
    assert B().decompressFile(z, None, True) == a + b
    assert B().decompressFile(z, None, False) == a
    assert B().decompressFile(z) == O.bzip2_decompress(z)


def test_multistream_members_with_different_levels():
    """`cat a-1.bz2 b-9.bz2 c-3.bz2`: the reference re-reads the level for every member (lib/Bzip2.js:105-124); a block of
    the level-9 member is far larger than the first member's dbufSize."""
    a, b, c3 = T.texty(250000, 11), T.ascii_random(1300000, 12), T.runs(200000, 13)
    z = bz2.compress(a, 1) + bz2.compress(b, 9) + bz2.compress(c3, 3)
    assert B().decompressFile(z, None, True) == a + b + c3
    assert B().decompressFile(z, None, True) == O.bzip2_decompress(z, multistream=True)
    # a block longer than its own member's limit is still a data error: level byte of a level-9 stream patched to 1
    big = bytearray(bz2.compress(T.ascii_random(400000, 14), 9))
    big[3] = ord("1")
    got, exp = _both(bytes(big))
    assert got == exp and got[0] == "err"


def _both(z, **kw):
    from compressjs_b200 import Bzip2Error
    try:
        exp = ("ok", O.bzip2_decompress(z, **kw))
    except O.OracleError as e:
        exp = ("err", e.errorCode)
    try:
        got = ("ok", B().decompressFile(z, None, kw.get("multistream", False)))
    except Bzip2Error as e:
        got = ("err", e.errorCode)
    return got, exp


def test_errors_match_reference():
    from compressjs_b200 import Bzip2Error
    with pytest.raises(Bzip2Error) as e:
        B().decompressFile(b"not a bzip2 file at all")
    assert e.value.errorCode == -2 and "bad magic" in str(e.value)
    with pytest.raises(Bzip2Error) as e:
        B().decompressFile(b"BZh0" + b"\x00" * 20)
    assert e.value.errorCode == -2 and "level out of range" in str(e.value)
    z = bytearray(T.fixture("sample1.bz2"))
    z[len(z) // 2] ^= 0x10
    got, exp = _both(bytes(z))
    assert got == exp and got[0] == "err"
    z = bytearray(T.fixture("sample0.bz2"))
    z[-5] ^= 1  # stream CRC
    with pytest.raises(Bzip2Error) as e:
        B().decompressFile(bytes(z))
    assert e.value.errorCode == -5 and "Bad stream CRC" in str(e.value)
    z = bytearray(T.fixture("sample0.bz2"))
    z[10] ^= 1  # block CRC field
    with pytest.raises(Bzip2Error) as e:
        B().decompressFile(bytes(z))
    assert e.value.errorCode == -5 and "Bad block CRC" in str(e.value)
    z = bytearray(T.fixture("sample0.bz2"))
    z[14] |= 0x80  # randomised bit
    with pytest.raises(Bzip2Error) as e:
        B().decompressFile(bytes(z))
    assert e.value.errorCode == -7


def test_corruption_fuzz_matches_oracle():
    z0 = T.fixture("sample1.bz2")
    g = T.rng(99)
    for _ in range(40):
        z = bytearray(z0)
        pos = int(g.integers(4, len(z)))
        z[pos] ^= 1 << int(g.integers(0, 8))
        got, exp = _both(bytes(z))
        assert got == exp, "corruption at byte %d" % pos
    for cut in (len(z0) - 1, len(z0) - 10, len(z0) // 2, 20):
        got, exp = _both(z0[:cut])
        assert got == exp, "truncation at %d" % cut
