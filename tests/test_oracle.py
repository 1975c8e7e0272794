"""CPU tests: the oracle (oracle/bz2_oracle.c, a restatement of the reference's JavaScript) against
every golden vector the reference's own tests hold for this path."""
import bz2
import hashlib

import pytest

from oracle import oracle as O
from tests import util as T

FIB = [0, 1, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2584, 4181, 6765, 10946, 17711, 28657, 46368, 75025,
       121393, 196418, 317811, 514229, 832040, 1346269, 2178309, 3524578, 5702887, 9227465, 14930352]

MARY = ("Mary had a little lamb, its fleece was white as snow" * 8 + "Nary had a little lamb, its fleece was white as snow")
MARY_OUT = ("dddddddddeeeeeeeeesssssssssyyyyyyyyy,,,,,,,,,eeeeeeeeeaaaaaaaaassssssssseeeeeeeeesss"
            "ssssssbbbbbbbbbwwwwwwwww         hhhhhhhhhlllllllllNMMMMMMMM         wwwwwwwwwmmmmmm"
            "mmmeeeeeeeeeaaaaaaaaatttttttttlllllllllccccccccceeeeeeeeelllllllll                  "
            "wwwwwwwwwhhhhhhhhh         lllllllll         tttttttttfffffffff         aaaaaaaaasss"
            "ssssssnnnnnnnnnaaaaaaaaatttttttttaaaaaaaaaaaaaaaaaa         iiiiiiiiitttttttttiiiiii"
            "iiiiiiiiiiiiooooooooo                  rrrrrrrrr")


def test_crc32_check_value():
    # CRC("This is a test\n") is bytes 10..13 of test/sample0.bz2
    assert O.crc32(b"This is a test\n") == 0xEA29357D
    assert O.crc32(b"") == 0
    assert O.crc32(b"123456789") == 0xFC891918  # CRC-32/BZIP2 catalogue check value


def test_bwt_kats():  # test/bwtest.js:39-79
    kats = [("bcababa", "cbbaaab", 5), ("ABCDEFGHIJKLMNOPQRSTUVWXYZ", "ZABCDEFGHIJKLMNOPQRSTUVWXY", 0),
            ("ZYXWVUTSRQPONMLKJIHGFEDCBA", "BCDEFGHIJKLMNOPQRSTUVWXYZA", 25),
            ("SIX.MIXED.PIXIES.SIFT.SIXTY.PIXIE.DUST.BOXES", "TEXYDST.E.IXIXIXXSSMPPS.B..E.S.EUSFXDIIOIIIT", 29),
            (MARY, MARY_OUT, 99)]
    for i, o, p in kats:
        assert O.bwt_cyclic(i.encode()) == (o.encode(), p)


def test_bwt_periodic_tie_rule():  # equal rotations in DESCENDING start index
    assert O.bwt_cyclic(b"abababab")[1] == 3
    assert O.bwt_cyclic(b"aaaa")[1] == 3
    assert O.bwt_cyclic(b"abcabcabc")[1] == 2


def test_sentinel_bwt_roundtrip():  # test/bwtest.js:10-36
    for name in ("sample0", "sample1", "sample3"):
        data = T.fixture(name + ".ref")
        u, p = O.bwt_sentinel(data)
        assert O.unbwt_sentinel(u, p) == data


def test_suffixsort_property():  # test/suftest.js:10-83 (order check against a direct comparison)
    data = T.fixture("sample1.ref")[:20000]
    sa = O.suffixsort(data)
    assert sorted(sa.tolist()) == list(range(len(data)))
    for a, b in zip(sa[:-1], sa[1:]):
        assert data[a:] < data[b:]


def test_huffman_allocator_kats():  # test/huffman.js:15-76
    f = O.huffman_code_lengths
    assert f([1], 32) == [1]
    assert f([1, 1], 32) == [1, 1]
    assert f([1] * 5, 32) == [3, 3, 2, 2, 2]
    assert f([0, 0, 1, 1, 1, 1], 3) == [3, 3, 3, 3, 2, 2]
    assert f(FIB[:36], 20) == [20] * 16 + [19, 19, 18, 17, 16, 16, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1]
    assert f(FIB[:22], 20) == [20, 20, 19, 19, 19, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1]
    assert f(FIB[:21], 20) == [20, 20, 19, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1]
    assert f(FIB[:36], 6) == [6] * 30 + [5, 5, 5, 4, 3, 2]


@pytest.mark.parametrize("k", range(5))
def test_decode_fixtures(k):  # test/bzip2-basic.js, test/bzip2-table.js
    ref = T.fixture("sample%d.ref" % k)
    z = T.fixture("sample%d.bz2" % k)
    assert O.bzip2_decompress(z) == ref
    exp = [tuple(map(int, l.split("\t"))) for l in T.fixture("sample%d.bzt" % k).decode().strip().split("\n")]
    assert O.bzip2_table(z) == exp


def test_decode_block_fixtures():  # test/bzip2-block.js:5-27
    for f, pos, blk in [("sample0", 32, "sample0.ref"), ("sample2", 544888, "sample2.544888"), ("sample4", 32, "sample4.32"),
                        ("sample4", 1596228, "sample4.1596228"), ("sample4", 2342106, "sample4.2342106")]:
        assert O.bzip2_decompress_block(T.fixture(f + ".bz2"), pos) == T.fixture(blk)


def test_decode_errors():
    with pytest.raises(O.OracleError) as e:
        O.bzip2_decompress(b"not a bzip2 file")
    assert e.value.errorCode == -2
    z = bytearray(T.fixture("sample0.bz2"))
    z[20] ^= 0x55
    with pytest.raises(O.OracleError) as e:
        O.bzip2_decompress(bytes(z))
    assert e.value.errorCode == -5


def test_encode_config1_worked_example():
    z = O.bzip2_compress(b"This is a test\n", 1)
    assert z.hex() == ("425a6831314159265359ea29357d000002538000104000040022600c00200021aa8f6f4a9ef5"
                       "0806059702fb58a70bb9229c284875149abe80")
    assert bz2.decompress(z) == b"This is a test\n"


@pytest.mark.parametrize("k", range(6))
@pytest.mark.parametrize("level", [1, 9])
def test_encode_samples_roundtrip_and_golden(k, level):  # test/file.js:5-46 + committed golden hashes
    data = T.fixture("sample%d.ref" % k)
    z = O.bzip2_compress(data, level)
    assert bz2.decompress(z) == data          # libbz2 accepts the stream
    assert O.bzip2_decompress(z) == data      # and so does the restated reference decoder
    g = T.golden()["bzip2_sample%d_-%d" % (k, level)]
    assert (len(z), hashlib.sha256(z).hexdigest()) == (g["size"], g["sha256"])


def test_readme_sizes_in_legacy_sort_mode():
    # README.md:42,45 of the reference: bzip2 -9 / -1 on sample5.ref -> 275087 / 341615 bytes.
    # They are reproduced exactly when Array.prototype.sort behaves like pre-7.0 V8.
    data = T.fixture("sample5.ref")
    assert len(O.bzip2_compress(data, 9, legacy_sort=True)) == 275087
    assert len(O.bzip2_compress(data, 1, legacy_sort=True)) == 341615


def test_mt_arm_is_byte_identical():
    data = T.texty(350000, 3)
    assert O.bzip2_compress(data, 1, threads=4) == O.bzip2_compress(data, 1)


def test_rle1_quirk_uncounted_run_of_four():
    # block fills on the 4th byte of a run -> no count byte; libbz2 rejects, the reference decoder accepts
    data = T.ascii_random(99977, 3).replace(b"aaaa", b"abab") + b"a" * 20 + b"tail"
    z = O.bzip2_compress(data, 1)
    assert O.bzip2_decompress(z) == data


# ---- BWTC container (lib/BWTC.js) -- oracle only, no CUDA path yet ----------------
BWTC9 = {  # size and SHA-256 of BWTC.compressFile(sample, null, 9)
    "sample1": (32903, "982488a6bb9282e93e848ff7e9eed798a580f2b3f3e963677573c9ad9fda9fcb"),
    "sample2": (74536, "b6109e3e2e40d0b143a2a72cf8f2302b67618a373f198a8530ba28b3c5d35088"),
    "sample3": (201, "0e06e8045a87124f2b0692c02b4a0ce0544b47655a2a54a09db6c04643ba970d"),
    "sample4": (335422, "9f5c636794242b9a24040e8c75fdc58044c861ded15a2f711e21812c5492738c"),
    "sample5": (272997, "01d8d0a0490c39ef86808be3a449a1f57c864c11ea1cf5b2057ba56a241c23ca"),  # = README.md:41
}


def test_bwtc_sample0_whole_file():
    z = O.bwtc_compress(T.fixture("sample0.ref"), 9)
    assert z.hex() == "627774639009625e4eff9d2a362e8184ba3a3eef321c245ae8adbbb300001c"
    assert O.bwtc_decompress(z) == T.fixture("sample0.ref")


@pytest.mark.parametrize("name", sorted(BWTC9))
def test_bwtc_level9_vectors(name):
    d = T.fixture(name + ".ref")
    z = O.bwtc_compress(d, 9)
    assert (len(z), hashlib.sha256(z).hexdigest()) == BWTC9[name]
    assert O.bwtc_decompress(z) == d


def test_bwtc_other_levels_and_edges():
    d = T.fixture("sample5.ref")
    z = O.bwtc_compress(d, 1)   # levels 1..5 switch to the deferred-sum model (lib/BWTC.js:22,111)
    assert len(z) == 345764 and O.bwtc_decompress(z) == d   # README.md:46
    z = O.bwtc_compress(d, 6)
    assert (len(z), hashlib.sha256(z).hexdigest()) == (279678, "29206d8d51e293ddee9e5fafb27552502bba5adc9730e27cefb56e962bb7c1b6")
    assert O.bwtc_decompress(z) == d
    for data in (b"", b"a", b"\x00" * 5000, bytes(range(256)) * 9, T.ascii_random(600000, 3) + T.runs(100001, 4)):
        for level in (1, 5, 6, 9):
            assert O.bwtc_decompress(O.bwtc_compress(data, level)) == data   # 700001 bytes: full + short blocks at -1, -5, -6
    with pytest.raises(O.OracleError):
        O.bwtc_decompress(b"bzzt" + b"\x81\x00\x00\x00\x00\x00")


@pytest.mark.parametrize("k", range(6))
def test_bwtc_committed_golden(k):
    d = T.fixture("sample%d.ref" % k)
    for level in (1, 6, 9):
        z = O.bwtc_compress(d, level)
        g = T.golden()["bwtc_sample%d_-%d" % (k, level)]
        assert (len(z), hashlib.sha256(z).hexdigest()) == (g["size"], g["sha256"])
