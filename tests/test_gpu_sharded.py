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
