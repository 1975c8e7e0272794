"""Quick on-GPU check of the experimental BWTC path against the oracle (no torch, no pytest: starts in seconds)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.time()
import numpy as np
from oracle import oracle as O
from compressjs_b200 import BWTC

out = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "bwtc_try.txt"), "w")


def say(*a):
    print("%.1fs" % (time.time() - t0), *a, file=out, flush=True)
    print("%.1fs" % (time.time() - t0), *a, flush=True)


g = np.random.Generator(np.random.PCG64(5))
cases = [("tiny", b"This is a test\n"), ("empty", b""), ("one", b"a"), ("zeros", b"\x00" * 5000),
         ("ascii250k", bytes(g.integers(32, 127, size=250001, dtype=np.uint8))),
         ("words", b" ".join(bytes(g.integers(97, 123, size=int(l), dtype=np.uint8)) for l in g.integers(2, 9, size=20000)))]
os.makedirs(os.path.dirname(out.name), exist_ok=True)
for name, d in cases:
    for level in (1, 9):
        try:
            exp = O.bwtc_compress(d, level)
            z = BWTC.compressFile(d, None, level)
            ok_enc = bytes(z) == exp
            back = BWTC.decompressFile(exp)
            say(name, level, "enc", ok_enc, "dec", bytes(back) == d, len(exp))
        except Exception as e:  # noqa: BLE001
            say(name, level, "EXC", repr(e)[:200])
say("done")
