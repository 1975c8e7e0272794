"""Small driver for ncu captures: one bzip2 -9 encode + one decode of MB MiB of a chosen workload."""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from compressjs_b200 import _native
from tests import util as T

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kind = sys.argv[2] if len(sys.argv) > 2 else "ascii"
mode = sys.argv[3] if len(sys.argv) > 3 else "both"
n = mb << 20
if kind == "ascii":
    data = T.ascii_random(n)
elif kind == "enwik":
    from tools.workloads import enwik_like
    n = mb * 1000000
    data = enwik_like(n).tobytes()
else:
    data = (T.texty(min(n, 8 << 20), 3) * (n // (8 << 20) + 1))[:n]
L = _native.lib()
L.b2_init(0)
d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
cap = L.b2_bzip2_bound(n)
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
out_n = C.c_size_t()
reps = 2 if mode != "dec" else 1
for _ in range(reps):
    rc = L.b2_bzip2_compress_dev(d_in.data_ptr(), n, 9, d_out.data_ptr(), cap, C.byref(out_n))
    assert rc == 0, _native.last_error()
print("enc", {k: round(v, 3) if isinstance(v, float) else v for k, v in _native.stats().items()})
if mode != "enc":
    d_dec = torch.empty(n, dtype=torch.uint8, device="cuda")
    dn = C.c_size_t()
    rc = L.b2_bzip2_decompress_dev(d_out.data_ptr(), out_n.value, 0, d_dec.data_ptr(), n, C.byref(dn))
    assert rc == 0 and dn.value == n and torch.equal(d_dec, d_in), _native.last_error()
    print("dec", {k: round(v, 3) if isinstance(v, float) else v for k, v in _native.stats().items()})
