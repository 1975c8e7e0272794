"""Host-buffer encode (b2_bzip2_compress, pinned input) of MB MiB of synthetic ASCII: wall time and stage times."""
import ctypes as C
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from compressjs_b200 import _native
from tests import util as T

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = mb << 20
L = _native.lib()
L.b2_init(0)
pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
pinned.numpy()[:] = np.frombuffer(T.ascii_random(n), dtype=np.uint8)
for r in range(reps + 1):
    out, on = C.POINTER(C.c_uint8)(), C.c_size_t()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = L.b2_bzip2_compress(pinned.data_ptr(), n, 9, C.byref(out), C.byref(on))
    dt = time.perf_counter() - t0
    assert rc == 0, _native.last_error()
    st = _native.stats()
    L.b2_free(out)
    if r:
        print("wall %.1f ms  %.0f MB/s " % (dt * 1e3, n / dt / 1e6), {k: round(v, 2) if isinstance(v, float) else v for k, v in st.items()
                                                                 if k.startswith("ms_") and v or k in ("blocks", "kernel_launches")})
