"""compressjs.BWT on the GPU (lib/BWT.js:305-417)."""
import numpy as np

from . import _native


class BWT:
    @staticmethod
    def bwtransform2(T, U, n, alphabetSize=256):
        """Cyclic BWT (lib/BWT.js:372-417): fills U[0:n], returns pidx (row of rotation 0)."""
        if alphabetSize is not None and alphabetSize > 256:
            raise ValueError("only byte alphabets are supported on this path")
        L = _native.lib()
        src = np.ascontiguousarray(np.frombuffer(T, dtype=np.uint8)[:n] if not isinstance(T, np.ndarray) else T[:n], dtype=np.uint8)
        dst = np.zeros(max(n, 1), dtype=np.uint8)
        p = L.b2_bwt_cyclic(src.ctypes.data if n else None, dst.ctypes.data, n)
        if p < 0:
            raise RuntimeError("libb2bz: " + _native.last_error())
        U[:n] = dst[:n] if isinstance(U, np.ndarray) else bytes(dst[:n])
        return int(p)
