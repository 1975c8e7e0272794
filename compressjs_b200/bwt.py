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

    @staticmethod
    def suffixsort(T, SA, n, alphabetSize=256):
        """Suffix array of T[0:n] (lib/BWT.js:305-321): fills SA[0:n] (int32), returns 0."""
        if alphabetSize is not None and alphabetSize > 256:
            raise ValueError("only byte alphabets are supported on this path")
        L = _native.lib()
        src = _bytes_view(T, n)
        sa = np.zeros(max(n, 1), dtype=np.int32)
        rc = L.b2_suffixsort(src.ctypes.data if n else None, sa.ctypes.data, n)
        if rc < 0:
            raise RuntimeError("libb2bz: " + _native.last_error())
        SA[:n] = sa[:n]
        return 0

    @staticmethod
    def bwtransform(T, U, A, n, alphabetSize=256):
        """Sentinel BWT (lib/BWT.js:328-350): fills U[0:n], returns pidx + 1.  A (the reference's int32 scratch
        array) is accepted and left untouched."""
        if alphabetSize is not None and alphabetSize > 256:
            raise ValueError("only byte alphabets are supported on this path")
        L = _native.lib()
        src = _bytes_view(T, n)
        dst = np.zeros(max(n, 1), dtype=np.uint8)
        p = L.b2_bwt_sentinel(src.ctypes.data if n else None, dst.ctypes.data, n)
        if p < 0:
            raise RuntimeError("libb2bz: " + _native.last_error())
        U[:n] = dst[:n] if isinstance(U, np.ndarray) else bytes(dst[:n])
        return int(p)

    @staticmethod
    def unbwtransform(T, U, LF, n, pidx):
        """Inverse of bwtransform (lib/BWT.js:352-363): T = transformed string, fills U[0:n].  LF (the reference's
        scratch array) is accepted and left untouched."""
        L = _native.lib()
        src = _bytes_view(T, n)
        dst = np.zeros(max(n, 1), dtype=np.uint8)
        rc = L.b2_bwt_inverse(src.ctypes.data if n else None, dst.ctypes.data, n, pidx)
        if rc < 0:
            raise RuntimeError("libb2bz: " + _native.last_error())
        U[:n] = dst[:n] if isinstance(U, np.ndarray) else bytes(dst[:n])


def _bytes_view(T, n):
    a = T[:n] if isinstance(T, np.ndarray) else np.frombuffer(T, dtype=np.uint8)[:n]
    return np.ascontiguousarray(a, dtype=np.uint8)
