"""compressjs.BWTC on the GPU (lib/BWTC.js:10-231): same two entry points and the same container bytes.

New in round 1 (see the header of csrc/bwtc.cu for what has been verified where).  The block stages (sentinel BWT, MTF, zero runs) are
the block-parallel kernels of the bzip2 path; the range coder is one serial chain over the file, so this path is
bound by a single GPU thread."""
import ctypes as C

import numpy as np

from . import _native
from ._streams import coerce_input, deliver_output


def _take(L, p, n):
    arr = np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, dtype=np.uint8)
    L.b2_free(p)
    return arr


class BWTC:
    MAGIC = "bwtc"  # lib/BWTC.js:11

    @staticmethod
    def compressFile(input, output=None, props=None):
        """lib/BWTC.js:12-139.  props: block size in units of 100 000 bytes, 1..9; anything else means 9 (:16-19)."""
        L = _native.lib()
        data = coerce_input(input)
        level = 9
        if isinstance(props, (int, float)) and not isinstance(props, bool) and 1 <= props <= 9:
            level = int(props)
        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        rc = L.b2_bwtc_compress(data.ctypes.data if data.size else None, data.size, level, C.byref(out), C.byref(n))
        if rc:
            raise RuntimeError("libb2bz: %s (code %d)" % (_native.last_error(), rc))
        return deliver_output(output, _take(L, out, n))

    @staticmethod
    def decompressFile(input, output=None):
        """lib/BWTC.js:141-231.  Raises ValueError("Bad magic") like lib/Util.js:151-153 throws Error("Bad magic")."""
        L = _native.lib()
        data = coerce_input(input)
        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        rc = L.b2_bwtc_decompress(data.ctypes.data if data.size else None, data.size, C.byref(out), C.byref(n))
        if rc == -102:
            raise ValueError("Bad magic")
        if rc:
            raise RuntimeError("libb2bz: %s (code %d)" % (_native.last_error(), rc))
        return deliver_output(output, _take(L, out, n))
