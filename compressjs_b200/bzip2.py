"""compressjs.Bzip2 on the GPU: same four entry points as lib/Bzip2.js:879-933."""
import ctypes as C

import numpy as np

from . import _native
from ._streams import coerce_input, deliver_output


class Bzip2Error(TypeError):
    """Decode errors are TypeErrors carrying ``errorCode`` (lib/Bzip2.js:82-88)."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.errorCode = code


class Err:  # lib/Bzip2.js:62-72
    OK = 0
    LAST_BLOCK = -1
    NOT_BZIP_DATA = -2
    UNEXPECTED_INPUT_EOF = -3
    UNEXPECTED_OUTPUT_EOF = -4
    DATA_ERROR = -5
    OUT_OF_MEMORY = -6
    OBSOLETE_INPUT = -7
    END_OF_BLOCK = -8


def _raise(rc):
    msg = _native.last_error()
    if rc == -100:
        raise ValueError(msg or "Invalid block size multiplier")  # `new Error(...)` lib/Bzip2.js:888-890
    if rc in (Err.NOT_BZIP_DATA, Err.DATA_ERROR, Err.OBSOLETE_INPUT):
        raise Bzip2Error(rc, msg)
    raise RuntimeError("libb2bz: %s (code %d)" % (msg, rc))


def _take(L, p, n):
    arr = np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, dtype=np.uint8)
    L.b2_free(p)
    return arr


class Bzip2:
    Err = Err

    @staticmethod
    def compressFile(input, output=None, props=None):
        """lib/Bzip2.js:879-929.  props: block size multiplier 1..9 (default 9)."""
        L = _native.lib()
        data = coerce_input(input)
        level = 9
        if isinstance(props, (int, float)) and not isinstance(props, bool):
            level = props
        if level < 1 or level > 9 or int(level) != level:
            raise ValueError("Invalid block size multiplier")
        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        rc = L.b2_bzip2_compress(data.ctypes.data if data.size else None, data.size, int(level), C.byref(out), C.byref(n))
        if rc:
            _raise(rc)
        return deliver_output(output, _take(L, out, n))

    @staticmethod
    def decompressFile(input, output=None, multistream=False):
        """lib/Bzip2.js:454-481 (Bunzip.decode)."""
        L = _native.lib()
        data = coerce_input(input)
        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        rc = L.b2_bzip2_decompress(data.ctypes.data if data.size else None, data.size, int(bool(multistream)), C.byref(out), C.byref(n))
        if rc:
            _raise(rc)
        return deliver_output(output, _take(L, out, n))

    @staticmethod
    def decompressBlock(input, pos, output=None):
        """lib/Bzip2.js:482-503 (Bunzip.decodeBlock): decode the single block whose magic starts at bit `pos`."""
        L = _native.lib()
        data = coerce_input(input)
        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        rc = L.b2_bzip2_decompress_block(data.ctypes.data if data.size else None, data.size, int(pos), C.byref(out), C.byref(n))
        if rc:
            _raise(rc)
        return deliver_output(output, _take(L, out, n))

    @staticmethod
    def table(input, callback, multistream=False):
        """lib/Bzip2.js:508-548: callback(bit position, decoded bytes) once per block."""
        L = _native.lib()
        data = coerce_input(input)
        bp, sz, cnt = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)(), C.c_size_t()
        rc = L.b2_bzip2_table(data.ctypes.data if data.size else None, data.size, int(bool(multistream)), C.byref(bp), C.byref(sz), C.byref(cnt))
        if rc:
            _raise(rc)
        rows = [(int(bp[i]), int(sz[i])) for i in range(cnt.value)]
        L.b2_free(bp)
        L.b2_free(sz)
        for pos, size in rows:
            callback(pos, size)
