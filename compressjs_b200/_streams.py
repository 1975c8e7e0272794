"""Argument coercion of the reference's L1 stream layer (lib/Util.js:9-101, lib/Stream.js).

``input``  : object with ``readByte()`` (drained until it returns -1, Stream.EOF lib/Stream.js:4)
             or any bytes-like / sequence of ints.
``output`` : object with ``writeByte(b)`` -> every byte is pushed to it and IT is returned;
             an ``int`` -> exact-size result, ``TypeError('outputsize does not match decoded input')``
             if the result does not fill it exactly (lib/Util.js:69-71, 90-92);
             a writable buffer (bytearray / numpy uint8) -> filled in place under the same rule;
             ``None`` -> a fresh ``bytes`` object (the reference returns a trimmed Uint8Array).
"""
import numpy as np

EOF_BYTE = -1


def coerce_input(inp):
    if hasattr(inp, "readByte"):
        out = bytearray()
        while True:
            b = inp.readByte()
            if b == EOF_BYTE or b is None:
                break
            out.append(b & 0xFF)
        return np.frombuffer(bytes(out), dtype=np.uint8)
    if isinstance(inp, np.ndarray):
        return np.ascontiguousarray(inp, dtype=np.uint8).ravel()
    if isinstance(inp, (bytes, bytearray, memoryview)):
        return np.frombuffer(inp, dtype=np.uint8)
    return np.array(list(inp), dtype=np.uint8)


def deliver_output(output, data):
    """data: numpy uint8 view of the result.  Implements coerceOutputStream + retval."""
    if output is None:
        return data.tobytes()
    if hasattr(output, "writeByte"):
        for b in data.tobytes():
            output.writeByte(b)
        return output
    if isinstance(output, bool):
        raise TypeError("output must be a stream, a size or a buffer")
    if isinstance(output, int):
        if output != data.size:
            raise TypeError("outputsize does not match decoded input")
        return data.tobytes()
    mv = memoryview(output)
    if mv.nbytes != data.size:
        raise TypeError("outputsize does not match decoded input")
    mv.cast("B")[:] = data.tobytes()
    return output
