"""Futhark binary data format (tools/png2data.py:50-57, tools/data2png.py:19-59 of the reference):
'b', version byte 2, rank byte, 4-character type, rank x u64 little-endian dims, little-endian values."""
import struct

import numpy as np

_TYPES = {b" i32": np.int32, b" u32": np.uint32, b" f32": np.float32, b"  u8": np.uint8, b"  i8": np.int8}


def dumps(a):
    a = np.ascontiguousarray(a)
    tag = {v: k for k, v in _TYPES.items()}[a.dtype.type]
    return b"b" + bytes([2, a.ndim]) + tag + b"".join(struct.pack("<Q", d) for d in a.shape) + a.astype(a.dtype.newbyteorder("<")).tobytes()


def loads(buf, offset=0):
    """-> (array, next_offset); skips white space between values like the Futhark reader."""
    while buf[offset:offset + 1] in (b" ", b"\n", b"\t", b"\r"):
        offset += 1
    if buf[offset:offset + 1] != b"b" or buf[offset + 1] != 2:
        raise ValueError("not a Futhark binary value")
    rank = buf[offset + 2]
    dt = _TYPES[bytes(buf[offset + 3:offset + 7])]
    shape = struct.unpack("<%dQ" % rank, buf[offset + 7:offset + 7 + 8 * rank])
    start = offset + 7 + 8 * rank
    n = int(np.prod(shape)) if rank else 1
    end = start + n * np.dtype(dt).itemsize
    return np.frombuffer(buf[start:end], dtype=np.dtype(dt).newbyteorder("<")).reshape(shape).astype(dt), end
