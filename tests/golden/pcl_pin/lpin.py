"""LPIN: the trivial record file shared by make_pin_inputs.py, pcl_pin.cpp and tests/test_pcl_pin.py.
file   = magic "LPIN1\\0\\0\\0" (8 bytes) | int32 n_records | records
record = name (32 bytes, NUL padded) | int32 dtype (0 f32, 1 f64, 2 i32, 3 u8) | int32 ndim | int64 shape[4] | raw data"""
import struct

import numpy as np

DT = {0: np.float32, 1: np.float64, 2: np.int32, 3: np.uint8}
CODE = {np.dtype(v): k for k, v in DT.items()}


def write(path, records: dict):
    with open(path, "wb") as f:
        f.write(b"LPIN1\0\0\0")
        f.write(struct.pack("<i", len(records)))
        for name, a in records.items():
            a = np.ascontiguousarray(a)
            assert a.ndim <= 4 and a.dtype in CODE, (name, a.dtype)
            shape = list(a.shape) + [0] * (4 - a.ndim)
            f.write(name.encode().ljust(32, b"\0")[:32])
            f.write(struct.pack("<ii4q", CODE[a.dtype], a.ndim, *shape))
            f.write(a.tobytes())


def read(path) -> dict:
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == b"LPIN1\0\0\0", "not an LPIN file"
        (n,) = struct.unpack("<i", f.read(4))
        for _ in range(n):
            name = f.read(32).rstrip(b"\0").decode()
            code, ndim, *shape = struct.unpack("<ii4q", f.read(40))
            shape = shape[:ndim]
            cnt = int(np.prod(shape)) if ndim else 1
            a = np.frombuffer(f.read(cnt * np.dtype(DT[code]).itemsize), dtype=DT[code]).reshape(shape)
            out[name] = a.copy()
    return out
