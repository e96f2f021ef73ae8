"""Write a synthetic seamount state as the flat file examples/pom_driver.cpp reads (the stand-in for
the netCDF inputs of the reference's `initialize`): magic, dims, blkcon scalars by name, arrays by
their COMMON-block names in Fortran (column-major) order.

    python scripts/dump_state.py OUT.bin IM JM KB [key=value ...]"""
import os
import struct
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from extpom_b200 import synthetic as syn  # noqa: E402


def dump(path, im, jm, kb, **kw):
    st = syn.make_state(im, jm, kb, **kw)
    with open(path, "wb") as f:
        f.write(b"POMSTAT1")
        f.write(struct.pack("<3i", im, jm, kb))
        consts = {k: float(v) for k, v in st["consts"].items()}
        f.write(struct.pack("<i", len(consts)))
        for k, v in consts.items():
            f.write(k.encode().ljust(32, b"\0")[:32])
            f.write(struct.pack("<d", v))
        fields = st["fields"]
        f.write(struct.pack("<i", len(fields)))
        for k, a in fields.items():
            a = np.asfortranarray(a, dtype=np.float64)
            f.write(k.encode().ljust(32, b"\0")[:32])
            f.write(struct.pack("<q", a.size))
            f.write(a.tobytes(order="F"))
    return st


def read_out(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            name = f.read(32)
            if len(name) < 32:
                break
            (cnt,) = struct.unpack("<q", f.read(8))
            out[name.rstrip(b"\0").decode()] = np.frombuffer(f.read(cnt * 8), dtype=np.float64)
    return out


if __name__ == "__main__":
    kw = {}
    for a in sys.argv[5:]:
        k, v = a.split("=")
        kw[k] = float(v) if "." in v else int(v)
    dump(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), **kw)
