"""Reader for the binary snapshots written by oracle/ref_harness.cpp (fix sphbvf/snapshot).

Test infrastructure only.  Layout is documented at the top of oracle/ref_harness.cpp.
"""
import struct

import numpy as np


def read_snapshot(path):
    with open(path, "rb") as fh:
        buf = fh.read()
    off = 0

    def take(fmt):
        nonlocal off
        vals = struct.unpack_from("<" + fmt, buf, off)
        off += struct.calcsize("<" + fmt)
        return vals

    magic = buf[:8]
    off = 8
    if magic != b"SPHBVF01":
        raise ValueError("bad snapshot magic in %s" % path)
    (step,) = take("q")
    n, S, dim, ntypes, nfields, ago = take("6i")
    boxlo = take("3d")
    boxhi = take("3d")
    periodic = take("3i")
    (dt,) = take("d")
    mass = take("%dd" % ntypes)
    out = {
        "step": step, "natoms": n, "S": S, "dim": dim, "ntypes": ntypes, "ago": ago,
        "boxlo": list(boxlo), "boxhi": list(boxhi), "periodic": list(periodic), "dt": dt,
        "mass": list(mass), "fields": {},
    }
    for _ in range(nfields):
        name = buf[off:off + 24].split(b"\0", 1)[0].decode()
        off += 24
        ncols, is_int = take("2i")
        dtype = np.int32 if is_int else np.float64
        count = n * ncols
        arr = np.frombuffer(buf, dtype=dtype, count=count, offset=off).copy()
        off += count * (4 if is_int else 8)
        out["fields"][name] = arr.reshape(n, ncols) if ncols != 1 else arr
    (npairs,) = take("q")
    pairs = np.frombuffer(buf, dtype=np.int32, count=2 * npairs, offset=off).reshape(npairs, 2).copy()
    out["pairs"] = pairs
    return out


def canonical_pairs(pairs):
    """Unordered tag pairs -> sorted array of (min, max) rows (duplicates kept)."""
    p = np.sort(np.asarray(pairs, dtype=np.int64).reshape(-1, 2), axis=1)
    order = np.lexsort((p[:, 1], p[:, 0]))
    return p[order]
