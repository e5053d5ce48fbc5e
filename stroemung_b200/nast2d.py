"""Reader for NaSt2D ``.out`` binaries (host-side I/O, no compute).

Layout as documented by the reference's converter
(/root/reference/python/generate_test_data.py:107-162): two C ints ``imax``,
``jmax``; then U, V, P, T as ``(imax+2)*(jmax+2)`` C doubles each, x-major
(``U[x][y]``); then the flag field as C ints (``0x10`` = fluid, else boundary).
`as_grid` applies the converter's boundary-kind reconstruction (:46-78): left
wall Inflow carrying the file's (u, v), right wall Outflow, everything else
NoSlip.
"""
import struct

import numpy as np

from . import refjson

FLAG_FLUID = 16


def parse_out(data: bytes, int_bytes=4, int_fmt="i"):
    imax, jmax = struct.unpack_from("<" + int_fmt * 2, data, 0)
    off = 2 * int_bytes
    n = (imax + 2) * (jmax + 2)
    shape = (imax + 2, jmax + 2)
    fields = {}
    for name in ("U", "V", "P", "T"):
        fields[name] = np.frombuffer(data, dtype="<f8", count=n, offset=off).reshape(shape).copy()
        off += 8 * n
    flags = np.array(struct.unpack_from("<" + int_fmt * n, data, off), dtype=np.int64)
    fields["flags"] = flags.reshape(shape)
    fields["imax"], fields["jmax"] = imax, jmax
    return fields


def as_grid(out):
    """NaSt2D output -> dict(size, p, u, v, kind, bu, bv) like refjson.grid_from_json."""
    imax, jmax = out["imax"], out["jmax"]
    nx, ny = imax + 2, jmax + 2
    kind = np.full((nx, ny), refjson.KIND_NOSLIP, dtype=np.uint8)
    bu = np.zeros((nx, ny))
    bv = np.zeros((nx, ny))
    # CONFORM IntFlag semantics of the converter: only bit 0x10 is meaningful
    fluid = (out["flags"] & FLAG_FLUID) != 0
    kind[fluid] = refjson.KIND_FLUID
    for y in range(1, jmax + 1):
        if not fluid[0, y]:
            kind[0, y] = refjson.KIND_INFLOW
            bu[0, y], bv[0, y] = out["U"][0, y], out["V"][0, y]
        if not fluid[imax + 1, y]:
            kind[imax + 1, y] = refjson.KIND_OUTFLOW
    return {"size": (nx, ny), "p": out["P"], "u": out["U"], "v": out["V"],
            "kind": kind, "bu": bu, "bv": bv}
