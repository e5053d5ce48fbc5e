"""The reference's on-disk JSON format (host-side I/O, no compute).

Mirrors what serde derives for `UnfinalizedSimulation` / `UnfinalizedSimulationGrid`
(/root/reference/src/simulation.rs:30-44, src/grid/mod.rs:85-92):

* arrays are ndarray-serde documents ``{"v": 1, "dim": [nx, ny], "data": [...]}``
  with ``data`` row-major, i.e. index ``(x, y)`` and y contiguous;
* a cell is ``"Fluid"``, ``{"Boundary": "NoSlip"}``, ``{"Boundary": "Outflow"}`` or
  ``{"Boundary": {"Inflow": {"velocity": [u, v]}}}`` (src/cell.rs:6-23).

Device-side the cell mask is a u8 kind (0 Fluid, 1 NoSlip, 2 Outflow, 3 Inflow,
4 MovingWall [extension]) plus a sparse table of boundary velocities.

Note: the reference parses with serde_json 1.0.140 (Cargo.lock:526-527)
without its `float_roundtrip` feature (Cargo.toml:18), whose number parser is fast but not
correctly rounded: e.g. the fixture literal -0.14603099243353101 is read 1 ulp low.  Python's
parser is correctly rounded.  `quirk_serde_json=True` parses numbers with `serde_json_f64`, a
restatement of that crate's published algorithm (the crate itself is a dependency that is not
vendored in /root/reference), so that a document loads to exactly the doubles the reference
computes with; it is pinned by the value the reference's own snapshot holds for that literal
(src/snapshots/stroemung__simulation__tests__deserialize-2.snap:42) and by
`initial_norm_squared` 899.9547140394143 of the same test.
"""
import json

import numpy as np

KIND_FLUID, KIND_NOSLIP, KIND_OUTFLOW, KIND_INFLOW, KIND_MOVING_WALL = range(5)

# ---- serde_json 1.0.140, default features: how a JSON number becomes an f64 ----------------
# src/de.rs of the crate: parse_integer / parse_decimal accumulate the digits into a u64
# significand and stop taking digits once the next one would overflow it (digits of the integer
# part that are dropped still count into the exponent, digits of the fraction are ignored);
# parse_exponent adds the written exponent (saturating i32); f64_from_parts then computes
# `significand as f64`, scaled by ONE multiplication or division by an exactly represented
# power of ten (two roundings in all, where a correctly rounded parser has one), with a loop of
# divisions by 1e308 for exponents below the table.
_U64_MAX = (1 << 64) - 1
_I32_MAX = (1 << 31) - 1
_POW10 = [float(f"1e{i}") for i in range(309)]


def _u64_overflows(significand, digit):
    """the crate's overflow!(significand * 10 + digit, u64::MAX)"""
    return significand >= _U64_MAX // 10 and (significand > _U64_MAX // 10
                                              or digit > _U64_MAX % 10)


def _f64_from_parts(positive, significand, exponent):
    f = float(significand)          # `significand as f64`: round to nearest even
    while True:
        k = abs(exponent)
        if k < len(_POW10):
            if exponent >= 0:
                f *= _POW10[k]
                if f == float("inf"):
                    raise ValueError("number out of range")
            else:
                f /= _POW10[k]
            break
        if f == 0.0:
            break
        if exponent >= 0:
            raise ValueError("number out of range")
        f /= 1e308
        exponent += 308
    return f if positive else -f


def serde_json_f64(text):
    """The f64 that serde_json 1.0.140 (no `float_roundtrip`, no `arbitrary_precision`)
    deserialises the JSON number `text` to."""
    s = text.strip()
    i, n = 0, len(s)
    positive = True
    if i < n and s[i] == "-":
        positive, i = False, i + 1
    if i >= n or not s[i].isdigit():
        raise ValueError(f"invalid number {text!r}")
    significand, exponent = 0, 0
    if s[i] == "0":                      # only one leading zero
        i += 1
        if i < n and s[i].isdigit():
            raise ValueError(f"invalid number {text!r}")
    else:
        while i < n and s[i].isdigit():
            d = ord(s[i]) - 48
            if _u64_overflows(significand, d):
                # parse_long_integer: the remaining integer digits only scale the value
                while i < n and s[i].isdigit():
                    exponent += 1
                    i += 1
                break
            significand = significand * 10 + d
            i += 1
    if i < n and s[i] == ".":            # parse_decimal
        i += 1
        start = i
        while i < n and s[i].isdigit():
            d = ord(s[i]) - 48
            if _u64_overflows(significand, d):
                while i < n and s[i].isdigit():   # parse_decimal_overflow: digits ignored
                    i += 1
                break
            significand = significand * 10 + d
            exponent -= 1
            i += 1
        if i == start:
            raise ValueError(f"invalid number {text!r}")
    if i < n and s[i] in "eE":           # parse_exponent
        i += 1
        positive_exp = True
        if i < n and s[i] in "+-":
            positive_exp = s[i] == "+"
            i += 1
        if i >= n or not s[i].isdigit():
            raise ValueError(f"invalid number {text!r}")
        exp = 0
        while i < n and s[i].isdigit():
            d = ord(s[i]) - 48
            if exp >= _I32_MAX // 10 and (exp > _I32_MAX // 10 or d > _I32_MAX % 10):
                # parse_exponent_overflow: an error instead of +/- infinity, else +/- 0
                if significand != 0 and positive_exp:
                    raise ValueError("number out of range")
                return 0.0 if positive else -0.0
            exp = exp * 10 + d
            i += 1
        exponent = (min(exponent + exp, _I32_MAX) if positive_exp
                    else max(exponent - exp, -_I32_MAX - 1))
    if i != n:
        raise ValueError(f"invalid number {text!r}")
    return _f64_from_parts(positive, significand, exponent)


def array_from_json(doc, dtype=np.float64):
    assert doc["v"] == 1, "unknown ndarray-serde version"
    nx, ny = doc["dim"]
    return np.asarray(doc["data"], dtype=dtype).reshape(nx, ny)


def array_to_json(a):
    a = np.asarray(a)
    return {"v": 1, "dim": [int(a.shape[0]), int(a.shape[1])],
            "data": [float(x) for x in a.reshape(-1)]}


def cells_from_json(doc):
    """cell_type document -> (kind u8 [nx,ny], bu f64 [nx,ny], bv f64 [nx,ny])."""
    nx, ny = doc["dim"]
    kind = np.zeros(nx * ny, dtype=np.uint8)
    bu = np.zeros(nx * ny)
    bv = np.zeros(nx * ny)
    for i, c in enumerate(doc["data"]):
        if c == "Fluid":
            kind[i] = KIND_FLUID
            continue
        b = c["Boundary"]
        if b == "NoSlip":
            kind[i] = KIND_NOSLIP
        elif b == "Outflow":
            kind[i] = KIND_OUTFLOW
        elif isinstance(b, dict) and "Inflow" in b:
            kind[i] = KIND_INFLOW
            bu[i], bv[i] = b["Inflow"]["velocity"]
        elif isinstance(b, dict) and "MovingWall" in b:  # extension
            kind[i] = KIND_MOVING_WALL
            bu[i], bv[i] = b["MovingWall"]["velocity"]
        else:
            raise ValueError(f"unknown cell {c!r}")
    return kind.reshape(nx, ny), bu.reshape(nx, ny), bv.reshape(nx, ny)


def cells_to_json(kind, bu, bv):
    kind = np.asarray(kind)
    data = []
    for k, a, b in zip(kind.reshape(-1), np.asarray(bu).reshape(-1), np.asarray(bv).reshape(-1)):
        if k == KIND_FLUID:
            data.append("Fluid")
        elif k == KIND_NOSLIP:
            data.append({"Boundary": "NoSlip"})
        elif k == KIND_OUTFLOW:
            data.append({"Boundary": "Outflow"})
        elif k == KIND_INFLOW:
            data.append({"Boundary": {"Inflow": {"velocity": [float(a), float(b)]}}})
        elif k == KIND_MOVING_WALL:
            data.append({"Boundary": {"MovingWall": {"velocity": [float(a), float(b)]}}})
        else:
            raise ValueError(k)
    return {"v": 1, "dim": [int(kind.shape[0]), int(kind.shape[1])], "data": data}


def grid_from_json(doc):
    """UnfinalizedSimulationGrid document -> dict(size, p, u, v, kind, bu, bv)."""
    kind, bu, bv = cells_from_json(doc["cell_type"])
    g = {"size": tuple(doc["size"]), "p": array_from_json(doc["pressure"]),
         "u": array_from_json(doc["u"]), "v": array_from_json(doc["v"]),
         "kind": kind, "bu": bu, "bv": bv}
    for k in ("p", "u", "v", "kind"):
        assert g[k].shape == g["size"], (k, g[k].shape, g["size"])
    return g


def simulation_from_json(doc):
    """UnfinalizedSimulation document -> (params dict, grid dict)."""
    prm = {
        "size": tuple(doc["size"]),
        "cell_size": tuple(doc["cell_size"]),
        "delt": doc["delt"], "gamma": doc["gamma"], "reynolds": doc["reynolds"],
        "initial_norm_squared": doc.get("initial_norm_squared"),
        "sor_absolute_epsilon": doc["sor_absolute_epsilon"],
        "max_iterations": doc["max_iterations"], "iterations": doc["iterations"],
        "time": doc["time"], "omega": doc["omega"],
    }
    return prm, grid_from_json(doc["grid"])


def loads(text, quirk_serde_json=False):
    """json.loads; optionally reading every float literal the way the reference's parser
    does (`serde_json_f64`).  Integer literals stay Python ints: converted to f64 they are
    rounded to nearest even by Python exactly as by Rust's `u64 as f64`."""
    if not quirk_serde_json:
        return json.loads(text)
    return json.loads(text, parse_float=serde_json_f64)


def simulation_to_json(prm, grid):
    """Serialize like `#[derive(Serialize)] Simulation` (f, g, rhs, boundaries skipped)."""
    return {
        "size": list(prm["size"]), "cell_size": list(prm["cell_size"]),
        "delt": prm["delt"], "gamma": prm["gamma"], "reynolds": prm["reynolds"],
        "initial_norm_squared": prm.get("initial_norm_squared"),
        "sor_absolute_epsilon": prm["sor_absolute_epsilon"],
        "max_iterations": prm["max_iterations"], "iterations": prm["iterations"],
        "time": prm["time"], "omega": prm["omega"],
        "grid": {
            "size": list(prm["size"]),
            "pressure": array_to_json(grid["p"]), "u": array_to_json(grid["u"]),
            "v": array_to_json(grid["v"]),
            "cell_type": cells_to_json(grid["kind"], grid["bu"], grid["bv"]),
        },
    }


# ---- binary sidecar (extension) -----------------------------------------------------------
# serde_json text costs ~20 bytes per f64 and minutes of parsing at 8192^2 (3 x 67 M values),
# so for large grids the arrays move into ONE little-endian binary file next to the JSON
# document and the document keeps everything else of the reference's format.  An array
# document then reads {"v": 1, "dim": [nx, ny], "sidecar": {"file": name, "offset": bytes,
# "dtype": "<f8"}} instead of carrying "data"; the cell types are u8 kinds (0 Fluid, 1 NoSlip,
# 2 Outflow, 3 Inflow, 4 MovingWall) plus the sparse list "velocities": [[x, y, u, v], ..] of
# the cells that carry one (src/cell.rs:12-16).  Documents without "sidecar" entries are the
# reference's own format and load in the reference unchanged.
SIDECAR_MIN_CELLS = 1 << 18     # grids from 512 x 512 on use the sidecar by default


def _resolve_array(doc, base_dir):
    """array document (inline or sidecar) -> ndarray"""
    import os
    if "sidecar" not in doc:
        return array_from_json(doc)
    sc = doc["sidecar"]
    nx, ny = doc["dim"]
    return np.fromfile(os.path.join(base_dir, sc["file"]), dtype=np.dtype(sc["dtype"]),
                       count=nx * ny, offset=sc["offset"]).reshape(nx, ny)


def save_simulation(path, prm, grid, sidecar=None):
    """Write the simulation as the reference's JSON document (`path`).  sidecar: True / False /
    None (= by size): arrays go to `path + ".bin"`.  Returns the document written."""
    import os
    nx, ny = prm["size"]
    if sidecar is None:
        sidecar = nx * ny >= SIDECAR_MIN_CELLS
    if not sidecar:
        doc = simulation_to_json(prm, grid)
        with open(path, "w") as f:
            json.dump(doc, f)
        return doc
    bin_name = os.path.basename(path) + ".bin"
    doc = simulation_to_json(prm, {"p": np.zeros((0, 0)), "u": np.zeros((0, 0)),
                                   "v": np.zeros((0, 0)), "kind": np.zeros((0, 0), np.uint8),
                                   "bu": np.zeros((0, 0)), "bv": np.zeros((0, 0))})
    offset = 0
    with open(path + ".bin", "wb") as f:
        for key, name in (("pressure", "p"), ("u", "u"), ("v", "v")):
            a = np.ascontiguousarray(grid[name], dtype="<f8")
            assert a.shape == (nx, ny), (name, a.shape)
            doc["grid"][key] = {"v": 1, "dim": [nx, ny],
                                "sidecar": {"file": bin_name, "offset": offset, "dtype": "<f8"}}
            f.write(a.tobytes())
            offset += a.nbytes
        kind = np.ascontiguousarray(grid["kind"], dtype=np.uint8)
        doc["grid"]["cell_type"] = {
            "v": 1, "dim": [nx, ny],
            "sidecar": {"file": bin_name, "offset": offset, "dtype": "|u1"},
            "velocities": [[int(x), int(y), float(grid["bu"][x, y]), float(grid["bv"][x, y])]
                           for x, y in zip(*np.nonzero((kind == KIND_INFLOW) |
                                                       (kind == KIND_MOVING_WALL)))]}
        f.write(kind.tobytes())
    with open(path, "w") as f:
        json.dump(doc, f)
    return doc


def load_simulation(path, quirk_serde_json=False):
    """(params dict, grid dict) of a document written by the reference or by save_simulation."""
    import os
    with open(path) as f:
        doc = loads(f.read(), quirk_serde_json=quirk_serde_json)
    base = os.path.dirname(os.path.abspath(path))
    g = doc["grid"]
    if "sidecar" not in g["cell_type"]:
        return simulation_from_json(doc)
    nx, ny = g["cell_type"]["dim"]
    kind = _resolve_array(g["cell_type"], base).astype(np.uint8)
    bu, bv = np.zeros((nx, ny)), np.zeros((nx, ny))
    for x, y, a, b in g["cell_type"].get("velocities", []):
        bu[x, y], bv[x, y] = a, b
    grid = {"size": tuple(g["size"]), "p": _resolve_array(g["pressure"], base),
            "u": _resolve_array(g["u"], base), "v": _resolve_array(g["v"], base),
            "kind": kind, "bu": bu, "bv": bv}
    prm = {"size": tuple(doc["size"]), "cell_size": tuple(doc["cell_size"]), "delt": doc["delt"],
           "gamma": doc["gamma"], "reynolds": doc["reynolds"],
           "initial_norm_squared": doc.get("initial_norm_squared"),
           "sor_absolute_epsilon": doc["sor_absolute_epsilon"],
           "max_iterations": doc["max_iterations"], "iterations": doc["iterations"],
           "time": doc["time"], "omega": doc["omega"]}
    return prm, grid
