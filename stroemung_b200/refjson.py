"""The reference's on-disk JSON format (host-side I/O, no compute).

Mirrors what serde derives for `UnfinalizedSimulation` / `UnfinalizedSimulationGrid`
(/root/reference/src/simulation.rs:30-44, src/grid/mod.rs:85-92):

* arrays are ndarray-serde documents ``{"v": 1, "dim": [nx, ny], "data": [...]}``
  with ``data`` row-major, i.e. index ``(x, y)`` and y contiguous;
* a cell is ``"Fluid"``, ``{"Boundary": "NoSlip"}``, ``{"Boundary": "Outflow"}`` or
  ``{"Boundary": {"Inflow": {"velocity": [u, v]}}}`` (src/cell.rs:6-23).

Device-side the cell mask is a u8 kind (0 Fluid, 1 NoSlip, 2 Outflow, 3 Inflow,
4 MovingWall [extension]) plus a sparse table of boundary velocities.

Note: the reference parses with serde_json 1.0.140 without
`float_roundtrip`; one fixture literal (-0.14603099243353101) is read 1 ulp low
by it.  Python's parser is correctly rounded; `quirk_serde_json=True` reproduces
the reference's value for that literal so golden snapshots can be matched.
"""
import json

import numpy as np

KIND_FLUID, KIND_NOSLIP, KIND_OUTFLOW, KIND_INFLOW, KIND_MOVING_WALL = range(5)

# literal -> double the reference's parser produced (pinned by
# src/snapshots/stroemung__simulation__tests__deserialize-2.snap:42)
_SERDE_QUIRKS = {"-0.14603099243353101": -0.146030992433531}


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
    """json.loads; optionally with the reference parser's 1-ulp quirk."""
    if not quirk_serde_json:
        return json.loads(text)
    return json.loads(text, parse_float=lambda s: _SERDE_QUIRKS.get(s, float(s)))


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
