"""Shared helpers for the parity tests (inputs, comparisons)."""
import numpy as np

from oracle import pyoracle as po

DEFAULTS = dict(delx=0.1, dely=0.2, delt=0.005, gamma=0.9, reynolds=100.0,
                sor_absolute_epsilon=0.001, max_iterations=100, omega=1.7)


def bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


def assert_bits_equal(a, b, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = np.nonzero(bits(a) != bits(b))
    if bad[0].size:
        first = tuple(int(x[0]) for x in bad)
        raise AssertionError(f"{what}: {bad[0].size} of {a.size} differ, first at {first}: "
                             f"{a[first]!r} vs {b[first]!r}")


def splitmix64(seed):
    """Deterministic RNG of SURVEY.md section 8d (seed 0x5EED5EED)."""
    state = np.uint64(seed)
    while True:
        state = np.uint64((int(state) + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
        z = int(state)
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        yield z ^ (z >> 31)


def random_mask(nx, ny, seed, n_blocks=6, kinds=(1, 1, 2, 3)):
    """Channel ring plus random >=2-wide blocks of mixed boundary kinds; always classifiable.

    Exercises every edge class (corners, flat edges, interior None cells) and interior
    Inflow / Outflow cells, which the reference's own tests never reach."""
    rng = np.random.default_rng(seed)
    kind, bu, bv = po.preset("simple_inflow", nx, ny)
    kind = kind.copy()
    bu = bu.copy()
    bv = bv.copy()
    for _ in range(n_blocks):
        for _try in range(20):
            w, h = int(rng.integers(2, max(3, nx // 4))), int(rng.integers(2, max(3, ny // 3)))
            if nx - 3 - w <= 2 or ny - 3 - h <= 2:
                break
            x0, y0 = int(rng.integers(2, nx - 2 - w)), int(rng.integers(2, ny - 2 - h))
            k = int(rng.choice(kinds))
            trial = kind.copy()
            trial[x0:x0 + w, y0:y0 + h] = k
            tbu, tbv = bu.copy(), bv.copy()
            if k == 3:
                tbu[x0:x0 + w, y0:y0 + h] = rng.uniform(-1, 1)
                tbv[x0:x0 + w, y0:y0 + h] = rng.uniform(-1, 1)
            try:
                po.OracleSim(nx, ny, kind=trial, bu=tbu, bv=tbv, **DEFAULTS)
            except po.BoundaryTooThin:
                continue
            kind, bu, bv = trial, tbu, tbv
            break
    return kind, bu, bv


def random_fields(nx, ny, seed, scale=1.0):
    rng = np.random.default_rng(seed + 1000)
    return (rng.uniform(-1, 1, (nx, ny)) * scale, rng.uniform(-1, 1, (nx, ny)) * scale,
            rng.uniform(-1, 1, (nx, ny)) * scale)


def unfinalized(nx, ny, kind, bu, bv, p=None, u=None, v=None, **over):
    prm = dict(DEFAULTS)
    prm.update(over)
    return {"size": (nx, ny), "cell_size": (prm["delx"], prm["dely"]), "delt": prm["delt"],
            "gamma": prm["gamma"], "reynolds": prm["reynolds"],
            "sor_absolute_epsilon": prm["sor_absolute_epsilon"],
            "max_iterations": prm["max_iterations"], "omega": prm["omega"],
            "initial_norm_squared": prm.get("initial_norm_squared"),
            "iterations": prm.get("iterations", 0), "time": prm.get("time", 0.0),
            "grid": {"p": p, "u": u, "v": v, "kind": kind, "bu": bu, "bv": bv}}


def oracle_from(unf, **kw):
    g = unf["grid"]
    return po.OracleSim(unf["size"][0], unf["size"][1], delx=unf["cell_size"][0],
                        dely=unf["cell_size"][1], delt=unf["delt"], gamma=unf["gamma"],
                        reynolds=unf["reynolds"],
                        sor_absolute_epsilon=unf["sor_absolute_epsilon"],
                        max_iterations=unf["max_iterations"], omega=unf["omega"],
                        kind=g["kind"], p=g["p"], u=g["u"], v=g["v"], bu=g["bu"], bv=g["bv"],
                        initial_norm_squared=unf.get("initial_norm_squared"),
                        iterations=unf.get("iterations", 0), time=unf.get("time", 0.0), **kw)
