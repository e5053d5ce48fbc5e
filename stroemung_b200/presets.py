"""Initial-condition generators (host side, integer work only).

`empty`, `simple_inflow`, `obstacle` restate /root/reference/src/grid/presets.rs:8-87,
including the reference's integer circle rasteriser (`draw_circle`, :42-62).  The other
masks are the parameterised shapes the BASELINE configs name; they reuse the same
building blocks.  Each returns the "grid" dict `Simulation.try_from` takes:
kind u8 [nx, ny], bu/bv f64 [nx, ny] (boundary velocities), p/u/v = None (zeros).
The same masks can be generated on the device with `Simulation.from_preset`.
"""
import numpy as np

from .refjson import (KIND_FLUID, KIND_INFLOW, KIND_MOVING_WALL, KIND_NOSLIP,  # noqa: F401
                      KIND_OUTFLOW)


def _grid(kind, bu, bv):
    return {"size": kind.shape, "p": None, "u": None, "v": None, "kind": kind, "bu": bu,
            "bv": bv}


def empty(size):
    """src/grid/presets.rs:8-17"""
    nx, ny = size
    return _grid(np.zeros((nx, ny), np.uint8), np.zeros((nx, ny)), np.zeros((nx, ny)))


def simple_inflow(size):
    """src/grid/presets.rs:19-40: NoSlip top/bottom, Inflow [1, 0] left, Outflow right."""
    nx, ny = size
    g = empty(size)
    kind, bu = g["kind"], g["bu"]
    kind[:, 0] = KIND_NOSLIP
    kind[:, ny - 1] = KIND_NOSLIP
    kind[0, 1:ny - 1] = KIND_INFLOW
    bu[0, 1:ny - 1] = 1.0
    kind[nx - 1, 1:ny - 1] = KIND_OUTFLOW
    return g


def draw_circle(kind, x, y, radius):
    """src/grid/presets.rs:42-62 (`radius as usize` truncates; saturating bounds)."""
    nx, ny = kind.shape
    r = int(radius)
    for xi in range(max(x - r, 0), x + r):
        if xi >= nx:
            continue
        x_dist = xi - x
        for yi in range(max(y - r, 0), y + r):
            if yi >= ny:
                continue
            y_dist = yi - y
            if float(np.sqrt(np.float64(x_dist * x_dist + y_dist * y_dist))) < radius:
                kind[xi, yi] = KIND_NOSLIP


def obstacle(size):
    """src/grid/presets.rs:64-87: simple_inflow + circle at (20, ny/2), r = 5."""
    g = simple_inflow(size)
    draw_circle(g["kind"], 20, size[1] // 2, 5.0)
    return g


def channel_circle(size, cx, cy, radius):
    """Karman configuration: simple_inflow channel + generalised draw_circle."""
    g = simple_inflow(size)
    draw_circle(g["kind"], cx, cy, radius)
    return g


def backward_step(size, step_len, step_top):
    """NoSlip walls and a NoSlip block x < step_len, y >= step_top; inflow above the block."""
    nx, ny = size
    g = empty(size)
    kind, bu = g["kind"], g["bu"]
    kind[:, 0] = KIND_NOSLIP
    kind[:, ny - 1] = KIND_NOSLIP
    kind[0, 1:ny - 1] = KIND_INFLOW
    kind[nx - 1, 1:ny - 1] = KIND_OUTFLOW
    kind[:step_len, step_top:] = KIND_NOSLIP
    bu[0, 1:min(step_top, ny - 1)] = 1.0
    return g


def cavity(size, lid_u=1.0):
    """Lid-driven cavity (extension kind MovingWall on y == 0; not expressible in the reference)."""
    nx, ny = size
    g = empty(size)
    kind, bu = g["kind"], g["bu"]
    kind[0, :] = kind[nx - 1, :] = KIND_NOSLIP
    kind[:, 0] = kind[:, ny - 1] = KIND_NOSLIP
    kind[1:nx - 1, 0] = KIND_MOVING_WALL
    bu[1:nx - 1, 0] = lid_u
    return g
