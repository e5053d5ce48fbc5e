"""The reference's JSON document with a binary sidecar for the arrays (refjson.save_simulation /
load_simulation): host-side I/O, no GPU."""
import json

import numpy as np

from stroemung_b200 import presets, refjson
from tests.util import assert_bits_equal, random_fields

PRM = {"size": (9, 14), "cell_size": (0.1, 0.2), "delt": 0.005, "gamma": 0.9, "reynolds": 100.0,
       "initial_norm_squared": 0.125, "sor_absolute_epsilon": 1e-3, "max_iterations": 100,
       "iterations": 7, "time": 0.035, "omega": 1.7}


def _grid():
    g = presets.simple_inflow(PRM["size"])
    p, u, v = random_fields(*PRM["size"], 3)
    return {"p": p, "u": u, "v": v, "kind": g["kind"], "bu": g["bu"], "bv": g["bv"]}


def test_inline_document_is_the_reference_format(tmp_path):
    grid = _grid()
    doc = refjson.save_simulation(str(tmp_path / "s.json"), PRM, grid, sidecar=False)
    assert doc == refjson.simulation_to_json(PRM, grid)
    assert not (tmp_path / "s.json.bin").exists()
    prm, g2 = refjson.load_simulation(str(tmp_path / "s.json"))
    assert prm == PRM
    for k in ("p", "u", "v", "bu", "bv"):
        assert_bits_equal(g2[k], grid[k], k)
    assert np.array_equal(g2["kind"], grid["kind"])


def test_sidecar_round_trip(tmp_path):
    grid = _grid()
    path = str(tmp_path / "big.json")
    doc = refjson.save_simulation(path, PRM, grid, sidecar=True)
    assert (tmp_path / "big.json.bin").stat().st_size == 9 * 14 * (3 * 8 + 1)
    on_disk = json.loads((tmp_path / "big.json").read_text())
    assert on_disk == json.loads(json.dumps(doc))
    assert "data" not in on_disk["grid"]["pressure"]
    assert on_disk["grid"]["cell_type"]["velocities"][0] == [0, 1, 1.0, 0.0]
    # everything but the arrays is the reference's document
    ref = refjson.simulation_to_json(PRM, grid)
    assert {k: v for k, v in on_disk.items() if k != "grid"} == \
        json.loads(json.dumps({k: v for k, v in ref.items() if k != "grid"}))
    prm, g2 = refjson.load_simulation(path)
    assert prm == PRM
    for k in ("p", "u", "v", "bu", "bv"):
        assert_bits_equal(g2[k], grid[k], k)
    assert np.array_equal(g2["kind"], grid["kind"])


def test_sidecar_default_by_size(tmp_path):
    assert refjson.SIDECAR_MIN_CELLS == 512 * 512
    grid = _grid()
    refjson.save_simulation(str(tmp_path / "a.json"), PRM, grid)
    assert not (tmp_path / "a.json.bin").exists()
