"""The C++ host-side mirror of the reference's Rust API (include/stroemung_b200.hpp).

The reference is compiled Rust and there is no rustc in this image, so the host layer above
the C ABI exists twice: in Python (stroemung_b200/simulation.py, what the other tests use) and
in C++ (the header above).  tests/cpp/test_reference_tests.cpp replays the reference's own unit
tests through the C++ one, linked against libstroemung_b200.so and -- as the checker -- the
CPU oracle.  Here: build it, check it on the CPU as far as that goes, run every test on the GPU.
(The file sorts last so a harness problem cannot mask the parity tests under `pytest -x`.)
"""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CPP = ROOT / "tests" / "cpp"
BIN = CPP / "_build" / "test_reference_tests"

# the names test_reference_tests --list prints (checked below); one pytest item each
NAMES = [
    "math_test_du2dx", "math_test_dv2dy", "math_test_duvdx", "math_test_duvdy",
    "math_test_laplacian", "simulation_test_calculate_f", "simulation_test_calculate_g",
    "simulation_simulation_tick", "simulation_serialize", "grid_thin_boundary",
    "grid_rebuild_boundary_list", "lib_draw_cells_rolls_back_thin_walls",
    "obstacle_preset_reference_order_vs_oracle", "obstacle_preset_red_black_vs_oracle",
    "obstacle_channel_red_black_pass_kernels_vs_oracle", "stage_functions_one_by_one_vs_oracle",
    "pub_fields_written_through_the_mirror", "device_preset_equals_host_preset",
    "host_serde_json_number_parser", "host_deserialize_fixtures_like_the_reference",
    "host_nast2d_out_file_like_the_reference_converter", "host_serialize_round_trip",
    "simulation_deserialize", "pipeline_three_requests_in_flight",
    "invalid_arguments_are_errors",
]
HOST_NAMES = [n for n in NAMES if n.startswith("host_")]   # file format only: no device needed


_MAKE = {}


@pytest.fixture(scope="module")
def binary():
    # the product library and the oracle are prebuilt (__graft_entry__.build()); only the test
    # binary is (re)made here -- a few seconds of g++, also on the GPU box.  A binary that
    # travelled with the tree is used as it is if the box cannot rebuild it; the build itself
    # is asserted in test_cpp_mirror_builds_and_lists_the_reference_tests.
    r = subprocess.run(["make", "-C", str(CPP), f"PYTHON={sys.executable}"], capture_output=True,
                       text=True, timeout=900)
    _MAKE["result"] = r
    assert BIN.exists(), r.stdout[-2000:] + r.stderr[-4000:]
    return BIN


def run(binary, *args):
    return subprocess.run([str(binary), *args], capture_output=True, text=True, timeout=600)


def test_cpp_mirror_builds_and_lists_the_reference_tests(binary):
    """the header compiles as C++17 with -Wall -Wextra, links against the C ABI, and --list
    (which touches no device) names every test this file runs"""
    m = _MAKE["result"]
    assert m.returncode == 0, m.stdout[-2000:] + m.stderr[-4000:]
    r = run(binary, "--list")
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split()
    assert lines[:len(NAMES)] == NAMES
    assert "sm_100a" in r.stdout  # sb_version() of the library it linked


def test_cpp_mirror_has_no_cpu_fallback(binary):
    """without a CUDA device the very first device call throws CudaError: a failed test,
    never a CPU answer"""
    from .conftest import _cuda_device_count
    if _cuda_device_count() > 0:
        pytest.skip("a GPU is present")
    r = run(binary, "math_test_laplacian", "simulation_simulation_tick")
    assert r.returncode == 1
    assert r.stdout.count("FAILED") == 2 and "ok " not in r.stdout


@pytest.mark.parametrize("name", HOST_NAMES)
def test_reference_file_format_through_the_cpp_mirror(binary, name):
    """include/stroemung_b200_json.hpp: the reference's fixture files parse to the doubles and
    cells its `deserialize` snapshots show (incl. serde_json's 1-ulp quirk), and what the mirror
    serialises reads back identically -- host code, runs here on the CPU"""
    r = run(binary, name)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert f"ok      {name}" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in NAMES if n not in HOST_NAMES])
def test_reference_test_through_the_cpp_mirror(binary, name):
    r = run(binary, name)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert f"ok      {name}" in r.stdout
