"""stroemung_b200.refjson.serde_json_f64: the restatement of serde_json 1.0.140's number parser
(no `float_roundtrip`; /root/reference/Cargo.toml:18, Cargo.lock:526-527).  The crate is an
un-vendored dependency, so parity is anchored on what the reference itself holds: the double
its snapshot shows for the fixture literal that a correctly rounded parser reads differently,
and the `initial_norm_squared` computed from the parsed state (tests/test_oracle_golden.py)."""
import math
import random
import struct

import pytest

from stroemung_b200.refjson import loads, serde_json_f64


def ulps_apart(a, b):
    ia, ib = (struct.unpack("<q", struct.pack("<d", x))[0] for x in (a, b))
    return abs(ia - ib)


def test_reference_held_literal():
    """src/test_data/small_simulation_with_boundaries.json holds -0.14603099243353101; the
    reference's snapshot of the parsed state (stroemung__simulation__tests__deserialize-2.snap:42)
    shows -0.146030992433531, one ulp away from the correctly rounded value"""
    got = serde_json_f64("-0.14603099243353101")
    assert got == -0.146030992433531
    assert got != float("-0.14603099243353101")
    assert ulps_apart(got, float("-0.14603099243353101")) == 1


def test_exact_fast_path_equals_correct_rounding():
    """<= 15 significant digits and a power of ten up to 1e22: significand and power are exact
    doubles and there is ONE rounding, so the algorithm is correctly rounded there"""
    rng = random.Random(5)
    for _ in range(20000):
        digits = rng.randint(1, 15)
        sig = rng.randint(0, 10 ** digits - 1)
        frac = rng.randint(1, digits)
        body = f"{sig:0{digits}d}"
        text = (body[:digits - frac].lstrip("0") or "0") + "." + body[digits - frac:]
        if rng.random() < 0.4:
            text += f"e{rng.randint(frac - 22, frac + 22)}"   # total power of ten within +-22
        if rng.random() < 0.5:
            text = "-" + text
        assert serde_json_f64(text) == float(text), text
        assert math.copysign(1.0, serde_json_f64(text)) == math.copysign(1.0, float(text)), text


def test_shortest_round_trip_literals_are_within_one_ulp():
    """what serde_json WRITES (the shortest literal that round-trips, up to 17 digits) read back
    by its own default parser, for magnitudes a flow field holds (the power of ten stays exact,
    so there are two roundings): never more than 1 ulp off -- and not always exact, which is the
    quirk the reference's fixtures bake in.  (Far outside that range the scaling power is itself
    rounded and the error can reach 2 ulps.)"""
    rng = random.Random(6)
    off = 0
    for _ in range(20000):
        x = rng.choice((-1.0, 1.0)) * 10.0 ** rng.uniform(-6.0, 6.0)
        got = serde_json_f64(repr(x))
        d = ulps_apart(got, x)
        assert d <= 1, (repr(x), got)
        off += d
    assert off > 0


def test_digits_beyond_u64_are_dropped_not_rounded():
    # 20 digits: the 20th would overflow the u64 significand of 1844674407370955161 * 10 + 6
    assert serde_json_f64("18446744073709551616") == float(1844674407370955161) * 10.0
    assert serde_json_f64("0.123456789012345678901234567890") == \
        float(12345678901234567890) / 1e20
    # integer digits that are dropped still scale the value, fraction digits do not
    assert serde_json_f64("123456789012345678901234.75") == float(12345678901234567890) * 1e4


@pytest.mark.parametrize("text", ["01", "1.", "-", "1e", "1e+", ".5", "1.5x", ""])
def test_malformed_numbers_are_errors(text):
    with pytest.raises(ValueError):
        serde_json_f64(text)


def test_out_of_range_is_an_error_not_infinity():
    for text in ("1e309", "1e400", "123e99999999999"):
        with pytest.raises(ValueError):
            serde_json_f64(text)
    assert serde_json_f64("1e-400") == 0.0
    assert math.copysign(1.0, serde_json_f64("-1e-99999999999")) == -1.0
    assert serde_json_f64("0e99999999999") == 0.0


def test_loads_uses_the_parser_for_every_float_literal():
    doc = loads('{"a": [-0.14603099243353101, 1, 2.5, -0.0]}', quirk_serde_json=True)
    assert doc["a"][0] == -0.146030992433531 and doc["a"][1] == 1 and doc["a"][2] == 2.5
    assert math.copysign(1.0, doc["a"][3]) == -1.0
    assert loads('[-0.14603099243353101]')[0] == float("-0.14603099243353101")


@pytest.mark.parametrize("fixture,snapshot,is_sim", [
    ("src/test_data/small_simulation_with_boundaries.json",
     "stroemung__simulation__tests__deserialize-2", True),
    ("src/test_data/simple_simulation.json", "stroemung__simulation__tests__deserialize", True),
    ("src/test_data/small_grid_with_boundaries.json",
     "stroemung__grid__tests__deserialize_boundaries", False),
    ("src/test_data/simple_grid.json", "stroemung__grid__tests__deserialize", False),
    ("tests/test_data/small_data.out.json", "test_file_parsing__deserialize", False),
])
def test_every_fixture_parses_to_the_doubles_the_reference_snapshots_show(
        fixtures, snapshots, fixture, snapshot, is_sim):
    """the reference's `deserialize` tests read these files with serde_json and snapshot the
    parsed structs; the insta snapshots print every f64 with its shortest round-trip literal, so
    they show exactly which doubles the reference's parser produced"""
    import numpy as np
    doc = loads(fixtures[fixture]["raw"], quirk_serde_json=True)
    snap = snapshots[snapshot]["json"]
    got, want = (doc["grid"], snap["grid"]) if is_sim else (doc, snap)
    for name in ("pressure", "u", "v"):
        a = np.asarray(got[name]["data"], dtype=np.float64)
        b = np.asarray(want[name]["data"], dtype=np.float64)
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (fixture, name)
    if is_sim:
        for key in ("delt", "gamma", "reynolds", "sor_absolute_epsilon", "omega", "time"):
            assert float(doc[key]) == snap[key], key
