#!/usr/bin/env python3
"""Extract the reference's golden vectors into small committed fixtures.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py [/root/reference]

Writes, next to this script:
  kat.json        the exact-equality known-answer tests of src/math.rs:192-402
                  and src/simulation.rs:449-569 (inputs + expected f64)
  snapshots.json  the JSON bodies of every insta snapshot on the hot path
                  (src/snapshots, src/grid/snapshots, tests/snapshots)
  fixtures.json   the reference's input fixtures (src/test_data/*.json,
                  tests/test_data/*.json, python/test_data/*), the 440-byte
                  NaSt2D binary as hex

Only test DATA is extracted (numbers and JSON documents); no reference source
code is copied.  Floats are serialised with repr() so they round-trip exactly.
/root/reference does not exist on the GPU box: tests read only these files.
"""
import ast
import json
import re
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent


def _test_cases(src: str, fn_name: str):
    """Return the `test_cases` array literal of one Rust #[test] fn as Python data."""
    start = src.index(f"fn {fn_name}()")
    start = src.index("let test_cases = [", start) + len("let test_cases = ")
    depth = 0
    for i in range(start, len(src)):
        ch = src[i]
        if ch in "[(":
            depth += 1
        elif ch in "])":
            depth -= 1
            if depth == 0:
                end = i + 1
                break
    text = src[start:end]
    text = re.sub(r"//[^\n]*", "", text)      # comments
    text = text.replace("array!", "")          # ndarray macro -> nested list
    return ast.literal_eval(text)


def extract_kats(ref: Path):
    math_rs = (ref / "src/math.rs").read_text()
    sim_rs = (ref / "src/simulation.rs").read_text()
    out = {}
    # (array, delx, gamma, expected)
    out["du2dx"] = [
        {"u": c[0], "delx": c[1], "gamma": c[2], "expected": c[3]}
        for c in _test_cases(math_rs, "test_du2dx")
    ]
    out["dv2dy"] = [
        {"v": c[0], "dely": c[1], "gamma": c[2], "expected": c[3]}
        for c in _test_cases(math_rs, "test_dv2dy")
    ]
    out["duvdx"] = [
        {"u": c[0], "v": c[1], "delx": c[2], "gamma": c[3], "expected": c[4]}
        for c in _test_cases(math_rs, "test_duvdx")
    ]
    out["duvdy"] = [
        {"u": c[0], "v": c[1], "dely": c[2], "gamma": c[3], "expected": c[4]}
        for c in _test_cases(math_rs, "test_duvdy")
    ]
    out["laplacian"] = [
        {"e": c[0], "delx": c[1], "dely": c[2], "expected": c[3]}
        for c in _test_cases(math_rs, "test_lapacian")
    ]
    for name in ("calculate_f", "calculate_g"):
        out[name] = [
            {"u": c[0], "v": c[1], "delx": c[2], "dely": c[3], "delt": c[4],
             "gamma": c[5], "reynolds": c[6], "expected": c[7]}
            for c in _test_cases(sim_rs, f"test_{name}")
        ]
    # the two assert pairs of simulation_tick (src/simulation.rs:597-598, 605-606)
    m = re.findall(r"assert_eq!\((?:last_)?sor_iterations, (\d+)\);\s*"
                   r"assert_eq!\((?:last_)?norm_squared, ([0-9.e+-]+)\);", sim_rs)
    out["simulation_tick_asserts"] = [
        {"sor_iterations": int(a), "norm_squared": float(b)} for a, b in m
    ]
    assert len(out["simulation_tick_asserts"]) == 2
    n = sum(len(v) for k, v in out.items() if k != "simulation_tick_asserts")
    assert n == 32, n
    return out


def _snap_body(path: Path):
    text = path.read_text()
    # insta format: '---\n<yaml header>\n---\n<body>'
    parts = text.split("\n---\n", 1)
    return parts[1]


def extract_snapshots(ref: Path):
    out = {}
    for d in ("src/snapshots", "src/grid/snapshots", "tests/snapshots"):
        for f in sorted((ref / d).glob("*.snap")):
            body = _snap_body(f)
            key = f.name[: -len(".snap")]
            try:
                out[key] = {"json": json.loads(body)}
            except json.JSONDecodeError:
                out[key] = {"text": body}
    return out


def extract_fixtures(ref: Path):
    out = {}
    for d in ("src/test_data", "tests/test_data"):
        for f in sorted((ref / d).glob("*.json")):
            # keep the raw text too: one literal is mis-parsed by the reference's
            # serde_json by 1 ulp and tests need to see it
            out[f"{d}/{f.name}"] = {"json": json.loads(f.read_text()), "raw": f.read_text()}
    pd = ref / "python/test_data"
    out["python/test_data/small_data.out"] = {"hex": (pd / "small_data.out").read_bytes().hex()}
    for name in ("small_data.out_expected.json", "small_data.out_rust_expected.json"):
        out[f"python/test_data/{name}"] = {"json": json.loads((pd / name).read_text())}
    return out


def main():
    ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    (HERE / "kat.json").write_text(json.dumps(extract_kats(ref), indent=1) + "\n")
    (HERE / "snapshots.json").write_text(json.dumps(extract_snapshots(ref), indent=1) + "\n")
    (HERE / "fixtures.json").write_text(json.dumps(extract_fixtures(ref), indent=1) + "\n")
    print("wrote kat.json, snapshots.json, fixtures.json")


if __name__ == "__main__":
    main()
