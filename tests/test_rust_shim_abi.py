"""rust/src/gpu.rs cannot be compiled here (no rustc in the image), so its `extern "C"` block and
`#[repr(C)]` structs are checked against include/stroemung_b200.h textually: every function the
header declares is bound, argument by argument and with the same return type, and the structs
carry the header's fields in the header's order and widths."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "stroemung_b200.h").read_text()
SHIM = (ROOT / "rust" / "src" / "gpu.rs").read_text()

SCALARS = {"size_t": "usize", "uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32",
           "uint8_t": "u8", "double": "f64", "sb_field": "i32", "sb_status": "i32", "void": "()",
           "char": "c_char"}
STRUCTS = {"sb_params": "SbParams", "sb_sim": "SbSim", "sb_state": "SbState",
           "sb_boundary_velocity": "SbBoundaryVelocity"}


def c_type_to_rust(decl, is_return=False):
    """'const double u[9]' -> '*const f64', 'sb_sim **out' -> '*mut *mut SbSim', ..."""
    decl = decl.strip()
    array = bool(re.search(r"\[[^\]]*\]\s*$", decl))
    decl = re.sub(r"\[[^\]]*\]\s*$", "", decl)
    const = decl.startswith("const ")
    decl = decl[6:] if const else decl
    stars = decl.count("*") + (1 if array else 0)
    base = decl.replace("*", " ").split()[0]
    rust = STRUCTS.get(base) or SCALARS.get(base) or ("c_void" if base == "void" else None)
    assert rust, decl
    if base == "void" and stars:
        rust = "c_void"
    for level in range(stars):
        # only the innermost pointee can be const in this header
        rust = ("*const " if const and level == 0 else "*mut ") + rust
    return rust


def header_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w \*]*?)\b(sb_\w+)\s*\(([^;{}]*?)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        if ret.startswith("typedef"):
            continue
        argl = [] if args in ("", "void") else [c_type_to_rust(a) for a in args.split(",")]
        out[name] = (argl, c_type_to_rust(ret + " x" if "*" not in ret else ret))
    return out


def shim_functions():
    block = re.search(r'extern "C" \{(.*?)\n\}', SHIM, flags=re.S).group(1)
    out = {}
    for m in re.finditer(r"fn (sb_\w+)\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block):
        name, args, ret = m.group(1), m.group(2).strip(), (m.group(3) or "()").strip()
        argl = [a.split(":", 1)[1].strip() for a in args.split(",")] if args else []
        out[name] = (argl, ret)
    return out


def test_every_header_function_is_bound_with_the_same_signature():
    h, r = header_functions(), shim_functions()
    from stroemung_b200 import _capi
    assert set(h) == set(_capi.SYMBOLS), set(h) ^ set(_capi.SYMBOLS)   # the parser saw them all
    assert set(r) == set(h), set(r) ^ set(h)
    for name in sorted(h):
        assert r[name] == h[name], (name, r[name], h[name])


def c_struct_fields(name):
    body = re.search(r"typedef struct \{([^{}]*?)\}\s*" + name + r"\s*;", HEADER, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for stmt in body.split(";"):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        ctype, rest = stmt.split(" ", 1)
        for item in rest.split(","):
            item = item.strip()
            m = re.match(r"(\w+)(?:\[(\d+)\])?$", item)
            rust = SCALARS[ctype]
            fields.append((m.group(1), f"[{rust}; {m.group(2)}]" if m.group(2) else rust))
    return fields


def rust_struct_fields(name):
    body = re.search(r"#\[repr\(C\)\][^{]*?pub struct " + name + r" \{(.*?)\n\}", SHIM,
                     flags=re.S).group(1)
    return [(m.group(1), m.group(2).strip())
            for m in re.finditer(r"pub (\w+):\s*([^,\n]+),", body)]


def test_repr_c_structs_match_the_header_field_for_field():
    for c, r in (("sb_params", "SbParams"), ("sb_boundary_velocity", "SbBoundaryVelocity"),
                 ("sb_state", "SbState")):
        assert rust_struct_fields(r) == c_struct_fields(c), (c, r)


def test_enum_constants_match_the_header():
    for name, value in re.findall(r"pub const (SB_\w+): \w+ = (\d+);", SHIM):
        m = re.search(r"\b" + name + r"\s*=\s*(\d+)", HEADER) or \
            re.search(r"#define " + name + r"\s+(\d+)", HEADER)
        assert m and int(m.group(1)) == int(value), name


def test_shim_delimiters_balance_and_items_are_well_formed():
    """the cheapest stand-in for a compiler: after stripping comments, strings and char
    literals every (, [ and { closes in order; every `fn` has a body or a `;`"""
    text = re.sub(r"//[^\n]*", "", SHIM)
    text = re.sub(r'"(?:\\.|[^"\\])*"', '""', text)
    text = re.sub(r"b?'(?:\\.|[^'\\])'", "' '", text)
    pairs = {")": "(", "]": "[", "}": "{"}
    stack = []
    for i, ch in enumerate(text):
        if ch in "([{":
            stack.append((ch, i))
        elif ch in pairs:
            assert stack and stack[-1][0] == pairs[ch], (ch, text[max(0, i - 80):i + 20])
            stack.pop()
    assert not stack, stack[-1]
    for m in re.finditer(r"\bfn\s+\w+\s*\(", text):
        depth, j = 0, m.end() - 1
        while True:
            depth += text[j] == "("
            depth -= text[j] == ")"
            j += 1
            if depth == 0:
                break
        tail = text[j:j + 200]
        assert re.match(r"\s*(->\s*[^;{]+)?\s*[;{]", tail), tail[:80]
