"""The Rust -sys crate (bindings/rust/jxlb200-sys) cannot be compiled here (no cargo/rustc), so this test
keeps its declarations in step with include/jxlb200.h: same exported functions, same struct fields in the same order
with matching widths, same constants."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "jxlb200.h")).read()
RS = open(os.path.join(ROOT, "bindings", "rust", "jxlb200-sys", "src", "lib.rs")).read()

C2RS = {"uint64_t": "u64", "uint32_t": "u32", "double": "f64", "float": "f32", "size_t": "usize",
        "const uint8_t*": "*const u8"}


def c_struct_fields(name):
    body = re.search(r"typedef struct \{((?:(?!typedef).)*?)\} " + name + ";", HDR, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(const uint8_t\*|\w+) (.*)", decl)
        ctype, names = m.group(1), m.group(2)
        for n in names.split(","):
            n = n.strip()
            arr = re.match(r"(\w+)\[(\d+)\]", n)
            out.append((arr.group(1), f"[{C2RS[ctype]}; {arr.group(2)}]") if arr else (n, C2RS[ctype]))
    return out


def rs_struct_fields(name):
    body = re.search(r"pub struct " + name + r" \{(.*?)\n\}", RS, re.S).group(1)
    return [(m.group(1), m.group(2)) for m in re.finditer(r"pub (\w+): ([^,\n]+),", body)]


def test_structs_match():
    for name in ("jxlb200_image", "jxlb200_params", "jxlb200_stats"):
        assert c_struct_fields(name) == rs_struct_fields(name), name


def test_functions_match():
    c_funcs = set(re.findall(r"\b(jxlb200_\w+)\(", re.sub(r"/\*.*?\*/", "", HDR, flags=re.S)))
    rs_funcs = set(re.findall(r"pub fn (jxlb200_\w+)\(", RS))
    assert c_funcs == rs_funcs, c_funcs ^ rs_funcs


def test_constants_match():
    for name, val in re.findall(r"(JXLB200_PROPOSAL_\w+) = (\d+)", HDR):
        assert re.search(rf"pub const {name}: u32 = {val};", RS), name
    for name, val in re.findall(r"#define (JXLB200_FLAG_\w+) (\d+)u", HDR):
        assert re.search(rf"pub const {name}: u32 = {val};", RS), name
    abi = re.search(r"#define JXLB200_ABI_VERSION (\d+)", HDR).group(1)
    assert re.search(rf"pub const JXLB200_ABI_VERSION: c_int = {abi};", RS)
