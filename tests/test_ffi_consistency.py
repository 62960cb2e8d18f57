"""The C header is the ABI; the Rust binding (rust/helio-voxel-cuda/src/ffi.rs) is generated from it.  These tests keep
the three views in step without a Rust toolchain: the generator's parse of include/hvx.h, gcc's own sizeof / offsetof
of the real header, and an independent reading of the committed ffi.rs with repr(C) layout rules."""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
import gen_rust_ffi as G  # noqa: E402

from helio_b200 import _ffi  # noqa: E402

FFI_RS = ROOT / "rust" / "helio-voxel-cuda" / "src" / "ffi.rs"
RUST_SCALARS = {"u8": (1, 1), "i8": (1, 1), "u16": (2, 2), "i16": (2, 2), "u32": (4, 4), "i32": (4, 4), "u64": (8, 8),
                "i64": (8, 8), "f32": (4, 4), "f64": (8, 8), "c_int": (4, 4), "c_char": (1, 1), "usize": (8, 8)}


def test_committed_ffi_rs_is_what_the_header_generates():
    assert FFI_RS.read_text() == G.emit(G.parse_header()), "run python tools/gen_rust_ffi.py"


def test_every_prototype_is_exported_and_bound():
    h = G.parse_header()
    names = [f.name for f in h.functions]
    assert len(names) == len(set(names)) and set(names) == set(_ffi.EXPORTS)
    rust_fns = dict(re.findall(r"pub fn (hvx_\w+)\((.*?)\)", FFI_RS.read_text()))
    assert set(rust_fns) == set(names)
    for f in h.functions:
        rust_args = [a for a in rust_fns[f.name].split(", ") if a]
        assert len(rust_args) == len(f.args), f.name
        for (ctype, cname), rust in zip(f.args, rust_args):
            rname, rtype = rust.split(": ")
            assert rname == cname, (f.name, cname, rname)
            assert rtype.count("*") == ctype.count("*"), (f.name, cname)        # same indirection
            assert ("*const" in rtype) == (ctype.startswith("const ") and "*" in ctype), (f.name, cname)


def _gcc_layout(h, tmp_path):
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{G.HEADER}"', "int main(void) {"]
    for s in h.structs:
        lines.append(f'  printf("S {s.name} %zu %zu\\n", sizeof({s.name}), _Alignof({s.name}));')
        for f in s.fields:
            lines.append(f'  printf("F {s.name} {f.name} %zu %zu\\n", offsetof({s.name}, {f.name}), sizeof((({s.name}*)0)->{f.name}));')
    lines += ["  return 0;", "}"]
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines))
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    sizes, fields = {}, {}
    for line in out.splitlines():
        parts = line.split()
        if parts[0] == "S":
            sizes[parts[1]] = (int(parts[2]), int(parts[3]))
        else:
            fields.setdefault(parts[1], {})[parts[2]] = (int(parts[3]), int(parts[4]))
    return sizes, fields


def _rust_layout():
    """repr(C) layout of every struct in the committed ffi.rs, read independently of the generator."""
    text = FFI_RS.read_text()
    done, order = {}, []
    for m in re.finditer(r"#\[repr\(C\)\]\n#\[derive\([^\]]*\)\]\npub struct (\w+) \{\n(.*?)\n\}", text, flags=re.S):
        name, body = m.group(1), m.group(2)
        offset, align, fields = 0, 1, {}
        for line in body.splitlines():
            fm = re.fullmatch(r"\s*pub (\w+): (.+),", line)
            fname, ftype = fm.group(1), fm.group(2)
            count = 1
            am = re.fullmatch(r"\[(\w+); (\d+)\]", ftype)
            if am:
                ftype, count = am.group(1), int(am.group(2))
            size, al = RUST_SCALARS[ftype] if ftype in RUST_SCALARS else (done[ftype][0], done[ftype][1])
            offset = (offset + al - 1) // al * al
            fields[fname] = (offset, size * count)
            offset += size * count
            align = max(align, al)
        done[name] = ((offset + align - 1) // align * align, align, fields)
        order.append(name)
    return done, order


def test_struct_layouts_agree_between_gcc_the_parser_and_the_rust_binding(tmp_path):
    h = G.parse_header()
    sizes, fields = _gcc_layout(h, tmp_path)
    parsed = G.c_layout(h)
    rust, order = _rust_layout()
    assert order == [s.name for s in h.structs]
    for s in h.structs:
        assert parsed[s.name][:2] == sizes[s.name], s.name
        assert rust[s.name][:2] == sizes[s.name], s.name
        assert parsed[s.name][2] == fields[s.name] == rust[s.name][2], s.name
    # the PODs that must equal the reference's byte for byte
    assert sizes["hvx_vertex"][0] == 32 and sizes["hvx_emission_counters"][0] == 32 and sizes["hvx_chunk_desc"][0] == 32
    assert sizes["hvx_voxel_edit"][0] == 32 and sizes["hvx_page_table_entry"][0] == 48 and sizes["hvx_gather_job"][0] == 64


def test_constants_match_the_python_binding():
    text = FFI_RS.read_text()
    consts = {n: int(v) for n, v in re.findall(r"pub const (HVX_\w+): (?:u32|c_int) = (-?\d+);", text)}
    for name in dir(_ffi):
        if name.startswith("HVX_") and isinstance(getattr(_ffi, name), int):
            assert consts[name] == getattr(_ffi, name), name
    assert consts["HVX_BUF_COUNT"] == 25 and consts["HVX_ABI_VERSION"] == _ffi.load().hvx_abi_version()


def test_the_hand_written_rust_only_names_things_the_binding_has():
    """No Rust toolchain here: at least every `ffi::name` the wrappers use must exist, with the fields they set."""
    text = FFI_RS.read_text()
    known = set(re.findall(r"pub (?:const|struct|fn) (\w+)", text))
    struct_fields = {m.group(1): set(re.findall(r"pub (\w+):", m.group(2)))
                     for m in re.finditer(r"pub struct (\w+) \{\n(.*?)\n\}", text, flags=re.S)}
    for rs in ("lib.rs", "extraction.rs"):
        src = (FFI_RS.parent / rs).read_text()
        for name in set(re.findall(r"(?<!core::)ffi::(\w+)", src)):
            assert name in known, f"{rs}: ffi::{name} does not exist"
        for block in re.findall(r"use crate::ffi::\{(.*?)\};", src, flags=re.S):
            for name in re.findall(r"\w+", block):
                assert name in known, f"{rs}: crate::ffi::{name} does not exist"
        for sname, body in re.findall(r"ffi::(hvx_\w+) \{(.*?)\}", src, flags=re.S):
            for fname in re.findall(r"(?<![:\w])(\w+):(?!:)", body):
                assert fname in struct_fields[sname], f"{rs}: {sname} has no field {fname}"
