#!/usr/bin/env python3
"""include/hvx.h -> rust/helio-voxel-cuda/src/ffi.rs, mechanically.

The header is the single source of truth of the C ABI; this script parses its (deliberately plain) C --
`#define` constants, `typedef enum`, `typedef struct`, opaque handles and function prototypes -- and emits the
`extern "C"` image the Rust crate binds.  tests/test_ffi_consistency.py regenerates the file and compares it with
the committed one, and checks every struct's size and field offsets three ways (this parser's C layout, gcc's
sizeof / offsetof on the real header, and the repr(C) layout of the emitted Rust).

  python tools/gen_rust_ffi.py            # rewrite rust/helio-voxel-cuda/src/ffi.rs
  python tools/gen_rust_ffi.py --check    # exit 1 if the committed file is stale
"""
from __future__ import annotations

import re
import sys
from dataclasses import dataclass, field
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "hvx.h"
OUT = ROOT / "rust" / "helio-voxel-cuda" / "src" / "ffi.rs"

SCALARS = {  # C type -> (Rust type, size, alignment)
    "uint8_t": ("u8", 1, 1), "int8_t": ("i8", 1, 1), "uint16_t": ("u16", 2, 2), "int16_t": ("i16", 2, 2),
    "uint32_t": ("u32", 4, 4), "int32_t": ("i32", 4, 4), "uint64_t": ("u64", 8, 8), "int64_t": ("i64", 8, 8),
    "float": ("f32", 4, 4), "double": ("f64", 8, 8), "int": ("c_int", 4, 4), "char": ("c_char", 1, 1),
    "size_t": ("usize", 8, 8),
}


@dataclass
class Field:
    name: str
    ctype: str
    count: int  # 0 = scalar, n = array of n


@dataclass
class Struct:
    name: str
    fields: list = field(default_factory=list)
    comment: str = ""


@dataclass
class Function:
    name: str
    ret: str
    args: list  # (ctype string incl. const / pointers, name)


@dataclass
class Header:
    defines: list = field(default_factory=list)   # (name, value text)
    enums: list = field(default_factory=list)     # (enum name, [(constant, int)])
    structs: list = field(default_factory=list)
    opaque: list = field(default_factory=list)
    functions: list = field(default_factory=list)


def strip_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def parse_header(path=HEADER) -> Header:
    text = strip_comments(path.read_text())
    h = Header()
    for m in re.finditer(r"^#define\s+(HVX_[A-Z0-9_]+)\s+(\S+)\s*$", text, flags=re.M):
        if m.group(1) != "HVX_H":
            h.defines.append((m.group(1), m.group(2)))
    for m in re.finditer(r"typedef\s+enum\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        value, items = -1, []
        for part in m.group(1).split(","):
            part = part.strip()
            if not part:
                continue
            if "=" in part:
                name, v = (x.strip() for x in part.split("="))
                value = int(v, 0)
            else:
                name, value = part, value + 1
            items.append((name, value))
        h.enums.append((m.group(2), items))
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        s = Struct(m.group(2))
        for decl in m.group(1).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            ctype, rest = decl.split(" ", 1)
            for item in rest.split(","):
                item = item.strip()
                am = re.fullmatch(r"(\w+)\[(\d+)\]", item)
                s.fields.append(Field(am.group(1), ctype, int(am.group(2))) if am else Field(item, ctype, 0))
        h.structs.append(s)
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s+(\w+)\s*;", text):
        h.opaque.append(m.group(2))
    body = re.sub(r"typedef\s+(enum|struct)\s*\{.*?\}\s*\w+\s*;", " ", text, flags=re.S)
    for m in re.finditer(r"^\s*((?:const\s+)?\w+\s*\**)\s*(hvx_\w+)\s*\((.*?)\)\s*;", body, flags=re.S | re.M):
        args = []
        raw = " ".join(m.group(3).split())
        if raw and raw != "void":
            for a in raw.split(","):
                a = a.strip()
                am = re.fullmatch(r"(.*?)(\w+)\[\d*\]", a)
                if am:   # `const int64_t focus[3]` is a pointer
                    args.append((am.group(1).strip() + "*", am.group(2)))
                    continue
                am = re.fullmatch(r"(.*?[\s\*])(\w+)", a)
                args.append((am.group(1).replace(" *", "*").strip(), am.group(2)))
        h.functions.append(Function(m.group(2), " ".join(m.group(1).split()).replace(" *", "*"), args))
    return h


def rust_type(ctype: str) -> str:
    """C parameter / field type -> Rust."""
    t = ctype.strip()
    stars = t.count("*")
    base = t.replace("*", "").strip()
    const = base.startswith("const ")
    base = base[6:].strip() if const else base
    if base == "void":
        core = "c_void"
    elif base in SCALARS:
        core = SCALARS[base][0]
    else:
        core = base
    if stars == 0:
        return core
    out = core
    for level in range(stars):
        innermost = level == 0
        out = ("*const " if (const and innermost) else "*mut ") + out
    return out


def c_layout(h: Header):
    """{struct: (size, align, {field: (offset, size)})} by the natural-alignment rule of the C ABI."""
    done = {}
    for s in h.structs:
        offset, align, fields = 0, 1, {}
        for f in s.fields:
            if f.ctype in SCALARS:
                size, al = SCALARS[f.ctype][1], SCALARS[f.ctype][2]
            else:
                size, al = done[f.ctype][0], done[f.ctype][1]
            total = size * max(f.count, 1)
            offset = (offset + al - 1) // al * al
            fields[f.name] = (offset, total)
            offset += total
            align = max(align, al)
        done[s.name] = ((offset + align - 1) // align * align, align, fields)
    return done


def emit(h: Header) -> str:
    out = ["//! Raw `extern \"C\"` image of include/hvx.h.  GENERATED by tools/gen_rust_ffi.py -- do not edit;",
           "//! tests/test_ffi_consistency.py fails when this file and the header disagree.",
           "#![allow(non_camel_case_types, non_upper_case_globals, clippy::too_many_arguments)]",
           "use core::ffi::{c_char, c_int, c_void};", ""]
    for name, value in h.defines:
        v = value.rstrip("uU")
        out.append(f"pub const {name}: u32 = {v};")
    out.append("")
    for ename, items in h.enums:
        out.append(f"// {ename}")
        for name, value in items:
            out.append(f"pub const {name}: c_int = {value};")
        out.append("")
    for name in h.opaque:
        out.append("#[repr(C)]")
        out.append(f"pub struct {name} {{ _private: [u8; 0] }}")
    out.append("")
    floaty = {s.name for s in h.structs if any(f.ctype in ("float", "double") for f in s.fields)}
    changed = True
    while changed:   # a struct holding a float-bearing struct cannot derive Eq either
        changed = False
        for s in h.structs:
            if s.name not in floaty and any(f.ctype in floaty for f in s.fields):
                floaty.add(s.name)
                changed = True
    for s in h.structs:
        out.append("#[repr(C)]")
        out.append("#[derive(Clone, Copy, Debug, Default, PartialEq)]" if s.name in floaty else "#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]")
        out.append(f"pub struct {s.name} {{")
        for f in s.fields:
            t = rust_type(f.ctype)
            out.append(f"    pub {f.name}: {'[' + t + '; ' + str(f.count) + ']' if f.count else t},")
        out.append("}")
        out.append("")
    out.append('extern "C" {')
    for fn in h.functions:
        args = ", ".join(f"{n}: {rust_type(t)}" for t, n in fn.args)
        ret = "" if fn.ret == "void" else f" -> {rust_type(fn.ret)}"
        out.append(f"    pub fn {fn.name}({args}){ret};")
    out.append("}")
    return "\n".join(out) + "\n"


def main():
    text = emit(parse_header())
    if "--check" in sys.argv:
        if not OUT.exists() or OUT.read_text() != text:
            print(f"{OUT} is stale: run python tools/gen_rust_ffi.py")
            sys.exit(1)
        return
    OUT.write_text(text)
    print(f"wrote {OUT}: {text.count(chr(10))} lines")


if __name__ == "__main__":
    main()
