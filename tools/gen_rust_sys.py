#!/usr/bin/env python
"""Generates rust/omb200-sys/src/lib.rs from include/omb200.h: every constant, struct, opaque handle, callback type and
`omb_*` function of the C ABI as Rust FFI declarations (what bindgen would emit, without needing libclang).

    python tools/gen_rust_sys.py            # rewrites rust/omb200-sys/src/lib.rs
    python tools/gen_rust_sys.py --check    # exit 1 if the committed file is stale

`parse_header()` is also what tests/test_abi.py uses to compare the crate's `extern "C"` block (names and arity) and the
exported symbols of libomb200.so against the header.
"""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "omb200.h")
OUT = os.path.join(ROOT, "rust", "omb200-sys", "src", "lib.rs")

PRIM = {"float": "f32", "double": "f64", "int": "i32", "int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64",
        "uint16_t": "u16", "uint8_t": "u8", "size_t": "usize", "char": "c_char", "void": "c_void"}


def strip_comments(src: str) -> str:
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)


def rust_type(ctype: str, structs) -> str:
    """'const float*' -> '*const f32', 'omb_spectrogram**' -> '*mut *mut omb_spectrogram'."""
    t = ctype.strip()
    stars = t.count("*")
    t = t.replace("*", " ").strip()
    const = False
    words = [w for w in t.split() if w not in ("struct",)]
    if words and words[0] == "const":
        const = True
        words = words[1:]
    base = " ".join(words)
    r = PRIM.get(base, base)
    for i in range(stars):
        r = ("*const " if (const and i == 0) else "*mut ") + r
    return r


def parse_param(p: str):
    """One C parameter -> (name, rust type)."""
    p = p.strip()
    if p == "void" or not p:
        return None
    m = re.match(r"^(.*?)([A-Za-z_]\w*)((?:\[[^\]]*\])*)$", p)
    ctype, name, arr = m.group(1).strip(), m.group(2), m.group(3)
    dims = re.findall(r"\[([^\]]*)\]", arr)
    if dims:
        const = ctype.startswith("const")
        inner = rust_type(ctype, None)
        for d in reversed(dims[1:]):
            inner = f"[{inner}; {d.replace('OMB_MAX_CHANNELS', 'OMB_MAX_CHANNELS as usize')}]"
        return name, ("*const " if const else "*mut ") + inner
    return name, rust_type(ctype, None)


def parse_header(path: str = HEADER):
    src = strip_comments(open(path).read())
    out = {"defines": [], "enums": [], "structs": [], "opaque": [], "callbacks": [], "functions": []}
    for m in re.finditer(r"#define\s+(OMB_[A-Z0-9_]+)\s+(-?\d+)", src):
        out["defines"].append((m.group(1), int(m.group(2))))
    for m in re.finditer(r"(?:typedef\s+)?enum\s*(\w*)\s*\{(.*?)\}\s*(\w*)\s*;", src, flags=re.S):
        vals = []
        nxt = 0
        for item in m.group(2).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = [x.strip() for x in item.split("=")]
                nxt = int(v, 0)
            else:
                k = item
            vals.append((k, nxt))
            nxt += 1
        out["enums"].append((m.group(3) or m.group(1) or "", vals))
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            mm = re.match(r"^(.*?[\s\*])([A-Za-z_]\w*(?:\[[^\]]*\])*(?:\s*,\s*[A-Za-z_]\w*(?:\[[^\]]*\])*)*)$", decl)
            ctype, names = mm.group(1).strip(), mm.group(2)
            for nm in names.split(","):
                nm = nm.strip()
                dims = re.findall(r"\[([^\]]*)\]", nm)
                nm = re.sub(r"\[.*", "", nm)
                rt = rust_type(ctype, None)
                for d in reversed(dims):
                    rt = f"[{rt}; {d.replace('OMB_MAX_CHANNELS', 'OMB_MAX_CHANNELS as usize')}]"
                fields.append((nm, rt))
        out["structs"].append((m.group(3), fields))
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s+(\w+)\s*;", src):
        out["opaque"].append(m.group(2))
    for m in re.finditer(r"typedef\s+(\w[\w\s\*]*?)\(\s*\*\s*(\w+)\s*\)\s*\((.*?)\)\s*;", src, flags=re.S):
        params = [parse_param(p) for p in split_params(m.group(3))]
        out["callbacks"].append((m.group(2), m.group(1).strip(), [p for p in params if p]))
    body = re.sub(r"typedef\s+.*?;", " ", src, flags=re.S)  # prototypes only
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(omb_\w+)\s*\(([^;{}]*?)\)\s*;", body, flags=re.S):
        ret = m.group(1).strip()
        if ret.startswith("extern") or "typedef" in ret:
            continue
        params = [parse_param(p) for p in split_params(m.group(3))]
        out["functions"].append((m.group(2), ret, [p for p in params if p]))
    return out


def split_params(s: str):
    return [p for p in (x.strip() for x in " ".join(s.split()).split(",")) if p]


def render(h) -> str:
    L = ["// GENERATED by tools/gen_rust_sys.py from include/omb200.h — do not edit by hand.",
         "//! Raw FFI declarations of libomb200 (the sm_100a CUDA implementation of OpenMeters' DSP hot path).",
         "//! The safe drop-in processor types live in the `omb200` crate.",
         "#![allow(non_camel_case_types, non_upper_case_globals, clippy::too_many_arguments)]",
         "use std::os::raw::{c_char, c_void};", ""]
    for k, v in h["defines"]:
        L.append(f"pub const {k}: u32 = {v};")
    L.append("")
    for name, vals in h["enums"]:
        if name:
            L.append(f"// enum {name}")
        for k, v in vals:
            L.append(f"pub const {k}: i32 = {v};")
        L.append("")
    for name in h["opaque"]:
        if name in [s[0] for s in h["structs"]]:
            continue
        L += ["#[repr(C)]", f"pub struct {name} {{ _private: [u8; 0] }}"]
    L.append("")
    for name, fields in h["structs"]:
        L += ["#[repr(C)]", "#[derive(Clone, Copy, Debug)]", f"pub struct {name} {{"]
        for fn_, ft in fields:
            L.append(f"    pub {fn_}: {ft},")
        L += ["}", ""]
    for name, ret, params in h["callbacks"]:
        ps = ", ".join(f"{n}: {t}" for n, t in params)
        r = "" if ret == "void" else f" -> {rust_type(ret, None)}"
        L.append(f"pub type {name} = Option<unsafe extern \"C\" fn({ps}){r}>;")
    L += ["", "extern \"C\" {"]
    for name, ret, params in h["functions"]:
        ps = ", ".join(f"{n}: {t}" for n, t in params)
        r = "" if ret == "void" else f" -> {rust_type(ret, None)}"
        L.append(f"    pub fn {name}({ps}){r};")
    L += ["}", ""]
    return "\n".join(L)


def parse_rust_externs(path: str = OUT):
    """name -> number of parameters, from the crate's extern block."""
    src = open(path).read()
    out = {}
    for m in re.finditer(r"pub fn (omb_\w+)\((.*?)\)(?:\s*->\s*[^;]+)?;", src, flags=re.S):
        ps = [p for p in m.group(2).split(",") if p.strip()]
        out[m.group(1)] = len(ps)
    return out


def main():
    text = render(parse_header())
    if "--check" in sys.argv:
        cur = open(OUT).read() if os.path.exists(OUT) else ""
        if cur != text:
            print("rust/omb200-sys/src/lib.rs is stale: run python tools/gen_rust_sys.py")
            return 1
        return 0
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    open(OUT, "w").write(text)
    print(OUT, len(text.splitlines()), "lines")
    return 0


if __name__ == "__main__":
    sys.exit(main())
