#!/usr/bin/env python3
"""Regenerates the `extern "C"` block of rust/qpzk-sys/src/lib.rs from include/qpzk.h, so that the FFI crate a
patched qp-plonky2 links cannot drift from the header. tests/test_abi.py checks the result.

    python scripts/gen_rust_decls.py            # rewrites rust/qpzk-sys/src/lib.rs in place
"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TYPES = {"int": "c_int", "void": "c_void", "char": "c_char", "uint8_t": "u8", "uint32_t": "u32", "uint64_t": "u64",
         "size_t": "usize", "float": "f32", "double": "f64"}
OPAQUE = ("qpzk_ctx", "qpzk_batch", "qpzk_tree", "qpzk_circuit", "qpzk_fri")
RESERVED = {"out": "out_", "in": "in_", "type": "type_", "ref": "ref_"}


def rust_type(c):
    c = c.strip()
    const = "const" in c.split()
    base = [t for t in c.replace("*", " * ").split() if t not in ("const", "struct")]
    stars = base.count("*")
    name = [t for t in base if t != "*"][0]
    t = TYPES.get(name, name)
    for i in range(stars):
        # `const T* const*` keeps constness on the innermost pointer only; good enough for this header
        t = ("*const " if const and i == 0 else "*mut ") + t
    return t


def parse(header):
    src = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    src = re.sub(r"#.*", "", src)
    decls = []
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(qpzk_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef"):
            continue
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)([A-Za-z_]\w*)(\s*\[[^\]]*\])?$", a)
                ctype, pname, arr = mm.group(1), mm.group(2), mm.group(3)
                if arr:
                    ctype += "*"
                params.append((RESERVED.get(pname, pname), rust_type(ctype)))
        decls.append((name, params, None if ret == "void" else rust_type(ret)))
    return decls


def main():
    header = open(os.path.join(ROOT, "include", "qpzk.h")).read()
    lines = []
    for name, params, ret in parse(header):
        sig = ", ".join("%s: %s" % p for p in params)
        lines.append("    pub fn %s(%s)%s;" % (name, sig, " -> %s" % ret if ret else ""))
    path = os.path.join(ROOT, "rust", "qpzk-sys", "src", "lib.rs")
    rs = open(path).read()
    head = rs[:rs.index('extern "C" {')]
    open(path, "w").write(head + 'extern "C" {\n' + "\n".join(lines) + "\n}\n")
    print("wrote %d declarations" % len(lines))


if __name__ == "__main__":
    main()
