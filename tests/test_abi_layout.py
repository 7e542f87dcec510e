"""The three descriptions of the C ABI agree: include/gears_physics_cuda.h (compiled with gcc here), the ctypes
binding (gears_b200/capi.py) and the Rust -sys crate (rust/gears-physics-cuda-sys/src/lib.rs, parsed as text: no Rust
toolchain exists in this image).  Also: the C++ host mirror compiles against the header without CUDA."""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gears_physics_cuda.h")
RUST_SYS = os.path.join(ROOT, "rust", "gears-physics-cuda-sys", "src", "lib.rs")


def header_text():
    return re.sub(r"/\*.*?\*/|//[^\n]*", "", open(HEADER).read(), flags=re.S)


def header_structs():
    """{struct name: [field names in order]} for every `typedef struct name { ... } name;` of the header."""
    out = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*\1\s*;", header_text(), flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(None, 1)[1] if " " in decl else decl
            for n in names.split(","):
                fields.append(re.sub(r"\[.*?\]|\*", "", n).strip().split()[-1])
        out[m.group(1)] = fields
    return out


def c_layout(tmp_path, structs):
    """sizeof and offsetof of every field, as the C compiler sees the header."""
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for s, fields in structs.items():
        src.append(f'  printf("{s} %zu\\n", sizeof({s}));')
        for f in fields:
            src.append(f'  printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    src += ["  return 0;", "}"]
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(c)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    return {k: int(v) for k, v in (line.split() for line in out.splitlines())}


def test_header_is_plain_c_and_matches_the_ctypes_binding(tmp_path):
    from gears_b200 import capi
    structs = header_structs()
    twins = {"gp_config": capi.GpConfig, "gp_step_stats": capi.GpStepStats, "gp_slab_config": capi.GpSlabConfig,
             "gp_kernel_times": capi.GpKernelTimes}
    assert set(twins) <= set(structs), sorted(structs)
    lay = c_layout(tmp_path, {k: structs[k] for k in twins})
    for name, cls in twins.items():
        assert [f for f, _ in cls._fields_] == structs[name], name
        assert C.sizeof(cls) == lay[name], name
        for f, _ in cls._fields_:
            assert getattr(cls, f).offset == lay[f"{name}.{f}"], f"{name}.{f}"


def rust_items():
    txt = re.sub(r"//[^\n]*", "", open(RUST_SYS).read())
    structs = {}
    for m in re.finditer(r"#\[repr\(C\)\][^{]*?pub struct (\w+)\s*\{(.*?)\}", txt, flags=re.S):
        structs[m.group(1)] = re.findall(r"pub (\w+)\s*:", m.group(2))
    fns = set(re.findall(r"pub fn (gp_\w+)\s*\(", txt))
    consts = dict(re.findall(r"pub const (GP_\w+)\s*:\s*\w+\s*=\s*([^;]+);", txt))
    return structs, fns, consts


def test_rust_sys_crate_mirrors_the_header():
    structs, fns, consts = rust_items()
    h = header_structs()
    for name in ("gp_config", "gp_step_stats", "gp_slab_config", "gp_kernel_times"):
        assert structs.get(name) == h[name], name  # same fields, same order, all repr(C)
    declared = set(re.findall(r"\b(gp_[a-z0-9_]+)\s*\(", header_text()))
    assert declared == fns, sorted(declared ^ fns)
    # flag and status values
    ht = header_text()
    for m in re.finditer(r"#define\s+(GP_(?:FLAG|SLAB)_\w+)\s+(\S+)", ht):
        assert m.group(1) in consts, m.group(1)
        assert int(consts[m.group(1)].strip().rstrip("u"), 0) == int(m.group(2).rstrip("u"), 0), m.group(1)
    for m in re.finditer(r"\b(GP_(?:OK|ERR_\w+))\s*=\s*(-?\d+)", ht):
        assert m.group(1) in consts, m.group(1)
        assert int(consts[m.group(1)]) == int(m.group(2)), m.group(1)


def test_cpp_host_mirror_compiles_without_cuda():
    """gears_b200/host/gears_physics.hpp needs only the C header: a maintainer can build it with any C++17 compiler."""
    src = os.path.join(ROOT, "tests", "host_cpp", "test_update_physics.cpp")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "gears_b200", "host"), src], check=True)
