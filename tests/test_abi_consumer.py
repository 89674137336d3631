"""The C ABI seen from a compiled non-Python consumer (tests/abi_consumer.c, plain C11 built by gcc in build()).

There is no Rust toolchain here, so this is the available stand-in for compiling bindings/rust/pasture_b200_sys.rs: the
struct layouts printed by the C program must equal the ctypes mirror's (pasture_b200/_lib.py) and the field order of the
#[repr(C)] structs in the Rust declarations; on a GPU box the program runs one conversion through the header alone."""
import ctypes as C
import os
import re
import subprocess

import pytest

from pasture_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "build", "abi_consumer")


def build_consumer():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi_consumer.c"), "-o", EXE, "-L" + os.path.join(ROOT, "pasture_b200"),
                           "-lpasture_b200", "-Wl,-rpath,$ORIGIN/../../pasture_b200"])


@pytest.fixture(scope="module")
def exe():
    src = os.path.join(ROOT, "tests", "abi_consumer.c")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(_lib.SO_PATH)):
        build_consumer()
    return EXE


CTYPES = {"pb200_attr": _lib.Attr, "pb200_buffer_desc": _lib.BufferDesc, "pb200_transform": _lib.Transform,
          "pb200_proj_op": _lib.ProjOp, "pb200_las_header": _lib.LasHeader, "pb200_las_write_stats": _lib.LasWriteStats,
          "pb200_voxel_partials_desc": _lib.VoxelPartialsDesc}


def test_struct_layouts_equal_the_ctypes_mirror(exe):
    out = subprocess.run([exe, "layout"], capture_output=True, text=True, check=True).stdout
    seen = 0
    for line in out.splitlines():
        f = line.split()
        if f[0] == "sizeof":
            assert C.sizeof(CTYPES[f[1]]) == int(f[2]), line
            seen += 1
        elif f[0] == "offsetof":
            assert getattr(CTYPES[f[1]], f[2]).offset == int(f[3]), line
            seen += 1
        elif f[0] == "abi_version":
            assert int(f[1]) == 1
    assert seen >= 40


def test_rust_declarations_list_the_same_fields_in_the_same_order():
    """#[repr(C)] makes Rust lay a struct out like C if the fields agree in order and type width; the declarations cannot
    be compiled here, so compare their field lists with the header's textually"""
    header = open(os.path.join(ROOT, "include", "pasture_b200.h")).read()
    rust = open(os.path.join(ROOT, "bindings", "rust", "pasture_b200_sys.rs")).read()
    width = {"u8": 1, "i8": 1, "u16": 2, "i16": 2, "u32": 4, "i32": 4, "c_int": 4, "u64": 8, "i64": 8, "f64": 8, "usize": 8}
    cwidth = {"char": 1, "uint8_t": 1, "uint16_t": 2, "uint32_t": 4, "int32_t": 4, "uint64_t": 8, "double": 8}
    for name in ("pb200_attr", "pb200_buffer_desc", "pb200_transform", "pb200_proj_op"):
        m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header, re.S)
        body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
        cf = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ty, rest = decl.rsplit(" ", 1) if "," not in decl else (decl.split(" ", 1)[0], decl.split(" ", 1)[1])
            for fld in rest.split(","):
                fld = fld.strip()
                ptr = "*" in fld or "*" in ty
                arr = re.search(r"\[(\w+)\]", fld)
                fname = re.sub(r"[\*\[\]\w]*$", "", "") or re.match(r"\**(\w+)", fld).group(1)
                base = ty.replace("const", "").replace("*", "").strip()
                w = 8 if ptr else cwidth[base]
                cf.append((fname, w))
        r = re.search(r"pub struct %s \{(.*?)\}" % name, rust, re.S)
        assert r, f"{name} missing from pasture_b200_sys.rs"
        rf = []
        rbody = re.sub(r"//[^\n]*", "", r.group(1))
        for mm in re.finditer(r"pub (\w+):\s*([^,\n}]+)", rbody):
            ty = mm.group(2).strip()
            if ty.startswith("*"):
                w = 8
            else:
                a = re.match(r"\[(\w+); .+\]", ty)
                base = a.group(1) if a else ty
                w = width.get(base, 1)  # c_char = 1
            rf.append((mm.group(1), w))
        assert [n for n, _ in cf] == [n for n, _ in rf], (name, cf, rf)
        assert [w for _, w in cf] == [w for _, w in rf], (name, cf, rf)


def test_consumer_plans_a_conversion_without_a_device(exe):
    """the C program builds the LAS read-path converter with ctx == NULL and prints its tile schedule: host logic only"""
    out = subprocess.run([exe, "plan"], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0].startswith("tiles tile_points=2048 threads=512") and "ops=12" in out[0]
    items = [l for l in out[1:] if l.startswith("item ")]
    assert len(items) == int(re.search(r"items=(\d+)", out[0]).group(1)) and 16 <= len(items) <= 28


def test_consumer_fails_loudly_without_a_device(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = subprocess.run([exe, "convert"], capture_output=True, text=True)
    assert p.returncode == 77 and "no CPU fallback" in p.stderr  # PB200_ERR_NO_DEVICE, never a silent host path


@pytest.mark.gpu
def test_consumer_converts_the_fixture_points_through_the_header_alone(exe):
    p = subprocess.run([exe, "convert"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "convert: ok" in p.stdout
