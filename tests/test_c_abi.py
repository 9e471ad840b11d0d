"""The drop-in boundary without a GPU: libsassy_b200.so loads, exports every function that
include/sassy.h and include/sassy_gpu.h declare, the record layouts are the reference's
(c/sassy.h:11-21: 40 bytes, cost @32, strand @36), the headers are valid C, and constructing a
searcher on a box without a CUDA device fails loudly instead of falling back to a CPU path."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def declared_functions():
    names = []
    for h in ("sassy.h", "sassy_gpu.h"):
        src = open(os.path.join(INC, h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        src = re.sub(r"^\s*#[^\n]*", "", src, flags=re.M)          # preprocessor lines
        src = re.sub(r'extern\s+"C"\s*\{', "", src)
        src = re.sub(r"typedef\s+struct[^;{]*\{[^}]*\}[^;]*;", "", src, flags=re.S)  # record definitions
        for decl in src.split(";"):
            if "typedef" in decl or "{" in decl or "(" not in decl:
                continue
            m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*\(", decl)
            if m:
                names.append(m.group(1))
    return names


def test_every_declared_symbol_is_exported_and_bound():
    from sassy_b200 import _native
    lib = _native.load()
    names = declared_functions()
    assert {"sassy_searcher", "sassy_searcher_free", "search", "sassy_matches_free"} <= set(names)  # c/sassy.h:38-63
    assert len(names) >= 40
    for n in names:
        assert getattr(lib, n) is not None, n
        assert n in _native.SIGNATURES, f"{n} is declared in include/ but not bound in _native.SIGNATURES"
    assert set(_native.SIGNATURES) <= set(names)


def test_record_layouts(tmp_path):
    from sassy_b200 import _native
    assert ctypes.sizeof(_native.CMatch) == 40
    assert _native.CMatch.cost.offset == 32 and _native.CMatch.strand.offset == 36
    assert ctypes.sizeof(_native.GpuMatch) == 72
    # the headers compile as C and agree with the ctypes mirrors
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sassy_gpu.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(sassy_Match), offsetof(sassy_Match, cost),'
                   'offsetof(sassy_Match, strand), sizeof(sassy_gpu_Match), offsetof(sassy_gpu_Match, ops_off),'
                   'sizeof(sassy_gpu_Stats)); return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", INC, str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert out[:5] == ["40", "32", "36", "72", "64"]
    assert int(out[5]) == ctypes.sizeof(_native.GpuStats)


def test_no_cpu_fallback_without_a_device():
    from sassy_b200 import _native
    lib = _native.load()
    if lib.sassy_gpu_device_count() > 0:
        return  # on a GPU box the parity tests cover construction
    h = lib.sassy_gpu_searcher(b"dna", True, float("nan"), 0)
    assert not h
    assert b"CUDA" in lib.sassy_gpu_last_error() or b"device" in lib.sassy_gpu_last_error()
    import sassy_b200
    try:
        sassy_b200.Searcher("dna")
    except RuntimeError:
        pass
    else:
        raise AssertionError("Searcher must not construct without a GPU")


def build_c_caller(tmp_path):
    exe = tmp_path / "abi_caller"
    libdir = os.path.join(ROOT, "sassy_b200", "lib")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", INC, os.path.join(ROOT, "tests", "c", "abi_caller.c"),
                           "-L", libdir, "-lsassy_b200", "-Wl,-rpath," + libdir, "-lm", "-o", str(exe)])
    return str(exe)


def test_c_caller_compiles_and_links(tmp_path):
    """A C program written against include/*.h links against libsassy_b200.so (like the reference's
    c/example.c against libsassy)."""
    from sassy_b200 import _native
    _native.load()  # the library is built
    assert os.path.exists(build_c_caller(tmp_path))
