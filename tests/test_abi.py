"""The C-ABI library loads without a GPU and exports exactly what include/splice_b200.h declares; the ctypes
mirrors of the argument structs have the C compiler's layout."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "splice_b200.h"


def declared_symbols():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"SPLICE_API\s+[\w\s\*]+?\b(splice_\w+)\s*\(", text)))


def test_library_loads_and_exports_header_symbols():
    from splice_b200 import _lib

    syms = declared_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.EXPORTS) == syms, set(syms) ^ set(_lib.EXPORTS)
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in the header but not exported by the .so"
    assert _lib.splice_version() == 100
    assert _lib.splice_last_error() == b""


def test_nm_shows_only_splice_symbols():
    from splice_b200 import _lib

    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert sorted(exported) == declared_symbols()


def test_struct_layouts_match_the_c_compiler(tmp_path):
    from splice_b200 import _lib

    structs = ["SpliceGemmArgs", "SpliceVitDesc", "SpliceImage", "SpliceVitForwardArgs", "SpliceVitBackwardArgs",
               "SpliceProfileEntry"]
    src = tmp_path / "sizes.c"
    body = "".join(f'  printf("{s} %zu\\n", sizeof({s}));\n' for s in structs)
    body += '  printf("off_fwd_pre_normalized %zu\\n", offsetof(SpliceVitForwardArgs, pre_normalized));\n'
    body += '  printf("off_gemm_bn_hint %zu\\n", offsetof(SpliceGemmArgs, bn_hint));\n'
    src.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "splice_b200.h"\nint main(void) {{\n{body}  return 0;\n}}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for s in structs:
        assert int(got[s]) == C.sizeof(getattr(_lib, s)), s
    assert int(got["off_fwd_pre_normalized"]) == _lib.SpliceVitForwardArgs.pre_normalized.offset
    assert int(got["off_gemm_bn_hint"]) == _lib.SpliceGemmArgs.bn_hint.offset


def test_argument_errors_are_reported_not_crashed():
    """No compute without a GPU: only argument validation paths are exercised here."""
    from splice_b200 import _lib

    a = _lib.SpliceGemmArgs()
    a.M, a.N, a.K = 128, 100, 64  # N not a multiple of 32
    rc = _lib.splice_gemm_bf16(a, None)
    assert rc < 0 and b"multiple of 32" in _lib.splice_last_error()
    rc = _lib.splice_vit_create(None, None, None, 0, None)
    assert rc < 0
    d = _lib.SpliceVitDesc(8, 768, 12, 12, 785, 1e-6)
    assert _lib.splice_vit_packed_floats(C.byref(d)) == 85_807_872 - 0  # all 150 DINO ViT-B/8 tensors
    d = _lib.SpliceVitDesc(16, 384, 6, 12, 197, 1e-6)
    assert _lib.splice_vit_packed_floats(C.byref(d)) == 21_665_664
