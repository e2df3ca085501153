"""The C-ABI shared library loads and exports every symbol include/misaki_b200.h declares (no compute calls:
this suite runs without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

from misaki_render_b200 import capi

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "misaki_b200.h").read_text()


def test_header_symbols_are_exported():
    declared = sorted(set(re.findall(r"\b(msk_gpu_[a-z_]+)\s*\(", HEADER)))
    assert declared, "no declarations found"
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(capi.EXPORTED_SYMBOLS) == declared
    assert lib.msk_gpu_abi_version() == int(re.search(r"#define MSK_ABI_VERSION (\d+)", HEADER).group(1))


def test_struct_layouts_match_header(tmp_path):
    """Compile the header with gcc and compare sizeof/offsetof with the ctypes mirror."""
    import subprocess
    structs = {"MskSpectrum": capi.MskSpectrum, "MskBsdf": capi.MskBsdf, "MskEmitter": capi.MskEmitter, "MskMesh": capi.MskMesh,
               "MskCamera": capi.MskCamera, "MskSceneDesc": capi.MskSceneDesc, "MskRenderDesc": capi.MskRenderDesc,
               "MskStats": capi.MskStats, "MskAccelInfo": capi.MskAccelInfo, "MskMedium": capi.MskMedium}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include" / "misaki_b200.h"}"', "int main(void){"]
    for name, st in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{name}.{fname} %zu\\n", offsetof({name}, {fname}));')
    lines += ['printf("MskRay %zu\\n", sizeof(MskRay));', 'printf("MskHit %zu\\n", sizeof(MskHit));', "return 0;}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, st in structs.items():
        assert int(out[name]) == ctypes.sizeof(st), name
        for fname, _ in st._fields_:
            assert int(out[f"{name}.{fname}"]) == getattr(st, fname).offset, f"{name}.{fname}"
    assert int(out["MskRay"]) == capi.RAY_DTYPE.itemsize == 32 and int(out["MskHit"]) == capi.HIT_DTYPE.itemsize == 20


def test_no_cpu_fallback_without_device():
    """Without a usable GPU the product refuses to run (it must never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(capi.MskError) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(capi.MskError):
        capi.load(tmp_path / "libmisaki_b200.so")


def test_product_never_imports_the_oracle():
    """A product path that routes through oracle/ would void every parity claim."""
    pat = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|pyoracle|oracle/|orc_[a-z_]+\s*\()")
    for p in (ROOT / "misaki_render_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp", ".hpp") and "lib" not in p.parts:
            m = pat.search(p.read_text())
            assert m is None, f"{p} references the oracle: {m.group(0)}"
