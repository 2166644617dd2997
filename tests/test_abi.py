"""CPU: the C-ABI library loads without a GPU, exports every symbol the headers declare, and its
public structs have the reference's binary layout (golden from the reference's p264.h)."""
import ctypes as C
import json
import re
import subprocess
from pathlib import Path

import p264decoder_b200 as P

ROOT = Path(__file__).resolve().parents[1]
INC = ROOT / "include"


def declared_functions():
    names = set()
    for h in INC.glob("*.h"):
        txt = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        for m in re.finditer(r"\b(p264b?2?0?0?_[a-z0-9_]+)\s*\(", txt):
            n = m.group(1)
            # skip function-pointer typedef names and struct members
            if re.search(r"\(\s*\*\s*" + re.escape(n), txt) or n.endswith("_t"):
                continue
            names.add(n)
    return names


def test_library_exports_every_declared_symbol():
    lib = P.load_library()
    names = declared_functions()
    assert len(names) > 40
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_no_cuda_runtime_dependency_leaks():
    out = subprocess.check_output(["ldd", str(P.LIB_PATH)]).decode()
    assert "libcudart" not in out and "libtorch" not in out  # static cudart, no torch types at the boundary


def test_struct_sizes():
    lib = P.load_library()
    assert lib.p264b200_abi_version() == 2
    assert P.MB_DTYPE.itemsize == 96
    assert C.sizeof(P.FrameHdr) == 124
    assert C.sizeof(P.FrameSyntax) == 144


def test_public_structs_match_reference_layout(tmp_path):
    golden = json.loads((ROOT / "tests" / "golden" / "p264_abi_layout.json").read_text())
    probe = (ROOT / "tests" / "golden" / "abi_probe.c.in").read_text().replace("HEADER", f'"{INC / "p264_b200.h"}"')
    c = tmp_path / "probe.c"
    c.write_text(probe)
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-w", "-o", str(exe), str(c)])
    ours = json.loads(subprocess.check_output([str(exe)]).decode())
    assert ours == golden


def test_engine_fails_loudly_without_gpu():
    lib = P.load_library()
    if lib.p264b200_device_count() > 0:
        return
    cfg = P.EngineCfg(0, 1, 4, 4, 2, 0, 1, 0)
    e = C.c_void_p()
    assert lib.p264b200_engine_create(C.byref(e), C.byref(cfg)) == -2  # P264B200_ENODEV
    assert b"no CPU fallback" in lib.p264b200_last_error()
    lib.p264_decoder_open.restype = C.c_void_p
    assert not lib.p264_decoder_open(None)
    assert lib.p264b200_tables_ready() == -2
