"""CPU check of the packed two-lines-per-register deblocking filters (csrc/cuda/swar.cuh).

swar.cuh compiles as plain C++ with bit-accurate emulations of the sm_100a packed-integer
instructions; tests/native/swar_check.cc compares every packed filter against a scalar
restatement of core/frame.c:302-470 on random and near-threshold lines.
"""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_packed_filters_equal_scalar_filters(tmp_path):
    exe = tmp_path / "swar_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "tests/native/swar_check.cc")], check=True)
    out = subprocess.run([str(exe), "1500000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "0 mismatches" in out.stdout
