# coding: utf-8
""" K4 divides its running sums by compile-time row counts with a 3-instruction sequence instead of the
generic IEEE division (`div_small`, csrc/reduce.cuh).  tools/divcheck.c compares the sequence with
`a / m` by enumeration; here: every 4099th float plus the operands around the guard's boundaries and
around every power of two, for all divisors 1..64 (the exhaustive run — all 2^32 operands per divisor,
0 failures — is profiles/r02_divcheck_exhaustive.txt). """

import pathlib
import shutil
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent

@pytest.mark.skipif(shutil.which("gcc") is None, reason="no C compiler")
def test_division_sequence_is_exact_inside_its_guard(tmp_path):
  exe = tmp_path / "divcheck"
  build = subprocess.run(["gcc", "-O2", "-mfma", "-fopenmp", "-ffp-contract=off", str(ROOT / "tools" / "divcheck.c"), "-o", str(exe), "-lm"],
                         capture_output=True, text=True)
  if build.returncode != 0:        # no FMA instruction set / OpenMP on this host: libm's fmaf is exact too
    build = subprocess.run(["gcc", "-O2", "-ffp-contract=off", str(ROOT / "tools" / "divcheck.c"), "-o", str(exe), "-lm"], capture_output=True, text=True)
  assert build.returncode == 0, build.stderr
  run = subprocess.run([str(exe), "1", "64", "4099"], capture_output=True, text=True, timeout=600)
  assert run.returncode == 0, run.stdout[-2000:]
  lines = run.stdout.strip().splitlines()
  assert lines[-1] == "failures 0" and len(lines) == 65
  # the guard is not decorative: outside its range the sequence does go wrong (inf -> NaN, -0 -> +0, tiny operands)
  assert all(int(line.split()[-1]) > 0 for line in lines[:-1])

def test_kernel_source_uses_the_checked_range():
  text = (ROOT / "byzantinemomentum_b200" / "csrc" / "reduce.cuh").read_text()
  assert "7.888609052210118e-31f" in text and "3.402823466e+38f" in text          # 2^-100 and FLT_MAX, as in divcheck.c
  check = (ROOT / "tools" / "divcheck.c").read_text()
  assert "0x1p-100f" in check and "3.402823466e+38f" in check
