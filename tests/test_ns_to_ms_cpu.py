"""CPU: the kernels' ns -> ms conversion (Markstein's division step, csrc/tml_engine.cu
ns_to_ms) equals true IEEE division -- i.e. Python's ``ns / 1e6`` -- checked with the
same three FMA-level operations compiled by gcc."""
import os
import shutil
import subprocess
import sys

import pytest

SRC = r"""
#include <math.h>
#include <stdint.h>
#include <stdio.h>
static inline double mk(double a) {
  const double y = 1.0e-6;
  double q = a * y;
  double r = fma(-1.0e6, q, a);
  return fma(r, y, q);
}
static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main(void) {
  uint64_t bad = 0, n = 0;
  for (uint64_t a = 0; a < 20000000ull; ++a, ++n) if (mk((double)a) != (double)a / 1.0e6) ++bad;
  for (uint64_t i = 0; i < 20000000ull; ++i, ++n) {
    uint64_t a = rnd() >> (11 + (int)(rnd() % 42));
    if (mk((double)a) != (double)a / 1.0e6) ++bad;
  }
  for (uint64_t k = 1; k < 2000000ull; ++k) for (int d = -2; d <= 2; ++d, ++n) {
    uint64_t a = k * 1000000ull + (uint64_t)d;
    if (mk((double)a) != (double)a / 1.0e6) ++bad;
  }
  printf("%llu %llu\n", (unsigned long long)n, (unsigned long long)bad);
  return 0;
}
"""


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_markstein_ns_to_ms_equals_true_division(tmp_path):
    src = tmp_path / "mk.c"
    src.write_text(SRC)
    exe = tmp_path / "mk"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(src), "-lm"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) > 45_000_000 and int(out[1]) == 0
