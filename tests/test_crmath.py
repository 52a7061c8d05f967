"""csrc/crmath.h (compiled for the CPU inside the oracle library) against correctly rounded values
from mpmath (tests/golden/crmath.npz, made by oracle/make_crmath_golden.py)."""
import os

import numpy as np

from conftest import GOLDEN


def _g():
    return np.load(os.path.join(GOLDEN, "crmath.npz"))


def test_atan2_sincos_correctly_rounded(oracle_lib):
    O = oracle_lib
    g = _g()
    for y, x, t, s, c in zip(g["y"], g["x"], g["atan2"], g["sin_t"], g["cos_t"]):
        th, ss, cc = O.cr_atan2_sincos(float(y), float(x))
        assert th == t and ss == s and cc == c, (y, x)


def test_sin_cos_hypot_correctly_rounded(oracle_lib):
    L = oracle_lib.lib()
    g = _g()
    for a, s, c in zip(g["arg"], g["sin"], g["cos"]):
        assert L.orc_cr_sin(float(a)) == s and L.orc_cr_cos(float(a)) == c, a
    for y, x, h in zip(g["y"], g["x"], g["hypot"]):
        assert L.orc_cr_hypot(float(y), float(x)) == h
        assert L.orc_hypot(float(y), float(x)) == h


def test_special_values(oracle_lib):
    import math
    O = oracle_lib
    assert O.cr_atan2_sincos(0.0, 1.0) == (0.0, 0.0, 1.0)
    assert O.cr_atan2_sincos(0.0, 0.0) == (0.0, 0.0, 1.0)          # math.atan2(0, 0) == 0
    th, s, c = O.cr_atan2_sincos(0.0, -1.0)
    assert th == math.pi and c == -1.0 and s == 1.2246467991473532e-16
    th, s, c = O.cr_atan2_sincos(-0.0, -1.0)
    assert th == -math.pi and c == -1.0
    assert O.lib().orc_cr_hypot(0.0, 0.0) == 0.0
    assert O.lib().orc_cr_hypot(3.0, 4.0) == 5.0
    assert O.lib().orc_cr_hypot(1e-320, 1e-320) == math.hypot(1e-320, 1e-320)
