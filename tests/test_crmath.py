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


def test_first_phase_evaluations_equal_the_double_double_path(tmp_path):
    """crmath.h with -DCRM_FAST (what the GPU build compiles): whenever the first-phase sin / cos / atan2 accepts, its
    result is the double-double path's bit for bit -- 4.8 M sin/cos and 3.6 M atan2 arguments (planner-sized angles, wide
    ranges, next to multiples of pi/2 and to the table points); almost all calls are accepted."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "crmath_fastcheck")
    cc = shutil.which("gcc")
    assert cc, "gcc is part of the image"
    subprocess.check_call([cc, "-O2", "-ffp-contract=off", "-mfma", "-DCRM_FAST", os.path.join(root, "oracle", "crmath_fastcheck.c"),
                           "-lm", "-o", exe])
    out = subprocess.run([exe, "600000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rows = {ln.split()[0]: [int(v) for v in ln.split()[1:]] for ln in out.stdout.splitlines()}
    for name, (n, accepted, bad) in rows.items():
        assert bad == 0 and n > 1_000_000 and accepted > 0.95 * n, (name, n, accepted, bad)
