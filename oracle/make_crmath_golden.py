"""TEST INFRASTRUCTURE ONLY -- correctly rounded reference values for csrc/crmath.h, computed with
mpmath at 300 bits and rounded once to binary64.  Writes tests/golden/crmath.npz.

    python oracle/make_crmath_golden.py
"""
import math
import os
import random

import mpmath
import numpy as np

mpmath.mp.prec = 300
HERE = os.path.dirname(os.path.abspath(__file__))


def rn(v):
    return float(v)


def main():
    rnd = random.Random(20261018)
    ys, xs = [], []
    for i in range(12000):
        sy = 10 ** rnd.uniform(-7, 2)
        sx = 10 ** rnd.uniform(-7, 2)
        ys.append(rnd.uniform(-1, 1) * sy)
        xs.append(rnd.uniform(-1, 1) * sx)
    # structured cases of the steer: multiples of the resolution, axis-aligned, tiny components
    for a in (0.1, 0.3, 0.5, 1.0, 2.0, 3.0):
        for b in (0.0, 0.1, 0.4, 1.0, -0.1, -1.0, 1e-17, -1e-17, 1e-300):
            for s1 in (1, -1):
                ys.append(s1 * b); xs.append(a)
                ys.append(a); xs.append(s1 * b)
                ys.append(s1 * b); xs.append(-a)
                ys.append(-a); xs.append(s1 * b)
    ys = np.array(ys); xs = np.array(xs)
    keep = ~((ys == 0) & (xs == 0))
    ys, xs = ys[keep], xs[keep]
    at, st, ct, hy = [], [], [], []
    for y, x in zip(ys, xs):
        t = rn(mpmath.atan2(mpmath.mpf(float(y)), mpmath.mpf(float(x)))) if y != 0 else math.atan2(y, x)
        at.append(t)
        st.append(rn(mpmath.sin(mpmath.mpf(t))))
        ct.append(rn(mpmath.cos(mpmath.mpf(t))))
        hy.append(math.hypot(y, x))  # CPython's hypot is the reference's (rrt_04:1235)
    args = [rnd.uniform(-7, 7) for _ in range(6000)] + [rnd.uniform(-1e4, 1e4) for _ in range(2000)]
    args += [math.pi, -math.pi, math.pi / 2, -math.pi / 2, 3 * math.pi / 2, 1e-10, -1e-300, 0.7853981633974483,
             0.785398163397448, 2.356194490192345, 1e5 * math.pi, 0.0]
    args = np.array(args)
    sn = [rn(mpmath.sin(mpmath.mpf(float(a)))) for a in args]
    cs = [rn(mpmath.cos(mpmath.mpf(float(a)))) for a in args]
    np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "crmath.npz"),
                        y=ys, x=xs, atan2=np.array(at), sin_t=np.array(st), cos_t=np.array(ct),
                        hypot=np.array(hy), arg=args, sin=np.array(sn), cos=np.array(cs))
    print(len(ys), "atan2 cases,", len(args), "sin/cos cases")


if __name__ == "__main__":
    main()
