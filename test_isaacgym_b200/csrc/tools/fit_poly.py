#!/usr/bin/env python3
"""Coefficients of the reduced-cost fp64 elementary functions of the reference-precision servo step
(servo_math.cuh: atan2_f32grade, sincos_quarter).

    python fit_poly.py atan 14      atan(t)  = t * P(t^2),  t in [0, 1]
    python fit_poly.py sin 8        sin(a)   = a * P(a^2),  |a| <= pi/2 (+1e-6)
    python fit_poly.py cos 8        cos(a)   =     P(a^2),  |a| <= pi/2 (+1e-6)

P is the Chebyshev interpolant of the even part on the interval of u = x^2, converted to the monomial basis and
evaluated by Horner in fp64.  The results only have to survive rounding to fp32 (torch.atan2 on fp32 tensors,
test10_servo_vecenv.py:407; the quaternion assignment into the fp32 state, :453): an fp64 error of 3e-13 changes the
rounded value for a few inputs in 10^6, far below torch's own Sleef atan2f (1 ulp, not correctly rounded).
Prints the coefficients and the max error of the fp64 Horner evaluation against long-double references.
"""
import sys

import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P

fn = sys.argv[1]
deg = int(sys.argv[2])
LD = np.longdouble
umax = 1.0 if fn == "atan" else (np.pi / 2 + 1e-6) ** 2


def even_part(u):
    u = np.asarray(u, dtype=LD)
    t = np.sqrt(u)
    if fn == "cos":
        return np.cos(t)
    g = np.arctan if fn == "atan" else np.sin
    out = np.ones_like(u)
    nz = t > 1e-5
    out[nz] = g(t[nz]) / t[nz]
    out[~nz] = 1 - u[~nz] / (3 if fn == "atan" else 6)
    return out


k = np.arange(deg + 1)
x = np.cos(np.pi * (k + 0.5) / (deg + 1))
u = (x + 1) / 2 * umax
V = C.chebvander(x, deg)
cheb = np.linalg.solve(V, np.asarray(even_part(u)).astype(np.float64))
mono_x = C.cheb2poly(cheb)                      # in x = 2u/umax - 1
coef = np.zeros(1)
for c in mono_x[::-1]:
    coef = P.polyadd(P.polymul(coef, np.array([-1.0, 2.0 / umax])), [c])

rng = np.random.default_rng(0)
t = np.concatenate([rng.uniform(0, np.sqrt(umax), 2_000_000), np.linspace(0, np.sqrt(umax), 200_001)])
uu = t * t
acc = np.zeros_like(uu)
for c in coef[::-1]:
    acc = acc * uu + c
approx = acc if fn == "cos" else t * acc
ref = {"atan": np.arctan, "sin": np.sin, "cos": np.cos}[fn](t.astype(LD))
err = np.abs(approx - ref) / (1.0 if fn == "cos" else np.maximum(np.abs(ref), 1e-300))
err[t == 0] = 0
print(f"// {fn}: degree {deg} in u = x^2, max {'abs' if fn == 'cos' else 'rel'} err of the fp64 Horner evaluation "
      f"= {float(err.max()):.2e} (fit_poly.py {fn} {deg})")
print(f"static constexpr double k{fn.capitalize()}P[{deg + 1}] = {{")
for c in coef:
    print(f"    {float(c):+.17e},")
print("};")
