#!/usr/bin/env python3
"""Generates the polynomial coefficients used by the kernels' exp() (scdd_b200/csrc/device_math.cuh).

exp(a) = 2^k * exp(r), r = a - k*ln2 in [-ln2/2, ln2/2];  exp(r) ~= 1 + r + r^2 * P(r),
P of degree DEG-2 obtained by Chebyshev interpolation of (exp(r)-1-r)/r^2 in 60-digit arithmetic
(near-minimax), coefficients rounded to double.  Prints the coefficients and the measured
max relative error of the double-precision Horner evaluation against mpmath.
"""
import sys
import numpy as np
import mpmath as mp

mp.mp.dps = 60
DEG = int(sys.argv[1]) if len(sys.argv) > 1 else 11
half = mp.log(2) / 2 * mp.mpf("1.0001")
m = DEG - 2 + 1  # number of coefficients of P


def f(r):
    if abs(r) < mp.mpf("1e-25"):
        return mp.mpf(1) / 2
    return (mp.e ** r - 1 - r) / (r * r)


nodes = [half * mp.cos(mp.pi * (2 * i + 1) / (2 * m)) for i in range(m)]
A = mp.matrix(m, m)
b = mp.matrix(m, 1)
for i, x in enumerate(nodes):
    for j in range(m):
        A[i, j] = x ** j
    b[i] = f(x)
c = mp.lu_solve(A, b)
coef = [float(c[j]) for j in range(m)]  # P(r) = sum coef[j] r^j
print("// degree", DEG, "coefficients of P, low order first")
for j, v in enumerate(coef):
    print(f"    {v!r},  // r^{j + 2}   ({float.hex(v)})")

# error of the double evaluation
rs = np.linspace(-float(mp.log(2) / 2), float(mp.log(2) / 2), 20001)
worst = 0.0
for r in rs:
    p = coef[-1]
    for v in coef[-2::-1]:
        p = p * r + v          # not fused here; the device fuses (smaller error)
    val = 1.0 + (r + r * r * p)
    ref = mp.e ** mp.mpf(float(r))
    err = abs((mp.mpf(val) - ref) / ref)
    worst = max(worst, float(err))
print("// max rel err (double Horner, unfused):", worst, "=", worst / 2 ** -53, "ulp(0.5)")
