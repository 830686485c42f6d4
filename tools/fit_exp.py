"""Coefficients of the degree-10 polynomial used by exp_scaled (celeste.jl_b200/csrc/elbo_math.cuh).

Chebyshev-node interpolant of 2^f on [-1/2, 1/2] (near-minimax), solved in 50-digit arithmetic;
prints the coefficients and the maximum relative error of the double-precision evaluation."""
import mpmath as mp
import numpy as np

mp.mp.dps = 50
a = 0.5 * 1.0001
deg = 10
n = deg + 1
xs = [mp.mpf(a) * mp.cos(mp.pi * (2 * i + 1) / (2 * n)) for i in range(n)]
A = mp.matrix(n, n)
b = mp.matrix(n, 1)
for i, x in enumerate(xs):
    for j in range(n):
        A[i, j] = x ** j
    b[i] = mp.mpf(2) ** x
c = [float(v) for v in mp.lu_solve(A, b)]
print(", ".join("%.17g" % v for v in c))
grid = np.linspace(-a / 1.0001, a / 1.0001, 4001)
p = np.full_like(grid, c[-1])
for k in range(deg - 1, -1, -1):
    p = p * grid + c[k]
ref = np.array([float(mp.mpf(2) ** mp.mpf(float(x))) for x in grid])
print("max relative error:", np.abs(p - ref).max() / 1.0)
