"""Generates tests/golden/unit_test.json — the known answers the reference's own tests hold for
the hot path (closed forms copied from /root/reference/test/unit-test.jl).

    G1  unit-test.jl:41-48, 179-186   n=10, f=x'x, c=sum(x)-1, sigma=0.5, x = ones/n
    G2  unit-test.jl:54-59, 193-198   same model, x = [0, 1, ..., 1]
    G3  unit-test.jl:100-126          Rosenbrock + circle, (sigma, rho, delta) = (0.5, 0.1, 0.25)
    G4  unit-test.jl:191, 202         Val(2) hprod = 2 v - 2 * 1 * (Ys' v)

Every value is a closed-form expression of the test file (no solver involved), evaluated with
numpy; the dense K-solve cross-check at the bottom guards against transcription errors.
Run:  python tests/golden/make_golden.py
"""
import json
import os

import numpy as np

n, sigma = 10, 0.5
ys = lambda x: ((2 - sigma) / n * x.sum() + sigma / n)
Ys = (2 - sigma) / n * np.ones(n)

xf = np.ones(n) / n
G1 = dict(x=xf.tolist(), obj=0.1, fx=0.1, gx=(0.2 * np.ones(n)).tolist(), ys=[ys(xf)], cx=[0.0],
          grad=np.zeros(n).tolist())
xr = np.r_[0.0, np.ones(9)]
cx = np.array([8.0])
G2 = dict(x=xr.tolist(), cx=cx.tolist(), ys=[ys(xr)], obj=float(9.0 - cx @ np.array([ys(xr)])),
          grad=(2 * xr - Ys * cx - np.ones(n) * ys(xr)).tolist())
rng = np.random.default_rng(1234)
v = rng.random(n)
G4 = dict(v=v.tolist(), hprod_val2=(2 * v - 2 * np.ones(n) * (Ys @ v)).tolist())

s, r, d = 0.5, 0.1, 0.25
x = np.array([np.sqrt(6) / 3, np.sqrt(3) / 3])
D = -(4 * x[0] ** 2 + 4 * x[1] ** 2 + d)
ys3 = (2 * x[0] * (-2 * (x[0] - 1) + 400 * x[0] * (x[1] - x[0] ** 2)) - 400 * x[1] * (x[1] - x[0] ** 2)
       + s * (x[0] ** 2 + x[1] ** 2 - 1)) / D
fx3 = (np.sqrt(6) - 3) ** 2 / 9 + 100 * (np.sqrt(3) - 2) ** 2 / 9
gx3 = np.array([2 * (np.sqrt(6) / 3 - 1) - 400 * np.sqrt(6) / 3 * (np.sqrt(3) / 3 - 6 / 9),
                200 * (np.sqrt(3) / 3 - 6 / 9)])
G3 = dict(x=x.tolist(), sigma=s, rho=r, delta=d, obj=fx3, fx=fx3, gx=gx3.tolist(), ys=[ys3], cx=[0.0])

# cross-check the closed forms against a dense solve of K [p; q] = rhs
A = np.ones((1, n))
K = np.block([[np.eye(n), A.T], [A, np.zeros((1, 1))]])
for xx, G in ((xf, G1), (xr, G2)):
    g = 2 * xx; c = np.array([xx.sum() - 1.0])
    s1 = np.linalg.solve(K, np.r_[g, 0.0]); s2 = np.linalg.solve(K, np.r_[np.zeros(n), c])
    assert abs(s1[n] + sigma * s2[n] - G["ys"][0]) < 1e-14
A3 = np.array([[2 * x[0], 2 * x[1]]])
K3 = np.block([[np.eye(2), A3.T], [A3, -d * np.eye(1)]])
c3 = np.array([x[0] ** 2 + x[1] ** 2 - 1])
s1 = np.linalg.solve(K3, np.r_[gx3, 0.0]); s2 = np.linalg.solve(K3, np.r_[0.0, 0.0, c3])
assert abs(s1[2] + s * s2[2] - ys3) < 1e-13

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "unit_test.json")
json.dump(dict(G1=G1, G2=G2, G3=G3, G4=G4), open(out, "w"), indent=1)
print("wrote", out)
