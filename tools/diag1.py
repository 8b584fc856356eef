import numpy as np, scipy.sparse as sp, sys
sys.path.insert(0, '.')
import fpsb200
from fpsb200 import models
from oracle import oracle as O
def rel(a,b): return np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-300)
def handle(A):
    coo = sp.coo_matrix(A); H = fpsb200.B200Handle(A.shape[1], A.shape[0], coo.row, coo.col); H.set_jac_values(coo.data); return H
m,n,k,w,delta = 20000,40000,20,64,1.4901161193847656e-08
A = models.window_random_jacobian(m,n,k,w=w,seed=7); H = handle(A)
rng = np.random.default_rng(1234); r1 = rng.standard_normal(n); r2 = rng.standard_normal(m)
got = H.iter_solve_two_mixed(delta, r1, r2); ref = O.IterativeOracle(A).solve_two_mixed(delta, r1, r2)
print("mixed big", [ (s['niter'], s['solved'], s['status']) for s in got[4]], [(s['niter'], s['solved'], s['status']) for s in ref[4]], [rel(a,b) for a,b in zip(got[:4], ref[:4])])
print(got[4][1], ref[4][1])
for (m,n,k,w,delta) in [(30,60,5,8,0.25),(2000,4000,10,32,0.0),(2000,4000,10,32,0.01)]:
    A = models.window_random_jacobian(m,n,k,w=w,seed=13); H = handle(A)
    rng = np.random.default_rng(99); r1 = rng.standard_normal(n); r2 = rng.standard_normal(m)
    u1,u2,st = H.iter_solve_two_extras(delta, r1, r2); ou1,ou2,ost = O.IterativeOracle(A).solve_two_extras(delta, r1, r2)
    print("extras", m, delta, [(s['niter'], s['solved'], s['status']) for s in st], [(s['niter'], s['solved'], s['status']) for s in ost], rel(u1,ou1), rel(u2,ou2))
    print("   ", st[1], ost[1])
