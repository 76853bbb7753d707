import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import numpy as np
from hypar_b200 import cases
from refrun import run_reference
from oracle import hpo
case = cases.ns3d_turbulence((16,12,10),'mapped')
o = run_reference(case,'rhs',exe='hypar_ref_mpi1')
S = hpo.Setup(case, mpi_semantics=True); O = hpo.Oracle(S)
u = S.local_u0()
rhs,hyp,par,src = O.rhs(u, parts=True)
ref = o['par']['data'].reshape(S.shape_g()); mine = par.reshape(S.shape_g())
d = np.abs(ref-mine).max(axis=-1)
print(d.max(), np.abs(ref).max())
idx = np.argwhere(d>1e-8)
print(len(idx), idx[:10], idx.min(axis=0), idx.max(axis=0))
for ax in range(3):
    print(ax, sorted(set(idx[:,ax].tolist())))
import collections
for ax in range(3):
    print(ax, sorted(collections.Counter(idx[:,ax].tolist()).items()))
# piecewise: compare with serial semantic
par0 = O.parabolic(u, mpi_semantics=False).reshape(S.shape_g())
print('vs serial-sem', np.abs(ref-par0).max())
