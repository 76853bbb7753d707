import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import numpy as np
from hypar_b200 import cases
from refrun import run_reference
from oracle import hpo

def rel(a,b):
    d = np.abs(a-b).max(); s = np.abs(b).max()
    return d/(s if s>0 else 1.0)

def check(case, exe='hypar_ref'):
    ms = exe.endswith('mpi1')
    o = run_reference(case,'rhs',exe=exe)
    S = hpo.Setup(case, mpi_semantics=ms); O = hpo.Oracle(S)
    print(case.name, exe, 'x', rel(S.x,o['x']['data']), 'dxinv', rel(S.dxinv,o['dxinv']['data']))
    u = S.local_u0()
    rhs,hyp,par,src = O.rhs(u, parts=True)
    for k,a in (('u',u),('hyp',hyp),('par',par),('source',src),('rhs',rhs)):
        print('   %-7s relLinf %.3e  (max %.3e)'%(k, rel(a,o[k]['data']), np.abs(o[k]['data']).max()))

check(cases.linear_advection_sine(64,'js'))
check(cases.linear_advection_sine(64,'mapped', diffusion=0.01, par_scheme='4'))
check(cases.euler1d_sod(101,'js'))
check(cases.euler1d_sod(101,'z', interp='components', upwinding='rusanov'))
check(cases.ns2d_vortex((32,24),'yc'))
for exe in ('hypar_ref','hypar_ref_mpi1'):
    check(cases.ns3d_turbulence((16,12,10),'mapped'), exe)
    check(cases.ns3d_turbulence((16,12,10),'js', upwinding='roe', viscous=False), exe)
    check(cases.ns3d_rising_bubble((12,16,10),'yc'), exe)
check(cases.ns3d_turbulence((12,12,12),'z', interp='characteristic', viscous=False))
