import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import numpy as np
from hypar_b200 import cases
from refrun import run_reference
from oracle import hpo
def rel(a,b):
    d = np.abs(a-b).max(); s = np.abs(b).max(); return d/(s if s>0 else 1.0)
for case in (cases.linear_advection_sine(64,'js'), cases.euler1d_sod(101,'js'), cases.ns2d_vortex((32,24),'yc'),
             cases.ns3d_turbulence((16,12,10),'mapped'), cases.ns3d_rising_bubble((12,16,10),'yc')):
    o = run_reference(case,'steps',[3])
    S = hpo.Setup(case); O = hpo.Oracle(S)
    u = S.local_u0()
    rk = hpo.RK_TYPES[case.solver['time_scheme_type']]
    for _ in range(3): O.time_step(u, float(case.solver['dt']), rk)
    print(case.name, rel(u, o['ufinal']['data']))
    # pieces
    o2 = run_reference(case,'pieces')
    u = S.local_u0(); O.apply_bc(u)
    print('  u', rel(u,o2['u']['data']), 'cfl', O.cfl(u,float(case.solver['dt'])), o2['cfl'])
    for d in range(S.ndims):
        f = O.flux(u,d); w = O.weno_weights(f,u,d); uc=O.modified_solution(u)
        uL=O.interp(uc,u,w,1,d,1); uR=O.interp(uc,u,w,-1,d,1); fL=O.interp(f,u,w,1,d,0); fR=O.interp(f,u,w,-1,d,0)
        fI=O.upwind(fL,fR,uL,uR,u,d)
        print('  dir',d, [ '%s %.1e'%(k,rel(a,o2[k+'_%d'%d]['data'])) for k,a in (('fluxC',f),('weights',w),('uC',uc),('uL',uL),('uR',uR),('fL',fL),('fR',fR),('fluxI',fI))])
        if 'D1_%d'%d in o2: print('     D1', rel(O.first_derivative(u,d), o2['D1_%d'%d]['data']))
        if 'D2_%d'%d in o2: print('     D2', rel(O.second_derivative(u,d,int(case.solver['par_space_scheme'])), o2['D2_%d'%d]['data']))
