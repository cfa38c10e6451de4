"""How many sites of the tie-sensitive golden cases differ from the reference with (a) the default rsqrt form of
directional_control and (b) the reference's own atan2 / cos / sin sequence (-DDABRY_REF_DIRECTIONAL), on tests/hostsim
(glibc libm, no FMA contraction).  python profiles/ref_directional_experiment.py default|ref"""
import os, sys, subprocess
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tests'); sys.path.insert(0, ROOT+'/oracle')
import numpy as np
import parity_util as pu
from dabry_b200 import _native
from dump import solver_record, differing_sites
variant = sys.argv[1]
lib = '/tmp/hostsim_%s.so' % variant
if not os.path.exists(lib):
    flags = ['-DDABRY_REF_DIRECTIONAL'] if variant == 'ref' else []
    subprocess.run(['g++','-O2','-std=c++17','-fPIC','-shared','-ffp-contract=off']+flags+['-x','c++',ROOT+'/tests/hostsim/hostsim.cpp','-o',lib],check=True)
_native.use_library(lib)
keys = ['trap_R','chertovskih_T','gyre_T','gyre_rhoads2010_T','linear_T','moving_vortex_T','obstacle_T','point_symmetric_techy2011_T','spherical_geodesic_T','three_vortices_T','trap_T']
tot=0
for k in keys:
    g = pu.golden(k)
    case, solver = pu.solve_case(k, horizon=float(g['total_duration']))
    rec = solver_record(solver)
    d = differing_sites(rec, g, 1e-9)
    cl = int(np.sum(np.asarray(rec['closure'])[:min(len(rec['closure']),len(g['closure']))] != np.asarray(g['closure'])[:min(len(rec['closure']),len(g['closure']))])) if int(rec['n_sites'])==int(g['n_sites']) else -1
    print('%-32s differing sites %3d of %3d   closure flags differing %d' % (k, len(d), int(g['n_sites']), cl))
    tot+=len(d)
    solver.close()
print('total', tot)
