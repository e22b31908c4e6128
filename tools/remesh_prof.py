"""Three cvtx_P3D/P2D_redistribute_on_grid calls on the reference benchmark's workload, for ncu launch lists
(tools/remesh_launch_summary.py) and CVTX_B200_TRACE=1 stage times.  python tools/remesh_prof.py DIM N INTERPOLANT"""
import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from cvortex_b200 import _native
from cvortex_b200.abi import CvtxLibrary, PointerRows
from util import remesh_particles
lib=CvtxLibrary(_native.LIB_PATH); lib.initialise()
dim=int(sys.argv[1]); n=int(sys.argv[2]); name=sys.argv[3]
p=PointerRows(remesh_particles(np.random.default_rng(1),n,dim,signed=False),7 if dim==3 else 4)
h=float(np.cbrt(2.0/n)); out=np.zeros((4*n,7 if dim==3 else 4),np.float32)
fn=lib.P3D_redistribute_on_grid if dim==3 else lib.P2D_redistribute_on_grid
for _ in range(3): fn(p,name,h,1e-4,max_output=4*n,out=out)
