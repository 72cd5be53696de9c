"""Host-side tooling (no GPU): accuracy of the first control against a fully converged run of the same solver, and work,
as a function of the stop-rule factors.

    python scripts/tolerance_scan.py c3 1,2,4
"""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neo_mpc_planner2_b200 import workloads
from tests.hostsim import HostSim
cfg=sys.argv[1]; n=2048; scs=[float(x) for x in sys.argv[2].split(',')]
wl=workloads.config(cfg,batch=n)
os.environ['HOSTSIM_TOLX']='0'
hs=HostSim(wl.params,wl.cells,wl.resolution,(wl.origin_x,wl.origin_y),0,wl.footprint,max_iterations=1000)
ref,pref=hs.solve(wl.requests,tol_pg=1e-7,tol_f=0.0)
if len(sys.argv)>3: os.environ['HOSTSIM_TOLX']=sys.argv[3]
else: del os.environ['HOSTSIM_TOLX']
hs=HostSim(wl.params,wl.cells,wl.resolution,(wl.origin_x,wl.origin_y),0,wl.footprint)
for sc in scs:
    out,plan=hs.solve(wl.requests,tol_pg=5e-5*sc,tol_f=1e-6*sc)
    d=np.abs(plan[:,:3]-pref[:,:3]).max(1)
    dc=out['cost'].astype(float)-ref['cost'].astype(float)
    ev=out['evals'].astype(int); it=out['iters'].astype(int)
    print('scale %5.1f iters %.1f evals %.1f p99 %d max %d| k8 max-evals %.1f | u0 diff med %.2e p90 %.2e p99 %.2e | dcost med %.1e p99 %.1e max %.1e'%(sc,it.mean(),ev.mean(),np.percentile(ev,99),ev.max(),ev.reshape(-1,8).max(1).mean(),np.median(d),np.percentile(d,90),np.percentile(d,99),np.median(dc),np.percentile(dc,99),dc.max()))
