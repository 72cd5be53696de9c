"""Host-side tooling (no GPU): what the lock step of a warp costs.  Solves a workload with the host emulation of the solver
core (tests/hostsim) with tracing on, reads the evaluations per pass of every instance and replays them as warps of k
instances: passes = max over the groups, evaluations = sum over passes of the max backtracking count.  The estimate
(passes * 1050 + evaluations * 375 warp instructions) tracked the ncu instruction counts within a few percent in round 1.

    python scripts/lockstep_sim.py c3 [n]
"""
import numpy as np, sys, os, re, tempfile, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neo_mpc_planner2_b200 import workloads
from tests.hostsim import HostSim
def run(cfg='c3',n=2048,k=8,**kn):
    wl=workloads.config(cfg,batch=n)
    hs=HostSim(wl.params,wl.cells,wl.resolution,(wl.origin_x,wl.origin_y),0,wl.footprint,**kn)
    tmp=tempfile.mktemp()
    fd=os.open(tmp,os.O_WRONLY|os.O_CREAT)
    save=os.dup(2); os.dup2(fd,2)
    hs.lib.hostsim_trace(1)
    out,plan=hs.solve(wl.requests)
    hs.lib.hostsim_trace(0)
    ctypes.CDLL(None).fflush(None)
    os.dup2(save,2); os.close(fd)
    per=[]; cur=None; prev=0
    for line in open(tmp):
        if line.startswith('it '):
            e=int(re.search(r'evals (\d+)',line).group(1))
            if e<=prev or cur is None:
                cur=[]; per.append(cur); prev=0
            cur.append(e-prev); prev=e
    os.unlink(tmp)
    tot_pass=0; tot_ev=0
    for w in range(0,len(per),k):
        grp=per[w:w+k]; L=max(len(p) for p in grp)
        tot_pass+=L
        for t in range(L): tot_ev+=max((p[t] if t<len(p) else 0) for p in grp)
    nw=len(per)/k
    P,E=tot_pass/nw,tot_ev/nw
    print('%s k=%d: inst passes %.1f evals %.1f | warp passes %.1f evals %.1f | est instr/warp %.0f | cost mean %.6f'%(cfg,k,np.mean([len(p) for p in per]),np.mean([sum(p) for p in per]),P,E,P*1050+E*375,out['cost'].astype(float).mean()))
if __name__=='__main__':
    cfg=sys.argv[1] if len(sys.argv)>1 else 'c3'
    k={'c2':16,'c3':8,'c4':4,'c5':8}[cfg]
    run(cfg,int(sys.argv[2]) if len(sys.argv)>2 else 2048,k)
