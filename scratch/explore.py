import numpy as np, time, sys
sys.path.insert(0,'/root/repo')
from oracle import node as onode
from oracle.mlp import ACT_ID, ACT_TANH, MLP
def ulp(p, seed):
    r = np.random.default_rng(seed).integers(-1, 2, p.size)
    return np.where(r > 0, np.nextafter(p, np.float32(np.inf)), np.where(r < 0, np.nextafter(p, np.float32(-np.inf)), p)).astype(np.float32)
def run(dims, scale, T, B, tol, sense, seed=2, bias=True):
    rng = np.random.default_rng(seed)
    mlp = MLP(dims, [ACT_TANH]*(len(dims)-2)+[ACT_ID]); p = mlp.init(rng, scale=scale)
    u0 = rng.standard_normal((B, dims[0])).astype(np.float32)
    prob = onode.NodeProblem(mlp)
    out=[]
    for pp in (p, ulp(p,11), ulp(p,12)):
        fw = onode.forward(prob, pp, u0, 0.0, T, abstol=tol, reltol=tol, save_everystep=False, save_start=False, dense=(sense==0), record_trace=True)
        adj = onode.adjoint(prob, pp, fw, np.ones_like(fw.us), sensealg=sense, abstol=tol, reltol=tol, record_trace=True)
        out.append((fw.stats.naccept, fw.stats.nreject, adj.stats.naccept, adj.stats.nreject))
    ee = [x[2] for x in fw.stats.trace]; ea=[x[2] for x in adj.stats.trace]
    print(dims, 'scale',scale,'T',T,'B',B,'tol',tol, out, 'fw EEst med %.3f max %.3f | adj med %.3f max %.3f'%(np.median(ee), max(ee), np.median(ea), max(ea)), 'tie', fw.stats.tie_dist, adj.stats.tie_dist, '|u|', np.abs(fw.us).max())
for scale in (1.0, 1.5, 2.0, 3.0):
    t=time.time(); run([64,256,256,64], scale, 1.0, 256, 1e-6, 1); print(time.time()-t)
