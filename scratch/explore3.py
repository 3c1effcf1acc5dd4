import numpy as np, time, sys
sys.path.insert(0,'/root/repo')
from oracle import node as onode
from oracle.mlp import ACT_ID, ACT_TANH, MLP
def ulp(p, seed, k=4):
    r = np.random.default_rng(seed).integers(-k, k+1, p.size)
    return (p * (1 + r * np.float32(6e-8))).astype(np.float32)
def run(dims, scale, T, B, tol, sense, seed=3, saveat=None):
    rng = np.random.default_rng(seed)
    mlp = MLP(dims, [ACT_TANH]*(len(dims)-2)+[ACT_ID]); p = mlp.init(rng, scale=scale)
    u0 = rng.standard_normal((B, dims[0])).astype(np.float32)
    prob = onode.NodeProblem(mlp)
    out=[]
    kw = dict(save_everystep=False, save_start=False) if saveat is None else dict(saveat=saveat)
    for pp in (p, ulp(p,11), ulp(p,12), ulp(p,13)):
        fw = onode.forward(prob, pp, u0, 0.0, T, abstol=tol, reltol=tol, dense=(sense==0), record_trace=True, **kw)
        adj = onode.adjoint(prob, pp, fw, np.ones_like(fw.us), sensealg=sense, abstol=tol, reltol=tol, record_trace=True)
        out.append((fw.stats.naccept, fw.stats.nreject, adj.stats.naccept, adj.stats.nreject))
    ee = [x[2] for x in fw.stats.trace]; ea=[x[2] for x in adj.stats.trace]
    print(dims, 'scale',scale,'T',T,'B',B,'tol',tol,'sense',sense, out, 'fw EEst med %.3f max %.3f | adj med %.3f max %.3f'%(np.median(ee), max(ee), np.median(ea), max(ea)), flush=True)
for scale, T in ((2.0,1.0),(3.0,1.0),(3.0,2.0),(4.0,2.0)):
    for tol in (1e-6,1e-4):
        run([16,64,64,16], scale, T, 100, tol, 0)
