import numpy as np, time, sys
sys.path.insert(0,'/root/repo')
from oracle import node as onode
from oracle.mlp import ACT_ID, ACT_TANH, MLP
def pert(p, seed, k):
    r = np.random.default_rng(seed).integers(-k, k+1, p.size)
    return (p * (1 + r * np.float32(6e-8))).astype(np.float32)
def run(dims, scale, T, B, tol, sense, seed, saveat=None, k=64):
    rng = np.random.default_rng(seed)
    mlp = MLP(dims, [ACT_TANH]*(len(dims)-2)+[ACT_ID]); p = mlp.init(rng, scale=scale)
    u0 = rng.standard_normal((B, dims[0])).astype(np.float32)
    prob = onode.NodeProblem(mlp)
    out=[]
    kw = dict(save_everystep=False, save_start=False) if saveat is None else dict(saveat=saveat)
    for pp in (p, pert(p,11,k), pert(p,12,k), pert(p,13,k)):
        fw = onode.forward(prob, pp, u0, 0.0, T, abstol=tol, reltol=tol, dense=(sense==0), **kw)
        adj = onode.adjoint(prob, pp, fw, np.ones_like(fw.us), sensealg=sense, abstol=tol, reltol=tol)
        out.append((fw.stats.naccept, fw.stats.nreject, adj.stats.naccept, adj.stats.nreject))
    print(dims,'seed',seed,'scale',scale,'T',T,'B',B,'tol',tol,'sense',sense,'saveat',saveat, out, 'STABLE' if len(set(out))==1 else 'unstable', flush=True)
for seed in (2,3,4,5):
    for scale in (2.0,2.5):
        run([64,256,256,64], scale, 1.0, 256, 1e-6, 1, seed)
