import numpy as np, time, sys
sys.path.insert(0,'/root/repo')
from oracle import node as onode
from oracle.mlp import ACT_ID, ACT_TANH, MLP
def ulp(p, seed, k=1):
    r = np.random.default_rng(seed).integers(-k, k+1, p.size)
    q = p.copy()
    for _ in range(k):
        q = np.where(r > 0, np.nextafter(q, np.float32(np.inf)), np.where(r < 0, np.nextafter(q, np.float32(-np.inf)), q)).astype(np.float32)
    return q
# (a) spiral kw4
rng = np.random.default_rng(1)
mlp = MLP([2,50,2],[ACT_TANH,ACT_ID]); p = mlp.init(rng)
u0 = np.array([[0,0],[1,0],[2,0]],np.float32)
prob = onode.NodeProblem(mlp)
for s in range(6):
    pp = p if s==0 else ulp(p, s)
    fw = onode.forward(prob, pp, u0, 0.0, 1.0, abstol=1e-6, reltol=1e-6, saveat=[0.3,0.7], dense=True, record_trace=True)
    print('spiral', s, fw.stats.naccept, fw.stats.nreject, ['%.4f'%x[1] for x in fw.stats.trace], ['%.2e'%x[2] for x in fw.stats.trace])
# (b) mid
rng = np.random.default_rng(3)
mlp = MLP([16,64,64,16],[ACT_TANH,ACT_TANH,ACT_ID]); p = mlp.init(rng, scale=0.7)
u0 = rng.standard_normal((100,16)).astype(np.float32)
prob = onode.NodeProblem(mlp)
for s in range(4):
    pp = p if s==0 else ulp(p, s)
    fw = onode.forward(prob, pp, u0, 0.0, 1.0, abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False, record_trace=True)
    print('mid', s, fw.stats.naccept, fw.stats.nreject, ['%.4f'%x[1] for x in fw.stats.trace], ['%.2e'%x[2] for x in fw.stats.trace])
