import numpy as np, sys
sys.path.insert(0,'/root/repo')
import diffeqflux_b200 as dfx
from diffeqflux_jl_b200 import layers
from oracle import node as onode
from oracle.mlp import ACT_ID, ACT_TANH, MLP
def make(dims, seed, scale):
    rng = np.random.default_rng(seed)
    acts = [ACT_TANH]*(len(dims)-2)+[ACT_ID]
    mlp = MLP(dims, acts); p = mlp.init(rng, scale=scale)
    chain = dfx.Chain(*[dfx.Dense(i,o,dfx.tanh if a==ACT_TANH else dfx.identity) for i,o,a in zip(dims[:-1],dims[1:],acts)])
    return rng, mlp, p, chain
rng, mlp, p, chain = make([64,256,256,64], 2, 2.0)
u0 = rng.standard_normal((300,64)).astype(np.float32)
kw = dict(abstol=1e-5, reltol=1e-5, saveat=[0.25,0.5,1.0])
node = dfx.NeuralODE(chain,(0.0,1.0),dfx.Tsit5(),**kw)
ctx = layers._context(0); ctx.trace(4096)
sol, back = node.pullback(u0,p)
ftr = ctx.trace().copy()
prob = onode.NodeProblem(mlp)
fw = onode.forward(prob,p,u0,0.0,1.0,dense=True,record_trace=True,**kw)
dL = np.sin(fw.us).astype(np.float32)
du0, dp = back(dL)
tr = ctx.trace().copy()
adj = onode.adjoint(prob,p,fw,dL,abstol=1e-5,reltol=1e-5,record_trace=True)
otr = np.array(adj.stats.trace); oftr=np.array(fw.stats.trace)
print('forward', len(ftr), len(oftr))
for i in range(max(len(ftr),len(oftr))):
    print(i, ftr[i] if i<len(ftr) else None, oftr[i] if i<len(oftr) else None)
print('adjoint', len(tr), len(otr))
for i in range(max(len(tr),len(otr))):
    print(i, tr[i] if i<len(tr) else None, otr[i] if i<len(otr) else None)
print('dp err', np.abs(dp-adj.dp).max()/np.abs(adj.dp).max(), 'du0 err', np.abs(du0-adj.du0).max()/np.abs(adj.du0).max())
