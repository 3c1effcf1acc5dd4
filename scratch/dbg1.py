import numpy as np, sys
sys.path.insert(0,'/root/repo')
import diffeqflux_b200 as dfx
from diffeqflux_jl_b200 import layers
from oracle import node as onode
from oracle.mlp import ACT_ID, ACT_TANH, MLP
rng = np.random.default_rng(1)
mlp = MLP([2,50,2],[ACT_TANH,ACT_ID]); p = mlp.init(rng)
chain = dfx.Chain(dfx.Dense(2,50,dfx.tanh), dfx.Dense(50,2))
u0 = np.array([2.0,0.0],np.float32)
ts = np.arange(0,1.01,0.1,dtype=np.float32)
node = dfx.NeuralODE(chain,(0.0,1.0),dfx.Tsit5(),saveat=ts)
ctx = layers._context(0); ctx.trace(4096)
sol, back = node.pullback(u0,p)
du0, dp = back(np.ones_like(sol.array()))
tr = ctx.trace()
prob = onode.NodeProblem(mlp)
fw = onode.forward(prob,p,u0,0.0,1.0,saveat=ts,dense=True)
adj = onode.adjoint(prob,p,fw,np.ones_like(fw.us),record_trace=True)
otr = np.array(adj.stats.trace)
print(len(tr), len(otr))
for i in range(max(len(tr),len(otr))):
    a = tr[i] if i < len(tr) else None; b = otr[i] if i < len(otr) else None
    print(i, a, b)
