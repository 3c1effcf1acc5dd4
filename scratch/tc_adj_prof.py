import numpy as np, sys
sys.path.insert(0,'/root/repo')
import diffeqflux_b200 as dfx
rng = np.random.default_rng(3)
dims=[64,256,256,64]
p = np.concatenate([np.concatenate([(rng.standard_normal(i*o)*2.5/np.sqrt(i)), rng.standard_normal(o)*0.25]) for i,o in zip(dims[:-1],dims[1:])]).astype(np.float32)
chain = dfx.Chain(dfx.Dense(64,256,dfx.tanh), dfx.Dense(256,256,dfx.tanh), dfx.Dense(256,64))
B = 16384
u0 = rng.standard_normal((B,64)).astype(np.float32)
node = dfx.NeuralODE(chain,(0.0,1.0),dfx.Tsit5(),abstol=1e-6,reltol=1e-6,save_everystep=False,save_start=False,path=2,sensealg=dfx.BacksolveAdjoint())
sol, back = node.pullback(u0,p); du0, dp = back(np.ones((1,B,64),np.float32))
sol, back = node.pullback(u0,p); du0, dp = back(np.ones((1,B,64),np.float32))
print(sol.stats, node.last["adjoint_stats"])
