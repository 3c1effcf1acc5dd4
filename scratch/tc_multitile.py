import numpy as np, sys
sys.path.insert(0,'/root/repo')
import diffeqflux_b200 as dfx
rng = np.random.default_rng(5)
dims=[64,128,128,64]
p = np.concatenate([np.concatenate([(rng.standard_normal(i*o)*2.0/np.sqrt(i)), rng.standard_normal(o)*0.2]) for i,o in zip(dims[:-1],dims[1:])]).astype(np.float32)
chain = dfx.Chain(dfx.Dense(64,128,dfx.tanh), dfx.Dense(128,128,dfx.tanh), dfx.Dense(128,64))
B = 148*128 + 3000     # > one tile per CTA, ragged
u0 = rng.standard_normal((B,64)).astype(np.float32)
dL = np.cos(np.arange(B*64,dtype=np.float32).reshape(1,B,64)*0.001)
out={}
for sense in (1,0):
  for path in (1,2):
    sa = dfx.BacksolveAdjoint() if sense else dfx.InterpolatingAdjoint()
    node = dfx.NeuralODE(chain,(0.0,1.0),dfx.Tsit5(),abstol=1e-5,reltol=1e-5,save_everystep=False,save_start=False,sensealg=sa,path=path,max_dense_steps=64)
    sol, back = node.pullback(u0,p); du0, dp = back(dL)
    a = node.last["adjoint_stats"]
    out[(sense,path)] = (sol.array(), du0, dp)
    print(sense, path, sol.stats.engine, (sol.stats.naccept, sol.stats.nreject), round(sol.stats.kernel_ms,2), (a.naccept,a.nreject), round(a.kernel_ms,2), sol.stats.grid, flush=True)
  rel=lambda x,y: float(np.abs(x-y).max()/np.abs(y).max())
  print("  sense",sense,"TC vs FFMA: state",rel(out[(sense,2)][0],out[(sense,1)][0]),"du0",rel(out[(sense,2)][1],out[(sense,1)][1]),"dp",rel(out[(sense,2)][2],out[(sense,1)][2]))
