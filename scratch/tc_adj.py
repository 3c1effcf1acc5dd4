import numpy as np, sys, time
sys.path.insert(0,'/root/repo')
import diffeqflux_b200 as dfx
from oracle import node as onode
from oracle.mlp import ACT_ID, ACT_TANH, MLP
def make(dims, seed, scale):
    rng = np.random.default_rng(seed)
    acts = [ACT_TANH]*(len(dims)-2)+[ACT_ID]
    mlp = MLP(dims, acts); p = mlp.init(rng, scale=scale)
    chain = dfx.Chain(*[dfx.Dense(i,o,dfx.tanh if a==ACT_TANH else dfx.identity) for i,o,a in zip(dims[:-1],dims[1:],acts)])
    return rng, mlp, p, chain
rng, mlp, p, chain = make([64,256,256,64], 3, 2.5)
sense = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sa = dfx.BacksolveAdjoint() if sense else dfx.InterpolatingAdjoint()
for B in [int(x) for x in (sys.argv[2:] or ['256', '16384'])]:
    u0 = rng.standard_normal((B,64)).astype(np.float32)
    dL = np.cos(np.arange(B*64, dtype=np.float32).reshape(1,B,64)*0.01)
    res = {}
    for path in (1, 2):
        node = dfx.NeuralODE(chain,(0.0,1.0),dfx.Tsit5(),abstol=1e-6,reltol=1e-6,save_everystep=False,save_start=False,sensealg=sa,path=path)
        for rep in range(2):
            sol, back = node.pullback(u0,p)
            du0, dp = back(dL)
        st, ast = sol.stats, node.last["adjoint_stats"]
        res[path] = (du0, dp)
        print(f"B={B} path={path} fwd eng={st.engine} steps=({st.naccept},{st.nreject}) {st.kernel_ms:.2f} ms | adj eng={ast.engine} steps=({ast.naccept},{ast.nreject}) nf={ast.nf} {ast.kernel_ms:.2f} ms grid={ast.grid}", flush=True)
    rel = lambda a,b: float(np.abs(a-b).max()/np.abs(b).max())
    print("   TC vs FFMA: du0", rel(res[2][0],res[1][0]), "dp", rel(res[2][1],res[1][1]))
    if B <= 256:
        prob = onode.NodeProblem(mlp)
        fw = onode.forward(prob,p,u0,0.0,1.0,abstol=1e-6,reltol=1e-6,save_everystep=False,save_start=False,dense=(sense==0))
        adj = onode.adjoint(prob,p,fw,dL,sensealg=sense,abstol=1e-6,reltol=1e-6)
        print("   oracle adj steps", adj.stats.naccept, adj.stats.nreject, "TC vs oracle du0", rel(res[2][0],adj.du0), "dp", rel(res[2][1],adj.dp), "| FFMA vs oracle dp", rel(res[1][1],adj.dp))
