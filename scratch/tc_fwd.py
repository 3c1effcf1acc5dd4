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
for B in (256, 1000, 16384):
    u0 = rng.standard_normal((B,64)).astype(np.float32)
    res = {}
    for path in (1, 2):
        node = dfx.NeuralODE(chain,(0.0,1.0),dfx.Tsit5(),abstol=1e-6,reltol=1e-6,save_everystep=False,save_start=False,path=path)
        sol,_ = node(u0,p,{})
        sol,_ = node(u0,p,{})
        res[path] = sol
        st = sol.stats
        print(f"B={B} path={path} engine={st.engine} retcode={sol.retcode} naccept={st.naccept} nreject={st.nreject} nf={st.nf} kernel_ms={st.kernel_ms:.3f} grid={st.grid} tile={st.tile}", flush=True)
    a, b = res[1].array(), res[2].array()
    print("   TC vs FFMA max rel diff", np.abs(a-b).max()/np.abs(a).max())
    if B <= 1000:
        fw = onode.forward(onode.NodeProblem(mlp), p, u0, 0.0, 1.0, abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False)
        print("   oracle steps", fw.stats.naccept, fw.stats.nreject, "TC vs oracle", np.abs(b-fw.us).max()/np.abs(fw.us).max(), "FFMA vs oracle", np.abs(a-fw.us).max()/np.abs(fw.us).max())
