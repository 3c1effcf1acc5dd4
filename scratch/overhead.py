import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import bench, diffeqflux_b200 as dfx
dev = torch.device("cuda:0")
B = 16384
_, u0_h = bench.synth(B, 2); p_h = bench.synth(8, 2)[0]
chain = dfx.Chain(dfx.Dense(64,256,dfx.tanh), dfx.Dense(256,256,dfx.tanh), dfx.Dense(256,64))
node = dfx.NeuralODE(chain,(0.0,1.0),dfx.Tsit5(),abstol=1e-6,reltol=1e-6,save_everystep=False,save_start=False,sensealg=dfx.BacksolveAdjoint())
p_d = torch.from_numpy(p_h).to(dev); u0_d = torch.from_numpy(u0_h).to(dev); dL_d = torch.ones((1,B,64),device=dev)
flush = torch.empty(256<<20, dtype=torch.uint8, device=dev)
def sync(): torch.cuda.synchronize()
for it in range(6):
    sync(); t0=time.perf_counter()
    if it >= 3: flush.fill_(1)
    sync(); t1=time.perf_counter()
    sol, back = node.pullback(u0_d, p_d)
    sync(); t2=time.perf_counter()
    du0, dp = back(dL_d)
    sync(); t3=time.perf_counter()
    fw = node.last
    print(f"it {it}: flush {1e3*(t1-t0):.2f} ms | fwd wall {1e3*(t2-t1):.2f} (kernel {fw['stats'].kernel_ms:.2f}) | adj wall {1e3*(t3-t2):.2f} (kernel {fw['adjoint_stats'].kernel_ms:.2f})")
