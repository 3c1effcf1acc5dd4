"""cfg4 (multiple shooting, 16 groups x 4096 trajectories): wall time of loss + backward with device tensors,
after a warm-up pass.  B200_MSHOOT_LANES=1 forces the groups to run back to back."""
import gc, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffeqflux_b200 as dfx
from tools.run_configs import mlp_params, chain_of

dims = [32, 128, 128, 32]
p, rng = mlp_params(dims, 4, 1.5)
B, T, gs = 4096, 65, 5
ts = np.linspace(0, 4, T, dtype=np.float32)
data = torch.from_numpy(rng.standard_normal((T, B, 32)).astype(np.float32)).cuda()
prob = dfx.ODEProblem(chain_of(dims), None, (0.0, 4.0), None)
for it in range(3):
    pt = torch.from_numpy(p).cuda().requires_grad_(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loss, preds = dfx.multiple_shoot(pt, data, ts, prob, lambda u, uh: ((u - uh) ** 2).sum(), dfx.Tsit5(), gs, abstol=1e-5, reltol=1e-5)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    loss.backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    run = dfx.multiple_shoot.last
    nf = sum(s["nf"] for s in run.stats) + sum(s["nf"] for s in run.adjoint_stats)
    kms = sum(s["kernel_ms"] for s in run.stats) + sum(s["kernel_ms"] for s in run.adjoint_stats)
    print(f"lanes={os.environ.get('B200_MSHOOT_LANES','auto')} pass {it}: fwd {1e3*(t1-t0):.1f} ms, bwd {1e3*(t2-t1):.1f} ms, sum of kernel times {kms:.1f} ms, "
          f"{B*nf/(t2-t0)/1e6:.1f} M RHS evals/s, loss {float(loss):.6e}, grad norm {float(pt.grad.norm()):.6e}")
    del loss, preds, run
    dfx.multiple_shoot.last = None
    gc.collect()
