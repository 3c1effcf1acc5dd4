import ctypes, torch, os
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tc_wgrad_test.so"))
torch.manual_seed(0)
for (R, N) in ((64, 256), (8, 64), (128, 64)):
  for variant in (0, 1):
    A = torch.randn(R, 128, device="cuda"); Bm = torch.randn(R, N, device="cuda")
    ref = A.double().T @ Bm.double()
    G = torch.full((128, N), 7.0, device="cuda")
    rc = lib.tc_wgrad_test(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(Bm.data_ptr()), ctypes.c_void_p(G.data_ptr()), R, 128, N, variant)
    err = (G.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"R={R} N={N} variant={variant} rc={rc} rel err {err:.3e} |G|max {G.abs().max().item():.3e} nan {torch.isnan(G).sum().item()}")
    print("  G[0,:6]  ", G[0,:6].tolist()); print("  ref[0,:6]", ref[0,:6].float().tolist())
    print("  G[:6,0]  ", G[:6,0].tolist()); print("  ref[:6,0]", ref[:6,0].float().tolist())
    # is it a permutation / partial?  correlation with ref
    print("  frac exact", ((G.double()-ref).abs() < 1e-3).double().mean().item())
