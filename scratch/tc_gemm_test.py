import ctypes, torch, sys, os
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tc_gemm_test.so"))
torch.manual_seed(0)
for (K, N) in ((32, 256), (64, 256), (256, 256), (256, 64)):
    A = torch.randn(128, K, device="cuda"); Bt = torch.randn(N, K, device="cuda")
    ref = (A.double() @ Bt.double().T)
    for nprod in (1, 3, 64):
        D = torch.zeros(128, N, device="cuda")
        rc = lib.tc_gemm_test(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(Bt.data_ptr()), ctypes.c_void_p(D.data_ptr()), K, N, nprod)
        err = (D.double() - ref).abs().max().item() / ref.abs().max().item()
        fp32 = ((A @ Bt.T).double() - ref).abs().max().item() / ref.abs().max().item()
        print(f"K={K} N={N} nprod={nprod} rc={rc} rel err {err:.3e} (torch fp32 matmul err {fp32:.3e})")
