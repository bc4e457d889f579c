import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the hardware experiments live in their own library (make -C bisinger_b200/csrc experiments), not in the product .so
L = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bisinger_b200', 'libbsg_experiments.so'))
L.bsg_experiment_rowoffset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
torch.manual_seed(0)
a = torch.randn(256, 64, device="cuda").bfloat16()
b = torch.randn(64, 64, device="cuda").bfloat16()
for mode in (0, 1):
    res = []
    for r in list(range(0, 18)) + [25, 50, 100, 127]:
        out = torch.full((128, 64), float("nan"), device="cuda")
        rc = L.bsg_experiment_rowoffset(a.data_ptr(), b.data_ptr(), r, mode, out.data_ptr())
        ref = a[r:r + 128].float() @ b.float().t()
        err = (out - ref).abs().max().item()
        res.append((r, "ok" if err < 1e-2 else "BAD %.2f" % err))
    print("mode", mode, res)
