"""Hardware experiment: kind::f8f6f4 (e4m3 x e5m2) MMAs accumulating onto kind::f16 MMAs in one TMEM accumulator."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the hardware experiments live in their own library (make -C bisinger_b200/csrc experiments), not in the product .so
L = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bisinger_b200', 'libbsg_experiments.so'))
L.bsg_experiment_f8.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p]
torch.manual_seed(0)
a16 = torch.randn(256, 64, device="cuda").half()
b16 = torch.randn(64, 64, device="cuda").half()
a8 = torch.randn(256, 128, device="cuda").to(torch.float8_e4m3fn)
b8 = (torch.randn(64, 128, device="cuda") * 0.01).to(torch.float8_e5m2)
for mode in (0, 1, 2):
    res = []
    for r in (0, 1, 3, 8, 13, 16, 100, 128):
        out = torch.full((128, 64), float("nan"), device="cuda")
        rc = L.bsg_experiment_f8(a16.data_ptr(), b16.data_ptr(), a8.data_ptr(), b8.data_ptr(), r, mode, out.data_ptr())
        ref = torch.zeros(128, 64, device="cuda", dtype=torch.float64)
        if not mode & 1: ref += a16[r:r + 128].double() @ b16.double().t()
        if not mode & 2: ref += a8[r:r + 128].double() @ b8.double().t()
        err = (out.double() - ref).abs().max().item()
        res.append((r, "ok %.1e" % err if err < 1e-3 else "BAD %.3f" % err))
    print("mode", mode, res)
