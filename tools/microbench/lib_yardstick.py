"""Yardstick (NOT product): cuBLAS dgemm / cuSOLVER dpotrf on the same box, to get the measured FP64
tensor roofline denominator and the library bar to beat.  Prints JSON lines."""
import json, sys, time, torch
torch.backends.cuda.preferred_linalg_library("cusolver")
dev = torch.device("cuda:0")
def ev(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for n in (4096, 8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    ms = ev(lambda: torch.matmul(a, b.t(), out=c))
    print(json.dumps({"op": "cublas_dgemm_nt", "n": n, "ms": ms, "tflops": 2 * n**3 / ms * 1e-9}), flush=True)
    del a, b, c
# sustained: 3 s of back-to-back dgemm
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev); c = torch.empty_like(a)
torch.cuda.synchronize(); t0 = time.time(); cnt = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
while time.time() - t0 < 3.0:
    for _ in range(5): torch.matmul(a, b.t(), out=c)
    cnt += 5; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
print(json.dumps({"op": "cublas_dgemm_nt_sustained", "n": n, "tflops": cnt * 2 * n**3 / e0.elapsed_time(e1) * 1e-9}), flush=True)
del a, b, c
for n in (8192, 16384, 40000):
    x = torch.randn(n, 64, dtype=torch.float64, device=dev)
    k = x @ x.t(); k.diagonal().add_(float(n)); del x
    L = torch.empty_like(k)
    ms = ev(lambda: torch.linalg.cholesky_ex(k, out=(L, torch.empty((), dtype=torch.int32, device=dev))), reps=2)
    print(json.dumps({"op": "cusolver_dpotrf", "n": n, "ms": ms, "tflops": n**3 / 3 / ms * 1e-9}), flush=True)
    del k, L; torch.cuda.empty_cache()
