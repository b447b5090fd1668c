import sys, os, time
sys.path.insert(0, "/root/repo")
import torch
B,H,T,d=64,4,399,64
q=torch.randn(B,H,T,d,device="cuda"); k=torch.randn(B,H,T,d,device="cuda"); v=torch.randn(B,H,T,d,device="cuda")
a=torch.randn(B,H,T,T,device="cuda")
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e3
for lib in ("default","cublas","cublaslt"):
    try:
        torch.backends.cuda.preferred_blas_library(lib)
    except Exception as e:
        print(lib, "unavailable", e); continue
    print(lib, "qk^T %.0f us" % t(lambda: torch.matmul(q,k.transpose(-2,-1))), "a@v %.0f us" % t(lambda: a@v),
          "a^T@v %.0f us" % t(lambda: a.transpose(-2,-1)@v))
x=torch.randn(25536,256,device="cuda"); w=torch.randn(5004,256,device="cuda")
for lib in ("cublas","cublaslt"):
    torch.backends.cuda.preferred_blas_library(lib)
    print(lib, "vocab proj %.0f us" % t(lambda: x@w.t()))
