"""one big row-major FP64 GEMM through axb_dgemm (for ncu traffic experiments): M N K [path]"""
import sys
import torch
sys.path.insert(0, ".")
from pyaxisymflow_b200 import _lib
from pyaxisymflow_b200.device import ptr, stream_ptr
M, N, K = (int(x) for x in sys.argv[1:4])
path = int(sys.argv[4]) if len(sys.argv) > 4 else 0
_lib.call("axb_dgemm_set_path", path)
A = torch.randn((M, K), dtype=torch.float64, device="cuda")
B = torch.randn((K, N), dtype=torch.float64, device="cuda")
C = torch.empty((M, N), dtype=torch.float64, device="cuda")
for _ in range(2):
    _lib.call("axb_dgemm", M, N, K, ptr(A), K, ptr(B), N, ptr(C), N, None, None, 0.0, 0.0, stream_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
_lib.call("axb_dgemm", M, N, K, ptr(A), K, ptr(B), N, ptr(C), N, None, None, 0.0, 0.0, stream_ptr())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"M{M} N{N} K{K} path{path}: {ms:.3f} ms {2.0*M*N*K/ms/1e9:.2f} TFLOP/s")
