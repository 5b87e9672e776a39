"""times the FP64 tensor-core GEMM of the dense engine for every block-tile shape (development aid):
    python scripts/bench_gemm.py [M N K]"""
import ctypes
import sys
import torch
sys.path.insert(0, ".")
from jitcode_b200 import _dense, _peaks
M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (1000, 8192, 1000)
lib = _dense.dense_library()
A = torch.randn(M, K, dtype=torch.float64, device="cuda")
B = torch.randn(K, N, dtype=torch.float64, device="cuda")
C = torch.empty(M, N, dtype=torch.float64, device="cuda")
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print(f"DMMA peak {_peaks.fp64_dmma_tflops():.1f} TFLOP/s, DFMA peak {_peaks.fp64_fma_tflops():.1f} TFLOP/s")
names = {-1: "chosen", 0: "128x64", 1: "112x64", 2: "96x64", 3: "64x64"}
for shape in (-1, 0, 1, 2, 3):
	call = (lambda: lib.jb_dense_gemm(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, stream)) if shape < 0 else \
		(lambda: lib.jb_dense_gemm_shape(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, shape, stream))
	for _ in range(3):
		call()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(20):
		call()
	e1.record(); torch.cuda.synchronize()
	ms = e0.elapsed_time(e1) / 20
	print(f"{names[shape]:8s} {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:6.2f} TFLOP/s")
ref = A @ B
print("max |C - A@B| =", float((C - ref).abs().max()))
