"""
Measured FP64 peak of the device (for the roofline of the register- and shared-memory-resident kernels;
MEASURED_PEAKS.json only has HBM and bf16 figures).  Compiles csrc/jb_fp64_peak.cu like any other module.
"""

import ctypes

import torch

from . import _build


def fp64_fma_tflops(device=None, repeats=5):
	"""best of `repeats` timings of a DFMA-only kernel, in TFLOP/s (2 flop per FMA)."""
	source = (_build.CSRC_DIR / "jb_fp64_peak.cu").read_text()
	lib = ctypes.CDLL(str(_build.compile_module(source)))
	lib.jb_fp64_peak_launch.restype = ctypes.c_double
	lib.jb_fp64_peak_launch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
	device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
	sink = torch.zeros(1, dtype=torch.float64, device=device)
	blocks = 8 * torch.cuda.get_device_properties(device).multi_processor_count
	stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
	best = 0.0
	for repeat in range(repeats + 1):
		start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		start.record()
		flops = lib.jb_fp64_peak_launch(sink.data_ptr(), blocks, 20000, stream)
		end.record()
		end.synchronize()
		if repeat:      # the first launch warms up
			best = max(best, flops / (start.elapsed_time(end) * 1e-3) / 1e12)
	return best
