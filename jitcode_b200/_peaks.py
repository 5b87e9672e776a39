"""
Measured FP64 peak of the device (for the roofline of the register- and shared-memory-resident kernels;
MEASURED_PEAKS.json only has HBM and bf16 figures).  Compiles csrc/jb_fp64_peak.cu like any other module.
"""

import ctypes

import torch

from . import _build


_cached = {}

def fp64_fma_tflops(device=None, repeats=5, entry="jb_fp64_peak_launch", iterations=20000):
	"""best of `repeats` timings of a DFMA-only kernel, in TFLOP/s (2 flop per FMA); measured once per device and process."""
	device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
	if (entry, device) not in _cached:
		_cached[(entry, device)] = _measure(device, repeats, entry, iterations)
	return _cached[(entry, device)]


def fp64_dmma_tflops(device=None, repeats=5):
	"""the same for the FP64 tensor cores (DMMA.8x8x4, 512 flop per warp instruction)."""
	return fp64_fma_tflops(device, repeats, "jb_dmma_peak_launch", 5000)


def _measure(device, repeats, entry, iterations):
	source = (_build.CSRC_DIR / "jb_fp64_peak.cu").read_text()
	lib = ctypes.CDLL(str(_build.compile_module(source)))
	launch = getattr(lib, entry)
	launch.restype = ctypes.c_double
	launch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
	sink = torch.zeros(1, dtype=torch.float64, device=device)
	blocks = 8 * torch.cuda.get_device_properties(device).multi_processor_count
	stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
	best = 0.0
	for repeat in range(repeats + 1):
		start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		start.record()
		flops = launch(sink.data_ptr(), blocks, iterations, stream)
		end.record()
		end.synchronize()
		if repeat:      # the first launch warms up
			best = max(best, flops / (start.elapsed_time(end) * 1e-3) / 1e12)
	return best
