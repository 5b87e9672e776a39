"""
Run-time compilation, caching and loading of generated CUDA modules — the counterpart of
jitcxde._compile_and_load / save_compiled / module_location (_jitcode.py:367-417, 110-118).

A module is one translation unit (see _codegen.module_source) compiled with
    nvcc -gencode arch=compute_100a,code=sm_100a -shared
into a content-addressed cache (`jitcode_b200/_cache/`, override with $JITCODE_B200_CACHE) and
loaded with ctypes through the C ABI of include/jitcode_b200.h.  There is no fallback: if nvcc
or the shared object is missing, this raises.
"""

import ctypes
import fcntl
import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PACKAGE_DIR = Path(__file__).resolve().parent
CSRC_DIR = PACKAGE_DIR / "csrc"
INCLUDE_DIR = PACKAGE_DIR.parent / "include"
JB_ABI_VERSION = 6
JB_TILE = 32

DEFAULT_COMPILE_ARGS = [
	"-gencode", "arch=compute_100a,code=sm_100a",
	"-O3", "-std=c++17", "-lineinfo",
	"-Xcompiler", "-fPIC",
]

def cache_dir():
	path = Path(os.environ.get("JITCODE_B200_CACHE", PACKAGE_DIR / "_cache"))
	path.mkdir(parents=True, exist_ok=True)
	return path

def find_nvcc():
	for candidate in (os.environ.get("JITCODE_B200_NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
		if candidate and os.path.exists(candidate):
			return candidate
	raise RuntimeError("nvcc not found: jitcode_b200 compiles its CUDA modules at run time and has no CPU fallback.")

_header_digest = None
def header_digest():
	"""the generated source includes our own headers; a change there must invalidate the cache."""
	global _header_digest
	if _header_digest is None:
		digest = hashlib.sha256()
		for path in sorted(CSRC_DIR.glob("*.cuh")) + [INCLUDE_DIR / "jitcode_b200.h"]:
			digest.update(path.read_bytes())
		_header_digest = digest.hexdigest()
	return _header_digest

_nvcc_version = None
def nvcc_version():
	"""release line of `nvcc --version` (part of the cache key: another toolkit must not reuse old binaries)."""
	global _nvcc_version
	if _nvcc_version is None:
		try:
			text = subprocess.run([find_nvcc(), "--version"], capture_output=True, text=True).stdout
		except OSError:
			text = ""
		lines = [line.strip() for line in text.splitlines() if "release" in line or line.startswith("Build")]
		_nvcc_version = " ".join(lines) or "unknown"
	return _nvcc_version

def module_key(source, compile_args):
	digest = hashlib.sha256()
	digest.update(source.encode())
	digest.update("\0".join(compile_args).encode())
	digest.update(header_digest().encode())
	digest.update(nvcc_version().encode())
	return digest.hexdigest()[:20]

def compile_module(source, extra_compile_args=(), extra_link_args=(), verbose=False):
	"""returns the path of the shared object for this source, compiling it unless cached."""
	args = [*DEFAULT_COMPILE_ARGS, *extra_compile_args]
	key = module_key(source, [*args, *extra_link_args])
	directory = cache_dir()
	target = directory / f"jb_{key}.so"
	if target.exists():
		return target
	# Several processes may want the same module at once (torchrun: every rank calls set_integrator).  One of them
	# compiles while the others wait on a file lock and then find the finished shared object; source and binary are
	# written under private names and renamed into place, so nobody ever reads a half-written file.
	with open(directory / f"jb_{key}.lock", "w") as lock:
		fcntl.flock(lock, fcntl.LOCK_EX)
		try:
			if target.exists():
				return target
			source_path = directory / f"jb_{key}.cu"
			private_source = directory / f"jb_{key}.{os.getpid()}.tmp.cu"
			private_source.write_text(source)
			os.replace(private_source, source_path)
			tmp = directory / f"jb_{key}.{os.getpid()}.tmp.so"
			cmd = [find_nvcc(), *args, "-shared", f"-I{INCLUDE_DIR}", f"-I{CSRC_DIR}", str(source_path), "-o", str(tmp), *extra_link_args]
			if verbose:
				cmd.insert(1, "-Xptxas=-v")
				print(" ".join(cmd))
			result = subprocess.run(cmd, capture_output=True, text=True)
			if verbose:
				print(result.stdout, result.stderr)
			if result.returncode != 0:
				raise RuntimeError(f"nvcc failed for {source_path}:\n{result.stderr[-4000:]}")
			os.replace(tmp, target)
			return target
		finally:
			fcntl.flock(lock, fcntl.LOCK_UN)


# ------------------------------------------------------------------ ctypes view of include/jitcode_b200.h

class ModuleInfo(ctypes.Structure):
	_fields_ = [(name, ctypes.c_int32) for name in (
		"abi_version", "n", "n_params", "n_basic", "n_lyap", "n_tangent", "method", "driver", "storage",
		"state_vectors", "state_scalars", "work_vectors", "threads_per_block", "shared_bytes",
	)] + [("reserved", ctypes.c_int32 * 2)]

class IntegrateArgs(ctypes.Structure):
	_fields_ = [
		("struct_size", ctypes.c_uint32), ("post", ctypes.c_int32),
		("B", ctypes.c_int64), ("ld", ctypes.c_int64),
		("n_times", ctypes.c_int32), ("n_proj", ctypes.c_int32),
		("times", ctypes.c_void_p),
		("rtol", ctypes.c_double), ("atol", ctypes.c_double), ("atol_vec", ctypes.c_void_p),
		("max_step", ctypes.c_double), ("first_step", ctypes.c_double),
		("nsteps", ctypes.c_int64),
		("safety", ctypes.c_double), ("ifactor", ctypes.c_double), ("dfactor", ctypes.c_double), ("beta", ctypes.c_double),
		("hairer_reject_by_dfactor", ctypes.c_int32), ("lyap_discard", ctypes.c_int32),
		("t", ctypes.c_void_p), ("Y", ctypes.c_void_p), ("P", ctypes.c_void_p),
		("state", ctypes.c_void_p), ("sstate", ctypes.c_void_p), ("work", ctypes.c_void_p),
		("Yres", ctypes.c_void_p), ("Yrec", ctypes.c_void_p),
		("lyap_rec", ctypes.c_void_p), ("lyap_acc", ctypes.c_void_p), ("proj", ctypes.c_void_p),
		("status", ctypes.c_void_p), ("n_accepted", ctypes.c_void_p), ("n_rejected", ctypes.c_void_p),
		("queue", ctypes.c_void_p),
		("stream", ctypes.c_void_p),
	]

EXPORTS = ("jb_abi_version", "jb_last_error", "jb_get_info", "jb_eval_f", "jb_integrate", "jb_pack", "jb_unpack", "jb_launch_count")


class CudaModuleError(RuntimeError):
	pass


class CudaModule:
	"""one loaded module; thin, typed access to its extern "C" symbols."""

	def __init__(self, path):
		self.path = str(path)
		if not os.path.exists(self.path):
			raise FileNotFoundError(f"compiled CUDA module {self.path} does not exist")
		self.lib = ctypes.CDLL(self.path)
		for name in EXPORTS:
			if not hasattr(self.lib, name):
				raise CudaModuleError(f"{self.path} does not export {name}")
		lib = self.lib
		lib.jb_abi_version.restype = ctypes.c_int
		lib.jb_last_error.restype = ctypes.c_char_p
		lib.jb_launch_count.restype = ctypes.c_int64
		lib.jb_get_info.argtypes = [ctypes.POINTER(ModuleInfo)]
		lib.jb_eval_f.argtypes = [ctypes.c_void_p]*4 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
		lib.jb_integrate.argtypes = [ctypes.POINTER(IntegrateArgs)]
		lib.jb_pack.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
		lib.jb_unpack.argtypes = lib.jb_pack.argtypes
		if lib.jb_abi_version() != JB_ABI_VERSION:
			raise CudaModuleError(f"{self.path} has ABI {lib.jb_abi_version()}, this package speaks {JB_ABI_VERSION}")
		self.info = ModuleInfo()
		self.check(lib.jb_get_info(ctypes.byref(self.info)), "jb_get_info")

	def check(self, code, what):
		if code != 0:
			text = self.lib.jb_last_error().decode(errors="replace")
			raise CudaModuleError(f"{what} failed with code {code}: {text}")

	def launch_count(self):
		return int(self.lib.jb_launch_count())

	def eval_f(self, t, Y, P, dY, B, ld, stream):
		self.check(self.lib.jb_eval_f(t, Y, P, dY, B, ld, stream), "jb_eval_f")

	def integrate(self, args):
		args.struct_size = ctypes.sizeof(IntegrateArgs)
		self.check(self.lib.jb_integrate(ctypes.byref(args)), "jb_integrate")

	def pack(self, rows, tiled, B, n, ld, stream):
		self.check(self.lib.jb_pack(rows, tiled, B, n, ld, stream), "jb_pack")

	def unpack(self, tiled, rows, B, n, ld, stream):
		self.check(self.lib.jb_unpack(tiled, rows, B, n, ld, stream), "jb_unpack")
