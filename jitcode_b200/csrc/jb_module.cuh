// The translation-unit tail of every generated module: kernels and the C ABI of include/jitcode_b200.h.
//
// A generated module defines, before including this file:
//   struct jb_generated::Sys   N, NP, NB, NL, NT, Par, rhs(t, y, dy, par), tangent_index(k)
//   JB_MODULE_METHOD           JB_DP5 | JB_DOP853
//   JB_MODULE_DRIVER           JB_DRV_ODE | JB_DRV_IVP_DENSE | JB_DRV_IVP_LAND
//   JB_MODULE_STORAGE          JB_STORE_REGISTERS | JB_STORE_GLOBAL
//   JB_MODULE_THREADS          threads per block of the integrate kernel
// It replaces the module table and entry points of the reference's jitced_template.c:213-240.
#pragma once

#include <atomic>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "jb_rk.cuh"

namespace jb {

using Sys = jb_generated::Sys;
using Tab = std::conditional<JB_MODULE_METHOD == JB_DP5, DP5, DOP853>::type;
constexpr bool MODULE_REG = JB_MODULE_STORAGE == JB_STORE_REGISTERS;
using ModuleLayout = Layout<Tab, JB_MODULE_DRIVER>;

__global__ void __launch_bounds__(JB_MODULE_THREADS)
jb_integrate_kernel(const __grid_constant__ Params a)
{
	const int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= a.B) return;
	run_trajectory<Sys, Tab, JB_MODULE_DRIVER, MODULE_REG>(a, b);
}

// dY = f(t, Y): the batched counterpart of py_f (jitced_template.c:67-116)
__global__ void __launch_bounds__(JB_MODULE_THREADS)
jb_eval_f_kernel(const double *t, const double *Y, const double *P, double *dY, int64_t B)
{
	const int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= B) return;
	Sys::Par par;
#pragma unroll
	for (int k = 0; k < Sys::NP; ++k) par.v[k] = P[tile_offset(b, Sys::NP) + k * TILE];
	const int64_t off = tile_offset(b, Sys::N);
	if constexpr (MODULE_REG) {
		RVec<Sys::N> y, dy;
#pragma unroll
		for (int i = 0; i < Sys::N; ++i) y.v[i] = Y[off + i * TILE];
		Sys::rhs(t[b], y, dy, par);
#pragma unroll
		for (int i = 0; i < Sys::N; ++i) dY[off + i * TILE] = dy.v[i];
	} else {
		rhs_outlined<Sys>(t[b], Y + off, dY + off, par);
	}
}

// (B, n) row-major  ->  tiles [B/32][n][32]; one block per tile, staged through shared memory so
// that both sides are coalesced.  Lanes beyond B replicate the last trajectory (harmless padding).
constexpr int PACK_CHUNK = 64;      // components per pass
__global__ void __launch_bounds__(256)
jb_pack_kernel(const double *__restrict__ rows, double *__restrict__ tiled, int64_t B, int64_t n)
{
	__shared__ double stage[TILE][PACK_CHUNK + 1];
	const int64_t b0 = (int64_t) blockIdx.x * TILE;
	for (int64_t c0 = 0; c0 < n; c0 += PACK_CHUNK) {
		const int width = (int) (n - c0 < PACK_CHUNK ? n - c0 : PACK_CHUNK);
		for (int e = threadIdx.x; e < TILE * width; e += blockDim.x) {
			const int r = e / width, c = e % width;
			int64_t b = b0 + r;
			if (b >= B) b = B - 1;
			stage[r][c] = rows[b * n + c0 + c];
		}
		__syncthreads();
		for (int e = threadIdx.x; e < TILE * width; e += blockDim.x) {
			const int c = e / TILE, r = e % TILE;
			tiled[(int64_t) blockIdx.x * n * TILE + (c0 + c) * TILE + r] = stage[r][c];
		}
		__syncthreads();
	}
}

__global__ void __launch_bounds__(256)
jb_unpack_kernel(const double *__restrict__ tiled, double *__restrict__ rows, int64_t B, int64_t n)
{
	__shared__ double stage[TILE][PACK_CHUNK + 1];
	const int64_t b0 = (int64_t) blockIdx.x * TILE;
	for (int64_t c0 = 0; c0 < n; c0 += PACK_CHUNK) {
		const int width = (int) (n - c0 < PACK_CHUNK ? n - c0 : PACK_CHUNK);
		for (int e = threadIdx.x; e < TILE * width; e += blockDim.x) {
			const int c = e / TILE, r = e % TILE;
			stage[r][c] = tiled[(int64_t) blockIdx.x * n * TILE + (c0 + c) * TILE + r];
		}
		__syncthreads();
		for (int e = threadIdx.x; e < TILE * width; e += blockDim.x) {
			const int r = e / width, c = e % width;
			if (b0 + r < B) rows[(b0 + r) * n + c0 + c] = stage[r][c];
		}
		__syncthreads();
	}
}

static std::atomic<int64_t> launch_counter{0};
static thread_local char error_text[256] = "";

static int fail(int code, const char *what)
{
	if (code > 0)
		snprintf(error_text, sizeof error_text, "%s: %s", what, cudaGetErrorString((cudaError_t) code));
	else
		snprintf(error_text, sizeof error_text, "%s", what);
	return code;
}

static int launched(const char *what)
{
	launch_counter.fetch_add(1, std::memory_order_relaxed);
	const cudaError_t err = cudaGetLastError();
	return err == cudaSuccess ? 0 : fail((int) err, what);
}

}  // namespace jb

extern "C" {

int jb_abi_version(void) { return JB_ABI_VERSION; }

const char *jb_last_error(void) { return jb::error_text; }

int64_t jb_launch_count(void) { return jb::launch_counter.load(std::memory_order_relaxed); }

int jb_get_info(jb_module_info *info)
{
	if (!info) return jb::fail(JB_E_BAD_ARGUMENT, "jb_get_info: null pointer");
	memset(info, 0, sizeof *info);
	info->abi_version = JB_ABI_VERSION;
	info->n = jb::Sys::N;
	info->n_params = jb::Sys::NP;
	info->n_basic = jb::Sys::NB;
	info->n_lyap = jb::Sys::NL;
	info->n_tangent = jb::Sys::NT;
	info->method = JB_MODULE_METHOD;
	info->driver = JB_MODULE_DRIVER;
	info->storage = JB_MODULE_STORAGE;
	info->state_vectors = jb::ModuleLayout::STATE_VECTORS;
	info->state_scalars = jb::ModuleLayout::STATE_SCALARS;
	info->work_vectors = jb::MODULE_REG ? 0 : jb::ModuleLayout::WORK_VECTORS;
	info->threads_per_block = JB_MODULE_THREADS;
	return 0;
}

int jb_eval_f(const double *t, const double *Y, const double *P, double *dY, int64_t B, int64_t ld, void *stream)
{
	if (B <= 0) return 0;
	if (!t || !Y || !dY || (jb::Sys::NP > 0 && !P) || ld < B || ld % JB_TILE)
		return jb::fail(JB_E_BAD_ARGUMENT, "jb_eval_f: bad argument");
	const unsigned blocks = (unsigned) ((B + JB_MODULE_THREADS - 1) / JB_MODULE_THREADS);
	jb::jb_eval_f_kernel<<<blocks, JB_MODULE_THREADS, 0, (cudaStream_t) stream>>>(t, Y, P, dY, B);
	return jb::launched("jb_eval_f");
}

int jb_integrate(const jb_integrate_args *x)
{
	if (!x || x->struct_size != sizeof(jb_integrate_args))
		return jb::fail(JB_E_ABI, "jb_integrate: argument block has the wrong size (ABI mismatch)");
	if (x->B <= 0 || x->n_times <= 0) return 0;
	if (x->ld < x->B || x->ld % JB_TILE || !x->times || !x->t || !x->Y || !x->status
			|| (jb::Sys::NP > 0 && !x->P)
			|| (jb::ModuleLayout::STATE_VECTORS > 0 && !x->state)
			|| (jb::ModuleLayout::STATE_SCALARS > 0 && !x->sstate)
			|| (!jb::MODULE_REG && !x->work)
			|| (jb::ModuleLayout::DENSE && !x->Yres))
		return jb::fail(JB_E_BAD_ARGUMENT, "jb_integrate: missing or inconsistent buffers");
	if (x->post != JB_POST_NONE) {
		if (jb::Sys::NL <= 0 || (x->post == JB_POST_LYAP_TRANSVERSAL && jb::Sys::NT <= 0)
				|| (x->post == JB_POST_LYAP_RESTRICTED && (jb::Sys::NL != 1 || (x->n_proj > 0 && !x->proj))))
			return jb::fail(JB_E_UNSUPPORTED, "jb_integrate: this module has no tangent vectors for the requested post-step");
	}
	jb::Params a;
	a.B = x->B; a.ld = x->ld;
	a.n_times = x->n_times; a.post = x->post; a.n_proj = x->n_proj; a.lyap_discard = x->lyap_discard;
	a.hairer_reject_by_dfactor = x->hairer_reject_by_dfactor;
	a.times = x->times;
	a.rtol = x->rtol; a.atol = x->atol; a.atol_vec = x->atol_vec;
	a.max_step = x->max_step; a.first_step = x->first_step; a.nsteps = x->nsteps;
	// scipy.integrate.ode's defaults for the Hairer controller (scipy _ode.py:1112-1236)
	const bool dp5 = JB_MODULE_METHOD == JB_DP5;
	a.safety = x->safety > 0.0 ? x->safety : 0.9;
	const double ifactor = x->ifactor > 0.0 ? x->ifactor : (dp5 ? 10.0 : 6.0);
	const double dfactor = x->dfactor > 0.0 ? x->dfactor : (dp5 ? 0.2 : 0.3);
	a.facc1 = 1.0 / dfactor;
	a.facc2 = 1.0 / ifactor;
	if (dp5) {
		a.beta = x->beta == 0.0 ? 0.04 : (x->beta < 0.0 ? 0.0 : x->beta);
		a.expo1 = 0.2 - a.beta * 0.75;
	} else {
		a.beta = x->beta <= 0.0 ? 0.0 : x->beta;
		a.expo1 = 1.0 / 8.0 - a.beta * 0.2;
	}
	if (JB_MODULE_DRIVER != JB_DRV_ODE && x->rtol < 100.0 * 2.220446049250313e-16)
		a.rtol = 100.0 * 2.220446049250313e-16;   // validate_tol, scipy _ivp/common.py:47-51
	a.t = x->t; a.Y = x->Y; a.P = x->P;
	a.state = x->state; a.sstate = x->sstate; a.work = x->work;
	a.Yres = x->Yres; a.Yrec = x->Yrec; a.lyap_rec = x->lyap_rec; a.lyap_acc = x->lyap_acc;
	a.proj = x->proj;
	a.status = x->status; a.n_accepted = x->n_accepted; a.n_rejected = x->n_rejected;
	const unsigned blocks = (unsigned) ((x->B + JB_MODULE_THREADS - 1) / JB_MODULE_THREADS);
	jb::jb_integrate_kernel<<<blocks, JB_MODULE_THREADS, 0, (cudaStream_t) x->stream>>>(a);
	return jb::launched("jb_integrate");
}

int jb_pack(const double *rows, double *tiled, int64_t B, int64_t n, int64_t ld, void *stream)
{
	if (B <= 0 || n <= 0) return 0;
	if (!rows || !tiled || ld < B || ld % JB_TILE) return jb::fail(JB_E_BAD_ARGUMENT, "jb_pack: bad argument");
	jb::jb_pack_kernel<<<(unsigned) (ld / JB_TILE), 256, 0, (cudaStream_t) stream>>>(rows, tiled, B, n);
	return jb::launched("jb_pack");
}

int jb_unpack(const double *tiled, double *rows, int64_t B, int64_t n, int64_t ld, void *stream)
{
	if (B <= 0 || n <= 0) return 0;
	if (!rows || !tiled || ld < B || ld % JB_TILE) return jb::fail(JB_E_BAD_ARGUMENT, "jb_unpack: bad argument");
	jb::jb_unpack_kernel<<<(unsigned) (ld / JB_TILE), 256, 0, (cudaStream_t) stream>>>(tiled, rows, B, n);
	return jb::launched("jb_unpack");
}

}  // extern "C"
