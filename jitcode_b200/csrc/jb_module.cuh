// The translation-unit tail of every generated module: kernels and the C ABI of include/jitcode_b200.h.
//
// A generated module defines, before including this file:
//   struct jb_generated::Sys   N, NP, NB, NL, NT, Par, rhs(t, y, dy, par), tangent_index(k)
//   JB_MODULE_METHOD           JB_DP5 | JB_DOP853 | JB_RK23
//   JB_MODULE_DRIVER           JB_DRV_ODE | JB_DRV_IVP_DENSE | JB_DRV_IVP_LAND
//   JB_MODULE_STORAGE          JB_STORE_REGISTERS | JB_STORE_GLOBAL | JB_STORE_WARP
//   JB_MODULE_THREADS          threads per block of the integrate kernel
//   JB_MODULE_KREGS            warp storage: how many stage rows (plus the state) are cached in registers
//   JB_MODULE_MIN_BLOCKS       resident blocks per SM the integrate kernel is compiled for (caps its registers)
// and, for JB_STORE_WARP, the tables of the vectorised right-hand side: jb_generated::jb_tf (doubles),
// jb_generated::jb_tu (unsigned indices / 16-bit byte offsets of type Sys::TU), jb_generated::jb_component,
// Sys::TF_COUNT, Sys::TU_COUNT.
// It replaces the module table and entry points of the reference's jitced_template.c:213-240.
#pragma once

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "jb_rk.cuh"

namespace jb {

using Sys = jb_generated::Sys;
using Tab = std::conditional<JB_MODULE_METHOD == JB_DP5, DP5, std::conditional<JB_MODULE_METHOD == JB_RK23, RK23, DOP853>::type>::type;
constexpr int MODULE_MODE = JB_MODULE_STORAGE;
constexpr bool MODULE_REG = MODULE_MODE == JB_STORE_REGISTERS;
constexpr bool MODULE_WARP = MODULE_MODE == JB_STORE_WARP;
constexpr bool MODULE_SHARED = MODULE_MODE == JB_STORE_SHARED;
constexpr bool MODULE_QUAD = MODULE_MODE == JB_STORE_QUAD;
constexpr int MODULE_GROUP = MODULE_QUAD ? Sys::GROUP : 1;      // threads per trajectory of the thread-storage kernels
// thread-per-trajectory with the stages in shared memory: (Y, Z, K rows, y_old) × N × threads doubles per block
constexpr size_t THREAD_SHARED_BYTES = MODULE_SHARED
		? sizeof(double) * (size_t) (2 + Layout<Tab, JB_MODULE_DRIVER>::KR + (Layout<Tab, JB_MODULE_DRIVER>::DENSE ? 1 : 0)) * Sys::N * JB_MODULE_THREADS : 0;
using ModuleLayout = Layout<Tab, JB_MODULE_DRIVER>;

// ---- warp storage: shared memory of a block = [tables of the right-hand side][stage vectors of warp 0][of warp 1]…
template <class S, bool W> struct WarpShared {
	static constexpr int TABLE_DOUBLES = 0, ROWS = 0, WARPS = 1, TEAM = 32, SCRATCH = 0;
	static constexpr size_t BYTES = 0, EVAL_BYTES = 0;
};
template <class S> struct WarpShared<S, true> {
	static constexpr int TABLE_DOUBLES = S::TF_COUNT + (int) (((S::TU_COUNT + 2 * S::N) * sizeof(typename S::TU) + 15) / 16) * 2;
	static constexpr int ROWS = WarpStore<S, Tab, JB_MODULE_DRIVER, JB_MODULE_KREGS>::ROWS;
	static constexpr int TEAM = 32 * S::TEAM_WARPS;              // threads that share one trajectory
	static constexpr int WARPS = JB_MODULE_THREADS / TEAM;       // trajectories (teams) per block
	static constexpr int SCRATCH = 8;                            // doubles per team for its reductions
	static constexpr size_t BYTES = sizeof(double) * (TABLE_DOUBLES + (size_t) WARPS * (ROWS * S::NROW + SCRATCH));
	static constexpr size_t EVAL_BYTES = sizeof(double) * (TABLE_DOUBLES + (size_t) WARPS * (2 * S::NROW + SCRATCH));
};
using Shared = WarpShared<Sys, MODULE_WARP>;

template <class S>
__device__ __forceinline__ void stage_tables(double *smem)
{
	for (int i = threadIdx.x; i < S::TF_COUNT; i += blockDim.x) smem[i] = jb_generated::jb_tf[i];
	typename S::TU *tu = reinterpret_cast<typename S::TU *>(smem + S::TF_COUNT);
	for (int i = threadIdx.x; i < S::TU_COUNT; i += blockDim.x) tu[i] = jb_generated::jb_tu[i];
	for (int i = threadIdx.x; i < S::N; i += blockDim.x) {
		const typename S::TU component = jb_generated::jb_component[i];
		tu[S::TU_COUNT + i] = component;                               // position → component
		tu[S::TU_COUNT + S::N + component] = (typename S::TU) i;         // component → position
	}
	__syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(JB_MODULE_THREADS, JB_MODULE_MIN_BLOCKS)
jb_integrate_kernel(const __grid_constant__ Params a)
{
	if constexpr (MODE == JB_STORE_WARP) {
		// persistent blocks; every team of warps integrates trajectories b, b + (teams in the grid), …
		extern __shared__ __align__(16) double jb_smem[];
		stage_tables<Sys>(jb_smem);
		using T = Team<Sys::TEAM_WARPS>;
		const int team = T::index();
		double *const mine = jb_smem + Shared::TABLE_DOUBLES + (size_t) team * (Shared::ROWS * Sys::NROW + Shared::SCRATCH);
		// with the caller's counter: work queue — a team takes the next trajectory when it has finished its own (adaptive
		// step counts differ); without: every team strides statically over the ensemble
		double *const slot = mine + Shared::ROWS * Sys::NROW;      // the team's reduction scratch doubles as mailbox
		const int64_t stride = (int64_t) gridDim.x * Shared::WARPS;
		int64_t b = (int64_t) blockIdx.x * Shared::WARPS + team - stride;
		while (true) {
			if (a.queue) {
				if (T::rank() == 0) *reinterpret_cast<unsigned long long *>(slot) = atomicAdd(a.queue, 1ull);
				T::sync();
				b = (int64_t) *reinterpret_cast<const unsigned long long *>(slot);
				T::sync();
			} else {
				b += stride;
			}
			if (b >= a.B) break;
			run_trajectory<Sys, Tab, JB_MODULE_DRIVER, MODE, JB_MODULE_KREGS>(a, b, mine, jb_smem);
			T::sync();
		}
	} else if constexpr (MODE == JB_STORE_SHARED) {
		extern __shared__ __align__(16) double jb_smem[];
		const int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
		if (b >= a.B) return;
		run_trajectory<Sys, Tab, JB_MODULE_DRIVER, MODE>(a, b, jb_smem + threadIdx.x);
	} else {
		// one thread per trajectory — or, in quad storage, one group of lanes (the whole group leaves together)
		const int64_t b = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) / MODULE_GROUP;
		if (b >= a.B) return;
		run_trajectory<Sys, Tab, JB_MODULE_DRIVER, MODE>(a, b);
	}
}

// lane refill for the register kernels (jb_rk.cuh, run_lanes): persistent warps, trajectories from the caller's queue
constexpr bool MODULE_REFILL = MODULE_MODE == JB_STORE_REGISTERS;

__global__ void __launch_bounds__(JB_MODULE_THREADS, JB_MODULE_MIN_BLOCKS)
jb_integrate_refill_kernel(const __grid_constant__ Params a)
{
	if constexpr (MODULE_REFILL) run_lanes<Sys, Tab, JB_MODULE_DRIVER>(a);
}

// dY = f(t, Y): the batched counterpart of py_f (jitced_template.c:67-116)
template <int MODE, class S = Sys>
__global__ void __launch_bounds__(JB_MODULE_THREADS)
jb_eval_f_kernel(const double *t, const double *Y, const double *P, double *dY, int64_t B)
{
	if constexpr (MODE == JB_STORE_WARP) {
		extern __shared__ __align__(16) double jb_smem[];
		stage_tables<S>(jb_smem);
		using T = Team<S::TEAM_WARPS>;
		const int team = T::index(), rank = T::rank();
		double *const in = jb_smem + Shared::TABLE_DOUBLES + (size_t) team * (2 * S::NROW + Shared::SCRATCH);
		double *const out = in + S::NROW;
		double *const scratch = out + S::NROW;
		const typename S::TU *const comp = S::components(jb_smem);
		if (rank == 0) for (int i = S::N; i < S::NROW; ++i) in[i] = 0.0;
		const int64_t stride = (int64_t) gridDim.x * Shared::WARPS;
		for (int64_t b = (int64_t) blockIdx.x * Shared::WARPS + team; b < B; b += stride) {
			typename S::Par par;
#pragma unroll
			for (int k = 0; k < S::NP; ++k) par.v[k] = P[b * S::NP + k];
			for (int i = rank; i < S::N; i += T::SIZE) in[i] = Y[b * S::N + comp[i]];
			rhs_warp_outlined<S>(t[b], in, out, par, jb_smem, scratch);
			for (int i = rank; i < S::N; i += T::SIZE) dY[b * S::N + comp[i]] = out[i];
			T::sync();
		}
	} else {
		const int64_t b = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) / MODULE_GROUP;
		if (b >= B) return;
		typename S::Par par;
#pragma unroll
		for (int k = 0; k < S::NP; ++k) par.v[k] = P[tile_offset(b, S::NP) + k * TILE];
		const int64_t off = tile_offset(b, S::N);
		if constexpr (MODE == JB_STORE_QUAD) {
			// lane k of the group evaluates the state's derivative and that of tangent vector k
			const int member = (int) (threadIdx.x % (unsigned) S::GROUP);
			const bool owner = member < S::NL;
			const QVec<S::NB> in{const_cast<double *>(Y) + off, S::NB * member, member == 0, owner}, out{dY + off, S::NB * member, member == 0, owner};
			RVec<2 * S::NB> y, dy;
#pragma unroll
			for (int i = 0; i < 2 * S::NB; ++i) y.v[i] = in(i);
			S::rhs(t[b], y, dy, par);
#pragma unroll
			for (int i = 0; i < 2 * S::NB; ++i) out.set(i, dy.v[i]);
		} else if constexpr (MODE == JB_STORE_REGISTERS) {
			RVec<S::N> y, dy;
#pragma unroll
			for (int i = 0; i < S::N; ++i) y.v[i] = Y[off + i * TILE];
			S::rhs(t[b], y, dy, par);
#pragma unroll
			for (int i = 0; i < S::N; ++i) dY[off + i * TILE] = dy.v[i];
		} else {
			rhs_outlined<S>(t[b], Y + off, dY + off, par);
		}
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

// what the launch code needs to know about the current device; queried once per device and process (the values never
// change; a race between two host threads writes the same numbers twice)
constexpr int MAX_DEVICES = 64;
struct DeviceFacts { std::atomic<int> sm_count{0}, refill_blocks{0}; };
static DeviceFacts device_facts[MAX_DEVICES];

static int current_device_facts(DeviceFacts **facts)
{
	int device = 0;
	cudaError_t err = cudaGetDevice(&device);
	if (err != cudaSuccess || device < 0 || device >= MAX_DEVICES) return fail(err != cudaSuccess ? (int) err : JB_E_UNSUPPORTED, "querying the device");
	DeviceFacts &f = device_facts[device];
	if (f.sm_count.load(std::memory_order_acquire) == 0) {
		int sms = 0;
		err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
		if (err != cudaSuccess || sms <= 0) return fail((int) err, "querying the device");
		f.sm_count.store(sms, std::memory_order_release);
	}
	*facts = &f;
	return 0;
}

// persistent grid of the warp-storage kernels: one block per SM (its warps share the staged tables)
static int warp_grid(int64_t B, size_t shared_bytes, const void *kernel, int *blocks)
{
	DeviceFacts *facts = nullptr;
	if (const int bad = current_device_facts(&facts)) return bad;
	const int sm_count = facts->sm_count.load(std::memory_order_relaxed);
	const cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) shared_bytes);
	if (err != cudaSuccess) return fail((int) err, "reserving shared memory");
	const int64_t needed = (B + Shared::WARPS - 1) / Shared::WARPS;
	*blocks = (int) (needed < sm_count ? needed : sm_count);
	return 0;
}

static int reset_queue(unsigned long long *queue, cudaStream_t stream)
{
	const cudaError_t err = cudaMemsetAsync(queue, 0, sizeof(unsigned long long), stream);
	return err == cudaSuccess ? 0 : fail((int) err, "resetting the trajectory queue");
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
	info->work_vectors = jb::MODULE_MODE == JB_STORE_GLOBAL ? jb::ModuleLayout::WORK_VECTORS : 0;
	info->threads_per_block = JB_MODULE_THREADS;
	info->shared_bytes = (int32_t) (jb::MODULE_SHARED ? jb::THREAD_SHARED_BYTES : jb::Shared::BYTES);
	return 0;
}

int jb_eval_f(const double *t, const double *Y, const double *P, double *dY, int64_t B, int64_t ld, void *stream)
{
	if (B <= 0) return 0;
	if (!t || !Y || !dY || (jb::Sys::NP > 0 && !P) || ld < B || (!jb::MODULE_WARP && ld % JB_TILE))
		return jb::fail(JB_E_BAD_ARGUMENT, "jb_eval_f: bad argument");
	if (jb::MODULE_WARP) {
		int blocks = 0;
		if (const int err = jb::warp_grid(B, jb::Shared::EVAL_BYTES, (const void *) jb::jb_eval_f_kernel<jb::MODULE_MODE>, &blocks)) return err;
		jb::jb_eval_f_kernel<jb::MODULE_MODE><<<blocks, JB_MODULE_THREADS, jb::Shared::EVAL_BYTES, (cudaStream_t) stream>>>(t, Y, P, dY, B);
		return jb::launched("jb_eval_f");
	}
	const unsigned blocks = (unsigned) ((B * jb::MODULE_GROUP + JB_MODULE_THREADS - 1) / JB_MODULE_THREADS);
	jb::jb_eval_f_kernel<jb::MODULE_MODE><<<blocks, JB_MODULE_THREADS, 0, (cudaStream_t) stream>>>(t, Y, P, dY, B);
	return jb::launched("jb_eval_f");
}

int jb_integrate(const jb_integrate_args *x)
{
	if (!x || x->struct_size != sizeof(jb_integrate_args))
		return jb::fail(JB_E_ABI, "jb_integrate: argument block has the wrong size (ABI mismatch)");
	if (x->B <= 0 || x->n_times <= 0) return 0;
	if (x->ld < x->B || (!jb::MODULE_WARP && x->ld % JB_TILE) || !x->times || !x->t || !x->Y || !x->status
			|| (jb::Sys::NP > 0 && !x->P)
			|| (jb::ModuleLayout::STATE_VECTORS > 0 && !x->state)
			|| (jb::ModuleLayout::STATE_SCALARS > 0 && !x->sstate)
			|| (jb::MODULE_MODE == JB_STORE_GLOBAL && !x->work)
			|| (jb::ModuleLayout::DENSE && !x->Yres))
		return jb::fail(JB_E_BAD_ARGUMENT, "jb_integrate: missing or inconsistent buffers");
	if (x->post != JB_POST_NONE) {
		if (jb::Sys::NL <= 0 || (jb::MODULE_QUAD && x->post != JB_POST_LYAP_QR) || (x->post == JB_POST_LYAP_TRANSVERSAL && jb::Sys::NT <= 0)
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
	a.dfactor = dfactor;
	a.ifactor = ifactor;
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
	a.queue = nullptr;
	if (jb::MODULE_WARP) {
		int blocks = 0;
		if (const int err = jb::warp_grid(x->B, jb::Shared::BYTES, (const void *) jb::jb_integrate_kernel<jb::MODULE_MODE>, &blocks)) return err;
		if (x->queue && x->B > (int64_t) blocks * jb::Shared::WARPS) {
			// more trajectories than resident teams: teams draw them from the caller's counter
			if (const int err = jb::reset_queue(x->queue, (cudaStream_t) x->stream)) return err;
			a.queue = x->queue;
		}
		jb::jb_integrate_kernel<jb::MODULE_MODE><<<blocks, JB_MODULE_THREADS, jb::Shared::BYTES, (cudaStream_t) x->stream>>>(a);
		return jb::launched("jb_integrate");
	}
	if (jb::MODULE_REFILL && x->queue && x->post == JB_POST_NONE && x->B >= 4 * JB_MODULE_THREADS && !std::getenv("JITCODE_B200_NO_LANE_REFILL")) {
		// large ensemble without post-step: persistent warps whose lanes take trajectories from the caller's queue
		jb::DeviceFacts *facts = nullptr;
		if (const int bad = jb::current_device_facts(&facts)) return bad;
		int resident = facts->refill_blocks.load(std::memory_order_acquire);
		if (resident == 0) {
			int per_sm = 0;
			const cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jb::jb_integrate_refill_kernel, JB_MODULE_THREADS, 0);
			if (err != cudaSuccess || per_sm <= 0) return jb::fail((int) err, "querying the device");
			resident = facts->sm_count.load(std::memory_order_relaxed) * per_sm;
			facts->refill_blocks.store(resident, std::memory_order_release);
		}
		if (const int err = jb::reset_queue(x->queue, (cudaStream_t) x->stream)) return err;
		a.queue = x->queue;
		const int64_t needed = (x->B + JB_MODULE_THREADS - 1) / JB_MODULE_THREADS;
		jb::jb_integrate_refill_kernel<<<(unsigned) (needed < resident ? needed : resident), JB_MODULE_THREADS, 0, (cudaStream_t) x->stream>>>(a);
		return jb::launched("jb_integrate");
	}
	const unsigned blocks = (unsigned) ((x->B * jb::MODULE_GROUP + JB_MODULE_THREADS - 1) / JB_MODULE_THREADS);
	if (jb::MODULE_SHARED) {
		const cudaError_t err = cudaFuncSetAttribute((const void *) jb::jb_integrate_kernel<jb::MODULE_MODE>,
				cudaFuncAttributeMaxDynamicSharedMemorySize, (int) jb::THREAD_SHARED_BYTES);
		if (err != cudaSuccess) return jb::fail((int) err, "reserving shared memory");
	}
	jb::jb_integrate_kernel<jb::MODULE_MODE><<<blocks, JB_MODULE_THREADS, jb::THREAD_SHARED_BYTES, (cudaStream_t) x->stream>>>(a);
	return jb::launched("jb_integrate");
}

int jb_pack(const double *rows, double *tiled, int64_t B, int64_t n, int64_t ld, void *stream)
{
	if (B <= 0 || n <= 0) return 0;
	if (!rows || !tiled || ld < B || (!jb::MODULE_WARP && ld % JB_TILE)) return jb::fail(JB_E_BAD_ARGUMENT, "jb_pack: bad argument");
	if (jb::MODULE_WARP) {      // row layout: the reference's own
		jb::launch_counter.fetch_add(1, std::memory_order_relaxed);
		const cudaError_t err = cudaMemcpyAsync(tiled, rows, sizeof(double) * B * n, cudaMemcpyDeviceToDevice, (cudaStream_t) stream);
		return err == cudaSuccess ? 0 : jb::fail((int) err, "jb_pack");
	}
	jb::jb_pack_kernel<<<(unsigned) (ld / JB_TILE), 256, 0, (cudaStream_t) stream>>>(rows, tiled, B, n);
	return jb::launched("jb_pack");
}

int jb_unpack(const double *tiled, double *rows, int64_t B, int64_t n, int64_t ld, void *stream)
{
	if (B <= 0 || n <= 0) return 0;
	if (!rows || !tiled || ld < B || (!jb::MODULE_WARP && ld % JB_TILE)) return jb::fail(JB_E_BAD_ARGUMENT, "jb_unpack: bad argument");
	if (jb::MODULE_WARP) {
		jb::launch_counter.fetch_add(1, std::memory_order_relaxed);
		const cudaError_t err = cudaMemcpyAsync(rows, tiled, sizeof(double) * B * n, cudaMemcpyDeviceToDevice, (cudaStream_t) stream);
		return err == cudaSuccess ? 0 : jb::fail((int) err, "jb_unpack");
	}
	jb::jb_unpack_kernel<<<(unsigned) (ld / JB_TILE), 256, 0, (cudaStream_t) stream>>>(tiled, rows, B, n);
	return jb::launched("jb_unpack");
}

}  // extern "C"
