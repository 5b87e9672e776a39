// Dense sine-coupled phase oscillators (Kuramoto-type networks) — the one coupling that really is a
// dense contraction (BASELINE config 5, reference examples/kuramoto_network.py:19-26):
//
//     dy_i/dt = ω_i + Σ_j W_ij · sin(y_j − y_i)                       n oscillators, W dense (n × n)
//
// The reference writes this out as n·deg sine terms of straight-line C (10⁵–10⁶ for n = 1000).  Here
//     sin(y_j − y_i) = sin y_j · cos y_i − cos y_j · sin y_i
// turns the coupling of a whole ensemble into ONE FP64 GEMM per right-hand-side evaluation,
//     [G_s | G_c] = W · [sin Z | cos Z]        (n × n) · (n × 2B),        f = ω + cos Z ∘ G_s − sin Z ∘ G_c
// with 2n sincos per trajectory instead of n·deg sines.  The GEMM runs on the FP64 tensor cores
// (mma.sync.m8n8k4.f64 → SASS DMMA.8x8x4; tcgen05 has no f64 kind), the rest is element-wise.
//
// Because the contraction couples nothing across trajectories, every trajectory keeps its own step size:
// the ensemble advances in lock-step ROUNDS — in each round every unfinished trajectory makes one step
// attempt (accepted or rejected by its own error estimate), the six stages of all of them share the six
// GEMMs.  State is structure-of-arrays, trajectory-minor: X[i][b] at X[i*ld + b].
//
// Controller: Hairer's DOPRI5 exactly as in jb_rk.cuh (driver JB_DRV_ODE): HINIT, PI control
// with β = 0.04, `1.01·h` landing rule, fresh start per integrate call.  The host loop of rounds lives in
// jb_dense_integrate (C++, no Python per launch).
#include <cuda_runtime.h>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include "jb_tableaux.cuh"
#include "jitcode_b200.h"

namespace jbd {

using jb::DP5;
using jb::DOP853;
using jb::Idx;
using jb::static_for;

// ---------------------------------------------------------------------------------------------
// C[M][N] = A[M][K] · B[K][N], all row-major FP64, on the FP64 tensor cores.
// Block tile 128 × 64, K chunk 16, 8 warps as 4 (M) × 2 (N), each warp 32 × 32 = 4 × 4 DMMA tiles.
// Column tiles whose flag in `skip` is non-zero are not computed (all their trajectories have finished).

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int LDA = BK + 4;     // shared-memory row lengths chosen so that a half-warp's fragment reads
constexpr int LDB = BN + 4;     // (4 rows × 4 consecutive doubles) hit 32 distinct banks
constexpr size_t GEMM_SHARED_BYTES = sizeof(double) * 2 * (BM * LDA + BK * LDB);     // double-buffered tiles

// Which column tiles a launch computes.  A whole matrix: tile t starts at column t·width.  The right-hand side of a
// SUB-ENSEMBLE (trajectories b0 … b1−1 of an ensemble with ld columns per half): its sine columns b0 … b1−1 and its
// cosine columns ld+b0 … ld+b1−1 — tiles_half tiles each.
struct Columns {
	int tile0, tiles_half, half_tiles;      // tiles_half == 0: the whole matrix
	__host__ __device__ int64_t first_column(int tile, int width) const
	{
		if (tiles_half == 0) return (int64_t) tile * width;
		return (int64_t) (tile < tiles_half ? tile0 + tile : half_tiles + tile0 + tile - tiles_half) * width;
	}
};

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
			: "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256)
dgemm_dmma_kernel(const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C,
		int M, int N, int K, const int *__restrict__ tile_skip, int64_t ld, Columns columns)
{
	const int64_t n0 = columns.first_column((int) blockIdx.x, BN);
	if (tile_skip) {
		// columns n0 … n0+63 are trajectories (n0 mod ld) … of the sin or of the cos half; skip_flags are per 32 trajectories
		const int64_t first = (n0 % ld) / 32;
		if (tile_skip[first] && tile_skip[first + 1]) return;
	}
	extern __shared__ __align__(16) double gemm_shared[];
	double (*sA)[BM * LDA] = reinterpret_cast<double (*)[BM * LDA]>(gemm_shared);
	double (*sB)[BK * LDB] = reinterpret_cast<double (*)[BK * LDB]>(gemm_shared + 2 * BM * LDA);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int wm = warp >> 1, wn = warp & 1;                 // warp position in the block tile
	const int64_t m0 = (int64_t) blockIdx.y * BM;

	double acc[4][4][2];
#pragma unroll
	for (int i = 0; i < 4; ++i)
#pragma unroll
		for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

	auto load_tiles = [&](int buffer, int k0) {
		// A tile: 128 rows × 16 doubles; thread loads 8 doubles (row = tid/2, half = tid%2)
		{
			const int row = tid >> 1, half = tid & 1;
			const int64_t gm = m0 + row;
#pragma unroll
			for (int e = 0; e < 8; ++e) {
				const int kk = half * 8 + e;
				const int gk = k0 + kk;
				sA[buffer][row * LDA + kk] = (gm < M && gk < K) ? A[gm * K + gk] : 0.0;
			}
		}
		// B tile: 16 rows × 64 doubles; thread loads 4 doubles (row = tid/16, cols 4*(tid%16)…)
		{
			const int row = tid >> 4, col = (tid & 15) * 4;
			const int gk = k0 + row;
#pragma unroll
			for (int e = 0; e < 4; ++e) {
				const int64_t gn = n0 + col + e;
				sB[buffer][row * LDB + col + e] = (gk < K && gn < N) ? B[(int64_t) gk * N + gn] : 0.0;
			}
		}
	};

	const int chunks = (K + BK - 1) / BK;
	load_tiles(0, 0);
	__syncthreads();
	for (int chunk = 0; chunk < chunks; ++chunk) {
		const int buffer = chunk & 1;
		if (chunk + 1 < chunks) load_tiles(buffer ^ 1, (chunk + 1) * BK);
		const double *a = sA[buffer] + (wm * 32 + (lane >> 2)) * LDA + (lane & 3);
		const double *b = sB[buffer] + (lane & 3) * LDB + wn * 32 + (lane >> 2);
#pragma unroll
		for (int kk = 0; kk < BK; kk += 4) {
			double fa[4], fb[4];
#pragma unroll
			for (int i = 0; i < 4; ++i) fa[i] = a[i * 8 * LDA + kk];
#pragma unroll
			for (int j = 0; j < 4; ++j) fb[j] = b[kk * LDB + j * 8];
#pragma unroll
			for (int i = 0; i < 4; ++i)
#pragma unroll
				for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
		}
		__syncthreads();
	}
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const int64_t gm = m0 + wm * 32 + i * 8 + (lane >> 2);
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int64_t gn = n0 + wn * 32 + j * 8 + 2 * (lane & 3);
			if (gm < M && gn < N) C[gm * N + gn] = acc[i][j][0];
			if (gm < M && gn + 1 < N) C[gm * N + gn + 1] = acc[i][j][1];
		}
	}
}

// The same product, software-pipelined: tiles travel global → shared memory with cp.async (16-byte copies, zero
// fill beyond the matrix edges) three chunks ahead of the tensor cores, so that the loads of chunk c+2 overlap
// the DMMAs of chunk c.  Needs 16-byte aligned rows (K and N even); the kernel above is the fallback.
//
// The block tile is a template parameter: WM × WN warps, each with MI × NI DMMA tiles (8 × 8), i.e. a block tile of
// (8·MI·WM) × (8·NI·WN).  Measured on 1000 × 8192 × 1000 (BASELINE config 5; profiles/r2_dgemm_*): 128 × 64 with two
// blocks per SM runs at 30.7 TFLOP/s with the DMMA pipe 87 % busy while a block is resident (ncu: top stall
// math_pipe_throttle, shared-memory bank conflicts 1.6 % of the wavefronts, L2 14 %) — 83 % of the measured DMMA peak
// of 37.1 TFLOP/s; 64 × 64 30.2, 96 × 64 28.7, 112 × 64 (which would fill the last wave of blocks better) 28.0,
// 128 × 128 and 256 × 64 with one block per SM 25.3 and 24.1.  So the tall tile is the default and the small one
// serves matrices with few rows.
constexpr int STAGES = 3;

template <int WM_, int WN_, int MI_, int NI_>
struct Shape {
	static constexpr int WM = WM_, WN = WN_, MI = MI_, NI = NI_;
	static constexpr int THREADS = 32 * WM * WN;
	static constexpr int TM = 8 * MI * WM, TN = 8 * NI * WN;      // block tile
	static constexpr int SA = BK + 4, SB = TN + 4;                 // padded rows: fragment reads hit 32 distinct banks
	static constexpr size_t SHARED_BYTES = sizeof(double) * STAGES * (TM * SA + BK * SB);
	static_assert(TN % 32 == 0, "column tiles are unions of 32-trajectory groups (tile_skip)");
};

__device__ __forceinline__ void cp_async_16(double *smem, const double *global, int bytes)
{
	const unsigned target = (unsigned) __cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(target), "l"(global), "r"(bytes) : "memory");
}

template <class S>
__global__ void __launch_bounds__(S::THREADS)
dgemm_dmma_pipelined_kernel(const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C,
		int M, int N, int K, const int *__restrict__ tile_skip, int64_t ld, Columns columns)
{
	constexpr int TM = S::TM, TN = S::TN, SA = S::SA, SB = S::SB, MI = S::MI, NI = S::NI, THREADS = S::THREADS;
	const int64_t n0 = columns.first_column((int) blockIdx.x, TN);
	if (tile_skip) {
		// columns n0 … n0+TN−1 are trajectories (n0 mod ld) … of the sin or of the cos half; flags are per 32 trajectories
		const int64_t first = (n0 % ld) / 32;
		bool all = true;
#pragma unroll
		for (int g = 0; g < TN / 32; ++g) all = all && tile_skip[first + g];
		if (all) return;
	}
	extern __shared__ __align__(16) double gemm_shared[];
	double *const sA = gemm_shared;                                  // [STAGES][TM * SA]
	double *const sB = gemm_shared + STAGES * TM * SA;                // [STAGES][BK * SB]
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int wm = warp / S::WN, wn = warp % S::WN;
	const int64_t m0 = (int64_t) blockIdx.y * TM;

	double acc[MI][NI][2];
#pragma unroll
	for (int i = 0; i < MI; ++i)
#pragma unroll
		for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

	auto issue = [&](int stage, int k0) {
		// A tile: TM rows × 8 chunks of 2 doubles
#pragma unroll
		for (int c = tid; c < TM * (BK / 2); c += THREADS) {
			const int row = c / (BK / 2), kk = (c % (BK / 2)) * 2;
			const int64_t gm = m0 + row;
			const int gk = k0 + kk;
			const int bytes = (gm < M && gk < K) ? 16 : 0;
			cp_async_16(sA + stage * TM * SA + row * SA + kk, A + (bytes ? gm * K + gk : 0), bytes);
		}
		// B tile: 16 rows × TN/2 chunks
#pragma unroll
		for (int c = tid; c < BK * (TN / 2); c += THREADS) {
			const int row = c / (TN / 2), col = (c % (TN / 2)) * 2;
			const int gk = k0 + row;
			const int64_t gn = n0 + col;
			const int bytes = (gk < K && gn < N) ? 16 : 0;
			cp_async_16(sB + stage * BK * SB + row * SB + col, B + (bytes ? (int64_t) gk * N + gn : 0), bytes);
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};

	const int chunks = (K + BK - 1) / BK;
#pragma unroll
	for (int s = 0; s < STAGES - 1; ++s) {
		if (s < chunks) issue(s, s * BK);
		else asm volatile("cp.async.commit_group;" ::: "memory");
	}
	for (int chunk = 0; chunk < chunks; ++chunk) {
		asm volatile("cp.async.wait_group %0;" :: "n"(STAGES - 2) : "memory");
		__syncthreads();                                           // chunk's tiles have landed; everybody is done with chunk-1's
		const int ahead = chunk + STAGES - 1;
		if (ahead < chunks) issue(ahead % STAGES, ahead * BK);
		else asm volatile("cp.async.commit_group;" ::: "memory");
		const int stage = chunk % STAGES;
		const double *a = sA + stage * TM * SA + (wm * 8 * MI + (lane >> 2)) * SA + (lane & 3);
		const double *b = sB + stage * BK * SB + (lane & 3) * SB + wn * 8 * NI + (lane >> 2);
#pragma unroll
		for (int kk = 0; kk < BK; kk += 4) {
			double fa[MI], fb[NI];
#pragma unroll
			for (int i = 0; i < MI; ++i) fa[i] = a[i * 8 * SA + kk];
#pragma unroll
			for (int j = 0; j < NI; ++j) fb[j] = b[kk * SB + j * 8];
#pragma unroll
			for (int i = 0; i < MI; ++i)
#pragma unroll
				for (int j = 0; j < NI; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
		}
	}
#pragma unroll
	for (int i = 0; i < MI; ++i) {
		const int64_t gm = m0 + wm * 8 * MI + i * 8 + (lane >> 2);
#pragma unroll
		for (int j = 0; j < NI; ++j) {
			const int64_t gn = n0 + wn * 8 * NI + j * 8 + 2 * (lane & 3);
			if (gm < M && gn < N) C[gm * N + gn] = acc[i][j][0];
			if (gm < M && gn + 1 < N) C[gm * N + gn + 1] = acc[i][j][1];
		}
	}
}

// ---------------------------------------------------------------------------------------------
// per-trajectory control block (structure of arrays, ctrl[k*ld + b])

enum { C_H = 0, C_HMAX, C_FACOLD, C_ERR2, C_DNF, C_DNY, C_DER2, C_FLAGS, C_NSTEP, C_ERR3,
		C_NACC /* accepted steps of this call */, C_STIFF /* IASTI + 16·NONSTI + 128·(last estimate above the limit) */,
		C_STNUM, C_STDEN /* sums of Hairer's stiffness estimate */,
		// SciPy's persistent stepper (solve_ivp drivers; these survive from call to call): |h| for the next step, start and
		// length of the last step, end of the attempt in flight, and IVP_* bits
		C_HABS, C_TOLD, C_HPREV, C_TNEW, C_IVP, C_COUNT };
enum { F_ACTIVE = 1, F_LAST = 2, F_REJECT = 4, F_ACCEPTED = 8, F_STIFF = 16 /* this attempt, if accepted, is looked at by the stiffness test */,
		F_INIT = 32 /* solve_ivp drivers: the stepper is being created (f(t, y), select_initial_step) */ };
enum { IVP_EXISTS = 1 /* the stepper exists: K[0] = f(t, Y), C_HABS valid */, IVP_STEPPED = 2 /* it has made a step: W_YOLD, W_Q0 … hold its interpolant */ };
// C_FLAGS is stored as a double holding a small int
// work vectors, work[k*n*ld + i*ld + b]
enum { W_Z = 0, W_K0 = 1 /* … W_K0+12 */, W_XS = 14 /* sin | cos: 2 vectors */, W_G = 16 /* G_s | G_c: 2 vectors */,
		W_YOLD = 18 /* solve_ivp with dense output: start of the last step */, W_Q0 = 19 /* … W_Q0+6: its interpolant (4 vectors for the 5(4) pair, 7 for the 8th-order pair) */,
		W_KX = 26 /* … W_KX+2: K[13 … 15], the extra stages of the 8th-order pair's dense output */,
		W_RES = 29 /* … and the sample at the target time */, W_COUNT = 30 };
// work vector of stage derivative K[J]
__host__ __device__ constexpr int krow(int j) { return j < 13 ? W_K0 + j : W_KX + (j - 13); }

// coefficient access as constant expressions (both pairs; jb_tableaux.cuh)
template <class Tab> struct Coef;
template <> struct Coef<DP5> {
	static __host__ __device__ constexpr double a(int s, int j) { return DP5::A[s][j]; }
};
template <> struct Coef<DOP853> {
	static __host__ __device__ constexpr double a(int s, int j) { return jb::dop853::A[s][j]; }
};

struct Dense {
	int64_t B, ld;          // trajectories b0 … B−1 of an ensemble whose arrays have ld columns
	int64_t b0, b1;         // the sub-ensemble this launch sequence works on: columns b0 … b1−1 (multiples of 64)
	int n;
	const double *W, *omega;
	const double *coupling;     // [ld] per-trajectory factor of the coupling sum, or null
	const double *atol_vec;     // [n] per-component atol, or null
	double T, rtol, atol, max_step, first_step;
	int64_t nmax;
	double safety, facc1, facc2, beta, expo1;
	int reject_by_dfactor;      // dop853: shrink by exactly dfactor on a rejection, like SciPy 1.18.1's compiled code
	int driver;                 // enum jb_driver
	double *t, *Y, *work, *ctrl;
	int32_t *status;
	int64_t *n_accepted, *n_rejected;
	int32_t *tile_skip;     // [ld/32]: 1 if no trajectory of the 32-column tile is active
	int32_t *active_count;  // [1]
};

__device__ __host__ __forceinline__ double *vec(const Dense &d, int k) { return d.work + (int64_t) k * d.n * d.ld; }
// the GEMM operands are matrices with 2·ld columns: row i = [sin(i, 0 … ld) | cos(i, 0 … ld)], likewise [G_s | G_c]
__device__ __forceinline__ double &xs_sin(const Dense &d, int i, int64_t b) { return vec(d, W_XS)[(int64_t) i * 2 * d.ld + b]; }
__device__ __forceinline__ double &xs_cos(const Dense &d, int i, int64_t b) { return vec(d, W_XS)[(int64_t) i * 2 * d.ld + d.ld + b]; }
__device__ __forceinline__ double g_sin(const Dense &d, int i, int64_t b) { return vec(d, W_G)[(int64_t) i * 2 * d.ld + b]; }
__device__ __forceinline__ double g_cos(const Dense &d, int i, int64_t b) { return vec(d, W_G)[(int64_t) i * 2 * d.ld + d.ld + b]; }

// element-wise kernels: blockDim = (32 trajectories, 8), each thread covers ROWS_PER_THREAD components
constexpr int ROWS_PER_BLOCK = 64;

template <class F>
__device__ __forceinline__ void for_rows(const Dense &d, F &&body)
{
	const int64_t b = d.b0 + (int64_t) blockIdx.x * 32 + threadIdx.x;
	const int i0 = blockIdx.y * ROWS_PER_BLOCK + threadIdx.y;
	if (b >= d.B) return;
#pragma unroll
	for (int r = 0; r < ROWS_PER_BLOCK / 8; ++r) {
		const int i = i0 + 8 * r;
		if (i < d.n) body(i, b, (int64_t) i * d.ld + b);
	}
}

__device__ __forceinline__ int flags_of(const Dense &d, int64_t b) { return (int) d.ctrl[C_FLAGS * d.ld + b]; }
__device__ __forceinline__ double atol_of(const Dense &d, int i) { return d.atol_vec ? d.atol_vec[i] : d.atol; }

// adds `value` (a per-thread partial sum over this thread's rows) to slot[b]: shared-memory reduction over the
// 8 row lanes of the block, then one atomic per (block, trajectory)
__device__ __forceinline__ void reduce_to(double *slot, int64_t b, bool valid, double value)
{
	__shared__ double partial[8][32];
	partial[threadIdx.y][threadIdx.x] = valid ? value : 0.0;
	__syncthreads();
	if (threadIdx.y == 0 && valid) {
		double total = 0.0;
#pragma unroll
		for (int r = 0; r < 8; ++r) total += partial[r][threadIdx.x];
		atomicAdd(slot + b, total);
	}
	__syncthreads();
}

// XS = [sin V | cos V] for the trajectories that take part in this evaluation
template <int WHICH>    // 0: V = Y, 1: V = Z
__global__ void __launch_bounds__(256) sincos_kernel(const Dense d, int need_flag)
{
	const double *V = WHICH == 0 ? d.Y : vec(d, W_Z);
	for_rows(d, [&](int i, int64_t b, int64_t at) {
		if (!(flags_of(d, b) & need_flag)) return;
		sincos(V[at], &xs_sin(d, i, b), &xs_cos(d, i, b));
	});
}

// K[J] = ω + cos ∘ G_s − sin ∘ G_c   (the right-hand side, given the GEMM result)
template <int J>
__device__ __forceinline__ double finish_rhs(const Dense &d, int i, int64_t b, int64_t at)
{
	const double sum = xs_cos(d, i, b) * g_sin(d, i, b) - xs_sin(d, i, b) * g_cos(d, i, b);
	const double value = d.coupling ? fma(d.coupling[b], sum, d.omega[i]) : d.omega[i] + sum;
	vec(d, krow(J))[at] = value;
	return value;
}

// after the GEMM of f(t, Y): K[0] = f; HINIT sums dnf = Σ(f/sk)², dny = Σ(y/sk)²
__global__ void __launch_bounds__(256) start_kernel(const Dense d, int need_flag)
{
	const int64_t b = d.b0 + (int64_t) blockIdx.x * 32 + threadIdx.x;
	double dnf = 0.0, dny = 0.0;
	const bool active = b < d.B && (flags_of(d, b) & need_flag);
	if (active)
		for_rows(d, [&](int i, int64_t, int64_t at) {
			const double f0 = finish_rhs<0>(d, i, b, at);
			const double sk = fma(d.rtol, fabs(d.Y[at]), atol_of(d, i));
			const double qf = f0 / sk, qy = d.Y[at] / sk;
			dnf = fma(qf, qf, dnf);
			dny = fma(qy, qy, dny);
		});
	reduce_to(d.ctrl + C_DNF * d.ld, b, active, dnf);
	reduce_to(d.ctrl + C_DNY * d.ld, b, active, dny);
}

// HINIT: first guess h0 and the Euler probe Z = Y + h0 K[0], XS = sincos(Z)
// (solve_ivp drivers: select_initial_step's h0 from the root-mean-square norms, scipy _ivp/common.py:108-117)
__device__ __forceinline__ double scipy_h0(const Dense &d, int64_t b)
{
	const double d0 = sqrt(d.ctrl[C_DNY * d.ld + b] / d.n), d1 = sqrt(d.ctrl[C_DNF * d.ld + b] / d.n);
	return (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
}

__global__ void __launch_bounds__(256) probe_kernel(const Dense d, int need_flag)
{
	for_rows(d, [&](int i, int64_t b, int64_t at) {
		if (!(flags_of(d, b) & need_flag)) return;
		const double dnf = d.ctrl[C_DNF * d.ld + b], dny = d.ctrl[C_DNY * d.ld + b];
		double h;
		if (d.driver == JB_DRV_ODE) {
			h = (dnf <= 1e-10 || dny <= 1e-10) ? 1e-6 : sqrt(dny / dnf) * 0.01;
			h = fmin(h, d.ctrl[C_HMAX * d.ld + b]);
		} else {
			h = scipy_h0(d, b);
		}
		const double z = fma(h, vec(d, W_K0)[at], d.Y[at]);
		vec(d, W_Z)[at] = z;
		sincos(z, &xs_sin(d, i, b), &xs_cos(d, i, b));
	});
}

// HINIT: der2 = Σ((f1 − f0)/sk)²
__global__ void __launch_bounds__(256) probe_finish_kernel(const Dense d, int need_flag)
{
	const int64_t b = d.b0 + (int64_t) blockIdx.x * 32 + threadIdx.x;
	double der2 = 0.0;
	const bool active = b < d.B && (flags_of(d, b) & need_flag);
	if (active)
		for_rows(d, [&](int i, int64_t, int64_t at) {
			const double f1 = finish_rhs<1>(d, i, b, at);
			const double q = (f1 - vec(d, W_K0)[at]) / fma(d.rtol, fabs(d.Y[at]), atol_of(d, i));
			der2 = fma(q, q, der2);
		});
	reduce_to(d.ctrl + C_DER2 * d.ld, b, active, der2);
}

// the top of Hairer's step loop for one trajectory: limits, landing rule; returns whether it attempts a step
__device__ __forceinline__ void begin_attempt(const Dense &d, int64_t b, double t, double &h, int &flags, double &nstep)
{
	if (nstep > (double) d.nmax) { d.status[b] = JB_TOO_MANY_STEPS; flags &= ~F_ACTIVE; return; }
	if (0.1 * fabs(h) <= fabs(t) * 2.3e-16) { d.status[b] = JB_STEP_TOO_SMALL; flags &= ~F_ACTIVE; return; }
	if (!(h == h)) { d.status[b] = JB_NONFINITE; flags &= ~F_ACTIVE; return; }
	if (t + 1.01 * h - d.T > 0.0) { h = d.T - t; flags |= F_LAST; }
	nstep += 1.0;
	// Hairer's stiffness test (dopri5.f / dop853.f) looks at every 1000th accepted step of a call and at every step while
	// a suspicion is pending; whether THIS attempt would be such a step is known before it is made
	const int64_t accepted_then = (int64_t) d.ctrl[C_NACC * d.ld + b] + 1;
	const int iasti = (int) d.ctrl[C_STIFF * d.ld + b] & 15;
	flags = (accepted_then % 1000 == 0 || iasti > 0) ? (flags | F_STIFF) : (flags & ~F_STIFF);
}

__device__ __forceinline__ void publish_activity(const Dense &d, int64_t b, bool active)
{
	// one flag per 32-trajectory tile for the GEMM, one counter for the host
	const unsigned any = __ballot_sync(0xffffffffu, active);
	if ((threadIdx.x & 31) == 0 && b < d.b1) {       // (the grid may overshoot the sub-ensemble: the next one's tiles are not ours)
		d.tile_skip[b / 32] = any == 0;
		if (any) atomicAdd(d.active_count, __popc(any));
	}
}

// one thread per trajectory: decide which trajectories integrate at all, reset their controllers
__global__ void __launch_bounds__(256) setup_kernel(const Dense d)
{
	const int64_t b = d.b0 + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	bool active = false;
	if (b < d.B) {
		active = d.status[b] == JB_OK && d.T > d.t[b];
		d.ctrl[C_FLAGS * d.ld + b] = active ? F_ACTIVE : 0;
		d.ctrl[C_HMAX * d.ld + b] = d.max_step > 0.0 ? d.max_step : fabs(d.T - d.t[b]);
		d.ctrl[C_FACOLD * d.ld + b] = 1e-4;
		d.ctrl[C_NSTEP * d.ld + b] = 0.0;
		d.ctrl[C_DNF * d.ld + b] = d.ctrl[C_DNY * d.ld + b] = d.ctrl[C_DER2 * d.ld + b] = d.ctrl[C_ERR2 * d.ld + b] = d.ctrl[C_ERR3 * d.ld + b] = 0.0;
		d.ctrl[C_NACC * d.ld + b] = d.ctrl[C_STIFF * d.ld + b] = d.ctrl[C_STNUM * d.ld + b] = d.ctrl[C_STDEN * d.ld + b] = 0.0;
	}
	publish_activity(d, b, active);
}

// HINIT's result (or first_step) and the first pass through the top of the step loop
template <class Tab>
__global__ void __launch_bounds__(256) first_step_kernel(const Dense d, int have_probe)
{
	const int64_t b = d.b0 + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	bool active = false;
	if (b < d.B) {
		int flags = flags_of(d, b);
		if (flags & F_ACTIVE) {
			const double hmax = d.ctrl[C_HMAX * d.ld + b];
			double h;
			if (!have_probe) {
				h = d.first_step;
			} else {
				const double dnf = d.ctrl[C_DNF * d.ld + b], dny = d.ctrl[C_DNY * d.ld + b];
				double h0 = (dnf <= 1e-10 || dny <= 1e-10) ? 1e-6 : sqrt(dny / dnf) * 0.01;
				h0 = fmin(h0, hmax);
				const double der2 = sqrt(d.ctrl[C_DER2 * d.ld + b]) / h0;
				const double der12 = fmax(fabs(der2), sqrt(dnf));
				const double h1 = der12 <= 1e-15 ? fmax(1e-6, fabs(h0) * 1e-3) : pow(0.01 / der12, 1.0 / Tab::HAIRER_ORDER);
				h = fmin(fmin(100.0 * fabs(h0), h1), hmax);
			}
			double nstep = 0.0;
			begin_attempt(d, b, d.t[b], h, flags, nstep);
			d.ctrl[C_H * d.ld + b] = h;
			d.ctrl[C_NSTEP * d.ld + b] = nstep;
			d.ctrl[C_FLAGS * d.ld + b] = flags;
			d.ctrl[C_ERR2 * d.ld + b] = d.ctrl[C_ERR3 * d.ld + b] = 0.0;
			active = flags & F_ACTIVE;
		}
	}
	publish_activity(d, b, active);
}

// stage ROW of an attempt: (ROW > 1: K[ROW-1] = f from the last GEMM;)  Z = Y + h Σ_j a(ROW, j) K[j];  XS = sincos(Z)
// ROW == 1 also applies the outcome of the previous round: accepted trajectories take Y ← Z, K[0] ← K[S].
// ROW == S of the 8th-order pair (whose error estimate does not need f(t+h, y_new)) also sums the error terms.
template <class Tab, int ROW>
__global__ void __launch_bounds__(256) stage_kernel(const Dense d)
{
	constexpr int S = Tab::NSTAGE;
	constexpr bool SUMS_ERROR = !Tab::ERR_NEEDS_FSAL && ROW == S;
	const int64_t bb = d.b0 + (int64_t) blockIdx.x * 32 + threadIdx.x;
	double s5 = 0.0, s3 = 0.0;
	for_rows(d, [&](int i, int64_t b, int64_t at) {
		const int flags = flags_of(d, b);
		double k[S + 1];
		double y = d.Y[at];
		if constexpr (ROW == 1) {
			// (solve_ivp with dense output: a trajectory that has arrived keeps y_old, y_new and all stages of its last
			// step until commit_ivp_kernel has turned them into the interpolant)
			if ((flags & F_ACCEPTED) && (d.driver != JB_DRV_IVP_DENSE || (flags & F_ACTIVE))) {
				y = vec(d, W_Z)[at];
				d.Y[at] = y;
				k[0] = vec(d, W_K0 + S)[at];
				vec(d, W_K0)[at] = k[0];
			} else {
				k[0] = vec(d, W_K0)[at];
			}
			if (!(flags & F_ACTIVE)) return;
		} else {
			if (!(flags & F_ACTIVE)) return;
			static_for<0, ROW - 1>([&](auto j) {
				constexpr double a = Coef<Tab>::a(ROW, j);
				constexpr bool in_error = SUMS_ERROR && (jb::dop853::E5[j] != 0.0 || jb::dop853::E3[j] != 0.0);
				if constexpr (a != 0.0 || in_error) k[j] = vec(d, W_K0 + j)[at];
			});
			k[ROW - 1] = finish_rhs<ROW - 1>(d, i, b, at);
		}
		const double h = d.ctrl[C_H * d.ld + b];
		double acc = 0.0;
		static_for<0, ROW>([&](auto j) {
			constexpr double a = Coef<Tab>::a(ROW, j);
			if constexpr (a != 0.0) acc = fma(a, k[j], acc);
		});
		const double z = fma(h, acc, y);
		vec(d, W_Z)[at] = z;
		sincos(z, &xs_sin(d, i, b), &xs_cos(d, i, b));
		if constexpr (SUMS_ERROR) {
			double e5 = 0.0, e3 = 0.0;
			static_for<0, S>([&](auto j) {
				constexpr double w5 = jb::dop853::E5[j], w3 = jb::dop853::E3[j];
				if constexpr (w5 != 0.0) e5 = fma(w5, k[j], e5);
				if constexpr (w3 != 0.0) e3 = fma(w3, k[j], e3);
			});
			const double sk = fma(d.rtol, fmax(fabs(y), fabs(z)), atol_of(d, i));
			e5 /= sk;
			e3 /= sk;
			s5 = fma(e5, e5, s5);
			s3 = fma(e3, e3, s3);
		}
	});
	if constexpr (SUMS_ERROR) {
		const bool active = bb < d.B && (flags_of(d, bb) & F_ACTIVE);
		reduce_to(d.ctrl + C_ERR2 * d.ld, bb, active, s5);
		reduce_to(d.ctrl + C_ERR3 * d.ld, bb, active, s3);
	}
}

// after the GEMM of f(t+h, y_new): K[S] = f; for the 5th-order pair also the error sum Σ (h Σ_j E_j K_j / sk)²
template <class Tab>
__global__ void __launch_bounds__(256) error_kernel(const Dense d)
{
	constexpr int S = Tab::NSTAGE;
	const int64_t b = d.b0 + (int64_t) blockIdx.x * 32 + threadIdx.x;
	const int flags = b < d.B ? flags_of(d, b) : 0;
	const bool active = flags & F_ACTIVE;
	const bool stiff_due = active && (flags & F_STIFF);
	double sum = 0.0, stnum = 0.0, stden = 0.0;
	if (active) {
		const double h = d.ctrl[C_H * d.ld + b];
		for_rows(d, [&](int i, int64_t, int64_t at) {
			const double last = finish_rhs<S>(d, i, b, at);
			if constexpr (Tab::ERR_NEEDS_FSAL) {
				double e = DP5::E[S] * last;
				static_for<0, S>([&](auto j) {
					constexpr double w = DP5::E[j];
					if constexpr (w != 0.0) e = fma(w, vec(d, W_K0 + j)[at], e);
				});
				const double sk = fma(d.rtol, fmax(fabs(d.Y[at]), fabs(vec(d, W_Z)[at])), atol_of(d, i));
				const double q = h * e / sk;
				sum = fma(q, q, sum);
			}
			if (stiff_due) {
				// ρ ≈ h·|f(t+h, y_new) − K[S-1]| / |y_new − g|, g = the input of the last stage (jb_rk.cuh, stiffness_sums)
				double acc = 0.0;
				static_for<0, S - 1>([&](auto j) {
					constexpr double a = Coef<Tab>::a(S - 1, j);
					if constexpr (a != 0.0) acc = fma(a, vec(d, W_K0 + j)[at], acc);
				});
				const double dk = last - vec(d, W_K0 + S - 1)[at];
				const double dy = vec(d, W_Z)[at] - fma(h, acc, d.Y[at]);
				stnum = fma(dk, dk, stnum);
				stden = fma(dy, dy, stden);
			}
		});
	}
	if constexpr (Tab::ERR_NEEDS_FSAL) reduce_to(d.ctrl + C_ERR2 * d.ld, b, active, sum);
	if (__syncthreads_or(stiff_due)) {
		reduce_to(d.ctrl + C_STNUM * d.ld, b, stiff_due, stnum);
		reduce_to(d.ctrl + C_STDEN * d.ld, b, stiff_due, stden);
	}
}

// one thread per trajectory: accept or reject, new step size, and the top of the next pass of the step loop
template <class Tab>
__global__ void __launch_bounds__(256) control_kernel(const Dense d)
{
	const int64_t b = d.b0 + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	bool active = false;
	if (b < d.B) {
		int flags = flags_of(d, b);
		flags &= ~F_ACCEPTED;
		if (flags & F_ACTIVE) {
			double t = d.t[b], h = d.ctrl[C_H * d.ld + b], facold = d.ctrl[C_FACOLD * d.ld + b];
			const double hmax = d.ctrl[C_HMAX * d.ld + b];
			double nstep = d.ctrl[C_NSTEP * d.ld + b];
			double err;
			if constexpr (Tab::ERR_NEEDS_FSAL) {
				err = sqrt(d.ctrl[C_ERR2 * d.ld + b] / d.n);
			} else {
				const double s5 = d.ctrl[C_ERR2 * d.ld + b], s3 = d.ctrl[C_ERR3 * d.ld + b];
				double deno = fma(0.01, s3, s5);
				if (deno <= 0.0) deno = 1.0;
				err = (s5 == 0.0 && s3 == 0.0) ? 0.0 : fabs(h) * s5 * sqrt(1.0 / (deno * d.n));
			}
			const double fac11 = pow(err, d.expo1);
			double fac = d.beta > 0.0 ? fac11 / pow(facold, d.beta) : fac11;
			fac = fmax(d.facc2, fmin(d.facc1, fac / d.safety));
			double hnew = h / fac;
			bool gave_up = false;
			if (err <= 1.0) {
				d.n_accepted[b] += 1;
				d.ctrl[C_NACC * d.ld + b] += 1.0;
				if (flags & F_STIFF) {
					int watch = (int) d.ctrl[C_STIFF * d.ld + b];
					int iasti = watch & 15, nonsti = (watch >> 4) & 7, above = watch >> 7;
					const double stnum = d.ctrl[C_STNUM * d.ld + b], stden = d.ctrl[C_STDEN * d.ld + b];
					if (stden > 0.0) above = fabs(h) * sqrt(stnum / stden) > Tab::STIFF_LIMIT;
					if (above) {
						nonsti = 0;
						gave_up = ++iasti == 15;
					} else {
						if (nonsti < 7) ++nonsti;
						if (nonsti == 6) iasti = 0;
					}
					d.ctrl[C_STIFF * d.ld + b] = (double) ((iasti & 15) | (nonsti << 4) | (above << 7));
				}
			}
			d.ctrl[C_STNUM * d.ld + b] = d.ctrl[C_STDEN * d.ld + b] = 0.0;
			if (gave_up) {
				// IDID = −4: the solver returns before it commits the step (t and Y stay)
				d.status[b] = JB_STIFF;
				flags &= ~(F_ACTIVE | F_LAST);
			} else if (err <= 1.0) {
				facold = fmax(err, 1e-4);
				flags |= F_ACCEPTED;
				t = (flags & F_LAST) ? d.T : t + h;
				if (fabs(hnew) > hmax) hnew = hmax;
				if (flags & F_REJECT) hnew = fmin(fabs(hnew), fabs(h));
				flags &= ~F_REJECT;
				if (flags & F_LAST) flags &= ~F_ACTIVE;
			} else {
				hnew = (Tab::ERR_NEEDS_FSAL || !d.reject_by_dfactor) ? h / fmin(d.facc1, fac11 / d.safety) : h / d.facc1;
				flags |= F_REJECT;
				flags &= ~F_LAST;
				d.n_rejected[b] += 1;
			}
			h = hnew;
			if (flags & F_ACTIVE) begin_attempt(d, b, t, h, flags, nstep);
			d.t[b] = t;
			d.ctrl[C_H * d.ld + b] = h;
			d.ctrl[C_FACOLD * d.ld + b] = facold;
			d.ctrl[C_NSTEP * d.ld + b] = nstep;
			d.ctrl[C_ERR2 * d.ld + b] = d.ctrl[C_ERR3 * d.ld + b] = 0.0;
			active = flags & F_ACTIVE;
		}
		d.ctrl[C_FLAGS * d.ld + b] = flags;
	}
	publish_activity(d, b, active);
}


// ---------------------------------------------------------------------------------------------
// SciPy's solve_ivp stepper (scipy _ivp/rk.py RungeKutta, common.py select_initial_step) as IVP_wrapper and
// IVP_wrapper_no_interpolation drive it (integrator_tools.py:98-128); the same arithmetic as the IVP branches of
// jb_rk.cuh's Lane, one trajectory per thread of these control kernels, the vectors in the element-wise kernels above.

// IVP_wrapper.integrate: `while t < T: step()` — and at least one step before the first dense output
__device__ __forceinline__ bool ivp_more_steps(const Dense &d, int64_t b, double t, int ivp)
{
	return t < d.T || (d.driver == JB_DRV_IVP_DENSE && !(ivp & IVP_STEPPED));
}

// the top of RungeKutta._step_impl (rk.py:111-124): limits of h_abs for a new step
__device__ __forceinline__ double ivp_begin_step(const Dense &d, double t, double h_abs)
{
	const double min_step = 10.0 * fabs(::nextafter(t, (double) INFINITY) - t);
	const double max_step = d.max_step > 0.0 ? d.max_step : INFINITY;
	if (h_abs > max_step) h_abs = max_step;
	else if (h_abs < min_step) h_abs = min_step;
	return h_abs;
}

// one pass of its `while not step_accepted` loop up to rk_step (rk.py:126-141): the step to attempt
__device__ __forceinline__ void ivp_begin_attempt(const Dense &d, int64_t b, double t, double &h_abs, int &flags)
{
	const double min_step = 10.0 * fabs(::nextafter(t, (double) INFINITY) - t);
	if (!(fabs(h_abs) <= 1.79769313486231571e308)) { d.status[b] = JB_NONFINITE; flags &= ~F_ACTIVE; return; }
	if (h_abs < min_step) { d.status[b] = JB_STEP_TOO_SMALL; flags &= ~F_ACTIVE; return; }
	const double t_bound = d.driver == JB_DRV_IVP_DENSE ? INFINITY : d.T;
	double t_new = t + h_abs;
	if (t_new - t_bound > 0.0) t_new = t_bound;
	const double h = t_new - t;
	h_abs = fabs(h);
	d.ctrl[C_H * d.ld + b] = h;
	d.ctrl[C_TNEW * d.ld + b] = t_new;
}

// who integrates in this call, and whose stepper has to be created first
__global__ void __launch_bounds__(256) setup_ivp_kernel(const Dense d)
{
	const int64_t b = d.b0 + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	bool active = false;
	if (b < d.B) {
		const int ivp = (int) d.ctrl[C_IVP * d.ld + b];
		active = d.status[b] == JB_OK && ivp_more_steps(d, b, d.t[b], ivp);
		d.ctrl[C_FLAGS * d.ld + b] = active ? (F_ACTIVE | ((ivp & IVP_EXISTS) ? 0 : F_INIT)) : 0;
		d.ctrl[C_DNF * d.ld + b] = d.ctrl[C_DNY * d.ld + b] = d.ctrl[C_DER2 * d.ld + b] = d.ctrl[C_ERR2 * d.ld + b] = d.ctrl[C_ERR3 * d.ld + b] = 0.0;
	}
	publish_activity(d, b, active);
}

// select_initial_step's result for the new steppers; the first attempt of everybody
template <class Tab>
__global__ void __launch_bounds__(256) first_step_ivp_kernel(const Dense d, int have_probe)
{
	const int64_t b = d.b0 + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	bool active = false;
	if (b < d.B) {
		int flags = flags_of(d, b);
		if (flags & F_ACTIVE) {
			double h_abs = d.ctrl[C_HABS * d.ld + b];
			if (flags & F_INIT) {
				if (!have_probe) {
					h_abs = d.first_step;
				} else {
					const double h0 = scipy_h0(d, b);
					const double d1 = sqrt(d.ctrl[C_DNF * d.ld + b] / d.n);
					const double d2 = sqrt(d.ctrl[C_DER2 * d.ld + b] / d.n) / h0;
					const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 1.0 / (Tab::ERR_ORDER + 1));
					h_abs = fmin(fmin(100.0 * h0, h1), d.max_step > 0.0 ? d.max_step : INFINITY);
				}
				d.ctrl[C_IVP * d.ld + b] = IVP_EXISTS;
				d.ctrl[C_TOLD * d.ld + b] = NAN;
				d.ctrl[C_HPREV * d.ld + b] = 0.0;
				flags &= ~F_INIT;
			}
			h_abs = ivp_begin_step(d, d.t[b], h_abs);
			ivp_begin_attempt(d, b, d.t[b], h_abs, flags);
			d.ctrl[C_HABS * d.ld + b] = h_abs;
			d.ctrl[C_FLAGS * d.ld + b] = flags;
			d.ctrl[C_ERR2 * d.ld + b] = d.ctrl[C_ERR3 * d.ld + b] = 0.0;
			active = flags & F_ACTIVE;
		}
	}
	publish_activity(d, b, active);
}

// the rest of the loop (rk.py:143-176): accept or reject, new h_abs, and the next attempt (of this step or the next)
template <class Tab>
__global__ void __launch_bounds__(256) control_ivp_kernel(const Dense d)
{
	const int64_t b = d.b0 + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	bool active = false;
	if (b < d.B) {
		int flags = flags_of(d, b);
		// (dense output: a trajectory that has arrived stays "accepted" until commit_ivp_kernel has formed its interpolant;
		// without dense output stage 1 of this round has already taken Y ← Z for it)
		if ((flags & F_ACTIVE) || d.driver != JB_DRV_IVP_DENSE) flags &= ~F_ACCEPTED;
		if (flags & F_ACTIVE) {
			double t = d.t[b], h_abs = d.ctrl[C_HABS * d.ld + b];
			const double h = d.ctrl[C_H * d.ld + b];
			double err;
			if constexpr (Tab::ERR_NEEDS_FSAL) {
				err = sqrt(d.ctrl[C_ERR2 * d.ld + b] / d.n);
			} else {
				const double s5 = d.ctrl[C_ERR2 * d.ld + b], s3 = d.ctrl[C_ERR3 * d.ld + b];
				double deno = fma(0.01, s3, s5);
				if (deno <= 0.0) deno = 1.0;
				err = (s5 == 0.0 && s3 == 0.0) ? 0.0 : fabs(h) * s5 * sqrt(1.0 / (deno * d.n));
			}
			constexpr double exponent = -1.0 / (Tab::ERR_ORDER + 1);
			if (err < 1.0) {
				double factor = err == 0.0 ? 10.0 : fmin(10.0, 0.9 * pow(err, exponent));
				if (flags & F_REJECT) factor = fmin(1.0, factor);
				h_abs *= factor;
				d.n_accepted[b] += 1;
				d.ctrl[C_TOLD * d.ld + b] = t;
				d.ctrl[C_HPREV * d.ld + b] = h;
				t = d.ctrl[C_TNEW * d.ld + b];
				const int ivp = IVP_EXISTS | IVP_STEPPED;
				d.ctrl[C_IVP * d.ld + b] = ivp;
				flags |= F_ACCEPTED;
				flags &= ~F_REJECT;
				if (ivp_more_steps(d, b, t, ivp)) {
					h_abs = ivp_begin_step(d, t, h_abs);
					ivp_begin_attempt(d, b, t, h_abs, flags);
				} else {
					flags &= ~F_ACTIVE;
				}
			} else {
				// (a NaN error lands here, as in SciPy: the step shrinks until it is too small)
				h_abs *= fmax(0.2, 0.9 * pow(err, exponent));
				flags |= F_REJECT;
				d.n_rejected[b] += 1;
				ivp_begin_attempt(d, b, t, h_abs, flags);
			}
			d.t[b] = t;
			d.ctrl[C_HABS * d.ld + b] = h_abs;
			d.ctrl[C_ERR2 * d.ld + b] = d.ctrl[C_ERR3 * d.ld + b] = 0.0;
			active = flags & F_ACTIVE;
		}
		d.ctrl[C_FLAGS * d.ld + b] = flags;
	}
	publish_activity(d, b, active);
}

// The 8th-order pair's dense output needs three more stages of the step just made (DOP853._dense_output_impl,
// rk.py:693-701): stage s = 13, 14, 15 from (t_old, y_old) with the step's own h, for the trajectories that stepped.
// The GEMMs in between run over the tiles that hold such trajectories.
__global__ void __launch_bounds__(256) mark_accepted_tiles_kernel(const Dense d)
{
	const int64_t b = d.b0 + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned any = __ballot_sync(0xffffffffu, b < d.B && (flags_of(d, b) & F_ACCEPTED));
	if ((threadIdx.x & 31) == 0 && b < d.b1) d.tile_skip[b / 32] = any == 0;
}

template <int ROW>
__global__ void __launch_bounds__(256) extra_stage_kernel(const Dense d)
{
	for_rows(d, [&](int i, int64_t b, int64_t at) {
		if (!(flags_of(d, b) & F_ACCEPTED)) return;
		double k[ROW];
		static_for<0, ROW - 1>([&](auto j) {
			constexpr double a = jb::dop853::A[ROW][j];
			if constexpr (a != 0.0) k[j] = vec(d, krow(j))[at];
		});
		if constexpr (ROW > 13) k[ROW - 1] = finish_rhs<ROW - 1>(d, i, b, at);
		else k[ROW - 1] = vec(d, krow(ROW - 1))[at];
		double acc = 0.0;
		static_for<0, ROW>([&](auto j) {
			constexpr double a = jb::dop853::A[ROW][j];
			if constexpr (a != 0.0) acc = fma(a, k[j], acc);
		});
		const double z = fma(d.ctrl[C_HPREV * d.ld + b], acc, d.Y[at]);
		sincos(z, &xs_sin(d, i, b), &xs_cos(d, i, b));
	});
}

// End of a call of the solve_ivp drivers.  Trajectories whose last attempt was accepted still have y_old in Y, y_new in
// Z and the stages of that step in K: with dense output the step's interpolant is formed first (Q = Kᵀ·P, rk.py:178-180
// for the 5(4) pair), then Y ← Z and K[0] ← K[S] = f(t, Y) for the next call.  With dense output every trajectory then
// delivers its sample from the interpolant of its last step (RkDenseOutput._call_impl), made in this call or an earlier one.
template <class Tab>
__global__ void __launch_bounds__(256) commit_ivp_kernel(const Dense d)
{
	constexpr int S = Tab::NSTAGE;
	for_rows(d, [&](int i, int64_t b, int64_t at) {
		(void) i;
		if (flags_of(d, b) & F_ACCEPTED) {
			if (d.driver == JB_DRV_IVP_DENSE) {
				const double y_old = d.Y[at];
				if constexpr (Tab::ERR_NEEDS_FSAL) {
					double k[S + 1];
					static_for<0, S + 1>([&](auto j) { k[j] = vec(d, W_K0 + j)[at]; });
					static_for<0, 4>([&](auto p) {
						double q = 0.0;
						static_for<0, S + 1>([&](auto j) {
							constexpr double w = DP5::P[j][p];
							if constexpr (w != 0.0) q = fma(w, k[j], q);
						});
						vec(d, W_Q0 + p)[at] = q;
					});
				} else {
					// F[0 … 6] of DOP853._dense_output_impl (rk.py:703-711); K[15] comes out of the last GEMM
					double k[16];
					static_for<0, 15>([&](auto j) { k[j] = vec(d, krow(j))[at]; });
					k[15] = finish_rhs<15>(d, i, b, at);
					const double h = d.ctrl[C_HPREV * d.ld + b];
					const double delta = vec(d, W_Z)[at] - y_old;
					vec(d, W_Q0)[at] = delta;
					vec(d, W_Q0 + 1)[at] = h * k[0] - delta;
					vec(d, W_Q0 + 2)[at] = 2.0 * delta - h * (k[S] + k[0]);
					static_for<0, 4>([&](auto row) {
						double q = 0.0;
						static_for<0, 16>([&](auto j) {
							constexpr double w = jb::dop853::D[row][j];
							if constexpr (w != 0.0) q = fma(w, k[j], q);
						});
						vec(d, W_Q0 + 3 + row)[at] = h * q;
					});
				}
				vec(d, W_YOLD)[at] = y_old;
			}
			d.Y[at] = vec(d, W_Z)[at];
			vec(d, W_K0)[at] = vec(d, W_K0 + S)[at];
		}
		if (d.driver == JB_DRV_IVP_DENSE && ((int) d.ctrl[C_IVP * d.ld + b] & IVP_STEPPED)) {
			const double t_old = d.ctrl[C_TOLD * d.ld + b];
			const double hh = d.t[b] - t_old;
			const double x = (d.T - t_old) / hh;
			if constexpr (Tab::ERR_NEEDS_FSAL) {
				double r = vec(d, W_Q0 + 3)[at];
				r = fma(r, x, vec(d, W_Q0 + 2)[at]);
				r = fma(r, x, vec(d, W_Q0 + 1)[at]);
				r = fma(r, x, vec(d, W_Q0)[at]);
				vec(d, W_RES)[at] = fma(hh * x, r, vec(d, W_YOLD)[at]);
			} else {
				// Dop853DenseOutput._call_impl (rk.py:747-764)
				double r = 0.0;
				static_for<0, 7>([&](auto k) {
					r += vec(d, W_Q0 + 6 - k)[at];
					r *= (k % 2 == 0) ? x : 1.0 - x;
				});
				vec(d, W_RES)[at] = r + vec(d, W_YOLD)[at];
			}
		}
	});
}

// the accepted trajectories of the very last round still have y_new in Z: Y ← Z
__global__ void __launch_bounds__(256) commit_kernel(const Dense d)
{
	for_rows(d, [&](int, int64_t b, int64_t at) {
		if (flags_of(d, b) & F_ACCEPTED) d.Y[at] = vec(d, W_Z)[at];
	});
}

__global__ void clear_accepted_kernel(const Dense d)
{
	const int64_t b = d.b0 + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (b < d.B) d.ctrl[C_FLAGS * d.ld + b] = (double) (flags_of(d, b) & ~F_ACCEPTED);
}

// (B, n) rows ↔ [n][ld] trajectory-minor
__global__ void __launch_bounds__(256) transpose_kernel(const double *__restrict__ in, double *__restrict__ out,
		int64_t rows, int64_t cols, int64_t ld_in, int64_t ld_out)
{
	__shared__ double tile[32][33];
	const int64_t c0 = (int64_t) blockIdx.x * 32, r0 = (int64_t) blockIdx.y * 32;
	for (int r = threadIdx.y; r < 32; r += 8)
		if (r0 + r < rows && c0 + threadIdx.x < cols) tile[r][threadIdx.x] = in[(r0 + r) * ld_in + c0 + threadIdx.x];
	__syncthreads();
	for (int c = threadIdx.y; c < 32; c += 8)
		if (c0 + c < cols && r0 + threadIdx.x < rows) out[(c0 + c) * ld_out + r0 + threadIdx.x] = tile[threadIdx.x][c];
}

static std::atomic<int64_t> launch_counter{0};
static thread_local char error_text[256] = "";

static int fail(int code, const char *what)
{
	if (code > 0) snprintf(error_text, sizeof error_text, "%s: %s", what, cudaGetErrorString((cudaError_t) code));
	else snprintf(error_text, sizeof error_text, "%s", what);
	return code;
}

#define JBD_LAUNCH(kernel, grid, block, ...) do { kernel<<<grid, block, 0, stream>>>(__VA_ARGS__); jbd::launch_counter.fetch_add(1, std::memory_order_relaxed); } while (0)

// the block-tile shapes that are compiled, tallest first
using Shape128 = Shape<4, 2, 4, 4>;      // 128 × 64, 256 threads
using Shape112 = Shape<7, 2, 2, 4>;      // 112 × 64, 448 threads
using Shape96 = Shape<4, 2, 3, 4>;       //  96 × 64, 256 threads
using Shape64 = Shape<4, 2, 2, 4>;       //  64 × 64, 256 threads
constexpr int N_SHAPES = 4;

template <class S>
static void launch_shape(const double *A, const double *B, double *C, int M, int N, int K, const int *tile_skip, int64_t ld, cudaStream_t stream, Columns columns)
{
	// (the attribute is a per-device property of the function; setting it costs a microsecond and keeps this re-entrant)
	cudaFuncSetAttribute(dgemm_dmma_pipelined_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) S::SHARED_BYTES);
	const unsigned tiles = columns.tiles_half ? 2u * columns.tiles_half : (unsigned) ((N + S::TN - 1) / S::TN);
	const dim3 grid(tiles, (M + S::TM - 1) / S::TM);
	dgemm_dmma_pipelined_kernel<S><<<grid, S::THREADS, S::SHARED_BYTES, stream>>>(A, B, C, M, N, K, tile_skip, ld, columns);
}

static void launch_gemm(const double *A, const double *B, double *C, int M, int N, int K, const int *tile_skip, int64_t ld, cudaStream_t stream,
		int shape = -1, Columns columns = Columns{0, 0, 0})
{
	const bool aligned = K % 2 == 0 && N % 2 == 0 && (((uintptr_t) A | (uintptr_t) B) & 15) == 0;
	if (!aligned) {
		cudaFuncSetAttribute(dgemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) GEMM_SHARED_BYTES);
		const unsigned tiles = columns.tiles_half ? 2u * columns.tiles_half : (unsigned) ((N + BN - 1) / BN);
		const dim3 grid(tiles, (M + BM - 1) / BM);
		dgemm_dmma_kernel<<<grid, 256, GEMM_SHARED_BYTES, stream>>>(A, B, C, M, N, K, tile_skip, ld, columns);
		launch_counter.fetch_add(1, std::memory_order_relaxed);
		return;
	}
	if (shape < 0) shape = M <= 64 ? 3 : (M <= 96 ? 2 : 0);
	switch (shape) {
	case 1: launch_shape<Shape112>(A, B, C, M, N, K, tile_skip, ld, stream, columns); break;
	case 2: launch_shape<Shape96>(A, B, C, M, N, K, tile_skip, ld, stream, columns); break;
	case 3: launch_shape<Shape64>(A, B, C, M, N, K, tile_skip, ld, stream, columns); break;
	default: launch_shape<Shape128>(A, B, C, M, N, K, tile_skip, ld, stream, columns); break;
	}
	launch_counter.fetch_add(1, std::memory_order_relaxed);
}

static void gemm(const Dense &d, cudaStream_t stream)
{
	// (all tile shapes are 64 columns wide, like the padding of ld)
	const Columns columns{(int) (d.b0 / 64), (int) ((d.b1 - d.b0) / 64), (int) (d.ld / 64)};
	launch_gemm(d.W, vec(d, W_XS), vec(d, W_G), d.n, (int) (2 * d.ld), d.n, d.tile_skip, d.ld, stream, -1, columns);
}

}  // namespace jbd

extern "C" {

const char *jb_dense_last_error(void) { return jbd::error_text; }
int64_t jb_dense_launch_count(void) { return jbd::launch_counter.load(std::memory_order_relaxed); }
int jb_dense_work_vectors(void) { return jbd::W_COUNT; }
int jb_dense_control_scalars(void) { return jbd::C_COUNT; }
int jb_dense_result_vector(void) { return jbd::W_RES; }

int jb_dense_gemm(const double *A, const double *B, double *C, int M, int N, int K, void *stream_)
{
	cudaStream_t stream = (cudaStream_t) stream_;
	jbd::launch_gemm(A, B, C, M, N, K, nullptr, 0, stream);
	const cudaError_t err = cudaGetLastError();
	return err == cudaSuccess ? 0 : jbd::fail((int) err, "jb_dense_gemm");
}

int jb_dense_gemm_shape(const double *A, const double *B, double *C, int M, int N, int K, int shape, void *stream_)
{
	if (shape < 0 || shape >= jbd::N_SHAPES) return jbd::fail(JB_E_BAD_ARGUMENT, "jb_dense_gemm_shape: shapes are 0 (128 x 64), 1 (112 x 64), 2 (96 x 64), 3 (64 x 64)");
	jbd::launch_gemm(A, B, C, M, N, K, nullptr, 0, (cudaStream_t) stream_, shape);
	const cudaError_t err = cudaGetLastError();
	return err == cudaSuccess ? 0 : jbd::fail((int) err, "jb_dense_gemm_shape");
}

int jb_dense_transpose(const double *in, double *out, int64_t rows, int64_t cols, int64_t ld_in, int64_t ld_out, void *stream_)
{
	cudaStream_t stream = (cudaStream_t) stream_;
	const dim3 grid((unsigned) ((cols + 31) / 32), (unsigned) ((rows + 31) / 32));
	JBD_LAUNCH(jbd::transpose_kernel, grid, dim3(32, 8), in, out, rows, cols, ld_in, ld_out);
	const cudaError_t err = cudaGetLastError();
	return err == cudaSuccess ? 0 : jbd::fail((int) err, "jb_dense_transpose");
}

// dY = f(Y) for B states (Y, dY: [n][ld]); work as for jb_dense_integrate
int jb_dense_eval_f(const jb_dense_args *x, double *dY)
{
	if (!x || x->struct_size != sizeof(jb_dense_args)) return jbd::fail(JB_E_ABI, "jb_dense_eval_f: argument block has the wrong size");
	jbd::Dense d{};
	d.B = x->B; d.ld = x->ld; d.n = x->n; d.W = x->W; d.omega = x->omega; d.coupling = x->coupling; d.atol_vec = x->atol_vec;
	d.b0 = 0; d.b1 = x->ld;
	d.Y = x->Y; d.work = x->work; d.ctrl = x->ctrl; d.status = x->status;
	d.tile_skip = x->scratch; d.active_count = x->scratch + x->ld / 32 + 1;
	d.t = x->t; d.T = INFINITY;
	cudaStream_t stream = (cudaStream_t) x->stream;
	const dim3 eblock(32, 8), egrid((unsigned) (d.ld / 32), (unsigned) ((d.n + jbd::ROWS_PER_BLOCK - 1) / jbd::ROWS_PER_BLOCK));
	JBD_LAUNCH(jbd::setup_kernel, (unsigned) ((d.B + 255) / 256), 256, d);
	JBD_LAUNCH(jbd::sincos_kernel<0>, egrid, eblock, d, jbd::F_ACTIVE);
	jbd::gemm(d, stream);
	JBD_LAUNCH(jbd::start_kernel, egrid, eblock, d, (int) jbd::F_ACTIVE);
	cudaError_t err = cudaMemcpyAsync(dY, jbd::vec(d, jbd::W_K0), sizeof(double) * d.n * d.ld, cudaMemcpyDeviceToDevice, stream);
	if (err == cudaSuccess) err = cudaGetLastError();
	return err == cudaSuccess ? 0 : jbd::fail((int) err, "jb_dense_eval_f");
}

}  // extern "C"

namespace jbd {

template <class Tab, int ROW = 1>
static void launch_stages(const Dense &d, dim3 egrid, dim3 eblock, cudaStream_t stream)
{
	if constexpr (ROW <= Tab::NSTAGE) {
		JBD_LAUNCH((stage_kernel<Tab, ROW>), egrid, eblock, d);
		gemm(d, stream);
		launch_stages<Tab, ROW + 1>(d, egrid, eblock, stream);
	}
}

// extra streams per device for the interleaved sub-ensembles (created once; read-only afterwards)
constexpr int MAX_CHAINS = 4;
static cudaStream_t side_stream(int which)
{
	static std::atomic<cudaStream_t> streams[64][MAX_CHAINS];
	int device = 0;
	if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 64 || which < 0 || which >= MAX_CHAINS) return nullptr;
	cudaStream_t stream = streams[device][which].load(std::memory_order_acquire);
	if (!stream) {
		if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
		cudaStream_t expected = nullptr;
		if (!streams[device][which].compare_exchange_strong(expected, stream)) {      // another thread was faster
			cudaStreamDestroy(stream);
			stream = expected;
		}
	}
	return stream;
}

// One of the interleaved sub-ensembles: its column range, stream, grids, and the host's view of its progress.
struct Chain {
	Dense d;
	cudaStream_t stream;
	cudaEvent_t polled;
	int32_t *host_flag;
	bool alive;
	int64_t rounds;
	dim3 egrid;             // element-wise kernels: 32 trajectories × 64 components per block
	unsigned tgrid;         // per-trajectory kernels: 256 trajectories per block
	void reset_count() const { cudaMemsetAsync(d.active_count, 0, sizeof(int32_t), stream); }
};
static const dim3 eblock(32, 8);

// Large ensembles are integrated as SEVERAL sub-ensembles whose launch sequences are interleaved on as many streams: the
// step of one is a chain of dependent kernels — element-wise stage kernel (memory-bound), GEMM (tensor-bound), stage
// kernel, … — so with a second, independent chain in flight the stage kernels of one run under the GEMM of the other
// and the partly filled last wave of a GEMM's blocks is topped up.  The host polls one chain while the other has a
// round queued.  start / round / finish queue the kernels of one chain (they end with the chain's counter of unfinished
// trajectories up to date); this function does the polling, the forking and joining of the streams.
template <class Start, class Round, class Finish>
static int run_chains(const jb_dense_args *x, const Dense &d, Start &&launch_start, Round &&launch_round, Finish &&launch_finish)
{
	// ---- the sub-ensembles: equal parts of ≥ 1024 trajectories each (bits 9-11 of method_flags: that many chains; 0: as
	// many as fit, at most four — measured on config 5: 1: 234.7 ms, 2: 224.1, 3: 216.9, 4: 215.9 per ten samples)
	Chain chains[MAX_CHAINS];
	cudaStream_t main_stream = (cudaStream_t) x->stream;
	const int wanted = (x->method_flags >> 9) & 7;
	int n_chains = wanted ? wanted : MAX_CHAINS;
	if (n_chains > MAX_CHAINS) n_chains = MAX_CHAINS;
	while (n_chains > 1 && d.B < 1024 * (int64_t) n_chains) --n_chains;
	for (int c = 1; c < n_chains; ++c)
		if (!side_stream(c)) n_chains = 1;
	const int64_t part = n_chains > 1 ? (d.B / n_chains + 63) / 64 * 64 : d.ld;
	for (int c = 0; c < n_chains; ++c) {
		Chain &chain = chains[c];
		chain.d = d;
		chain.d.b0 = c * part;
		chain.d.b1 = c == n_chains - 1 ? d.ld : (c + 1) * part;
		if (chain.d.B > chain.d.b1) chain.d.B = chain.d.b1;       // its kernels must not touch the next chain's trajectories
		chain.d.active_count = x->scratch + x->ld / 32 + 1 + c;
		chain.stream = c == 0 ? main_stream : side_stream(c);
		chain.host_flag = x->host_flag + c;
		chain.alive = true;
		chain.rounds = 0;
		const int64_t width = chain.d.b1 - chain.d.b0;
		chain.egrid = dim3((unsigned) (width / 32), (unsigned) ((d.n + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK));
		chain.tgrid = (unsigned) ((width + 255) / 256);
		if (cudaEventCreateWithFlags(&chain.polled, cudaEventDisableTiming) != cudaSuccess)
			return fail((int) cudaGetLastError(), "jb_dense_integrate: creating an event");
	}
	cudaError_t err = cudaSuccess;
	for (int c = 1; c < n_chains && err == cudaSuccess; ++c) {
		// the extra streams join the caller's: they start after what the caller has queued so far
		err = cudaEventRecord(chains[c].polled, main_stream);
		if (err == cudaSuccess) err = cudaStreamWaitEvent(chains[c].stream, chains[c].polled, 0);
	}
	// queue the copy of the chain's counter of unfinished trajectories and mark the point to wait for
	auto queue_poll = [&](Chain &chain) {
		cudaMemcpyAsync(chain.host_flag, chain.d.active_count, sizeof(int32_t), cudaMemcpyDeviceToHost, chain.stream);
		cudaEventRecord(chain.polled, chain.stream);
	};
	for (int c = 0; c < n_chains; ++c) {
		launch_start(chains[c]);
		queue_poll(chains[c]);
	}
	int alive = n_chains;
	while (alive > 0 && err == cudaSuccess) {
		for (int c = 0; c < n_chains && err == cudaSuccess; ++c) {
			Chain &chain = chains[c];
			if (!chain.alive) continue;
			err = cudaEventSynchronize(chain.polled);       // (the other chains have their rounds queued meanwhile)
			if (err != cudaSuccess) break;
			if (*chain.host_flag == 0) {
				launch_finish(chain);
				chain.alive = false;
				--alive;
			} else if (chain.rounds > d.nmax + 8) {
				err = cudaErrorUnknown;
			} else {
				launch_round(chain);
				++chain.rounds;
				queue_poll(chain);
			}
		}
	}
	for (int c = 1; c < n_chains && err == cudaSuccess; ++c) {
		// the caller's stream continues after the other chains have finished
		err = cudaEventRecord(chains[c].polled, chains[c].stream);
		if (err == cudaSuccess) err = cudaStreamWaitEvent(main_stream, chains[c].polled, 0);
	}
	int64_t rounds = 0;
	for (int c = 0; c < n_chains; ++c) {
		rounds = chains[c].rounds > rounds ? chains[c].rounds : rounds;
		cudaEventDestroy(chains[c].polled);
	}
	if (x->rounds_out) *x->rounds_out = rounds;
	if (err == cudaSuccess) err = cudaGetLastError();
	return err == cudaSuccess ? 0 : fail((int) err, "jb_dense_integrate");
}

// checks the argument block and fills the kernels' view of it; 0 or an error code
template <class Tab>
static int describe(const jb_dense_args *x, Dense &d)
{
	if (!x || x->struct_size != sizeof(jb_dense_args)) return jbd::fail(JB_E_ABI, "jb_dense_integrate: argument block has the wrong size");
	if (x->B <= 0) return 0;
	if (x->ld < x->B || x->ld % 64 || x->n <= 0 || !x->W || !x->omega || !x->t || !x->Y || !x->work || !x->ctrl
			|| !x->status || !x->n_accepted || !x->n_rejected || !x->scratch || !x->host_flag)
		return jbd::fail(JB_E_BAD_ARGUMENT, "jb_dense_integrate: missing or inconsistent buffers");
	d.B = x->B; d.ld = x->ld; d.n = x->n; d.W = x->W; d.omega = x->omega; d.coupling = x->coupling; d.atol_vec = x->atol_vec;
	d.T = x->T; d.rtol = x->rtol; d.atol = x->atol; d.max_step = x->max_step; d.first_step = x->first_step;
	d.nmax = x->nsteps > 0 ? x->nsteps : 100000;
	d.safety = x->safety > 0.0 ? x->safety : 0.9;
	// scipy.integrate.ode's defaults for the two codes (scipy _ode.py:1112-1236)
	constexpr bool dp5 = Tab::ERR_NEEDS_FSAL;
	d.facc1 = 1.0 / (x->dfactor > 0.0 ? x->dfactor : (dp5 ? 0.2 : 0.3));
	d.facc2 = 1.0 / (x->ifactor > 0.0 ? x->ifactor : (dp5 ? 10.0 : 6.0));
	if (dp5) {
		d.beta = x->beta == 0.0 ? 0.04 : (x->beta < 0.0 ? 0.0 : x->beta);
		d.expo1 = 0.2 - d.beta * 0.75;
	} else {
		d.beta = x->beta <= 0.0 ? 0.0 : x->beta;
		d.expo1 = 1.0 / 8.0 - d.beta * 0.2;
	}
	d.reject_by_dfactor = (x->method_flags >> 8) & 1;
	d.driver = (x->method_flags >> 12) & 3;
	d.t = x->t; d.Y = x->Y; d.work = x->work; d.ctrl = x->ctrl; d.status = x->status;
	d.n_accepted = x->n_accepted; d.n_rejected = x->n_rejected;
	d.tile_skip = x->scratch;
	return 0;
}

// advance every trajectory to x->T (Hairer's DOPRI5 / DOP853, fresh start, lands exactly on T)
template <class Tab>
static int integrate(const jb_dense_args *x)
{
	jbd::Dense d{};
	if (const int code = describe<Tab>(x, d)) return code;
	if (x->B <= 0) return 0;
	const int have_probe = !(d.first_step > 0.0);

	return run_chains(x, d,
		// ---- fresh start: f(t, Y), HINIT
		[&](Chain &chain) {
			const Dense &sub = chain.d;
			cudaStream_t stream = chain.stream;
			chain.reset_count();
			JBD_LAUNCH(jbd::setup_kernel, chain.tgrid, 256, sub);
			JBD_LAUNCH(jbd::sincos_kernel<0>, chain.egrid, eblock, sub, (int) jbd::F_ACTIVE);
			jbd::gemm(sub, stream);
			JBD_LAUNCH(jbd::start_kernel, chain.egrid, eblock, sub, (int) jbd::F_ACTIVE);
			if (have_probe) {
				JBD_LAUNCH(jbd::probe_kernel, chain.egrid, eblock, sub, (int) jbd::F_ACTIVE);
				jbd::gemm(sub, stream);
				JBD_LAUNCH(jbd::probe_finish_kernel, chain.egrid, eblock, sub, (int) jbd::F_ACTIVE);
			}
			chain.reset_count();
			JBD_LAUNCH(jbd::first_step_kernel<Tab>, chain.tgrid, 256, sub, have_probe);
		},
		// ---- a round: one step attempt of every unfinished trajectory of the chain
		[&](Chain &chain) {
			const Dense &sub = chain.d;
			cudaStream_t stream = chain.stream;
			launch_stages<Tab>(sub, chain.egrid, eblock, stream);
			JBD_LAUNCH(jbd::error_kernel<Tab>, chain.egrid, eblock, sub);
			chain.reset_count();
			JBD_LAUNCH(jbd::control_kernel<Tab>, chain.tgrid, 256, sub);
		},
		// ---- the accepted trajectories of the last round still have y_new in Z
		[&](Chain &chain) {
			const Dense &sub = chain.d;
			cudaStream_t stream = chain.stream;
			if (chain.rounds == 0) return;
			JBD_LAUNCH(jbd::commit_kernel, chain.egrid, eblock, sub);
			JBD_LAUNCH(jbd::clear_accepted_kernel, chain.tgrid, 256, sub);
		});
}

// the same for SciPy's solve_ivp stepper (RK45 / DOP853): persistent between calls; with dense output (JB_DRV_IVP_DENSE)
// the sample at T goes to work vector jb_dense_result_vector() and Y stays the solver's state
template <class Tab>
static int integrate_ivp(const jb_dense_args *x)
{
	jbd::Dense d{};
	if (const int code = describe<Tab>(x, d)) return code;
	if (x->B <= 0) return 0;
	d.nmax = (int64_t) 1 << 26;       // solve_ivp has no step limit; a call must end nevertheless
	const int have_probe = !(d.first_step > 0.0);
	const int fresh = (x->method_flags >> 14) & 1;      // some stepper may have to be created (the caller set a new state)

	return run_chains(x, d,
		[&](Chain &chain) {
			const Dense &sub = chain.d;
			cudaStream_t stream = chain.stream;
			chain.reset_count();
			JBD_LAUNCH(jbd::setup_ivp_kernel, chain.tgrid, 256, sub);
			if (fresh) {
				// RungeKutta.__init__: f = fun(t, y), select_initial_step (the Euler probe of common.py:119-122)
				JBD_LAUNCH(jbd::sincos_kernel<0>, chain.egrid, eblock, sub, (int) jbd::F_INIT);
				jbd::gemm(sub, stream);
				JBD_LAUNCH(jbd::start_kernel, chain.egrid, eblock, sub, (int) jbd::F_INIT);
				if (have_probe) {
					JBD_LAUNCH(jbd::probe_kernel, chain.egrid, eblock, sub, (int) jbd::F_INIT);
					jbd::gemm(sub, stream);
					JBD_LAUNCH(jbd::probe_finish_kernel, chain.egrid, eblock, sub, (int) jbd::F_INIT);
				}
			}
			chain.reset_count();
			JBD_LAUNCH(jbd::first_step_ivp_kernel<Tab>, chain.tgrid, 256, sub, have_probe);
		},
		[&](Chain &chain) {
			const Dense &sub = chain.d;
			cudaStream_t stream = chain.stream;
			launch_stages<Tab>(sub, chain.egrid, eblock, stream);
			JBD_LAUNCH(jbd::error_kernel<Tab>, chain.egrid, eblock, sub);
			chain.reset_count();
			JBD_LAUNCH(jbd::control_ivp_kernel<Tab>, chain.tgrid, 256, sub);
		},
		[&](Chain &chain) {
			const Dense &sub = chain.d;
			cudaStream_t stream = chain.stream;
			if constexpr (!Tab::ERR_NEEDS_FSAL) {
				if (sub.driver == JB_DRV_IVP_DENSE && chain.rounds > 0) {
					JBD_LAUNCH(jbd::mark_accepted_tiles_kernel, chain.tgrid, 256, sub);
					JBD_LAUNCH(jbd::extra_stage_kernel<13>, chain.egrid, eblock, sub);
					jbd::gemm(sub, stream);
					JBD_LAUNCH(jbd::extra_stage_kernel<14>, chain.egrid, eblock, sub);
					jbd::gemm(sub, stream);
					JBD_LAUNCH(jbd::extra_stage_kernel<15>, chain.egrid, eblock, sub);
					jbd::gemm(sub, stream);
				}
			}
			JBD_LAUNCH(jbd::commit_ivp_kernel<Tab>, chain.egrid, eblock, sub);
			JBD_LAUNCH(jbd::clear_accepted_kernel, chain.tgrid, 256, sub);
		});
}

}  // namespace jbd

extern "C" {

int jb_dense_integrate(const jb_dense_args *x)
{
	if (!x || x->struct_size != sizeof(jb_dense_args)) return jbd::fail(JB_E_ABI, "jb_dense_integrate: argument block has the wrong size");
	const bool dop853 = (x->method_flags & 0xff) == JB_DOP853;
	if (((x->method_flags >> 12) & 3) != JB_DRV_ODE) return dop853 ? jbd::integrate_ivp<jbd::DOP853>(x) : jbd::integrate_ivp<jbd::DP5>(x);
	return dop853 ? jbd::integrate<jbd::DOP853>(x) : jbd::integrate<jbd::DP5>(x);
}

}  // extern "C"
