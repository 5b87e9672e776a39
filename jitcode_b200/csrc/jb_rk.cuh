// Ensemble adaptive Runge–Kutta kernels, one thread per trajectory.
//
// What is computed follows the reference's integrators exactly:
//   JB_DRV_ODE        Hairer's DOPRI5 / DOP853 core (HINIT + DOPCOR / DP86CO) as scipy.integrate.ode
//                     configures it (scipy _ode.py:1112-1236) and as ODE_wrapper.integrate drives it
//                     (integrator_tools.py:134-144): fresh start per sample time, lands exactly on it.
//   JB_DRV_IVP_DENSE  scipy solve_ivp's RungeKutta._step_impl (scipy _ivp/rk.py:111-176) with
//                     select_initial_step (common.py:68-134), driven like IVP_wrapper.integrate
//                     (integrator_tools.py:98-105): persistent stepper, t_bound = ∞, dense output.
//   JB_DRV_IVP_LAND   same stepper driven like IVP_wrapper_no_interpolation (:110-128).
// followed, per sample time, by the optional Lyapunov post-step of _jitcode.py:743-775 / :886-893 /
// :937-946.  oracle/rk_numpy.py is the line-by-line CPU statement of the same arithmetic.
//
// How it is computed is B200-first: the generated right-hand side (Sys::rhs) is fused into the
// stepper; a trajectory never leaves its thread between the first and the last sample time of a
// launch; memory is laid out in trajectory-minor tiles of 32 (one tile = one warp), so that every
// global access of a warp is one contiguous 256-byte segment at a compile-time offset from a
// per-thread base pointer (no address arithmetic in the generated code).
//   JB_STORE_REGISTERS  thread per trajectory; y and all stages live in registers for the whole launch
//                       (small systems, FP64-pipe bound);
//   JB_STORE_GLOBAL     thread per trajectory; they live in a caller-provided tiled workspace
//                       (mid-size or irregular systems, HBM/L2 bound); buffers rotate by pointer swap;
//   JB_STORE_WARP       WARP per trajectory, lanes across components; y and all stages live in shared
//                       memory for the whole launch, the right-hand side is the vectorised form printed
//                       by _vectorise.py, norms are shuffle reductions, and every control decision is
//                       warp-uniform — no divergence between trajectories and no HBM traffic between
//                       the first and the last sample time (large regular networks).
#pragma once

#include <cstdint>
#include <cmath>
#include <type_traits>
#include "jb_tableaux.cuh"
#include "jitcode_b200.h"

namespace jb {

constexpr int TILE = JB_TILE;

// ---------------------------------------------------------------------------------------------
// vectors: the generated code reads `y(i)` and writes `dy.set(i, v)`

template <int N>
struct RVec {               // thread-private, scalarised into registers
	double v[N > 0 ? N : 1];
	__device__ __forceinline__ double operator()(int i) const { return v[i]; }
	__device__ __forceinline__ void set(int i, double x) { v[i] = x; }
};

template <int STRIDE>
struct StrideVec {          // element i at p[i*STRIDE]: consecutive threads touch consecutive words
	double *p;
	__device__ __forceinline__ double operator()(int i) const { return p[i * STRIDE]; }
	__device__ __forceinline__ void set(int i, double x) const { p[i * STRIDE] = x; }
};
using GVec = StrideVec<TILE>;       // one trajectory's column of a tile in global memory

// Quad storage: what lane k of a trajectory's group sees of a caller-visible (tiled) vector of the full Lyapunov
// system — local component i < NB is state component i (every lane reads it, lane 0 writes it), local component
// NB + i is component NB·(1 + k) + i, i.e. its own tangent vector.  Lanes beyond the last tangent vector ("phantom"
// lanes of a group padded to a power of two) read zeros there and write nothing.
template <int NB>
struct QVec {
	double *p;
	int shift;          // NB · k
	bool first, real;   // lane 0 of the group / a lane that owns a tangent vector
	__device__ __forceinline__ double operator()(int i) const
	{
		if (i < NB) return p[i * TILE];
		return real ? p[(i + shift) * TILE] : 0.0;
	}
	__device__ __forceinline__ void set(int i, double x) const
	{
		if (i < NB) { if (first) p[i * TILE] = x; }
		else if (real) p[(i + shift) * TILE] = x;
	}
};

struct SVec {               // contiguous vector: shared memory (warp storage) or one row of a (B, n) global array
	double *p;
	__device__ __forceinline__ double operator()(int i) const { return p[i]; }
	// element at a byte offset (the vectorised right-hand side keeps pre-scaled offsets in its tables)
	__device__ __forceinline__ double at(unsigned bytes) const
	{
		return *reinterpret_cast<const double *>(reinterpret_cast<const char *>(p) + bytes);
	}
	__device__ __forceinline__ void set(int i, double x) const { p[i] = x; }
};

// caller-visible vector of a warp-storage module: position q of the kernel's vectors holds component comp[q]
template <class TU>
struct PVec {
	double *p;
	const TU *comp;
	__device__ __forceinline__ double operator()(int q) const { return p[comp[q]]; }
	__device__ __forceinline__ void set(int q, double x) const { p[comp[q]] = x; }
};

// v with its sign flipped where bit 31 of `word` is set (signed entries of the vectoriser's gather tables)
__device__ __forceinline__ double flip_sign(double v, unsigned word)
{
	return __hiloint2double(__double2hiint(v) ^ (int) (word & 0x80000000u), __double2loint(v));
}

// value of `v` in lane `source` of the warp (all lanes take part)
__device__ __forceinline__ double shfl(double v, unsigned source) { return __shfl_sync(0xffffffffu, v, source); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int offset = 16; offset > 0; offset >>= 1) v += __shfl_xor_sync(0xffffffffu, v, offset);
	return v;
}

// Σ over the G = 2^k lanes of a group (quad storage); only the group's own lanes take part, so groups of one warp
// may be in different places of the code
template <int G>
__device__ __forceinline__ double group_sum(double v)
{
	if constexpr (G > 1) {
		const unsigned lane = threadIdx.x & 31u;
		const unsigned mask = (G >= 32 ? 0xffffffffu : ((1u << G) - 1u)) << (lane & ~(unsigned) (G - 1));
#pragma unroll
		for (int offset = G / 2; offset > 0; offset >>= 1) v += __shfl_xor_sync(mask, v, offset);
	}
	return v;
}
template <int G>
__device__ __forceinline__ double group_broadcast(double v, int source)
{
	if constexpr (G > 1) {
		const unsigned lane = threadIdx.x & 31u;
		const unsigned mask = (G >= 32 ? 0xffffffffu : ((1u << G) - 1u)) << (lane & ~(unsigned) (G - 1));
		v = __shfl_sync(mask, v, source, G);
	}
	return v;
}

// A team: the W warps (W = 1, 2, 4, 8) that integrate ONE trajectory together in warp storage, lanes across
// components.  One warp suffices while a trajectory's rows leave room for many trajectories per SM; for larger
// systems several warps share the rows, so that the SM still has enough warps in flight.
template <int W>
struct Team {
	static constexpr int SIZE = 32 * W;
	static __device__ __forceinline__ int rank() { return W == 1 ? (threadIdx.x & 31) : (threadIdx.x % SIZE); }
	static __device__ __forceinline__ int index() { return threadIdx.x / SIZE; }      // team within the block
	static __device__ __forceinline__ void sync()
	{
		if constexpr (W == 1) __syncwarp();
		else asm volatile("bar.sync %0, %1;" :: "r"(1 + index()), "r"(SIZE) : "memory");   // named barrier per team
	}
	// the same sum in five steps that the caller may spread over other work (a single warp: step K of the butterfly;
	// the additions are those of sum(), in the same order)
	template <int K> static __device__ __forceinline__ double butterfly(double v, double *scratch)
	{
		if constexpr (W == 1) return v + __shfl_xor_sync(0xffffffffu, v, 16 >> K);
		else return K == 0 ? sum(v, scratch) : v;
	}
	// Σ over the team, known to all its threads; `scratch`: W doubles of shared memory owned by the team
	static __device__ __forceinline__ double sum(double v, double *scratch)
	{
		v = warp_sum(v);
		if constexpr (W > 1) {
			if ((threadIdx.x & 31) == 0) scratch[(threadIdx.x % SIZE) >> 5] = v;
			sync();
			v = 0.0;
#pragma unroll
			for (int w = 0; w < W; ++w) v += scratch[w];
			sync();
		}
		return v;
	}
};

// dst takes the value of src; src becomes scratch (registers: copy; global: pointer swap)
template <int N>
__device__ __forceinline__ void take(RVec<N> &dst, RVec<N> &src)
{
#pragma unroll
	for (int i = 0; i < N; ++i) dst.v[i] = src.v[i];
}
template <int STRIDE>
__device__ __forceinline__ void take(StrideVec<STRIDE> &dst, StrideVec<STRIDE> &src)
{
	double *const p = dst.p;
	dst.p = src.p;
	src.p = p;
}
__device__ __forceinline__ void take(SVec &dst, SVec &src)
{
	double *const p = dst.p;
	dst.p = src.p;
	src.p = p;
}

// position of trajectory b in a tiled array of `rows` rows
__device__ __forceinline__ int64_t tile_offset(int64_t b, int rows)
{
	return (b / TILE) * ((int64_t) rows * TILE) + (b % TILE);
}

// ---------------------------------------------------------------------------------------------

// The coefficients of the hot loops (stage combinations, error estimates) once more, in the constant bank: a 64-bit
// literal costs two UMOV per use (23 % of all issued instructions of the Lotka–Volterra kernel were UMOV), a constant
// costs one LDCU.64.  Measured (scripts/ab_variants.py … -DJB_LITERAL_COEF): Lotka–Volterra +3.3 %, the same under
// RK45 +3.4 %, FHN-Lyapunov +2.9 %, Rössler network +3.0 %.
#define JB_ROW4(M, s) { M[s][0], M[s][1], M[s][2], M[s][3] }
#define JB_ROW7(M, s) { M[s][0], M[s][1], M[s][2], M[s][3], M[s][4], M[s][5], M[s][6] }
#define JB_ROW16(M, s) { M[s][0], M[s][1], M[s][2], M[s][3], M[s][4], M[s][5], M[s][6], M[s][7], M[s][8], M[s][9], M[s][10], M[s][11], M[s][12], M[s][13], M[s][14], M[s][15] }
namespace bank {
__constant__ double dp5_a[7][7] = { JB_ROW7(DP5::A, 0), JB_ROW7(DP5::A, 1), JB_ROW7(DP5::A, 2), JB_ROW7(DP5::A, 3), JB_ROW7(DP5::A, 4), JB_ROW7(DP5::A, 5), JB_ROW7(DP5::A, 6) };
__constant__ double dp5_e[7] = { DP5::E[0], DP5::E[1], DP5::E[2], DP5::E[3], DP5::E[4], DP5::E[5], DP5::E[6] };
__constant__ double rk23_a[4][4] = { JB_ROW4(RK23::A, 0), JB_ROW4(RK23::A, 1), JB_ROW4(RK23::A, 2), JB_ROW4(RK23::A, 3) };
__constant__ double rk23_e[4] = { RK23::E[0], RK23::E[1], RK23::E[2], RK23::E[3] };
__constant__ double dop853_a[13][16] = {
	JB_ROW16(dop853::A, 0), JB_ROW16(dop853::A, 1), JB_ROW16(dop853::A, 2), JB_ROW16(dop853::A, 3), JB_ROW16(dop853::A, 4),
	JB_ROW16(dop853::A, 5), JB_ROW16(dop853::A, 6), JB_ROW16(dop853::A, 7), JB_ROW16(dop853::A, 8), JB_ROW16(dop853::A, 9),
	JB_ROW16(dop853::A, 10), JB_ROW16(dop853::A, 11), JB_ROW16(dop853::A, 12) };
__constant__ double dop853_e5[13] = { dop853::E5[0], dop853::E5[1], dop853::E5[2], dop853::E5[3], dop853::E5[4], dop853::E5[5], dop853::E5[6],
	dop853::E5[7], dop853::E5[8], dop853::E5[9], dop853::E5[10], dop853::E5[11], dop853::E5[12] };
__constant__ double dop853_e3[13] = { dop853::E3[0], dop853::E3[1], dop853::E3[2], dop853::E3[3], dop853::E3[4], dop853::E3[5], dop853::E3[6],
	dop853::E3[7], dop853::E3[8], dop853::E3[9], dop853::E3[10], dop853::E3[11], dop853::E3[12] };
}  // namespace bank
#undef JB_ROW4
#undef JB_ROW7
#undef JB_ROW16

template <class Tab> struct Coef;
// a(s, j), c(s): constant expressions (which coefficients vanish is decided at compile time);
// a_bank(s, j): the same number as an operand from the constant bank (rows 0 … NSTAGE)
template <> struct Coef<DP5> {
	static __host__ __device__ constexpr double a(int s, int j) { return DP5::A[s][j]; }
	static __host__ __device__ constexpr double c(int s) { return DP5::C[s]; }
	static __device__ __forceinline__ double a_bank(int s, int j) { return bank::dp5_a[s][j]; }
};
template <> struct Coef<DOP853> {
	static __host__ __device__ constexpr double a(int s, int j) { return dop853::A[s][j]; }
	static __host__ __device__ constexpr double c(int s) { return dop853::C[s]; }
	static __device__ __forceinline__ double a_bank(int s, int j) { return bank::dop853_a[s][j]; }
};
template <> struct Coef<RK23> {
	static __host__ __device__ constexpr double a(int s, int j) { return RK23::A[s][j]; }
	static __host__ __device__ constexpr double c(int s) { return RK23::C[s]; }
	static __device__ __forceinline__ double a_bank(int s, int j) { return bank::rk23_a[s][j]; }
};
#ifdef JB_LITERAL_COEF
#define JB_COEF(literal, banked) (literal)
#else
#define JB_COEF(literal, banked) (banked)
#endif

struct Params {             // kernel argument block (by value)
	int64_t B, ld;
	int32_t n_times, post, n_proj, lyap_discard, hairer_reject_by_dfactor;
	const double *times;
	double rtol, atol;
	const double *atol_vec;
	double max_step, first_step;
	int64_t nsteps;
	double safety, facc1, facc2, beta, expo1;
	double dfactor, ifactor;       // = 1/facc1, 1/facc2: the limits of h_new/h (the step-size rules below multiply, never divide)
	double *t, *Y;
	const double *P;
	double *state, *sstate, *work, *Yres, *Yrec, *lyap_rec, *lyap_acc;
	const double *proj;
	int32_t *status;
	int64_t *n_accepted, *n_rejected;
	unsigned long long *queue;     // caller-owned counter from which persistent kernels draw trajectories, or null
};

// persistent scalars of the IVP drivers, sstate[k*ld + b]
enum { SS_H_ABS = 0, SS_T_OLD = 1, SS_H_PREV = 2, SS_FLAGS = 3, SS_COUNT = 4 };
enum { FLAG_INITIALISED = 1, FLAG_INTERPOLANT = 2 };
// persistent vectors of the IVP drivers (tiled, n rows each): f(t,y), y_old, interpolant
enum { SV_F = 0, SV_YOLD = 1, SV_DENSE0 = 2 };

template <class Tab, int DRIVER> struct Layout {
	static constexpr bool DENSE = DRIVER == JB_DRV_IVP_DENSE;
	static constexpr bool IVP = DRIVER != JB_DRV_ODE;
	static constexpr int KR = DENSE ? Tab::KROWS_DENSE : Tab::KROWS;
	static constexpr int STATE_VECTORS = DENSE ? 2 + Tab::NDENSE : (IVP ? 1 : 0);
	static constexpr int STATE_SCALARS = IVP ? SS_COUNT : 0;
	static constexpr int WORK_VECTORS = 1 + KR + (DENSE ? 1 : 0);   // Z, K rows, y_old
};

// ---------------------------------------------------------------------------------------------
// calling the generated right-hand side

// global / shared: one out-of-line copy of the (large) right-hand side
template <class Sys, int STRIDE = TILE>
__device__ __noinline__ void rhs_outlined(double t, const double *__restrict__ in, double *__restrict__ out,
		const typename Sys::Par par)
{
	const StrideVec<STRIDE> vi{const_cast<double *>(in)};
	StrideVec<STRIDE> vo{out};
	Sys::rhs(t, vi, vo, par);
}
// warp storage: one out-of-line copy of the vectorised right-hand side; threads read what other threads of their
// team wrote, hence the team barriers around it
template <class Sys>
__device__ __noinline__ void rhs_warp_outlined(double t, const double *__restrict__ in, double *__restrict__ out,
		const typename Sys::Par par, const void *tables, double *scratch)
{
	using T = Team<Sys::TEAM_WARPS>;
	T::sync();
	const SVec vi{const_cast<double *>(in)};
	SVec vo{out};
	Sys::rhs_warp(t, vi, vo, par, T::rank(), tables, scratch);
	T::sync();
}

// ---------------------------------------------------------------------------------------------
// Storage policies.  Common interface used by the stepper and by run_trajectory:
//   for_each(body(i, r)), sum(value(i, r)), sum2(value(i, r, s0, s1))     component loops / reductions
//   y, z, yo, k<J> (i, r)  and  set_y, set_z, set_yo, set_k<J> (i, r, v)  element access
//   rhs_y<J>(t), rhs_z<J>(t)      K[J] = f(t, Y) / f(t, Z)
//   accept_y()                    Y ← Z            (Z becomes scratch)
//   save_old()                    y_old ← Y        (before accept_y, dense driver)
//   fsal<JD, JS>()                K[JD] ← K[JS]    (K[JS] becomes scratch)
// `i` is the position of the component in the vector, `r` the pass of the lane-strided loop (a
// compile-time constant where rows are cached in registers, otherwise a dummy).

template <class Sys, class Tab, int DRIVER, int MODE>
struct ThreadStore {        // one thread per trajectory: registers (MODE 0), the caller's tiled workspace (MODE 1)
	                            // or the block's shared memory (MODE 3: rows of one word per thread);
	                            // quad storage (MODE 5): a group of lanes per trajectory, each with the state and ONE
	                            // tangent vector in registers (Sys::rhs is then the system with a single tangent vector)
	using L = Layout<Tab, DRIVER>;
	static constexpr bool QUAD = MODE == JB_STORE_QUAD;
	static constexpr bool REG = MODE == JB_STORE_REGISTERS || QUAD;
	static constexpr bool SHARED = MODE == JB_STORE_SHARED;
	static constexpr bool WARP = false;
	static constexpr int N = QUAD ? 2 * Sys::NB : Sys::N;      // length of this thread's vectors
	static constexpr int GROUP = QUAD ? Sys::GROUP : 1;         // lanes per trajectory
	static constexpr int NSPLIT = QUAD ? Sys::NB : N;           // components [NSPLIT, N) are summed over the group
	static constexpr int KR = L::KR;
	static constexpr int UNR = REG ? (N > 0 ? N : 1) : 4;   // unroll factor of component loops
	static constexpr int STRIDE = SHARED ? Sys::THREADS : TILE;
	static constexpr int ROWS = 2 + KR + (L::DENSE ? 1 : 0);    // shared storage: Y, Z, K rows, y_old
	using Vec = typename std::conditional<REG, RVec<N>, StrideVec<STRIDE>>::type;
	using IVec = typename std::conditional<QUAD, QVec<Sys::NB>, GVec>::type;      // persistent solver vectors in the caller's memory
	using UVec = IVec;      // the caller's state and samples

	Vec Y, Z, YO, K[KR];
	typename Sys::Par par;

	template <class F> __device__ __forceinline__ void for_each(F &&body)
	{
#pragma unroll UNR
		for (int i = 0; i < N; ++i) body(i, Idx<0>{});
	}
	// quad storage: every lane of the group holds the state's contribution, the tangent vectors' are added up over
	// the group (butterfly: all lanes get bit-identical totals, so the group's control flow stays uniform)
	template <class F> __device__ __forceinline__ double sum(F &&value)
	{
		double total = 0.0;
#pragma unroll UNR
		for (int i = 0; i < NSPLIT; ++i) total += value(i, Idx<0>{});
		if constexpr (QUAD) {
			double own = 0.0;
#pragma unroll UNR
			for (int i = NSPLIT; i < N; ++i) own += value(i, Idx<0>{});
			total += group_sum<GROUP>(own);
		}
		return total;
	}
	template <class F> __device__ __forceinline__ void sum2(F &&value, double &s0, double &s1)
	{
		s0 = s1 = 0.0;
#pragma unroll UNR
		for (int i = 0; i < NSPLIT; ++i) value(i, Idx<0>{}, s0, s1);
		if constexpr (QUAD) {
			double o0 = 0.0, o1 = 0.0;
#pragma unroll UNR
			for (int i = NSPLIT; i < N; ++i) value(i, Idx<0>{}, o0, o1);
			s0 += group_sum<GROUP>(o0);
			s1 += group_sum<GROUP>(o1);
		}
	}

	template <class R> __device__ __forceinline__ double y(int i, R) const { return Y(i); }
	template <class R> __device__ __forceinline__ double z(int i, R) const { return Z(i); }
	template <class R> __device__ __forceinline__ double yo(int i, R) const { return YO(i); }
	template <int J, class R> __device__ __forceinline__ double k(int i, R) const { return K[J](i); }
	template <class R> __device__ __forceinline__ void set_y(int i, R, double v) { Y.set(i, v); }
	template <class R> __device__ __forceinline__ void set_z(int i, R, double v) { Z.set(i, v); }
	template <class R> __device__ __forceinline__ void set_yo(int i, R, double v) { YO.set(i, v); }
	template <int J, class R> __device__ __forceinline__ void set_k(int i, R, double v) { K[J].set(i, v); }

	template <int J> __device__ __forceinline__ void call(double t, const Vec &in)
	{
		if constexpr (REG) Sys::rhs(t, in, K[J], par);
		else rhs_outlined<Sys, STRIDE>(t, in.p, K[J].p, par);
	}
	template <int J> __device__ __forceinline__ void rhs_y(double t) { call<J>(t, Y); }
	template <int J> __device__ __forceinline__ void rhs_z(double t) { call<J>(t, Z); }
	__device__ __forceinline__ void accept_y() { take(Y, Z); }
	__device__ __forceinline__ void save_old() { take(YO, Y); }
	template <int JD, int JS> __device__ __forceinline__ void fsal() { take(K[JD], K[JS]); }
};

template <class Sys, class Tab, int DRIVER, int KREGS>
struct WarpStore {          // one team of warps per trajectory, lanes across components
	using T = Team<Sys::TEAM_WARPS>;
	static constexpr int N = Sys::N;
	static constexpr int NROW = Sys::NROW;
	using L = Layout<Tab, DRIVER>;
	static constexpr bool REG = false;
	static constexpr bool WARP = true;
	static constexpr bool QUAD = false;
	static constexpr int KR = L::KR;
	static constexpr bool CACHED = KREGS > 0;           // Y and K[0 … KREGS) live in registers, lane-strided
	static constexpr int NP = (N + T::SIZE - 1) / T::SIZE;   // passes of a loop strided over the team
	// The staging row through which the right-hand side fills a register row shares the row of K[S]: register rows
	// are only filled while K[S] (f at the end of the last step) is dead — it moves to K[0] before the next stage 1.
	static constexpr bool F_ALIASED = CACHED && KREGS <= Tab::NSTAGE;
	// shared-memory rows of one warp: Z, [staging row], [Y], [y_old], K[KREGS … KR)
	static constexpr int ROWS = 1 + (CACHED ? (F_ALIASED ? 0 : 1) : 1) + (L::DENSE ? 1 : 0) + (KR - KREGS);
	using IVec = SVec;
	using UVec = PVec<typename Sys::TU>;

	double ry[CACHED ? NP : 1];
	double rk[CACHED ? KREGS : 1][CACHED ? NP : 1];
	double *pY, *pZ, *pYO, *pF, *pK[KR];
	double *scratch;            // T::W doubles for team reductions
	typename Sys::Par par;
	const void *tables;

	__device__ __forceinline__ void bind(double *smem)
	{
		scratch = smem + NROW * ROWS;
		int row = 0;
		pZ = smem + NROW * row++;
		pY = pF = nullptr;
		if constexpr (!CACHED) pY = smem + NROW * row++;
		else if constexpr (!F_ALIASED) pF = smem + NROW * row++;
		pYO = L::DENSE ? smem + NROW * row++ : nullptr;
#pragma unroll
		for (int j = 0; j < KR; ++j) pK[j] = j < KREGS ? nullptr : smem + NROW * (row + j - KREGS);
		// position N … NROW-1 of every row holds 0: padding target of the vectorised loops
		for (int r = T::rank(); r < ROWS; r += T::SIZE)
			for (int i = N; i < NROW; ++i) smem[r * NROW + i] = 0.0;
	}

	template <class F> __device__ __forceinline__ void for_each(F &&body)
	{
		const int rank = T::rank();
		if constexpr (CACHED) {
			static_for<0, NP>([&](auto r) {
				const int i = rank + T::SIZE * r;
				if (T::SIZE * (r + 1) <= N || i < N) body(i, r);
			});
		} else {
			for (int i = rank; i < N; i += T::SIZE) body(i, Idx<0>{});
		}
	}
	template <class F> __device__ __forceinline__ double sum(F &&value)
	{
		double total = 0.0;
		for_each([&](int i, auto r) { total += value(i, r); });
		return T::sum(total, scratch);
	}
	template <class F> __device__ __forceinline__ void sum2(F &&value, double &s0, double &s1)
	{
		double a = 0.0, b = 0.0;
		for_each([&](int i, auto r) { value(i, r, a, b); });
		s0 = T::sum(a, scratch);
		s1 = T::sum(b, scratch);
	}

	template <class R> __device__ __forceinline__ double y(int i, R r) const
	{
		if constexpr (CACHED) return ry[r];
		else return pY[i];
	}
	template <class R> __device__ __forceinline__ double z(int i, R) const { return pZ[i]; }
	template <class R> __device__ __forceinline__ double yo(int i, R) const { return pYO[i]; }
	template <int J, class R> __device__ __forceinline__ double k(int i, R r) const
	{
		if constexpr (J < KREGS) return rk[J][r];
		else return pK[J][i];
	}
	template <class R> __device__ __forceinline__ void set_y(int i, R r, double v)
	{
		if constexpr (CACHED) ry[r] = v;
		else pY[i] = v;
	}
	template <class R> __device__ __forceinline__ void set_z(int i, R, double v) { pZ[i] = v; }
	template <class R> __device__ __forceinline__ void set_yo(int i, R, double v) { pYO[i] = v; }
	template <int J, class R> __device__ __forceinline__ void set_k(int i, R r, double v)
	{
		if constexpr (J < KREGS) rk[J][r] = v;
		else pK[J][i] = v;
	}

	template <int J> __device__ __forceinline__ void call(double t, const double *in)
	{
		if constexpr (J < KREGS) {
			double *const staging = F_ALIASED ? pK[Tab::NSTAGE] : pF;
			rhs_warp_outlined<Sys>(t, in, staging, par, tables, scratch);
			for_each([&](int i, auto r) { rk[J][r] = staging[i]; });
		} else {
			rhs_warp_outlined<Sys>(t, in, pK[J], par, tables, scratch);
		}
	}
	template <int J> __device__ __forceinline__ void rhs_z(double t) { call<J>(t, pZ); }
	template <int J> __device__ __forceinline__ void rhs_y(double t)
	{
		if constexpr (CACHED) {
			for_each([&](int i, auto r) { pZ[i] = ry[r]; });     // Z is scratch wherever f(t, Y) is needed
			call<J>(t, pZ);
		} else {
			call<J>(t, pY);
		}
	}
	__device__ __forceinline__ void accept_y()
	{
		if constexpr (CACHED) for_each([&](int i, auto r) { ry[r] = pZ[i]; });
		else { double *const p = pY; pY = pZ; pZ = p; }
	}
	__device__ __forceinline__ void save_old()
	{
		if constexpr (CACHED) for_each([&](int i, auto r) { pYO[i] = ry[r]; });
		else { double *const p = pYO; pYO = pY; pY = p; }
	}
	template <int JD, int JS> __device__ __forceinline__ void fsal()
	{
		if constexpr (JD >= KREGS && JS >= KREGS) { double *const p = pK[JD]; pK[JD] = pK[JS]; pK[JS] = p; }
		else for_each([&](int i, auto r) { set_k<JD>(i, r, k<JS>(i, r)); });
	}
};

// ---------------------------------------------------------------------------------------------
// the Runge–Kutta arithmetic, written once against the storage interface

template <class Sys, class Tab, int DRIVER, class Store>
struct Stepper : Store {
	static constexpr int N = Sys::N;
	using L = Layout<Tab, DRIVER>;
	static constexpr bool DENSE = L::DENSE;
	static constexpr bool IVP = L::IVP;
	static constexpr int S = Tab::NSTAGE;
	static constexpr bool WARP = Store::WARP;
	static constexpr bool MODE_QUAD = Store::QUAD;

	double atol_s, rtol;
	const double *atol_vec;
	const typename Sys::TU *comp;     // warp storage: component stored at each position
	int shift;                        // quad storage: offset of this lane's tangent vector in the full system

	__device__ __forceinline__ double atol(int i) const
	{
		if constexpr (WARP) return atol_vec ? atol_vec[comp[i]] : atol_s;
		else if constexpr (MODE_QUAD) return atol_vec ? atol_vec[i < Sys::NB ? i : i + shift] : atol_s;
		else return atol_vec ? atol_vec[i] : atol_s;
	}

	// Z = Y + h * Σ_{j<ROW} a(ROW,j) K[j]   (stage input for ROW < S, solution for ROW == S)
	template <int ROW>
	__device__ __forceinline__ void combine(double h)
	{
		this->for_each([&](int i, auto r) {
			double acc = 0.0;
			static_for<0, ROW>([&](auto j) {
				constexpr double a = Coef<Tab>::a(ROW, j);
				if constexpr (a != 0.0) acc = fma(JB_COEF(a, Coef<Tab>::a_bank(ROW, j)), this->template k<j>(i, r), acc);
			});
			this->set_z(i, r, fma(h, acc, this->y(i, r)));
		});
	}

	// What attempt() returns: the error norm err of the 8th-order pair, but err² of the pairs whose norm is a plain root
	// mean square — the controllers only need log(err) = ½·log(err²) and the comparison with 1, and a double-precision
	// square root is some twenty instructions per attempt.  sqrt(x) <= 1 ⇔ x <= 1 + 2⁻⁵² and sqrt(x) < 1 ⇔ x < 1 in
	// round-to-nearest, so the decisions are exactly those of the reference's err <= 1 (Hairer) and err < 1 (SciPy).
	static constexpr bool ERR_SQUARED = Tab::ERR_NEEDS_FSAL;
	static __device__ __forceinline__ bool accepted_hairer(double e) { return ERR_SQUARED ? e <= 1.0000000000000002 : e <= 1.0; }
	static __device__ __forceinline__ bool accepted_scipy(double e) { return e < 1.0; }
	static __device__ __forceinline__ double log_error(double e) { return ERR_SQUARED ? 0.5 * log(e) : log(e); }

	// one step attempt from (t, Y, K[0]) with step h: fills K[1..], Z = y_new; returns the error measure (see above)
	__device__ __forceinline__ double attempt(double t, double h)
	{
		static_for<1, S>([&](auto s) {
			combine<s>(h);
			constexpr double cs = Coef<Tab>::c(s);
			this->template rhs_z<s>(fma(cs, h, t));
		});
		combine<S>(h);
		if constexpr (Tab::ERR_NEEDS_FSAL) {
			this->template rhs_z<S>(t + h);
			const double sum = this->sum([&](int i, auto r) {
				double e = 0.0;
				static_for<0, Tab::KROWS>([&](auto j) {
					constexpr double w = Tab::E[j];
					if constexpr (w != 0.0) e = fma(JB_COEF(w, (std::is_same<Tab, DP5>::value ? bank::dp5_e[j] : bank::rk23_e[j])), this->template k<j>(i, r), e);
				});
				const double sk = fma(rtol, fmax(fabs(this->y(i, r)), fabs(this->z(i, r))), atol(i));
				const double q = h * e / sk;
				return q * q;
			});
			return sum * (1.0 / N);
		} else {
			double s5, s3;
			this->sum2([&](int i, auto r, double &a5, double &a3) {
				double e5 = 0.0, e3 = 0.0;
				static_for<0, S>([&](auto j) {
					constexpr double w5 = dop853::E5[j];
					constexpr double w3 = dop853::E3[j];
					if constexpr (w5 != 0.0) e5 = fma(JB_COEF(w5, bank::dop853_e5[j]), this->template k<j>(i, r), e5);
					if constexpr (w3 != 0.0) e3 = fma(JB_COEF(w3, bank::dop853_e3[j]), this->template k<j>(i, r), e3);
				});
				const double sk = fma(rtol, fmax(fabs(this->y(i, r)), fabs(this->z(i, r))), atol(i));
				e5 /= sk;
				e3 /= sk;
				a5 = fma(e5, e5, a5);
				a3 = fma(e3, e3, a3);
			}, s5, s3);
			if (s5 == 0.0 && s3 == 0.0) return 0.0;
			double deno = fma(0.01, s3, s5);
			if (deno <= 0.0) deno = 1.0;
			return fabs(h) * s5 * sqrt(1.0 / (deno * N));
		}
	}

	// the 8th-order pair evaluates f(t+h, y_new) only once the step is accepted
	__device__ __forceinline__ void finish_accepted(double t, double h)
	{
		if constexpr (!Tab::ERR_NEEDS_FSAL) this->template rhs_z<S>(t + h);
	}

	// Hairer's stiffness estimate (dopri5.f / dop853.f "STIFFNESS DETECTION"), evaluated on an accepted step before it
	// is committed: ρ ≈ h·|f(t+h, y_new) − K[S-1]| / |y_new − g|, g = the input of the last stage (c = 1).  K[S] must
	// hold f(t+h, y_new), Z = y_new.  g is recomputed (the check runs once in 1000 steps unless a suspicion is pending).
	__device__ __forceinline__ void stiffness_sums(double h, double &stnum, double &stden)
	{
		this->sum2([&](int i, auto r, double &num, double &den) {
			double acc = 0.0;
			static_for<0, S - 1>([&](auto j) {
				constexpr double a = Coef<Tab>::a(S - 1, j);
				if constexpr (a != 0.0) acc = fma(a, this->template k<j>(i, r), acc);
			});
			const double dk = this->template k<S>(i, r) - this->template k<S - 1>(i, r);
			const double dy = this->z(i, r) - fma(h, acc, this->y(i, r));
			num = fma(dk, dk, num);
			den = fma(dy, dy, den);
		}, stnum, stden);
	}

	// Σ_i (v_i / (atol_i + rtol |Y_i|))², the weighted norms of both start-up rules
	template <class F>
	__device__ __forceinline__ double weighted_sq(F &&value)
	{
		return this->sum([&](int i, auto r) {
			const double q = value(i, r) / fma(rtol, fabs(this->y(i, r)), atol(i));
			return q * q;
		});
	}

	// Z = Y + h K[0]
	__device__ __forceinline__ void euler_probe(double h)
	{
		this->for_each([&](int i, auto r) { this->set_z(i, r, fma(h, this->template k<0>(i, r), this->y(i, r))); });
	}
};

// Bookkeeping of Hairer's stiffness test for one call of the solver: it looks at every NSTIFF-th accepted step and
// at every step while a suspicion is pending; 15 suspicious steps in a row end the call with IDID = −4 (which
// ODE_wrapper.integrate, integrator_tools.py:137-140, turns into UnsuccessfulIntegration), 6 harmless ones clear it.
struct StiffnessWatch {
	static constexpr unsigned NSTIFF = 1000;
	// one register: bits 0-9 accepted steps until the next look, 10-13 IASTI (suspicious steps in a row), 14-16 NONSTI
	// (harmless steps in a row), 17 whether the last estimate exceeded the limit (HLAMB is only ever compared with it)
	unsigned bits;
	__device__ __forceinline__ void reset() { bits = NSTIFF; }
	// called once per accepted step; true if the sums are wanted for this step
	__device__ __forceinline__ bool due()
	{
		if (((--bits) & 0x3ffu) == 0u) { bits |= NSTIFF; return true; }
		return (bits & 0x3c00u) != 0u;
	}
	// returns true if the solver gives up
	__device__ __forceinline__ bool judge(double h, double stnum, double stden, double limit)
	{
		if (stden > 0.0) bits = (fabs(h) * sqrt(stnum / stden) > limit) ? (bits | 0x20000u) : (bits & ~0x20000u);
		if (bits & 0x20000u) {
			bits = (bits & ~0x1c000u) + 0x400u;                 // NONSTI = 0, ++IASTI
			return ((bits >> 10) & 0xfu) == 15u;
		}
		if (((bits >> 14) & 0x7u) < 7u) bits += 0x4000u;       // ++NONSTI (the Fortran counts on; only "== 6" matters, so 7 = "more")
		if (((bits >> 14) & 0x7u) == 6u) bits &= ~0x3c00u;     // IASTI = 0
		return false;
	}
};

// spacing of the doubles just above t: |nextafter(t, ∞) − t| (RungeKutta._step_impl: min_step = 10·that)
__device__ __forceinline__ double ulp_above(double t)
{
	// positive normal t whose spacing is normal too: 2^(exponent − 52), straight from the bits
	const unsigned biased = ((unsigned) __double2hiint(t) >> 20) & 0x7ffu;
	if (t > 0.0 && biased > 52u && biased < 0x7feu) return __hiloint2double((int) ((biased - 52u) << 20), 0);
	return fabs(::nextafter(t, (double) INFINITY) - t);
}

// ---------------------------------------------------------------------------------------------
// Lyapunov post-steps on the state of one trajectory (component access through get/set lambdas)

// jitcode_lyap.norms (_jitcode.py:743-749): Gram–Schmidt of the n_lyap tangent blocks;
// norms_k = |R_kk| of the QR decomposition that orthonormalise_qr computes.
// U = unroll factor (full for register vectors, whose indices must be compile-time constants).
template <int NB, int NL, bool REG, class Vec>
__device__ __forceinline__ void lyap_orthonormalise(Vec &Y, double *norms)
{
	constexpr int UK = REG ? (NL > 0 ? NL : 1) : 1;
	constexpr int UI = REG ? (NB > 0 ? NB : 1) : 4;
#pragma unroll UK
	for (int k = 0; k < NL; ++k) {
		const int ok = NB * (1 + k);
#pragma unroll UK
		for (int j = 0; j < k; ++j) {
			const int oj = NB * (1 + j);
			double dot = 0.0;
#pragma unroll UI
			for (int i = 0; i < NB; ++i) dot = fma(Y(oj + i), Y(ok + i), dot);
#pragma unroll UI
			for (int i = 0; i < NB; ++i) Y.set(ok + i, fma(-dot, Y(oj + i), Y(ok + i)));
		}
		double sq = 0.0;
#pragma unroll UI
		for (int i = 0; i < NB; ++i) sq = fma(Y(ok + i), Y(ok + i), sq);
		const double norm = sqrt(sq);
		norms[k] = norm;
		const double inv = 1.0 / norm;
#pragma unroll UI
		for (int i = 0; i < NB; ++i) Y.set(ok + i, Y(ok + i) * inv);
	}
}

// The same Gram–Schmidt for quad storage: tangent vector k lives in the registers of lane k of the group (components
// NB … 2·NB−1 of its vector).  For j = 0, 1, …: lane j normalises its vector and hands it round (shuffles within the
// group), every later lane removes its projection on it — the same operations in the same order as above.  Returns
// this lane's norm |R_kk|.
template <int NB, int NL, int G, class Vec>
__device__ __forceinline__ double group_orthonormalise(Vec &Y, int member)
{
	double own_norm = 1.0;
#pragma unroll
	for (int j = 0; j < NL; ++j) {
		double q[NB];
		if (member == j) {
			double sq = 0.0;
#pragma unroll
			for (int i = 0; i < NB; ++i) sq = fma(Y(NB + i), Y(NB + i), sq);
			own_norm = sqrt(sq);
			const double inv = 1.0 / own_norm;
#pragma unroll
			for (int i = 0; i < NB; ++i) Y.set(NB + i, Y(NB + i) * inv);
		}
#pragma unroll
		for (int i = 0; i < NB; ++i) q[i] = group_broadcast<G>(Y(NB + i), j);
		if (member > j) {
			double dot = 0.0;
#pragma unroll
			for (int i = 0; i < NB; ++i) dot = fma(q[i], Y(NB + i), dot);
#pragma unroll
			for (int i = 0; i < NB; ++i) Y.set(NB + i, fma(-dot, q[i], Y(NB + i)));
		}
	}
	return own_norm;
}

// jitcode_restricted_lyap.norms (_jitcode.py:937-946): remove the projections onto the given
// (orthonormalised) vectors from the single tangent vector, then normalise.
template <int NB, bool REG, class Vec>
__device__ __forceinline__ double lyap_restricted(Vec &Y, const double *proj, int n_proj)
{
	constexpr int UI = REG ? (NB > 0 ? NB : 1) : 4;
	for (int v = 0; v < n_proj; ++v) {
		const double *vec = proj + (int64_t) v * NB;
		double dot = 0.0;
#pragma unroll UI
		for (int i = 0; i < NB; ++i) dot = fma(vec[i], Y(NB + i), dot);
#pragma unroll UI
		for (int i = 0; i < NB; ++i) Y.set(NB + i, fma(-dot, vec[i], Y(NB + i)));
	}
	double sq = 0.0;
#pragma unroll UI
	for (int i = 0; i < NB; ++i) sq = fma(Y(NB + i), Y(NB + i), sq);
	const double norm = sqrt(sq);
	const double inv = 1.0 / norm;
#pragma unroll UI
	for (int i = 0; i < NB; ++i) Y.set(NB + i, Y(NB + i) * inv);
	return norm;
}

// jitcode_transversal_lyap.norm (_jitcode.py:886-893): norm over the tangent (difference) coordinates
template <class Sys, bool REG, class Vec>
__device__ __forceinline__ double lyap_transversal(Vec &Y)
{
	constexpr int UT = REG ? (Sys::NT > 0 ? Sys::NT : 1) : 4;
	double sq = 0.0;
#pragma unroll UT
	for (int k = 0; k < Sys::NT; ++k) {
		const double v = Y(Sys::tangent_index(k));
		sq = fma(v, v, sq);
	}
	const double norm = sqrt(sq);
	const double inv = 1.0 / norm;
#pragma unroll UT
	for (int k = 0; k < Sys::NT; ++k) Y.set(Sys::tangent_index(k), Y(Sys::tangent_index(k)) * inv);
	return norm;
}

// The same three post-steps for warp storage: the state sits in one shared-memory row `v` (component c at
// v[where[c]]), the threads of the team stride over the components of a tangent vector, and every scalar product
// is a shuffle reduction (two-level for teams of several warps) — the warp-level batched Gram–Schmidt of the
// north star.  All threads of the team return the same norms.
template <class Sys, class WT>
__device__ __forceinline__ void team_lyap(int post, double *v, const WT *where, const double *proj, int n_proj,
		double *norms, double *scratch)
{
	using T = Team<Sys::TEAM_WARPS>;
	constexpr int NB = Sys::NB;
	const int rank = T::rank();
	auto dot = [&](auto &&left, auto &&right, int count) {
		double partial = 0.0;
		for (int i = rank; i < count; i += T::SIZE) partial = fma(left(i), right(i), partial);
		return T::sum(partial, scratch);
	};
	auto tangent = [&](int k) { return [=](int i) -> double & { return v[where[NB * (1 + k) + i]]; }; };
	T::sync();
	if (post == JB_POST_LYAP_QR) {
		for (int k = 0; k < Sys::NL; ++k) {
			auto vk = tangent(k);
			for (int j = 0; j < k; ++j) {
				auto vj = tangent(j);
				const double overlap = dot(vj, vk, NB);
				for (int i = rank; i < NB; i += T::SIZE) vk(i) = fma(-overlap, vj(i), vk(i));
				T::sync();
			}
			const double norm = sqrt(dot(vk, vk, NB));
			norms[k] = norm;
			const double inverse = 1.0 / norm;
			for (int i = rank; i < NB; i += T::SIZE) vk(i) *= inverse;
			T::sync();
		}
	} else if (post == JB_POST_LYAP_RESTRICTED) {
		auto v0 = tangent(0);
		for (int p = 0; p < n_proj; ++p) {
			const double *plane = proj + (int64_t) p * NB;
			const double overlap = dot([&](int i) { return plane[i]; }, v0, NB);
			for (int i = rank; i < NB; i += T::SIZE) v0(i) = fma(-overlap, plane[i], v0(i));
			T::sync();
		}
		const double norm = sqrt(dot(v0, v0, NB));
		norms[0] = norm;
		const double inverse = 1.0 / norm;
		for (int i = rank; i < NB; i += T::SIZE) v0(i) *= inverse;
		T::sync();
	} else if (post == JB_POST_LYAP_TRANSVERSAL) {
		auto vt = [=](int k) -> double & { return v[where[Sys::tangent_index(k)]]; };
		const double norm = sqrt(dot(vt, vt, Sys::NT));
		norms[0] = norm;
		const double inverse = 1.0 / norm;
		for (int k = rank; k < Sys::NT; k += T::SIZE) vt(k) *= inverse;
		T::sync();
	}
}

// ---------------------------------------------------------------------------------------------
// Everything one trajectory does during one launch, as an object: `load` binds the working storage and reads the
// solver's state, `begin_sample` / `attempt_once` advance it towards a sample time one step attempt at a time,
// `sample` delivers the state there (dense output, recording, Lyapunov post-step), `store` writes the solver back.
// run_trajectory strings them together for one trajectory; run_lanes (register storage) keeps the lanes of a warp
// busy by handing a lane the next trajectory as soon as its own has arrived.
// Thread storage: one thread owns the object.  Warp storage: all threads of the trajectory's team hold identical
// scalar state; `smem` is the team's private stage area, `tables` the staged tables of the right-hand side.

template <class Sys, class Tab, int DRIVER, int MODE, int KREGS = 0>
struct Lane {
	static constexpr bool WARP = MODE == JB_STORE_WARP;
	using Store = typename std::conditional<WARP, WarpStore<Sys, Tab, DRIVER, KREGS>, ThreadStore<Sys, Tab, DRIVER, MODE>>::type;
	using ST = Stepper<Sys, Tab, DRIVER, Store>;
	using IVec = typename ST::IVec;
	using UVec = typename ST::UVec;
	static constexpr int N = Sys::N;
	static constexpr int S = Tab::NSTAGE;
	static constexpr bool QUAD = MODE == JB_STORE_QUAD;
	static constexpr bool REG = MODE == JB_STORE_REGISTERS || QUAD;
	static constexpr int NV = Store::N;                                     // length of this thread's vectors (N unless QUAD)
	static constexpr bool DENSE = ST::DENSE;
	static constexpr bool IVP = ST::IVP;
	static constexpr int NLY = QUAD ? 1 : (Sys::NL > 0 ? Sys::NL : 1);      // exponents this thread computes
	static constexpr double LOG_FACOLD_MIN = -9.210340371976182;           // log(1e-4)

	const Params &a;
	ST st;
	int64_t b, off;             // the trajectory and its place inside a caller-visible vector
	bool leader, owner;         // writes the per-trajectory scalars / (quad storage) owns a tangent vector
	int member;                 // quad storage: place in the trajectory's group
	int status;
	double t;
	int n_acc, n_rej;           // of this launch (one register each; the caller's totals are 64-bit)
	// Hairer's driver: one call of the solver per sample time
	double h, hmax, log_facold;
	bool reject, last;
	int nstep;
	StiffnessWatch stiff;
	// SciPy's stepper: persistent between sample times and launches
	double h_abs, t_old, h_prev;
	bool initialised;           // the persistent stepper exists (f(t, y) and h_abs are valid)
	bool interpolant;           // DENSE: the interpolant of the last step is in the canonical state
	bool have_f;                // K[0] holds f(t, Y) for the step to come
	bool pending_f;             // K[S] holds f(t, Y) (fresh from an accepted step), K[0..S] are its stages
	bool in_step, rejected;     // inside RungeKutta._step_impl: a step has begun / has seen a rejection
	double min_step;
	int attempts_here;
	// Lyapunov exponents
	double t_start;             // solver time before the current sample (Δt of the Lyapunov step for ODE/LAND)
	double lyap_sum[NLY], lyap_sq[NLY];

	__device__ __forceinline__ explicit Lane(const Params &params) : a(params) {}

	// ---- canonical (caller-visible) locations
	__device__ __forceinline__ UVec user(double *base) const
	{
		if constexpr (WARP) return UVec{base, st.comp};
		else if constexpr (QUAD) return UVec{base, Sys::NB * member, member == 0, owner};
		else return UVec{base};
	}
	__device__ __forceinline__ IVec persistent(double *base) const
	{
		if constexpr (QUAD) return IVec{base, Sys::NB * member, member == 0, owner};
		else return IVec{base};
	}
	__device__ __forceinline__ int64_t vec() const { return (int64_t) N * a.ld; }      // one caller-visible vector
	__device__ __forceinline__ UVec Yc() const { return user(a.Y + off); }
	__device__ __forceinline__ IVec Fc() const { return persistent(IVP ? a.state + SV_F * vec() + off : nullptr); }
	__device__ __forceinline__ IVec Oc() const { return persistent(DENSE ? a.state + SV_YOLD * vec() + off : nullptr); }
	__device__ __forceinline__ IVec Qc(int p) const { return persistent(a.state + (SV_DENSE0 + p) * vec() + off); }

	// ================= bind the working storage, read the solver
	__device__ __forceinline__ void load(const int64_t trajectory, double *smem = nullptr, const void *tables = nullptr)
	{
		b = trajectory;
		off = WARP ? b * (int64_t) N : tile_offset(b, N);
		leader = true;
		if constexpr (WARP) leader = Team<Sys::TEAM_WARPS>::rank() == 0;
		member = QUAD ? (int) (threadIdx.x % (unsigned) Sys::GROUP) : 0;
		owner = !QUAD || member < Sys::NL;
		if constexpr (QUAD) leader = member == 0;

		st.rtol = a.rtol;
		st.atol_s = a.atol;
		st.atol_vec = a.atol_vec;
		st.comp = nullptr;
		st.shift = owner ? Sys::NB * member : 0;
#pragma unroll
		for (int k = 0; k < Sys::NP; ++k)
			st.par.v[k] = WARP ? a.P[b * Sys::NP + k] : a.P[tile_offset(b, Sys::NP) + k * TILE];
		status = a.status[b];
		t = a.t[b];
		n_acc = n_rej = 0;
		if constexpr (WARP) {
			st.tables = tables;
			st.comp = Sys::components(tables);
		}
		const UVec Y0 = Yc();
		if constexpr (REG) {
#pragma unroll
			for (int i = 0; i < NV; ++i) st.Y.v[i] = Y0(i);
		} else if constexpr (WARP) {
			st.bind(smem);
			st.for_each([&](int i, auto r) { st.set_y(i, r, Y0(i)); });
		} else if constexpr (MODE == JB_STORE_SHARED) {
			// rows of N × (threads per block) doubles in the block's shared memory; this thread's column starts at `smem`
			constexpr int ROW = N * Sys::THREADS;
			st.Y.p = smem;
			st.Z.p = smem + ROW;
#pragma unroll
			for (int j = 0; j < ST::KR; ++j) st.K[j].p = smem + (2 + j) * ROW;
			if constexpr (DENSE) st.YO.p = smem + (2 + ST::KR) * ROW;
			st.for_each([&](int i, auto r) { st.set_y(i, r, Y0(i)); });
		} else {
			st.Y.p = Y0.p;
			st.Z.p = a.work + off;
#pragma unroll
			for (int j = 0; j < ST::KR; ++j) st.K[j].p = a.work + (int64_t) (1 + j) * vec() + off;
			if constexpr (DENSE) st.YO.p = a.work + (int64_t) (1 + ST::KR) * vec() + off;
		}
		h_abs = 0.0, t_old = NAN, h_prev = 0.0;
		initialised = interpolant = have_f = pending_f = false;
		if constexpr (IVP) {
			h_abs = a.sstate[SS_H_ABS * a.ld + b];
			t_old = a.sstate[SS_T_OLD * a.ld + b];
			h_prev = a.sstate[SS_H_PREV * a.ld + b];
			const int flags = (int) a.sstate[SS_FLAGS * a.ld + b];
			initialised = flags & FLAG_INITIALISED;
			interpolant = flags & FLAG_INTERPOLANT;
		}
#pragma unroll
		for (int k = 0; k < NLY; ++k) lyap_sum[k] = lyap_sq[k] = 0.0;
		t_start = t;
	}

	// ================= start towards the sample time T; false: there is nothing to integrate
	__device__ __forceinline__ bool begin_sample(const double T)
	{
		t_start = t;
		if constexpr (DRIVER == JB_DRV_ODE) {
			// Hairer DOPRI5 / DOP853, fresh start (dopri5.f DOPCOR, dop853.f DP86CO)
			if (!(T > t)) return false;
			st.template rhs_y<0>(t);
			hmax = a.max_step > 0.0 ? a.max_step : fabs(T - t);
			if (a.first_step > 0.0) {
				h = a.first_step;
			} else {
				// HINIT
				const double dnf = st.weighted_sq([&](int i, auto r) { return st.template k<0>(i, r); });
				const double dny = st.weighted_sq([&](int i, auto r) { return st.y(i, r); });
				h = (dnf <= 1e-10 || dny <= 1e-10) ? 1e-6 : sqrt(dny / dnf) * 0.01;
				h = fmin(h, hmax);
				st.euler_probe(h);
				st.template rhs_z<1>(t + h);
				const double der2 = sqrt(st.weighted_sq([&](int i, auto r) { return st.template k<1>(i, r) - st.template k<0>(i, r); })) / h;
				const double der12 = fmax(fabs(der2), sqrt(dnf));
				const double h1 = der12 <= 1e-15 ? fmax(1e-6, fabs(h) * 1e-3)
						: pow(0.01 / der12, 1.0 / Tab::HAIRER_ORDER);
				h = fmin(fmin(100.0 * fabs(h), h1), hmax);
			}
			log_facold = LOG_FACOLD_MIN;      // facold = 1e-4
			reject = last = false;
			nstep = 0;
			stiff.reset();
			return true;
		} else {
			// scipy solve_ivp stepper (rk.py RungeKutta), persistent between launches
			if (!initialised) {
				// RungeKutta.__init__: f = fun(t, y); select_initial_step with t_bound = ∞
				st.template rhs_y<0>(t);
				if (a.first_step > 0.0) {
					h_abs = a.first_step;
				} else {
					const double d0 = sqrt(st.weighted_sq([&](int i, auto r) { return st.y(i, r); }) / N);
					const double d1 = sqrt(st.weighted_sq([&](int i, auto r) { return st.template k<0>(i, r); }) / N);
					const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
					st.euler_probe(h0);
					st.template rhs_z<1>(t + h0);
					const double d2 = sqrt(st.weighted_sq([&](int i, auto r) { return st.template k<1>(i, r) - st.template k<0>(i, r); }) / N) / h0;
					const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3)
							: pow(0.01 / fmax(d1, d2), 1.0 / (Tab::ERR_ORDER + 1));
					h_abs = fmin(fmin(100.0 * h0, h1), a.max_step > 0.0 ? a.max_step : INFINITY);
				}
				t_old = NAN;
				initialised = true;
				interpolant = false;
				have_f = true;
				pending_f = false;
			}
			attempts_here = 0;      // solve_ivp has no step limit; a launch must end nevertheless
			in_step = false;
			return more_steps(T);
		}
	}

	// IVP_wrapper.integrate (integrator_tools.py:98-105): step until the solver has passed T (and has made a step at all)
	__device__ __forceinline__ bool more_steps(const double T) const
	{
		return status == JB_OK && (t < T || (DENSE && !(t_old == t_old)));
	}

	// ================= one step attempt; true: the stepping for this sample time is over (arrived or failed)
	__device__ __forceinline__ bool attempt_once(const double T)
	{
		if constexpr (DRIVER == JB_DRV_ODE) {
			const int nmax = a.nsteps > 0 ? (int) (a.nsteps < 2000000000 ? a.nsteps : 2000000000) : 100000;
			if (nstep > nmax) { status = JB_TOO_MANY_STEPS; return true; }
			if (0.1 * fabs(h) <= fabs(t) * 2.3e-16) { status = JB_STEP_TOO_SMALL; return true; }
			if (!(h == h)) { status = JB_NONFINITE; return true; }
			if (t + 1.01 * h - T > 0.0) { h = T - t; last = true; }
			++nstep;
			const double err = st.attempt(t, h);
			// fac11 = err^expo1 and facold^beta through one logarithm per attempt (facold is the previous error)
			const double log_err = ST::log_error(err);
			if (ST::accepted_hairer(err)) {
				++n_acc;
				st.finish_accepted(t, h);
#ifndef JB_NO_STIFFNESS
				if (stiff.due()) {      // (before the new step size is worked out: fewer values alive across the sums)
					double stnum, stden;
					st.stiffness_sums(h, stnum, stden);
					if (stiff.judge(h, stnum, stden, Tab::STIFF_LIMIT)) { status = JB_STIFF; return true; }
				}
#endif
				// h_new = h / clamp(fac11 / facold^β / safety, 1/ifactor, 1/dfactor), written without a division
				double hnew = h * fmin(a.ifactor, fmax(a.dfactor, a.safety * exp(a.beta * log_facold - a.expo1 * log_err)));
				log_facold = fmax(log_err, LOG_FACOLD_MIN);     // facold = max(err, 1e-4)
				st.accept_y();
				st.template fsal<0, S>();
				t = last ? T : t + h;
				if (fabs(hnew) > hmax) hnew = hmax;
				if (reject) hnew = fmin(fabs(hnew), fabs(h));
				reject = false;
				if (last) return true;
				h = hnew;
			} else {
				if (Tab::ERR_NEEDS_FSAL || !a.hairer_reject_by_dfactor)
					h = h * fmax(a.dfactor, a.safety * exp(-a.expo1 * log_err));      // h / min(1/dfactor, fac11/safety)
				else
					h = h * a.dfactor;
				reject = true;
				last = false;
				++n_rej;
			}
			return false;
		} else {
			constexpr int ATTEMPT_CAP = 1 << 26;
			const double max_step = a.max_step > 0.0 ? a.max_step : INFINITY;
			const double t_bound = DENSE ? INFINITY : T;
			if (!in_step) {
				// ---- begin a step: K[0] = f(t, Y)
				if (pending_f) {
					st.template fsal<0, S>();
					pending_f = false;
				} else if (!have_f) {
					const IVec F = Fc();
					st.for_each([&](int i, auto r) { st.template set_k<0>(i, r, F(i)); });
				}
				have_f = true;
				// ---- RungeKutta._step_impl
				min_step = 10.0 * ulp_above(t);
				if (h_abs > max_step) h_abs = max_step;
				else if (h_abs < min_step) h_abs = min_step;
				rejected = false;
				in_step = true;
			}
			if (!(fabs(h_abs) <= 1.79769313486231571e308)) { status = JB_NONFINITE; return true; }   // NaN or ∞ (non-finite state)
			if (h_abs < min_step) { status = JB_STEP_TOO_SMALL; return true; }
			if (++attempts_here > ATTEMPT_CAP) { status = JB_TOO_MANY_STEPS; return true; }
			double t_new = t + h_abs;
			if (t_new - t_bound > 0.0) t_new = t_bound;
			const double step = t_new - t;
			h_abs = fabs(step);
			const double err = st.attempt(t, step);
			if (ST::accepted_scipy(err)) {
				double factor = err == 0.0 ? 10.0 : fmin(10.0, 0.9 * exp(ST::log_error(err) * (-1.0 / (Tab::ERR_ORDER + 1))));
				if (rejected) factor = fmin(1.0, factor);
				h_abs *= factor;
				++n_acc;
				st.finish_accepted(t, step);
				if constexpr (DENSE) st.save_old();
				st.accept_y();
				t_old = t;
				h_prev = step;
				t = t_new;
				pending_f = true;
				interpolant = false;
				in_step = false;
				return !more_steps(T);
			}
			h_abs *= fmax(0.2, 0.9 * exp(ST::log_error(err) * (-1.0 / (Tab::ERR_ORDER + 1))));
			rejected = true;
			++n_rej;
			return false;
		}
	}

	// ================= the sample at T (sample index `it` of the launch)
	__device__ __forceinline__ void sample(const int it, const double T)
	{
		const UVec rec = user(a.Yrec ? a.Yrec + (int64_t) it * vec() + off : nullptr);
		const UVec res = user(a.Yres ? a.Yres + off : nullptr);
		const double dt_lyap = T - t_start;
		if constexpr (DENSE) {
			const IVec O = Oc();
			if (!interpolant) {
				// ---- materialise the interpolant of the last step (rk.py:178-180, 693-712)
				if constexpr (!std::is_same<Tab, DOP853>::value) {
					st.for_each([&](int i, auto r) {
						static_for<0, Tab::NDENSE>([&](auto p) {
							double q = 0.0;
							static_for<0, Tab::KROWS>([&](auto j) {
								constexpr double w = Tab::P[j][p];
								if constexpr (w != 0.0) q = fma(w, st.template k<j>(i, r), q);
							});
							Qc(p).set(i, q);
						});
						O.set(i, st.yo(i, r));
					});
				} else {
					// three extra stages from (t_old, y_old), then F[0..6]
					static_for<S + 1, S + 4>([&](auto s) {
						st.for_each([&](int i, auto r) {
							double acc = 0.0;
							static_for<0, s>([&](auto j) {
								constexpr double c = Coef<DOP853>::a(s, j);
								if constexpr (c != 0.0) acc = fma(c, st.template k<j>(i, r), acc);
							});
							st.set_z(i, r, fma(h_prev, acc, st.yo(i, r)));
						});
						constexpr double cs = Coef<DOP853>::c(s);
						st.template rhs_z<s>(fma(cs, h_prev, t_old));
					});
					st.for_each([&](int i, auto r) {
						const double delta = st.y(i, r) - st.yo(i, r);
						const double f_old = st.template k<0>(i, r), f_new = st.template k<S>(i, r);
						Qc(0).set(i, delta);
						Qc(1).set(i, h_prev * f_old - delta);
						Qc(2).set(i, 2.0 * delta - h_prev * (f_new + f_old));
						static_for<0, 4>([&](auto row) {
							double q = 0.0;
							static_for<0, 16>([&](auto j) {
								constexpr double w = dop853::D[row][j];
								if constexpr (w != 0.0) q = fma(w, st.template k<j>(i, r), q);
							});
							Qc(3 + row).set(i, h_prev * q);
						});
						O.set(i, st.yo(i, r));
					});
				}
				interpolant = true;
			}
			// ---- evaluate it at T (rk.py:723-737, 747-764)
			const double hh = t - t_old;
			const double x = (T - t_old) / hh;
			const bool restart = a.post != JB_POST_NONE;
			st.for_each([&](int i, auto pass) {
				double r;
				if constexpr (!std::is_same<Tab, DOP853>::value) {
					r = Qc(Tab::NDENSE - 1)(i);
					static_for<1, Tab::NDENSE>([&](auto p) { r = fma(r, x, Qc(Tab::NDENSE - 1 - p)(i)); });
					r = fma(hh * x, r, O(i));
				} else {
					r = 0.0;
					static_for<0, 7>([&](auto k) {
						r += Qc(6 - k)(i);
						r *= (k % 2 == 0) ? x : 1.0 - x;
					});
					r += O(i);
				}
				if (restart) st.set_z(i, pass, r);
				else {
					res.set(i, r);
					if (rec.p) rec.set(i, r);
				}
			});
			if (restart) {
				// jitcode_lyap.integrate re-creates the stepper at (T, sample) (_jitcode.py:773)
				st.accept_y();
				t = T;
				initialised = false;
				interpolant = false;
				pending_f = have_f = false;
			}
		}

		if constexpr (Sys::NL > 0)      // (modules without tangent vectors: nothing of this is compiled)
		if (a.post != JB_POST_NONE) {
			double norms[NLY];
			if constexpr (WARP) {
				// through the Z row: the state may live in registers, the tangent vectors are needed by every thread
				st.for_each([&](int i, auto r) { st.set_z(i, r, st.y(i, r)); });
				team_lyap<Sys>(a.post, st.pZ, Sys::positions(st.tables), a.proj, a.n_proj, norms, st.scratch);
				st.for_each([&](int i, auto r) { st.set_y(i, r, st.z(i, r)); });
			} else if constexpr (QUAD) {
				norms[0] = group_orthonormalise<Sys::NB, Sys::NL, Sys::GROUP>(st.Y, member);
			} else {
				if (a.post == JB_POST_LYAP_QR)
					lyap_orthonormalise<Sys::NB, Sys::NL, REG>(st.Y, norms);
				else if (a.post == JB_POST_LYAP_RESTRICTED)
					norms[0] = lyap_restricted<Sys::NB, REG>(st.Y, a.proj, a.n_proj);
				else if (a.post == JB_POST_LYAP_TRANSVERSAL)
					norms[0] = lyap_transversal<Sys, REG>(st.Y);
			}
			const int n_out = a.post == JB_POST_LYAP_QR ? Sys::NL : 1;
			if constexpr (QUAD) {
				// every owning lane reports the exponent of its own tangent vector
				const double lyap = log(norms[0]) / dt_lyap;
				if (a.lyap_rec && owner)
					a.lyap_rec[tile_offset(b, a.n_times * n_out) + (int64_t) (it * n_out + member) * TILE] = lyap;
				if (it >= a.lyap_discard) {
					lyap_sum[0] += lyap;
					lyap_sq[0] = fma(lyap, lyap, lyap_sq[0]);
				}
			} else
			for (int k = 0; k < NLY; ++k) {
				if (k >= n_out) break;
				const double lyap = log(norms[k]) / dt_lyap;
				if (a.lyap_rec && leader) {
					// one vector of n_times · n_out entries per trajectory, entry (it, k) at row it·n_out + k
					if constexpr (WARP) a.lyap_rec[((int64_t) b * a.n_times + it) * n_out + k] = lyap;
					else a.lyap_rec[tile_offset(b, a.n_times * n_out) + (int64_t) (it * n_out + k) * TILE] = lyap;
				}
				if (it >= a.lyap_discard) {
					lyap_sum[k] += lyap;
					lyap_sq[k] = fma(lyap, lyap, lyap_sq[k]);
				}
			}
			if constexpr (IVP) {
				// the tangent vectors changed: the stepper restarts from the new state
				initialised = false;
				interpolant = false;
				pending_f = have_f = false;
			}
		}
		if constexpr (!DENSE) {
			if (rec.p) st.for_each([&](int i, auto r) { rec.set(i, st.y(i, r)); });
		} else if (a.post != JB_POST_NONE) {
			st.for_each([&](int i, auto r) {
				res.set(i, st.y(i, r));
				if (rec.p) rec.set(i, st.y(i, r));
			});
		}
	}

	// ================= write the solver back
	__device__ __forceinline__ void store()
	{
		const UVec Y0 = Yc();
		if constexpr (REG) {
#pragma unroll
			for (int i = 0; i < NV; ++i) Y0.set(i, st.Y.v[i]);
		} else {
			bool copy_back = WARP || MODE == JB_STORE_SHARED;
			if constexpr (MODE == JB_STORE_GLOBAL) copy_back = st.Y.p != Y0.p;
			if (copy_back) st.for_each([&](int i, auto r) { Y0.set(i, st.y(i, r)); });
		}
		if constexpr (IVP) {
			if (pending_f || have_f) {
				const IVec F = Fc();
				if (pending_f) st.for_each([&](int i, auto r) { F.set(i, st.template k<S>(i, r)); });
				else st.for_each([&](int i, auto r) { F.set(i, st.template k<0>(i, r)); });
			}
			if (leader) {
				a.sstate[SS_H_ABS * a.ld + b] = h_abs;
				a.sstate[SS_T_OLD * a.ld + b] = t_old;
				a.sstate[SS_H_PREV * a.ld + b] = h_prev;
				a.sstate[SS_FLAGS * a.ld + b] = (double) ((initialised ? FLAG_INITIALISED : 0) | (interpolant ? FLAG_INTERPOLANT : 0));
			}
		}
		if (leader) {
			a.t[b] = t;
			a.status[b] = status;
			if (a.n_accepted) a.n_accepted[b] += n_acc;
			if (a.n_rejected) a.n_rejected[b] += n_rej;
		}
		if constexpr (Sys::NL <= 0) {
			// no tangent vectors, no exponents
		} else if constexpr (QUAD) {
			if (a.lyap_acc && a.post != JB_POST_NONE && owner) {
				const int n_out = Sys::NL;
				a.lyap_acc[tile_offset(b, n_out) + member * TILE] += lyap_sum[0];
				a.lyap_acc[(int64_t) n_out * a.ld + tile_offset(b, n_out) + member * TILE] += lyap_sq[0];
			}
		} else if (a.lyap_acc && a.post != JB_POST_NONE && leader) {
			const int n_out = a.post == JB_POST_LYAP_QR ? Sys::NL : 1;
			for (int k = 0; k < NLY; ++k) {
				if (k >= n_out) break;
				if constexpr (WARP) {
					a.lyap_acc[b * n_out + k] += lyap_sum[k];
					a.lyap_acc[((int64_t) a.ld + b) * n_out + k] += lyap_sq[k];
				} else {
					a.lyap_acc[tile_offset(b, n_out) + k * TILE] += lyap_sum[k];
					a.lyap_acc[(int64_t) n_out * a.ld + tile_offset(b, n_out) + k * TILE] += lyap_sq[k];
				}
			}
		}
	}
};

template <class Sys, class Tab, int DRIVER, int MODE, int KREGS = 0>
__device__ __forceinline__ void run_trajectory(const Params &a, const int64_t b, double *smem = nullptr, const void *tables = nullptr)
{
	Lane<Sys, Tab, DRIVER, MODE, KREGS> lane(a);
	lane.load(b, smem, tables);
	for (int it = 0; it < a.n_times; ++it) {
		const double T = a.times[it];
		if (lane.status != JB_OK) continue;       // failures are sticky: a failed trajectory is not advanced further
		if (lane.begin_sample(T))
			while (!lane.attempt_once(T)) {}
		if (lane.status != JB_OK) continue;
		lane.sample(it, T);
	}
	lane.store();
}

// ---------------------------------------------------------------------------------------------
#ifndef JB_REFILL_LANES
#define JB_REFILL_LANES 4      // waiting lanes that change trajectories together (measured: see DESIGN.md)
#endif
// Lane refill (register storage, no post-step; all three drivers, any number of sample times).  With one thread per
// trajectory the lanes of a warp wait for the trajectory that needs most steps (Lotka–Volterra ensemble: 24 % of the
// lane slots idle).  Here the warps are persistent and a lane whose trajectory has arrived at its last sample time
// delivers it and takes the next one from a queue while its neighbours keep stepping.  Delivering a sample (dense
// output, recording), writing the solver back and starting anew (f(t, y), initial step size) are code the stepping
// lanes cannot share, so lanes that have arrived at a sample time wait until REFILL of them can do that together (or
// nobody is stepping).  The arithmetic of a trajectory is exactly that of run_trajectory: the same calls of the same
// Lane in the same order.
template <class Sys, class Tab, int DRIVER>
__device__ __forceinline__ void run_lanes(const Params &a)
{
	constexpr unsigned FULL = 0xffffffffu;
	constexpr int REFILL = JB_REFILL_LANES;
	Lane<Sys, Tab, DRIVER, JB_STORE_REGISTERS> lane(a);
	const int index = threadIdx.x & 31;
	bool occupied = false;      // the lane holds a trajectory
	bool stepping = false;      // … which is on its way to sample time `it`
	bool more = true;           // the queue may still hold trajectories
	int it = 0;
	double T = 0.0;             // a.times[it], in a register (the step loop compares with it in every attempt)
	while (true) {
		const unsigned waiting = __ballot_sync(FULL, !stepping);
		const unsigned able = __ballot_sync(FULL, !stepping && (occupied || more));      // lanes with something to deliver or to fetch
		if (waiting == FULL && able == 0) break;
		if (able != 0 && (__popc(able) >= REFILL || waiting == FULL)) {
			// ---- change of trajectories.  Pass 0: the lanes that have arrived at sample time `it` (or failed) deliver
			// the sample and go on to their next sample time, or write the solver back and become free.  Then the free
			// lanes take the next trajectories of the queue (one atomic per warp).  Pass 1: those start.  (One loop, not
			// two copies of its body: delivering and starting are the bulky parts of the kernel.)
			bool arrived = !stepping && occupied, fresh = false;
#pragma unroll 1
			for (int pass = 0; pass < 2; ++pass) {
				if (pass == 0 ? arrived : fresh) {
					bool deliver = arrived;
					while (true) {
						if (deliver) {
							if (lane.status == JB_OK) lane.sample(it, T);
							++it;
						}
						if (it >= a.n_times || lane.status != JB_OK) {      // through with all sample times, or failed (failures are sticky)
							lane.store();
							occupied = false;
							break;
						}
						T = a.times[it];
						if (lane.begin_sample(T)) {
							stepping = true;
							break;
						}
						deliver = true;       // nothing to integrate for this sample time: it is delivered at once
					}
				}
				if (pass == 0) {
					arrived = false;
					const unsigned hungry = __ballot_sync(FULL, !occupied && more);
					if (hungry == 0) break;
					const int first = __ffs(hungry) - 1;
					unsigned long long base = 0;
					if (index == first) base = atomicAdd(a.queue, (unsigned long long) __popc(hungry));
					base = __shfl_sync(FULL, base, first);
					if (!occupied && more) {
						const int64_t mine = (int64_t) base + __popc(hungry & ((1u << index) - 1u));
						if (mine >= a.B) {
							more = false;
						} else {
							lane.load(mine);
							it = 0;
							occupied = fresh = true;
						}
					}
				}
			}
		}
		if (stepping) stepping = !lane.attempt_once(T);
	}
}

}  // namespace jb
