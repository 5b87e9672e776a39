/*
 * jitcode_b200 — C ABI of one run-time compiled ODE module (sm_100a).
 *
 * One module = one symbolic system (its right-hand side and, for the Lyapunov classes, its
 * tangent equations) fused into one adaptive Runge–Kutta ensemble kernel, compiled with
 * `nvcc -gencode arch=compute_100a,code=sm_100a -shared` and loaded with ctypes.  It replaces
 * the CPython extension the reference builds from jitcode/jitced_template.c:
 *
 *   reference entry point (jitced_template.c)          replaced by
 *   ------------------------------------------------   ------------------------------------
 *   f(t, Y) -> dY             py_f        :67-116       jb_eval_f   (B states per call)
 *   initialise(*pars)         py_initialise :184-209    the P argument (per-trajectory parameters,
 *                                                       no module-level globals :10-12)
 *   `dimension`               :8                        jb_get_info().n
 *   module table              :213-240                  the extern "C" symbols below
 *
 * and — because the batched integrator is fused with the right-hand side — also the SciPy
 * steppers the reference reaches through jitcode/integrator_tools.py:
 *
 *   ODE_wrapper.integrate                 :134-144  ->  jb_integrate, driver JB_DRV_ODE
 *   IVP_wrapper.integrate                 :98-105   ->  jb_integrate, driver JB_DRV_IVP_DENSE
 *   IVP_wrapper_no_interpolation.integrate :110-128 ->  jb_integrate, driver JB_DRV_IVP_LAND
 *   jitcode_lyap.norms / integrate (_jitcode.py:743-775), jitcode_restricted_lyap.norms
 *   (:937-946), jitcode_transversal_lyap.norm (:886-893)  ->  the JB_POST_* stage of jb_integrate
 *
 * Conventions
 *   - All array arguments are DEVICE pointers owned by the caller; the module allocates nothing
 *     and keeps no mutable global state.  Kernels are launched on `stream` (a cudaStream_t).
 *   - Layout is structure-of-arrays, trajectory-minor, in tiles of JB_TILE = 32 trajectories (one
 *     tile = one warp): component i of trajectory b of an n-vector lives at
 *         X[(b / 32) * n * 32  +  i * 32  +  b % 32],
 *     so that a warp touches one contiguous 256-byte segment per component and a tile's whole
 *     vector is one contiguous n*256-byte block.  ld = B rounded up to a multiple of 32 is the
 *     padded number of trajectories; a "[k][n][ld]" array below is k such tiled vectors back to
 *     back (k * n * ld doubles).  Per-trajectory scalars are plain arrays [ld].  jb_pack /
 *     jb_unpack convert from / to the reference's row-per-trajectory (B, n) layout.
 *     Modules with storage JB_STORE_WARP (one warp per trajectory, lanes across components) read a
 *     trajectory's vector as one contiguous row instead: every n-vector, parameter block and
 *     "[k][n][ld]" array is row-major (k, ld, n) — the reference's own layout, X[b*n + i], ld >= B
 *     without padding requirement — and jb_pack / jb_unpack are plain copies.
 *   - Every function returns 0 on success, otherwise a cudaError_t value (>0) or a negative
 *     JB_E_* code; jb_last_error() gives the text for the calling thread.
 *   - Per-trajectory outcome is reported in `status[b]` (enum jb_status).
 */
#ifndef JITCODE_B200_H
#define JITCODE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JB_ABI_VERSION 6
#define JB_TILE 32

enum jb_method  { JB_DP5 = 0, JB_DOP853 = 1, JB_RK23 = 2 /* solve_ivp drivers only */ };
enum jb_driver  {
	JB_DRV_ODE = 0,        /* Hairer controller; fresh start per sample time; lands exactly on it */
	JB_DRV_IVP_DENSE = 1,  /* scipy solve_ivp controller; persistent stepper; dense output at the sample time */
	JB_DRV_IVP_LAND = 2    /* scipy solve_ivp controller; persistent stepper; last step clipped to the sample time */
};
enum jb_storage {
	JB_STORE_REGISTERS = 0, /* thread per trajectory, state and stages in registers; tiled layout */
	JB_STORE_GLOBAL = 1,    /* thread per trajectory, stages in the caller's tiled workspace */
	JB_STORE_WARP = 2,      /* warp per trajectory, stages in shared memory; ROW layout (see below) */
	JB_STORE_SHARED = 3,    /* thread per trajectory, stages in the block's shared memory; tiled layout */
	JB_STORE_QUAD = 5       /* Lyapunov modules: a group of 2^k >= n_lyap lanes per trajectory, lane j integrates the state
	                           (redundantly) and tangent vector j in registers; norms and Gram–Schmidt are shuffles within
	                           the group; tiled layout.  (4 is the dense engine's tag in bench.py / _dense.DenseInfo.) */
};
enum jb_status  { JB_OK = 0, JB_STEP_TOO_SMALL = 1, JB_TOO_MANY_STEPS = 2, JB_STIFF = 3, JB_NONFINITE = 4 };
enum jb_post    {
	JB_POST_NONE = 0,
	JB_POST_LYAP_QR = 1,         /* Gram–Schmidt of the n_lyap tangent blocks, exponents = ln(norm)/dt */
	JB_POST_LYAP_RESTRICTED = 2, /* n_lyap = 1: project out `proj` vectors, then norm */
	JB_POST_LYAP_TRANSVERSAL = 3 /* norm over the module's tangent indices */
};
enum jb_error   { JB_E_BAD_ARGUMENT = -1, JB_E_ABI = -2, JB_E_UNSUPPORTED = -3 };

typedef struct jb_module_info {
	int32_t abi_version;
	int32_t n;            /* dimension of the integrated system (incl. tangent vectors) */
	int32_t n_params;     /* control parameters per trajectory */
	int32_t n_basic;      /* dimension of the underlying system (== n for plain jitcode) */
	int32_t n_lyap;       /* tangent vectors integrated alongside */
	int32_t n_tangent;    /* transversal: number of tangent (difference) coordinates, else 0 */
	int32_t method;       /* enum jb_method this module was compiled for */
	int32_t driver;       /* enum jb_driver this module was compiled for */
	int32_t storage;      /* enum jb_storage */
	int32_t state_vectors;/* persistent solver vectors per trajectory besides Y: caller provides state[state_vectors][n][ld] */
	int32_t state_scalars;/* persistent solver scalars per trajectory: sstate[state_scalars][ld] */
	int32_t work_vectors; /* scratch vectors (global storage only): work[work_vectors][n][ld] */
	int32_t threads_per_block;
	int32_t shared_bytes;  /* dynamic shared memory per block (JB_STORE_WARP: tables + stage vectors of its warps) */
	int32_t reserved[2];
} jb_module_info;

typedef struct jb_integrate_args {
	uint32_t struct_size;     /* sizeof(jb_integrate_args), checked */
	int32_t  post;            /* enum jb_post */
	int64_t  B;               /* trajectories */
	int64_t  ld;              /* padded number of trajectories: multiple of JB_TILE, >= B */
	int32_t  n_times;         /* sample times handled by this one launch */
	int32_t  n_proj;          /* JB_POST_LYAP_RESTRICTED: number of projection vectors */
	const double *times;      /* [n_times] increasing sample times (device) */
	double   rtol;
	double   atol;            /* used when atol_vec == NULL */
	const double *atol_vec;   /* [n] per-component atol or NULL (device) */
	double   max_step;        /* <= 0: unbounded (ode: |T - t|) */
	double   first_step;      /* <= 0: chosen by the method's own rule */
	int64_t  nsteps;          /* step-attempt limit per sample time; 0 = method default */
	double   safety, ifactor, dfactor, beta;   /* Hairer controller (0 = scipy.integrate.ode's defaults) */
	int32_t  hairer_reject_by_dfactor;         /* dop853: shrink by dfactor on rejection like scipy 1.18.1's _dop */
	int32_t  lyap_discard;    /* accumulate exponents only for sample indices >= lyap_discard */
	double  *t;               /* [B]       in/out: solver time of every trajectory */
	double  *Y;               /* [n][ld]   in/out: solver state at t */
	const double *P;          /* [n_params][ld] tiled control parameters (may be NULL if n_params == 0) */
	double  *state;           /* [state_vectors][n][ld] persistent solver vectors (IVP drivers) */
	double  *sstate;          /* [state_scalars][ld]    persistent solver scalars */
	double  *work;            /* [work_vectors][n][ld]  scratch (global storage) */
	double  *Yres;            /* [n][ld] state at the last sample time (required for JB_DRV_IVP_DENSE, else optional) */
	double  *Yrec;            /* [n_times][n][ld] every sample, or NULL */
	double  *lyap_rec;        /* [n_times * (n_lyap or 1)][ld] local exponents: one (tiled / row) vector per trajectory with entry (sample, exponent) at row sample * n_out + exponent, or NULL */
	double  *lyap_acc;        /* [2][n_lyap or 1][ld] += sum and sum of squares of the local exponents of this launch, or NULL */
	const double *proj;       /* [n_proj][n_basic] orthonormalised projection vectors (restricted), or NULL */
	int32_t *status;          /* [B] enum jb_status, sticky: a failed trajectory is not advanced further */
	int64_t *n_accepted;      /* [B] += accepted steps, or NULL */
	int64_t *n_rejected;      /* [B] += rejected step attempts, or NULL (a trajectory-step = accepted + rejected) */
	unsigned long long *queue;/* [1] device scratch owned by the caller, or NULL: the counter from which persistent kernels draw
	                             trajectories (lane refill of the register kernels, work queue of the warp-storage kernels).
	                             One counter per stream with launches in flight; reset by jb_integrate itself on `stream`.
	                             NULL selects the static distribution (one thread / one strided sequence per trajectory). */
	void    *stream;          /* cudaStream_t */
} jb_integrate_args;

int         jb_abi_version(void);
const char *jb_last_error(void);
int         jb_get_info(jb_module_info *info);

/* dY(i,b) = f_i(t[b], Y(:,b), P[:,b]) in the tiled layout — the batched py_f */
int jb_eval_f(const double *t, const double *Y, const double *P, double *dY,
		int64_t B, int64_t ld, void *stream);

/* advance every trajectory through args->times (one launch) */
int jb_integrate(const jb_integrate_args *args);

/* (B,n) row-major <-> tiled trajectory-minor (device pointers), for buffers in the reference's layout;
 * n need not be the module's dimension (parameters and Lyapunov outputs use the same kernels) */
int jb_pack(const double *rows, double *soa, int64_t B, int64_t n, int64_t ld, void *stream);
int jb_unpack(const double *soa, double *rows, int64_t B, int64_t n, int64_t ld, void *stream);

/* how many kernels this module has launched since it was loaded (bench.py's gpu_launches) */
int64_t jb_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Dense sine-coupled phase oscillators (csrc/jb_dense.cu; one library, not generated per system):
 *     dy_i/dt = omega_i + c * sum_j W_ij sin(y_j - y_i)          (c: optional per-trajectory control parameter)
 * — what the reference writes out as n*deg sine terms (examples/kuramoto_network.py:19-26) — as one FP64
 * tensor-core GEMM per right-hand-side evaluation for the whole ensemble.  Replaces, for such systems, the same
 * reference entry points as jb_eval_f / jb_integrate: driver JB_DRV_ODE ("dopri5" / "dop853", ODE_wrapper.integrate,
 * integrator_tools.py:134-144) and SciPy's solve_ivp stepper with the same pairs — JB_DRV_IVP_LAND ("RK45" / "DOP853"
 * without interpolation, IVP_wrapper_no_interpolation.integrate :110-128) and JB_DRV_IVP_DENSE (the same with dense
 * output, IVP_wrapper.integrate :98-105).  Layout: plain structure of arrays, X[i*ld + b], ld a multiple of 64.
 */
typedef struct jb_dense_args {
	uint32_t struct_size;      /* sizeof(jb_dense_args), checked */
	int32_t  n;                /* oscillators */
	int64_t  B, ld;            /* trajectories, padded (multiple of 64) */
	const double *W;           /* [n][n] row-major coupling matrix (device) */
	const double *omega;       /* [n] natural frequencies (device) */
	const double *coupling;    /* [ld] per-trajectory factor c_b of the coupling sum (a control parameter: dy_i/dt = omega_i + c_b * sum_j …), or NULL for 1 */
	const double *atol_vec;    /* [n] per-component atol (device), or NULL: `atol` for all */
	double   T;                /* target time */
	double   rtol, atol, max_step, first_step;
	int64_t  nsteps;           /* step-attempt limit per call; 0 = Hairer's default */
	double   safety, ifactor, dfactor, beta;   /* 0 = scipy.integrate.ode's defaults */
	double  *t;                /* [ld] in/out */
	double  *Y;                /* [n][ld] in/out */
	double  *work;             /* [jb_dense_work_vectors()][n][ld] */
	double  *ctrl;             /* [jb_dense_control_scalars()][ld] */
	int32_t *status;           /* [ld] enum jb_status, sticky */
	int64_t *n_accepted, *n_rejected;   /* [ld] += */
	int32_t *scratch;          /* [ld/32 + 8] device scratch */
	int32_t *host_flag;        /* four int32 in (pinned) host memory: the host polls the number of unfinished trajectories (of each of the interleaved sub-ensembles) */
	int32_t  rounds_per_check; /* lock-step rounds between two polls; 0 = default (1: a round takes milliseconds, a poll microseconds) */
	int32_t  method_flags;     /* bits 0-7: enum jb_method; bit 8: dop853 shrinks by dfactor on rejection (SciPy 1.18.1); bits 9-11: number of interleaved sub-ensembles (0: default, up to four of >= 1024 trajectories each; 1: none); bits 12-13: the driver, enum jb_driver — with the solve_ivp drivers work and ctrl persist between calls; bit 14: a new state was set, steppers have to be created */
	int64_t *rounds_out;       /* host: number of lock-step rounds of this call, or NULL */
	void    *stream;
} jb_dense_args;

int         jb_dense_work_vectors(void);
int         jb_dense_control_scalars(void);
int         jb_dense_result_vector(void);   /* solve_ivp with dense output: the work vector that holds the sample at T ([n][ld], like Y) */
const char *jb_dense_last_error(void);
int64_t     jb_dense_launch_count(void);
int jb_dense_integrate(const jb_dense_args *args);
int jb_dense_eval_f(const jb_dense_args *args, double *dY);
/* C[M][N] = A[M][K] * B[K][N], row-major FP64 on the FP64 tensor cores (exported for tests) */
int jb_dense_gemm(const double *A, const double *B, double *C, int M, int N, int K, void *stream);
/* the same with the block-tile shape given instead of chosen (0: 128 x 64, 1: 112 x 64, 2: 96 x 64, 3: 64 x 64; needs even K and N) */
int jb_dense_gemm_shape(const double *A, const double *B, double *C, int M, int N, int K, int shape, void *stream);
/* out[c*ld_out + r] = in[r*ld_in + c]: (B, n) rows <-> [n][ld] */
int jb_dense_transpose(const double *in, double *out, int64_t rows, int64_t cols, int64_t ld_in, int64_t ld_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif
