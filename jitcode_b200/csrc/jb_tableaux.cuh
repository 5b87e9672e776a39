// Butcher tableaux of the two explicit pairs the reference reaches through SciPy
// (SURVEY.md §8a a4/a7): Dormand–Prince 5(4) ("dopri5" / "RK45") and Hairer's 8(5,3)
// ("dop853" / "DOP853").  Published coefficients; DOP853's come from the generated header.
#pragma once

#define JB_TABLE static constexpr

#include "jb_dop853_coeffs.cuh"

namespace jb {

// Row s of A combines K[0..s-1] into the input of stage s; row NSTAGE is the solution
// weights b (first-same-as-last: K[NSTAGE] = f(t+h, y_new)).
struct DP5 {
	static constexpr int NSTAGE = 6;        // RHS evaluations per attempt (FSAL)
	static constexpr int KROWS = 7;         // K[0..6]
	static constexpr int KROWS_DENSE = 7;   // no extra stages for dense output
	static constexpr int ERR_ORDER = 4;     // error estimator order → exponent -1/5
	static constexpr int HAIRER_ORDER = 5;  // iord of HINIT
	static constexpr double STIFF_LIMIT = 3.25;   // dopri5.f: HLAMB > 3.25 counts as a sign of stiffness
	static constexpr int NDENSE = 4;        // dense-output vectors persisted between calls
	static constexpr bool ERR_NEEDS_FSAL = true;
	JB_TABLE double C[7] = { 0.0, 1.0/5, 3.0/10, 4.0/5, 8.0/9, 1.0, 1.0 };
	JB_TABLE double A[7][7] = {
		{ 0, 0, 0, 0, 0, 0, 0 },
		{ 1.0/5, 0, 0, 0, 0, 0, 0 },
		{ 3.0/40, 9.0/40, 0, 0, 0, 0, 0 },
		{ 44.0/45, -56.0/15, 32.0/9, 0, 0, 0, 0 },
		{ 19372.0/6561, -25360.0/2187, 64448.0/6561, -212.0/729, 0, 0, 0 },
		{ 9017.0/3168, -355.0/33, 46732.0/5247, 49.0/176, -5103.0/18656, 0, 0 },
		{ 35.0/384, 0, 500.0/1113, 125.0/192, -2187.0/6784, 11.0/84, 0 },
	};
	JB_TABLE double E[7] = { -71.0/57600, 0, 71.0/16695, -71.0/1920, 17253.0/339200, -22.0/525, 1.0/40 };
	// quartic dense output (Shampine), scipy/integrate/_ivp/rk.py:554-565
	JB_TABLE double P[7][4] = {
		{ 1.0, -8048581381.0/2820520608.0, 8663915743.0/2820520608.0, -12715105075.0/11282082432.0 },
		{ 0, 0, 0, 0 },
		{ 0, 131558114200.0/32700410799.0, -68118460800.0/10900136933.0, 87487479700.0/32700410799.0 },
		{ 0, -1754552775.0/470086768.0, 14199869525.0/1410260304.0, -10690763975.0/1880347072.0 },
		{ 0, 127303824393.0/49829197408.0, -318862633887.0/49829197408.0, 701980252875.0/199316789632.0 },
		{ 0, -282668133.0/205662961.0, 2019193451.0/616988883.0, -1453857185.0/822651844.0 },
		{ 0, 40617522.0/29380423.0, -110615467.0/29380423.0, 69997945.0/29380423.0 },
	};
};

// Bogacki–Shampine 3(2) ("RK23" of scipy.integrate.solve_ivp, scipy/integrate/_ivp/rk.py:183-276); published coefficients
struct RK23 {
	static constexpr int NSTAGE = 3;
	static constexpr int KROWS = 4;
	static constexpr int KROWS_DENSE = 4;
	static constexpr int ERR_ORDER = 2;     // exponent -1/3
	static constexpr int HAIRER_ORDER = 3;  // (no Hairer code of this order; only the solve_ivp drivers use the pair)
	static constexpr double STIFF_LIMIT = 1e300;
	static constexpr int NDENSE = 3;        // cubic dense output
	static constexpr bool ERR_NEEDS_FSAL = true;
	JB_TABLE double C[4] = { 0.0, 1.0/2, 3.0/4, 1.0 };
	JB_TABLE double A[4][4] = {
		{ 0, 0, 0, 0 },
		{ 1.0/2, 0, 0, 0 },
		{ 0, 3.0/4, 0, 0 },
		{ 2.0/9, 1.0/3, 4.0/9, 0 },
	};
	JB_TABLE double E[4] = { 5.0/72, -1.0/12, -1.0/9, 1.0/8 };
	JB_TABLE double P[4][3] = {
		{ 1.0, -4.0/3, 5.0/9 },
		{ 0, 1.0, -2.0/3 },
		{ 0, 4.0/3, -8.0/9 },
		{ 0, -1.0, 1.0 },
	};
};

struct DOP853 {
	static constexpr int NSTAGE = 12;
	static constexpr int KROWS = 13;
	static constexpr int KROWS_DENSE = 16;  // three extra stages for the 7th-order dense output
	static constexpr int ERR_ORDER = 7;
	static constexpr int HAIRER_ORDER = 8;
	static constexpr double STIFF_LIMIT = 6.1;    // dop853.f
	static constexpr int NDENSE = 7;
	static constexpr bool ERR_NEEDS_FSAL = false;  // error uses K[0..11] only (Hairer evaluates it before f(t+h,y_new))
};

// compile-time loop with the index available as a constant expression
template <int I> struct Idx {
	static constexpr int value = I;
	__host__ __device__ constexpr operator int() const { return I; }
};

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f)
{
	if constexpr (I < N) {
		f(Idx<I>{});
		static_for<I + 1, N>(f);
	}
}

}  // namespace jb
