// FP64 FMA peak of the device, measured (MEASURED_PEAKS.json holds no FP64 figure; SURVEY.md §8d asks the
// builder to measure one): every thread runs 8 independent DFMA chains, so the FP64 pipe is the only limit.
#include <cuda_runtime.h>
#include <cstdint>

__global__ void __launch_bounds__(256) jb_fp64_peak_kernel(double *sink, int iterations, double a, double b)
{
	double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	for (int i = 0; i < iterations; ++i) {
		x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
		x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
	}
	const double total = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
	if (total == 12345.678) sink[0] = total;      // never true; keeps the chains alive
}

extern "C" {

// launches the kernel on `stream`; returns the number of floating-point operations it performs (2 per FMA)
double jb_fp64_peak_launch(double *sink, int blocks, int iterations, void *stream)
{
	jb_fp64_peak_kernel<<<blocks, 256, 0, (cudaStream_t) stream>>>(sink, iterations, 0.999999, 1e-9);
	return 2.0 * 8.0 * (double) iterations * 256.0 * (double) blocks;
}

}
