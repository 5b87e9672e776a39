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

// FP64 tensor-core peak: every warp runs 8 independent chains of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; 512 flop per
// warp instruction) — the denominator for the dense engine's GEMM.
__global__ void __launch_bounds__(256) jb_dmma_peak_kernel(double *sink, int iterations, double a, double b)
{
	double c[8][2];
#pragma unroll
	for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = threadIdx.x + k;
	for (int i = 0; i < iterations; ++i) {
#pragma unroll
		for (int k = 0; k < 8; ++k)
			asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
					: "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
	}
	double total = 0.0;
#pragma unroll
	for (int k = 0; k < 8; ++k) total += c[k][0] + c[k][1];
	if (total == 12345.678) sink[0] = total;
}

extern "C" {

double jb_dmma_peak_launch(double *sink, int blocks, int iterations, void *stream)
{
	jb_dmma_peak_kernel<<<blocks, 256, 0, (cudaStream_t) stream>>>(sink, iterations, 1e-9, 1e-9);
	return 512.0 * 8.0 * (double) iterations * 8.0 * (double) blocks;      // 8 warps per block
}

// launches the kernel on `stream`; returns the number of floating-point operations it performs (2 per FMA)
double jb_fp64_peak_launch(double *sink, int blocks, int iterations, void *stream)
{
	jb_fp64_peak_kernel<<<blocks, 256, 0, (cudaStream_t) stream>>>(sink, iterations, 0.999999, 1e-9);
	return 2.0 * 8.0 * (double) iterations * 256.0 * (double) blocks;
}

}
