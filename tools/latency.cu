// Dependent-issue latency of the instruction kinds the control warp's chain is made of (one warp per SM,
// one chain): cycles per dependent instruction.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template<int OP>
__global__ void lat(double *out, long long *cycles, float seedf, double seedd) {
	double d = seedd + threadIdx.x;
	float f = seedf + threadIdx.x;
	double const a = seedd * 1.0000001, b = seedd * 0.5;
	float const af = seedf * 1.0000001f, bf = seedf * 0.5f;
	long long t0 = clock64();
#pragma unroll 16
	for (int it = 0; it < ITERS; ++it) {
		if (OP == 0) d = fma(d, a, b);
		if (OP == 1) d = d + b;
		if (OP == 2) f = __fmaf_rn(f, af, bf);
		if (OP == 3) f = (float) ((double) f + b);   // cvt up, DADD, cvt down
		if (OP == 4) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); d = r + b; }
		if (OP == 5) f = __shfl_xor_sync(0xffffffffu, f, 1) + bf;
		if (OP == 6) d = fmax(d, a) + b;
	}
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = d + f;
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template<int OP> void run(char const *name, double per_iter) {
	double *out; long long *cyc; cudaMalloc(&out, 8 * 148 * 32); cudaMalloc(&cyc, 8 * 148);
	lat<OP><<<148, 32>>>(out, cyc, 1.5f, 1.25); cudaDeviceSynchronize();
	lat<OP><<<148, 32>>>(out, cyc, 1.5f, 1.25); cudaDeviceSynchronize();
	long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
	double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
	printf("%-40s %6.1f cycles per step (%g dependent instructions)\n", name, avg / ITERS, per_iter);
	cudaFree(out); cudaFree(cyc);
}
int main() {
	run<0>("DFMA chain", 1); run<1>("DADD chain", 1); run<2>("FFMA chain", 1);
	run<3>("F2F.F64.F32 + DADD + F2F.F32.F64", 3); run<4>("MUFU.RCP64H + DADD", 2);
	run<5>("SHFL + FADD", 2); run<6>("fmax(double) + DADD", 2);
	return 0;
}
