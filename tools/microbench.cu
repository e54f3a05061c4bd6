// Pipe-throughput microbenchmark for sm_100a (development tool, not part of the library).
// Prints results/clk/SM for the instruction kinds the fit kernels are made of, and for
// mixes, to see which pipes are shared.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template<int OP>
__global__ void __launch_bounds__(1024) bench(double *out, long long *cycles, float seedf, double seedd) {
	double d[CHAINS];
	float f[CHAINS];
	for (int i = 0; i < CHAINS; ++i) {
		d[i] = seedd + i + threadIdx.x;
		f[i] = seedf + i + threadIdx.x;
	}
	double const a = seedd * 1.0000001, b = seedd * 0.5;
	float const af = seedf * 1.0000001f, bf = seedf * 0.5f;
	__syncthreads();
	long long t0 = clock64();
	for (int it = 0; it < ITERS; ++it) {
#pragma unroll
		for (int i = 0; i < CHAINS; ++i) {
			if (OP == 0) d[i] = fma(d[i], a, b);                       // DFMA
			if (OP == 1) d[i] = d[i] + b;                              // DADD
			if (OP == 2) d[i] = d[i] * a;                              // DMUL
			if (OP == 3) f[i] = __fmaf_rn(f[i], af, bf);               // FFMA
			if (OP == 4) { d[i] = (double) f[i]; f[i] = __int_as_float(__double2loint(d[i]) ^ __float_as_int(f[i])); } // F2F.F64.F32 (+LOP)
			if (OP == 5) { f[i] = (float) d[i]; d[i] = __hiloint2double(__double2hiint(d[i]), __float_as_int(f[i])); } // F2F.F32.F64
			if (OP == 6) { d[i] = fma(d[i], a, b); f[i] = __fmaf_rn(f[i], af, bf); } // DFMA + FFMA
			if (OP == 7) { d[i] = fma(d[i], a, (double) f[i]); f[i] = __fmaf_rn(f[i], af, bf);} // DFMA + cvt + FFMA
			if (OP == 8) { f[i] = __shfl_xor_sync(0xffffffffu, f[i], 1); }  // SHFL
			if (OP == 9) { d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, (double) f[i]); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); f[i] = __fmaf_rn(f[i], af, bf);} // 4 DFMA + 1 cvt
			if (OP == 10) { int x = __float_as_int(f[i]); int hi = ((x >> 3) & 0x0fffffff) + 0x38000000; hi |= (x & 0x80000000); d[i] = fma(d[i], a, __hiloint2double(hi, x << 29)); f[i] = __fmaf_rn(f[i], af, bf);} // manual cvt + DFMA
			if (OP == 11) { d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); f[i] = (float) d[i]; } // 2 DFMA + cvt.f32.f64
			if (OP == 12) { d[i] = d[i] / a; } // DDIV
			if (OP == 13) { d[i] = sqrt(d[i]); } // DSQRT
		}
	}
	long long t1 = clock64();
	double s = 0;
	for (int i = 0; i < CHAINS; ++i) s += d[i] + f[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template<int OP>
void run(char const *name, double ops_per_iter) {
	int const grid = 148, block = 1024;
	double *out; long long *cyc;
	cudaMalloc(&out, sizeof(double) * grid * block);
	cudaMalloc(&cyc, sizeof(long long) * grid);
	bench<OP><<<grid, block>>>(out, cyc, 1.5f, 1.25);
	cudaDeviceSynchronize();
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	bench<OP><<<grid, block>>>(out, cyc, 1.5f, 1.25);
	cudaEventRecord(e1);
	cudaDeviceSynchronize();
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
	double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
	double ops = ops_per_iter * (double) ITERS * CHAINS * block;
	printf("%-34s %8.2f results/clk/SM   (%.0f cyc, %.3f ms, eff clk %.0f MHz)\n", name, ops / avg, avg, ms, avg / ms / 1e3);
	cudaFree(out); cudaFree(cyc);
}

int main() {
	run<0>("DFMA", 1);
	run<1>("DADD", 1);
	run<2>("DMUL", 1);
	run<3>("FFMA", 1);
	run<4>("F2F.F64.F32 (+LOP)", 1);
	run<5>("F2F.F32.F64", 1);
	run<6>("DFMA+FFMA (pairs)", 1);
	run<7>("DFMA+cvt64+FFMA (triples)", 1);
	run<8>("SHFL", 1);
	run<9>("4 DFMA + 1 cvt64 + FFMA (groups)", 1);
	run<10>("manual cvt + DFMA + FFMA (groups)", 1);
	run<11>("2 DFMA + cvt32 (groups)", 1);
	run<12>("DDIV", 1);
	run<13>("DSQRT", 1);
	return 0;
}
