// accuracy of the reciprocal used by the in-register solve (poly_device.cuh::fast_rcp)
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../sakura_b200/csrc/poly_device.cuh"
__device__ double rsq2(double v) {
	double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
	double const hv = 0.5 * v;
	r = fma(r, fma(-hv * r, r, 0.5), r);
	r = fma(r, fma(-hv * r, r, 0.5), r);
	return r;
}
__global__ void k2(double *out, double *est, int n) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double x = exp((i - n / 2) * 1e-4) * (1.0 + 1e-7 * i);
	double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	est[i] = fabs(r * sqrt(x) - 1.0);
	out[i] = fabs(rsq2(x) * sqrt(x) - 1.0);
}
__global__ void k(double *out, double *est, int n) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double x = exp((i - n / 2) * 1e-3) * (1.0 + 1e-7 * i);
	double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	est[i] = fabs(fma(-x, r, 1.0));
	out[i] = fabs(fma(-x, sakura_b200::fast_rcp(x), 1.0));
}
int main() {
	int n = 1 << 20; double *o, *e; cudaMallocManaged(&o, 8 * n); cudaMallocManaged(&e, 8 * n);
	k<<<n / 256, 256>>>(o, e, n); cudaDeviceSynchronize();
	double m = 0, me = 0; for (int i = 0; i < n; ++i) { m = fmax(m, o[i]); me = fmax(me, e[i]); }
	printf("max |1 - x*rcp|: hardware estimate %.3g (2^%.1f), fast_rcp %.3g (2^%.1f)\n", me, log2(me), m, log2(m));
	k2<<<n / 256, 256>>>(o, e, n); cudaDeviceSynchronize();
	m = 0; me = 0; for (int i = 0; i < n; ++i) { m = fmax(m, o[i]); me = fmax(me, e[i]); }
	printf("max |rsqrt(x) sqrt(x) - 1|: hardware estimate %.3g (2^%.1f), two Newton steps %.3g (2^%.1f)\n", me, log2(me), m, log2(m));
	return 0;
}
