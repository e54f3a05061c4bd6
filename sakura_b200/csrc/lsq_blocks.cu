// Standalone normal-equation building blocks on the GPU:
//   sakura_GetLSQCoefficientsDouble       (numeric_operation.cc:867-903, 192-225, 370-396)
//   sakura_UpdateLSQCoefficientsDouble    (numeric_operation.cc:905-946, 283-314, 448-479)
//   sakura_SolveSimultaneousEquationsByLUDouble (numeric_operation.cc:948-968, 609-622)
// One CTA each; every matrix/vector entry is accumulated by a single thread over the
// channels in ascending order with fma, i.e. in the reference's own summation order.
#include "lsq_blocks.cuh"
#include "lsq_device.cuh"

namespace sakura_b200 {

namespace {

constexpr int kBlockThreads = 256;

__global__ void __launch_bounds__(kBlockThreads)
NormalEquationsKernel(int n, int K, double const *table, int stride, int const *use_idx,
		uint8_t const *mask, float const *data, double *G, double *b, int *num_unmasked) {
	__shared__ double part[kBlockThreads];
	__shared__ int iscr[64];
	SmemView s;
	s.A = G;
	s.bvec = b;
	s.cvec = nullptr;
	s.cfull = nullptr;
	s.part = part;
	s.bound = nullptr;
	s.row_tr = nullptr;
	s.col_tr = nullptr;
	s.iscr = iscr;
	BuildNormalEquations(n, K, table, stride, use_idx, mask, data, s, true);
	int cnt = 0;
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		cnt += mask[i] ? 1 : 0;
	}
	cnt = block_sum_int(cnt, iscr);
	if (threadIdx.x == 0) {
		*num_unmasked = cnt;
	}
}

__global__ void __launch_bounds__(kBlockThreads)
UpdateNormalEquationsKernel(int K, double const *table, int stride, int const *use_idx,
		uint8_t const *mask, float const *data, int num_exclude,
		unsigned long long const *exclude, double *G, double *b) {
	int const ne = K * K + K;
	for (int e = threadIdx.x; e < ne; e += blockDim.x) {
		if (e < K * K) {
			int const cj = use_idx[e / K];
			int const ck = use_idx[e % K];
			double acc = G[e];
			for (int t = 0; t < num_exclude; ++t) {
				unsigned long long const i = exclude[t];
				if (mask[i]) {
					double const *row = table + i * stride;
					acc = fma(-row[cj], row[ck], acc);
				}
			}
			G[e] = acc;
		} else {
			int const ck = use_idx[e - K * K];
			double acc = b[e - K * K];
			for (int t = 0; t < num_exclude; ++t) {
				unsigned long long const i = exclude[t];
				if (mask[i]) {
					double const *row = table + i * stride;
					acc = fma(-static_cast<double>(data[i]), row[ck], acc);
				}
			}
			b[e - K * K] = acc;
		}
	}
}

// A (K*K, overwritten), bvec (K, overwritten), x (K), tr (2K ints) live in global memory.
__global__ void __launch_bounds__(kBlockThreads)
SolveKernel(int K, double *A, double *bvec, double *x, int *tr) {
	__shared__ double part[kBlockThreads];
	__shared__ int iscr[64];
	SmemView s;
	s.A = A;
	s.bvec = bvec;
	s.cvec = x;
	s.cfull = nullptr;
	s.part = part;
	s.bound = nullptr;
	s.row_tr = tr;
	s.col_tr = tr + K;
	s.iscr = iscr;
	SolveFullPivLU(K, s);
}

} // namespace

cudaError_t LaunchNormalEquations(int n, int K, double const *table, int stride, int const *use_idx,
		uint8_t const *mask, float const *data, double *G, double *b, int *num_unmasked,
		cudaStream_t stream) {
	NormalEquationsKernel<<<1, kBlockThreads, 0, stream>>>(n, K, table, stride, use_idx, mask, data,
			G, b, num_unmasked);
	return cudaGetLastError();
}

cudaError_t LaunchUpdateNormalEquations(int K, double const *table, int stride, int const *use_idx,
		uint8_t const *mask, float const *data, int num_exclude, unsigned long long const *exclude,
		double *G, double *b, cudaStream_t stream) {
	UpdateNormalEquationsKernel<<<1, kBlockThreads, 0, stream>>>(K, table, stride, use_idx, mask,
			data, num_exclude, exclude, G, b);
	return cudaGetLastError();
}

cudaError_t LaunchSolve(int K, double *A, double *bvec, double *x, int *tr, cudaStream_t stream) {
	SolveKernel<<<1, kBlockThreads, 0, stream>>>(K, A, bvec, x, tr);
	return cudaGetLastError();
}

} // namespace sakura_b200
