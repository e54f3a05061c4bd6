// Interpolation kernels: one thread per output element (point, array).
//
// The reference walks every array with a merge of the two sorted position lists
// (interpolation.cc:164-232, 1100-1230). What it computes for one point x of one array is a
// function of x alone: with b_0 < ... < b_{m-1} the positions of that array's VALID base points,
//   m == 0            -> mask false, data untouched        (:1182-1185)
//   x <  b_0          -> y_0   (no extrapolation)            (:1207-1209)
//   x >= b_{m-1}      -> y_{m-1}                             (:1222-1225)
//   b_j <= x < b_{j+1}-> nearest: x > (b_{j+1} + b_j)/2 ? y_{j+1} : y_j   (:625-637)
//                        linear : (float)(y_j + dydx (x - b_j)), dydx = (double)(y_{j+1} - y_j [float])
//                                 / (b_{j+1} - b_j), the last step fused as compiled      (:691-704)
// so each thread finds j by bisection over all base positions and steps over flagged points.
// Descending base positions are read back to front (the reference copies them reversed, :905-922).
#include "interpolation.cuh"

namespace sakura_b200 {

namespace {

struct BaseView {
	InterpolateParams const &p;
	bool ascending;
	__device__ __forceinline__ double position(size_t k) const {
		return p.base_position[ascending ? k : p.num_base - 1 - k];
	}
	// element of array a at ascending base index k
	__device__ __forceinline__ size_t element(size_t k, size_t a) const {
		size_t const kb = ascending ? k : p.num_base - 1 - k;
		return p.y_axis ? kb * p.num_array + a : a * p.num_base + kb;
	}
};

/** number of base positions <= x (bisection over all base points, flagged or not) */
__device__ __forceinline__ size_t CountNotAbove(BaseView const &v, double x) {
	size_t lo = 0, hi = v.p.num_base;
	while (lo < hi) {
		size_t const mid = (lo + hi) / 2;
		if (v.position(mid) <= x) {
			lo = mid + 1;
		} else {
			hi = mid;
		}
	}
	return lo;
}

/** one output element: array a at position x, `lo` = CountNotAbove(x); false: no valid base point */
__device__ __forceinline__ bool InterpolateOne(BaseView const &v, size_t a, double x, size_t lo, float *out) {
	size_t const nb = v.p.num_base;
	// nearest valid base point at or below x, and above x
	long long kl = static_cast<long long>(lo) - 1;
	while (kl >= 0 && v.p.base_mask[v.element(static_cast<size_t>(kl), a)] == 0) {
		--kl;
	}
	size_t kr = lo;
	while (kr < nb && v.p.base_mask[v.element(kr, a)] == 0) {
		++kr;
	}
	if (kl < 0 && kr >= nb) {
		return false;
	}
	if (kl < 0) {
		*out = v.p.base_data[v.element(kr, a)];
	} else if (kr >= nb) {
		*out = v.p.base_data[v.element(static_cast<size_t>(kl), a)];
	} else {
		float const y0 = v.p.base_data[v.element(static_cast<size_t>(kl), a)];
		float const y1 = v.p.base_data[v.element(kr, a)];
		double const x0 = v.position(static_cast<size_t>(kl));
		double const x1 = v.position(kr);
		if (v.p.linear) {
			double const dydx = __ddiv_rn(static_cast<double>(__fsub_rn(y1, y0)), __dsub_rn(x1, x0));
			*out = static_cast<float>(fma(dydx, __dsub_rn(x, x0), static_cast<double>(y0)));
		} else {
			double const midpoint = __ddiv_rn(__dadd_rn(x1, x0), 2.0);
			*out = (x > midpoint) ? y1 : y0;
		}
	}
	return true;
}

// Y layout ([point][array]): threads along the arrays (coalesced), blocks in y stride over the points;
// the position and its bisection are uniform within a block.
__global__ void __launch_bounds__(256)
InterpolateYKernel(InterpolateParams const p) {
	BaseView const v = { p, p.base_position[p.num_base - 1] > p.base_position[0] }; // :1128-1129
	size_t const a = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (a >= p.num_array) {
		return;
	}
	for (size_t i = blockIdx.y; i < p.num_interpolated; i += gridDim.y) {
		double const x = p.interpolated_position[i];
		size_t const lo = CountNotAbove(v, x);
		size_t const e = i * p.num_array + a;
		float value;
		if (InterpolateOne(v, a, x, lo, &value)) {
			p.interpolated_data[e] = value;
			p.interpolated_mask[e] = 1;
		} else {
			p.interpolated_mask[e] = 0; // no valid base point: the data is left as it is
		}
	}
}

// Y layout, few base points (the Tsys / sky case: a handful of measurements, many points): every thread
// keeps its array's interval table in shared memory -- for each gap k between consecutive base
// positions (k = number of base positions <= x, 0..num_base) the valid neighbours' (x0, y0, y1) and the
// slope dydx, computed ONCE per array with the reference's operations (:691-704) -- so that a point costs
// one uniform bisection, three shared-memory loads and one fused multiply-add.
constexpr int kTableThreads = 128;
constexpr int kTableMaxBase = 15;

struct IntervalEntry {
	double a; // linear: dydx; nearest: midpoint
	double x0;
	float y0;
	float y1; // nearest: right value; linear: 1.0f marks "no interpolation, the value is y0"
};

__global__ void __launch_bounds__(kTableThreads)
InterpolateYTableKernel(InterpolateParams const p) {
	extern __shared__ __align__(16) unsigned char table_raw[];
	IntervalEntry *table = reinterpret_cast<IntervalEntry *>(table_raw); // [num_base + 1][kTableThreads]
	BaseView const v = { p, p.base_position[p.num_base - 1] > p.base_position[0] };
	size_t const a = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	size_t const nb = p.num_base;
	bool any_valid = false;
	for (size_t lo = 0; lo <= nb && a < p.num_array; ++lo) {
		long long kl = static_cast<long long>(lo) - 1;
		while (kl >= 0 && p.base_mask[v.element(static_cast<size_t>(kl), a)] == 0) {
			--kl;
		}
		size_t kr = lo;
		while (kr < nb && p.base_mask[v.element(kr, a)] == 0) {
			++kr;
		}
		IntervalEntry e;
		e.a = p.linear ? 0.0 : __longlong_as_double(0x7ff0000000000000LL); // nearest: midpoint +Inf -> y0
		e.x0 = 0.0;
		e.y0 = 0.0f;
		e.y1 = 1.0f;
		if (kl >= 0 || kr < nb) {
			any_valid = true;
			if (kl < 0) {
				e.y0 = p.base_data[v.element(kr, a)];
			} else if (kr >= nb) {
				e.y0 = p.base_data[v.element(static_cast<size_t>(kl), a)];
			} else {
				float const y0 = p.base_data[v.element(static_cast<size_t>(kl), a)];
				float const y1 = p.base_data[v.element(kr, a)];
				double const x0 = v.position(static_cast<size_t>(kl));
				double const x1 = v.position(kr);
				e.x0 = x0;
				e.y0 = y0;
				if (p.linear) {
					e.a = __ddiv_rn(static_cast<double>(__fsub_rn(y1, y0)), __dsub_rn(x1, x0));
					e.y1 = 0.0f;
				} else {
					e.a = __ddiv_rn(__dadd_rn(x1, x0), 2.0);
					e.y1 = y1;
				}
			}
		}
		table[lo * kTableThreads + threadIdx.x] = e;
	}
	// The points of this block row by row, kTableThreads at a time: first every thread locates one
	// point (position and gap index go to shared memory), then every thread walks the located
	// points for its own array.
	bool const linear = p.linear != 0;
	double *sx = reinterpret_cast<double *>(table + (nb + 1) * kTableThreads); // [kTableThreads]
	int *slo = reinterpret_cast<int *>(sx + kTableThreads); // [kTableThreads]
	size_t const per_block = (p.num_interpolated + gridDim.y - 1) / gridDim.y;
	size_t const i_begin = per_block * blockIdx.y;
	size_t const i_end = (i_begin + per_block < p.num_interpolated) ? i_begin + per_block : p.num_interpolated;
	for (size_t base = i_begin; base < i_end; base += kTableThreads) {
		__syncthreads();
		if (base + threadIdx.x < i_end) {
			double const x = p.interpolated_position[base + threadIdx.x];
			sx[threadIdx.x] = x;
			slo[threadIdx.x] = static_cast<int>(CountNotAbove(v, x));
		}
		__syncthreads();
		if (a >= p.num_array) {
			continue;
		}
		int const count = static_cast<int>((i_end - base < kTableThreads) ? i_end - base : kTableThreads);
		float *out = p.interpolated_data + base * p.num_array + a;
		uint8_t *outm = p.interpolated_mask + base * p.num_array + a;
		if (!any_valid) {
			for (int r = 0; r < count; ++r) {
				outm[static_cast<size_t>(r) * p.num_array] = 0; // the data is left as it is
			}
			continue;
		}
#pragma unroll 4
		for (int r = 0; r < count; ++r) {
			double const x = sx[r];
			IntervalEntry const t = table[slo[r] * kTableThreads + threadIdx.x];
			float value;
			if (linear) {
				value = (t.y1 != 0.0f) ? t.y0 :
						static_cast<float>(fma(t.a, __dsub_rn(x, t.x0), static_cast<double>(t.y0)));
			} else {
				value = (x > t.a) ? t.y1 : t.y0;
			}
			out[static_cast<size_t>(r) * p.num_array] = value;
			outm[static_cast<size_t>(r) * p.num_array] = 1;
		}
	}
}

// X layout ([array][point]): threads along the points, blocks in y stride over the arrays
__global__ void __launch_bounds__(256)
InterpolateXKernel(InterpolateParams const p) {
	BaseView const v = { p, p.base_position[p.num_base - 1] > p.base_position[0] };
	size_t const i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i >= p.num_interpolated) {
		return;
	}
	double const x = p.interpolated_position[i];
	size_t const lo = CountNotAbove(v, x);
	for (size_t a = blockIdx.y; a < p.num_array; a += gridDim.y) {
		size_t const e = a * p.num_interpolated + i;
		float value;
		if (InterpolateOne(v, a, x, lo, &value)) {
			p.interpolated_data[e] = value;
			p.interpolated_mask[e] = 1;
		} else {
			p.interpolated_mask[e] = 0;
		}
	}
}

} // namespace

cudaError_t LaunchInterpolate(InterpolateParams const &p, int num_sms, cudaStream_t stream) {
	if (p.num_interpolated == 0 || p.num_array == 0 || p.num_base == 0) {
		return cudaSuccess;
	}
	size_t const fast = p.y_axis ? p.num_array : p.num_interpolated;
	size_t const slow = p.y_axis ? p.num_interpolated : p.num_array;
	size_t const bx = (fast + 255) / 256;
	if (bx > 0x7fffffffu) {
		return cudaErrorInvalidValue;
	}
	size_t by = static_cast<size_t>(num_sms) * 16 / bx + 1;
	if (by > slow) {
		by = slow;
	}
	if (by > 65535) {
		by = 65535;
	}
	if (p.y_axis && p.num_base <= static_cast<size_t>(kTableMaxBase) && p.num_interpolated >= 64) {
		size_t const tbx = (p.num_array + kTableThreads - 1) / kTableThreads;
		size_t tby = static_cast<size_t>(num_sms) * 8 / tbx + 1;
		if (tby > p.num_interpolated / 32) { // enough points per thread to pay for its table
			tby = p.num_interpolated / 32;
		}
		if (tby > 65535) {
			tby = 65535;
		}
		size_t const smem = (p.num_base + 1) * kTableThreads * sizeof(IntervalEntry)
				+ kTableThreads * (sizeof(double) + sizeof(int));
		// 15 base points need 50 688 bytes: above the 48 KiB a kernel gets without opting in (set on every
		// launch: the attribute is per device and a process may drive several)
		cudaError_t const attr = cudaFuncSetAttribute(InterpolateYTableKernel,
				cudaFuncAttributeMaxDynamicSharedMemorySize,
				static_cast<int>((kTableMaxBase + 1) * kTableThreads * sizeof(IntervalEntry)
						+ kTableThreads * (sizeof(double) + sizeof(int))));
		if (attr != cudaSuccess) {
			return attr;
		}
		InterpolateYTableKernel<<<dim3(static_cast<unsigned>(tbx), static_cast<unsigned>(tby)), kTableThreads,
				smem, stream>>>(p);
		return cudaGetLastError();
	}
	dim3 const grid(static_cast<unsigned>(bx), static_cast<unsigned>(by));
	if (p.y_axis) {
		InterpolateYKernel<<<grid, 256, 0, stream>>>(p);
	} else {
		InterpolateXKernel<<<grid, 256, 0, stream>>>(p);
	}
	return cudaGetLastError();
}

} // namespace sakura_b200
