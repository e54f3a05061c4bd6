// Interpolation kernel: one thread per output element (point, array).
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

__global__ void __launch_bounds__(256)
InterpolateKernel(InterpolateParams const p) {
	size_t const nb = p.num_base;
	size_t const total = p.num_interpolated * p.num_array;
	bool const ascending = p.base_position[nb - 1] > p.base_position[0]; // :1128-1129
	size_t const stride = static_cast<size_t>(gridDim.x) * blockDim.x;
	for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
		size_t i, a;
		if (p.y_axis) {
			i = e / p.num_array;
			a = e - i * p.num_array;
		} else {
			a = e / p.num_interpolated;
			i = e - a * p.num_interpolated;
		}
		auto position = [&](size_t k) {
			return p.base_position[ascending ? k : nb - 1 - k];
		};
		auto element = [&](size_t k) {
			size_t const kb = ascending ? k : nb - 1 - k;
			return p.y_axis ? kb * p.num_array + a : a * nb + kb;
		};
		double const x = p.interpolated_position[i];
		// number of base positions <= x
		size_t lo = 0, hi = nb;
		while (lo < hi) {
			size_t const mid = (lo + hi) / 2;
			if (position(mid) <= x) {
				lo = mid + 1;
			} else {
				hi = mid;
			}
		}
		// nearest valid base point at or below x, and above x
		long long kl = static_cast<long long>(lo) - 1;
		while (kl >= 0 && p.base_mask[element(static_cast<size_t>(kl))] == 0) {
			--kl;
		}
		size_t kr = lo;
		while (kr < nb && p.base_mask[element(kr)] == 0) {
			++kr;
		}
		if (kl < 0 && kr >= nb) {
			p.interpolated_mask[e] = 0; // no valid base point: the data is left as it is
			continue;
		}
		float v;
		if (kl < 0) {
			v = p.base_data[element(kr)];
		} else if (kr >= nb) {
			v = p.base_data[element(static_cast<size_t>(kl))];
		} else {
			float const y0 = p.base_data[element(static_cast<size_t>(kl))];
			float const y1 = p.base_data[element(kr)];
			double const x0 = position(static_cast<size_t>(kl));
			double const x1 = position(kr);
			if (p.linear) {
				double const dydx = __ddiv_rn(static_cast<double>(__fsub_rn(y1, y0)), __dsub_rn(x1, x0));
				v = static_cast<float>(fma(dydx, __dsub_rn(x, x0), static_cast<double>(y0)));
			} else {
				double const midpoint = __ddiv_rn(__dadd_rn(x1, x0), 2.0);
				v = (x > midpoint) ? y1 : y0;
			}
		}
		p.interpolated_data[e] = v;
		p.interpolated_mask[e] = 1;
	}
}

} // namespace

cudaError_t LaunchInterpolate(InterpolateParams const &p, int num_sms, cudaStream_t stream) {
	size_t const total = p.num_interpolated * p.num_array;
	if (total == 0 || p.num_base == 0) {
		return cudaSuccess;
	}
	size_t blocks = (total + 255) / 256;
	size_t const cap = static_cast<size_t>(num_sms) * 16;
	if (blocks > cap) {
		blocks = cap;
	}
	InterpolateKernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
	return cudaGetLastError();
}

} // namespace sakura_b200
