// Host-side FP64 basis tables, built with the reference's expressions so that the
// values uploaded to the GPU are bit-identical to the reference's context tables
// (libsakura/src/baseline.cc:251-366). Compiled with -ffp-contract=off; the one
// place where the reference's Release build fuses (the Chebyshev recurrence,
// vfmsub231sd in the compiled SetBasisDataChebyshev) is written as std::fma.
#include "basis_tables.h"

#include <cmath>

namespace sakura_b200 {

// baseline.cc:251-289 -- x = i/(n-1), powers by repeated multiplication.
void FillPolynomialTable(size_t num_data, size_t num_bases, double *table) {
	double const max_x = static_cast<double>(num_data - 1);
	size_t idx = 0;
	for (size_t i = 0; i < num_data; ++i) {
		double const x = static_cast<double>(i) / max_x;
		double val = 1.0;
		for (size_t j = 0; j < num_bases; ++j) {
			table[idx++] = val;
			val *= x;
		}
	}
}

// baseline.cc:299-327 -- x = 2i/(n-1) - 1, T_j = 2x T_{j-1} - T_{j-2}.
void FillChebyshevTable(size_t num_data, size_t num_bases, double *table) {
	size_t idx = 0;
	if (num_bases == 1) {
		for (size_t i = 0; i < num_data; ++i) {
			table[idx++] = 1.0;
		}
		return;
	}
	double const max_x = static_cast<double>(num_data - 1);
	for (size_t i = 0; i < num_data; ++i) {
		table[idx++] = 1.0;
		double const x = 2.0 * static_cast<double>(i) / max_x - 1.0;
		table[idx++] = x;
		double const two_x = 2.0 * x;
		for (size_t j = 2; j < num_bases; ++j) {
			table[idx] = std::fma(two_x, table[idx - 1], -table[idx - 2]);
			++idx;
		}
	}
}

// baseline.cc:343-366 -- [1, sin(k theta_i), cos(k theta_i)], theta_i = 2 pi i/(n-1), libm sin/cos.
void FillSinusoidTable(size_t num_data, size_t max_nwave, double *table) {
	size_t idx = 0;
	if (num_data == 1) {
		table[idx] = 1.0;
		return;
	}
	double const max_x = static_cast<double>(num_data - 1);
	double const norm = 2.0 * M_PI / max_x;
	for (size_t i = 0; i < num_data; ++i) {
		table[idx++] = 1.0;
		for (size_t nwave = 1; nwave <= max_nwave; ++nwave) {
			double const x = static_cast<double>(i) * static_cast<double>(nwave) * norm;
			table[idx++] = sin(x);
			table[idx++] = cos(x);
		}
	}
}

} // namespace sakura_b200
