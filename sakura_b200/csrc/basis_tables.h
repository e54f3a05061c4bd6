// Host-side FP64 basis tables (see basis_tables.cc).
#ifndef SAKURA_B200_BASIS_TABLES_H_
#define SAKURA_B200_BASIS_TABLES_H_

#include <stddef.h>

namespace sakura_b200 {

/** table[i*num_bases + j] = (i/(num_data-1))^j ; baseline.cc:274-289 */
void FillPolynomialTable(size_t num_data, size_t num_bases, double *table);
/** table[i*num_bases + j] = T_j(2i/(num_data-1) - 1) ; baseline.cc:299-327 */
void FillChebyshevTable(size_t num_data, size_t num_bases, double *table);
/** table[i*(2*max_nwave+1) + {0, 2k-1, 2k}] = {1, sin(k th_i), cos(k th_i)} ; baseline.cc:343-366 */
void FillSinusoidTable(size_t num_data, size_t max_nwave, double *table);

} // namespace sakura_b200

#endif
