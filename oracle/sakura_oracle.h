/* TEST INFRASTRUCTURE ONLY (oracle). Declarations of the plain-C restatement
 * (oracle/sakura_oracle.c). */
#ifndef SAKURA_B200_ORACLE_H_
#define SAKURA_B200_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

/* same layout as sakura_StatisticsResultFloat (libsakura/src/libsakura/sakura.h:208-237) */
typedef struct {
	size_t count;
	double sum;
	double square_sum;
	float min;
	float max;
	long index_of_min;
	long index_of_max;
} OracleStatistics;

void oracle_compute_accurate_statistics(size_t num_data, float const *data,
		unsigned char const *mask, OracleStatistics *result);
void oracle_compute_statistics(size_t num_data, float const *data, unsigned char const *mask,
		OracleStatistics *result);

#endif
