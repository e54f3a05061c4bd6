// Batched sinusoid baseline fit with FP64 tensor-core (DMMA) contractions; see lsqfit_sinusoid.cu.
#ifndef SAKURA_B200_LSQFIT_SINUSOID_CUH_
#define SAKURA_B200_LSQFIT_SINUSOID_CUH_

#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>

namespace sakura_b200 {

constexpr int kSinRowsPerTile = 32; // spectra fitted together by one CTA (4 DMMA row tiles)
constexpr int kSinMaxKp = 48; // fitted bases, padded to a multiple of 8
constexpr int kSinThreads = 256; // two CTAs per SM: the scalar phases of one overlap the tensor-core phases of the other
constexpr int kSinCtasPerSm = 2;

struct SinusoidFitParams {
	size_t num_spectra;
	int n; // channels (multiple of 8)
	int K; // fitted bases
	int Kp; // slots of the fit, a multiple of 8: [even bases | padding | odd bases | padding]
	int kc_p; // slots of the even group (a multiple of 8)
	int next; // columns of the extended trigonometric table: 2*mmax + 1
	int cos_only; // extended table holds only C_m (column m): Chebyshev bases T_m = cos(m t)
	// Selected basis columns (zero padded to Kp), built per call, in two row strides so that tiles
	// of 16 channels are contiguous (one bulk copy) and their shared-memory images are read
	// without bank conflicts: ld_p = Kp + 1 for the channel-contracting passes (trig sums, data
	// moments), ld_a = SinusoidLdA(Kp) for the model pass.
	double const *table_p; // [n][ld_p]
	double const *table_a; // [n][ld_a]
	int ld_p;
	int ld_a;
	int ext_ld; // columns of the trig sums kept per spectrum: SinusoidExtLd(next) = 32 * chunks
	// [ext_ld / 32][n][kSinExtChunkLd]: 1, then (sin, cos)(m theta_i) for m = 1..mmax, 32 columns
	// per chunk (+1 padding column), the columns beyond `next` are zero
	double const *ext_p;
	// [ceil(n / kSinSumStep) + 1][ext_ld]: column sums of the extended table over blocks of kSinSumStep
	// channels, last row = over all channels
	double const *ext_sums;
	// slot s holds: kind[s] = 0 constant, 1 sin, 2 cos, -1 padding; wave[s] its wave number; out_pos[s] the
	// position of its coefficient in the caller's array (-1 padding)
	int8_t kind[kSinMaxKp];
	int16_t wave[kSinMaxKp];
	int8_t out_pos[kSinMaxKp];
	float const *data;
	size_t data_stride;
	uint8_t const *mask;
	size_t mask_stride;
	float clip_threshold_sigma;
	int num_fitting_max;
	double *coeff; // may be NULL, [num_spectra][K]
	float *best_fit; // may be NULL
	float *residual; // may be NULL
	uint8_t *final_mask;
	float *rms;
	int32_t *status;
	int32_t *lsqfit_status;
	int32_t *flags; // may be NULL
	float *ws_residual; // [gridDim.x][32][n]
	float *ws_best_fit; // [gridDim.x][32][n] (when best_fit != NULL)
	uint8_t *ws_mask; // [gridDim.x][32][n] working copy of the mask
	// which two trig sums (and signs) make up each Gram matrix entry: SinusoidGramEntries(Kp) words,
	// filled by LaunchGatherTable into the buffer behind the two tables
	unsigned const *gram_table;
	int32_t *fallback_rows; // rows whose normal equations are (nearly) singular
	int32_t *fallback_count;
};

constexpr int kSinSumStep = 512; // channels per block of the extended table's precomputed column sums
constexpr int kSinExtChunkLd = 33; // row stride of one 32-column chunk of the extended table

size_t SinusoidSharedBytes(int Kp, int next);
int SinusoidExtLd(int next);
int SinusoidLdA(int Kp);
/** doubles of the two per-call tables together (table_p first, table_a at n * (Kp + 1)). */
size_t SinusoidTableDoubles(int n, int Kp);
/** 32-bit words of the Gram index table that follows the two tables in the same buffer. */
size_t SinusoidGramEntries(int Kp);

/** Gathers the selected columns of the context table into table_p and table_a and builds the Gram index
 *  table of `p` (kind / wave / K / cos_only / next must be set) behind them. */
cudaError_t LaunchGatherTable(double const *ctx_table, int nb_ctx, int n, int K, int Kp,
		int16_t const *use_idx_dev, double *tables, SinusoidFitParams const &p, cudaStream_t stream);

/** True when the kernel is instantiated for this split of the slots into an even and an odd group. */
bool SinusoidLayoutSupported(int kc_p, int Kp);

cudaError_t LaunchSinusoidFit(SinusoidFitParams const &p, int grid, cudaStream_t stream);

} // namespace sakura_b200

#endif
