// C ABI of libsakura_b200.so, part 4: statistics and the standalone normal-equation building blocks.
#include "capi_internal.cuh"

using namespace sakura_b200;
using namespace sakura_b200::capi;

// ---------------------------------------------------------------- statistics

namespace {
struct StatScratch {
	DeviceBuffer data, mask, result, partial;
};
thread_local StatScratch *g_stat_scratch = nullptr;

sakura_Status RunStatsDevice(size_t num_spectra, size_t num_data, float const *d_data,
		size_t data_stride, uint8_t const *d_mask, size_t mask_stride, StatResult *d_result,
		DeviceBuffer *partial, cudaStream_t stream) {
	DeviceInfo info;
	sakura_Status st = CurrentDevice(&info);
	if (st != sakura_Status_kOK) {
		return st;
	}
	size_t chunks, chunk_len;
	size_t const pbytes = StatsPartialBytes(num_spectra, num_data, &chunks, &chunk_len);
	if (pbytes > 0) {
		CUDA_TRY(partial->Reserve(pbytes));
	}
	int launches = 0;
	cudaError_t e = LaunchStats(num_spectra, num_data, d_data, data_stride, d_mask, mask_stride,
			d_result, partial->ptr, info.num_sms, stream, &launches);
	g_launch_count += launches;
	g_last_kernel = "statistics";
	if (e != cudaSuccess) {
		SetError("LaunchStats", e);
		return sakura_Status_kUnknownError;
	}
	return sakura_Status_kOK;
}
} // namespace

sakura_Status sakura_ComputeStatisticsFloatBatchDevice(size_t num_spectra, size_t num_data,
		float const d_data[], size_t data_stride, bool const d_is_valid[], size_t mask_stride,
		sakura_StatisticsResultFloat d_result[], void *cuda_stream) noexcept {
	CHECK_ARGS(num_data <= static_cast<size_t>(INT32_MAX)); // statistics.cc:855
	CHECK_ARGS(d_data != nullptr && d_is_valid != nullptr && d_result != nullptr);
	if (g_stat_scratch == nullptr) {
		g_stat_scratch = new StatScratch();
	}
	return RunStatsDevice(num_spectra, num_data, d_data, data_stride,
			reinterpret_cast<uint8_t const *>(d_is_valid), mask_stride,
			reinterpret_cast<StatResult *>(d_result), &g_stat_scratch->partial,
			static_cast<cudaStream_t>(cuda_stream));
}

sakura_Status sakura_ComputeStatisticsFloatBatch(size_t num_spectra, size_t num_data,
		float const data[], size_t data_stride, bool const is_valid[], size_t mask_stride,
		sakura_StatisticsResultFloat result[]) noexcept {
	CHECK_ARGS(num_data <= static_cast<size_t>(INT32_MAX));
	CHECK_ARGS(data != nullptr);
	CHECK_ARGS(is_valid != nullptr);
	CHECK_ARGS(RowsAligned(data, data_stride * sizeof(float), num_spectra));
	CHECK_ARGS(RowsAligned(is_valid, mask_stride, num_spectra));
	CHECK_ARGS(result != nullptr);
	CHECK_ARGS(num_spectra <= 1 || (data_stride >= num_data && mask_stride >= num_data));
	if (num_spectra == 0) {
		return sakura_Status_kOK;
	}
	static_assert(sizeof(sakura_StatisticsResultFloat) == sizeof(StatResult), "layout");
	if (g_stat_scratch == nullptr) {
		g_stat_scratch = new StatScratch();
	}
	StatScratch &s = *g_stat_scratch;
	cudaStream_t const stream = 0;
	size_t const n = num_data;
	size_t const ds = RoundUp(n > 0 ? n : 1, 32);
	CUDA_TRY(s.data.Reserve(num_spectra * ds * sizeof(float)));
	CUDA_TRY(s.mask.Reserve(num_spectra * ds));
	CUDA_TRY(s.result.Reserve(num_spectra * sizeof(StatResult)));
	if (n > 0) {
		CUDA_TRY(cudaMemcpy2DAsync(s.data.ptr, ds * sizeof(float), data, data_stride * sizeof(float),
				n * sizeof(float), num_spectra, cudaMemcpyHostToDevice, stream));
		CUDA_TRY(cudaMemcpy2DAsync(s.mask.ptr, ds, is_valid, mask_stride, n, num_spectra,
				cudaMemcpyHostToDevice, stream));
	}
	sakura_Status st = RunStatsDevice(num_spectra, n, s.data.As<float>(), ds, s.mask.As<uint8_t>(),
			ds, s.result.As<StatResult>(), &s.partial, stream);
	if (st != sakura_Status_kOK) {
		return st;
	}
	CUDA_TRY(cudaMemcpyAsync(result, s.result.ptr, num_spectra * sizeof(StatResult),
			cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	return sakura_Status_kOK;
}

sakura_Status sakura_ComputeStatisticsFloat(size_t num_data, float const data[],
		bool const is_valid[], sakura_StatisticsResultFloat *result) noexcept {
	return sakura_ComputeStatisticsFloatBatch(1, num_data, data, num_data, is_valid, num_data,
			result);
}

sakura_Status sakura_ComputeAccurateStatisticsFloat(size_t num_data, float const data[],
		bool const is_valid[], sakura_StatisticsResultFloat *result) noexcept {
	return sakura_ComputeStatisticsFloatBatch(1, num_data, data, num_data, is_valid, num_data,
			result);
}


// ---------------------------------------------------------------- normal-equation building blocks

namespace {
struct BlockScratch {
	DeviceBuffer data, mask, basis, use_idx, G, b, x, tr, exclude, count;
};
thread_local BlockScratch *g_block_scratch = nullptr;

BlockScratch &BlockScratchGet() {
	if (g_block_scratch == nullptr) {
		g_block_scratch = new BlockScratch();
	}
	return *g_block_scratch;
}

sakura_Status UploadModel(BlockScratch &s, size_t num_data, float const *data, bool const *mask,
		size_t num_model_bases, double const *basis, size_t num_lsq_bases, size_t const *use_idx,
		cudaStream_t stream) {
	CUDA_TRY(s.data.Reserve(num_data * sizeof(float)));
	CUDA_TRY(s.mask.Reserve(num_data));
	CUDA_TRY(s.basis.Reserve(num_data * num_model_bases * sizeof(double)));
	CUDA_TRY(s.use_idx.Reserve(num_lsq_bases * sizeof(int)));
	CUDA_TRY(s.G.Reserve(num_lsq_bases * num_lsq_bases * sizeof(double)));
	CUDA_TRY(s.b.Reserve(num_lsq_bases * sizeof(double)));
	CUDA_TRY(s.count.Reserve(sizeof(int)));
	std::vector<int> idx(num_lsq_bases);
	for (size_t i = 0; i < num_lsq_bases; ++i) {
		idx[i] = static_cast<int>(use_idx[i]);
	}
	CUDA_TRY(cudaMemcpyAsync(s.data.ptr, data, num_data * sizeof(float), cudaMemcpyHostToDevice,
			stream));
	CUDA_TRY(cudaMemcpyAsync(s.mask.ptr, mask, num_data, cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(s.basis.ptr, basis, num_data * num_model_bases * sizeof(double),
			cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpy(s.use_idx.ptr, idx.data(), num_lsq_bases * sizeof(int),
			cudaMemcpyHostToDevice));
	return sakura_Status_kOK;
}
} // namespace

sakura_Status sakura_GetLSQCoefficientsDouble(size_t const num_data, float const data[],
		bool const mask[], size_t const num_model_bases, double const basis_data[],
		size_t const num_lsq_bases, size_t const use_bases_idx[], double lsq_matrix[],
		double lsq_vector[]) noexcept {
	CHECK_ARGS(num_data != 0 && num_data <= static_cast<size_t>(INT32_MAX));
	CHECK_ARGS(data != nullptr && IsAligned(data));
	CHECK_ARGS(mask != nullptr && IsAligned(mask));
	CHECK_ARGS(num_model_bases != 0);
	CHECK_ARGS(num_model_bases <= num_data);
	CHECK_ARGS(basis_data != nullptr && IsAligned(basis_data));
	CHECK_ARGS(num_lsq_bases != 0);
	CHECK_ARGS(num_lsq_bases <= num_model_bases);
	CHECK_ARGS(use_bases_idx != nullptr && IsAligned(use_bases_idx));
	CHECK_ARGS(lsq_matrix != nullptr && IsAligned(lsq_matrix));
	CHECK_ARGS(lsq_vector != nullptr && IsAligned(lsq_vector));
	for (size_t i = 0; i < num_lsq_bases; ++i) {
		CHECK_ARGS(use_bases_idx[i] < num_model_bases);
	}
	BlockScratch &s = BlockScratchGet();
	cudaStream_t const stream = 0;
	sakura_Status st = UploadModel(s, num_data, data, mask, num_model_bases, basis_data,
			num_lsq_bases, use_bases_idx, stream);
	if (st != sakura_Status_kOK) {
		return st;
	}
	cudaError_t e = LaunchNormalEquations(static_cast<int>(num_data), static_cast<int>(num_lsq_bases),
			s.basis.As<double>(), static_cast<int>(num_model_bases), s.use_idx.As<int>(),
			s.mask.As<uint8_t>(), s.data.As<float>(), s.G.As<double>(), s.b.As<double>(),
			s.count.As<int>(), stream);
	g_launch_count += 1;
	g_last_kernel = "normal_equations";
	if (e != cudaSuccess) {
		SetError("LaunchNormalEquations", e);
		return sakura_Status_kUnknownError;
	}
	int count = 0;
	CUDA_TRY(cudaMemcpyAsync(lsq_matrix, s.G.ptr, num_lsq_bases * num_lsq_bases * sizeof(double),
			cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaMemcpyAsync(lsq_vector, s.b.ptr, num_lsq_bases * sizeof(double),
			cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaMemcpyAsync(&count, s.count.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	// numeric_operation.cc:221-224 (templates, <= 100 bases) / :266-269 (generic)
	size_t const need = (num_lsq_bases <= 100) ? num_lsq_bases : num_model_bases;
	if (static_cast<size_t>(count) < need) {
		return sakura_Status_kNG;
	}
	return sakura_Status_kOK;
}

sakura_Status sakura_UpdateLSQCoefficientsDouble(size_t const num_data, float const data[],
		bool const mask[], size_t const num_exclude_indices, size_t const exclude_indices[],
		size_t const num_model_bases, double const basis_data[], size_t const num_lsq_bases,
		size_t const use_bases_idx[], double lsq_matrix[], double lsq_vector[]) noexcept {
	CHECK_ARGS(num_data != 0 && num_data <= static_cast<size_t>(INT32_MAX));
	CHECK_ARGS(data != nullptr && IsAligned(data));
	CHECK_ARGS(mask != nullptr && IsAligned(mask));
	CHECK_ARGS(exclude_indices != nullptr && IsAligned(exclude_indices));
	CHECK_ARGS(num_exclude_indices <= num_data);
	for (size_t i = 0; i < num_exclude_indices; ++i) {
		CHECK_ARGS(exclude_indices[i] < num_data);
	}
	CHECK_ARGS(num_model_bases != 0);
	CHECK_ARGS(num_model_bases <= num_data);
	CHECK_ARGS(basis_data != nullptr && IsAligned(basis_data));
	CHECK_ARGS(num_lsq_bases != 0);
	CHECK_ARGS(num_lsq_bases <= num_model_bases);
	CHECK_ARGS(use_bases_idx != nullptr && IsAligned(use_bases_idx));
	CHECK_ARGS(lsq_matrix != nullptr && IsAligned(lsq_matrix));
	CHECK_ARGS(lsq_vector != nullptr && IsAligned(lsq_vector));
	for (size_t i = 0; i < num_lsq_bases; ++i) {
		CHECK_ARGS(use_bases_idx[i] < num_model_bases);
	}
	BlockScratch &s = BlockScratchGet();
	cudaStream_t const stream = 0;
	sakura_Status st = UploadModel(s, num_data, data, mask, num_model_bases, basis_data,
			num_lsq_bases, use_bases_idx, stream);
	if (st != sakura_Status_kOK) {
		return st;
	}
	CUDA_TRY(s.exclude.Reserve((num_exclude_indices + 1) * sizeof(unsigned long long)));
	CUDA_TRY(cudaMemcpyAsync(s.exclude.ptr, exclude_indices, num_exclude_indices * sizeof(size_t),
			cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(s.G.ptr, lsq_matrix, num_lsq_bases * num_lsq_bases * sizeof(double),
			cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(s.b.ptr, lsq_vector, num_lsq_bases * sizeof(double),
			cudaMemcpyHostToDevice, stream));
	cudaError_t e = LaunchUpdateNormalEquations(static_cast<int>(num_lsq_bases), s.basis.As<double>(),
			static_cast<int>(num_model_bases), s.use_idx.As<int>(), s.mask.As<uint8_t>(),
			s.data.As<float>(), static_cast<int>(num_exclude_indices),
			s.exclude.As<unsigned long long>(), s.G.As<double>(), s.b.As<double>(), stream);
	g_launch_count += 1;
	g_last_kernel = "update_normal_equations";
	if (e != cudaSuccess) {
		SetError("LaunchUpdateNormalEquations", e);
		return sakura_Status_kUnknownError;
	}
	CUDA_TRY(cudaMemcpyAsync(lsq_matrix, s.G.ptr, num_lsq_bases * num_lsq_bases * sizeof(double),
			cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaMemcpyAsync(lsq_vector, s.b.ptr, num_lsq_bases * sizeof(double),
			cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	return sakura_Status_kOK;
}

sakura_Status sakura_SolveSimultaneousEquationsByLUDouble(size_t num_equations,
		double const in_matrix[], double const in_vector[], double out[]) noexcept {
	CHECK_ARGS(in_matrix != nullptr && IsAligned(in_matrix));
	CHECK_ARGS(in_vector != nullptr && IsAligned(in_vector));
	CHECK_ARGS(in_vector != out);
	CHECK_ARGS(out != nullptr && IsAligned(out));
	if (num_equations == 0) {
		return sakura_Status_kOK;
	}
	CHECK_ARGS(num_equations <= 32768);
	BlockScratch &s = BlockScratchGet();
	cudaStream_t const stream = 0;
	size_t const K = num_equations;
	CUDA_TRY(s.G.Reserve(K * K * sizeof(double)));
	CUDA_TRY(s.b.Reserve(K * sizeof(double)));
	CUDA_TRY(s.x.Reserve(K * sizeof(double)));
	CUDA_TRY(s.tr.Reserve(2 * K * sizeof(int)));
	CUDA_TRY(cudaMemcpyAsync(s.G.ptr, in_matrix, K * K * sizeof(double), cudaMemcpyHostToDevice,
			stream));
	CUDA_TRY(cudaMemcpyAsync(s.b.ptr, in_vector, K * sizeof(double), cudaMemcpyHostToDevice, stream));
	cudaError_t e = LaunchSolve(static_cast<int>(K), s.G.As<double>(), s.b.As<double>(),
			s.x.As<double>(), s.tr.As<int>(), stream);
	g_launch_count += 1;
	g_last_kernel = "solve_full_piv_lu";
	if (e != cudaSuccess) {
		SetError("LaunchSolve", e);
		return sakura_Status_kUnknownError;
	}
	CUDA_TRY(cudaMemcpyAsync(out, s.x.ptr, K * sizeof(double), cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	return sakura_Status_kOK;
}

