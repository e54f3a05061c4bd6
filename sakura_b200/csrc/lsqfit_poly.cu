// placeholder until the register-resident kernel lands (next commit)
#include "lsqfit_poly.cuh"
namespace sakura_b200 {
bool PolyFastSupported(size_t, size_t, size_t, size_t, void const *, void const *, void const *,
		void const *, void const *) {
	return false;
}
cudaError_t LaunchPolyFit(PolyFitParams const &, int, cudaStream_t, int *) {
	return cudaErrorNotSupported;
}
}
