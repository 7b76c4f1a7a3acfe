// kernels_fast.cu - CMC_MODE_FAST line sweeps (placeholder until the partitioned kernels land)
#include "kernels.h"
namespace cmc {
template <typename FT> bool launch_fast_sweep(int, const SweepArgs<FT> &, cudaStream_t, long long *) { return false; }
template <typename FT> bool launch_pcr_batch(int, int, const FT *, const FT *, const FT *, const FT *, FT *, cudaStream_t) { return false; }
template bool launch_fast_sweep<float>(int, const SweepArgs<float> &, cudaStream_t, long long *);
template bool launch_fast_sweep<double>(int, const SweepArgs<double> &, cudaStream_t, long long *);
template bool launch_pcr_batch<float>(int, int, const float *, const float *, const float *, const float *, float *, cudaStream_t);
template bool launch_pcr_batch<double>(int, int, const double *, const double *, const double *, const double *, double *, cudaStream_t);
}
