// dist.cu - see dist.h.  (multi-GPU path; single-GPU handles never enter this file)
#include <string>
#include "dist.h"

namespace cmc {

static thread_local std::string g_dist_err;
const char *dist_error() { return g_dist_err.c_str(); }

struct DistContext { int rank, nranks; };

int dist_unique_id(void *) { g_dist_err = "multi-GPU support not built yet"; return -1; }
DistContext *dist_create(int, int, int, const void *, const Layout &, int, cudaStream_t) { g_dist_err = "multi-GPU support not built yet"; return nullptr; }
void dist_destroy(DistContext *d) { delete d; }
template <typename FT> int dist_halo_exchange(DistContext *, const Layout &, FT *const[4], cudaStream_t, long long *) { return -1; }
template <typename FT> int dist_sweep_x(DistContext *, const SweepArgs<FT> &, cudaStream_t, long long *) { return -1; }
int dist_allreduce_f64(DistContext *, double *, int, cudaStream_t) { return -1; }
int dist_sum_i64(DistContext *, long long *, int, cudaStream_t) { return -1; }
template <typename FT> int dist_gather_layer(DistContext *, const Layout &, int, int, int, const FT *, const double *, int, int, FT *, double *, cudaStream_t) { return -1; }

#define INST(FT) \
	template int dist_halo_exchange<FT>(DistContext *, const Layout &, FT *const[4], cudaStream_t, long long *); \
	template int dist_sweep_x<FT>(DistContext *, const SweepArgs<FT> &, cudaStream_t, long long *); \
	template int dist_gather_layer<FT>(DistContext *, const Layout &, int, int, int, const FT *, const double *, int, int, FT *, double *, cudaStream_t);
INST(float)
INST(double)
} // namespace cmc
