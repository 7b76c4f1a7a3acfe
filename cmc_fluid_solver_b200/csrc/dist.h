// dist.h - multi-GPU plumbing of the x-slab decomposition: NCCL halo exchange, residual all-reduce,
// readback gather and the partitioned (SPIKE) x-sweep.  Replaces the reference's peer-copy halo sync
// (src/FluidSolver3D/TimeLayer3D.h:159-247, 272-335), MPI reductions (:630-637) and the pipelined
// distributed Thomas (src/FluidSolver3D/AdiSolver3D.cu:524-640).  NCCL is loaded with dlopen so that a
// single-GPU process has no NCCL dependency and a torch process shares torch's own libnccl.
#pragma once
#include "common.cuh"

namespace cmc {

struct DistContext;

const char *dist_error();
int dist_unique_id(void *id128);
DistContext *dist_create(int device, int rank, int nranks, const void *nccl_id, const Layout &L, int fp_bytes, cudaStream_t s);
void dist_destroy(DistContext *d);

// exchange the boundary x-planes of the 4 fields of one layer with the neighbour slabs (into the guard planes)
template <typename FT>
int dist_halo_exchange(DistContext *d, const Layout &L, FT *const fields[4], cudaStream_t s, long long *launches);
// partitioned x-sweep (local elimination -> reduced system exchange -> back substitution) + merged temp_out
template <typename FT>
int dist_sweep_x(DistContext *d, const SweepArgs<FT> &A, cudaStream_t s, long long *launches);
int dist_allreduce_f64(DistContext *d, double *dev_buf, int n, cudaStream_t s);
int dist_sum_i64(DistContext *d, long long *host_vals, int n, cudaStream_t s);
template <typename FT>
int dist_gather_layer(DistContext *d, const Layout &G, int ox, int oy, int oz, const FT *d_vel, const double *d_T,
                      int oi0, int oi1, FT *h_vel, double *h_T, cudaStream_t s);

} // namespace cmc
