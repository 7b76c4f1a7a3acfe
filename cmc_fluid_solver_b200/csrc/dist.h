// dist.h - NCCL transport of the x-slab decomposition (one process per GPU).  Replaces the reference's
// peer-copy halo sync (src/FluidSolver3D/TimeLayer3D.h:159-247, 272-335), its MPI reductions (:630-637) and
// the hand-offs of the pipelined distributed Thomas (src/FluidSolver3D/AdiSolver3D.cu:524-640).
// NCCL is loaded with dlopen: a single-GPU process has no NCCL dependency, and a torch process shares
// torch's own libnccl.so.2.  All operations are stream-ordered; nothing here synchronises with the host.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace cmc {

struct NcclComm;   // opaque

const char *nccl_error();
int nccl_unique_id(void *id128);
NcclComm *nccl_create(int rank, int nranks, const void *id128);
void nccl_destroy(NcclComm *c);

// grouped point-to-point: every (buffer, peer) pair is one ncclSend / ncclRecv inside one group
struct P2P { const void *send; void *recv; size_t bytes; int peer; };   // send or recv may be null
int nccl_exchange(NcclComm *c, const P2P *ops, int nops, cudaStream_t s);
int nccl_allreduce_sum_f64(NcclComm *c, double *dev_buf, int n, cudaStream_t s);

} // namespace cmc
