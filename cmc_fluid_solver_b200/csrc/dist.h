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
// every rank contributes `bytes` from send_dev; recv_dev receives nranks * bytes in rank order (device buffers)
int nccl_allgather_bytes(NcclComm *c, const void *send_dev, void *recv_dev, size_t bytes, cudaStream_t s);

// ---- peer memory over NVLink / NVSwitch (one process per GPU) ------------------------------------------------------
// The exchange steps of a slab-decomposed run are stores into the other ranks' buffers issued by the sweep kernels
// themselves (SweepArgs::push_*, xcoef_to; k_x_interface): every rank maps every other rank's exchange arena with
// CUDA IPC.  What remains between two kernels is ordering: a monotone epoch per rank, published to the peers' flag
// words after a kernel (peer_signal) and awaited before the next one (peer_wait).  Both are stream-ordered one-warp
// kernels; nothing synchronises with the host.
struct PeerMap {
	int rank = 0, nranks = 1;
	void *base[16] = {};            // base[r]: rank r's arena as mapped into this process (base[rank] = the local arena)
	bool mapped[16] = {};
};
// maps the arenas of all ranks (collective over `c`); returns 0 and fills `pm`, or -1 when any rank could not
// (all ranks then agree to stay on the NCCL transport)
int peer_map_arenas(NcclComm *c, void *arena, size_t arena_bytes, int rank, int nranks, PeerMap *pm, cudaStream_t s);
void peer_unmap(PeerMap *pm);
// flags: `nranks` words in every arena at byte offset flag_off; word [src] of rank dst is written by rank src
void peer_signal(const PeerMap &pm, size_t flag_off, unsigned epoch, cudaStream_t s);
// waits until the local words of all ranks in `src_mask` are >= epoch; *timeout_flag (device) is set on a 10 s timeout
void peer_wait(const PeerMap &pm, size_t flag_off, unsigned src_mask, unsigned epoch, int *timeout_flag, cudaStream_t s);

} // namespace cmc
