// dist.cu - see dist.h
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "dist.h"

namespace cmc {

static thread_local std::string g_err;
const char *nccl_error() { return g_err.c_str(); }

// the handful of NCCL entry points we use (stable C ABI; declared here so that no NCCL header is needed)
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void *ncclComm_t;
typedef int ncclResult_t;
enum { kNcclUint8 = 1, kNcclFloat64 = 8, kNcclSum = 0 };
struct Api {
	void *lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_t, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static Api *api()
{
	static Api a;
	static bool tried = false;
	if (tried) return a.lib ? &a : nullptr;
	tried = true;
	const char *names[] = {getenv("CMC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
	for (const char *n : names) {
		if (!n || !*n) continue;
		a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if (a.lib) break;
	}
	if (!a.lib) { g_err = std::string("cannot load NCCL (set CMC_NCCL_LIB): ") + (dlerror() ? dlerror() : ""); return nullptr; }
#define LOAD(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, sym)); if (!a.field) { g_err = std::string("NCCL symbol missing: ") + sym; a.lib = nullptr; return nullptr; }
	LOAD(GetUniqueId, "ncclGetUniqueId") LOAD(CommInitRank, "ncclCommInitRank") LOAD(CommDestroy, "ncclCommDestroy")
	LOAD(Send, "ncclSend") LOAD(Recv, "ncclRecv") LOAD(AllReduce, "ncclAllReduce")
	LOAD(GroupStart, "ncclGroupStart") LOAD(GroupEnd, "ncclGroupEnd") LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
	return &a;
}

#define NCCL_TRY(call) do { ncclResult_t r__ = (call); if (r__ != 0) { g_err = std::string(#call) + ": " + a->GetErrorString(r__); return -1; } } while (0)

struct NcclComm { ncclComm_t comm; int rank, nranks; };

int nccl_unique_id(void *id128)
{
	Api *a = api();
	if (!a) return -1;
	NCCL_TRY(a->GetUniqueId(reinterpret_cast<ncclUniqueId_t *>(id128)));
	return 0;
}

NcclComm *nccl_create(int rank, int nranks, const void *id128)
{
	Api *a = api();
	if (!a) return nullptr;
	ncclUniqueId_t id;
	memcpy(&id, id128, sizeof id);
	ncclComm_t c = nullptr;
	ncclResult_t r = a->CommInitRank(&c, nranks, id, rank);
	if (r != 0) { g_err = std::string("ncclCommInitRank: ") + a->GetErrorString(r); return nullptr; }
	return new NcclComm{c, rank, nranks};
}

void nccl_destroy(NcclComm *c)
{
	if (!c) return;
	Api *a = api();
	if (a && c->comm) a->CommDestroy(c->comm);
	delete c;
}

int nccl_exchange(NcclComm *c, const P2P *ops, int nops, cudaStream_t s)
{
	Api *a = api();
	if (!a || !c) { g_err = "NCCL not initialised"; return -1; }
	NCCL_TRY(a->GroupStart());
	for (int i = 0; i < nops; i++) {
		if (ops[i].send) NCCL_TRY(a->Send(ops[i].send, ops[i].bytes, kNcclUint8, ops[i].peer, c->comm, s));
		if (ops[i].recv) NCCL_TRY(a->Recv(ops[i].recv, ops[i].bytes, kNcclUint8, ops[i].peer, c->comm, s));
	}
	NCCL_TRY(a->GroupEnd());
	return 0;
}

int nccl_allreduce_sum_f64(NcclComm *c, double *dev_buf, int n, cudaStream_t s)
{
	Api *a = api();
	if (!a || !c) { g_err = "NCCL not initialised"; return -1; }
	NCCL_TRY(a->AllReduce(dev_buf, dev_buf, (size_t)n, kNcclFloat64, kNcclSum, c->comm, s));
	return 0;
}

int nccl_allgather_bytes(NcclComm *c, const void *send_dev, void *recv_dev, size_t bytes, cudaStream_t s)
{
	Api *a = api();
	if (!a || !c) { g_err = "NCCL not initialised"; return -1; }
	if (cudaMemcpyAsync((char *)recv_dev + (size_t)c->rank * bytes, send_dev, bytes, cudaMemcpyDeviceToDevice, s) != cudaSuccess) { g_err = "allgather: local copy failed"; return -1; }
	NCCL_TRY(a->GroupStart());
	for (int p = 0; p < c->nranks; p++) {
		if (p == c->rank) continue;
		NCCL_TRY(a->Send(send_dev, bytes, kNcclUint8, p, c->comm, s));
		NCCL_TRY(a->Recv((char *)recv_dev + (size_t)p * bytes, bytes, kNcclUint8, p, c->comm, s));
	}
	NCCL_TRY(a->GroupEnd());
	return 0;
}

// ---- peer memory ---------------------------------------------------------------------------------------------------
int peer_map_arenas(NcclComm *c, void *arena, size_t arena_bytes, int rank, int nranks, PeerMap *pm, cudaStream_t s)
{
	(void)arena_bytes;
	pm->rank = rank; pm->nranks = nranks;
	for (int r = 0; r < 16; r++) { pm->base[r] = nullptr; pm->mapped[r] = false; }
	pm->base[rank] = arena;
	struct Msg { cudaIpcMemHandle_t h; int ok; int pad[15]; };
	static_assert(sizeof(Msg) == 128, "message layout");
	Msg mine;
	memset(&mine, 0, sizeof mine);
	mine.ok = cudaIpcGetMemHandle(&mine.h, arena) == cudaSuccess;
	if (!mine.ok) cudaGetLastError();
	Msg *d_all = nullptr, *d_mine = nullptr;
	std::vector<Msg> all((size_t)nranks);
	bool good = cudaMalloc((void **)&d_all, sizeof(Msg) * nranks) == cudaSuccess && cudaMalloc((void **)&d_mine, sizeof(Msg)) == cudaSuccess;
	if (good) good = cudaMemcpyAsync(d_mine, &mine, sizeof mine, cudaMemcpyHostToDevice, s) == cudaSuccess;
	// the collective must run on every rank even if this one already failed locally
	if (nccl_allgather_bytes(c, d_mine, d_all, sizeof(Msg), s)) good = false;
	if (good) good = cudaMemcpyAsync(all.data(), d_all, sizeof(Msg) * nranks, cudaMemcpyDeviceToHost, s) == cudaSuccess;
	if (cudaStreamSynchronize(s) != cudaSuccess) good = false;
	if (good) {
		for (int r = 0; r < nranks && good; r++) {
			if (!all[r].ok) { good = false; break; }
			if (r == rank) continue;
			void *p = nullptr;
			if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = false; break; }
			pm->base[r] = p; pm->mapped[r] = true;
		}
	}
	// agree: everybody maps everybody, or nobody uses peer memory
	double *d_ok = reinterpret_cast<double *>(d_all);
	double v = good ? 1.0 : 0.0;
	bool agreed = false;
	if (d_all && cudaMemcpyAsync(d_ok, &v, sizeof v, cudaMemcpyHostToDevice, s) == cudaSuccess && nccl_allreduce_sum_f64(c, d_ok, 1, s) == 0 &&
	    cudaMemcpyAsync(&v, d_ok, sizeof v, cudaMemcpyDeviceToHost, s) == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess)
		agreed = v > nranks - 0.5;
	if (d_all) cudaFree(d_all);
	if (d_mine) cudaFree(d_mine);
	if (!agreed) { peer_unmap(pm); g_err = "peer mapping of the exchange arenas failed on at least one rank"; return -1; }
	return 0;
}

void peer_unmap(PeerMap *pm)
{
	for (int r = 0; r < 16; r++)
		if (pm->mapped[r] && pm->base[r]) { cudaIpcCloseMemHandle(pm->base[r]); pm->base[r] = nullptr; pm->mapped[r] = false; }
}

struct FlagTargets { unsigned *p[16]; };

__global__ void k_peer_signal(const FlagTargets ft, int nranks, int me, unsigned epoch)
{
	// everything this rank stored into peer memory before this kernel (stream order) becomes visible system-wide
	// before the flag does
	__threadfence_system();
	const int r = threadIdx.x;
	if (r < nranks && r != me && ft.p[r]) {
		volatile unsigned *f = ft.p[r] + me;
		*f = epoch;
	}
	__threadfence_system();
}

__global__ void k_peer_wait(const volatile unsigned *flags, int nranks, unsigned src_mask, unsigned epoch, int *timeout_flag)
{
	const int r = threadIdx.x;
	if (r < nranks && (src_mask >> r & 1u)) {
		unsigned long long t0;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
		for (;;) {
			unsigned v;
			asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
			if ((int)(v - epoch) >= 0) break;
			__nanosleep(200);
			unsigned long long t1;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
			if (t1 - t0 > 10000000000ull) { if (timeout_flag) atomicExch(timeout_flag, 1); break; }    // 10 s: a peer died
		}
	}
	__threadfence_system();
}

void peer_signal(const PeerMap &pm, size_t flag_off, unsigned epoch, cudaStream_t s)
{
	FlagTargets ft;
	for (int r = 0; r < 16; r++) ft.p[r] = (r < pm.nranks && pm.base[r]) ? reinterpret_cast<unsigned *>((char *)pm.base[r] + flag_off) : nullptr;
	k_peer_signal<<<1, 32, 0, s>>>(ft, pm.nranks, pm.rank, epoch);
}

void peer_wait(const PeerMap &pm, size_t flag_off, unsigned src_mask, unsigned epoch, int *timeout_flag, cudaStream_t s)
{
	src_mask &= ~(1u << pm.rank);
	if (!src_mask) return;
	const volatile unsigned *flags = reinterpret_cast<const volatile unsigned *>((char *)pm.base[pm.rank] + flag_off);
	k_peer_wait<<<1, 32, 0, s>>>(flags, pm.nranks, src_mask, epoch, timeout_flag);
}

} // namespace cmc
