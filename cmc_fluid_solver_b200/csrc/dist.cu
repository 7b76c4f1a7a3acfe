// dist.cu - see dist.h
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
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

} // namespace cmc
