// kernels_tma.cu - CMC_MODE_FAST sweeps along the strided axes (x, y) as persistent CTAs fed by the Tensor Memory
// Accelerator (cp.async.bulk.tensor / UTMALDG): BASELINE.json north_star's "TMA-staged tiles" for the strided-axis sweeps.
//
// Same arithmetic as k_fast_sweep (kernels_fast.cu: partition method with 8-row chunks in registers, CR + PCR reduced
// solve in shared memory, coefficient build / boundary rows / mask / relaxation fused), different data movement.  The
// direct-load kernel is latency bound (ncu, profiles/r01_ncu_sweeps_final.md): one 512-thread CTA per SM owns the
// register file, ~150 LDG and ~64 STG per thread per tile go through the LSU (lg_throttle 2.0-2.3 stall cycles per
// issue), its load phases are separated by barriers and HBM idles while the SM solves (DRAM 42 % busy).  Here
//   * persistent CTAs walk over their tiles (tile = 8 or 16 neighbouring lines x all rows of the line; one 512-thread CTA per
//     SM at 512 rows; tile shapes, the CTA-pair form and the one-pass slab coupling: see the kernel's own comment);
//   * every input field of a tile arrives by ONE bulk-tensor copy per field (a 5-D box that gathers the 64-byte row
//     segments of the tile - rows 256 KB (x) / 4 KB (y) apart - straight into a shared-memory slot laid out
//     [row-in-chunk][chunk][line], which the compute threads read without bank conflicts); thread 0 issues them, an
//     mbarrier per slot counts the bytes in;
//   * five slots (5 x 32 KB in fp64), eleven copies per tile, each issued one solve phase (or more) before its use:
//         slot 0: temp[DIR]            (resident from the u,v,w phase to the dissipation function)
//         slot 1: temp.T   -> the first other temp component
//         slot 2: cur.u    -> the second other temp component
//         slot 3: cur.v    -> temp[DIR] of the cross-line neighbour below (j-1 for x lines, i-1 for y lines) -> cur.T
//         slot 4: cur.w    -> temp[DIR] of the cross-line neighbour above -> temp.T (for the relaxation of T at the end)
//     so HBM streams while the SM eliminates / solves the reduced systems, and the LSU only sees shared-memory loads and
//     the result stores;
//   * no global load in the common path except the k +- 1 neighbours of the two edge lines of a tile.
// Requirements (launch_tma_sweep returns false otherwise and the caller uses k_fast_sweep): x or y sweep of a slab whose
// lines stay inside the slab (MODE 0), line length a multiple of 8 and at most 512 - or a multiple of 16 and at most 1024,
// which a pair of CTAs handles (CL 2 below; the direct-load kernel stops at 512 rows).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <algorithm>
#include <cuda.h>
#include "kernels.h"
#include "fast_core.cuh"
#include "ring_slots.cuh"

namespace cmc {

// ---- tensor maps ------------------------------------------------------------------------------------------------------
struct TmaMaps {
	CUtensorMap temp[4];     // linearisation layer (read)
	CUtensorMap cur[4];      // the sweep's "cur" layer
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
	static EncodeTiledFn fn = nullptr;
	static bool tried = false;
	if (!tried) {
		tried = true;
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
			fn = reinterpret_cast<EncodeTiledFn>(p);
		else
			cudaGetLastError();
	}
	return fn;
}

// One map per (field buffer, direction, layout): the layers rotate through a handful of physical buffers, so the maps
// are built once and cached.
struct MapKey {
	const void *ptr; long long bstride; int dir, fp, gp, nl, nx, ny, nz, jbs;
	bool operator<(const MapKey &o) const { return memcmp(this, &o, sizeof(MapKey)) < 0; }
};

template <typename FT>
static bool tensor_map_for(const FT *field, const Layout &L, int dir, int GP, int NL, CUtensorMap *out)
{
	static std::map<MapKey, CUtensorMap> cache;
	static std::mutex mu;
	MapKey key;
	memset(&key, 0, sizeof key);
	key.ptr = field; key.bstride = L.bstride; key.dir = dir; key.fp = (int)sizeof(FT); key.gp = GP; key.nl = NL; key.nx = L.nx; key.ny = L.ny; key.nz = L.nz; key.jbs = L.jbs;
	std::lock_guard<std::mutex> lock(mu);
	auto it = cache.find(key);
	if (it != cache.end()) { *out = it->second; return true; }
	EncodeTiledFn enc = encode_fn();
	if (!enc) return false;
	const cuuint64_t es = sizeof(FT);
	const int rows_per_block = L.nblk > 1 ? (1 << L.jbs) : L.ny;
	cuuint64_t dims[5], strides[4];
	cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
	void *base;
	if (dir == 0) {
		// (k, chunk along x, row in chunk, j inside its y-block, y-block); plane 0 is the first real plane
		dims[0] = (cuuint64_t)L.nzp; dims[1] = (cuuint64_t)(L.nx / M); dims[2] = M; dims[3] = (cuuint64_t)rows_per_block; dims[4] = (cuuint64_t)L.nblk;
		strides[0] = (cuuint64_t)(M * L.plane) * es; strides[1] = (cuuint64_t)L.plane * es; strides[2] = (cuuint64_t)L.nzp * es; strides[3] = (cuuint64_t)L.bstride * es;
		box[0] = (cuuint32_t)NL; box[1] = (cuuint32_t)GP; box[2] = M; box[3] = 1; box[4] = 1;
		base = (void *)(field + L.plane);
	} else {
		// (k, chunk inside a y-block, y-block, row in chunk, x-plane incl. the guard planes)
		const int gb = rows_per_block / M;
		dims[0] = (cuuint64_t)L.nzp; dims[1] = (cuuint64_t)gb; dims[2] = (cuuint64_t)L.nblk; dims[3] = M; dims[4] = (cuuint64_t)(L.nx + 2);
		strides[0] = (cuuint64_t)(M * L.nzp) * es; strides[1] = (cuuint64_t)L.bstride * es; strides[2] = (cuuint64_t)L.nzp * es; strides[3] = (cuuint64_t)L.plane * es;
		const int bg = L.nblk > 1 ? gb : GP;            // chunks of one y-block in the box; GP / bg blocks
		box[0] = (cuuint32_t)NL; box[1] = (cuuint32_t)bg; box[2] = (cuuint32_t)(GP / bg); box[3] = M; box[4] = 1;
		base = (void *)field;
	}
	const CUtensorMapDataType dt = sizeof(FT) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
	CUtensorMap m;
	const CUresult r = enc(&m, dt, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
	                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		static bool warned = false;
		if (!warned) { warned = true; fprintf(stderr, "[cmc] cuTensorMapEncodeTiled failed (%d) for dir %d: using the direct-load sweep kernel\n", (int)r, dir); }
		return false;
	}
	cache[key] = m;
	*out = m;
	return true;
}

// ---- device helpers -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
	// try_wait suspends the thread for up to the given time (ns) per attempt; a copy that never completes (a tensor map
	// that does not describe the buffer) must not hang the device: trap after about two seconds
	long long t0 = 0;
	for (int tries = 0;; tries++) {
		unsigned ok;
		asm volatile(
			"{\n\t"
			".reg .pred P1;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x100000;\n\t"
			"selp.u32 %0, 1, 0, P1;\n\t"
			"}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
		if (ok) return;
		if ((tries & 63) == 63) {
			const long long now = clock64();
			if (!t0) t0 = now;
			else if (now - t0 > 4000000000ll) __trap();       // ~2 s at 1.9 GHz
		}
	}
}
// Orders this thread's earlier shared-memory reads (generic proxy) before bulk copies (async proxy) that are issued after
// the next barrier and overwrite the same slot.  __syncthreads() alone orders the threads among themselves, not against
// the copy engine: a shared-memory load that is still queued in the load/store unit when the barrier resolves can be
// overtaken by the incoming copy (seen as one corrupted tile in ~10^5 with two CTAs per SM; PTX memory model, proxies).
__device__ __forceinline__ void slot_reads_done() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *map, unsigned long long *bar, int c0, int c1, int c2, int c3, int c4)
{
	asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
	             ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)),
	             "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

__device__ __forceinline__ void tma_load_5d_hint(void *dst, const CUtensorMap *map, unsigned long long *bar, int c0, int c1, int c2, int c3, int c4,
                                                 unsigned long long policy)
{
	asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
	             ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)),
	             "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "l"(policy) : "memory");
}
__device__ __forceinline__ unsigned long long l2_evict_first()
{
	unsigned long long p;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ unsigned long long l2_evict_last()
{
	unsigned long long p;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}

// result stores: written once, read again only by the next sweep (17 GB later at 512^3): streaming (L2 evict-first), so
// that they do not push the temp tiles that neighbouring tiles are about to re-read out of L2
template <typename FT>
__device__ __forceinline__ void store8_stream(FT *__restrict__ p, const int (&off)[M], unsigned mask, const FT (&v)[M], int streaming)
{
	if (streaming) {
#pragma unroll
		for (int i = 0; i < M; i++)
			if (mask & (1u << i)) __stcs(p + off[i], v[i]);
	} else {
#pragma unroll
		for (int i = 0; i < M; i++)
			if (mask & (1u << i)) p[off[i]] = v[i];
	}
}

// ---- cluster helpers (CL == 2: two CTAs on neighbouring SMs share a tile, each takes half of every line) --------------------
__device__ __forceinline__ unsigned cluster_ctarank()
{
	unsigned r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in the CTA `rank` of the cluster
__device__ __forceinline__ unsigned peer_smem(const void *p, unsigned rank)
{
	unsigned r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
	return r;
}
__device__ __forceinline__ void st_cluster(unsigned addr, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void st_cluster(unsigned addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void st_cluster(unsigned addr, int v) { asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void mbar_arrive_peer(unsigned bar_addr)
{
	asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long *bar, unsigned parity)
{
	long long t0 = 0;
	for (int tries = 0;; tries++) {
		unsigned ok;
		asm volatile(
			"{\n\t"
			".reg .pred P1;\n\t"
			"mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2, 0x100000;\n\t"
			"selp.u32 %0, 1, 0, P1;\n\t"
			"}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
		if (ok) return;
		if ((tries & 63) == 63) {
			const long long now = clock64();
			if (!t0) t0 = now;
			else if (now - t0 > 4000000000ll) __trap();
		}
	}
}

// ---- helpers of the fused slab-coupled x-sweep (XS): coefficient tables in the other slabs' memory ----------------------------
// A value crosses to another GPU as self-validating 8-byte words - 32 bits of payload + the sweep's 32-bit epoch, one aligned
// store each (the low-latency protocol of collective libraries): the reader polls the word until it carries the epoch, no
// flag word, no fence on either side.  (A release store at system scope has to wait until every earlier store of the thread -
// the sweep's result stores - is acknowledged: measured 20 us per exchange under load, against ~3 us this way.)
template <typename FT> struct LLWords;
template <> struct LLWords<double> { static constexpr int W = 2; };
template <> struct LLWords<float> { static constexpr int W = 1; };
__device__ __forceinline__ void ll_store(unsigned long long *p, size_t wstride, double v, unsigned epoch)
{
	const unsigned long long b = (unsigned long long)__double_as_longlong(v), ep = (unsigned long long)epoch << 32;
	*reinterpret_cast<volatile unsigned long long *>(p) = (b & 0xffffffffull) | ep;
	*reinterpret_cast<volatile unsigned long long *>(p + wstride) = (b >> 32) | ep;
}
__device__ __forceinline__ void ll_store(unsigned long long *p, size_t, float v, unsigned epoch)
{
	*reinterpret_cast<volatile unsigned long long *>(p) = (unsigned long long)__float_as_uint(v) | ((unsigned long long)epoch << 32);
}
// word w of a value with the epoch on top
__device__ __forceinline__ unsigned long long ll_word(double v, int w, unsigned epoch)
{
	const unsigned long long b = (unsigned long long)__double_as_longlong(v);
	return (w ? b >> 32 : b & 0xffffffffull) | ((unsigned long long)epoch << 32);
}
__device__ __forceinline__ unsigned long long ll_word(float v, int, unsigned epoch) { return (unsigned long long)__float_as_uint(v) | ((unsigned long long)epoch << 32); }
// poll one word until it carries `epoch`; a peer that never arrives must not hang the device: trap after about two seconds
__device__ __forceinline__ unsigned ll_poll(const unsigned long long *p, unsigned epoch)
{
	long long t0 = 0;
	for (int tries = 0;; tries++) {
		const unsigned long long w = *reinterpret_cast<const volatile unsigned long long *>(p);
		if ((unsigned)(w >> 32) == epoch) return (unsigned)w;
		if ((tries & 255) == 255) {
			const long long now = clock64();
			if (!t0) t0 = now;
			else if (now - t0 > 4000000000ll) __trap();
		}
	}
}
// a value from its polled payload words (shared memory)
template <typename FT> __device__ __forceinline__ FT ll_value(const unsigned *p, int wstride);
template <> __device__ __forceinline__ double ll_value<double>(const unsigned *p, int wstride)
{
	return __longlong_as_double((long long)(((unsigned long long)p[wstride] << 32) | p[0]));
}
template <> __device__ __forceinline__ float ll_value<float>(const unsigned *p, int) { return __uint_as_float(p[0]); }

// Interface system of one line across P slabs (same algebra as k_x_interface, kernels_fast.cu): unknowns per slab r the first
// row F_r and the last row L_r,   F_r + pf_r L_{r-1} + qf_r F_{r+1} = f_r ,   L_r + pl_r L_{r-1} + ql_r F_{r+1} = l_r ,
// block-tridiagonal in Z_r = (L_r, F_{r+1}), block Thomas.  Every slab solves it for itself (all slabs' coefficients are in
// its table) and keeps the two values it needs: xl = L_{me-1}, xr = F_{me+1}.  C(r, v) = coefficient v of slab r for this line.
template <typename FT, int NR, typename Coef>
__device__ __noinline__ void xs_interface(Coef C, int P, int me, int vf, int vl, int vp, FT (&xl)[NR], FT (&xr)[NR])
{
	constexpr int MP = MAX_SLABS;
	FT m01[MP], i00[MP], i01[MP], i10[MP], i11[MP], r0[NR][MP], r1[NR][MP];
	FT p_i01 = 0, p_s[NR];
#pragma unroll
	for (int q = 0; q < NR; q++) { p_s[q] = 0; xl[q] = 0; xr[q] = 0; }
	for (int r = 0; r + 1 < P; r++) {
		const FT pl = C(r, vp + 2), ql = C(r, vp + 3), pf1 = C(r + 1, vp + 0), qf_r = C(r, vp + 1);
		m01[r] = ql - pl * p_i01 * qf_r;             // (r == 0: pl == 0)
		const FT id = rcp<FT>(FT(1) - m01[r] * pf1);
		i00[r] = id; i01[r] = -m01[r] * id; i10[r] = -pf1 * id; i11[r] = id;
#pragma unroll
		for (int q = 0; q < NR; q++) {
			r0[q][r] = C(r, vl + q) - pl * p_s[q];
			r1[q][r] = C(r + 1, vf + q);
			p_s[q] = i00[r] * r0[q][r] + i01[r] * r1[q][r];
		}
		p_i01 = i01[r];
	}
#pragma unroll
	for (int q = 0; q < NR; q++) {
		FT nextF = FT(0);                            // F_{r+2} of the block above
		for (int r = P - 2; r >= 0; r--) {
			const FT qf1 = C(r + 1, vp + 1);
			const FT b0 = r0[q][r], b1 = r1[q][r] - qf1 * nextF;
			const FT Lr = i00[r] * b0 + i01[r] * b1, Fr1 = i10[r] * b0 + i11[r] * b1;
			if (r == me - 1) xl[q] = Lr;
			if (r == me) xr[q] = Fr1;
			nextF = Fr1;
		}
	}
}

// ---- the kernel -----------------------------------------------------------------------------------------------------------
// GP chunks x NL lines per CTA (GP * NL threads).  CL = CTAs per tile:
//   CL 1: the CTA holds whole lines (up to GP * 8 rows);
//   CL 2: a cluster of two CTAs holds a tile of NL lines, each CTA one half of every line.  What that buys is the tile
//         shape: 16 lines x 256 rows fit the same 32 KB slots as 8 lines x 512 rows, and a 16-line fp64 tile has 128-byte
//         rows, which the copy engine delivers twice as fast as 64-byte rows (tools/tma_probe.cu).  The two halves of a
//         line are coupled like two slabs of a decomposed grid (kernels_fast.cu MODE 1 / 2), except that nothing is
//         read twice: each CTA eliminates its half, solves its reduced system with one extra right-hand side (the
//         response to the unknown row of the other half), the CTAs swap 4 numbers per line and phase through distributed
//         shared memory, and every thread finishes its rows from registers.
//   XS 1: the slab-coupled x-sweep of a decomposed grid in ONE pass (replaces spike pass + interface kernel + coupled pass,
//         kernels_fast.cu MODE 1 / 2): the same open-ended elimination with two spike columns; the first / last row's
//         coefficients of every line go straight into every other slab's table (peer memory over NVLink, 128-byte stores)
//         followed by a release flag per line; the tile's first NL threads wait for the flags of all slabs, solve the
//         interface system of their line, and every thread finishes its rows from registers.  No slab waits before it has
//         sent, tiles are handed out in index order on every GPU, so the smallest unfinished tile can always complete.
template <typename FT, int DIR, int GP, int NL, int CL, int XS>
__global__ void __launch_bounds__(GP * NL, GP * NL >= 512 ? 1 : 512 / (GP * NL))
k_tma_sweep(const SweepArgs<FT> A, const FastConst<FT> K, const __grid_constant__ TmaMaps TM, const int ntiles, const int hints)
{
	static_assert(DIR == 0 || DIR == 1, "strided axes only (z lines are contiguous: kernels_fast.cu)");
	static_assert(NL == 8 || NL == 16, "8 or 16 lines per tile");
	static_assert(CL == 1 || CL == 2, "one CTA or a CTA pair per tile");
	static_assert(XS == 0 || (DIR == 0 && CL == 1), "slabs couple along x; one CTA per tile");
	constexpr int STR = GP * NL;
	constexpr int GS = NL;                          // shared-memory distance of neighbouring chunks of a line
	constexpr int SLOT = STR * M;                   // elements per slot: one field of one tile (part)
	constexpr int NW = STR / 32;                    // warps
	constexpr int NS = 5;                           // slots
	constexpr int XV = CL == 2 ? 1 : XS ? 2 : 0;    // extra right-hand sides of the reduced solves (spike columns)
	constexpr unsigned SLOT_BYTES = SLOT * (unsigned)sizeof(FT);
	// the two temp components other than the one along the sweep, and their slots
	constexpr int QO1 = DIR == 0 ? 1 : 0, QO2 = 2;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	FT *slots = reinterpret_cast<FT *>(smem_raw);                                   // NS slots
	FT *sys = slots + NS * SLOT;                                                    // reduced-solve scratch
	FT *sol = sys;                                                                  // aliases the CR publications (see reduced_solve)
	FT *headx = sys + reduced_scratch_elems<3 + XV, GP, NL>();                      // heads that cross a warp: 5 x (NW * NL)
	FT *edge = headx + 5 * NW * NL;                                                 // CL 2: the row next to this half, 4 fields x NL
	FT *xown = edge + (CL == 2 ? 4 * NL : XS ? 8 * NL : 0);                         // CL 2: this half's interface coefficients [6][NL]; XS: xl / xr [8][NL]
	FT *xin = xown + (CL == 2 ? 6 * NL : XS ? 8 * NL : 0);                          // CL 2: the other half's, written by the peer CTA; XS: this slab's own 16 coefficients
	uint8_t *roles = reinterpret_cast<uint8_t *>(xin + (CL == 2 ? 6 * NL : XS ? 16 * NL : 0));     // descriptor bytes of the tile: NL * GP * 8
	unsigned long long *full = reinterpret_cast<unsigned long long *>(roles + NL * GP * 8);   // NS mbarriers (+ 2 for the exchange)
	unsigned long long *xbar = full + NS;                                                 // CL 2: [0] u,v,w phase, [1] T phase
	int *next_box = reinterpret_cast<int *>(full + NS + 2);                               // the tile after the current one
	unsigned *xw = reinterpret_cast<unsigned *>(next_box + 4);                            // XS: the other slabs' words of this tile, [slab][value][word][line]
#define SLOTP(k) (slots + (k) * SLOT)

	const Layout &L = A.L;
	const int t = threadIdx.x;
	const int l = t % NL, g = t / NL;               // line in tile, chunk (of this CTA's part of the line)
	const int e = t;                                // == g * NL + l
	const int lane = t & 31, warp = t >> 5;
	const unsigned crank = CL == 2 ? cluster_ctarank() : 0u;
	const int n = DIR == 0 ? L.nx : L.ny;           // rows of a line
	const int nloc = n / CL;                        // rows of this CTA's part
	const int row0 = (int)crank * nloc;             // first row of this CTA's part
	const int GL = nloc / M;                        // chunks that hold real rows (nloc % 8 == 0)
	const int r0 = row0 + g * M;                    // first row of this chunk in the line
	const int ktiles = (L.nz + NL - 1) / NL;
	const int stride = DIR == 0 ? (int)L.plane : (int)L.nzp;        // between the rows of a chunk
	const int cid = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // tile-processing unit: CTA or CTA pair
	const int nunits = (int)gridDim.x / CL;
	// open ends of this part (CL 2): the other half continues the line there
	const bool open_lo = XS ? A.xs_me > 0 : (CL == 2 && crank == 1), open_hi = XS ? A.xs_me + 1 < A.xs_P : (CL == 2 && crank == 0);

	if (t == 0) {
#pragma unroll
		for (int s = 0; s < NS; s++) mbar_init(full + s, 1);
		if (CL == 2) { mbar_init(xbar + 0, NL + (crank == 1 ? 1 : 0)); mbar_init(xbar + 1, NL); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();
	if (CL == 2) cluster_sync_all();                // the peer's barriers exist before anything arrives on them
	unsigned ph = 0;                                // phase parity of every slot's mbarrier (bits NS, NS + 1: the exchange barriers)
#define WAIT_SLOT(s) do { mbar_wait(full + (s), (ph >> (s)) & 1u); ph ^= 1u << (s); } while (0)
#define WAIT_XCHG(w) do { mbar_wait_cluster(xbar + (w), (ph >> (NS + (w))) & 1u); ph ^= 1u << (NS + (w)); } while (0)

	// L2 policies (hints bit 0): the `cur` fields are read exactly once per sweep -> evict-first; the temp fields are
	// read again within a tile time (the cross-line neighbours of the adjacent tiles, the second copy of temp.T) -> evict-last
	const unsigned long long pol_once = l2_evict_first(), pol_keep = l2_evict_last();
	const int streaming = hints & 2;               // hints bit 1: streaming result stores
	// box coordinates of (field tile, cross-line shift dj) of a tile
	auto issue = [&](int s, const CUtensorMap *map, int tile, int shift, int keep) {
		const int a = tile / ktiles, k0 = (tile - a * ktiles) * NL;       // a = j (x lines) or i (y lines)
		mbar_expect_tx(full + s, SLOT_BYTES);
		int c1 = 0, c2 = 0, c3, c4;
		if (DIR == 0) {
			const int j = min(max(a + shift, 0), L.ny - 1);
			c1 = (int)crank * GL; c3 = j & L.jbm; c4 = j >> L.jbs;
		} else {
			// (chunk inside a y-block, y-block): the upper half starts GL chunks further along the line
			if (L.nblk > 1) c2 = ((int)crank * nloc) >> L.jbs; else c1 = (int)crank * GL;
			c3 = 0; c4 = a + 1 + shift;
		}
		if (hints & 1) tma_load_5d_hint(SLOTP(s), map, full + s, k0, c1, c2, c3, c4, keep ? pol_keep : pol_once);
		else tma_load_5d(SLOTP(s), map, full + s, k0, c1, c2, c3, c4);
	};
	// Slot plan (five slots, eleven copies per tile, each issued a solve phase or more before its use):
	//   slot 0: temp[DIR]            (resident from the u,v,w phase to the dissipation function)
	//   slot 1: temp.T -> the first other temp component
	//   slot 2: cur.u  -> the second other temp component
	//   slot 3: cur.v  -> temp[DIR] of the cross-line neighbour below -> cur.T
	//   slot 4: cur.w  -> temp[DIR] of the cross-line neighbour above -> temp.T (relaxation of T at the end)
	auto issue_A = [&](int tile) {                 // u,v,w phase inputs but cur.w (its slot is busy until the tile ends)
		issue(0, &TM.temp[DIR], tile, 0, 1);
		issue(1, &TM.temp[3], tile, 0, 1);
		issue(2, &TM.cur[0], tile, 0, 0);
		issue(3, &TM.cur[1], tile, 0, 0);
	};
	// descriptor bytes: NL per tile row, thread <-> tile row; CL 2: plus the row of the linearisation layer next to this half
	auto issue_roles = [&](int tile) {
		const int a = tile / ktiles, k0 = (tile - a * ktiles) * NL;
		if (t < GP * M) {
			const int r = min(row0 + t, n - 1);
			const long long o = DIR == 0 ? L.idx(r, a, k0) : L.idx(a, r, k0);
#pragma unroll
			for (int c = 0; c < NL / 8; c++) cp_async8(roles + (size_t)t * NL + c * 8, A.role + o + c * 8);
		}
		if (CL == 2 && t < 4 * NL) {
			const int q = t / NL, ll = t - q * NL;
			const int r = crank == 0 ? nloc : row0 - 1, kk = min(k0 + ll, L.nz - 1);
			const long long o = DIR == 0 ? L.idx(r, a, kk) : L.idx(a, r, kk);
			cp_async_elem<FT>(edge + t, A.temp[q] + o);
		}
		if (XS && t < 8 * NL) {              // the guard planes: rows -1 and n of the slab (kept current by the neighbours' sweeps)
			const int side = t / (4 * NL), q = (t / NL) & 3, ll = t % NL;
			cp_async_elem<FT>(edge + t, A.temp[q] + L.idx(side ? n : -1, a, min(k0 + ll, L.nz - 1)));
		}
		cp_async_commit();
	};

	// Tiles are handed out dynamically (the first ones statically, then in index order through an atomic counter):
	// SMs that run a little faster take more tiles, and - what matters for HBM traffic - tiles that are neighbours in
	// memory (adjacent k-tiles share their 128-byte lines, adjacent rows are each other's cross-line neighbours) are in
	// flight at about the same time on different SMs, so the second request for a line finds it in L2.  With a static
	// stride the CTAs drift apart and every line is fetched from HBM twice (ncu: 15.5 GB read against 8.7 algorithmic).
	int tile = cid;
	if (tile < ntiles) {
		if (t == 0) { issue_A(tile); issue(4, &TM.cur[2], tile, 0, 0); }
		issue_roles(tile);
	}

	// (thread 0 asks for the next tile at the top of an iteration and first looks at the answer when it issues that tile's
	// copies, half a tile later: the latency of the atomic is never waited for; the other threads learn the index behind
	// a later barrier.  CL 2: the first CTA of the pair asks and passes the answer on with the first exchange.)
	while (tile < ntiles) {
		int fetched = 0;
		if (t == 0 && crank == 0) fetched = nunits + atomicAdd(A.tile_counter, 1);
		int next_tile = ntiles;
		const int a = tile / ktiles, k0 = (tile - a * ktiles) * NL;
		const int k = k0 + l;
		const bool line_ok = k < L.nz;
		// global offsets of the chunk's rows (stores, boundary-row node values, edge-line k +- 1 loads)
		const int rc = min(r0, n - 1);              // (padding chunks: clamped, never stored)
		const bool chunk_ok = g < GL;
		const int off0 = (int)(DIR == 0 ? L.idx(rc, a, line_ok ? k : 0) : L.idx(a, rc, line_ok ? k : 0));
		int off[M];
#pragma unroll
		for (int i = 0; i < M; i++) off[i] = off0 + (chunk_ok ? i : 0) * stride;
		// distance from the chunk's last row to the next row of the line (y lines: may cross into the next y-block)
		const int rn = min(r0 + M, n - 1);
		const int step_last = (int)(DIR == 0 ? L.idx(rn, a, line_ok ? k : 0) : L.idx(a, rn, line_ok ? k : 0)) - off[M - 1];
		unsigned rowmask = (line_ok && chunk_ok) ? 0xffu : 0u;
		// rows r0 - 1 and r0 + 8 of the line inside a slot (clamped into the line like the direct-load kernel; at an open end
		// of a half the value comes from `edge` instead)
		const int e_lo = g > 0 ? (M - 1) * STR + e - GS : e;               // row r0 - 1 = last row of chunk g - 1
		const int e_hi = g + 1 < GL ? e + GS : (M - 1) * STR + e;          // row r0 + 8 = first row of chunk g + 1
		const bool edge_lo = open_lo && g == 0, edge_hi = open_hi && g == GL - 1;

#ifdef CMC_XS_TRACE
		const long long tt0 = clock64();
#endif
		cp_async_wait_all();
		WAIT_SLOT(0); WAIT_SLOT(1); WAIT_SLOT(2); WAIT_SLOT(3); WAIT_SLOT(4);
		__syncthreads();            // roles (cp.async of every thread) have landed
#ifdef CMC_XS_TRACE
		const long long tt1 = clock64();
		long long tlast = tt1;
#define TRACE(k) do { if (t == 0) { const long long now__ = clock64(); atomicAdd(A.tile_counter + 12 + (k), (int)((now__ - tlast) >> 6)); tlast = now__; } } while (0)
#else
#define TRACE(k) do { } while (0)
#endif

		unsigned rw0 = 0, rw1 = 0;
#pragma unroll
		for (int i = 0; i < 4; i++) {
			rw0 |= (unsigned)roles[(min(g * M, nloc - M) + i) * NL + l] << (8 * i);
			rw1 |= (unsigned)roles[(min(g * M, nloc - M) + 4 + i) * NL + l] << (8 * i);
		}
		if (!rowmask) { rw0 = 0; rw1 = 0; }
#define ROLE(i) (((i) < 4 ? rw0 >> (8 * (i)) : rw1 >> (8 * ((i) - 4))) & 0xffu)
		unsigned segmask = 0, inmask = 0;
#pragma unroll
		for (int i = 0; i < M; i++) {
			segmask |= (ROLE(i) & R_SEG) ? (1u << i) : 0u;
			inmask |= (ROLE(i) & R_IN) ? (1u << i) : 0u;
		}
		const bool any_int = ((rw0 | rw1) & (R_INT * 0x01010101u)) != 0;
		const unsigned holes = inmask & ~segmask;      // fluid cells outside every segment (dropped runs)
		const unsigned full_m = rowmask;
		const unsigned segfull = segmask;

		// ======================================= phase V: u, v, w ==========================================
		FT cp[M], lp[M], dp[3][M];
		FT b7 = FT(1), rr;
		{
			FT V[M], Tl[M];
#pragma unroll
			for (int i = 0; i < M; i++) {
				V[i] = SLOTP(0)[i * STR + e];
				Tl[i] = SLOTP(1)[i * STR + e];
				dp[0][i] = SLOTP(2)[i * STR + e];
				dp[1][i] = SLOTP(3)[i * STR + e];
				dp[2][i] = SLOTP(4)[i * STR + e];
			}
			const FT Tlo = edge_lo ? edge[3 * NL + l] : SLOTP(1)[e_lo], Thi = edge_hi ? edge[(XS ? 7 : 3) * NL + l] : SLOTP(1)[e_hi];
			slot_reads_done();
			__syncthreads();        // every thread has its inputs in registers: slots 1-4 are free (slot 0 stays)
			if (t == 0) {
				issue(1, &TM.temp[QO1], tile, 0, 1);
				issue(2, &TM.temp[QO2], tile, 0, 1);
				issue(3, &TM.temp[DIR], tile, -1, 1);
				issue(4, &TM.temp[DIR], tile, +1, 1);
			}
#pragma unroll
			for (int i = 0; i < M; i++) {
				dp[0][i] *= K.c3dt; dp[1][i] *= K.c3dt; dp[2][i] *= K.c3dt;
				dp[DIR][i] -= K.v_T * cdiff<FT>(Tl, Tlo, Thi, i, K.inv2h);
			}
#pragma unroll
			for (int i = 0; i < M; i++) {
				const unsigned r = ROLE(i);
				const FT Vh = V[i] * K.inv2h;
				FT a_ = -Vh - K.vis_v, c = Vh - K.vis_v, b = K.b_v;
				FT d0 = dp[0][i], d1 = dp[1][i], d2 = dp[2][i];
				if ((r & (R_SEG | R_PRE)) != R_INT) {       // rare: boundary row, cell outside every segment, or shared-cell fold
					const bool vfree = r & R_VFREE;
					if (r & R_INT) {                        // R_PRE: the next cell ends this segment AND starts the next one
						if (vfree) b += FT(0.5) * c;
						else {
							const int idn = off[i] + (i == M - 1 ? step_last : stride);
							d0 -= c * A.nodev[0][idn]; d1 -= c * A.nodev[1][idn]; d2 -= c * A.nodev[2][idn];
						}
						c = FT(0);
					} else if (r & (R_START | R_END)) {     // ApplyBC0 / ApplyBC1 (a shared cell keeps its start row only)
						a_ = ((r & (R_END | R_START)) == R_END && vfree) ? FT(-1) : FT(0);
						c = ((r & R_START) && vfree) ? FT(-1) : FT(0);
						b = vfree ? FT(2) : FT(1);
						d0 = d1 = d2 = FT(0);
						if (!vfree) { d0 = A.nodev[0][off[i]]; d1 = A.nodev[1][off[i]]; d2 = A.nodev[2][off[i]]; }
					} else { a_ = FT(0); c = FT(0); b = FT(1); d0 = d1 = d2 = FT(0); }
				}
				CMC_ELIM_ROW(i, a_, b, c)
				if (i == M - 1) { dp[0][i] = d0; dp[1][i] = d1; dp[2][i] = d2; }
				else if (i == 0) { dp[0][0] = d0 * rr; dp[1][0] = d1 * rr; dp[2][0] = d2 * rr; }
				else {
					dp[0][i] = (d0 - a_ * dp[0][i - 1]) * rr;
					dp[1][i] = (d1 - a_ * dp[1][i - 1]) * rr;
					dp[2][i] = (d2 - a_ * dp[2][i - 1]) * rr;
				}
			}
		}
		// head of the NEXT chunk of the line (thread t + NL): the lower lanes by shuffle, the last NL lanes of a warp from the
		// next warp's first NL lanes through headx
#define NEXT_HEAD(dst, val, slot_)                                                     \
		do {                                                                           \
			const FT sh__ = __shfl_down_sync(0xffffffffu, (val), NL);                  \
			(dst) = lane < 32 - NL ? sh__ : (warp + 1 < NW ? headx[(slot_) * (NW * NL) + (warp + 1) * NL + (lane - (32 - NL))] : FT(0)); \
		} while (0)
		FT E[3];
		FT xl[3] = {FT(0), FT(0), FT(0)}, xr[3] = {FT(0), FT(0), FT(0)}, Es = FT(0), Es2 = FT(0);      // CL 2 / XS: the adjacent rows of the other half / slabs, this separator's spikes
		{
			// coupling of the first interior row to the two separators: x_0 = y0 - v0*E(g-1) - w0*E(g)
			FT y0[3] = {dp[0][M - 2], dp[1][M - 2], dp[2][M - 2]}, v0 = lp[M - 2], w0 = cp[M - 2];
#pragma unroll
			for (int i = M - 3; i >= 0; i--) {
				y0[0] = dp[0][i] - cp[i] * y0[0]; y0[1] = dp[1][i] - cp[i] * y0[1]; y0[2] = dp[2][i] - cp[i] * y0[2];
				v0 = lp[i] - cp[i] * v0; w0 = -cp[i] * w0;
			}
			if (lane < NL) {
				headx[0 * (NW * NL) + warp * NL + lane] = y0[0]; headx[1 * (NW * NL) + warp * NL + lane] = y0[1];
				headx[2 * (NW * NL) + warp * NL + lane] = y0[2]; headx[3 * (NW * NL) + warp * NL + lane] = v0;
				headx[4 * (NW * NL) + warp * NL + lane] = w0;
			}
			__syncthreads();
			FT ny0, ny1, ny2, nv, nw;
			NEXT_HEAD(ny0, y0[0], 0); NEXT_HEAD(ny1, y0[1], 1); NEXT_HEAD(ny2, y0[2], 2); NEXT_HEAD(nv, v0, 3); NEXT_HEAD(nw, w0, 4);
			const FT a7 = lp[M - 1], c7 = cp[M - 1];
			rr = rcp<FT>(b7 - a7 * cp[M - 2] - c7 * nv);
			FT Rd[3];
			Rd[0] = (dp[0][M - 1] - a7 * dp[0][M - 2] - c7 * ny0) * rr;
			Rd[1] = (dp[1][M - 1] - a7 * dp[1][M - 2] - c7 * ny1) * rr;
			Rd[2] = (dp[2][M - 1] - a7 * dp[2][M - 2] - c7 * ny2) * rr;
			if (DIR == 0) TRACE(1);
			if (XS) {
				// open-ended slab: E = Y - P * x_left - Q * x_right (kernels_fast.cu MODE 1), P / Q two more right-hand sides
				FT Re[5] = {Rd[0], Rd[1], Rd[2], FT(0), FT(0)}, Xe[5];
				FT Ra = -a7 * lp[M - 2] * rr;
				if (g == 0) { Re[3] = Ra; Ra = FT(0); }
				if (g == GL - 1) Re[4] = c7 * rr;            // (the chunk after it is an identity chunk or absent: nv == nw == 0)
				reduced_solve<FT, 5, GP, GS, NL>(sys, sol, g, e, Ra, -c7 * nw * rr, Re, Xe);
				TRACE(2);
				constexpr int W = LLWords<FT>::W;
				if (g == 0 || g == GL - 1) {
					// first row: x = f - pf x_left - qf x_right (table entries 0-2, 6, 7); last row: x = l - pl x_left - ql x_right (3-5, 8, 9)
					const bool first = g == 0;
					FT c5[5];
					if (first) {
#pragma unroll
						for (int q = 0; q < 3; q++) c5[q] = y0[q] - w0 * Xe[q];
						c5[3] = v0 - w0 * Xe[3]; c5[4] = -w0 * Xe[4];
					} else {
#pragma unroll
						for (int q = 0; q < 5; q++) c5[q] = Xe[q];
					}
					const int vi[5] = {first ? 0 : 3, first ? 1 : 4, first ? 2 : 5, first ? 6 : 8, first ? 7 : 9};
#pragma unroll
					for (int q = 0; q < 5; q++) xin[vi[q] * NL + l] = c5[q];
				}
				__syncthreads();                             // this slab's coefficients are in xin
				// ... and go to every other slab's table, one word per thread and turn (128-byte stores, all warps share the work)
				for (int idx = t; idx < (A.xs_P - 1) * 10 * W * NL; idx += STR) {
					const int ll = idx % NL, r1 = idx / NL, w = r1 % W, r2 = r1 / W, v = r2 % 10, ro = r2 / 10, r = ro < A.xs_me ? ro : ro + 1;
					*reinterpret_cast<volatile unsigned long long *>(A.xs_tab_to[r] + ((((size_t)A.xs_me * ntiles + tile) * 16 + v) * W + w) * NL + ll) =
						ll_word(xin[v * NL + ll], w, (unsigned)A.xs_epoch);
				}
				// all threads fetch the other slabs' words of this tile side by side (one poll each, instead of a chain of dependent
				// polls in the thread that solves the interface)
				TRACE(3);
#ifdef CMC_XS_TRACE
				const long long tw0 = clock64();
#endif
				for (int idx = t; idx < (A.xs_P - 1) * 10 * W * NL; idx += STR) {
					const int ll = idx % NL, r1 = idx / NL, w = r1 % W, r2 = r1 / W, v = r2 % 10, ro = r2 / 10, r = ro < A.xs_me ? ro : ro + 1;
					xw[idx] = ll_poll(A.xs_tab + ((((size_t)r * ntiles + tile) * 16 + v) * W + w) * NL + ll, (unsigned)A.xs_epoch);
				}
				__syncthreads();                             // this slab's coefficients are in xin, the others' in xw
				TRACE(4);
#ifdef CMC_XS_TRACE
				const long long tw1 = clock64();
#endif
				if (g == 0) {                                // this line's interface system
					auto C = [&](int r, int v) -> FT {
						if (r == A.xs_me) return xin[v * NL + l];
						const int ro = r < A.xs_me ? r : r - 1;
						return ll_value<FT>(xw + ((ro * 10 + v) * W) * NL + l, NL);
					};
					FT il[3] = {FT(0), FT(0), FT(0)}, ir[3] = {FT(0), FT(0), FT(0)};
					if (A.xs_P == 2) {
						// two slabs: L_0 + ql_0 F_1 = l_0, F_1 + pf_1 L_0 = f_1
						const FT ql0 = C(0, 9), pf1 = C(1, 6), den = rcp<FT>(FT(1) - ql0 * pf1);
#pragma unroll
						for (int q = 0; q < 3; q++) {
							const FT l0 = C(0, 3 + q), f1 = C(1, q), L0 = (l0 - ql0 * f1) * den;
							if (A.xs_me == 0) ir[q] = f1 - pf1 * L0; else il[q] = L0;
						}
					} else
						xs_interface<FT, 3>(C, A.xs_P, A.xs_me, 0, 3, 6, il, ir);
#pragma unroll
					for (int q = 0; q < 3; q++) { xown[q * NL + l] = il[q]; xown[(4 + q) * NL + l] = ir[q]; }
				}
#ifdef CMC_XS_TRACE
				if (t == 0) { atomicAdd(A.tile_counter + 2, (int)((tw1 - tw0) >> 6)); atomicAdd(A.tile_counter + 3, (int)((clock64() - tw1) >> 6)); atomicAdd(A.tile_counter + 4, 1); }
#endif
				__syncthreads();
#pragma unroll
				for (int q = 0; q < 3; q++) { xl[q] = xown[q * NL + l]; xr[q] = xown[(4 + q) * NL + l]; }
				TRACE(5);
				Es = Xe[3]; Es2 = Xe[4];
#pragma unroll
				for (int q = 0; q < 3; q++) E[q] = Xe[q] - Es * xl[q] - Es2 * xr[q];
			} else if (CL == 2) {
				// open-ended half: E = Y - S * x_other, S = one more right-hand side: the left spike of chunk 0 (upper half,
				// x_other = last row of the lower half) or the last separator's coupling c7 (lower half, x_other = first row
				// of the upper half).  Same algebra as the slab-coupled x-sweep (kernels_fast.cu MODE 1).
				FT Re[4] = {Rd[0], Rd[1], Rd[2], FT(0)}, Xe[4];
				FT Ra = -a7 * lp[M - 2] * rr;
				if (open_lo && g == 0) { Re[3] = Ra; Ra = FT(0); }
				if (open_hi && g == GL - 1) Re[3] = c7 * rr;         // (the chunk after it is an identity chunk or absent: nv == nw == 0)
				reduced_solve<FT, 4, GP, GS, NL>(sys, sol, g, e, Ra, -c7 * nw * rr, Re, Xe);
				// interface coefficients of this half: lower: last row x = c[q] - s * x_first(upper); upper: first row
				// x = c[q] - s * x_last(lower)
				if ((open_hi && g == GL - 1) || (open_lo && g == 0)) {
					FT cq[3], sp;
					if (open_hi) { cq[0] = Xe[0]; cq[1] = Xe[1]; cq[2] = Xe[2]; sp = Xe[3]; }
					else { cq[0] = y0[0] - w0 * Xe[0]; cq[1] = y0[1] - w0 * Xe[1]; cq[2] = y0[2] - w0 * Xe[2]; sp = v0 - w0 * Xe[3]; }
					const unsigned pa = peer_smem(xin, crank ^ 1u);
#pragma unroll
					for (int q = 0; q < 3; q++) { xown[q * NL + l] = cq[q]; st_cluster(pa + (unsigned)((q * NL + l) * sizeof(FT)), cq[q]); }
					xown[3 * NL + l] = sp; st_cluster(pa + (unsigned)((3 * NL + l) * sizeof(FT)), sp);
					mbar_arrive_peer(peer_smem(xbar + 0, crank ^ 1u));
				}
				if (t == 0 && crank == 0) {          // the next tile's index travels with the exchange
					st_cluster(peer_smem(next_box, 1u), fetched);
					mbar_arrive_peer(peer_smem(xbar + 0, 1u));
				}
				__syncthreads();                     // xown is written
				WAIT_XCHG(0);                        // xin is written (and, upper half, next_box)
				if (crank == 1 && t == 0) fetched = *next_box;
				{
					// x_last(lower) = (cl - sl * cu) / (1 - sl * su),  x_first(upper) = cu - su * x_last(lower)
					const FT *lo_ = open_hi ? xown : xin, *up_ = open_hi ? xin : xown;
					const FT sl = lo_[3 * NL + l], su = up_[3 * NL + l];
					const FT den = rcp<FT>(FT(1) - sl * su);
#pragma unroll
					for (int q = 0; q < 3; q++) {
						const FT xlast = (lo_[q * NL + l] - sl * up_[q * NL + l]) * den;
						const FT xfirst = up_[q * NL + l] - su * xlast;
						if (open_hi) xr[q] = xfirst; else xl[q] = xlast;
					}
				}
				Es = Xe[3];
#pragma unroll
				for (int q = 0; q < 3; q++) E[q] = Xe[q] - Es * (open_hi ? xr[q] : xl[q]);
			} else {
				reduced_solve<FT, 3, GP, GS, NL>(sys, sol, g, e, -a7 * lp[M - 2] * rr, -c7 * nw * rr, Rd, E);    // every row's solution -> sol[]
				if (DIR == 0) TRACE(2);
			}
		}
		// back substitution in place (dp[q] <- x: retires cp / lp), then store u, v, w and the relaxed linearisation layer
		{
			const FT Sl = ((CL == 2 || XS) && g > 0) ? sol[3 * STR + e - GS] : FT(0);
			const FT Sl2 = (XS && g > 0) ? sol[4 * STR + e - GS] : FT(0);
			(void)Es; (void)Es2;
#pragma unroll
			for (int q = 0; q < 3; q++) {
				FT El = g > 0 ? sol[q * STR + e - GS] : FT(0);
				if (CL == 2) El = g > 0 ? El - Sl * (open_hi ? xr[q] : xl[q]) : xl[q];      // chunk 0 of the upper half: row -1 is x_last(lower)
				if (XS) El = g > 0 ? El - Sl * xl[q] - Sl2 * xr[q] : xl[q];                 // chunk 0: row -1 is the lower slab's last row
				dp[q][M - 1] = E[q];
#pragma unroll
				for (int i = M - 2; i >= 0; i--) dp[q][i] = dp[q][i] - lp[i] * El - cp[i] * dp[q][i + 1];
			}
		}
		WAIT_SLOT(1); WAIT_SLOT(2);
#pragma unroll
		for (int q = 0; q < 3; q++) {
			FT (&x)[M] = dp[q];
			const FT *sq = q == DIR ? SLOTP(0) : q == QO1 ? SLOTP(1) : SLOTP(2);
			FT tq[M];
#pragma unroll
			for (int i = 0; i < M; i++) tq[i] = sq[i * STR + e];
			if (holes) {                      // merge those with the OLD value of `next` (MergeFieldTo reads whatever is there)
#pragma unroll
				for (int i = 0; i < M; i++)
					if (holes & (1u << i)) x[i] = A.next[q][off[i]];
			}
			relax8<FT, DIR>(tq, x, inmask, A.extra_merge);
			store8_stream<FT>(A.temp_out[q], off, full_m, tq, streaming);
			store8_stream<FT>(A.next[q], off, segfull, x, streaming);
			push_planes<FT, DIR, XS ? 2 : 0>(A, q, a, g, XS ? GL : GP, off, full_m, segfull, tq, x);
		}

		if (DIR == 0) TRACE(6);
		// ======================================= phase T ======================================================
		FT (&dT)[M] = dp[0];
		{
			FT diss[M], V[M];
			{
				// dissipation function of the sweep direction (TimeLayer3D.h:554-588), accumulated component by component:
				//   X: 2 u_x^2 + v_x^2 + w_x^2 + v_x u_y + w_x u_z ; Y: u_y^2 + 2 v_y^2 + w_y^2 + u_y v_x + w_y v_z
				// c1, c2: the two cross-line derivatives of temp[DIR]; c1 pairs with component QA, c2 with QB
				constexpr int QA = DIR == 0 ? 1 : 0, QB = 2;
				FT c1[M], c2[M];
				WAIT_SLOT(3); WAIT_SLOT(4);
#pragma unroll
				for (int i = 0; i < M; i++) c1[i] = (SLOTP(4)[i * STR + e] - SLOTP(3)[i * STR + e]) * K.inv2h1;
				slot_reads_done();
				__syncthreads();        // the cross-line neighbours are consumed: slots 3 / 4 take cur.T and temp.T
				if (t == 0) {
					issue(3, &TM.cur[3], tile, 0, 0);
					issue(4, &TM.temp[3], tile, 0, 0);       // last use of this tile's temp.T
				}
				{
					// second cross direction (k +- 1): the neighbouring lines of the tile, the two edge lines from L2 / HBM
					// (issuing those loads earlier - in registers, or as cp.async into shared memory together with the
					// descriptor bytes - costs more than it hides: measured, profiles/r02_variants.md)
					FT p2[M], m2[M];
					if (l == NL - 1) {
#pragma unroll
						for (int i = 0; i < M; i++) p2[i] = A.temp[DIR][off[i] + 1];
					} else {
#pragma unroll
						for (int i = 0; i < M; i++) p2[i] = SLOTP(0)[i * STR + e + 1];
					}
					if (l == 0) {
#pragma unroll
						for (int i = 0; i < M; i++) m2[i] = A.temp[DIR][off[i] - 1];
					} else {
#pragma unroll
						for (int i = 0; i < M; i++) m2[i] = SLOTP(0)[i * STR + e - 1];
					}
#pragma unroll
					for (int i = 0; i < M; i++) c2[i] = (p2[i] - m2[i]) * K.inv2h2;
				}
#pragma unroll
				for (int i = 0; i < M; i++) diss[i] = FT(0);
#pragma unroll
				for (int q = 0; q < 3; q++) {
					const FT *sq = q == DIR ? SLOTP(0) : q == QO1 ? SLOTP(1) : SLOTP(2);
					FT f[M];
#pragma unroll
					for (int i = 0; i < M; i++) f[i] = sq[i * STR + e];
					const FT lo = edge_lo ? edge[q * NL + l] : sq[e_lo], hi = edge_hi ? edge[((XS ? 4 : 0) + q) * NL + l] : sq[e_hi];
#pragma unroll
					for (int i = 0; i < M; i++) {
						const FT d = cdiff<FT>(f, lo, hi, i, K.inv2h);
						FT w = q == DIR ? d + d : d;
						if (q == QA) w += c1[i];
						if (q == QB) w += c2[i];
						diss[i] += d * w;
					}
					if (q == DIR) {
#pragma unroll
						for (int i = 0; i < M; i++) V[i] = f[i];
					}
				}
				if (!any_int) {      // no interior row: the (clamped) neighbour values above were never meant to be used
#pragma unroll
					for (int i = 0; i < M; i++) diss[i] = FT(0);
				}
			}
			FT cT[M];
			WAIT_SLOT(3);
#pragma unroll
			for (int i = 0; i < M; i++) cT[i] = SLOTP(3)[i * STR + e];
			slot_reads_done();
			__syncthreads();        // slots 0 - 3 have been consumed (slot 4 keeps temp.T for the end of the tile)
			if (t == 0) {
				if (fetched < ntiles) issue_A(fetched);
				if (CL == 1 || crank == 0) *next_box = fetched;      // read by everybody after the next barrier (upper half: already there)
			}
#pragma unroll
			for (int i = 0; i < M; i++) {
				const unsigned r = ROLE(i);
				const FT Vh = V[i] * K.inv2h;
				FT a_ = -Vh - K.vis_T, c = Vh - K.vis_T, b = K.b_T;
				FT d = cT[i] * K.c3dt + K.t_phi * diss[i];
				if ((r & (R_SEG | R_PRE)) != R_INT) {
					const bool tfree = r & R_TFREE;
					if (r & R_INT) {
						if (tfree) b += FT(0.5) * c;
						else d -= c * A.nodev[3][off[i] + (i == M - 1 ? step_last : stride)];
						c = FT(0);
					} else if (r & (R_START | R_END)) {
						a_ = ((r & (R_END | R_START)) == R_END && tfree) ? FT(-1) : FT(0);
						c = ((r & R_START) && tfree) ? FT(-1) : FT(0);
						b = tfree ? FT(2) : FT(1);
						d = tfree ? FT(0) : A.nodev[3][off[i]];
					} else { a_ = FT(0); c = FT(0); b = FT(1); d = FT(0); }
				}
				CMC_ELIM_ROW(i, a_, b, c)
				if (i == M - 1) dT[i] = d;
				else if (i == 0) dT[0] = d * rr;
				else dT[i] = (d - a_ * dT[i - 1]) * rr;
			}
		}
		{
			FT y0 = dT[M - 2], v0 = lp[M - 2], w0 = cp[M - 2];
#pragma unroll
			for (int i = M - 3; i >= 0; i--) { y0 = dT[i] - cp[i] * y0; v0 = lp[i] - cp[i] * v0; w0 = -cp[i] * w0; }
			if (lane < NL) {
				headx[0 * (NW * NL) + warp * NL + lane] = y0; headx[3 * (NW * NL) + warp * NL + lane] = v0; headx[4 * (NW * NL) + warp * NL + lane] = w0;
			}
			__syncthreads();
			next_tile = *next_box;
			if (next_tile < ntiles) issue_roles(next_tile);
			FT ny0, nv, nw;
			NEXT_HEAD(ny0, y0, 0); NEXT_HEAD(nv, v0, 3); NEXT_HEAD(nw, w0, 4);
			const FT a7 = lp[M - 1], c7 = cp[M - 1];
			rr = rcp<FT>(b7 - a7 * cp[M - 2] - c7 * nv);
			FT Rd[1] = {(dT[M - 1] - a7 * dT[M - 2] - c7 * ny0) * rr}, ET[1];
			FT El;
			if (DIR == 0) TRACE(7);
			if (XS) {
				FT Re[3] = {Rd[0], FT(0), FT(0)}, Xe[3];
				FT Ra = -a7 * lp[M - 2] * rr;
				if (g == 0) { Re[1] = Ra; Ra = FT(0); }
				if (g == GL - 1) Re[2] = c7 * rr;
				reduced_solve<FT, 3, GP, GS, NL>(sys, sol, g, e, Ra, -c7 * nw * rr, Re, Xe);
				TRACE(8);
				constexpr int W = LLWords<FT>::W;
				if (g == 0 || g == GL - 1) {
					// table entries: 10 f, 12 pf, 13 qf (first row); 11 l, 14 pl, 15 ql (last row)
					const bool first = g == 0;
					const FT c3[3] = {first ? y0 - w0 * Xe[0] : Xe[0], first ? v0 - w0 * Xe[1] : Xe[1], first ? -w0 * Xe[2] : Xe[2]};
					const int vi[3] = {first ? 10 : 11, first ? 12 : 14, first ? 13 : 15};
#pragma unroll
					for (int q = 0; q < 3; q++) xin[vi[q] * NL + l] = c3[q];
				}
				__syncthreads();
				for (int idx = t; idx < (A.xs_P - 1) * 6 * W * NL; idx += STR) {
					const int ll = idx % NL, r1 = idx / NL, w = r1 % W, r2 = r1 / W, v = 10 + r2 % 6, ro = r2 / 6, r = ro < A.xs_me ? ro : ro + 1;
					*reinterpret_cast<volatile unsigned long long *>(A.xs_tab_to[r] + ((((size_t)A.xs_me * ntiles + tile) * 16 + v) * W + w) * NL + ll) =
						ll_word(xin[v * NL + ll], w, (unsigned)A.xs_epoch);
				}
#ifdef CMC_XS_TRACE
				const long long tw0 = clock64();
#endif
				for (int idx = t; idx < (A.xs_P - 1) * 6 * W * NL; idx += STR) {
					const int ll = idx % NL, r1 = idx / NL, w = r1 % W, r2 = r1 / W, v = r2 % 6, ro = r2 / 6, r = ro < A.xs_me ? ro : ro + 1;
					xw[idx] = ll_poll(A.xs_tab + ((((size_t)r * ntiles + tile) * 16 + 10 + v) * W + w) * NL + ll, (unsigned)A.xs_epoch);
				}
				__syncthreads();
#ifdef CMC_XS_TRACE
				if (t == 0) atomicAdd(A.tile_counter + 5, (int)((clock64() - tw0) >> 6));
#endif
				if (g == 0) {
					auto C = [&](int r, int v) -> FT {
						if (r == A.xs_me) return xin[v * NL + l];
						const int ro = r < A.xs_me ? r : r - 1;
						return ll_value<FT>(xw + ((ro * 6 + (v - 10)) * W) * NL + l, NL);
					};
					FT il[1] = {FT(0)}, ir[1] = {FT(0)};
					if (A.xs_P == 2) {
						const FT ql0 = C(0, 15), pf1 = C(1, 12), l0 = C(0, 11), f1 = C(1, 10);
						const FT L0 = (l0 - ql0 * f1) * rcp<FT>(FT(1) - ql0 * pf1);
						if (A.xs_me == 0) ir[0] = f1 - pf1 * L0; else il[0] = L0;
					} else
						xs_interface<FT, 1>(C, A.xs_P, A.xs_me, 10, 11, 12, il, ir);
					xown[3 * NL + l] = il[0]; xown[7 * NL + l] = ir[0];
				}
				__syncthreads();
				const FT xlT = xown[3 * NL + l], xrT = xown[7 * NL + l];
				TRACE(9);
				ET[0] = Xe[0] - Xe[1] * xlT - Xe[2] * xrT;
				El = g > 0 ? sol[e - GS] - sol[STR + e - GS] * xlT - sol[2 * STR + e - GS] * xrT : xlT;
			} else if (CL == 2) {
				FT Re[2] = {Rd[0], FT(0)}, Xe[2];
				FT Ra = -a7 * lp[M - 2] * rr;
				if (open_lo && g == 0) { Re[1] = Ra; Ra = FT(0); }
				if (open_hi && g == GL - 1) Re[1] = c7 * rr;
				reduced_solve<FT, 2, GP, GS, NL>(sys, sol, g, e, Ra, -c7 * nw * rr, Re, Xe);
				if ((open_hi && g == GL - 1) || (open_lo && g == 0)) {
					const FT cq = open_hi ? Xe[0] : y0 - w0 * Xe[0], sp = open_hi ? Xe[1] : v0 - w0 * Xe[1];
					const unsigned pa = peer_smem(xin, crank ^ 1u);
					xown[4 * NL + l] = cq; st_cluster(pa + (unsigned)((4 * NL + l) * sizeof(FT)), cq);
					xown[5 * NL + l] = sp; st_cluster(pa + (unsigned)((5 * NL + l) * sizeof(FT)), sp);
					mbar_arrive_peer(peer_smem(xbar + 1, crank ^ 1u));
				}
				__syncthreads();
				WAIT_XCHG(1);
				const FT *lo_ = open_hi ? xown : xin, *up_ = open_hi ? xin : xown;
				const FT sl = lo_[5 * NL + l], su = up_[5 * NL + l];
				const FT xlast = (lo_[4 * NL + l] - sl * up_[4 * NL + l]) * rcp<FT>(FT(1) - sl * su);
				const FT xo = open_hi ? up_[4 * NL + l] - su * xlast : xlast;          // the other half's adjacent row
				ET[0] = Xe[0] - Xe[1] * xo;
				El = g > 0 ? sol[e - GS] - sol[STR + e - GS] * xo : (open_lo ? xo : FT(0));
			} else {
				reduced_solve<FT, 1, GP, GS, NL>(sys, sol, g, e, -a7 * lp[M - 2] * rr, -c7 * nw * rr, Rd, ET);
				El = g > 0 ? sol[e - GS] : FT(0);
				if (DIR == 0) TRACE(8);
			}
			FT x[M], tq[M];
			x[M - 1] = ET[0];
#pragma unroll
			for (int i = M - 2; i >= 0; i--) x[i] = dT[i] - lp[i] * El - cp[i] * x[i + 1];
			WAIT_SLOT(4);
#pragma unroll
			for (int i = 0; i < M; i++) tq[i] = SLOTP(4)[i * STR + e];
			slot_reads_done();
			__syncthreads();        // slot 4 consumed (and sol / headx are free for the next tile)
			if (t == 0 && next_tile < ntiles) issue(4, &TM.cur[2], next_tile, 0, 0);
			if (holes) {
#pragma unroll
				for (int i = 0; i < M; i++)
					if (holes & (1u << i)) x[i] = A.next[3][off[i]];
			}
			relax8<FT, DIR>(tq, x, inmask, A.extra_merge);
			store8_stream<FT>(A.temp_out[3], off, full_m, tq, streaming);
			store8_stream<FT>(A.next[3], off, segfull, x, streaming);
			push_planes<FT, DIR, XS ? 2 : 0>(A, 3, a, g, XS ? GL : GP, off, full_m, segfull, tq, x);
		}
		if (DIR == 0) TRACE(10);
#ifdef CMC_XS_TRACE
		if (t == 0) { atomicAdd(A.tile_counter + 6 + 3 * DIR, (int)((clock64() - tt0) >> 6)); atomicAdd(A.tile_counter + 7 + 3 * DIR, (int)((tt1 - tt0) >> 6)); atomicAdd(A.tile_counter + 8 + 3 * DIR, 1); }
#endif
		tile = next_tile;
	}
	cp_async_wait_all();
	if (CL == 2) cluster_sync_all();                // the peer may still be storing into this CTA's shared memory
#undef ROLE
#undef SLOTP
#undef WAIT_SLOT
#undef WAIT_XCHG
#undef NEXT_HEAD
#undef TRACE
}

template <typename FT, int GP, int NL, int CL, int XS>
static size_t tma_smem_bytes()
{
	const size_t STR = (size_t)GP * NL;
	return sizeof(FT) * (5 * STR * M + reduced_scratch_elems<3 + (CL == 2 ? 1 : XS ? 2 : 0), GP, NL>() + 5 * (STR / 32) * NL + (CL == 2 ? 16 * NL : XS ? 32 * NL : 0)) + (size_t)NL * GP * 8 +
	       7 * sizeof(unsigned long long) + 16;
}

// shape of the tile a sweep uses: lines per tile, chunks per CTA, CTAs per tile (0 chunks: not supported).  Measured on B200
// (profiles/r02_variants.md, fp64): 512-row lines - 8 lines x 1 CTA (512 threads) wins over the 16-line CTA pair although the
// pair's copies run twice as fast: the tile time is set by the solve, not by the copy engine; 256-row lines - 16 lines x 1
// CTA along x (0.51 against 0.66 ms at 256^3), 8 lines and two independent CTAs per SM along y (0.48 against 0.52 ms).
struct TmaShape { int gp, nl, cl; };
static TmaShape tma_shape(const Layout &L, int dir, int forced)
{
	TmaShape S = {0, 8, 1};
	const int n = dir == 0 ? L.nx : L.ny;
	if (n % M != 0 || n / M > 128 || n / M <= 16) return S;              // 136 .. 1024 rows
	int nl = (n / M <= 32 && dir == 0) ? 16 : 8, cl = 1;
	if (n / M > 64) { forced = 0; nl = 8; cl = 2; }                      // 520 .. 1024 rows: only a CTA pair holds the line
	static const char *env = getenv("CMC_TMA_SHAPE");                    // (experiments) "8x1" / "16x1" / "16x2": lines x CTAs per tile
	if (env && !forced && n / M <= 64) {
		int fnl = 0, fcl = 0;
		if (sscanf(env, "%dx%d", &fnl, &fcl) == 2 && (fnl == 8 || fnl == 16) && (fcl == 1 || fcl == 2)) forced = fnl + 256 * fcl;
	}
	if (forced) { nl = forced & 255; cl = forced >> 8; }
	// a CTA pair takes halves of whole chunks and, along y in blocked storage, of whole y-blocks
	if (cl == 2 && (n % (2 * M) != 0 || (dir == 1 && L.nblk > 1 && (n / 2) % (1 << L.jbs) != 0))) {
		if (n / M > 64) return S;                                        // (no single-CTA alternative)
		cl = 1;
	}
	const int chunks = n / M / cl;
	if (chunks > 32 && nl == 16) nl = 8;                                 // 1024 threads do not fit
	S.gp = chunks > 32 ? 64 : 32; S.nl = nl; S.cl = cl;
	return S;
}

bool tma_sweep_supported(const Layout &L, int dir)
{
	if (dir != 0 && dir != 1) return false;
	if (L.total >= (1ll << 31)) return false;                            // 32-bit element offsets
	if (L.nblk > 1 && ((1 << L.jbs) % M != 0)) return false;
	if (tma_shape(L, dir, 0).gp == 0) return false;                      // 136 .. 1024 rows, a multiple of 8 (16 above 512)
	return encode_fn() != nullptr;
}

template <typename FT, int DIR, int GP, int NL, int CL, int XS = 0>
static bool launch_tma_one(const SweepArgs<FT> &A, cudaStream_t s)
{
	const Layout &L = A.L;
	if (DIR == 1 && L.nblk > 1 && (GP % ((1 << L.jbs) / M) != 0)) return false;
	TmaMaps TM;
	for (int q = 0; q < 4; q++) {
		if (!tensor_map_for<FT>(A.temp[q], L, DIR, GP, NL, &TM.temp[q])) return false;
		if (!tensor_map_for<FT>(A.cur[q], L, DIR, GP, NL, &TM.cur[q])) return false;
	}
	const int ntiles = (DIR == 0 ? L.ny : L.nx) * ((L.nz + NL - 1) / NL);
	const size_t smem = tma_smem_bytes<FT, GP, NL, CL, XS>() + (XS ? (size_t)(A.xs_P - 1) * 10 * (sizeof(FT) / 4) * NL * sizeof(unsigned) : 0);
	static int ctas_of[64] = {};
	static size_t smem_of[64] = {};          // (XS: the shared-memory size grows with the number of slabs)
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) return false;
	auto kern = k_tma_sweep<FT, DIR, GP, NL, CL, XS>;
	if (!ctas_of[dev] || smem != smem_of[dev]) {
		smem_of[dev] = smem;
		if (cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
		int per_sm = 0, sms = 0;
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GP * NL, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); return false; }
		if (getenv("CMC_TMA_CTAS")) per_sm = std::max(1, std::min(per_sm, atoi(getenv("CMC_TMA_CTAS"))));    // (experiments)
		ctas_of[dev] = (per_sm * sms) / CL * CL;
	}
	FastConst<FT> K; K.init(A, DIR);
	// CMC_TMA_HINTS: bit 0 = L2 eviction hints on the bulk copies, bit 1 = streaming result stores (measured: neither helps)
	static const int hints = getenv("CMC_TMA_HINTS") ? atoi(getenv("CMC_TMA_HINTS")) : 0;
	if (!A.tile_counter || cudaMemsetAsync(A.tile_counter, 0, sizeof(int), s) != cudaSuccess) return false;
	int grid = std::min(ctas_of[dev], ntiles * CL);
	// XS: the kernels of all slabs wait for each other; slabs that share a device (a test configuration) must be resident together
	if (XS && A.xs_share > 1) grid = std::max(1, std::min(grid, ctas_of[dev] / A.xs_share));
	if (CL == 1) {
		kern<<<grid, GP * NL, smem, s>>>(A, K, TM, ntiles, hints);
		return true;
	}
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(GP * NL); cfg.dynamicSmemBytes = smem; cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	if (cudaLaunchKernelEx(&cfg, kern, A, K, TM, ntiles, hints) != cudaSuccess) { cudaGetLastError(); return false; }
	return true;
}

template <typename FT, int DIR>
static bool launch_tma_dir(const SweepArgs<FT> &A, cudaStream_t s)
{
	const TmaShape S = tma_shape(A.L, DIR, A.tma_shape);
	if (S.gp == 64) return S.cl == 2 ? launch_tma_one<FT, DIR, 64, 8, 2>(A, s) : launch_tma_one<FT, DIR, 64, 8, 1>(A, s);
	if (S.gp != 32) return false;
	if (S.nl == 16) return S.cl == 2 ? launch_tma_one<FT, DIR, 32, 16, 2>(A, s) : launch_tma_one<FT, DIR, 32, 16, 1>(A, s);
	return S.cl == 2 ? launch_tma_one<FT, DIR, 32, 8, 2>(A, s) : launch_tma_one<FT, DIR, 32, 8, 1>(A, s);
}

// fused slab-coupled x-sweep: chunks per slab line -> lines per tile (the same for all slabs of a grid: tiles index the
// coefficient tables).  CMC_XS_NL=8|16 overrides (experiments).
int tma_xs_lines(const Layout &L)
{
	const int n = L.nx;
	if (n % M != 0 || n < 64 || n > 512 || L.total >= (1ll << 31) || (L.nblk > 1 && ((1 << L.jbs) % M != 0)) || !encode_fn()) return 0;
	if (n / M > 32) return 8;
	static const int force = getenv("CMC_XS_NL") ? atoi(getenv("CMC_XS_NL")) : 0;
	if (force == 8 && n / M > 8) return 8;
	if (force == 16) return 16;
	return 16;
}
template <typename FT>
bool launch_tma_xs(const SweepArgs<FT> &A, cudaStream_t s, long long *launches)
{
	const int nl = tma_xs_lines(A.L), chunks = A.L.nx / M;
	if (!nl) return false;
	bool ok;
	if (chunks > 32) ok = launch_tma_one<FT, 0, 64, 8, 1, 1>(A, s);
	else if (chunks > 16) ok = nl == 16 ? launch_tma_one<FT, 0, 32, 16, 1, 1>(A, s) : launch_tma_one<FT, 0, 32, 8, 1, 1>(A, s);
	else if (chunks > 8) ok = nl == 16 ? launch_tma_one<FT, 0, 16, 16, 1, 1>(A, s) : launch_tma_one<FT, 0, 16, 8, 1, 1>(A, s);
	else ok = launch_tma_one<FT, 0, 8, 16, 1, 1>(A, s);
	if (ok && launches) (*launches)++;
	return ok;
}
template bool launch_tma_xs<float>(const SweepArgs<float> &, cudaStream_t, long long *);
template bool launch_tma_xs<double>(const SweepArgs<double> &, cudaStream_t, long long *);

template <typename FT>
bool launch_tma_sweep(int dir, const SweepArgs<FT> &A, cudaStream_t s, long long *launches)
{
	if (!tma_sweep_supported(A.L, dir)) return false;
	const bool ok = dir == 0 ? launch_tma_dir<FT, 0>(A, s) : launch_tma_dir<FT, 1>(A, s);
	if (ok && launches) (*launches)++;
	return ok;
}
template bool launch_tma_sweep<float>(int, const SweepArgs<float> &, cudaStream_t, long long *);
template bool launch_tma_sweep<double>(int, const SweepArgs<double> &, cudaStream_t, long long *);

} // namespace cmc
