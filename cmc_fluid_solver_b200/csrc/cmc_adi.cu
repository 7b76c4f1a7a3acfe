// cmc_adi.cu - C ABI (include/cmc_adi.h) and host orchestration of the ADI time step.
//
// Host-side restatement of AdiSolver3D::{Init, CreateSegments, UpdateBoundaries, TimeStep,
// SolveDirection} (reference src/FluidSolver3D/AdiSolver3D.cpp:166-268, 286-391, 553-666) and
// Solver3D::GetLayer (Solver3D.cpp:21-25) on top of the sm_100a kernels.
//
// A handle (Engine) owns one or more x-slabs of the grid (Slab: all device buffers of one slab):
//   * single GPU            : one slab = the whole grid;
//   * one process per GPU   : one slab per handle, neighbours reached through NCCL (dist.h);
//   * emulated slabs        : N slabs of ONE device in one process, exchanging with device-to-device copies -
//                             the same code path as the NCCL case, testable on a single GPU (the counterpart of
//                             the reference's MGPU_EMU switch, src/Common/GPUplan.h:10-15).
// The step logic is written once, in lockstep over the local slabs; every exchange is stream-ordered.  A time
// step does not synchronise with the host unless the caller asks for the residual.  No CPU fallback anywhere.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <new>
#include <thread>
#include <atomic>

#include "../../include/cmc_adi.h"
#include "kernels.h"
#include "dist.h"

using namespace cmc;

static thread_local std::string g_err;

static int fail(int code, const std::string &msg)
{
	g_err = msg;
	return code;
}

#define CU_TRY(call)                                                                                        \
	do {                                                                                                    \
		cudaError_t e__ = (call);                                                                           \
		if (e__ != cudaSuccess) {                                                                           \
			char buf__[512];                                                                                \
			snprintf(buf__, sizeof buf__, "%s failed on device %d: %s (%d) at %s:%d", #call, device,        \
			         cudaGetErrorString(e__), (int)e__, __FILE__, __LINE__);                               \
			return fail(CMC_ERR_CUDA, buf__);                                                               \
		}                                                                                                   \
	} while (0)

struct cmc_adi3d {
	virtual ~cmc_adi3d() {}
	virtual int set_nodes(const int32_t *type, const int32_t *bc_vel, const int32_t *bc_temp,
	                      const void *vx, const void *vy, const void *vz, const void *T, size_t aos_stride, bool keep_layers) = 0;
	virtual int set_nodes_slab(const int32_t *type, const int32_t *bc_vel, const int32_t *bc_temp,
	                           const void *vx, const void *vy, const void *vz, const void *T, int halo, bool keep_layers) = 0;
	virtual int build_lines() = 0;
	virtual int update_boundaries() = 0;
	virtual int time_step(double dt, int ng, int nl, int ce, double *err, bool async) = 0;
	virtual int sync(double *err) = 0;
	virtual int get_layer(void *vel, double *T, int ox, int oy, int oz) = 0;
	virtual int get_layer_async(void *vel, double *T, int ox, int oy, int oz) = 0;
	virtual int get_layer_wait() = 0;
	virtual int write_layer_async(const void *const src[4]) = 0;
	virtual int write_layer_commit(int layer) = 0;
	virtual int read_field(int layer, int var, void *dst) = 0;
	virtual int write_field(int layer, int var, const void *src) = 0;
	virtual int step_prologue() = 0;
	virtual int solve_direction(int dir, double dt, int nl, int cur_layer, int next_layer) = 0;
	virtual int eval_div_error(int layer, double *err) = 0;
	virtual int field_sums(int layer, double *sums8) = 0;
	virtual int exchange_kind() const = 0;
	virtual int kernel_kind(int dir) const = 0;
	virtual int debug_counter(int i, int64_t *value) = 0;

	int device = 0, fp = 8;
	int rank = 0, nranks = 1;       // position of this handle's (first) slab among all slabs of the grid
	Layout L{}, G{};                // L: extent of the locally held planes [x0, x0 + nx); G: the whole grid
	cudaStream_t stream = nullptr;
	long long launches = 0;
	long long dev_bytes = 0;
	long long num_segs[3] = {0, 0, 0};
	long long shared_free[3] = {0, 0, 0};   // cells shared by two segments with a BC_FREE row (informational)
	int mode = CMC_MODE_FAST;
	int tma_mask = default_tma_mask();
	// option "local_output": GetLayer of a run with one process per GPU leaves every rank's output rows on that rank (rows
	// [lo, hi) of the arrays the rank passed, cmc_adi3d_output_rows) instead of gathering everything on rank 0 - N host links
	// carry the result instead of one
	int local_output = 0;
	int tma_shape = 0;               // option "tma_shape": 0 = automatic, else lines per tile + 256 * CTAs per tile (kernels_tma.cu)
	// option "xs" / CMC_XS=1: one-pass slab-coupled x-sweep (kernels_tma.cu XS).  Off by default: measured on 2 and 4 B200s it
	// loses to the two-pass form (4.81 against 4.07 ms and 3.28 against 1.92 ms per x-sweep at 512^3, profiles/r02_variants.md)
	int xs_enabled = getenv("CMC_XS") ? atoi(getenv("CMC_XS")) : 0;
	static int default_tma_mask()
	{
		// default: both strided axes (measured at 512^3 fp64 on B200: x 5.14 -> 4.44 ms, y 5.07 -> 4.36 ms per launch against
		// the direct-load kernel, profiles/r02_variants.md); CMC_TMA=<subset of "xy"> overrides, CMC_TMA= turns it off
		const char *env = getenv("CMC_TMA");
		if (!env) return 3;
		return (strchr(env, 'x') ? 1 : 0) | (strchr(env, 'y') ? 2 : 0);
	}
	bool have_nodes = false, have_lines = false;
	bool window_counts = false;     // the x-segment count was taken per slab (slab-local node arrays)

	// optional per-kernel-kind device timing (cmc_adi3d_set_option "profile"): CUDA event pairs on `stream`
	int profile = 0;
	struct Span { int kind; cudaEvent_t a, b; };
	std::vector<Span> spans;
	std::vector<size_t> open_spans;          // indices of the spans begun and not yet ended (innermost last)
	std::vector<cudaEvent_t> span_events;    // events of collected spans, reused: nothing is created inside a timed loop after warm-up
	cudaEvent_t span_event()
	{
		cudaEvent_t e = nullptr;
		if (!span_events.empty()) { e = span_events.back(); span_events.pop_back(); }
		else cudaEventCreate(&e);
		return e;
	}
	double kind_ms[CMC_TIMING_KINDS] = {};
	long long kind_calls[CMC_TIMING_KINDS] = {};
	void span_begin(int kind)
	{
		if (!profile) return;
		Span sp; sp.kind = kind;
		cudaSetDevice(device);                  // (a handle with slabs on several devices may have left another one current)
		sp.a = span_event(); sp.b = span_event();
		cudaEventRecord(sp.a, stream);          // (slabs on several devices: the first slab's stream)
		open_spans.push_back(spans.size());     // (spans nest: the residual span contains a halo exchange on the NCCL transport)
		spans.push_back(sp);
	}
	void span_end()
	{
		if (!profile || open_spans.empty()) return;
		cudaSetDevice(device);
		cudaEventRecord(spans[open_spans.back()].b, stream);
		open_spans.pop_back();
	}
	void spans_collect()
	{
		open_spans.clear();
		if (spans.empty()) return;
		cudaStreamSynchronize(stream);
		for (auto &sp : spans) {
			float ms = 0.f;
			if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { kind_ms[sp.kind] += ms; kind_calls[sp.kind]++; }
			else cudaGetLastError();            // (a span that an error return left open: not an error of a later call)
			span_events.push_back(sp.a); span_events.push_back(sp.b);
		}
		spans.clear();
	}
};

namespace {

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

// host loops over all cells of the grid (node packing): chunks of the index range on all hardware threads
template <typename F>
static void parallel_for(size_t n, F body)
{
	unsigned nt = std::thread::hardware_concurrency();
	if (nt > 32) nt = 32;
	if (nt < 2 || n < (1u << 20)) { body(0, n); return; }
	std::vector<std::thread> pool;
	const size_t chunk = (n + nt - 1) / nt;
	for (unsigned t = 0; t < nt; t++) {
		const size_t a = t * chunk, b = std::min(n, a + chunk);
		if (a >= b) break;
		pool.emplace_back([=] { body(a, b); });
	}
	for (auto &th : pool) th.join();
}

// Default x-split of a grid over n slabs: the even split of GPUplan::splitEven1D (reference GPUplan.cpp:122-141) with the
// cut positions moved to multiples of 8 planes - the partitioned x-sweep works on 8-row chunks, so every slab but the
// last holds a multiple of 8 planes; the last one takes what is left.  (n == 1: the whole grid.)
static void split_default(int dimx, int n, std::vector<int> &planes)
{
	planes.assign(n, 0);
	int prev = 0;
	for (int r = 0; r < n; r++) {
		int cut = r + 1 == n ? dimx : (int)(((long long)dimx * (r + 1) / n + 4) / 8 * 8);
		if (cut < prev + 8) cut = prev + 8;
		if (cut > dimx) cut = dimx;
		planes[r] = cut - prev;
		prev = cut;
	}
	if (n == 1) planes[0] = dimx;
}

// ---- one x-slab: every device buffer of the planes [x0, x0 + nx) ------------------------------------------------
template <typename FT>
struct Slab {
	int device = 0;
	cudaStream_t stream = nullptr;
	Layout L{}, G{};
	int index = 0;                 // position among all slabs of the grid
	FT *field[5][4] = {};          // physical buffers: four layers + the spare linearisation buffer
	int slot[4] = {0, 1, 2, 3};    // logical layer (CMC_LAYER_*) -> physical buffer
	int spare = 4;
	FT *nodev[4] = {};
	uint8_t *role[3] = {};
	uint8_t *ncode = nullptr;      // node codes (only until the line descriptors are built): the whole grid, or a window of it
	int ncode_w0 = 0;              // global plane held by plane 0 of ncode
	bool ncode_window = false;     // ncode covers [x0 - 2, x0 + nx + 2) (clipped to the grid) instead of the whole grid
	FT *cv = nullptr, *cT = nullptr;
	double *d_partials = nullptr, *d_err2 = nullptr, *d_sums8 = nullptr;
	unsigned long long *d_segcount = nullptr;
	int *d_tilectr = nullptr;      // tile counter of the persistent sweep kernels
	cudaEvent_t done = nullptr;    // slabs on different devices of ONE process: "everything enqueued so far" of this slab
	bool owns_stream = false;
	// GetLayer staging (device): two sets, so that the device-to-host copy of one readback can still be running - on the
	// slab's copy stream - while the next one is produced
	FT *d_outvel2[2] = {nullptr, nullptr};
	double *d_outT2[2] = {nullptr, nullptr};
	size_t out_cap2[2] = {0, 0};
	FT *d_outvel = nullptr;        // the set in use by the current readback
	double *d_outT = nullptr;
	int out_idx = 0;
	cudaStream_t io_out = nullptr, io_in = nullptr;      // copy streams: results to the host / inputs from the host
	cudaEvent_t ev_filtered = nullptr, ev_out_done[2] = {nullptr, nullptr}, ev_in_done = nullptr, ev_in_free = nullptr;
	bool out_pending[2] = {false, false};
	FT *d_stage_in[4] = {nullptr, nullptr, nullptr, nullptr};   // dense [nx][ny][nz] staging of an asynchronously uploaded layer
	bool in_pending = false;
	// partitioned x-sweep exchange buffers: [peer][16 | 8][lpo].  The *_recv tables are filled by the other slabs'
	// kernels directly (they live in the arena); the *_send staging buffers exist only for the NCCL transport.
	FT *xcoef_send = nullptr, *xcoef_recv = nullptr, *xbnd_send = nullptr, *xbnd_recv = nullptr;
	// fused one-pass x-sweep (kernels_tma.cu XS): [slab][tile][16][lines per tile] coefficients of every slab's first / last
	// row and [slab][tile][4][lines per tile] flag words, written by all slabs' kernels (arena)
	unsigned long long *xs_tab = nullptr;
	// exchange arena: everything another slab stores into - the 20 field buffers (guard planes), the two interface
	// tables and the flag words - in ONE allocation with the same layout on every rank, so that one CUDA IPC mapping
	// per peer makes all of it addressable over NVLink
	char *arena = nullptr;
	size_t arena_bytes = 0, flag_off = 0;
	long long bytes = 0;
	static const int kMaxErrBlocks = 148 * 8;

	~Slab()
	{
		cudaSetDevice(device);
		if (arena) cudaFree(arena);
		for (auto &p : nodev) if (p) cudaFree(p);
		for (auto &p : role) if (p) cudaFree(p);
		if (done) cudaEventDestroy(done);
		if (owns_stream && stream) cudaStreamDestroy(stream);
		if (io_out) { cudaStreamSynchronize(io_out); cudaStreamDestroy(io_out); }
		if (io_in) { cudaStreamSynchronize(io_in); cudaStreamDestroy(io_in); }
		cudaEvent_t evs[] = {ev_filtered, ev_out_done[0], ev_out_done[1], ev_in_done, ev_in_free};
		for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
		void *misc[] = {cv, cT, d_partials, d_err2, d_sums8, d_segcount, d_tilectr, d_outvel2[0], d_outT2[0], d_outvel2[1], d_outT2[1],
		                d_stage_in[0], d_stage_in[1], d_stage_in[2], d_stage_in[3], ncode, xcoef_send, xbnd_send};
		for (void *p : misc) if (p) cudaFree(p);
	}

	template <typename T>
	int dalloc(T *&p, size_t count)
	{
		CU_TRY(cudaMalloc((void **)&p, count * sizeof(T)));
		CU_TRY(cudaMemsetAsync(p, 0, count * sizeof(T), stream));
		bytes += (long long)(count * sizeof(T));
		return CMC_OK;
	}

	int init(const Layout &g, int x0, int nx, int alloc_nx, int dev, cudaStream_t s, int idx, int nslabs, bool with_xs = false)
	{
		device = dev; stream = s; G = g; index = idx;
		L = G; L.x0 = x0; L.shape(nx, G.ny, G.nz, G.nzp, G.jbs, alloc_nx);
		int rc;
		{
			auto up = [](size_t v) { return (v + 255) / 256 * 256; };
			const size_t fb = up(sizeof(FT) * (size_t)L.total), lpo = lines_per_owner(nslabs);
			const size_t cb = nslabs > 1 ? up(sizeof(FT) * lpo * 16 * nslabs) : 0, bb = nslabs > 1 ? up(sizeof(FT) * lpo * 8 * nslabs) : 0;
			// (lines padded to whole 16-line tiles)
			const size_t xlines = (size_t)G.ny * (size_t)((G.nz + 15) / 16 * 16);
			// (only when the one-pass slab-coupled x-sweep is switched on at creation: 0.5 GB per slab at 512^2 lines and 8 slabs)
			const size_t xtb = (nslabs > 1 && with_xs) ? up(sizeof(unsigned long long) * xlines * 16 * (sizeof(FT) / 4) * nslabs) : 0;
			flag_off = 20 * fb + cb + bb + xtb;
			arena_bytes = flag_off + 256;
			CU_TRY(cudaMalloc((void **)&arena, arena_bytes));
			CU_TRY(cudaMemsetAsync(arena, 0, arena_bytes, stream));
			bytes += (long long)arena_bytes;
			for (int l = 0; l < 5; l++)
				for (int q = 0; q < 4; q++) field[l][q] = reinterpret_cast<FT *>(arena + (size_t)(l * 4 + q) * fb);
			if (nslabs > 1) {
				xcoef_recv = reinterpret_cast<FT *>(arena + 20 * fb);
				xbnd_recv = reinterpret_cast<FT *>(arena + 20 * fb + cb);
				if (xtb) xs_tab = reinterpret_cast<unsigned long long *>(arena + 20 * fb + cb + bb);
			}
		}
		for (int q = 0; q < 4; q++) if ((rc = dalloc(nodev[q], (size_t)L.total))) return rc;
		for (int d = 0; d < 3; d++) if ((rc = dalloc(role[d], (size_t)L.total))) return rc;
		if ((rc = dalloc(cv, (size_t)L.total))) return rc;
		if ((rc = dalloc(cT, (size_t)L.total))) return rc;
		if ((rc = dalloc(d_partials, (size_t)2 * kMaxErrBlocks))) return rc;
		if ((rc = dalloc(d_err2, 2))) return rc;
		if ((rc = dalloc(d_sums8, 8))) return rc;
		if ((rc = dalloc(d_segcount, 8))) return rc;
		if ((rc = dalloc(d_tilectr, 32))) return rc;
		if (nslabs > 1) {
			const size_t lpo = lines_per_owner(nslabs);
			if ((rc = dalloc(xcoef_send, lpo * 16 * nslabs))) return rc;
			if ((rc = dalloc(xbnd_send, lpo * 8 * nslabs))) return rc;
		}
		return CMC_OK;
	}

	int io_setup()
	{
		if (io_out) return CMC_OK;
		CU_TRY(cudaStreamCreateWithFlags(&io_out, cudaStreamNonBlocking));
		CU_TRY(cudaStreamCreateWithFlags(&io_in, cudaStreamNonBlocking));
		cudaEvent_t *evs[] = {&ev_filtered, &ev_out_done[0], &ev_out_done[1], &ev_in_done, &ev_in_free};
		for (cudaEvent_t *e : evs) CU_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
		return CMC_OK;
	}

	size_t lines_per_owner(int nslabs) const { return ((size_t)G.ny * G.nz + nslabs - 1) / nslabs; }

	// dense host array of the WHOLE grid ((i * ny + j) * nz + k) <-> the local planes [p0, p1) of one field buffer;
	// host_x0 = global plane held by host plane 0 of `host`.  One 3-D copy per y-block.
	int copy_planes(FT *dev, const FT *host_c, FT *host_m, int p0, int p1, int host_x0)
	{
		for (int b = 0; b < L.nblk; b++) {
			const int j0 = b << (L.nblk == 1 ? 0 : L.jbs), rows = L.nblk == 1 ? L.ny : std::min(1 << L.jbs, L.ny - j0);
			cudaMemcpy3DParms P;
			memset(&P, 0, sizeof P);
			const size_t hoff = ((size_t)(L.x0 + p0 - host_x0) * L.ny + j0) * L.nz;
			cudaPitchedPtr hp = make_cudaPitchedPtr((void *)((host_c ? host_c : host_m) + hoff), sizeof(FT) * L.nz, L.nz, L.ny);
			cudaPitchedPtr dp = make_cudaPitchedPtr((void *)(dev + L.idx(p0, j0, 0)), sizeof(FT) * L.nzp, L.nzp, (size_t)(L.plane / L.nzp));
			if (host_c) { P.srcPtr = hp; P.dstPtr = dp; P.kind = cudaMemcpyHostToDevice; }
			else { P.srcPtr = dp; P.dstPtr = hp; P.kind = cudaMemcpyDeviceToHost; }
			P.extent = make_cudaExtent(sizeof(FT) * L.nz, (size_t)rows, (size_t)(p1 - p0));
			CU_TRY(cudaMemcpy3DAsync(&P, stream));
		}
		return CMC_OK;
	}

	ConstLayerPtrs<FT> clayer(int logical) const
	{
		ConstLayerPtrs<FT> r;
		for (int q = 0; q < 4; q++) r.f[q] = field[slot[logical]][q];
		return r;
	}
	LayerPtrs<FT> layer(int logical)
	{
		LayerPtrs<FT> r;
		for (int q = 0; q < 4; q++) r.f[q] = field[slot[logical]][q];
		return r;
	}

	// dense host (whole grid) -> padded device slab, plus the neighbour planes into the guard (halo) planes
	// w0 = global plane of host plane 0 (0: the arrays cover the whole grid)
	int upload_nodes(const uint8_t *code, size_t N, const FT *const src[4], bool keep_layers, int w0 = 0, bool window = false)
	{
		if (ncode) { cudaFree(ncode); ncode = nullptr; }
		CU_TRY(cudaMalloc((void **)&ncode, N));
		CU_TRY(cudaMemcpyAsync(ncode, code, N, cudaMemcpyHostToDevice, stream));
		ncode_w0 = w0; ncode_window = window;
		for (int q = 0; q < 4; q++) {
			CU_TRY(cudaMemsetAsync(nodev[q], 0, sizeof(FT) * (size_t)L.total, stream));
			const int p0 = L.x0 > 0 ? -1 : 0, p1 = L.x0 + L.nx < G.nx ? L.nx + 1 : L.nx;   // include the halo planes that exist
			int rc = copy_planes(nodev[q], src[q], nullptr, p0, p1, w0);
			if (rc) return rc;
		}
		if (keep_layers) return CMC_OK;        // Grid3D::Prepare(t): new nodes, same time layers
		// cur = TimeLayer3D(grid) (TimeLayer3D.h:734-751); half/next/temp are uninitialised in the reference
		// (TimeLayer3D.h:353) and are defined here as copies of cur.
		for (int l = 0; l < 5; l++)
			for (int q = 0; q < 4; q++)
				CU_TRY(cudaMemcpyAsync(field[l][q], nodev[q], sizeof(FT) * (size_t)L.total, cudaMemcpyDeviceToDevice, stream));
		slot[0] = 0; slot[1] = 1; slot[2] = 2; slot[3] = 3; spare = 4;
		return CMC_OK;
	}
};

// ---- the handle ----------------------------------------------------------------------------------------------------
template <typename FT>
struct Engine : cmc_adi3d {
	std::vector<Slab<FT> *> slabs;     // local slabs: 1 (single GPU / NCCL rank) or all of them (emulation)
	NcclComm *nccl = nullptr;          // one process per GPU
	int nslabs_total = 1;
	cmc_fluid_params params{};
	double dx = 0, dy = 0, dz = 0;
	double diffError = 0.0;
	// residuals of enqueued (asynchronous) steps: pinned ring of kErrRing entries of [2 * nlocal] (sum, count) per local
	// slab.  sync() looks at EVERY pending entry, so a step that diverged in the middle of an asynchronous run is not
	// hidden by the steps that followed it (the reference checks after every step, AdiSolver3D.cpp:371-374)
	static constexpr int kErrRing = 64;
	int err_head = 0, err_pending = 0; // next entry to write / entries written since the last fetch
	double worst_err = 0.0;            // largest residual among the entries of the last fetch
	double *h_err2 = nullptr;
	// exchanges as stores into the other slabs' buffers (see SweepArgs::push_*): always when all slabs share this
	// device (emulation), and between processes once every rank has mapped every other rank's arena (peer memory)
	PeerMap pm;
	bool p2p = false;
	unsigned epoch = 0;                // ordering of peer stores: last epoch this rank has published
	int *d_timeout = nullptr;
	bool halos_dirty = true;           // the guard planes of `cur` may not match the neighbours' boundary planes
	// all slabs in ONE process on DIFFERENT devices (cmc_adi3d_create_multi: the reference's "GPU <n>" mode, one host
	// thread driving n devices, FluidSolver3D.cpp:88-95, GPUplan.cpp:35-77): every slab has its own stream on its own
	// device, the sweeps store into the neighbours' buffers through peer access, and ordering between the slabs' streams
	// is by events - publish() records one per slab, await() makes the slabs' streams wait for their neighbours' (or for
	// everybody's)
	bool multi_device = false;
	std::vector<int> slab_devices;
	// the x-split of the grid over ALL slabs (local or not): planes per slab and first plane of every slab
	std::vector<int> split_nx, split_x0;
	std::vector<int> planes_arg;       // caller-supplied split (cmc_adi3d_create_*_split), empty = the default

	void use(const Slab<FT> *s) const { if (multi_device) cudaSetDevice(s->device); }
	int sync_all()
	{
		for (auto *s : slabs) {
			use(s);
			const cudaError_t e = cudaStreamSynchronize(s->stream);
			if (e != cudaSuccess) return fail(CMC_ERR_CUDA, std::string("cudaStreamSynchronize failed on device ") + std::to_string(s->device) + ": " + cudaGetErrorString(e));
		}
		if (multi_device) cudaSetDevice(device);
		return CMC_OK;
	}
	// every slab's stream waits for what all slabs have enqueued so far (exchanges that run outside the sweeps)
	void join_streams()
	{
		if (!multi_device) return;
		for (auto *s : slabs) { use(s); cudaEventRecord(s->done, s->stream); }
		for (auto *s : slabs) { use(s); for (auto *r : slabs) if (r != s) cudaStreamWaitEvent(s->stream, r->done, 0); }
	}

	~Engine() override
	{
		if (multi_device) sync_all();
		cudaSetDevice(device);
		if (p2p && stream) {
			// the neighbours' last sweeps may still be storing into this rank's arena: every rank publishes once more
			// when it gets here and waits for the others before anything is unmapped or freed (10 s timeout inside)
			publish();
			await(all_mask());
		}
		if (stream) cudaStreamSynchronize(stream);
		spans_collect();
		for (cudaEvent_t e : span_events) cudaEventDestroy(e);
		span_events.clear();
		if (p2p) peer_unmap(&pm);
		if (d_timeout) cudaFree(d_timeout);
		for (auto *s : slabs) delete s;
		if (h_err2) cudaFreeHost(h_err2);
		if (nccl) nccl_destroy(nccl);
		if (stream) cudaStreamDestroy(stream);
	}

	bool multi() const { return nslabs_total > 1; }
	int exchange_kind() const override { return !multi() ? 0 : !push_mode() ? 1 : nccl ? 3 : multi_device ? 4 : 2; }
	// (stores into another slab's buffers assume its layout equals this slab's: equal numbers of planes)
	bool push_mode() const { return multi() && (!nccl || p2p); }

	// the address, in the slab that holds slab index `r`, of the buffer that is `mine` in slab `s`
	template <typename T>
	T *in_slab(Slab<FT> *s, int r, T *mine) const
	{
		const size_t o = (size_t)((char *)mine - s->arena);
		if (nccl) return reinterpret_cast<T *>((char *)pm.base[r] + o);
		return reinterpret_cast<T *>(slabs[r]->arena + o);
	}
	// One-pass slab-coupled x-sweep (kernels_tma.cu XS) instead of spike pass + interface kernel + coupled pass: needs the
	// slabs' kernels to run at the same time (one process per GPU with mapped peer memory, or one process driving several
	// devices - not the emulation of several slabs on one stream), whole 8-row chunks in every slab and the same tile
	// shape for all slabs.  Every rank evaluates this from the same global split, so all take the same path.
	// CMC_XS=1 at creation switches it on (see xs_enabled; option "xs" can switch it off and on again afterwards).
	int xs_epoch = 0;
	bool fused_x() const
	{
		if (!multi() || !xs_enabled || mode != CMC_MODE_FAST || !want_tma(CMC_DIR_X)) return false;
		if (!slabs[0]->xs_tab) return false;                 // (the tables exist only when the option was on at creation: CMC_XS=1)
		if (!((nccl && p2p) || multi_device)) return false;
		int nl = -1;
		for (int r = 0; r < nslabs_total; r++) {
			Layout Lr = G; Lr.shape(split_nx[r], G.ny, G.nz, G.nzp, G.jbs, slabs[0]->L.bstride / slabs[0]->L.plane - 2);
			const int v = tma_xs_lines(Lr);
			if (!v || (nl >= 0 && v != nl)) return false;
			nl = v;
		}
		return true;
	}

	// peer ordering (processes): publish `epoch` after this rank's kernel / wait for the ranks in `mask`
	void publish()
	{
		if (multi_device) { for (auto *s : slabs) { use(s); cudaEventRecord(s->done, s->stream); } return; }
		if (!p2p) return;
		epoch++;
		peer_signal(pm, slabs[0]->flag_off, epoch, stream);
		launches++;
	}
	void await(unsigned mask)
	{
		if (multi_device) {
			// (mask is built for `rank`, the first slab: neighbour_mask() / all_mask(); here it only says which of the two)
			const bool all = mask == all_mask();
			for (size_t i = 0; i < slabs.size(); i++) {
				use(slabs[i]);
				for (size_t r = 0; r < slabs.size(); r++) {
					if (r == i) continue;
					if (all || r + 1 == i || r == i + 1) cudaStreamWaitEvent(slabs[i]->stream, slabs[r]->done, 0);
				}
			}
			return;
		}
		if (!p2p) return;
		peer_wait(pm, slabs[0]->flag_off, mask, epoch, d_timeout, stream);
		launches++;
	}
	unsigned neighbour_mask() const { return (rank > 0 ? 1u << (rank - 1) : 0u) | (rank + 1 < nranks ? 1u << (rank + 1) : 0u); }
	unsigned all_mask() const { return (nranks >= 32 ? ~0u : (1u << nranks) - 1u) & ~(1u << rank); }

	int init(const cmc_grid_desc *g, const cmc_fluid_params *p, int first_slab, int nlocal, int ntotal)
	{
		CU_TRY(cudaSetDevice(device));
		CU_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
		params = *p;
		dx = g->dx; dy = g->dy; dz = g->dz;
		G.nx = g->dimx; G.ny = g->dimy; G.nz = g->dimz; G.gx = g->dimx; G.x0 = 0;
		{
			// y-blocking of the storage (Layout).  Default: blocks of about 256 KB per x-plane once a whole plane is
			// larger than 512 KB (measured on B200, profiles/r01_variants.md: the x-sweep pays per touched page, 512^3 fp64
			// 7.0 -> 5.2 ms per launch with 64-row blocks).  CMC_JB = rows per block (a power of two >= 8), 0 = one block.
			int jbs = 30;
			const int nzp = round_up(g->dimz, 16);
			int jb = 0;
			if ((size_t)g->dimy * nzp * sizeof(FT) > (512u << 10)) {
				jb = 8;
				while ((size_t)(2 * jb) * nzp * sizeof(FT) <= (256u << 10)) jb *= 2;
			}
			if (getenv("CMC_JB")) jb = atoi(getenv("CMC_JB"));
			if (jb >= 8 && (jb & (jb - 1)) == 0) { jbs = 0; while ((1 << jbs) < jb) jbs++; }
			G.shape(g->dimx, g->dimy, g->dimz, round_up(g->dimz, 16), jbs);
		}
		nslabs_total = ntotal; rank = first_slab; nranks = ntotal;
		if ((int)planes_arg.size() == ntotal) split_nx = planes_arg; else split_default(G.nx, ntotal, split_nx);
		split_x0.assign(ntotal, 0);
		int nx_max = 0, covered = 0;
		for (int r = 0; r < ntotal; r++) {
			split_x0[r] = covered; covered += split_nx[r];
			nx_max = std::max(nx_max, split_nx[r]);
			if (split_nx[r] < 1) return fail(CMC_ERR_INVALID, "create: a slab without planes (grid too small for this many slabs, or a bad split)");
		}
		if (covered != G.nx) return fail(CMC_ERR_INVALID, "create: the slab sizes do not add up to dimx");
		int lo = 0, hi = 0;
		for (int i = 0; i < nlocal; i++) {
			const int x0 = split_x0[first_slab + i], nx = split_nx[first_slab + i];
			if (i == 0) lo = x0;
			hi = x0 + nx;
			auto *s = new (std::nothrow) Slab<FT>();
			if (!s) return fail(CMC_ERR_INVALID, "out of host memory");
			slabs.push_back(s);
			int sdev = device;
			cudaStream_t sstream = stream;
			if (multi_device && i > 0) {              // (slab 0 lives on the handle's own device and stream: the timing spans see it)
				sdev = slab_devices[i];
				{ const int device = sdev; CU_TRY(cudaSetDevice(sdev)); CU_TRY(cudaStreamCreateWithFlags(&sstream, cudaStreamNonBlocking)); }
				s->owns_stream = true;
			}
			int rc = s->init(G, x0, nx, ntotal > 1 ? nx_max : nx, sdev, sstream, first_slab + i, ntotal, xs_enabled != 0);
			if (rc) return rc;
			if (multi_device) { const int device = sdev; CU_TRY(cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming)); }
			dev_bytes += s->bytes;
		}
		if (multi_device) {
			// every device stores into every other device's slab buffers (guard planes, interface tables)
			for (size_t i = 0; i < slabs.size(); i++) {
				cudaSetDevice(slabs[i]->device);
				for (size_t j = 0; j < slabs.size(); j++) {
					if (i == j || slabs[i]->device == slabs[j]->device) continue;
					int can = 0;
					cudaDeviceCanAccessPeer(&can, slabs[i]->device, slabs[j]->device);
					if (!can) return fail(CMC_ERR_UNSUPPORTED, "create_multi: the devices cannot access each other's memory (peer access)");
					const cudaError_t e = cudaDeviceEnablePeerAccess(slabs[j]->device, 0);
					if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
						return fail(CMC_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
					cudaGetLastError();
				}
			}
			cudaSetDevice(device);
		}
		L = G; L.x0 = lo; L.shape(hi - lo, G.ny, G.nz, G.nzp, G.jbs);
		CU_TRY(cudaHostAlloc((void **)&h_err2, 2 * sizeof(double) * nlocal * kErrRing, cudaHostAllocPortable));
		memset(h_err2, 0, 2 * sizeof(double) * nlocal * kErrRing);
		CU_TRY(cudaMalloc((void **)&d_timeout, sizeof(int)));
		CU_TRY(cudaMemsetAsync(d_timeout, 0, sizeof(int), stream));
		CU_TRY(cudaStreamSynchronize(stream));
		if (nccl) {
			// (every rank's arena has the layout of the largest slab, so one mapping per peer addresses everything);
			// CMC_P2P=0 keeps the NCCL transport
			const bool want = !(getenv("CMC_P2P") && atoi(getenv("CMC_P2P")) == 0);
			if (want) p2p = peer_map_arenas(nccl, slabs[0]->arena, slabs[0]->arena_bytes, rank, nranks, &pm, stream) == 0;
		}
		return CMC_OK;
	}

	// ------------------------------------------------------------------------------------------- exchanges
	// boundary x-planes of the 4 fields of one logical layer -> the neighbours' guard planes
	int halo_exchange(int logical)
	{
		if (!multi()) return CMC_OK;
		span_begin(CMC_TIMING_COMM);
		// an x-plane is one contiguous piece per y-block
		Slab<FT> *s0 = slabs[0];
		const int nb = s0->L.nblk;
		const size_t pb = sizeof(FT) * (size_t)s0->L.plane;
		if (nccl) {
			Slab<FT> *s = s0;
			std::vector<P2P> ops;
			for (int q = 0; q < 4; q++) {
				FT *f = s->field[s->slot[logical]][q];
				for (int b = 0; b < nb; b++) {
					const long long bo = (long long)b * s->L.bstride;
					if (rank > 0) ops.push_back(P2P{f + bo + s->L.idx(0, 0, 0), f + bo + s->L.idx(-1, 0, 0), pb, rank - 1});
					if (rank + 1 < nranks) ops.push_back(P2P{f + bo + s->L.idx(s->L.nx - 1, 0, 0), f + bo + s->L.idx(s->L.nx, 0, 0), pb, rank + 1});
				}
			}
			if (nccl_exchange(nccl, ops.data(), (int)ops.size(), stream)) return fail(CMC_ERR_COMM, nccl_error());
			launches += 1;
		} else {
			join_streams();          // (slabs on different devices: the copies below read what the other streams produced)
			for (size_t i = 0; i + 1 < slabs.size(); i++) {
				Slab<FT> *a = slabs[i], *b = slabs[i + 1];
				use(a);
				for (int q = 0; q < 4; q++) {
					FT *fa = a->field[a->slot[logical]][q], *fb = b->field[b->slot[logical]][q];
					for (int k = 0; k < nb; k++) {
						const long long ao = (long long)k * a->L.bstride, bo = (long long)k * b->L.bstride;
						CU_TRY(cudaMemcpyAsync(fb + bo + b->L.idx(-1, 0, 0), fa + ao + a->L.idx(a->L.nx - 1, 0, 0), pb, cudaMemcpyDefault, a->stream));
						CU_TRY(cudaMemcpyAsync(fa + ao + a->L.idx(a->L.nx, 0, 0), fb + bo + b->L.idx(0, 0, 0), pb, cudaMemcpyDefault, a->stream));
					}
				}
			}
			join_streams();
		}
		span_end();
		return CMC_OK;
	}

	// all-to-all of per-peer blocks: block p of `send` goes to slab p, which stores it as block <sender> of `recv`
	int all_to_all(FT *Slab<FT>::*send, FT *Slab<FT>::*recv, size_t block_elems)
	{
		span_begin(CMC_TIMING_COMM);
		const size_t bb = sizeof(FT) * block_elems;
		if (nccl) {
			Slab<FT> *s = slabs[0];
			std::vector<P2P> ops;
			for (int p = 0; p < nranks; p++) {
				if (p == rank) { CU_TRY(cudaMemcpyAsync((s->*recv) + p * block_elems, (s->*send) + p * block_elems, bb, cudaMemcpyDeviceToDevice, stream)); continue; }
				ops.push_back(P2P{(s->*send) + p * block_elems, (s->*recv) + p * block_elems, bb, p});
			}
			if (nccl_exchange(nccl, ops.data(), (int)ops.size(), stream)) return fail(CMC_ERR_COMM, nccl_error());
			launches += 1;
		} else {
			join_streams();
			for (size_t i = 0; i < slabs.size(); i++) {
				use(slabs[i]);
				for (size_t j = 0; j < slabs.size(); j++)
					CU_TRY(cudaMemcpyAsync((slabs[j]->*recv) + i * block_elems, (slabs[i]->*send) + j * block_elems, bb, cudaMemcpyDefault, slabs[i]->stream));
			}
			join_streams();
		}
		span_end();
		return CMC_OK;
	}

	// ------------------------------------------------------------------------------------------- set up
	int set_nodes(const int32_t *type, const int32_t *bc_vel, const int32_t *bc_temp,
	              const void *vx, const void *vy, const void *vz, const void *T, size_t aos_stride, bool keep_layers) override
	{
		if (keep_layers && !have_nodes) return fail(CMC_ERR_INVALID, "update_nodes: call cmc_adi3d_set_nodes first");
		CU_TRY(cudaSetDevice(device));
		const size_t N = (size_t)G.nx * G.ny * G.nz;
		std::vector<uint8_t> code(N);
		std::vector<FT> tmp[4];
		const FT *src[4] = {(const FT *)vx, (const FT *)vy, (const FT *)vz, (const FT *)T};
		if (aos_stride) {
			// reference Node (Grid3D.h:73-88): {int type; int bc_vel; int bc_temp; FTYPE v[3]; FTYPE T}
			const char *base = (const char *)type;
			for (int q = 0; q < 4; q++) tmp[q].resize(N);
			FT *t0 = tmp[0].data(), *t1 = tmp[1].data(), *t2 = tmp[2].data(), *t3 = tmp[3].data();
			uint8_t *cd = code.data();
			std::atomic<int> bad(0);
			parallel_for(N, [&](size_t a, size_t b) {
				for (size_t id = a; id < b; id++) {
					const int32_t *hd = (const int32_t *)(base + id * aos_stride);
					const FT *fv = (const FT *)(base + id * aos_stride + (sizeof(FT) == 8 ? 16 : 12));
					if (hd[0] < 0 || hd[0] > 3) bad = 1;
					if ((hd[1] != CMC_BC_NOSLIP && hd[1] != CMC_BC_FREE) || (hd[2] != CMC_BC_NOSLIP && hd[2] != CMC_BC_FREE)) bad = 2;
					cd[id] = (uint8_t)((hd[0] & 3) | (hd[1] == CMC_BC_FREE ? 4 : 0) | (hd[2] == CMC_BC_FREE ? 8 : 0));
					t0[id] = fv[0]; t1[id] = fv[1]; t2[id] = fv[2]; t3[id] = fv[3];
				}
			});
			if (bad == 1) return fail(CMC_ERR_INVALID, "set_nodes_aos: node type out of range");
			if (bad == 2) return fail(CMC_ERR_INVALID, "set_nodes_aos: boundary-condition type out of range");
			for (int q = 0; q < 4; q++) src[q] = tmp[q].data();
		} else {
			uint8_t *cd = code.data();
			std::atomic<int> bad(0);
			parallel_for(N, [&](size_t a, size_t b) {
				for (size_t id = a; id < b; id++) {
					if (type[id] < 0 || type[id] > 3) bad = 1;
					if ((bc_vel[id] != CMC_BC_NOSLIP && bc_vel[id] != CMC_BC_FREE) || (bc_temp[id] != CMC_BC_NOSLIP && bc_temp[id] != CMC_BC_FREE)) bad = 2;
					cd[id] = (uint8_t)((type[id] & 3) | (bc_vel[id] == CMC_BC_FREE ? 4 : 0) | (bc_temp[id] == CMC_BC_FREE ? 8 : 0));
				}
			});
			if (bad == 1) return fail(CMC_ERR_INVALID, "set_nodes: node type out of range");
			if (bad == 2) return fail(CMC_ERR_INVALID, "set_nodes: boundary-condition type out of range");
		}
		for (auto *s : slabs) {
			use(s);
			int rc = s->upload_nodes(code.data(), N, src, keep_layers);
			if (rc) return rc;
		}
		{ int rc = sync_all(); if (rc) return rc; }
		halos_dirty = true;
		have_nodes = true; have_lines = false;
		if (!keep_layers) { diffError = 0.0; err_pending = 0; worst_err = 0.0; }
		return CMC_OK;
	}

	// cmc_adi3d_set_nodes_slab: the arrays cover the planes [x0 - hlo, x0 + nx + hhi) of the grid only, hlo = min(halo, x0),
	// hhi = min(halo, dimx - x0 - nx), halo >= 2 (the line descriptors of a cell look two cells ahead, the guard planes
	// need one) - what Grid3D::Init_GPU uploads per device (node slices, Grid3D.cpp:567-596)
	int set_nodes_slab(const int32_t *type, const int32_t *bc_vel, const int32_t *bc_temp,
	                   const void *vx, const void *vy, const void *vz, const void *T, int halo, bool keep_layers) override
	{
		if (slabs.size() != 1) return fail(CMC_ERR_INVALID, "set_nodes_slab: the handle holds several slabs (it takes whole-grid arrays: cmc_adi3d_set_nodes)");
		if (halo < 2) return fail(CMC_ERR_INVALID, "set_nodes_slab: halo must be at least 2 planes");
		if (keep_layers && !have_nodes) return fail(CMC_ERR_INVALID, "update_nodes_slab: call cmc_adi3d_set_nodes_slab first");
		CU_TRY(cudaSetDevice(device));
		Slab<FT> *s = slabs[0];
		const int hlo = std::min(halo, s->L.x0), hhi = std::min(halo, G.nx - s->L.x0 - s->L.nx);
		const int w0 = s->L.x0 - hlo, planes = hlo + s->L.nx + hhi;
		const size_t N = (size_t)planes * G.ny * G.nz;
		std::vector<uint8_t> code(N);
		uint8_t *cd = code.data();
		std::atomic<int> bad(0);
		parallel_for(N, [&](size_t a, size_t b) {
			for (size_t id = a; id < b; id++) {
				if (type[id] < 0 || type[id] > 3) bad = 1;
				if ((bc_vel[id] != CMC_BC_NOSLIP && bc_vel[id] != CMC_BC_FREE) || (bc_temp[id] != CMC_BC_NOSLIP && bc_temp[id] != CMC_BC_FREE)) bad = 2;
				cd[id] = (uint8_t)((type[id] & 3) | (bc_vel[id] == CMC_BC_FREE ? 4 : 0) | (bc_temp[id] == CMC_BC_FREE ? 8 : 0));
			}
		});
		if (bad == 1) return fail(CMC_ERR_INVALID, "set_nodes_slab: node type out of range");
		if (bad == 2) return fail(CMC_ERR_INVALID, "set_nodes_slab: boundary-condition type out of range");
		const FT *src[4] = {(const FT *)vx, (const FT *)vy, (const FT *)vz, (const FT *)T};
		int rc = s->upload_nodes(code.data(), N, src, keep_layers, w0, true);
		if (rc) return rc;
		CU_TRY(cudaStreamSynchronize(stream));
		halos_dirty = true;
		have_nodes = true; have_lines = false;
		if (!keep_layers) { diffError = 0.0; err_pending = 0; worst_err = 0.0; }
		return CMC_OK;
	}

	int build_lines() override
	{
		if (!have_nodes) return fail(CMC_ERR_INVALID, "build_lines: call cmc_adi3d_set_nodes first");
		CU_TRY(cudaSetDevice(device));
		for (int d = 0; d < 3; d++) { num_segs[d] = 0; shared_free[d] = 0; }
		for (auto *s : slabs) {
			use(s);
			cudaStream_t stream = s->stream;
			CU_TRY(cudaMemsetAsync(s->d_segcount, 0, 8 * sizeof(unsigned long long), stream));
			for (int d = 0; d < 3; d++) CU_TRY(cudaMemsetAsync(s->role[d], 0, (size_t)s->L.total, stream));
			// the kernels address the node codes by GLOBAL plane: a window is handed over with its base shifted accordingly
			// (only planes inside the window are touched)
			const uint8_t *nc = s->ncode - (long long)s->ncode_w0 * G.ny * G.nz;
			launch_role_type_bits(G, nc, s->L, s->role[0], s->role[1], s->role[2], stream, &launches);
			if (s->ncode_window) {
				// x-lines leave the window: local rule instead of the scan; valid when no fluid cell lies on an x-face of the grid
				if (s->L.x0 == 0) launch_count_in_plane(G, nc, 0, s->d_segcount + 7, stream);
				if (s->L.x0 + s->L.nx == G.nx) launch_count_in_plane(G, nc, G.nx - 1, s->d_segcount + 7, stream);
				launch_build_roles_x_local(G, nc, s->L, s->role[0], s->d_segcount + 0, stream, &launches);
				for (int d = 1; d < 3; d++) launch_build_roles(d, G, nc, s->L, s->role[d], s->d_segcount + d, stream, &launches);
			} else
				for (int d = 0; d < 3; d++) launch_build_roles(d, G, nc, s->L, s->role[d], s->d_segcount + d, stream, &launches);
			unsigned long long h[8];
			CU_TRY(cudaMemcpyAsync(h, s->d_segcount, sizeof h, cudaMemcpyDeviceToHost, stream));
			CU_TRY(cudaStreamSynchronize(stream));
			CU_TRY(cudaGetLastError());
			if (s->ncode_window && h[7])
				return fail(CMC_ERR_UNSUPPORTED, "set_nodes_slab: NODE_IN cells on an x-face of the grid (slab-local node arrays need the NODE_OUT rim the reference's loaders produce; use cmc_adi3d_set_nodes)");
			window_counts = s->ncode_window;
			// x-lines are scanned over the whole grid by every slab (same count everywhere) - or, from a window, counted by
			// the slab that holds their end cell; y/z-lines per slab
			if (s->ncode_window) { num_segs[0] += (long long)h[0]; shared_free[0] += (long long)h[4]; }
			else { num_segs[0] = (long long)h[0]; shared_free[0] = (long long)h[4]; }
			for (int d = 1; d < 3; d++) { num_segs[d] += (long long)h[d]; shared_free[d] += (long long)h[4 + d]; }
			cudaFree(s->ncode); s->ncode = nullptr;          // the descriptors carry everything from here on
		}
		if (nccl) {      // y / z counts of the other ranks (and the x counts when every rank counted its own window)
			double *tmp = slabs[0]->d_sums8;
			double v[6] = {(double)num_segs[1], (double)num_segs[2], (double)shared_free[1], (double)shared_free[2],
			               window_counts ? (double)num_segs[0] : 0.0, window_counts ? (double)shared_free[0] : 0.0};
			CU_TRY(cudaMemcpyAsync(tmp, v, sizeof v, cudaMemcpyHostToDevice, stream));
			if (nccl_allreduce_sum_f64(nccl, tmp, 6, stream)) return fail(CMC_ERR_COMM, nccl_error());
			CU_TRY(cudaMemcpyAsync(v, tmp, sizeof v, cudaMemcpyDeviceToHost, stream));
			CU_TRY(cudaStreamSynchronize(stream));
			num_segs[1] = (long long)v[0]; num_segs[2] = (long long)v[1];
			shared_free[1] = (long long)v[2]; shared_free[2] = (long long)v[3];
			if (window_counts) { num_segs[0] = (long long)v[4]; shared_free[0] = (long long)v[5]; }
		}
		if (multi()) {
			for (int r = 0; r < nslabs_total; r++)
				if (split_nx[r] > 512 || split_nx[r] < 8 || (split_nx[r] % 8 != 0 && r + 1 < nslabs_total))
					return fail(CMC_ERR_UNSUPPORTED, "slab-decomposed runs need 8 <= planes per slab <= 512, and a multiple of 8 in every slab but the last");
			if (!fast_sweep_supported(slabs[0]->L, 1) || !fast_sweep_supported(slabs[0]->L, 2))
				return fail(CMC_ERR_UNSUPPORTED, "slab-decomposed runs need 8 <= dimy, dimz <= 512");
		}
		have_lines = true;
		return CMC_OK;
	}

	int update_boundaries() override
	{
		if (!have_lines) return fail(CMC_ERR_INVALID, "update_boundaries: call cmc_adi3d_build_lines first");
		CU_TRY(cudaSetDevice(device));
		// (slabs on several devices: the neighbours' last sweep stores into the guard planes that are refreshed here)
		if (multi_device) await(neighbour_mask());
		span_begin(CMC_TIMING_BOUNDARY);
		for (auto *s : slabs) {
			use(s);
			ConstLayerPtrs<FT> nv; for (int q = 0; q < 4; q++) nv.f[q] = s->nodev[q];
			launch_update_boundaries<FT>(s->L, s->role[2], nv, s->layer(CMC_LAYER_CUR), s->stream, &launches);
		}
		span_end();
		return CMC_OK;
	}

	// ------------------------------------------------------------------------------------------- sweeps
	SweepArgs<FT> sweep_args(Slab<FT> *s, int dir, FT dt, int cur_layer, int next_layer)
	{
		SweepArgs<FT> A;
		A.L = s->L; A.dt = dt;
		A.h[0] = (FT)dx; A.h[1] = (FT)dy; A.h[2] = (FT)dz;
		A.v_T = (FT)params.v_T; A.v_vis = (FT)params.v_vis; A.t_vis = (FT)params.t_vis; A.t_phi = (FT)params.t_phi;
		A.role = s->role[dir];
		for (int q = 0; q < 4; q++) {
			A.cur[q] = s->field[s->slot[cur_layer]][q];
			A.temp[q] = s->field[s->slot[CMC_LAYER_TEMP]][q];
			A.next[q] = s->field[s->slot[next_layer]][q];
			A.temp_out[q] = s->field[s->spare][q];
			A.nodev[q] = s->nodev[q];
		}
		A.cv = s->cv; A.cT = s->cT;
		A.xcoef = s->xcoef_send; A.xbnd = s->xbnd_recv; A.lpo = (int)s->lines_per_owner(nslabs_total);
		A.extra_merge = 0;
		A.tile_counter = s->d_tilectr;
		A.tma_shape = tma_shape;
		for (int q = 0; q < 4; q++) A.push_lo[q] = A.push_hi[q] = A.pushn_lo[q] = A.pushn_hi[q] = nullptr;
		for (int r = 0; r < MAX_SLABS; r++) { A.xcoef_to[r] = nullptr; A.xs_tab_to[r] = nullptr; }
		A.xs_P = nslabs_total; A.xs_me = s->index; A.xs_epoch = xs_epoch; A.xs_share = 1; A.xs_tab = s->xs_tab;
		if (multi() && push_mode() && dir == CMC_DIR_X) {
			for (int r = 0; r < nslabs_total; r++) A.xs_tab_to[r] = in_slab(s, r, s->xs_tab);
			if (multi_device) {
				int share = 0;
				for (auto *o : slabs) share += o->device == s->device;
				A.xs_share = share;
			}
		}
		if (multi()) {
			const size_t lpo = (size_t)A.lpo;
			const int me = s->index;
			if (push_mode()) {
				// element (plane p, block 0, row 0, k 0) has the same offset in every slab's buffers (common block stride): the
				// lower neighbour's upper guard plane is ITS plane nx, the upper neighbour's lower guard plane is plane -1
				const long long hi_plane = me > 0 ? (long long)(split_nx[me - 1] + 1) * s->L.plane : 0, lo_plane = s->L.idx(-1, 0, 0);
				for (int q = 0; q < 4; q++) {
					if (me > 0) {
						A.push_lo[q] = in_slab(s, me - 1, s->field[s->spare][q]) + hi_plane;
						A.pushn_lo[q] = in_slab(s, me - 1, s->field[s->slot[next_layer]][q]) + hi_plane;
					}
					if (me + 1 < nslabs_total) {
						A.push_hi[q] = in_slab(s, me + 1, s->field[s->spare][q]) + lo_plane;
						A.pushn_hi[q] = in_slab(s, me + 1, s->field[s->slot[next_layer]][q]) + lo_plane;
					}
				}
				for (int r = 0; r < nslabs_total; r++) A.xcoef_to[r] = in_slab(s, r, s->xcoef_recv) + (size_t)me * 16 * lpo;
			} else {
				for (int r = 0; r < nslabs_total; r++) A.xcoef_to[r] = s->xcoef_send + (size_t)r * 16 * lpo;
			}
		}
		return A;
	}

	// fast mode covers lines of up to 512 rows in every kernel, and x / y lines of up to 1024 rows on one slab through the
	// CTA-pair form of the TMA kernel; longer lines run the exact kernels + a separate merge
	bool fast_ok(int dir) const
	{
		if (mode != CMC_MODE_FAST) return false;
		// (fp32, lines above 512 rows: the partition solve's rounding reaches 1.4e-5 of the field there - measured - against the
		// 1e-5 the fast mode promises; those lines take the exact kernels)
		if (sizeof(FT) == 4 && (dir == CMC_DIR_X ? slabs[0]->L.nx : dir == CMC_DIR_Y ? slabs[0]->L.ny : slabs[0]->L.nz) > 512) return false;
		return fast_sweep_supported(slabs[0]->L, dir) || (!multi() && want_tma(dir) && tma_sweep_supported(slabs[0]->L, dir));
	}

	// which kernel a sweep along `dir` runs (get_option "kernel_x|y|z"): 0 exact Thomas kernels + merge, 1 direct-load
	// partition kernel (kernels_fast.cu), 2 cp.async ring kernel (kernels_ring.cu), 4 slab-coupled x-sweep (spike pass +
	// interface solve + coupled pass)
	int debug_counter(int i, int64_t *value) override       // (kernel instrumentation builds: words of the tile-counter block)
	{
		int v = 0;
		if (i < 0 || i >= 32) return fail(CMC_ERR_INVALID, "debug counter index");
		use(slabs[0]);
		CU_TRY(cudaMemcpy(&v, slabs[0]->d_tilectr + i, sizeof(int), cudaMemcpyDeviceToHost));
		*value = v;
		return CMC_OK;
	}
	int kernel_kind(int dir) const override
	{
		if (multi() && dir == CMC_DIR_X && mode == CMC_MODE_FAST) return fused_x() ? 5 : 4;
		if (!fast_ok(dir)) return 0;
		if (want_tma(dir) && tma_sweep_supported(slabs[0]->L, dir)) return 3;
		return want_ring(dir) && ring_sweep_supported(slabs[0]->L, dir) ? 2 : 1;
	}
	// TMA-staged persistent tiles for the strided axes (kernels_tma.cu): option "tma" (bit 0 = x, bit 1 = y), default from
	// CMC_TMA=<subset of "xy"> ("" = off)
	bool want_tma(int dir) const { return dir != CMC_DIR_Z && ((tma_mask >> dir) & 1); }
	// two data-movement variants of the same arithmetic (kernels_ring.cu / kernels_fast.cu).  Measured on B200 at 512^3
	// (profiles/r01_variants.md): fp64 - the direct-load kernel wins everywhere (z: 2-line tiles, four independent CTAs per
	// SM, 4.00 ms against 4.58 ms for the cp.async ring); fp32 - the ring wins along z (2.42 against 2.71 ms).
	// CMC_RING=<subset of "xyz"> overrides.
	static bool want_ring(int dir)
	{
		static const char *ring_env = getenv("CMC_RING");
		return ring_env ? strchr(ring_env, "xyz"[dir]) != nullptr : (dir == CMC_DIR_Z && sizeof(FT) == 4);
	}

	// AdiSolver3D::SolveDirection (AdiSolver3D.cpp:564-666): num_local x { solve every line for u,v,w,T ; merge }
	// temp_is_cur: the linearisation layer still equals `cur` (first sweep of a step; the temp<-cur copy is folded
	// away).  fold_post_merge: the last local iteration also applies the post-X MergeLayerTo (AdiSolver3D.cpp:354).
	// halos_ready: the guard planes the sweeps read are already valid (time_step in push mode: every sweep stores its
	// boundary planes into the neighbours' guard planes itself)
	int solve_direction_impl(int dir, FT dt, int nl, int cur_layer, int next_layer, bool temp_is_cur = false, bool fold_post_merge = false,
	                         bool halos_ready = false)
	{
		for (int it = 0; it < nl; it++) {
			const bool t_is_c = temp_is_cur && it == 0;
			int rc;
			if (!(push_mode() && halos_ready))
				if ((rc = halo_exchange(t_is_c ? CMC_LAYER_CUR : CMC_LAYER_TEMP))) return rc;   // x-stencils of the sweep read the neighbours' planes
			const bool coupled = multi() && dir == CMC_DIR_X;
			const bool exact_x = coupled && mode != CMC_MODE_FAST;          // bit-exact mode: the Thomas recurrence runs through the slabs as a chain
			const bool fused = coupled && !exact_x && fused_x();
			if (fused) xs_epoch++;
			// the neighbours have finished their previous sweep: their stores into this slab's guard planes are complete,
			// and they no longer read the guard planes this sweep is about to overwrite on their side
			await(neighbour_mask());
			if (coupled && !fused && !exact_x) {
				// partitioned solve along the decomposed axis: spike pass -> coefficients to the line owners -> interface
				// solve -> neighbour values back -> coupled sweep.  (replaces LaunchSolveSegments_X, AdiSolver3D.cu:524-640)
				span_begin(CMC_TIMING_X_SPIKE);
				for (auto *s : slabs) {
					SweepArgs<FT> A = sweep_args(s, dir, dt, cur_layer, next_layer);
					if (t_is_c)
						for (int q = 0; q < 4; q++) A.temp[q] = s->field[s->slot[CMC_LAYER_CUR]][q];
					use(s);
					if (!launch_x_spike<FT>(A, s->stream, &launches)) return fail(CMC_ERR_UNSUPPORTED, "x-spike pass: unsupported slab shape");
				}
				span_end();
				const size_t lpo = slabs[0]->lines_per_owner(nslabs_total);
				if (push_mode()) { span_begin(CMC_TIMING_COMM); publish(); await(all_mask()); span_end(); }
				else if ((rc = all_to_all(&Slab<FT>::xcoef_send, &Slab<FT>::xcoef_recv, lpo * 16))) return rc;
				span_begin(CMC_TIMING_X_INTERFACE);
				const long long nlines = (long long)G.ny * G.nz;
				for (auto *s : slabs) {
					const long long first = (long long)s->index * (long long)lpo;
					const int owned = (int)std::max(0ll, std::min((long long)lpo, nlines - first));
					FT *to[MAX_SLABS] = {};
					for (int r = 0; r < nslabs_total; r++)
						to[r] = push_mode() ? in_slab(s, r, s->xbnd_recv) + (size_t)s->index * 8 * lpo : s->xbnd_send + (size_t)r * 8 * lpo;
					use(s);
					launch_x_interface<FT>(nslabs_total, (int)lpo, owned, s->xcoef_recv, to, s->stream, &launches);
				}
				span_end();
				if (push_mode()) { span_begin(CMC_TIMING_COMM); publish(); await(all_mask()); span_end(); }
				else if ((rc = all_to_all(&Slab<FT>::xbnd_send, &Slab<FT>::xbnd_recv, lpo * 8))) return rc;
			}
			span_begin(CMC_TIMING_SWEEP_X + dir);
			std::vector<char> swapped(slabs.size(), 0);
			if (exact_x) {
				if ((rc = exact_x_chain(dt, cur_layer, next_layer, t_is_c))) return rc;
				for (auto *s : slabs) {
					use(s);
					launch_merge<FT>(s->L, s->role[dir], s->clayer(next_layer), s->layer(CMC_LAYER_TEMP), s->stream, &launches);
				}
			}
			for (size_t si = 0; si < slabs.size() && !exact_x; si++) {
				Slab<FT> *s = slabs[si];
				use(s);
				cudaStream_t stream = s->stream;
				SweepArgs<FT> A = sweep_args(s, dir, dt, cur_layer, next_layer);
				if (t_is_c)
					for (int q = 0; q < 4; q++) A.temp[q] = s->field[s->slot[CMC_LAYER_CUR]][q];
				A.extra_merge = (fold_post_merge && it == nl - 1) ? 1 : 0;
				bool done = false;
				if (fused) {
					if (!launch_tma_xs<FT>(A, stream, &launches)) return fail(CMC_ERR_UNSUPPORTED, "one-pass slab-coupled x-sweep: launch failed");
					done = true;
				} else if (coupled) {
					if (!launch_x_coupled<FT>(A, stream, &launches)) return fail(CMC_ERR_UNSUPPORTED, "coupled x-sweep: unsupported slab shape");
					done = true;
				} else if (fast_ok(dir)) {
					if (want_tma(dir)) done = launch_tma_sweep<FT>(dir, A, stream, &launches);
					if (!done && want_ring(dir)) done = launch_ring_sweep<FT>(dir, A, stream, &launches);
					if (!done) done = launch_fast_sweep<FT>(dir, A, stream, &launches);
					if (!done) return fail(CMC_ERR_UNSUPPORTED, "fast sweep: no kernel for this line length (lines above 512 rows need the TMA kernel)");
				}
				if (done) swapped[si] = 1;                               // merged temp went to the other buffer
				else {
					launch_exact_sweep<FT>(dir, A, stream, &launches);
					launch_merge<FT>(s->L, s->role[dir], s->clayer(next_layer), s->layer(CMC_LAYER_TEMP), stream, &launches);
				}
			}
			{   // debug facility: CMC_DEBUG_SYNC=1 synchronises after every sweep and reports the kernel that faulted
				static const bool dbg = getenv("CMC_DEBUG_SYNC") != nullptr;
				if (dbg) {
					cudaError_t e = cudaSuccess;
					for (auto *s : slabs) { use(s); const cudaError_t es = cudaStreamSynchronize(s->stream); if (es != cudaSuccess) e = es; }
					if (e != cudaSuccess) {
						char buf[160];
						snprintf(buf, sizeof buf, "sweep along %c (kernel kind %d, local iteration %d) failed: %s", "xyz"[dir], kernel_kind(dir), it, cudaGetErrorString(e));
						return fail(CMC_ERR_CUDA, buf);
					}
				}
			}
			// (after ALL slabs are launched: the push targets above are computed from the neighbours' buffer roles)
			for (size_t si = 0; si < slabs.size(); si++)
				if (swapped[si]) std::swap(slabs[si]->slot[CMC_LAYER_TEMP], slabs[si]->spare);
			span_end();
			publish();
		}
		return CMC_OK;
	}

	// CMC_MODE_EXACT along the decomposed axis: forward elimination from the first slab to the last, back substitution from the
	// last to the first; between neighbours one x-plane of (c' of both matrices, d' of u, v, w, T) goes up and one plane of the
	// carried solution comes down.  The operations are those of the undivided line, in the same order: the result is
	// bit-identical with the single-slab run (and the reference CPU solver); the slabs work one after the other along x.
	int exact_x_chain(FT dt, int cur_layer, int next_layer, bool t_is_c)
	{
		auto args = [&](Slab<FT> *s) {
			SweepArgs<FT> A = sweep_args(s, CMC_DIR_X, dt, cur_layer, next_layer);
			if (t_is_c)
				for (int q = 0; q < 4; q++) A.temp[q] = s->field[s->slot[CMC_LAYER_CUR]][q];
			return A;
		};
		const int nb = slabs[0]->L.nblk;
		const size_t pb = sizeof(FT) * (size_t)slabs[0]->L.plane;
		auto six = [&](Slab<FT> *s, FT *(&f)[6]) { f[0] = s->cv; f[1] = s->cT; for (int q = 0; q < 4; q++) f[2 + q] = s->field[s->slot[next_layer]][q]; };
		auto piece = [&](Slab<FT> *s, FT *f, int plane, int b) { return f + (long long)b * s->L.bstride + s->L.idx(plane, 0, 0); };
		if (nccl) {
			Slab<FT> *s = slabs[0];
			FT *f[6]; six(s, f);
			SweepArgs<FT> A = args(s);
			auto xfer = [&](int first, int plane, int peer, bool send) -> int {
				std::vector<P2P> ops;
				for (int k = first; k < 6; k++)
					for (int b = 0; b < nb; b++)
						ops.push_back(send ? P2P{piece(s, f[k], plane, b), nullptr, pb, peer} : P2P{nullptr, piece(s, f[k], plane, b), pb, peer});
				if (nccl_exchange(nccl, ops.data(), (int)ops.size(), stream)) return fail(CMC_ERR_COMM, nccl_error());
				launches += 1;
				return CMC_OK;
			};
			int rc;
			if (rank > 0 && (rc = xfer(0, -1, rank - 1, false))) return rc;
			launch_exact_x_pass<FT>(true, A, stream, &launches);
			if (rank + 1 < nranks && (rc = xfer(0, s->L.nx - 1, rank + 1, true))) return rc;
			if (rank + 1 < nranks && (rc = xfer(2, s->L.nx, rank + 1, false))) return rc;
			launch_exact_x_pass<FT>(false, A, stream, &launches);
			if (rank > 0 && (rc = xfer(2, -1, rank - 1, true))) return rc;
			return CMC_OK;
		}
		for (size_t si = 0; si < slabs.size(); si++) {
			Slab<FT> *s = slabs[si];
			if (si > 0) {
				Slab<FT> *p = slabs[si - 1];
				FT *fs[6], *fp[6]; six(s, fs); six(p, fp);
				join_streams();
				use(s);
				for (int k = 0; k < 6; k++)
					for (int b = 0; b < nb; b++) CU_TRY(cudaMemcpyAsync(piece(s, fs[k], -1, b), piece(p, fp[k], p->L.nx - 1, b), pb, cudaMemcpyDefault, s->stream));
			}
			use(s);
			launch_exact_x_pass<FT>(true, args(s), s->stream, &launches);
		}
		for (size_t si = slabs.size(); si-- > 0;) {
			Slab<FT> *s = slabs[si];
			if (si + 1 < slabs.size()) {
				Slab<FT> *u = slabs[si + 1];
				FT *fs[6], *fu[6]; six(s, fs); six(u, fu);
				join_streams();
				use(s);
				for (int k = 2; k < 6; k++)
					for (int b = 0; b < nb; b++) CU_TRY(cudaMemcpyAsync(piece(s, fs[k], s->L.nx, b), piece(u, fu[k], -1, b), pb, cudaMemcpyDefault, s->stream));
			}
			use(s);
			launch_exact_x_pass<FT>(false, args(s), s->stream, &launches);
		}
		join_streams();
		return CMC_OK;
	}

	int step_prologue() override { return step_prologue_impl(true); }

	int step_prologue_impl(bool copy_temp)
	{
		if (!have_lines) return fail(CMC_ERR_INVALID, "time_step: call cmc_adi3d_build_lines first");
		CU_TRY(cudaSetDevice(device));
		// cur -> next on BOUND and VALVE cells (AdiSolver3D.cpp:310-311); temp <- cur (:320)
		span_begin(CMC_TIMING_COPY);
		for (auto *s : slabs) {
			use(s);
			launch_copy_masked<FT>(s->L, s->role[2], R_BV, s->clayer(CMC_LAYER_CUR), s->layer(CMC_LAYER_NEXT), s->stream, &launches);
			if (copy_temp) launch_copy_full<FT>(s->L, s->clayer(CMC_LAYER_CUR), s->layer(CMC_LAYER_TEMP), s->stream, &launches);
		}
		span_end();
		return CMC_OK;
	}

	int enqueue_div_error(int logical_layer, bool halos_ready = false)
	{
		if (err_pending == kErrRing) {         // ring full: drain it (one stream synchronisation every kErrRing residuals)
			int rc = fetch_error();
			if (rc) return rc;
		}
		double *slot_h = h_err2 + (size_t)err_head * 2 * slabs.size();
		// the residual reads the i-1 plane (TimeLayer3D.h:614-621)
		if (push_mode() && halos_ready) await(neighbour_mask());
		else {
			int rc = halo_exchange(logical_layer);
			if (rc) return rc;
		}
		for (size_t i = 0; i < slabs.size(); i++) {
			Slab<FT> *s = slabs[i];
			use(s);
			cudaStream_t stream = s->stream;
			const int l = s->slot[logical_layer];
			launch_div_error<FT>(s->L, s->role[2], s->field[l][0], s->field[l][1], s->field[l][2], (FT)dx, (FT)dy, (FT)dz,
			                     s->d_partials, Slab<FT>::kMaxErrBlocks, s->d_err2, stream, &launches);
			if (nccl && nccl_allreduce_sum_f64(nccl, s->d_err2, 2, stream)) return fail(CMC_ERR_COMM, nccl_error());
			CU_TRY(cudaMemcpyAsync(slot_h + 2 * i, s->d_err2, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
		}
		err_head = (err_head + 1) % kErrRing;
		err_pending++;
		return CMC_OK;
	}

	double err_from_host(int entry) const
	{
		const double *slot_h = h_err2 + (size_t)entry * 2 * slabs.size();
		double e = 0.0, c = 0.0;
		for (size_t i = 0; i < slabs.size(); i++) { e += slot_h[2 * i]; c += slot_h[2 * i + 1]; }
		return e / c;           // err / count (TimeLayer3D.h:639); 0/0 = NaN like the reference
	}

	// waits for the enqueued residuals: diffError = the latest one, worst_err = the largest of them
	int fetch_error()
	{
		worst_err = diffError;
		if (err_pending) {
			{ int rc = sync_all(); if (rc) return rc; }
			worst_err = 0.0;
			for (int k = err_pending; k >= 1; k--) {
				const double e = err_from_host((err_head - k + 2 * kErrRing) % kErrRing);
				diffError = e;
				if (e > worst_err || e != e) worst_err = e;
			}
			err_pending = 0;
		}
		return CMC_OK;
	}

	// a peer that never published its epoch (dead or stalled rank): k_peer_wait gave up after 10 s and the kernels after it
	// ran on stale guard planes / interface tables.  Called after every host-visible synchronisation.
	int check_peers()
	{
		if (!p2p) return CMC_OK;
		int t = 0;
		CU_TRY(cudaMemcpy(&t, d_timeout, sizeof t, cudaMemcpyDeviceToHost));
		if (t) {
			CU_TRY(cudaMemset(d_timeout, 0, sizeof(int)));
			return fail(CMC_ERR_COMM, "a peer rank did not publish its epoch within 10 s (peer_wait timed out): the fields of this step are invalid");
		}
		return CMC_OK;
	}

	// AdiSolver3D::TimeStep (AdiSolver3D.cpp:306-391)
	int time_step(double dt_in, int ng, int nl, int ce, double *err, bool async) override
	{
		int rc;
		if (ng < 0 || nl < 0) return fail(CMC_ERR_INVALID, "time_step: negative iteration count");
		const FT dt = (FT)dt_in;                                   // FluidSolver3D.cpp:242 casts to FTYPE
		// fast mode folds two full-field passes into the sweeps (identical arithmetic, fewer HBM round trips):
		//  * temp <- cur (:320): the first Z sweep reads `cur` as its linearisation layer;
		//  * the post-X MergeLayerTo (:354): the last X sweep of a global iteration relaxes twice.
		const bool fold_copy = ng > 0 && nl > 0 && fast_ok(CMC_DIR_Z);
		const bool fold_merge = nl > 0 && mode == CMC_MODE_FAST && (multi() || fast_ok(CMC_DIR_X));
		if ((rc = step_prologue_impl(!fold_copy))) return rc;
		// push mode: every sweep stores its boundary planes into the neighbours' guard planes itself (temp' always,
		// `next` by the x-sweep, which makes the guard planes of the next step's `cur` valid as well), so one explicit
		// exchange is needed only when something else touched the layers (first step, write_field, GetLayer)
		const bool pushing = push_mode() && fold_copy && fold_merge;
		if (pushing && halos_dirty) {
			if ((rc = halo_exchange(CMC_LAYER_CUR))) return rc;
		}
		halos_dirty = !pushing;
		for (int it = 0; it < ng; it++) {                          // :335-358
			if ((rc = solve_direction_impl(CMC_DIR_Z, dt, nl, CMC_LAYER_CUR, CMC_LAYER_NEXT, fold_copy && it == 0, false, pushing))) return rc;
			if ((rc = solve_direction_impl(CMC_DIR_Y, dt, nl, CMC_LAYER_NEXT, CMC_LAYER_HALF, false, false, pushing))) return rc;
			if ((rc = solve_direction_impl(CMC_DIR_X, dt, nl, CMC_LAYER_HALF, CMC_LAYER_NEXT, false, fold_merge, pushing))) return rc;
			if (!fold_merge) {
				// update non-linear layer once more (:354): temp = (temp + next) / 2 on NODE_IN
				span_begin(CMC_TIMING_MERGE);
				for (auto *s : slabs) {
					use(s);
					launch_merge<FT>(s->L, s->role[2], s->clayer(CMC_LAYER_NEXT), s->layer(CMC_LAYER_TEMP), s->stream, &launches);
				}
				span_end();
			}
		}
		if (ce) {
			span_begin(CMC_TIMING_RESIDUAL);
			rc = enqueue_div_error(CMC_LAYER_NEXT, pushing);
			span_end();
			if (rc) return rc;
		}
		if (!async) {
			if ((rc = fetch_error())) return rc;
			if ((rc = sync_all())) return rc;
			CU_TRY(cudaGetLastError());
			if ((rc = check_peers())) return rc;
			if (err) *err = diffError;
			if (diffError > CMC_ERR_THRESHOLD) {                   // :371-374 (layers are not swapped)
				char buf[96];
				snprintf(buf, sizeof buf, "Error is too big! %f", diffError);
				return fail(CMC_ERR_DIVERGED, buf);
			}
		}
		for (auto *s : slabs) std::swap(s->slot[CMC_LAYER_CUR], s->slot[CMC_LAYER_NEXT]);      // :388-390
		return CMC_OK;
	}

	int sync(double *err) override
	{
		CU_TRY(cudaSetDevice(device));
		int rc = fetch_error();
		if (rc) return rc;
		if ((rc = sync_all())) return rc;
		CU_TRY(cudaGetLastError());
		if ((rc = check_peers())) return rc;
		if (err) *err = diffError;
		// every residual enqueued since the last synchronisation is checked, not only the latest one.  (The layers of an
		// asynchronous run have been swapped by then - unlike cmc_adi3d_time_step, which leaves them unswapped.)
		if (worst_err > CMC_ERR_THRESHOLD) {
			char buf[96];
			snprintf(buf, sizeof buf, "Error is too big! %f", worst_err);
			worst_err = 0.0;
			return fail(CMC_ERR_DIVERGED, buf);
		}
		return CMC_OK;
	}

	int solve_direction(int dir, double dt, int nl, int cur_layer, int next_layer) override
	{
		if (!have_lines) return fail(CMC_ERR_INVALID, "solve_direction: call cmc_adi3d_build_lines first");
		if (dir < 0 || dir > 2 || cur_layer < 0 || cur_layer > 3 || next_layer < 0 || next_layer > 3 || cur_layer == next_layer)
			return fail(CMC_ERR_INVALID, "solve_direction: bad direction or layers");
		CU_TRY(cudaSetDevice(device));
		halos_dirty = true;
		int rc = solve_direction_impl(dir, (FT)dt, nl, cur_layer, next_layer);
		if (rc) return rc;
		if ((rc = sync_all())) return rc;
		CU_TRY(cudaGetLastError());
		return CMC_OK;
	}

	int eval_div_error(int logical, double *err) override
	{
		if (!have_lines) return fail(CMC_ERR_INVALID, "eval_div_error: call cmc_adi3d_build_lines first");
		CU_TRY(cudaSetDevice(device));
		int rc = fetch_error();                  // residuals of earlier asynchronous steps stay accounted for
		if (rc) return rc;
		const double keep = diffError, keep_worst = worst_err;
		if ((rc = enqueue_div_error(logical))) return rc;
		if ((rc = fetch_error())) return rc;
		if (err) *err = diffError;
		diffError = keep; worst_err = keep_worst;
		CU_TRY(cudaGetLastError());
		return check_peers();
	}

	int field_sums(int logical, double *sums8) override
	{
		if (!have_lines) return fail(CMC_ERR_INVALID, "field_sums: call cmc_adi3d_build_lines first");
		CU_TRY(cudaSetDevice(device));
		double tot[8] = {};
		for (auto *s : slabs) {
			// (d_partials holds 2 doubles per block for the residual: use a quarter of the blocks here)
			use(s);
			cudaStream_t stream = s->stream;
			launch_field_sums<FT>(s->L, s->role[2], s->clayer(logical), s->d_partials, Slab<FT>::kMaxErrBlocks / 4, s->d_sums8, stream, &launches);
			if (nccl && nccl_allreduce_sum_f64(nccl, s->d_sums8, 8, stream)) return fail(CMC_ERR_COMM, nccl_error());
			double h8[8];
			CU_TRY(cudaMemcpyAsync(h8, s->d_sums8, sizeof h8, cudaMemcpyDeviceToHost, stream));
			CU_TRY(cudaStreamSynchronize(stream));
			for (int q = 0; q < 8; q++) tot[q] += h8[q];
		}
		for (int q = 0; q < 8; q++) sums8[q] = tot[q];
		CU_TRY(cudaGetLastError());
		return check_peers();
	}

	// Solver3D::GetLayer (Solver3D.cpp:21-25)
	int get_layer(void *vel, double *T, int ox, int oy, int oz) override { return get_layer_impl(vel, T, ox, oy, oz, false); }
	int get_layer_async(void *vel, double *T, int ox, int oy, int oz) override { return get_layer_impl(vel, T, ox, oy, oz, true); }

	// waits for every readback started by get_layer_async
	int get_layer_wait() override
	{
		for (auto *s : slabs) {
			use(s);
			for (int k = 0; k < 2; k++)
				if (s->out_pending[k]) { CU_TRY(cudaEventSynchronize(s->ev_out_done[k])); s->out_pending[k] = false; }
		}
		if (multi_device) cudaSetDevice(device);
		CU_TRY(cudaGetLastError());
		return CMC_OK;
	}

	// async: the device part (Clear + FilterToArrays + gather) is enqueued on the solver's stream and the copy to the host
	// on the slab's copy stream behind it, into a second staging set - the caller goes on with the next time step and
	// collects the result with get_layer_wait (reference: TimeLayer3D::FilterToArrays + OutputNetCDF3D_layer run between
	// two time steps, TimeLayer3D.h:819-924, IO.h:350-388)
	int get_layer_impl(void *vel, double *T, int ox, int oy, int oz, bool async)
	{
		if (!have_lines) return fail(CMC_ERR_INVALID, "get_layer: call cmc_adi3d_build_lines first");
		CU_TRY(cudaSetDevice(device));
		if (ox == 0) ox = G.nx;
		if (oy == 0) oy = G.ny;
		if (oz == 0) oz = G.nz;
		if (ox < 0 || oy < 0 || oz < 0) return fail(CMC_ERR_INVALID, "get_layer: negative output dims");
		halos_dirty = true;                // Clear() below rewrites the NODE_OUT cells of a layer
		const size_t outN = (size_t)ox * oy * oz, rowN = (size_t)oy * oz;
		span_begin(CMC_TIMING_READBACK);
		std::vector<int> lo(nslabs_total, ox), hi(nslabs_total, 0);
		// output rows i whose source plane x = i*dimx/outdimx lies in slab r (TimeLayer3D.h:842-854)
		for (int r = 0; r < nslabs_total; r++) {
			const int x0 = split_x0[r], nx = split_nx[r];
			for (int i = 0; i < ox; i++) {
				const int x = (int)((long long)i * G.nx / ox);
				if (x >= x0 && x < x0 + nx) { if (i < lo[r]) lo[r] = i; if (i + 1 > hi[r]) hi[r] = i + 1; }
			}
		}
		for (auto *s : slabs) {
			use(s);
			cudaStream_t stream = s->stream;
			launch_clear_out<FT>(s->L, s->role[2], s->layer(CMC_LAYER_NEXT), (FT)CMC_MISSING_VALUE, stream, &launches);
			const bool gather = nccl && !local_output;
			const size_t need = (gather && rank == 0) ? outN : (size_t)std::max(0, hi[s->index] - lo[s->index]) * rowN;
			if (async) { int rc = s->io_setup(); if (rc) return rc; }
			const int k = async ? (s->out_idx ^= 1) : s->out_idx;
			if (s->out_pending[k]) {          // this staging set still feeds an earlier asynchronous copy
				if (need > s->out_cap2[k]) { CU_TRY(cudaEventSynchronize(s->ev_out_done[k])); }
				else CU_TRY(cudaStreamWaitEvent(stream, s->ev_out_done[k], 0));
				s->out_pending[k] = false;
			}
			if (need > s->out_cap2[k]) {
				if (s->d_outvel2[k]) cudaFree(s->d_outvel2[k]);
				if (s->d_outT2[k]) cudaFree(s->d_outT2[k]);
				s->d_outvel2[k] = nullptr; s->d_outT2[k] = nullptr;
				CU_TRY(cudaMalloc((void **)&s->d_outvel2[k], need * 3 * sizeof(FT)));
				CU_TRY(cudaMalloc((void **)&s->d_outT2[k], need * sizeof(double)));
				s->out_cap2[k] = need;
			}
			s->d_outvel = s->d_outvel2[k]; s->d_outT = s->d_outT2[k];
			// slab-local output buffer starts at output row lo (rank 0 of an NCCL run: at row 0, it also receives)
			const int oi0 = lo[s->index], oi1 = hi[s->index];
			const size_t shift = (gather && rank == 0) ? 0 : (size_t)std::max(oi0, 0) * rowN;
			if (oi1 > oi0)
				launch_filter<FT>(s->L, s->clayer(CMC_LAYER_NEXT), ox, oy, oz, oi0, oi1, s->d_outvel - 3 * shift, s->d_outT - shift, stream, &launches);
		}
		span_end();
		if (nccl && !local_output) {
			Slab<FT> *s = slabs[0];
			std::vector<P2P> ops;
			if (rank == 0) {
				for (int r = 1; r < nranks; r++)
					if (hi[r] > lo[r]) {
						const size_t o0 = (size_t)lo[r] * rowN, cnt = (size_t)(hi[r] - lo[r]) * rowN;
						ops.push_back(P2P{nullptr, s->d_outvel + 3 * o0, cnt * 3 * sizeof(FT), r});
						ops.push_back(P2P{nullptr, s->d_outT + o0, cnt * sizeof(double), r});
					}
			} else if (hi[rank] > lo[rank]) {
				const size_t cnt = (size_t)(hi[rank] - lo[rank]) * rowN;
				ops.push_back(P2P{s->d_outvel, nullptr, cnt * 3 * sizeof(FT), 0});
				ops.push_back(P2P{s->d_outT, nullptr, cnt * sizeof(double), 0});
			}
			if (!ops.empty() && nccl_exchange(nccl, ops.data(), (int)ops.size(), stream)) return fail(CMC_ERR_COMM, nccl_error());
			if (rank == 0) {
				cudaStream_t cs = stream;
				if (async) { CU_TRY(cudaEventRecord(s->ev_filtered, stream)); CU_TRY(cudaStreamWaitEvent(s->io_out, s->ev_filtered, 0)); cs = s->io_out; }
				CU_TRY(cudaMemcpyAsync(vel, s->d_outvel, outN * 3 * sizeof(FT), cudaMemcpyDeviceToHost, cs));
				CU_TRY(cudaMemcpyAsync(T, s->d_outT, outN * sizeof(double), cudaMemcpyDeviceToHost, cs));
				if (async) { CU_TRY(cudaEventRecord(s->ev_out_done[s->out_idx], cs)); s->out_pending[s->out_idx] = true; }
			}
		} else {
			for (auto *s : slabs) {
				const int oi0 = lo[s->index], oi1 = hi[s->index];
				if (oi1 <= oi0) continue;
				use(s);
				// (one process per GPU with local_output: the caller's arrays hold this rank's rows only)
				const size_t o0 = nccl ? 0 : (size_t)oi0 * rowN, cnt = (size_t)(oi1 - oi0) * rowN;
				cudaStream_t cs = s->stream;
				if (async) { CU_TRY(cudaEventRecord(s->ev_filtered, s->stream)); CU_TRY(cudaStreamWaitEvent(s->io_out, s->ev_filtered, 0)); cs = s->io_out; }
				CU_TRY(cudaMemcpyAsync((FT *)vel + 3 * o0, s->d_outvel, cnt * 3 * sizeof(FT), cudaMemcpyDeviceToHost, cs));
				CU_TRY(cudaMemcpyAsync(T + o0, s->d_outT, cnt * sizeof(double), cudaMemcpyDeviceToHost, cs));
				if (async) { CU_TRY(cudaEventRecord(s->ev_out_done[s->out_idx], cs)); s->out_pending[s->out_idx] = true; }
			}
		}
		if (multi_device) cudaSetDevice(device);
		if (async) return CMC_OK;
		{ int rc = sync_all(); if (rc) return rc; }
		CU_TRY(cudaGetLastError());
		return check_peers();
	}

	// Asynchronous upload of a whole layer (u, v, w, T: dense host arrays of this handle's planes): the copies run on the
	// copy stream into a dense staging buffer while the solver's stream keeps computing; write_layer_commit makes the
	// solver's stream wait for them and scatters the staging buffer into the (padded, y-blocked) layer - one pass at HBM
	// speed.  (Solver3D::SetLayer-style state injection; TimeLayer3D::CopyFromGrid for the fields, TimeLayer3D.h:926-951.)
	int write_layer_async(const void *const src[4]) override
	{
		for (auto *s : slabs) {
			use(s);
			int rc = s->io_setup();
			if (rc) return rc;
			const size_t n = (size_t)s->L.nx * s->L.ny * s->L.nz;
			const size_t o = (size_t)(s->L.x0 - L.x0) * s->L.ny * s->L.nz;
			if (s->in_pending) CU_TRY(cudaStreamWaitEvent(s->io_in, s->ev_in_free, 0));   // the previous commit has read the staging buffer
			for (int q = 0; q < 4; q++) {
				if (!s->d_stage_in[q]) { CU_TRY(cudaMalloc((void **)&s->d_stage_in[q], n * sizeof(FT))); s->bytes += (long long)(n * sizeof(FT)); }
				CU_TRY(cudaMemcpyAsync(s->d_stage_in[q], (const FT *)src[q] + o, n * sizeof(FT), cudaMemcpyHostToDevice, s->io_in));
			}
			CU_TRY(cudaEventRecord(s->ev_in_done, s->io_in));
			s->in_pending = true;
		}
		if (multi_device) cudaSetDevice(device);
		return CMC_OK;
	}

	int write_layer_commit(int logical) override
	{
		if (logical < 0 || logical > 3) return fail(CMC_ERR_INVALID, "write_layer_commit: bad layer");
		for (auto *s : slabs) {
			if (!s->in_pending) return fail(CMC_ERR_INVALID, "write_layer_commit: no upload pending (call cmc_adi3d_write_layer_async first)");
			use(s);
			CU_TRY(cudaStreamWaitEvent(s->stream, s->ev_in_done, 0));
			ConstLayerPtrs<FT> st; for (int q = 0; q < 4; q++) st.f[q] = s->d_stage_in[q];
			launch_scatter_dense<FT>(s->L, st, s->layer(logical), s->stream, &launches);
			CU_TRY(cudaEventRecord(s->ev_in_free, s->stream));
		}
		if (multi_device) cudaSetDevice(device);
		halos_dirty = true;
		return CMC_OK;
	}

	// dense host copy of the locally held planes (all local slabs, in x order)
	int read_field(int logical, int var, void *dst) override
	{
		if (logical < 0 || logical > 3 || var < 0 || var > 3) return fail(CMC_ERR_INVALID, "read_field: bad layer/var");
		CU_TRY(cudaSetDevice(device));
		for (auto *s : slabs) {
			use(s);
			int rc = s->copy_planes(s->field[s->slot[logical]][var], nullptr, (FT *)dst, 0, s->L.nx, L.x0);
			if (rc) return rc;
		}
		{ int rc = sync_all(); if (rc) return rc; }
		return CMC_OK;
	}

	int write_field(int logical, int var, const void *src) override
	{
		if (logical < 0 || logical > 3 || var < 0 || var > 3) return fail(CMC_ERR_INVALID, "write_field: bad layer/var");
		CU_TRY(cudaSetDevice(device));
		halos_dirty = true;
		for (auto *s : slabs) {
			use(s);
			int rc = s->copy_planes(s->field[s->slot[logical]][var], (const FT *)src, nullptr, 0, s->L.nx, L.x0);
			if (rc) return rc;
		}
		{ int rc = sync_all(); if (rc) return rc; }
		return CMC_OK;
	}
};

static int check_device(int device)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n < 1)
		return fail(CMC_ERR_NO_DEVICE, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
	if (device < 0 || device >= n) return fail(CMC_ERR_INVALID, "device index out of range");
	cudaDeviceProp pr;
	if (cudaGetDeviceProperties(&pr, device) != cudaSuccess) return fail(CMC_ERR_CUDA, "cudaGetDeviceProperties failed");
	if (pr.major < 10)
		return fail(CMC_ERR_NO_DEVICE, std::string("device ") + pr.name + " is not sm_100 class; this library is built for sm_100a only");
	return CMC_OK;
}

// first_slab / nlocal / ntotal: which slabs of the x-split this handle holds
static int create_impl(const cmc_grid_desc *grid, const cmc_fluid_params *params, int fp_bytes, int device,
                       int first_slab, int nlocal, int ntotal, const void *nccl_id, cmc_adi3d **out,
                       const int *devices = nullptr, const int *planes = nullptr)
{
	if (!grid || !params || !out) return fail(CMC_ERR_INVALID, "create: null argument");
	*out = nullptr;
	if (fp_bytes != 4 && fp_bytes != 8) return fail(CMC_ERR_INVALID, "create: fp_bytes must be 4 or 8");
	if (grid->dimx < 3 || grid->dimy < 3 || grid->dimz < 3) return fail(CMC_ERR_INVALID, "create: every grid dimension must be >= 3");
	if (!(grid->dx > 0) || !(grid->dy > 0) || !(grid->dz > 0)) return fail(CMC_ERR_INVALID, "create: grid spacing must be positive");
	if (ntotal < 1 || ntotal > 16 || first_slab < 0 || nlocal < 1 || first_slab + nlocal > ntotal) return fail(CMC_ERR_INVALID, "create: bad slab / rank count (1..16)");
	if (ntotal > grid->dimx) return fail(CMC_ERR_INVALID, "create: more slabs than x-planes");
	int rc = check_device(device);
	if (rc) return rc;
	std::vector<int> devs;
	if (devices) {
		for (int i = 0; i < nlocal; i++) {
			if ((rc = check_device(devices[i]))) return rc;
			// (CMC_SHARE_DEVICE=1: a test configuration - several slabs, each with its own stream, on one GPU)
			static const bool share_ok = getenv("CMC_SHARE_DEVICE") && atoi(getenv("CMC_SHARE_DEVICE")) != 0;
			for (int j = 0; j < i && !share_ok; j++) if (devices[j] == devices[i]) return fail(CMC_ERR_INVALID, "create_multi: a device is listed twice");
			devs.push_back(devices[i]);
		}
	}
	cmc_adi3d *h = nullptr;
	NcclComm *comm = nullptr;
	if (nccl_id) {
		cudaSetDevice(device);
		comm = nccl_create(first_slab, ntotal, nccl_id);
		if (!comm) return fail(CMC_ERR_COMM, nccl_error());
	}
	if (fp_bytes == 4) {
		auto *s = new (std::nothrow) Engine<float>();
		if (!s) return fail(CMC_ERR_INVALID, "out of host memory");
		s->device = device; s->fp = 4; s->nccl = comm;
		if (devs.size() > 1) { s->multi_device = true; s->slab_devices = devs; }
		if (planes) s->planes_arg.assign(planes, planes + ntotal);
		rc = s->init(grid, params, first_slab, nlocal, ntotal);
		h = s;
	} else {
		auto *s = new (std::nothrow) Engine<double>();
		if (!s) return fail(CMC_ERR_INVALID, "out of host memory");
		s->device = device; s->fp = 8; s->nccl = comm;
		if (devs.size() > 1) { s->multi_device = true; s->slab_devices = devs; }
		if (planes) s->planes_arg.assign(planes, planes + ntotal);
		rc = s->init(grid, params, first_slab, nlocal, ntotal);
		h = s;
	}
	if (rc) { delete h; return rc; }
	*out = h;
	return CMC_OK;
}

} // namespace

// shared with cmc_adi2d.cu
int cmc_set_error(int code, const std::string &msg) { return fail(code, msg); }
int cmc_check_device(int device) { return check_device(device); }

// =================================================== C ABI ===================================================
extern "C" {

const char *cmc_last_error(void) { return g_err.c_str(); }
int cmc_abi_version(void) { return CMC_ADI_ABI_VERSION; }

int cmc_device_count(void)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n < 1) return fail(CMC_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
	return n;
}

int cmc_adi3d_create(const cmc_grid_desc *grid, const cmc_fluid_params *params, int fp_bytes, int device, cmc_adi3d **out)
{
	return create_impl(grid, params, fp_bytes, device, 0, 1, 1, nullptr, out);
}

int cmc_adi3d_create_emulated(const cmc_grid_desc *grid, const cmc_fluid_params *params, int fp_bytes, int device,
                              int n_slabs, cmc_adi3d **out)
{
	return create_impl(grid, params, fp_bytes, device, 0, n_slabs, n_slabs, nullptr, out);
}

int cmc_adi3d_create_multi(const cmc_grid_desc *grid, const cmc_fluid_params *params, int fp_bytes, const int *devices, int n_devices,
                           cmc_adi3d **out)
{
	if (!devices || n_devices < 1) return fail(CMC_ERR_INVALID, "create_multi: no devices");
	if (n_devices == 1) return create_impl(grid, params, fp_bytes, devices[0], 0, 1, 1, nullptr, out);
	return create_impl(grid, params, fp_bytes, devices[0], 0, n_devices, n_devices, nullptr, out, devices);
}

int cmc_adi3d_create_ex(const cmc_grid_desc *grid, const cmc_fluid_params *params, int fp_bytes, const cmc_decomp *d, cmc_adi3d **out)
{
	if (!d) return fail(CMC_ERR_INVALID, "create_ex: null decomposition");
	const int n = d->n_slabs < 1 ? 1 : d->n_slabs;
	switch (d->kind) {
	case 0: return create_impl(grid, params, fp_bytes, d->device, 0, 1, 1, nullptr, out);
	case 1: return create_impl(grid, params, fp_bytes, d->device, 0, n, n, nullptr, out, nullptr, d->planes);
	case 2:
		if (n > 1 && !d->nccl_unique_id) return fail(CMC_ERR_INVALID, "create_ex: nccl_unique_id required when n_slabs > 1");
		return create_impl(grid, params, fp_bytes, d->device, d->rank, 1, n, n > 1 ? d->nccl_unique_id : nullptr, out, nullptr, d->planes);
	case 3:
		if (!d->devices) return fail(CMC_ERR_INVALID, "create_ex: no devices");
		if (n == 1) return create_impl(grid, params, fp_bytes, d->devices[0], 0, 1, 1, nullptr, out);
		return create_impl(grid, params, fp_bytes, d->devices[0], 0, n, n, nullptr, out, d->devices, d->planes);
	}
	return fail(CMC_ERR_INVALID, "create_ex: unknown decomposition kind");
}

// Grid3D::SplitSegments_X (reference Grid3D.cpp:148-235) with the cuts on multiples of 8 planes
int cmc_split_planes(int policy, const cmc_grid_desc *grid, const int32_t *type, int n, int32_t *planes_out)
{
	if (!grid || !planes_out || n < 1) return fail(CMC_ERR_INVALID, "split_planes: bad argument");
	const int dimx = grid->dimx, dimy = grid->dimy, dimz = grid->dimz;
	if (dimx < 8 * n && n > 1) return fail(CMC_ERR_UNSUPPORTED, "split_planes: fewer than 8 planes per slab");
	std::vector<int> planes;
	if (policy == CMC_SPLIT_EVEN_X || n == 1) {
		split_default(dimx, n, planes);
	} else {
		if (!type) return fail(CMC_ERR_INVALID, "split_planes: node types required for this policy");
		std::vector<double> w((size_t)dimx, 0.0);
		double total = 0.0;
		auto T = [&](int i, int j, int k) { return type[((size_t)i * dimy + j) * dimz + k]; };
		if (policy == CMC_SPLIT_EVEN_VOLUME) {
			for (int i = 0; i < dimx; i++) {
				double c = 0.0;
				for (size_t id = (size_t)i * dimy * dimz, e = id + (size_t)dimy * dimz; id < e; id++) c += type[id] == CMC_NODE_IN;
				w[i] = c; total += c;
			}
		} else if (policy == CMC_SPLIT_EVEN_SEGMENTS) {
			// the segment rule of GenerateListSegments (Grid3D.cpp:47-127): a run of NODE_IN cells closed by a non-IN cell
			auto scan = [&](int len, auto cell, auto emit) {
				int state = 0, start = 0;
				for (int p = 0; p + 1 < len; p++) {
					if (cell(p + 1) == CMC_NODE_IN) { if (!state) start = p; state = 1; }
					else if (state) { emit(start, p + 1); state = 0; }
				}
			};
			for (int i = 0; i < dimx; i++) {
				for (int k = 0; k < dimz; k++) scan(dimy, [&](int p) { return T(i, p, k); }, [&](int, int) { w[i] += 1.0; total += 1.0; });
				for (int j = 0; j < dimy; j++) scan(dimz, [&](int p) { return T(i, j, p); }, [&](int, int) { w[i] += 1.0; total += 1.0; });
			}
			for (int j = 0; j < dimy; j++)
				for (int k = 0; k < dimz; k++)
					scan(dimx, [&](int p) { return T(p, j, k); }, [&](int a, int b) {
						const double sz = (double)(b - a + 1);
						for (int i = a; i <= b; i++) w[i] += 1.0 / sz;
						total += 1.0;
					});
		} else
			return fail(CMC_ERR_INVALID, "split_planes: unknown policy");
		// the reference cuts where the running load passes total / n (Grid3D.cpp:214-229); here: cut r where the cumulated
		// load reaches r * total / n, moved to the nearest multiple of 8 that leaves every slab at least 8 planes
		const double per = total / n;
		planes.assign(n, 0);
		int prev = 0, i = 0;
		double cum = 0.0;
		for (int r = 1; r < n; r++) {
			while (i < dimx && cum + w[i] <= per * r) cum += w[i++];
			int cut = (i + 4) / 8 * 8;
			if (cut < prev + 8) cut = prev + 8;
			const int room = dimx - 8 * (n - r);               // the slabs still to come need 8 planes each
			if (cut > room) cut = room / 8 * 8;
			if (cut < prev + 8) return fail(CMC_ERR_UNSUPPORTED, "split_planes: the grid is too small for this many slabs");
			planes[r - 1] = cut - prev;
			prev = cut;
		}
		planes[n - 1] = dimx - prev;
		if (planes[n - 1] < 8) return fail(CMC_ERR_UNSUPPORTED, "split_planes: the grid is too small for this many slabs");
	}
	for (int r = 0; r < n; r++) planes_out[r] = planes[r];
	return CMC_OK;
}

int cmc_nccl_unique_id(void *id128)
{
	if (!id128) return fail(CMC_ERR_INVALID, "null id");
	if (nccl_unique_id(id128)) return fail(CMC_ERR_COMM, nccl_error());
	return CMC_OK;
}

int cmc_adi3d_create_dist(const cmc_grid_desc *grid, const cmc_fluid_params *params, int fp_bytes, int device,
                          int rank, int nranks, const void *nccl_unique_id, cmc_adi3d **out)
{
	if (nranks > 1 && !nccl_unique_id) return fail(CMC_ERR_INVALID, "create_dist: nccl_unique_id required when nranks > 1");
	return create_impl(grid, params, fp_bytes, device, rank, 1, nranks, nranks > 1 ? nccl_unique_id : nullptr, out);
}

int cmc_adi3d_destroy(cmc_adi3d *h)
{
	delete h;
	return CMC_OK;
}

#define H_CHECK(h) if (!(h)) return fail(CMC_ERR_INVALID, "null handle")

int cmc_adi3d_output_rows(const cmc_adi3d *h, int outdimx, int *row_lo, int *row_hi)
{
	H_CHECK(h);
	if (outdimx == 0) outdimx = h->G.nx;
	if (outdimx < 0) return fail(CMC_ERR_INVALID, "output_rows: negative output dimension");
	int lo = outdimx, hi = 0;
	for (int i = 0; i < outdimx; i++) {          // source plane of output row i: i * dimx / outdimx (TimeLayer3D.h:842-854)
		const int x = (int)((long long)i * h->G.nx / outdimx);
		if (x >= h->L.x0 && x < h->L.x0 + h->L.nx) { if (i < lo) lo = i; if (i + 1 > hi) hi = i + 1; }
	}
	if (hi <= lo) lo = hi = 0;
	if (row_lo) *row_lo = lo;
	if (row_hi) *row_hi = hi;
	return CMC_OK;
}

int cmc_adi3d_slab(const cmc_adi3d *h, int *x0, int *nx)
{
	H_CHECK(h);
	if (x0) *x0 = h->L.x0;
	if (nx) *nx = h->L.nx;
	return CMC_OK;
}

int cmc_adi3d_set_nodes(cmc_adi3d *h, const int32_t *type, const int32_t *bc_vel, const int32_t *bc_temp,
                        const void *vx, const void *vy, const void *vz, const void *T)
{
	H_CHECK(h);
	if (!type || !bc_vel || !bc_temp || !vx || !vy || !vz || !T) return fail(CMC_ERR_INVALID, "set_nodes: null array");
	return h->set_nodes(type, bc_vel, bc_temp, vx, vy, vz, T, 0, false);
}

int cmc_adi3d_set_nodes_slab(cmc_adi3d *h, const int32_t *type, const int32_t *bc_vel, const int32_t *bc_temp,
                             const void *vx, const void *vy, const void *vz, const void *T, int halo)
{
	H_CHECK(h);
	if (!type || !bc_vel || !bc_temp || !vx || !vy || !vz || !T) return fail(CMC_ERR_INVALID, "set_nodes_slab: null array");
	return h->set_nodes_slab(type, bc_vel, bc_temp, vx, vy, vz, T, halo, false);
}

int cmc_adi3d_update_nodes(cmc_adi3d *h, const int32_t *type, const int32_t *bc_vel, const int32_t *bc_temp,
                           const void *vx, const void *vy, const void *vz, const void *T)
{
	H_CHECK(h);
	if (!type || !bc_vel || !bc_temp || !vx || !vy || !vz || !T) return fail(CMC_ERR_INVALID, "update_nodes: null array");
	return h->set_nodes(type, bc_vel, bc_temp, vx, vy, vz, T, 0, true);
}

int cmc_adi3d_update_nodes_aos(cmc_adi3d *h, const void *nodes, size_t stride)
{
	H_CHECK(h);
	if (!nodes) return fail(CMC_ERR_INVALID, "update_nodes_aos: null array");
	if (stride < (size_t)((h->fp == 8 ? 16 : 12) + 4 * h->fp)) return fail(CMC_ERR_INVALID, "update_nodes_aos: stride smaller than a Node");
	return h->set_nodes((const int32_t *)nodes, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stride, true);
}

int cmc_adi3d_set_nodes_aos(cmc_adi3d *h, const void *nodes, size_t stride)
{
	H_CHECK(h);
	if (!nodes) return fail(CMC_ERR_INVALID, "set_nodes_aos: null array");
	if (stride < (size_t)((h->fp == 8 ? 16 : 12) + 4 * h->fp)) return fail(CMC_ERR_INVALID, "set_nodes_aos: stride smaller than a Node");
	return h->set_nodes((const int32_t *)nodes, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stride, false);
}

int cmc_adi3d_build_lines(cmc_adi3d *h) { H_CHECK(h); return h->build_lines(); }

int cmc_adi3d_num_segments(const cmc_adi3d *h, int dir, int64_t *n)
{
	H_CHECK(h);
	if (dir < 0 || dir > 2 || !n) return fail(CMC_ERR_INVALID, "num_segments: bad argument");
	if (!h->have_lines) return fail(CMC_ERR_INVALID, "num_segments: call cmc_adi3d_build_lines first");
	*n = h->num_segs[dir];
	return CMC_OK;
}

int cmc_adi3d_update_boundaries(cmc_adi3d *h) { H_CHECK(h); return h->update_boundaries(); }

int cmc_adi3d_time_step(cmc_adi3d *h, double dt, int ng, int nl, int ce, double *err)
{
	H_CHECK(h);
	return h->time_step(dt, ng, nl, ce, err, false);
}

int cmc_adi3d_time_step_async(cmc_adi3d *h, double dt, int ng, int nl, int ce)
{
	H_CHECK(h);
	return h->time_step(dt, ng, nl, ce, nullptr, true);
}

int cmc_adi3d_sync(cmc_adi3d *h, double *err) { H_CHECK(h); return h->sync(err); }

int cmc_adi3d_get_layer(cmc_adi3d *h, void *vel, double *T, int ox, int oy, int oz)
{
	H_CHECK(h);
	if (h->rank == 0 && (!vel || !T)) return fail(CMC_ERR_INVALID, "get_layer: null output");
	return h->get_layer(vel, T, ox, oy, oz);
}

int cmc_adi3d_get_layer_async(cmc_adi3d *h, void *vel, double *T, int ox, int oy, int oz)
{
	H_CHECK(h);
	if (h->rank == 0 && (!vel || !T)) return fail(CMC_ERR_INVALID, "get_layer_async: null output");
	return h->get_layer_async(vel, T, ox, oy, oz);
}

int cmc_adi3d_get_layer_wait(cmc_adi3d *h) { H_CHECK(h); return h->get_layer_wait(); }

int cmc_adi3d_write_layer_async(cmc_adi3d *h, const void *u, const void *v, const void *w, const void *T)
{
	H_CHECK(h);
	if (!u || !v || !w || !T) return fail(CMC_ERR_INVALID, "write_layer_async: null source");
	const void *src[4] = {u, v, w, T};
	return h->write_layer_async(src);
}

int cmc_adi3d_write_layer_commit(cmc_adi3d *h, int layer) { H_CHECK(h); return h->write_layer_commit(layer); }

int cmc_adi3d_set_option(cmc_adi3d *h, const char *key, int64_t value)
{
	H_CHECK(h);
	if (!key) return fail(CMC_ERR_INVALID, "set_option: null key");
	if (!strcmp(key, "mode")) {
		if (value != CMC_MODE_FAST && value != CMC_MODE_EXACT) return fail(CMC_ERR_INVALID, "set_option: unknown mode");
		h->mode = (int)value;
		return CMC_OK;
	}
	if (!strcmp(key, "tma")) { h->tma_mask = (int)value & 3; return CMC_OK; }
	if (!strcmp(key, "xs")) { h->xs_enabled = value != 0; return CMC_OK; }
	if (!strcmp(key, "local_output")) { h->local_output = value != 0; return CMC_OK; }
	if (!strcmp(key, "tma_shape")) {
		const int nl = (int)value & 255, cl = (int)value >> 8;
		if (value != 0 && !((nl == 8 || nl == 16) && (cl == 1 || cl == 2))) return fail(CMC_ERR_INVALID, "set_option tma_shape: 0, or lines per tile (8 | 16) + 256 * CTAs per tile (1 | 2)");
		h->tma_shape = (int)value; return CMC_OK;
	}
	if (!strcmp(key, "profile")) {
		h->spans_collect();
		h->profile = value != 0;
		if (value == 2) for (int k = 0; k < CMC_TIMING_KINDS; k++) { h->kind_ms[k] = 0.0; h->kind_calls[k] = 0; }
		return CMC_OK;
	}
	return fail(CMC_ERR_INVALID, std::string("set_option: unknown key ") + key);
}

int cmc_adi3d_get_option(const cmc_adi3d *h, const char *key, int64_t *value)
{
	H_CHECK(h);
	if (!key || !value) return fail(CMC_ERR_INVALID, "get_option: null argument");
	if (!strcmp(key, "mode")) { *value = h->mode; return CMC_OK; }
	if (!strcmp(key, "tma")) { *value = h->tma_mask; return CMC_OK; }
	if (!strcmp(key, "tma_shape")) { *value = h->tma_shape; return CMC_OK; }
	if (!strcmp(key, "xs")) { *value = h->xs_enabled; return CMC_OK; }
	if (!strcmp(key, "local_output")) { *value = h->local_output; return CMC_OK; }
	if (!strncmp(key, "tilectr", 7) && key[7] >= '0' && key[7] <= '9') { return const_cast<cmc_adi3d *>(h)->debug_counter(atoi(key + 7), value); }
	if (!strcmp(key, "nzp")) { *value = h->L.nzp; return CMC_OK; }
	if (!strcmp(key, "jb")) { *value = h->L.nblk == 1 ? 0 : (1 << h->L.jbs); return CMC_OK; }   // rows per y-block, 0 = one block
	if (!strncmp(key, "kernel_", 7) && key[7] >= 'x' && key[7] <= 'z' && !key[8]) { *value = h->kernel_kind(key[7] - 'x'); return CMC_OK; }
	if (!strcmp(key, "exchange")) { *value = h->exchange_kind(); return CMC_OK; }   // 0 none, 1 NCCL send/recv, 2 fused stores (same device), 3 fused stores (peer memory)
	if (!strcmp(key, "shared_free_cells")) { *value = h->shared_free[0] + h->shared_free[1] + h->shared_free[2]; return CMC_OK; }
	return fail(CMC_ERR_INVALID, std::string("get_option: unknown key ") + key);
}

int cmc_adi3d_read_field(cmc_adi3d *h, int layer, int var, void *dst)
{
	H_CHECK(h);
	if (!dst) return fail(CMC_ERR_INVALID, "read_field: null destination");
	return h->read_field(layer, var, dst);
}

int cmc_adi3d_write_field(cmc_adi3d *h, int layer, int var, const void *src)
{
	H_CHECK(h);
	if (!src) return fail(CMC_ERR_INVALID, "write_field: null source");
	return h->write_field(layer, var, src);
}

int cmc_adi3d_step_prologue(cmc_adi3d *h) { H_CHECK(h); return h->step_prologue(); }

int cmc_adi3d_solve_direction(cmc_adi3d *h, int dir, double dt, int nl, int cur_layer, int next_layer)
{
	H_CHECK(h);
	return h->solve_direction(dir, dt, nl, cur_layer, next_layer);
}

int cmc_adi3d_eval_div_error(cmc_adi3d *h, int layer, double *err)
{
	H_CHECK(h);
	if (layer < 0 || layer > 3) return fail(CMC_ERR_INVALID, "eval_div_error: bad layer");
	return h->eval_div_error(layer, err);
}

int cmc_adi3d_field_sums(cmc_adi3d *h, int layer, double *sums8)
{
	H_CHECK(h);
	if (layer < 0 || layer > 3 || !sums8) return fail(CMC_ERR_INVALID, "field_sums: bad argument");
	return h->field_sums(layer, sums8);
}

int cmc_adi3d_stream(const cmc_adi3d *h, void **s)
{
	H_CHECK(h);
	if (!s) return fail(CMC_ERR_INVALID, "null argument");
	*s = (void *)h->stream;
	return CMC_OK;
}

int cmc_adi3d_launch_count(const cmc_adi3d *h, int64_t *n, int reset)
{
	H_CHECK(h);
	if (n) *n = h->launches;
	if (reset) const_cast<cmc_adi3d *>(h)->launches = 0;
	return CMC_OK;
}

int cmc_adi3d_get_timing(cmc_adi3d *h, int kind, double *total_ms, int64_t *calls)
{
	H_CHECK(h);
	if (kind < 0 || kind >= CMC_TIMING_KINDS) return fail(CMC_ERR_INVALID, "get_timing: bad kind");
	cudaSetDevice(h->device);
	h->spans_collect();
	if (total_ms) *total_ms = h->kind_ms[kind];
	if (calls) *calls = h->kind_calls[kind];
	return CMC_OK;
}

int cmc_adi3d_device_bytes(const cmc_adi3d *h, int64_t *n)
{
	H_CHECK(h);
	if (n) *n = h->dev_bytes;
	return CMC_OK;
}

int cmc_solve_tridiagonal_batch(int fp_bytes, int mode, int nsys, int n,
                                const void *a, const void *b, const void *c, const void *d, void *x)
{
	int device = 0;
	if (nsys < 1 || n < 2 || !a || !b || !c || !d || !x) return fail(CMC_ERR_INVALID, "solve_tridiagonal_batch: bad argument");
	if (fp_bytes != 4 && fp_bytes != 8) return fail(CMC_ERR_INVALID, "fp_bytes must be 4 or 8");
	int rc = check_device(device);
	if (rc) return rc;
	CU_TRY(cudaSetDevice(device));
	const size_t bytes = (size_t)nsys * n * fp_bytes;
	void *dv[5] = {};
	const void *hv[4] = {a, b, c, d};
	for (int i = 0; i < 5; i++) CU_TRY(cudaMalloc(&dv[i], bytes));
	for (int i = 0; i < 4; i++) CU_TRY(cudaMemcpy(dv[i], hv[i], bytes, cudaMemcpyHostToDevice));
	bool ok = true;
	if (mode == CMC_MODE_EXACT) {
		if (fp_bytes == 4) launch_thomas_batch<float>(nsys, n, (float *)dv[0], (float *)dv[1], (float *)dv[2], (float *)dv[3], (float *)dv[4], 0);
		else launch_thomas_batch<double>(nsys, n, (double *)dv[0], (double *)dv[1], (double *)dv[2], (double *)dv[3], (double *)dv[4], 0);
	} else {
		if (fp_bytes == 4) ok = launch_pcr_batch<float>(nsys, n, (float *)dv[0], (float *)dv[1], (float *)dv[2], (float *)dv[3], (float *)dv[4], 0);
		else ok = launch_pcr_batch<double>(nsys, n, (double *)dv[0], (double *)dv[1], (double *)dv[2], (double *)dv[3], (double *)dv[4], 0);
	}
	cudaError_t e = cudaDeviceSynchronize();
	if (e == cudaSuccess) e = cudaMemcpy(x, dv[4], bytes, cudaMemcpyDeviceToHost);
	for (int i = 0; i < 5; i++) cudaFree(dv[i]);
	if (!ok) return fail(CMC_ERR_UNSUPPORTED, "solve_tridiagonal_batch: size not supported by the fast line solver");
	if (e != cudaSuccess) return fail(CMC_ERR_CUDA, std::string("solve_tridiagonal_batch: ") + cudaGetErrorString(e));
	return CMC_OK;
}

} // extern "C"
