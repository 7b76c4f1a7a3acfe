// cmc_adi.cu - C ABI (include/cmc_adi.h) and host orchestration of the ADI time step.
//
// Host-side restatement of AdiSolver3D::{Init, CreateSegments, UpdateBoundaries, TimeStep,
// SolveDirection} (reference src/FluidSolver3D/AdiSolver3D.cpp:166-268, 286-391, 553-666) and
// Solver3D::GetLayer (Solver3D.cpp:21-25) on top of the sm_100a kernels.  Everything is ordered on
// ONE CUDA stream per handle; a time step does not synchronise with the host unless the caller
// asks for the residual.  There is no CPU fallback anywhere in this file.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <new>

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
	                      const void *vx, const void *vy, const void *vz, const void *T, size_t aos_stride) = 0;
	virtual int build_lines() = 0;
	virtual int update_boundaries() = 0;
	virtual int time_step(double dt, int ng, int nl, int ce, double *err, bool async) = 0;
	virtual int sync(double *err) = 0;
	virtual int get_layer(void *vel, double *T, int ox, int oy, int oz) = 0;
	virtual int read_field(int layer, int var, void *dst) = 0;
	virtual int write_field(int layer, int var, const void *src) = 0;
	virtual int step_prologue() = 0;
	virtual int solve_direction(int dir, double dt, int nl, int cur_layer, int next_layer) = 0;
	virtual int eval_div_error(int layer, double *err) = 0;

	int device = 0, fp = 8;
	int rank = 0, nranks = 1;
	Layout L{}, G{};
	cudaStream_t stream = nullptr;
	long long launches = 0;
	long long dev_bytes = 0;
	long long num_segs[3] = {0, 0, 0};
	long long shared_free[3] = {0, 0, 0};   // cells shared by two segments with a BC_FREE row (fast solver falls back)
	int mode = CMC_MODE_FAST;
	int fold_boundaries = 0;
	bool have_nodes = false, have_lines = false;
	DistContext *dist = nullptr;

	// optional per-kernel-kind device timing (cmc_adi3d_set_option "profile"): CUDA event pairs on `stream`
	int profile = 0;
	struct Span { int kind; cudaEvent_t a, b; };
	std::vector<Span> spans;
	double kind_ms[CMC_TIMING_KINDS] = {};
	long long kind_calls[CMC_TIMING_KINDS] = {};
	void span_begin(int kind)
	{
		if (!profile) return;
		Span sp; sp.kind = kind;
		cudaEventCreate(&sp.a); cudaEventCreate(&sp.b);
		cudaEventRecord(sp.a, stream);
		spans.push_back(sp);
	}
	void span_end()
	{
		if (!profile || spans.empty()) return;
		cudaEventRecord(spans.back().b, stream);
	}
	void spans_collect()
	{
		if (spans.empty()) return;
		cudaStreamSynchronize(stream);
		for (auto &sp : spans) {
			float ms = 0.f;
			if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { kind_ms[sp.kind] += ms; kind_calls[sp.kind]++; }
			cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);
		}
		spans.clear();
	}
};

namespace {

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <typename FT>
struct Solver : cmc_adi3d {
	FT *field[5][4] = {};        // physical buffers: four layers + the spare linearisation buffer
	int slot[4] = {0, 1, 2, 3};  // logical layer (CMC_LAYER_*) -> physical buffer
	int spare = 4;
	FT *nodev[4] = {};
	uint8_t *role[3] = {};
	FT *cv = nullptr, *cT = nullptr;
	double *d_partials = nullptr, *d_err2 = nullptr, *h_err2 = nullptr;
	unsigned long long *d_segcount = nullptr;
	FT *d_outvel = nullptr;
	double *d_outT = nullptr;
	size_t out_cap = 0;
	cmc_fluid_params params{};
	double dx = 0, dy = 0, dz = 0;
	double diffError = 0.0;
	bool err_pending = false;
	static const int kMaxErrBlocks = 148 * 8;

	~Solver() override
	{
		cudaSetDevice(device);
		if (stream) cudaStreamSynchronize(stream);
		spans_collect();
		for (auto &l : field) for (auto &p : l) if (p) cudaFree(p);
		for (auto &p : nodev) if (p) cudaFree(p);
		for (auto &p : role) if (p) cudaFree(p);
		if (cv) cudaFree(cv);
		if (cT) cudaFree(cT);
		if (d_partials) cudaFree(d_partials);
		if (d_err2) cudaFree(d_err2);
		if (h_err2) cudaFreeHost(h_err2);
		if (d_segcount) cudaFree(d_segcount);
		if (d_outvel) cudaFree(d_outvel);
		if (d_outT) cudaFree(d_outT);
		if (ncode) cudaFree(ncode);
		if (dist) dist_destroy(dist);
		if (stream) cudaStreamDestroy(stream);
	}

	template <typename T>
	int dalloc(T *&p, size_t count)
	{
		CU_TRY(cudaMalloc((void **)&p, count * sizeof(T)));
		CU_TRY(cudaMemsetAsync(p, 0, count * sizeof(T), stream));
		dev_bytes += (long long)(count * sizeof(T));
		return CMC_OK;
	}

	int init(const cmc_grid_desc *g, const cmc_fluid_params *p, int x0, int nx)
	{
		CU_TRY(cudaSetDevice(device));
		CU_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
		params = *p;
		dx = g->dx; dy = g->dy; dz = g->dz;
		G.nx = g->dimx; G.ny = g->dimy; G.nz = g->dimz; G.gx = g->dimx; G.x0 = 0;
		G.nzp = round_up(g->dimz, 16); G.plane = (long long)G.ny * G.nzp; G.total = (long long)(G.nx + 2) * G.plane;
		L = G; L.nx = nx; L.x0 = x0; L.total = (long long)(nx + 2) * L.plane;
		int rc;
		for (int l = 0; l < 5; l++)
			for (int q = 0; q < 4; q++)
				if ((rc = dalloc(field[l][q], (size_t)L.total))) return rc;
		for (int q = 0; q < 4; q++) if ((rc = dalloc(nodev[q], (size_t)L.total))) return rc;
		for (int d = 0; d < 3; d++) if ((rc = dalloc(role[d], (size_t)L.total))) return rc;
		if ((rc = dalloc(cv, (size_t)L.total))) return rc;
		if ((rc = dalloc(cT, (size_t)L.total))) return rc;
		if ((rc = dalloc(d_partials, (size_t)2 * kMaxErrBlocks))) return rc;
		if ((rc = dalloc(d_err2, 2))) return rc;
		if ((rc = dalloc(d_segcount, 8))) return rc;
		CU_TRY(cudaHostAlloc((void **)&h_err2, 2 * sizeof(double), cudaHostAllocDefault));
		h_err2[0] = h_err2[1] = 0.0;
		CU_TRY(cudaStreamSynchronize(stream));
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

	// dense host (global grid) <-> padded device slab
	int upload_dense(FT *dst_field, const FT *src_global_dense)
	{
		const FT *src = src_global_dense + (size_t)L.x0 * L.ny * L.nz;
		CU_TRY(cudaMemcpy2DAsync(dst_field + L.idx(0, 0, 0), sizeof(FT) * L.nzp, src, sizeof(FT) * L.nz,
		                         sizeof(FT) * L.nz, (size_t)L.nx * L.ny, cudaMemcpyHostToDevice, stream));
		return CMC_OK;
	}

	int set_nodes(const int32_t *type, const int32_t *bc_vel, const int32_t *bc_temp,
	              const void *vx, const void *vy, const void *vz, const void *T, size_t aos_stride) override
	{
		CU_TRY(cudaSetDevice(device));
		const size_t N = (size_t)G.nx * G.ny * G.nz;
		std::vector<uint8_t> code(N);
		std::vector<FT> tmp[4];
		const FT *src[4] = {(const FT *)vx, (const FT *)vy, (const FT *)vz, (const FT *)T};
		long long face_in = 0;
		if (aos_stride) {
			// reference Node (Grid3D.h:73-88): {int type; int bc_vel; int bc_temp; FTYPE v[3]; FTYPE T}
			const char *base = (const char *)type;
			for (int q = 0; q < 4; q++) tmp[q].resize(N);
			for (size_t id = 0; id < N; id++) {
				const int32_t *hd = (const int32_t *)(base + id * aos_stride);
				const FT *fv = (const FT *)(base + id * aos_stride + (sizeof(FT) == 8 ? 16 : 12));
				code[id] = (uint8_t)((hd[0] & 3) | (hd[1] == CMC_BC_FREE ? 4 : 0) | (hd[2] == CMC_BC_FREE ? 8 : 0));
				tmp[0][id] = fv[0]; tmp[1][id] = fv[1]; tmp[2][id] = fv[2]; tmp[3][id] = fv[3];
			}
			for (int q = 0; q < 4; q++) src[q] = tmp[q].data();
		} else {
			for (size_t id = 0; id < N; id++) {
				if (type[id] < 0 || type[id] > 3) return fail(CMC_ERR_INVALID, "set_nodes: node type out of range");
				code[id] = (uint8_t)((type[id] & 3) | (bc_vel[id] == CMC_BC_FREE ? 4 : 0) | (bc_temp[id] == CMC_BC_FREE ? 8 : 0));
			}
		}
		// NODE_IN cells on a domain face make the reference read out of bounds (stencils, EvalDivError):
		// the guard planes / line padding keep our accesses in bounds, values there are unspecified.
		for (int i = 0; i < G.nx; i++)
			for (int j = 0; j < G.ny; j++)
				for (int k = 0; k < G.nz; k++)
					if (i == 0 || j == 0 || k == 0 || i == G.nx - 1 || j == G.ny - 1 || k == G.nz - 1)
						if ((code[((size_t)i * G.ny + j) * G.nz + k] & 3) == CMC_NODE_IN) face_in++;
		(void)face_in;
		if (ncode) { cudaFree(ncode); ncode = nullptr; }
		CU_TRY(cudaMalloc((void **)&ncode, N));
		CU_TRY(cudaMemcpyAsync(ncode, code.data(), N, cudaMemcpyHostToDevice, stream));
		int rc;
		for (int q = 0; q < 4; q++) {
			CU_TRY(cudaMemsetAsync(nodev[q], 0, sizeof(FT) * (size_t)L.total, stream));
			if ((rc = upload_dense(nodev[q], src[q]))) return rc;
		}
		// halo planes of the node values (neighbour slabs) so that layers start with valid halos
		if (nranks > 1) {
			for (int q = 0; q < 4; q++) {
				if (L.x0 > 0)
					CU_TRY(cudaMemcpy2DAsync(nodev[q] + L.idx(-1, 0, 0), sizeof(FT) * L.nzp, src[q] + (size_t)(L.x0 - 1) * L.ny * L.nz,
					                         sizeof(FT) * L.nz, sizeof(FT) * L.nz, (size_t)L.ny, cudaMemcpyHostToDevice, stream));
				if (L.x0 + L.nx < G.nx)
					CU_TRY(cudaMemcpy2DAsync(nodev[q] + L.idx(L.nx, 0, 0), sizeof(FT) * L.nzp, src[q] + (size_t)(L.x0 + L.nx) * L.ny * L.nz,
					                         sizeof(FT) * L.nz, sizeof(FT) * L.nz, (size_t)L.ny, cudaMemcpyHostToDevice, stream));
			}
		}
		// cur = TimeLayer3D(grid) (TimeLayer3D.h:734-751); half/next/temp are uninitialised in the reference
		// (TimeLayer3D.h:353) and are defined here as copies of cur.
		for (int l = 0; l < 5; l++)
			for (int q = 0; q < 4; q++)
				CU_TRY(cudaMemcpyAsync(field[l][q], nodev[q], sizeof(FT) * (size_t)L.total, cudaMemcpyDeviceToDevice, stream));
		slot[0] = 0; slot[1] = 1; slot[2] = 2; slot[3] = 3; spare = 4;
		CU_TRY(cudaStreamSynchronize(stream));
		have_nodes = true; have_lines = false;
		diffError = 0.0; err_pending = false;
		return CMC_OK;
	}
	uint8_t *ncode = nullptr;

	int build_lines() override
	{
		if (!have_nodes) return fail(CMC_ERR_INVALID, "build_lines: call cmc_adi3d_set_nodes first");
		CU_TRY(cudaSetDevice(device));
		CU_TRY(cudaMemsetAsync(d_segcount, 0, 8 * sizeof(unsigned long long), stream));
		for (int d = 0; d < 3; d++) CU_TRY(cudaMemsetAsync(role[d], 0, (size_t)L.total, stream));
		launch_role_type_bits(G, ncode, L, role[0], role[1], role[2], stream, &launches);
		for (int d = 0; d < 3; d++) launch_build_roles(d, G, ncode, L, role[d], d_segcount + d, stream, &launches);
		unsigned long long h[8];
		CU_TRY(cudaMemcpyAsync(h, d_segcount, sizeof h, cudaMemcpyDeviceToHost, stream));
		CU_TRY(cudaStreamSynchronize(stream));
		CU_TRY(cudaGetLastError());
		for (int d = 0; d < 3; d++) { num_segs[d] = (long long)h[d]; shared_free[d] = (long long)h[4 + d]; }
		if (nranks > 1 && dist) {
			int rc = dist_sum_i64(dist, &num_segs[1], 2, stream);   // Y/Z counted per slab, X counted globally
			if (rc) return fail(CMC_ERR_COMM, dist_error());
		}
		have_lines = true;
		return CMC_OK;
	}

	int update_boundaries() override
	{
		if (!have_lines) return fail(CMC_ERR_INVALID, "update_boundaries: call cmc_adi3d_build_lines first");
		CU_TRY(cudaSetDevice(device));
		ConstLayerPtrs<FT> nv; for (int q = 0; q < 4; q++) nv.f[q] = nodev[q];
		span_begin(CMC_TIMING_BOUNDARY);
		launch_update_boundaries<FT>(L, role[2], nv, layer(CMC_LAYER_CUR), stream, &launches);
		span_end();
		return CMC_OK;
	}

	SweepArgs<FT> sweep_args(int dir, FT dt, int cur_layer, int next_layer)
	{
		SweepArgs<FT> A;
		A.L = L; A.dt = dt;
		A.h[0] = (FT)dx; A.h[1] = (FT)dy; A.h[2] = (FT)dz;
		A.v_T = (FT)params.v_T; A.v_vis = (FT)params.v_vis; A.t_vis = (FT)params.t_vis; A.t_phi = (FT)params.t_phi;
		A.role = role[dir];
		for (int q = 0; q < 4; q++) {
			A.cur[q] = field[slot[cur_layer]][q];
			A.temp[q] = field[slot[CMC_LAYER_TEMP]][q];
			A.next[q] = field[slot[next_layer]][q];
			A.temp_out[q] = field[spare][q];
			A.nodev[q] = nodev[q];
		}
		A.cv = cv; A.cT = cT;
		return A;
	}

	// AdiSolver3D::SolveDirection (AdiSolver3D.cpp:564-666): num_local x { solve every line for u,v,w,T ; merge }
	int solve_direction_impl(int dir, FT dt, int nl, int cur_layer, int next_layer)
	{
		for (int it = 0; it < nl; it++) {
			if (nranks > 1) {
				int rc = dist_halo_exchange<FT>(dist, L, field[slot[CMC_LAYER_TEMP]], stream, &launches);
				if (rc) return fail(CMC_ERR_COMM, dist_error());
			}
			SweepArgs<FT> A = sweep_args(dir, dt, cur_layer, next_layer);
			bool done = false;
			if (nranks > 1 && dir == CMC_DIR_X) {
				int rc = dist_sweep_x<FT>(dist, A, stream, &launches);
				if (rc) return fail(CMC_ERR_COMM, dist_error());
				std::swap(slot[CMC_LAYER_TEMP], spare);
				done = true;
			}
			if (!done && mode == CMC_MODE_FAST && shared_free[dir] == 0) {
				span_begin(CMC_TIMING_SWEEP_X + dir);
				done = launch_fast_sweep<FT>(dir, A, stream, &launches);
				span_end();
				if (done) std::swap(slot[CMC_LAYER_TEMP], spare);      // merged temp went to the other buffer
				else if (profile) spans.pop_back();
			}
			if (!done) {
				span_begin(CMC_TIMING_SWEEP_X + dir);
				launch_exact_sweep<FT>(dir, A, stream, &launches);
				span_end();
				span_begin(CMC_TIMING_MERGE);
				launch_merge<FT>(L, role[dir], clayer(next_layer), layer(CMC_LAYER_TEMP), stream, &launches);
				span_end();
			}
		}
		return CMC_OK;
	}

	int step_prologue() override
	{
		if (!have_lines) return fail(CMC_ERR_INVALID, "time_step: call cmc_adi3d_build_lines first");
		CU_TRY(cudaSetDevice(device));
		// cur -> next on BOUND and VALVE cells (AdiSolver3D.cpp:310-311); temp <- cur (:320)
		span_begin(CMC_TIMING_COPY);
		launch_copy_masked<FT>(L, role[2], R_BV, clayer(CMC_LAYER_CUR), layer(CMC_LAYER_NEXT), stream, &launches);
		launch_copy_full<FT>(L, clayer(CMC_LAYER_CUR), layer(CMC_LAYER_TEMP), stream, &launches);
		span_end();
		return CMC_OK;
	}

	int enqueue_div_error(int logical_layer)
	{
		if (nranks > 1) {
			int rc = dist_halo_exchange<FT>(dist, L, field[slot[logical_layer]], stream, &launches);
			if (rc) return fail(CMC_ERR_COMM, dist_error());
		}
		const int l = slot[logical_layer];
		launch_div_error<FT>(L, role[2], field[l][0], field[l][1], field[l][2], (FT)dx, (FT)dy, (FT)dz,
		                     d_partials, kMaxErrBlocks, d_err2, stream, &launches);
		if (nranks > 1) {
			int rc = dist_allreduce_f64(dist, d_err2, 2, stream);
			if (rc) return fail(CMC_ERR_COMM, dist_error());
		}
		CU_TRY(cudaMemcpyAsync(h_err2, d_err2, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
		err_pending = true;
		return CMC_OK;
	}

	int fetch_error()
	{
		if (err_pending) {
			CU_TRY(cudaStreamSynchronize(stream));
			diffError = h_err2[0] / h_err2[1];      // err / count (TimeLayer3D.h:639); 0/0 = NaN like the reference
			err_pending = false;
		}
		return CMC_OK;
	}

	// AdiSolver3D::TimeStep (AdiSolver3D.cpp:306-391)
	int time_step(double dt_in, int ng, int nl, int ce, double *err, bool async) override
	{
		int rc;
		if (ng < 0 || nl < 0) return fail(CMC_ERR_INVALID, "time_step: negative iteration count");
		const FT dt = (FT)dt_in;                                   // FluidSolver3D.cpp:242 casts to FTYPE
		if ((rc = step_prologue())) return rc;
		for (int it = 0; it < ng; it++) {                          // :335-358
			if ((rc = solve_direction_impl(CMC_DIR_Z, dt, nl, CMC_LAYER_CUR, CMC_LAYER_NEXT))) return rc;
			if ((rc = solve_direction_impl(CMC_DIR_Y, dt, nl, CMC_LAYER_NEXT, CMC_LAYER_HALF))) return rc;
			if ((rc = solve_direction_impl(CMC_DIR_X, dt, nl, CMC_LAYER_HALF, CMC_LAYER_NEXT))) return rc;
			// update non-linear layer once more (:354): temp = (temp + next) / 2 on NODE_IN
			span_begin(CMC_TIMING_MERGE);
			launch_merge<FT>(L, role[2], clayer(CMC_LAYER_NEXT), layer(CMC_LAYER_TEMP), stream, &launches);
			span_end();
		}
		if (ce) {
			span_begin(CMC_TIMING_RESIDUAL);
			rc = enqueue_div_error(CMC_LAYER_NEXT);
			span_end();
			if (rc) return rc;
		}
		if (!async) {
			if ((rc = fetch_error())) return rc;
			CU_TRY(cudaGetLastError());
			if (err) *err = diffError;
			if (diffError > CMC_ERR_THRESHOLD) {                   // :371-374 (layers are not swapped)
				char buf[96];
				snprintf(buf, sizeof buf, "Error is too big! %f", diffError);
				return fail(CMC_ERR_DIVERGED, buf);
			}
		}
		std::swap(slot[CMC_LAYER_CUR], slot[CMC_LAYER_NEXT]);      // :388-390
		return CMC_OK;
	}

	int sync(double *err) override
	{
		CU_TRY(cudaSetDevice(device));
		int rc = fetch_error();
		if (rc) return rc;
		CU_TRY(cudaStreamSynchronize(stream));
		CU_TRY(cudaGetLastError());
		if (err) *err = diffError;
		if (diffError > CMC_ERR_THRESHOLD) {
			char buf[96];
			snprintf(buf, sizeof buf, "Error is too big! %f", diffError);
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
		int rc = solve_direction_impl(dir, (FT)dt, nl, cur_layer, next_layer);
		if (rc) return rc;
		CU_TRY(cudaStreamSynchronize(stream));
		CU_TRY(cudaGetLastError());
		return CMC_OK;
	}

	int eval_div_error(int logical, double *err) override
	{
		if (!have_lines) return fail(CMC_ERR_INVALID, "eval_div_error: call cmc_adi3d_build_lines first");
		CU_TRY(cudaSetDevice(device));
		int rc = enqueue_div_error(logical);
		if (rc) return rc;
		CU_TRY(cudaStreamSynchronize(stream));
		err_pending = false;
		if (err) *err = h_err2[0] / h_err2[1];
		return CMC_OK;
	}

	// Solver3D::GetLayer (Solver3D.cpp:21-25)
	int get_layer(void *vel, double *T, int ox, int oy, int oz) override
	{
		if (!have_lines) return fail(CMC_ERR_INVALID, "get_layer: call cmc_adi3d_build_lines first");
		CU_TRY(cudaSetDevice(device));
		if (ox == 0) ox = G.nx;
		if (oy == 0) oy = G.ny;
		if (oz == 0) oz = G.nz;
		if (ox < 0 || oy < 0 || oz < 0) return fail(CMC_ERR_INVALID, "get_layer: negative output dims");
		span_begin(CMC_TIMING_READBACK);
		launch_clear_out<FT>(L, role[2], layer(CMC_LAYER_NEXT), (FT)CMC_MISSING_VALUE, stream, &launches);
		// output rows i whose source plane x = i*dimx/outdimx lies in this slab
		int oi0 = ox, oi1 = 0;
		for (int i = 0; i < ox; i++) {
			const int x = (int)((long long)i * G.nx / ox);
			if (x >= L.x0 && x < L.x0 + L.nx) { if (i < oi0) oi0 = i; if (i + 1 > oi1) oi1 = i + 1; }
		}
		const size_t outN = (size_t)ox * oy * oz;
		if (outN > out_cap) {
			if (d_outvel) cudaFree(d_outvel);
			if (d_outT) cudaFree(d_outT);
			d_outvel = nullptr; d_outT = nullptr;
			CU_TRY(cudaMalloc((void **)&d_outvel, outN * 3 * sizeof(FT)));
			CU_TRY(cudaMalloc((void **)&d_outT, outN * sizeof(double)));
			out_cap = outN;
		}
		launch_filter<FT>(L, clayer(CMC_LAYER_NEXT), ox, oy, oz, oi0, oi1, d_outvel, d_outT, stream, &launches);
		span_end();
		if (oi1 > oi0) {
			const size_t o0 = (size_t)oi0 * oy * oz, cnt = (size_t)(oi1 - oi0) * oy * oz;
			if (nranks == 1 || rank == 0) {
				CU_TRY(cudaMemcpyAsync((FT *)vel + 3 * o0, d_outvel + 3 * o0, cnt * 3 * sizeof(FT), cudaMemcpyDeviceToHost, stream));
				CU_TRY(cudaMemcpyAsync(T + o0, d_outT + o0, cnt * sizeof(double), cudaMemcpyDeviceToHost, stream));
			}
		}
		if (nranks > 1) {
			int rc = dist_gather_layer<FT>(dist, G, ox, oy, oz, d_outvel, d_outT, oi0, oi1, (FT *)vel, T, stream);
			if (rc) return fail(CMC_ERR_COMM, dist_error());
		}
		CU_TRY(cudaStreamSynchronize(stream));
		CU_TRY(cudaGetLastError());
		return CMC_OK;
	}

	int read_field(int logical, int var, void *dst) override
	{
		if (logical < 0 || logical > 3 || var < 0 || var > 3) return fail(CMC_ERR_INVALID, "read_field: bad layer/var");
		CU_TRY(cudaSetDevice(device));
		const FT *src = field[slot[logical]][var] + L.idx(0, 0, 0);
		CU_TRY(cudaMemcpy2DAsync(dst, sizeof(FT) * L.nz, src, sizeof(FT) * L.nzp, sizeof(FT) * L.nz, (size_t)L.nx * L.ny,
		                         cudaMemcpyDeviceToHost, stream));
		CU_TRY(cudaStreamSynchronize(stream));
		return CMC_OK;
	}

	int write_field(int logical, int var, const void *src) override
	{
		if (logical < 0 || logical > 3 || var < 0 || var > 3) return fail(CMC_ERR_INVALID, "write_field: bad layer/var");
		CU_TRY(cudaSetDevice(device));
		FT *dst = field[slot[logical]][var] + L.idx(0, 0, 0);
		CU_TRY(cudaMemcpy2DAsync(dst, sizeof(FT) * L.nzp, src, sizeof(FT) * L.nz, sizeof(FT) * L.nz, (size_t)L.nx * L.ny,
		                         cudaMemcpyHostToDevice, stream));
		CU_TRY(cudaStreamSynchronize(stream));
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

static int create_impl(const cmc_grid_desc *grid, const cmc_fluid_params *params, int fp_bytes, int device,
                       int rank, int nranks, const void *nccl_id, cmc_adi3d **out)
{
	if (!grid || !params || !out) return fail(CMC_ERR_INVALID, "create: null argument");
	*out = nullptr;
	if (fp_bytes != 4 && fp_bytes != 8) return fail(CMC_ERR_INVALID, "create: fp_bytes must be 4 or 8");
	if (grid->dimx < 3 || grid->dimy < 3 || grid->dimz < 3) return fail(CMC_ERR_INVALID, "create: every grid dimension must be >= 3");
	if (!(grid->dx > 0) || !(grid->dy > 0) || !(grid->dz > 0)) return fail(CMC_ERR_INVALID, "create: grid spacing must be positive");
	if (nranks < 1 || rank < 0 || rank >= nranks) return fail(CMC_ERR_INVALID, "create: bad rank / nranks");
	if (nranks > grid->dimx) return fail(CMC_ERR_INVALID, "create: more ranks than x-planes");
	int rc = check_device(device);
	if (rc) return rc;
	// GPUplan::splitEven1D (reference GPUplan.cpp:122-141): dimx / n planes each, remainder spread over the first ranks
	int x0 = 0, nx = grid->dimx;
	if (nranks > 1) {
		const int base = grid->dimx / nranks, rem = grid->dimx % nranks;
		nx = base + (rank < rem ? 1 : 0);
		x0 = rank * base + (rank < rem ? rank : rem);
	}
	cmc_adi3d *h = nullptr;
	if (fp_bytes == 4) {
		auto *s = new (std::nothrow) Solver<float>();
		if (!s) return fail(CMC_ERR_INVALID, "out of host memory");
		s->device = device; s->fp = 4; s->rank = rank; s->nranks = nranks;
		rc = s->init(grid, params, x0, nx);
		h = s;
	} else {
		auto *s = new (std::nothrow) Solver<double>();
		if (!s) return fail(CMC_ERR_INVALID, "out of host memory");
		s->device = device; s->fp = 8; s->rank = rank; s->nranks = nranks;
		rc = s->init(grid, params, x0, nx);
		h = s;
	}
	if (rc) { delete h; return rc; }
	if (nranks > 1) {
		h->dist = dist_create(device, rank, nranks, nccl_id, h->L, fp_bytes, h->stream);
		if (!h->dist) { delete h; return fail(CMC_ERR_COMM, dist_error()); }
	}
	*out = h;
	return CMC_OK;
}

} // namespace

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
	return create_impl(grid, params, fp_bytes, device, 0, 1, nullptr, out);
}

int cmc_nccl_unique_id(void *id128)
{
	if (!id128) return fail(CMC_ERR_INVALID, "null id");
	if (dist_unique_id(id128)) return fail(CMC_ERR_COMM, dist_error());
	return CMC_OK;
}

int cmc_adi3d_create_dist(const cmc_grid_desc *grid, const cmc_fluid_params *params, int fp_bytes, int device,
                          int rank, int nranks, const void *nccl_unique_id, cmc_adi3d **out)
{
	if (nranks > 1 && !nccl_unique_id) return fail(CMC_ERR_INVALID, "create_dist: nccl_unique_id required when nranks > 1");
	return create_impl(grid, params, fp_bytes, device, rank, nranks, nccl_unique_id, out);
}

int cmc_adi3d_destroy(cmc_adi3d *h)
{
	delete h;
	return CMC_OK;
}

#define H_CHECK(h) if (!(h)) return fail(CMC_ERR_INVALID, "null handle")

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
	return h->set_nodes(type, bc_vel, bc_temp, vx, vy, vz, T, 0);
}

int cmc_adi3d_set_nodes_aos(cmc_adi3d *h, const void *nodes, size_t stride)
{
	H_CHECK(h);
	if (!nodes) return fail(CMC_ERR_INVALID, "set_nodes_aos: null array");
	if (stride < (size_t)((h->fp == 8 ? 16 : 12) + 4 * h->fp)) return fail(CMC_ERR_INVALID, "set_nodes_aos: stride smaller than a Node");
	return h->set_nodes((const int32_t *)nodes, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stride);
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
	if ((h->nranks == 1 || h->rank == 0) && (!vel || !T)) return fail(CMC_ERR_INVALID, "get_layer: null output");
	return h->get_layer(vel, T, ox, oy, oz);
}

int cmc_adi3d_set_option(cmc_adi3d *h, const char *key, int64_t value)
{
	H_CHECK(h);
	if (!key) return fail(CMC_ERR_INVALID, "set_option: null key");
	if (!strcmp(key, "mode")) {
		if (value != CMC_MODE_FAST && value != CMC_MODE_EXACT) return fail(CMC_ERR_INVALID, "set_option: unknown mode");
		h->mode = (int)value;
		return CMC_OK;
	}
	if (!strcmp(key, "fold_boundaries")) { h->fold_boundaries = value != 0; return CMC_OK; }
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
	if (!strcmp(key, "fold_boundaries")) { *value = h->fold_boundaries; return CMC_OK; }
	if (!strcmp(key, "nzp")) { *value = h->L.nzp; return CMC_OK; }
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
