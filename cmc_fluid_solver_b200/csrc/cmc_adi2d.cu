// cmc_adi2d.cu - the 2D ADI time step (reference src/FluidSolver2D/AdiSolver2D.cpp:74-323, TimeLayer2D.h,
// Solver2D.cpp) behind the C ABI of include/cmc_adi.h (cmc_adi2d_*).
//
// The 2D cases of the reference are tiny (data/2D/box_pipe: 120 x 135 cells, 3 variables), so one time step is ONE
// kernel launch by ONE thread block: every phase of AdiSolver2D::TimeStep - segment scan, layer copies, the
// while-not-converged outer loop with its two directional solves per iteration, the residual and the final clear /
// copy - runs inside the kernel separated by block barriers, on layers that stay in L2.  The directional solve uses
// one thread per (segment, variable) with a sequential Thomas in the reference's operation order; this translation
// unit is compiled with -fmad=false, and the residual is accumulated in FTYPE in the reference's cell order by one
// thread, so fields AND the data-dependent iteration count are bit-identical with the reference CPU solver.
// (Batching many 2D cases per launch is the throughput path and comes next - SURVEY 8(f).)
//
// The grid (Grid2D::GetType / GetData) is an input refreshed by the host before every step, because the reference
// driver calls grid.Prepare(t) per step (FluidSolver2D.cpp:129).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <new>

#include "../../include/cmc_adi.h"
#include <cuda_runtime.h>

extern int cmc_set_error(int code, const std::string &msg);   // cmc_adi.cu

namespace {

enum { L2_CUR = 0, L2_HALF = 1, L2_NEXT = 2, L2_TEMP = 3, L2_NEXT_LOCAL = 4, L2_TEMP_LOCAL = 5, L2_COUNT = 6 };
constexpr double kErrThreshold2D = 0.1;      // AdiSolver2D.h:24
constexpr int kMaxGlobalIters2D = 100;       // AdiSolver2D.h:25

struct Seg2 { int pos, end, valid; };        // first / last cell along the line (AdiSolver2D.h:32-38); valid = row has a segment

template <typename FT>
struct Dev2 {
	int dimx, dimy;
	FT dx, dy;             // TimeLayer2D members ((FTYPE)grid->dx)
	FT gdx, gdy;           // (FTYPE)grid->dx as used by BuildMatrix
	FT v_T, v_vis, t_vis, t_phi;
	FT startT;
	const int *type, *bc;
	const FT *gvx, *gvy, *gT;
	FT *f[L2_COUNT][3];
	Seg2 *listX, *listY;   // listX[i]: segment of column i (direction Y); listY[j]: segment of row j (direction X)
	FT *scratch;           // per (segment, variable): a, b, c, d, x of max(dimx, dimy) rows
	FT *resid;             // per cell: its term of the divergence residual (0 where the reference skips the cell)
	FT dt;                 // time step of the batched launch (cmc_adi2d_time_step_batch)
	double *result;        // [0] err, [1] iterations, [2] status (0 ok, 1 exceeded MAX_GLOBAL_ITERS, 2 error too big)
};

#define ID2(i, j) ((i) * P.dimy + (j))

// TimeLayer2D::Copy*to / Merge*to (TimeLayer2D.h:104-166): cells i < dimx-1, j < dimy-1 of one node type (-1: all four)
template <typename FT>
__device__ void copy_cells(const Dev2<FT> &P, int src, int dst, int type, bool merge)
{
	const int n = (P.dimx - 1) * (P.dimy - 1);
	for (int t = threadIdx.x; t < n; t += blockDim.x) {
		const int i = t / (P.dimy - 1), j = t % (P.dimy - 1), id = ID2(i, j);
		if (type >= 0 && P.type[id] != type) continue;
		for (int q = 0; q < 3; q++)
			P.f[dst][q][id] = merge ? (P.f[dst][q][id] + P.f[src][q][id]) / 2 : P.f[src][q][id];
	}
	__syncthreads();
}

// TimeLayer2D::EvalDivError (TimeLayer2D.h:88-102).  The reference adds the cells' terms into one FTYPE accumulator in
// row-major order, and floating-point addition does not commute with a tree: the terms (8 loads, 2 type tests and a dozen
// operations per cell - the expensive part) are computed by the whole block, the additions are then done by one thread
// in the reference's order.  Skipped cells contribute +0.0, which leaves an FTYPE sum of non-negative terms unchanged.
template <typename FT>
__device__ double eval_div_error(const Dev2<FT> &P, int l, double *sh)
{
	const FT *U = P.f[l][0], *V = P.f[l][1];
	const int ncell = (P.dimx - 1) * (P.dimy - 1);
	int mine = 0;
	for (int t = threadIdx.x; t < ncell; t += blockDim.x) {
		const int i = t / (P.dimy - 1), j = t % (P.dimy - 1);
		FT term = 0.0;
		if (P.type[ID2(i, j)] == CMC_NODE_IN && P.type[ID2(i + 1, j)] == CMC_NODE_IN && P.type[ID2(i, j + 1)] == CMC_NODE_IN && P.type[ID2(i + 1, j + 1)] == CMC_NODE_IN) {
			const FT tx = P.dy * (U[ID2(i + 1, j)] - U[ID2(i, j)]) + (U[ID2(i + 1, j + 1)] - U[ID2(i, j + 1)]) / 2;
			const FT ty = P.dx * (V[ID2(i, j + 1)] - V[ID2(i, j)]) + (V[ID2(i + 1, j + 1)] - V[ID2(i + 1, j)]) / 2;
			const FT s = tx + ty;
			term = s < 0 ? -s : s;
			mine++;
		}
		P.resid[t] = term;
	}
	const int count = __syncthreads_count(0) + 0;      // (barrier: the terms are visible)
	(void)count;
	__shared__ int s_count;
	if (threadIdx.x == 0) s_count = 0;
	__syncthreads();
	if (mine) atomicAdd(&s_count, mine);
	__syncthreads();
	if (threadIdx.x == 0) {
		FT err = 0.0;
		int t = 0;
		for (; t + 8 <= ncell; t += 8) {              // loads first, then the eight dependent additions
			const FT a0 = P.resid[t], a1 = P.resid[t + 1], a2 = P.resid[t + 2], a3 = P.resid[t + 3];
			const FT a4 = P.resid[t + 4], a5 = P.resid[t + 5], a6 = P.resid[t + 6], a7 = P.resid[t + 7];
			err += a0; err += a1; err += a2; err += a3; err += a4; err += a5; err += a6; err += a7;
		}
		for (; t < ncell; t++) err += P.resid[t];
		*sh = err / s_count;
	}
	__syncthreads();
	const double e = *sh;
	__syncthreads();
	return e;
}

// AdiSolver2D::SolveSegment (AdiSolver2D.cpp:180-203): ApplyBC0, BuildMatrix, ApplyBC1, SolveTridiagonal, UpdateSegment
template <typename FT>
__device__ void solve_segment(const Dev2<FT> &P, FT dt, int line, const Seg2 &sg, int var, int dirX, int cur, FT *w, int maxn)
{
	const int n = sg.end - sg.pos + 1;
	FT *a = w, *b = w + maxn, *c = w + 2 * maxn, *d = w + 3 * maxn, *x = w + 4 * maxn;
	const int tl = L2_TEMP_LOCAL;
	const int i0 = dirX ? sg.pos : line, j0 = dirX ? line : sg.pos;
	const FT *TU = P.f[tl][0], *TV = P.f[tl][1], *TT = P.f[tl][2];
	{   // ApplyBC0 (:74-95)
		const int id = ID2(i0, j0);
		if (P.bc[id] == CMC_BC_NOSLIP) { b[0] = 1.0; c[0] = 0.0; d[0] = var == 0 ? P.gvx[id] : var == 1 ? P.gvy[id] : P.gT[id]; }
		else { b[0] = 1.0; c[0] = -1.0; d[0] = 0.0; }
	}
	const FT h = dirX ? P.gdx : P.gdy;
	const FT vis = (var == 2 ? P.t_vis : P.v_vis) / (h * h);
	for (int p = 1; p < n - 1; p++) {   // BuildMatrix (:118-178)
		const int i = dirX ? i0 + p : i0, j = dirX ? j0 : j0 + p, id = ID2(i, j);
		const FT vel = dirX ? TU[id] : TV[id];
		a[p] = -vel / (2 * h) - vis;
		b[p] = 1 / dt + 2 * vis;
		c[p] = vel / (2 * h) - vis;
		if (var == 2) {
			// DissFuncX / DissFuncY (TimeLayer2D.h:64-84)
			const FT ux = (TU[ID2(i + 1, j)] - TU[ID2(i - 1, j)]) / (2 * P.dx), vx = (TV[ID2(i + 1, j)] - TV[ID2(i - 1, j)]) / (2 * P.dx);
			const FT uy = (TU[ID2(i, j + 1)] - TU[ID2(i, j - 1)]) / (2 * P.dy), vy = (TV[ID2(i, j + 1)] - TV[ID2(i, j - 1)]) / (2 * P.dy);
			const FT diss = dirX ? 2 * ux * ux + vx * vx + uy * vx : uy * uy + 2 * vy * vy + vx * uy;
			d[p] = P.f[cur][2][id] / dt + P.t_phi * diss;
		} else if ((var == 0) == (dirX != 0)) {
			const FT grad = dirX ? (TT[ID2(i + 1, j)] - TT[ID2(i - 1, j)]) / (2 * P.dx) : (TT[ID2(i, j + 1)] - TT[ID2(i, j - 1)]) / (2 * P.dy);
			d[p] = P.f[cur][var][id] / dt - P.v_T * grad;
		} else
			d[p] = P.f[cur][var][id] / dt;
	}
	{   // ApplyBC1 (:97-116)
		const int id = dirX ? ID2(sg.end, line) : ID2(line, sg.end);
		if (P.bc[id] == CMC_BC_NOSLIP) { a[n - 1] = 0.0; b[n - 1] = 1.0; d[n - 1] = var == 0 ? P.gvx[id] : var == 1 ? P.gvy[id] : P.gT[id]; }
		else { a[n - 1] = 1.0; b[n - 1] = -1.0; d[n - 1] = 0.0; }
	}
	// Common::SolveTridiagonal (src/Common/Algorithms.h:21-38)
	c[n - 1] = 0.0;
	c[0] = c[0] / b[0];
	d[0] = d[0] / b[0];
	for (int p = 1; p < n; p++) {
		c[p] = c[p] / (b[p] - a[p] * c[p - 1]);
		d[p] = (d[p] - d[p - 1] * a[p]) / (b[p] - a[p] * c[p - 1]);
	}
	x[n - 1] = d[n - 1];
	for (int p = n - 2; p >= 0; p--) x[p] = d[p] - c[p] * x[p + 1];
	// UpdateSegment (:52-72): all n cells
	FT *out = P.f[L2_NEXT_LOCAL][var];
	for (int p = 0; p < n; p++) out[dirX ? ID2(i0 + p, j0) : ID2(i0, j0 + p)] = x[p];
}

// AdiSolver2D::SolveDirection (:205-226)
template <typename FT>
__device__ void solve_direction(const Dev2<FT> &P, FT dt, int num_local, int dirX, int cur, int temp, int next)
{
	const int N = P.dimx * P.dimy, maxn = max(P.dimx, P.dimy);
	const Seg2 *list = dirX ? P.listY : P.listX;
	const int nlines = dirX ? P.dimy : P.dimx;
	for (int t = threadIdx.x; t < N; t += blockDim.x)
		for (int q = 0; q < 3; q++) P.f[L2_TEMP_LOCAL][q][t] = FT(0);      // `new TimeLayer2D`: uninitialised in the reference
	__syncthreads();
	copy_cells(P, temp, L2_TEMP_LOCAL, -1, false);
	for (int it = 0; it < num_local; it++) {
		for (int t = threadIdx.x; t < nlines * 3; t += blockDim.x) {
			const int line = t / 3, var = t % 3;
			if (list[line].valid) solve_segment(P, dt, line, list[line], var, dirX, cur, P.scratch + (size_t)t * 5 * maxn, maxn);
		}
		__syncthreads();
		copy_cells(P, L2_NEXT_LOCAL, L2_TEMP_LOCAL, CMC_NODE_IN, it != 0);
	}
	copy_cells(P, L2_TEMP_LOCAL, temp, CMC_NODE_IN, false);
	copy_cells(P, L2_NEXT_LOCAL, next, CMC_NODE_IN, false);
}

// AdiSolver2D::TimeStep (:279-323)
// One thread block per case: blockIdx.x indexes the batch (cmc_adi2d_time_step_batch: many independent 2D cases - a
// parameter study, an ensemble - advance in ONE launch, one case per SM; a single case is a batch of one).
template <typename FT>
__global__ void __launch_bounds__(1024, 1) k_adi2d_time_step(const Dev2<FT> *batch, int num_global, int num_local)
{
	__shared__ double sh_err;
	__shared__ Dev2<FT> P;
	if (threadIdx.x == 0) P = batch[blockIdx.x];
	__syncthreads();
	const FT dt = P.dt;
	// CreateSegments (:228-277)
	for (int t = threadIdx.x; t < P.dimx + P.dimy; t += blockDim.x) {
		const bool col = t < P.dimx;            // listX: column i, scanned along j
		const int line = col ? t : t - P.dimx, n = col ? P.dimy : P.dimx;
		auto ty = [&](int p) { return col ? P.type[ID2(line, p)] : P.type[ID2(p, line)]; };
		Seg2 s; s.valid = 0; s.pos = 0; s.end = 0;
		int p = 0;
		while (p < n && ty(p) == CMC_NODE_OUT) p++;
		while (p + 1 < n && ty(p + 1) != CMC_NODE_IN) p++;
		if (p + 1 < n) {
			s.pos = p;
			p = n - 1;
			while (p >= 0 && ty(p) == CMC_NODE_OUT) p--;
			while (p - 1 >= 0 && ty(p - 1) != CMC_NODE_IN) p--;
			s.end = p; s.valid = 1;
		}
		(col ? P.listX : P.listY)[line] = s;
	}
	__syncthreads();
	copy_cells(P, L2_CUR, L2_NEXT, -1, false);
	copy_cells(P, L2_CUR, L2_HALF, -1, false);
	copy_cells(P, L2_CUR, L2_TEMP, -1, false);
	int it, status = 0;
	double err = eval_div_error(P, L2_NEXT, &sh_err);
	for (it = 0; (it < num_global) || (err > kErrThreshold2D); it++) {
		solve_direction(P, dt, num_local, 1, L2_CUR, L2_TEMP, L2_HALF);      // listY: direction X
		solve_direction(P, dt, num_local, 0, L2_HALF, L2_TEMP, L2_NEXT);     // listX: direction Y
		err = eval_div_error(P, L2_NEXT, &sh_err);
		copy_cells(P, L2_NEXT, L2_TEMP, CMC_NODE_IN, it != 0);
		if (it > kMaxGlobalIters2D) { status = 1; break; }
		if (err > kErrThreshold2D * 10) { status = 2; break; }
	}
	// Solver2D::ClearOutterCells (Solver2D.cpp:73-84), then next -> cur
	for (int t = threadIdx.x; t < P.dimx * P.dimy; t += blockDim.x)
		if (P.type[t] == CMC_NODE_OUT) { P.f[L2_NEXT][0][t] = 0.0; P.f[L2_NEXT][1][t] = 0.0; P.f[L2_NEXT][2][t] = P.startT; }
	__syncthreads();
	copy_cells(P, L2_NEXT, L2_CUR, -1, false);
	if (threadIdx.x == 0) { P.result[0] = err; P.result[1] = (double)it; P.result[2] = (double)status; }
}

// Solver2D::UpdateBoundaries (Solver2D.cpp:48-62)
template <typename FT>
__global__ void k_adi2d_update_boundaries(const Dev2<FT> P)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= P.dimx * P.dimy) return;
	const int ty = P.type[t];
	if (ty != CMC_NODE_BOUND && ty != CMC_NODE_VALVE) return;
	P.f[L2_CUR][0][t] = P.gvx[t]; P.f[L2_CUR][1][t] = P.gvy[t]; P.f[L2_CUR][2][t] = P.gT[t];
	const int i = t / P.dimy, j = t % P.dimy;
	if (i < P.dimx - 1 && j < P.dimy - 1)
		for (int q = 0; q < 3; q++) P.f[L2_NEXT][q][t] = P.f[L2_CUR][q][t];
}

template <typename FT>
__global__ void k_adi2d_update_boundaries_batch(const Dev2<FT> *batch)
{
	const Dev2<FT> &P = batch[blockIdx.y];
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= P.dimx * P.dimy) return;
	const int ty = P.type[t];
	if (ty != CMC_NODE_BOUND && ty != CMC_NODE_VALVE) return;
	P.f[L2_CUR][0][t] = P.gvx[t]; P.f[L2_CUR][1][t] = P.gvy[t]; P.f[L2_CUR][2][t] = P.gT[t];
	const int i = t / P.dimy, j = t % P.dimy;
	if (i < P.dimx - 1 && j < P.dimy - 1)
		for (int q = 0; q < 3; q++) P.f[L2_NEXT][q][t] = P.f[L2_CUR][q][t];
}

} // namespace

struct cmc_adi2d {
	virtual ~cmc_adi2d() {}
	virtual int set_grid(const int32_t *type, const int32_t *bc, const void *vx, const void *vy, const void *T) = 0;
	virtual int init_layer() = 0;
	virtual int update_boundaries() = 0;
	virtual int time_step(double dt, int ng, int nl, double *err, int *iters) = 0;
	virtual int step_host(const int32_t *type, const int32_t *bc, const void *vx, const void *vy, const void *T, void *const cur[3], void *const next[3],
	                      double dt, int ng, int nl, double *err, int *iters) = 0;
	virtual int get_layer(void *vel, double *T, int ox, int oy) = 0;
	virtual int rw_field(int layer, int var, void *dst, const void *src) = 0;
	virtual const void *dev_params(double dt) = 0;
	virtual size_t dev_params_size() const = 0;
	virtual int launch_batch(const void *dev_array, int n, int ng, int nl, bool with_boundaries, void *stream) = 0;
	virtual int collect(double *err, int *iters) = 0;
	virtual void *stream_handle() = 0;
	int device = 0, fp = 4, dimx = 0, dimy = 0;
	long long launches = 0;
};

namespace {

#define CU2(call)                                                                                           \
	do {                                                                                                    \
		cudaError_t e__ = (call);                                                                           \
		if (e__ != cudaSuccess) {                                                                           \
			char buf__[384];                                                                                \
			snprintf(buf__, sizeof buf__, "%s failed on device %d: %s (%d)", #call, device, cudaGetErrorString(e__), (int)e__); \
			return cmc_set_error(CMC_ERR_CUDA, buf__);                                                      \
		}                                                                                                   \
	} while (0)

template <typename FT>
struct Engine2D : cmc_adi2d {
	Dev2<FT> P{};
	Dev2<FT> *d_self = nullptr;     // device copy of P: the batch of one
	std::vector<void *> allocs;
	cudaStream_t stream = nullptr;
	bool have_grid = false;
	// everything a host-driven step moves lives in ONE device block, mirrored by one pinned host block:
	// [cur u v T | next u v T | grid vx vy T | type | bc] - one copy up, one copy (the first six arrays) down (step_host)
	unsigned char *io_dev = nullptr, *io_host = nullptr;
	size_t io_bytes = 0;

	~Engine2D() override
	{
		cudaSetDevice(device);
		if (stream) cudaStreamSynchronize(stream);
		for (void *p : allocs) cudaFree(p);
		if (io_host) cudaFreeHost(io_host);
		if (stream) cudaStreamDestroy(stream);
	}
	template <typename T>
	int dalloc(T *&p, size_t n)
	{
		CU2(cudaMalloc((void **)&p, n * sizeof(T)));
		CU2(cudaMemset(p, 0, n * sizeof(T)));
		allocs.push_back(p);
		return CMC_OK;
	}
	int init(int dx_, int dy_, double dx, double dy, const cmc_fluid_params *fp_, double startT)
	{
		CU2(cudaSetDevice(device));
		CU2(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
		dimx = dx_; dimy = dy_;
		P.dimx = dx_; P.dimy = dy_; P.dx = (FT)dx; P.dy = (FT)dy; P.gdx = (FT)dx; P.gdy = (FT)dy;
		P.v_T = (FT)fp_->v_T; P.v_vis = (FT)fp_->v_vis; P.t_vis = (FT)fp_->t_vis; P.t_phi = (FT)fp_->t_phi; P.startT = (FT)startT;
		const size_t N = (size_t)dimx * dimy, maxn = (size_t)std::max(dimx, dimy);
		int rc;
		io_bytes = 9 * N * sizeof(FT) + 2 * N * sizeof(int);
		if ((rc = dalloc(io_dev, io_bytes))) return rc;
		CU2(cudaMallocHost((void **)&io_host, io_bytes));
		FT *fb = reinterpret_cast<FT *>(io_dev);
		for (int q = 0; q < 3; q++) { P.f[L2_CUR][q] = fb + q * N; P.f[L2_NEXT][q] = fb + (3 + q) * N; }
		P.gvx = fb + 6 * N; P.gvy = fb + 7 * N; P.gT = fb + 8 * N;
		P.type = reinterpret_cast<const int *>(fb + 9 * N); P.bc = P.type + N;
		// the reference allocates half / next / temp / next_local uninitialised (TimeLayer2D.h:176-181): zero here
		for (int l = 0; l < L2_COUNT; l++)
			if (l != L2_CUR && l != L2_NEXT)
				for (int q = 0; q < 3; q++) if ((rc = dalloc(P.f[l][q], N))) return rc;
		if ((rc = dalloc(P.listX, (size_t)dimx)) || (rc = dalloc(P.listY, (size_t)dimy))) return rc;
		if ((rc = dalloc(P.scratch, 3 * maxn * 5 * maxn))) return rc;
		if ((rc = dalloc(P.resid, N))) return rc;
		if ((rc = dalloc(P.result, 4))) return rc;
		if ((rc = dalloc(d_self, 1))) return rc;
		return CMC_OK;
	}
	int set_grid(const int32_t *type, const int32_t *bc, const void *vx, const void *vy, const void *T) override
	{
		CU2(cudaSetDevice(device));
		const size_t N = (size_t)dimx * dimy;
		for (size_t i = 0; i < N; i++)
			if (type[i] < 0 || type[i] > 3 || bc[i] < 0 || bc[i] > 1) return cmc_set_error(CMC_ERR_INVALID, "adi2d set_grid: node type / boundary type out of range");
		CU2(cudaMemcpyAsync((void *)P.type, type, N * sizeof(int), cudaMemcpyHostToDevice, stream));
		CU2(cudaMemcpyAsync((void *)P.bc, bc, N * sizeof(int), cudaMemcpyHostToDevice, stream));
		CU2(cudaMemcpyAsync((void *)P.gvx, vx, N * sizeof(FT), cudaMemcpyHostToDevice, stream));
		CU2(cudaMemcpyAsync((void *)P.gvy, vy, N * sizeof(FT), cudaMemcpyHostToDevice, stream));
		CU2(cudaMemcpyAsync((void *)P.gT, T, N * sizeof(FT), cudaMemcpyHostToDevice, stream));
		CU2(cudaStreamSynchronize(stream));
		have_grid = true;
		return CMC_OK;
	}
	int init_layer() override      // AdiSolver2D::Init (:36-50): cur <- grid data in every cell
	{
		if (!have_grid) return cmc_set_error(CMC_ERR_INVALID, "adi2d: call cmc_adi2d_set_grid first");
		CU2(cudaSetDevice(device));
		const size_t B = (size_t)dimx * dimy * sizeof(FT);
		CU2(cudaMemcpyAsync(P.f[L2_CUR][0], P.gvx, B, cudaMemcpyDeviceToDevice, stream));
		CU2(cudaMemcpyAsync(P.f[L2_CUR][1], P.gvy, B, cudaMemcpyDeviceToDevice, stream));
		CU2(cudaMemcpyAsync(P.f[L2_CUR][2], P.gT, B, cudaMemcpyDeviceToDevice, stream));
		CU2(cudaStreamSynchronize(stream));
		return CMC_OK;
	}
	int update_boundaries() override
	{
		if (!have_grid) return cmc_set_error(CMC_ERR_INVALID, "adi2d: call cmc_adi2d_set_grid first");
		CU2(cudaSetDevice(device));
		const int N = dimx * dimy;
		k_adi2d_update_boundaries<FT><<<(N + 255) / 256, 256, 0, stream>>>(P);
		launches++;
		return CMC_OK;
	}
	int time_step(double dt, int ng, int nl, double *err, int *iters) override
	{
		if (!have_grid) return cmc_set_error(CMC_ERR_INVALID, "adi2d: call cmc_adi2d_set_grid first");
		if (ng < 0 || nl < 0) return cmc_set_error(CMC_ERR_INVALID, "adi2d time_step: negative iteration count");
		CU2(cudaSetDevice(device));
		P.dt = (FT)dt;                                                       // (FTYPE)dt: FluidSolver2D.cpp:131
		CU2(cudaMemcpyAsync(d_self, &P, sizeof P, cudaMemcpyHostToDevice, stream));
		k_adi2d_time_step<FT><<<1, 1024, 0, stream>>>(d_self, ng, nl);
		launches++;
		return collect(err, iters);
	}
	// set_grid + the host's cur / next layers up, one step, both layers down: two copies and one synchronisation
	int step_host(const int32_t *type, const int32_t *bc, const void *vx, const void *vy, const void *T, void *const cur[3], void *const next[3],
	              double dt, int ng, int nl, double *err, int *iters) override
	{
		if (ng < 0 || nl < 0) return cmc_set_error(CMC_ERR_INVALID, "adi2d step_host: negative iteration count");
		CU2(cudaSetDevice(device));
		const size_t N = (size_t)dimx * dimy, B = N * sizeof(FT);
		for (size_t i = 0; i < N; i++)
			if (type[i] < 0 || type[i] > 3 || bc[i] < 0 || bc[i] > 1) return cmc_set_error(CMC_ERR_INVALID, "adi2d step_host: node type / boundary type out of range");
		for (int q = 0; q < 3; q++) { memcpy(io_host + q * B, cur[q], B); memcpy(io_host + (3 + q) * B, next[q], B); }
		memcpy(io_host + 6 * B, vx, B); memcpy(io_host + 7 * B, vy, B); memcpy(io_host + 8 * B, T, B);
		memcpy(io_host + 9 * B, type, N * sizeof(int)); memcpy(io_host + 9 * B + N * sizeof(int), bc, N * sizeof(int));
		CU2(cudaMemcpyAsync(io_dev, io_host, io_bytes, cudaMemcpyHostToDevice, stream));
		have_grid = true;
		P.dt = (FT)dt;
		CU2(cudaMemcpyAsync(d_self, &P, sizeof P, cudaMemcpyHostToDevice, stream));
		k_adi2d_time_step<FT><<<1, 1024, 0, stream>>>(d_self, ng, nl);
		launches++;
		CU2(cudaMemcpyAsync(io_host, io_dev, 6 * B, cudaMemcpyDeviceToHost, stream));
		const int rc = collect(err, iters);          // synchronises
		if (rc != CMC_OK && rc != CMC_ERR_DIVERGED) return rc;
		for (int q = 0; q < 3; q++) { memcpy(cur[q], io_host + q * B, B); memcpy(next[q], io_host + (3 + q) * B, B); }
		return rc;
	}
	int collect(double *err, int *iters)
	{
		double r[4] = {};
		CU2(cudaMemcpyAsync(r, P.result, sizeof r, cudaMemcpyDeviceToHost, stream));
		CU2(cudaStreamSynchronize(stream));
		CU2(cudaGetLastError());
		if (err) *err = r[0];
		if (iters) *iters = (int)r[1];
		if (r[2] == 1.0) return cmc_set_error(CMC_ERR_DIVERGED, "Exceeded max number of iterations (100)");   // AdiSolver2D.cpp:303-307 (the reference exits)
		if (r[2] == 2.0) return cmc_set_error(CMC_ERR_DIVERGED, "Error is too big!");                         // :309-313
		return CMC_OK;
	}
	// batch: the Dev2 of every member on this (the first) handle's stream; update_boundaries of all members rides in the same launch sequence
	const void *dev_params(double dt) override { P.dt = (FT)dt; return &P; }
	size_t dev_params_size() const override { return sizeof P; }
	int launch_batch(const void *dev_array, int n, int ng, int nl, bool with_boundaries, void *stream_) override
	{
		CU2(cudaSetDevice(device));
		cudaStream_t st = (cudaStream_t)stream_;
		if (with_boundaries) k_adi2d_update_boundaries_batch<FT><<<dim3((dimx * dimy + 255) / 256, n), 256, 0, st>>>((const Dev2<FT> *)dev_array);
		k_adi2d_time_step<FT><<<n, 1024, 0, st>>>((const Dev2<FT> *)dev_array, ng, nl);
		CU2(cudaGetLastError());
		launches += with_boundaries ? 2 : 1;
		return CMC_OK;
	}
	void *stream_handle() override { return (void *)stream; }
	int get_layer(void *vel, double *T, int ox, int oy) override      // Solver2D::GetLayer (Solver2D.cpp:20-34)
	{
		CU2(cudaSetDevice(device));
		if (ox == 0) ox = dimx;
		if (oy == 0) oy = dimy;
		if (ox < 0 || oy < 0) return cmc_set_error(CMC_ERR_INVALID, "adi2d get_layer: negative output dims");
		const size_t N = (size_t)dimx * dimy;
		std::vector<FT> h[3];
		for (int q = 0; q < 3; q++) { h[q].resize(N); CU2(cudaMemcpyAsync(h[q].data(), P.f[L2_NEXT][q], N * sizeof(FT), cudaMemcpyDeviceToHost, stream)); }
		CU2(cudaStreamSynchronize(stream));
		FT *v = (FT *)vel;
		for (int i = 0; i < ox; i++)
			for (int j = 0; j < oy; j++) {
				const size_t id = (size_t)(i * dimx / ox) * dimy + (size_t)(j * dimy / oy);
				v[2 * ((size_t)i * oy + j)] = h[0][id]; v[2 * ((size_t)i * oy + j) + 1] = h[1][id]; T[(size_t)i * oy + j] = h[2][id];
			}
		return CMC_OK;
	}
	int rw_field(int layer, int var, void *dst, const void *src) override
	{
		if (layer < 0 || layer >= L2_COUNT || var < 0 || var > 2) return cmc_set_error(CMC_ERR_INVALID, "adi2d field: bad layer/var");
		CU2(cudaSetDevice(device));
		const size_t B = (size_t)dimx * dimy * sizeof(FT);
		if (dst) CU2(cudaMemcpyAsync(dst, P.f[layer][var], B, cudaMemcpyDeviceToHost, stream));
		else CU2(cudaMemcpyAsync(P.f[layer][var], src, B, cudaMemcpyHostToDevice, stream));
		CU2(cudaStreamSynchronize(stream));
		return CMC_OK;
	}
};

} // namespace

extern int cmc_check_device(int device);   // cmc_adi.cu

extern "C" {

int cmc_adi2d_create(int dimx, int dimy, double dx, double dy, const cmc_fluid_params *params, double startT, int fp_bytes, int device, cmc_adi2d **out)
{
	if (!params || !out) return cmc_set_error(CMC_ERR_INVALID, "adi2d create: null argument");
	*out = nullptr;
	if (fp_bytes != 4 && fp_bytes != 8) return cmc_set_error(CMC_ERR_INVALID, "adi2d create: fp_bytes must be 4 or 8");
	if (dimx < 3 || dimy < 3 || dimx > 4096 || dimy > 4096) return cmc_set_error(CMC_ERR_INVALID, "adi2d create: grid dimensions must be in 3..4096");
	if (!(dx > 0) || !(dy > 0)) return cmc_set_error(CMC_ERR_INVALID, "adi2d create: grid spacing must be positive");
	int rc = cmc_check_device(device);
	if (rc) return rc;
	cmc_adi2d *h;
	if (fp_bytes == 4) { auto *e = new (std::nothrow) Engine2D<float>(); if (!e) return cmc_set_error(CMC_ERR_INVALID, "out of host memory"); e->device = device; e->fp = 4; rc = e->init(dimx, dimy, dx, dy, params, startT); h = e; }
	else { auto *e = new (std::nothrow) Engine2D<double>(); if (!e) return cmc_set_error(CMC_ERR_INVALID, "out of host memory"); e->device = device; e->fp = 8; rc = e->init(dimx, dimy, dx, dy, params, startT); h = e; }
	if (rc) { delete h; return rc; }
	*out = h;
	return CMC_OK;
}
int cmc_adi2d_destroy(cmc_adi2d *h) { delete h; return CMC_OK; }
#define H2(h) if (!(h)) return cmc_set_error(CMC_ERR_INVALID, "null handle")
int cmc_adi2d_set_grid(cmc_adi2d *h, const int32_t *type, const int32_t *bc_type, const void *vx, const void *vy, const void *T)
{
	H2(h);
	if (!type || !bc_type || !vx || !vy || !T) return cmc_set_error(CMC_ERR_INVALID, "adi2d set_grid: null array");
	return h->set_grid(type, bc_type, vx, vy, T);
}
int cmc_adi2d_init_layer(cmc_adi2d *h) { H2(h); return h->init_layer(); }
int cmc_adi2d_update_boundaries(cmc_adi2d *h) { H2(h); return h->update_boundaries(); }
int cmc_adi2d_time_step(cmc_adi2d *h, double dt, int num_global, int num_local, double *err_out, int *iters_out)
{
	H2(h);
	return h->time_step(dt, num_global, num_local, err_out, iters_out);
}
int cmc_adi2d_step_host(cmc_adi2d *h, const int32_t *type, const int32_t *bc_type, const void *vx, const void *vy, const void *T, void *const cur_uvT[3],
                        void *const next_uvT[3], double dt, int num_global, int num_local, double *err_out, int *iters_out)
{
	H2(h);
	if (!type || !bc_type || !vx || !vy || !T || !cur_uvT || !next_uvT) return cmc_set_error(CMC_ERR_INVALID, "adi2d step_host: null argument");
	for (int q = 0; q < 3; q++)
		if (!cur_uvT[q] || !next_uvT[q]) return cmc_set_error(CMC_ERR_INVALID, "adi2d step_host: null layer array");
	return h->step_host(type, bc_type, vx, vy, T, cur_uvT, next_uvT, dt, num_global, num_local, err_out, iters_out);
}
int cmc_adi2d_get_layer(cmc_adi2d *h, void *vel_xy, double *T, int outdimx, int outdimy)
{
	H2(h);
	if (!vel_xy || !T) return cmc_set_error(CMC_ERR_INVALID, "adi2d get_layer: null output");
	return h->get_layer(vel_xy, T, outdimx, outdimy);
}
int cmc_adi2d_read_field(cmc_adi2d *h, int layer, int var, void *dst)
{
	H2(h);
	if (!dst) return cmc_set_error(CMC_ERR_INVALID, "adi2d read_field: null destination");
	return h->rw_field(layer, var, dst, nullptr);
}
int cmc_adi2d_write_field(cmc_adi2d *h, int layer, int var, const void *src)
{
	H2(h);
	if (!src) return cmc_set_error(CMC_ERR_INVALID, "adi2d write_field: null source");
	return h->rw_field(layer, var, nullptr, src);
}
// Many independent cases in one launch (one thread block - one SM - per case).  All handles: same precision, same grid
// dimensions, same device.  update_boundaries != 0 runs Solver2D::UpdateBoundaries of every case first.  err_out /
// iters_out / status_out: per case (status: CMC_OK or CMC_ERR_DIVERGED); returns the first non-OK status.
int cmc_adi2d_time_step_batch(cmc_adi2d *const *handles, int n, double dt, int num_global, int num_local, int update_boundaries,
                              double *err_out, int *iters_out, int *status_out)
{
	if (!handles || n < 1) return cmc_set_error(CMC_ERR_INVALID, "adi2d batch: no handles");
	for (int i = 0; i < n; i++) {
		if (!handles[i]) return cmc_set_error(CMC_ERR_INVALID, "adi2d batch: null handle");
		if (handles[i]->fp != handles[0]->fp || handles[i]->device != handles[0]->device || handles[i]->dimx != handles[0]->dimx || handles[i]->dimy != handles[0]->dimy)
			return cmc_set_error(CMC_ERR_INVALID, "adi2d batch: every case needs the same precision, grid dimensions and device");
	}
	if (num_global < 0 || num_local < 0) return cmc_set_error(CMC_ERR_INVALID, "adi2d batch: negative iteration count");
	cudaSetDevice(handles[0]->device);
	const size_t psz = handles[0]->dev_params_size();
	std::vector<char> host((size_t)n * psz);
	for (int i = 0; i < n; i++) memcpy(host.data() + (size_t)i * psz, handles[i]->dev_params(dt), psz);
	void *dev = nullptr;
	cudaStream_t st = (cudaStream_t)handles[0]->stream_handle();
	if (cudaMalloc(&dev, host.size()) != cudaSuccess) return cmc_set_error(CMC_ERR_CUDA, "adi2d batch: out of device memory");
	if (cudaMemcpyAsync(dev, host.data(), host.size(), cudaMemcpyHostToDevice, st) != cudaSuccess) { cudaFree(dev); return cmc_set_error(CMC_ERR_CUDA, "adi2d batch: upload failed"); }
	int rc = handles[0]->launch_batch(dev, n, num_global, num_local, update_boundaries != 0, st);
	int first = rc;
	if (cudaStreamSynchronize(st) != cudaSuccess) first = cmc_set_error(CMC_ERR_CUDA, std::string("adi2d batch: ") + cudaGetErrorString(cudaGetLastError()));
	cudaFree(dev);
	if (first) return first;
	for (int i = 0; i < n; i++) {
		double e = 0.0; int it = 0;
		const int r = handles[i]->collect(&e, &it);
		if (err_out) err_out[i] = e;
		if (iters_out) iters_out[i] = it;
		if (status_out) status_out[i] = r;
		if (r && !first) first = r;
	}
	return first;
}

int cmc_adi2d_launch_count(const cmc_adi2d *h, int64_t *n) { H2(h); if (n) *n = h->launches; return CMC_OK; }

} // extern "C"
