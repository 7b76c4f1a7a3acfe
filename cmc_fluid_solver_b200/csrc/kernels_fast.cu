// kernels_fast.cu - CMC_MODE_FAST directional sweeps for sm_100a.
//
// One kernel per sweep does what the reference does in 4 solve kernels + merge kernels
// (src/FluidSolver3D/AdiSolver3D.cu:376-457, TimeLayer3D.cu:133-146): coefficient build with boundary
// conditions and obstacle mask folded in, the tridiagonal solves of u, v, w (one shared matrix, three
// right-hand sides) and T, the scatter into `next` and the nonlinear-layer relaxation
// temp' = (temp + next)/2 on fluid cells (written to the second temp buffer so that neighbouring lines
// keep reading the old linearisation state: CPU/Jacobi semantics).
//
// Line solve = partition method (Wang / SPIKE) so that no Thomas intermediate ever leaves the SM:
//   * every grid line (all segments of it at once: boundary rows decouple them) is cut into chunks of
//     M = 8 rows, one thread per chunk; the last row of each chunk is its separator;
//   * the 7 interior rows are eliminated in registers (forward sweep carrying the left spike), a short
//     backward recurrence gives the chunk's coupling to its two separators;
//   * the separators form a reduced tridiagonal system of G = n/8 unknowns per line, solved by
//     parallel cyclic reduction (PCR) in shared memory, one thread per unknown, normalised rows
//     (one reciprocal per step);
//   * back substitution of the interior rows from registers, results stored once.
//
// Thread mapping.  X and Y sweeps (strided lines): a CTA owns NL = 8 neighbouring k-columns of one
// (j or i) row-set; lanes run along k first (64-byte segments per row for fp64), chunks across warps.
// Z sweep (contiguous lines): a CTA owns NL = 8 neighbouring lines; lanes run along the chunk index, so a
// warp reads 32 consecutive 64-byte chunks = 2 KB of one line with 128-bit loads.
// HBM traffic per cell and sweep is the algorithmic 16 values + 1 descriptor byte: stencil neighbours and
// re-reads for the merge are served by L1/L2.
#include "kernels.h"
#include "rows.cuh"

namespace cmc {

constexpr int M = 8;          // rows per chunk
constexpr int NL = 8;         // lines per CTA

template <typename FT> __device__ __forceinline__ FT rcp(FT x) { return FT(1) / x; }

// ---- PCR over the reduced systems of a CTA ---------------------------------------------------------------
// Rows are normalised (B == 1).  sys[(slot*NR + r) * GP * NL + g * NL + l], slot in {0,1} ping-pong.
// NR = 2 + NRHS (A, C, D...).  Element (g, l) = chunk g of line l.  GP = G rounded up to a power of two
// (rows >= G are identity rows).
template <typename FT, int NRHS>
__device__ __forceinline__ void pcr_solve(FT *sys, int GP, int g, int l, bool active, FT Ain, FT Cin, const FT (&Din)[NRHS], FT (&X)[NRHS])
{
	constexpr int NR = 2 + NRHS;
	const int stride = GP * NL;
	const int e = g * NL + l;
	FT A = Ain, Cc = Cin, D[NRHS];
#pragma unroll
	for (int q = 0; q < NRHS; q++) D[q] = Din[q];
	int buf = 0;
	for (int s = 1; s < GP; s <<= 1) {
		FT *w = sys + buf * NR * stride;
		if (active) {
			w[0 * stride + e] = A; w[1 * stride + e] = Cc;
#pragma unroll
			for (int q = 0; q < NRHS; q++) w[(2 + q) * stride + e] = D[q];
		}
		__syncthreads();
		if (active) {
			const bool lo = g - s >= 0, hi = g + s < GP;
			const int el = e - s * NL, eh = e + s * NL;
			const FT Al = lo ? w[0 * stride + el] : FT(0), Cl = lo ? w[1 * stride + el] : FT(0);
			const FT Ah = hi ? w[0 * stride + eh] : FT(0), Ch = hi ? w[1 * stride + eh] : FT(0);
			const FT r = rcp<FT>(FT(1) - A * Cl - Cc * Ah);
#pragma unroll
			for (int q = 0; q < NRHS; q++) {
				const FT Dl = lo ? w[(2 + q) * stride + el] : FT(0), Dh = hi ? w[(2 + q) * stride + eh] : FT(0);
				D[q] = (D[q] - A * Dl - Cc * Dh) * r;
			}
			A = -A * Al * r;
			Cc = -Cc * Ch * r;
		}
		buf ^= 1;
	}
#pragma unroll
	for (int q = 0; q < NRHS; q++) X[q] = D[q];
}

// shared-memory exchange area: separators' solutions and chunk heads
template <typename FT>
struct Smem {
	FT *sys;      // PCR ping-pong: 2 * (2 + 3) * GP * NL
	FT *head;     // per chunk: y0[3], v0, w0 -> 5 * GP * NL  (read by the previous chunk)
	FT *sol;      // separator solutions E: 3 * (GP + 1) * NL (entry 0 = virtual chunk -1)
};

template <typename FT, int DIR>
__global__ void __launch_bounds__(512, 1) k_fast_sweep(const SweepArgs<FT> A, const int G, const int GP)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	FT *sys = reinterpret_cast<FT *>(smem_raw);
	FT *head = sys + 2 * 5 * GP * NL;
	FT *sol = head + 5 * GP * NL;

	const Layout &L = A.L;
	const int t = threadIdx.x;
	int g, l;                       // chunk, line-in-CTA
	if (DIR == 2) { g = t % GP; l = t / GP; } else { l = t % NL; g = t / NL; }
	const bool active = g < GP && l < NL;     // always true by construction (blockDim = GP * NL)

	// ---- which line ----------------------------------------------------------------------------------------
	const long long sx = L.plane, sy = L.nzp, sz = 1;
	long long base;                 // element index of row 0 of this thread's line
	long long stride;               // along the line
	int n;                          // rows of the line
	bool line_ok;
	if (DIR == 0) {                 // lines along x: CTA = (j, k-tile)
		const int ktiles = (L.nz + NL - 1) / NL;
		const int j = blockIdx.x / ktiles, k = (blockIdx.x % ktiles) * NL + l;
		line_ok = k < L.nz; n = L.nx; stride = sx; base = L.idx(0, j, line_ok ? k : 0);
	} else if (DIR == 1) {          // lines along y: CTA = (i, k-tile)
		const int ktiles = (L.nz + NL - 1) / NL;
		const int i = blockIdx.x / ktiles, k = (blockIdx.x % ktiles) * NL + l;
		line_ok = k < L.nz; n = L.ny; stride = sy; base = L.idx(i, 0, line_ok ? k : 0);
	} else {                        // lines along z: CTA = (i, j-tile)
		const int jtiles = (L.ny + NL - 1) / NL;
		const int i = blockIdx.x / jtiles, j = (blockIdx.x % jtiles) * NL + l;
		line_ok = j < L.ny; n = L.nz; stride = sz; base = L.idx(i, line_ok ? j : 0, 0);
	}
	const int r0 = g * M;           // first row of this chunk
	RowConst<FT> K; K.init(A, DIR);

	// roles of the chunk's rows (rows >= n or lines outside the grid: no segment, no store)
	unsigned role[M];
#pragma unroll
	for (int i = 0; i < M; i++) {
		const int r = r0 + i;
		role[i] = (line_ok && r < n) ? (unsigned)A.role[base + (long long)r * stride] : 0u;
	}

	const int e = g * NL + l;
	const int stride_s = GP * NL;

	// ======================================= phase V: u, v, w ==============================================
	FT x3[3][M];
	{
		FT cp[M], lp[M], dp[3][M];
		FT b7 = FT(1);
#pragma unroll
		for (int i = 0; i < M; i++) {
			const long long id = base + (long long)(r0 + i) * stride;
			FT a, b, c, d[3];
			const unsigned r = role[i];
			if (r & R_INT) {
				const FT V = A.temp[DIR][id];
				a = -V / K.two_h - K.vis_v; c = V / K.two_h - K.vis_v; b = K.b_v;
				const FT grad = (A.temp[3][id + stride] - A.temp[3][id - stride]) / K.two_h;
				d[0] = A.cur[0][id] * 3 / K.dt; d[1] = A.cur[1][id] * 3 / K.dt; d[2] = A.cur[2][id] * 3 / K.dt;
				d[DIR] -= K.v_T * grad;
			} else if (r & (R_START | R_END)) {
				const bool free_bc = r & R_VFREE;
				b = free_bc ? FT(2) : FT(1);
				a = (r & R_END) ? (free_bc ? FT(-1) : FT(0)) : FT(0);
				c = (r & R_START) ? (free_bc ? FT(-1) : FT(0)) : FT(0);
				if (free_bc) { d[0] = d[1] = d[2] = FT(0); }
				else { d[0] = A.nodev[0][id]; d[1] = A.nodev[1][id]; d[2] = A.nodev[2][id]; }
			} else { a = FT(0); b = FT(1); c = FT(0); d[0] = d[1] = d[2] = FT(0); }
			if (i == M - 1) {       // separator row stays raw
				lp[i] = a; cp[i] = c; b7 = b; dp[0][i] = d[0]; dp[1][i] = d[1]; dp[2][i] = d[2];
			} else if (i == 0) {
				const FT rr = rcp<FT>(b);
				cp[0] = c * rr; lp[0] = a * rr; dp[0][0] = d[0] * rr; dp[1][0] = d[1] * rr; dp[2][0] = d[2] * rr;
			} else {
				const FT rr = rcp<FT>(b - a * cp[i - 1]);
				cp[i] = c * rr; lp[i] = -a * lp[i - 1] * rr;
				dp[0][i] = (d[0] - a * dp[0][i - 1]) * rr;
				dp[1][i] = (d[1] - a * dp[1][i - 1]) * rr;
				dp[2][i] = (d[2] - a * dp[2][i - 1]) * rr;
			}
		}
		// coupling of the first interior row to the two separators: x_0 = y0 - v0*E(g-1) - w0*E(g)
		FT y0[3] = {dp[0][M - 2], dp[1][M - 2], dp[2][M - 2]}, v0 = lp[M - 2], w0 = cp[M - 2];
#pragma unroll
		for (int i = M - 3; i >= 0; i--) {
			y0[0] = dp[0][i] - cp[i] * y0[0]; y0[1] = dp[1][i] - cp[i] * y0[1]; y0[2] = dp[2][i] - cp[i] * y0[2];
			v0 = lp[i] - cp[i] * v0; w0 = -cp[i] * w0;
		}
		head[0 * stride_s + e] = y0[0]; head[1 * stride_s + e] = y0[1]; head[2 * stride_s + e] = y0[2];
		head[3 * stride_s + e] = v0; head[4 * stride_s + e] = w0;
		__syncthreads();
		// reduced row of this chunk's separator (row M-1): needs the head of chunk g+1
		FT Ra, Rc, Rd[3];
		{
			const bool has_next = g + 1 < GP;
			const int en = e + NL;
			const FT ny0 = has_next ? head[0 * stride_s + en] : FT(0), ny1 = has_next ? head[1 * stride_s + en] : FT(0),
			         ny2 = has_next ? head[2 * stride_s + en] : FT(0);
			const FT nv = has_next ? head[3 * stride_s + en] : FT(0), nw = has_next ? head[4 * stride_s + en] : FT(0);
			const FT a7 = lp[M - 1], c7 = cp[M - 1];
			const FT yL0 = dp[0][M - 2], yL1 = dp[1][M - 2], yL2 = dp[2][M - 2], vL = lp[M - 2], wL = cp[M - 2];
			const FT rr = rcp<FT>(b7 - a7 * wL - c7 * nv);
			Ra = -a7 * vL * rr; Rc = -c7 * nw * rr;
			Rd[0] = (dp[0][M - 1] - a7 * yL0 - c7 * ny0) * rr;
			Rd[1] = (dp[1][M - 1] - a7 * yL1 - c7 * ny1) * rr;
			Rd[2] = (dp[2][M - 1] - a7 * yL2 - c7 * ny2) * rr;
		}
		FT E[3];
		pcr_solve<FT, 3>(sys, GP, g, l, active, Ra, Rc, Rd, E);
		// publish separator solutions; entry g+1 (entry 0 = no chunk on the left)
		sol[0 * (GP + 1) * NL + (g + 1) * NL + l] = E[0];
		sol[1 * (GP + 1) * NL + (g + 1) * NL + l] = E[1];
		sol[2 * (GP + 1) * NL + (g + 1) * NL + l] = E[2];
		if (g == 0) { sol[0 * (GP + 1) * NL + l] = FT(0); sol[1 * (GP + 1) * NL + l] = FT(0); sol[2 * (GP + 1) * NL + l] = FT(0); }
		__syncthreads();
#pragma unroll
		for (int q = 0; q < 3; q++) {
			const FT El = sol[q * (GP + 1) * NL + g * NL + l];
			x3[q][M - 1] = E[q];
#pragma unroll
			for (int i = M - 2; i >= 0; i--) x3[q][i] = dp[q][i] - lp[i] * El - cp[i] * x3[q][i + 1];
		}
	}
	// store u, v, w and the relaxed linearisation layer
#pragma unroll
	for (int i = 0; i < M; i++) {
		const int r = r0 + i;
		if (!(line_ok && r < n)) continue;
		const long long id = base + (long long)r * stride;
		const unsigned ro = role[i];
		const bool seg = ro & R_SEG, in = ro & R_IN;
#pragma unroll
		for (int q = 0; q < 3; q++) {
			if (seg) A.next[q][id] = x3[q][i];
			const FT tq = A.temp[q][id];
			FT o = tq;
			if (in) o = (tq + (seg ? x3[q][i] : A.next[q][id])) / 2;
			A.temp_out[q][id] = o;
		}
	}

	// ======================================= phase T ======================================================
	__syncthreads();                // head / sol / sys are reused
	{
		FT cp[M], lp[M], dp[M];
		FT b7 = FT(1);
#pragma unroll
		for (int i = 0; i < M; i++) {
			const long long id = base + (long long)(r0 + i) * stride;
			FT a, b, c, d;
			const unsigned r = role[i];
			if (r & R_INT) {
				FT av, cv, aT, cT, dd[4];
				build_interior_row<FT, DIR>(A, K, id, sx, sy, sz, av, cv, aT, cT, dd);
				a = aT; c = cT; b = K.b_T; d = dd[3];
			} else if (r & (R_START | R_END)) {
				const bool free_bc = r & R_TFREE;
				b = free_bc ? FT(2) : FT(1);
				a = (r & R_END) ? (free_bc ? FT(-1) : FT(0)) : FT(0);
				c = (r & R_START) ? (free_bc ? FT(-1) : FT(0)) : FT(0);
				d = free_bc ? FT(0) : A.nodev[3][id];
			} else { a = FT(0); b = FT(1); c = FT(0); d = FT(0); }
			if (i == M - 1) { lp[i] = a; cp[i] = c; b7 = b; dp[i] = d; }
			else if (i == 0) { const FT rr = rcp<FT>(b); cp[0] = c * rr; lp[0] = a * rr; dp[0] = d * rr; }
			else {
				const FT rr = rcp<FT>(b - a * cp[i - 1]);
				cp[i] = c * rr; lp[i] = -a * lp[i - 1] * rr; dp[i] = (d - a * dp[i - 1]) * rr;
			}
		}
		FT y0 = dp[M - 2], v0 = lp[M - 2], w0 = cp[M - 2];
#pragma unroll
		for (int i = M - 3; i >= 0; i--) { y0 = dp[i] - cp[i] * y0; v0 = lp[i] - cp[i] * v0; w0 = -cp[i] * w0; }
		head[0 * stride_s + e] = y0; head[3 * stride_s + e] = v0; head[4 * stride_s + e] = w0;
		__syncthreads();
		FT Ra, Rc, Rd[1];
		{
			const bool has_next = g + 1 < GP;
			const int en = e + NL;
			const FT ny0 = has_next ? head[0 * stride_s + en] : FT(0);
			const FT nv = has_next ? head[3 * stride_s + en] : FT(0), nw = has_next ? head[4 * stride_s + en] : FT(0);
			const FT a7 = lp[M - 1], c7 = cp[M - 1];
			const FT rr = rcp<FT>(b7 - a7 * cp[M - 2] - c7 * nv);
			Ra = -a7 * lp[M - 2] * rr; Rc = -c7 * nw * rr;
			Rd[0] = (dp[M - 1] - a7 * dp[M - 2] - c7 * ny0) * rr;
		}
		FT E[1];
		pcr_solve<FT, 1>(sys, GP, g, l, active, Ra, Rc, Rd, E);
		sol[(g + 1) * NL + l] = E[0];
		if (g == 0) sol[l] = FT(0);
		__syncthreads();
		const FT El = sol[g * NL + l];
		FT x[M];
		x[M - 1] = E[0];
#pragma unroll
		for (int i = M - 2; i >= 0; i--) x[i] = dp[i] - lp[i] * El - cp[i] * x[i + 1];
#pragma unroll
		for (int i = 0; i < M; i++) {
			const int r = r0 + i;
			if (!(line_ok && r < n)) continue;
			const long long id = base + (long long)r * stride;
			const unsigned ro = role[i];
			const bool seg = ro & R_SEG, in = ro & R_IN;
			if (seg) A.next[3][id] = x[i];
			const FT tq = A.temp[3][id];
			FT o = tq;
			if (in) o = (tq + (seg ? x[i] : A.next[3][id])) / 2;
			A.temp_out[3][id] = o;
		}
	}
}

template <typename FT>
static size_t fast_smem_bytes(int GP) { return sizeof(FT) * (size_t)(2 * 5 * GP * NL + 5 * GP * NL + 3 * (GP + 1) * NL); }

template <typename FT>
bool launch_fast_sweep(int dir, const SweepArgs<FT> &A, cudaStream_t s, long long *launches)
{
	const Layout &L = A.L;
	const int n = dir == 0 ? L.nx : dir == 1 ? L.ny : L.nz;
	const int G = (n + M - 1) / M;
	int GP = 1;
	while (GP < G) GP <<= 1;
	if (GP * NL > 512) return false;                      // lines longer than 512 rows: caller falls back
	if (GP * NL < 32) GP = 32 / NL;                       // at least one warp
	const size_t smem = fast_smem_bytes<FT>(GP);
	const int threads = GP * NL;
	unsigned grid;
	if (dir == 0) grid = (unsigned)L.ny * (unsigned)((L.nz + NL - 1) / NL);
	else if (dir == 1) grid = (unsigned)L.nx * (unsigned)((L.nz + NL - 1) / NL);
	else grid = (unsigned)L.nx * (unsigned)((L.ny + NL - 1) / NL);
	static bool attr_set[2][3] = {};
	const int fi = sizeof(FT) == 4 ? 0 : 1;
	auto set_attr = [&](const void *fn) {
		if (!attr_set[fi][dir]) {
			cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes<FT>(64));
			attr_set[fi][dir] = true;
		}
	};
	switch (dir) {
	case 0: set_attr((const void *)k_fast_sweep<FT, 0>); k_fast_sweep<FT, 0><<<grid, threads, smem, s>>>(A, G, GP); break;
	case 1: set_attr((const void *)k_fast_sweep<FT, 1>); k_fast_sweep<FT, 1><<<grid, threads, smem, s>>>(A, G, GP); break;
	default: set_attr((const void *)k_fast_sweep<FT, 2>); k_fast_sweep<FT, 2><<<grid, threads, smem, s>>>(A, G, GP); break;
	}
	if (launches) (*launches)++;
	return true;
}

template bool launch_fast_sweep<float>(int, const SweepArgs<float> &, cudaStream_t, long long *);
template bool launch_fast_sweep<double>(int, const SweepArgs<double> &, cudaStream_t, long long *);

// ---- standalone batched solver with the same partition + PCR machinery (unit tests) -------------------------
// One CTA per NL systems; system-major host layout.
template <typename FT>
__global__ void __launch_bounds__(512, 1) k_pcr_batch(int nsys, int n, int GP, const FT *a, const FT *b, const FT *c, const FT *d, FT *x)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	FT *sys = reinterpret_cast<FT *>(smem_raw);
	FT *head = sys + 2 * 5 * GP * NL;
	FT *sol = head + 5 * GP * NL;
	const int t = threadIdx.x;
	const int g = t % GP, l = t / GP;
	const int sidx = blockIdx.x * NL + l;
	const bool ok = sidx < nsys;
	const size_t base = (size_t)(ok ? sidx : 0) * n;
	const int r0 = g * M, e = g * NL + l, stride_s = GP * NL;
	FT cp[M], lp[M], dp[M], b7 = FT(1);
#pragma unroll
	for (int i = 0; i < M; i++) {
		const int r = r0 + i;
		FT ra = FT(0), rb = FT(1), rc = FT(0), rd = FT(0);
		if (ok && r < n) { ra = r == 0 ? FT(0) : a[base + r]; rb = b[base + r]; rc = r == n - 1 ? FT(0) : c[base + r]; rd = d[base + r]; }
		if (i == M - 1) { lp[i] = ra; cp[i] = rc; b7 = rb; dp[i] = rd; }
		else if (i == 0) { const FT rr = rcp<FT>(rb); cp[0] = rc * rr; lp[0] = ra * rr; dp[0] = rd * rr; }
		else { const FT rr = rcp<FT>(rb - ra * cp[i - 1]); cp[i] = rc * rr; lp[i] = -ra * lp[i - 1] * rr; dp[i] = (rd - ra * dp[i - 1]) * rr; }
	}
	FT y0 = dp[M - 2], v0 = lp[M - 2], w0 = cp[M - 2];
#pragma unroll
	for (int i = M - 3; i >= 0; i--) { y0 = dp[i] - cp[i] * y0; v0 = lp[i] - cp[i] * v0; w0 = -cp[i] * w0; }
	head[0 * stride_s + e] = y0; head[3 * stride_s + e] = v0; head[4 * stride_s + e] = w0;
	__syncthreads();
	const bool has_next = g + 1 < GP;
	const int en = e + NL;
	const FT ny0 = has_next ? head[0 * stride_s + en] : FT(0), nv = has_next ? head[3 * stride_s + en] : FT(0), nw = has_next ? head[4 * stride_s + en] : FT(0);
	const FT a7 = lp[M - 1], c7 = cp[M - 1];
	const FT rr = rcp<FT>(b7 - a7 * cp[M - 2] - c7 * nv);
	FT Rd[1] = {(dp[M - 1] - a7 * dp[M - 2] - c7 * ny0) * rr}, E[1];
	pcr_solve<FT, 1>(sys, GP, g, l, true, -a7 * lp[M - 2] * rr, -c7 * nw * rr, Rd, E);
	sol[(g + 1) * NL + l] = E[0];
	if (g == 0) sol[l] = FT(0);
	__syncthreads();
	const FT El = sol[g * NL + l];
	FT xx[M];
	xx[M - 1] = E[0];
#pragma unroll
	for (int i = M - 2; i >= 0; i--) xx[i] = dp[i] - lp[i] * El - cp[i] * xx[i + 1];
#pragma unroll
	for (int i = 0; i < M; i++) if (ok && r0 + i < n) x[base + r0 + i] = xx[i];
}

template <typename FT>
bool launch_pcr_batch(int nsys, int n, const FT *a, const FT *b, const FT *c, const FT *d, FT *x, cudaStream_t s)
{
	const int G = (n + M - 1) / M;
	int GP = 1;
	while (GP < G) GP <<= 1;
	if (GP * NL > 512) return false;
	if (GP * NL < 32) GP = 32 / NL;
	cudaFuncSetAttribute((const void *)k_pcr_batch<FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes<FT>(64));
	k_pcr_batch<FT><<<(nsys + NL - 1) / NL, GP * NL, fast_smem_bytes<FT>(GP), s>>>(nsys, n, GP, a, b, c, d, x);
	return true;
}
template bool launch_pcr_batch<float>(int, int, const float *, const float *, const float *, const float *, float *, cudaStream_t);
template bool launch_pcr_batch<double>(int, int, const double *, const double *, const double *, const double *, double *, cudaStream_t);

} // namespace cmc
