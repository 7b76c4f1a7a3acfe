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

// Reciprocal without the IEEE division slow path: hardware seed + Newton steps (fp64: MUFU.RCP64H seed, two
// fused Newton iterations -> < 1 ulp for the well-scaled pivots of a diagonally dominant system).
template <typename FT> __device__ __forceinline__ FT rcp(FT x);
template <> __device__ __forceinline__ float rcp<float>(float x) { return __frcp_rn(x); }
template <> __device__ __forceinline__ double rcp<double>(double x)
{
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	double e = fma(-x, r, 1.0);
	r = fma(r, e, r);
	e = fma(-x, r, 1.0);
	r = fma(r, e, r);
	return r;
}

// Division-free constants of one sweep (fast mode multiplies by reciprocals; exact mode keeps the reference's
// divisions, see rows.cuh).
template <typename FT>
struct FastConst {
	FT inv2h;            // 1 / (2 h_D)
	FT inv2hx, inv2hy, inv2hz;
	FT vis_v, vis_T, b_v, b_T;
	FT c3dt;             // 3 / dt
	FT v_T, t_phi;
	__device__ __forceinline__ void init(const SweepArgs<FT> &A, int dir)
	{
		const FT h = A.h[dir];
		inv2h = FT(1) / (2 * h);
		inv2hx = FT(1) / (2 * A.h[0]); inv2hy = FT(1) / (2 * A.h[1]); inv2hz = FT(1) / (2 * A.h[2]);
		vis_v = A.v_vis / (h * h); vis_T = A.t_vis / (h * h);
		c3dt = 3 / A.dt;
		b_v = c3dt + 2 * vis_v; b_T = c3dt + 2 * vis_T;
		v_T = A.v_T; t_phi = A.t_phi;
	}
};

// RHS of the temperature row: cur.T*3/dt + t_phi * DissFunc_D(temp) (TimeLayer3D.h:554-588)
template <typename FT, int DIR>
__device__ __forceinline__ FT temperature_rhs(const SweepArgs<FT> &A, const FastConst<FT> &K, long long id, long long sx, long long sy, long long sz)
{
	const FT *tu = A.temp[0], *tv = A.temp[1], *tw = A.temp[2];
	FT diss;
	if (DIR == 0) {
		const FT u_x = (tu[id + sx] - tu[id - sx]) * K.inv2hx, v_x = (tv[id + sx] - tv[id - sx]) * K.inv2hx, w_x = (tw[id + sx] - tw[id - sx]) * K.inv2hx;
		const FT u_y = (tu[id + sy] - tu[id - sy]) * K.inv2hy, u_z = (tu[id + sz] - tu[id - sz]) * K.inv2hz;
		diss = 2 * u_x * u_x + v_x * v_x + w_x * w_x + v_x * u_y + w_x * u_z;
	} else if (DIR == 1) {
		const FT u_y = (tu[id + sy] - tu[id - sy]) * K.inv2hy, v_y = (tv[id + sy] - tv[id - sy]) * K.inv2hy, w_y = (tw[id + sy] - tw[id - sy]) * K.inv2hy;
		const FT v_x = (tv[id + sx] - tv[id - sx]) * K.inv2hx, v_z = (tv[id + sz] - tv[id - sz]) * K.inv2hz;
		diss = u_y * u_y + 2 * v_y * v_y + w_y * w_y + u_y * v_x + w_y * v_z;
	} else {
		const FT u_z = (tu[id + sz] - tu[id - sz]) * K.inv2hz, v_z = (tv[id + sz] - tv[id - sz]) * K.inv2hz, w_z = (tw[id + sz] - tw[id - sz]) * K.inv2hz;
		const FT w_x = (tw[id + sx] - tw[id - sx]) * K.inv2hx, w_y = (tw[id + sy] - tw[id - sy]) * K.inv2hy;
		diss = u_z * u_z + v_z * v_z + 2 * w_z * w_z + u_z * w_x + v_z * w_y;
	}
	return A.cur[3][id] * K.c3dt + K.t_phi * diss;
}

// ---- PCR over the reduced systems of a CTA ---------------------------------------------------------------
// Rows are normalised (B == 1).  sys[(slot*NR + r) * GP * NL + g * NL + l], slot in {0,1} ping-pong.
// NR = 2 + NRHS (A, C, D...).  Element (g, l) = chunk g of line l.  GP = G rounded up to a power of two
// (rows >= G are identity rows).
template <typename FT, int NRHS>
__device__ __forceinline__ void pcr_solve(FT *sys, int GP, int g, int e, int gstep, FT Ain, FT Cin, const FT (&Din)[NRHS], FT (&X)[NRHS])
{
	constexpr int NR = 2 + NRHS;
	const int stride = GP * NL;
	FT A = Ain, Cc = Cin, D[NRHS];
#pragma unroll
	for (int q = 0; q < NRHS; q++) D[q] = Din[q];
	int buf = 0;
	for (int s = 1; s < GP; s <<= 1) {
		FT *w = sys + buf * NR * stride;
		w[0 * stride + e] = A; w[1 * stride + e] = Cc;
#pragma unroll
		for (int q = 0; q < NRHS; q++) w[(2 + q) * stride + e] = D[q];
		__syncthreads();
		const bool lo = g - s >= 0, hi = g + s < GP;
		const int el = e - s * gstep, eh = e + s * gstep;
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
		buf ^= 1;
	}
#pragma unroll
	for (int q = 0; q < NRHS; q++) X[q] = D[q];
}

// ---- line I/O: 8 consecutive rows of this thread's chunk ----------------------------------------------------
// X / Y sweeps: rows are `stride` apart, lanes of a warp sit on neighbouring k (coalesced 64-byte segments).
// Z sweep: the 8 rows are 8 contiguous elements (64 bytes in fp64) -> 128-bit vector accesses.
template <typename FT> struct Vec16;
template <> struct Vec16<double> { typedef double2 type; static constexpr int N = 2; };
template <> struct Vec16<float> { typedef float4 type; static constexpr int N = 4; };

template <typename FT, int DIR>
__device__ __forceinline__ void load8(const FT *__restrict__ p, long long base, long long stride, int r0, int n, FT (&o)[M])
{
	if (DIR == 2) {
		typedef typename Vec16<FT>::type V;
		constexpr int N = Vec16<FT>::N;
		const V *q = reinterpret_cast<const V *>(p + base + r0);     // r0 % 8 == 0 and lines are 128-byte aligned
#pragma unroll
		for (int v = 0; v < M / N; v++) {
			const V t = q[v];
			const FT *e = reinterpret_cast<const FT *>(&t);
#pragma unroll
			for (int k = 0; k < N; k++) o[v * N + k] = e[k];
		}
	} else {
#pragma unroll
		for (int i = 0; i < M; i++) {
			const int r = min(r0 + i, n - 1);                         // rows past the end: any valid address (value unused)
			o[i] = p[base + (long long)r * stride];
		}
	}
}

template <typename FT, int DIR>
__device__ __forceinline__ void store8(FT *__restrict__ p, long long base, long long stride, int r0, int n, const FT (&v)[M])
{
	if (DIR == 2) {
		typedef typename Vec16<FT>::type V;
		constexpr int N = Vec16<FT>::N;
		V *q = reinterpret_cast<V *>(p + base + r0);                  // the padded tail of a z-line may be written freely
#pragma unroll
		for (int w = 0; w < M / N; w++) {
			V t;
			FT *e = reinterpret_cast<FT *>(&t);
#pragma unroll
			for (int k = 0; k < N; k++) e[k] = v[w * N + k];
			q[w] = t;
		}
	} else {
#pragma unroll
		for (int i = 0; i < M; i++)
			if (r0 + i < n) p[base + (long long)(r0 + i) * stride] = v[i];
	}
}

// rows r0-1 and r0+8 (clamped into the line; only interior rows use them and their neighbours always exist)
template <typename FT>
__device__ __forceinline__ void load_ends(const FT *__restrict__ p, long long base, long long stride, int r0, int n, FT &lo, FT &hi)
{
	lo = p[base + (long long)max(r0 - 1, 0) * stride];
	hi = p[base + (long long)min(r0 + M, n - 1) * stride];
}

// central difference along the line for the 8 rows of a chunk
template <typename FT>
__device__ __forceinline__ FT cdiff(const FT (&f)[M], FT lo, FT hi, int i, FT inv2h)
{
	const FT m = i == 0 ? lo : f[i == 0 ? 0 : i - 1], p = i == M - 1 ? hi : f[i == M - 1 ? M - 1 : i + 1];
	return (p - m) * inv2h;
}

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
	// a chunk takes part in the loads/stores when it overlaps the line (z-lines: the 128-byte padded line)
	const bool chunk_ok = line_ok && r0 < (DIR == 2 ? L.nzp : n);
	FastConst<FT> K; K.init(A, DIR);

	// shared-memory element of (chunk g, line l): consecutive lanes -> consecutive addresses in both mappings
	const int e = DIR == 2 ? l * GP + g : g * NL + l;
	const int e_next = DIR == 2 ? e + 1 : e + NL;      // chunk g+1 of the same line
	const int e_prev = DIR == 2 ? e - 1 : e - NL;
	const int gstep = DIR == 2 ? 1 : NL;               // shared-memory distance of neighbouring chunks
	const int stride_s = GP * NL;

	// roles of the chunk's rows (rows >= n or lines outside the grid: no segment, no store)
	unsigned role[M];
	{
		uint8_t rb[M];
#pragma unroll
		for (int i = 0; i < M; i++) rb[i] = 0;
		if (chunk_ok) {
			if (DIR == 2) {
				const uint2 w = *reinterpret_cast<const uint2 *>(A.role + base + r0);
#pragma unroll
				for (int i = 0; i < 4; i++) { rb[i] = (uint8_t)(w.x >> (8 * i)); rb[4 + i] = (uint8_t)(w.y >> (8 * i)); }
			} else {
#pragma unroll
				for (int i = 0; i < M; i++) rb[i] = A.role[base + (long long)min(r0 + i, n - 1) * stride];
			}
		}
#pragma unroll
		for (int i = 0; i < M; i++) role[i] = (r0 + i < n) ? (unsigned)rb[i] : 0u;
	}

	// ======================================= phase V: u, v, w ==============================================
	FT cp[M], lp[M], dp[3][M];
	FT b7 = FT(1);
	{
		FT V[M], Tl[M], Tlo = FT(0), Thi = FT(0);
#pragma unroll
		for (int i = 0; i < M; i++) { V[i] = FT(0); Tl[i] = FT(0); dp[0][i] = FT(0); dp[1][i] = FT(0); dp[2][i] = FT(0); }
		if (chunk_ok) {
			load8<FT, DIR>(A.temp[DIR], base, stride, r0, n, V);
			load8<FT, DIR>(A.cur[0], base, stride, r0, n, dp[0]);
			load8<FT, DIR>(A.cur[1], base, stride, r0, n, dp[1]);
			load8<FT, DIR>(A.cur[2], base, stride, r0, n, dp[2]);
			load8<FT, DIR>(A.temp[3], base, stride, r0, n, Tl);
			load_ends<FT>(A.temp[3], base, stride, r0, n, Tlo, Thi);
		}
#pragma unroll
		for (int i = 0; i < M; i++) {
			const unsigned r = role[i];
			const bool is_int = r & R_INT, is_bc = r & (R_START | R_END), vfree = r & R_VFREE;
			FT bd0 = FT(0), bd1 = FT(0), bd2 = FT(0);
			if (is_bc && !vfree) {        // no-slip boundary row: value of the node (rare: two rows per segment)
				const long long id = base + (long long)(r0 + i) * stride;
				bd0 = A.nodev[0][id]; bd1 = A.nodev[1][id]; bd2 = A.nodev[2][id];
			}
			const FT Vh = V[i] * K.inv2h;
			const FT a = is_int ? -Vh - K.vis_v : ((r & R_END) && vfree ? FT(-1) : FT(0));
			const FT c = is_int ? Vh - K.vis_v : ((r & R_START) && vfree ? FT(-1) : FT(0));
			const FT b = is_int ? K.b_v : (is_bc && vfree ? FT(2) : FT(1));
			FT d[3];
			d[0] = is_int ? dp[0][i] * K.c3dt : bd0;
			d[1] = is_int ? dp[1][i] * K.c3dt : bd1;
			d[2] = is_int ? dp[2][i] * K.c3dt : bd2;
			if (is_int) d[DIR] -= K.v_T * cdiff<FT>(Tl, Tlo, Thi, i, K.inv2h);
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
	}
	FT E[3];
	{
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
		const bool has_next = g + 1 < GP;
		const FT ny0 = has_next ? head[0 * stride_s + e_next] : FT(0), ny1 = has_next ? head[1 * stride_s + e_next] : FT(0),
		         ny2 = has_next ? head[2 * stride_s + e_next] : FT(0);
		const FT nv = has_next ? head[3 * stride_s + e_next] : FT(0), nw = has_next ? head[4 * stride_s + e_next] : FT(0);
		const FT a7 = lp[M - 1], c7 = cp[M - 1];
		const FT rr = rcp<FT>(b7 - a7 * cp[M - 2] - c7 * nv);
		Ra = -a7 * lp[M - 2] * rr; Rc = -c7 * nw * rr;
		Rd[0] = (dp[0][M - 1] - a7 * dp[0][M - 2] - c7 * ny0) * rr;
		Rd[1] = (dp[1][M - 1] - a7 * dp[1][M - 2] - c7 * ny1) * rr;
		Rd[2] = (dp[2][M - 1] - a7 * dp[2][M - 2] - c7 * ny2) * rr;
		pcr_solve<FT, 3>(sys, GP, g, e, gstep, Ra, Rc, Rd, E);
		// publish separator solutions for the chunk on the right
		sol[0 * stride_s + e] = E[0]; sol[1 * stride_s + e] = E[1]; sol[2 * stride_s + e] = E[2];
		__syncthreads();
	}
	// back substitution, store u, v, w and the relaxed linearisation layer
#pragma unroll
	for (int q = 0; q < 3; q++) {
		const FT El = g > 0 ? sol[q * stride_s + e_prev] : FT(0);
		FT x[M], tq[M];
		x[M - 1] = E[q];
#pragma unroll
		for (int i = M - 2; i >= 0; i--) x[i] = dp[q][i] - lp[i] * El - cp[i] * x[i + 1];
		if (chunk_ok) {
			load8<FT, DIR>(A.temp[q], base, stride, r0, n, tq);
			bool all_seg = true, hole = false;
#pragma unroll
			for (int i = 0; i < M; i++) {
				const bool seg = role[i] & R_SEG, in = role[i] & R_IN;
				all_seg &= seg || (DIR == 2 && r0 + i >= n);
				hole |= in && !seg;
			}
			if (hole) {                       // fluid cell outside every segment (dropped run): merge with the old `next`
#pragma unroll
				for (int i = 0; i < M; i++)
					if ((role[i] & R_IN) && !(role[i] & R_SEG)) x[i] = A.next[q][base + (long long)(r0 + i) * stride];
			}
#pragma unroll
			for (int i = 0; i < M; i++) tq[i] = (role[i] & R_IN) ? (tq[i] + x[i]) * FT(0.5) : tq[i];
			store8<FT, DIR>(A.temp_out[q], base, stride, r0, n, tq);
			if (all_seg) store8<FT, DIR>(A.next[q], base, stride, r0, n, x);
			else {
#pragma unroll
				for (int i = 0; i < M; i++)
					if (role[i] & R_SEG) A.next[q][base + (long long)(r0 + i) * stride] = x[i];
			}
		}
	}

	// ======================================= phase T ======================================================
	__syncthreads();                // head / sol / sys are reused
	FT (&dT)[M] = dp[0];
	{
		FT V[M], cT[M];
#pragma unroll
		for (int i = 0; i < M; i++) { V[i] = FT(0); cT[i] = FT(0); }
		FT diss[M];
#pragma unroll
		for (int i = 0; i < M; i++) diss[i] = FT(0);
		if (chunk_ok) {
			load8<FT, DIR>(A.temp[DIR], base, stride, r0, n, V);
			load8<FT, DIR>(A.cur[3], base, stride, r0, n, cT);
			// dissipation function of the sweep direction (TimeLayer3D.h:554-588): derivatives along the line of
			// u, v, w plus the two cross-line derivatives of the component aligned with the sweep
			FT f[M], lo, hi, d_u[M], d_v[M], d_w[M];
			load8<FT, DIR>(A.temp[0], base, stride, r0, n, f); load_ends<FT>(A.temp[0], base, stride, r0, n, lo, hi);
#pragma unroll
			for (int i = 0; i < M; i++) d_u[i] = cdiff<FT>(f, lo, hi, i, K.inv2h);
			load8<FT, DIR>(A.temp[1], base, stride, r0, n, f); load_ends<FT>(A.temp[1], base, stride, r0, n, lo, hi);
#pragma unroll
			for (int i = 0; i < M; i++) d_v[i] = cdiff<FT>(f, lo, hi, i, K.inv2h);
			load8<FT, DIR>(A.temp[2], base, stride, r0, n, f); load_ends<FT>(A.temp[2], base, stride, r0, n, lo, hi);
#pragma unroll
			for (int i = 0; i < M; i++) d_w[i] = cdiff<FT>(f, lo, hi, i, K.inv2h);
			// cross-line derivatives of temp[DIR] in the two other directions
			const long long s1 = DIR == 0 ? sy : sx, s2 = DIR == 2 ? sy : sz;
			const FT i1 = DIR == 0 ? K.inv2hy : K.inv2hx, i2 = DIR == 2 ? K.inv2hy : K.inv2hz;
			FT p1[M], m1[M], p2[M], m2[M];
			const bool any_int = (role[0] | role[1] | role[2] | role[3] | role[4] | role[5] | role[6] | role[7]) & R_INT;
			if (any_int) {          // neighbouring lines exist around every interior cell
				if (DIR == 2) {
					load8<FT, DIR>(A.temp[DIR], base + s1, stride, r0, n, p1); load8<FT, DIR>(A.temp[DIR], base - s1, stride, r0, n, m1);
					load8<FT, DIR>(A.temp[DIR], base + s2, stride, r0, n, p2); load8<FT, DIR>(A.temp[DIR], base - s2, stride, r0, n, m2);
				} else {
					load8<FT, DIR>(A.temp[DIR], base + s1, stride, r0, n, p1); load8<FT, DIR>(A.temp[DIR], base - s1, stride, r0, n, m1);
					load8<FT, DIR>(A.temp[DIR], base + s2, stride, r0, n, p2); load8<FT, DIR>(A.temp[DIR], base - s2, stride, r0, n, m2);
				}
#pragma unroll
				for (int i = 0; i < M; i++) {
					const FT c1 = (p1[i] - m1[i]) * i1, c2 = (p2[i] - m2[i]) * i2;
					// X: 2 u_x^2 + v_x^2 + w_x^2 + v_x u_y + w_x u_z ; Y: u_y^2 + 2 v_y^2 + w_y^2 + u_y v_x + w_y v_z ;
					// Z: u_z^2 + v_z^2 + 2 w_z^2 + u_z w_x + v_z w_y
					FT s;
					if (DIR == 0) s = 2 * d_u[i] * d_u[i] + d_v[i] * d_v[i] + d_w[i] * d_w[i] + d_v[i] * c1 + d_w[i] * c2;
					else if (DIR == 1) s = d_u[i] * d_u[i] + 2 * d_v[i] * d_v[i] + d_w[i] * d_w[i] + d_u[i] * c1 + d_w[i] * c2;
					else s = d_u[i] * d_u[i] + d_v[i] * d_v[i] + 2 * d_w[i] * d_w[i] + d_u[i] * c1 + d_v[i] * c2;
					diss[i] = s;
				}
			}
		}
#pragma unroll
		for (int i = 0; i < M; i++) {
			const unsigned r = role[i];
			const bool is_int = r & R_INT, is_bc = r & (R_START | R_END), tfree = r & R_TFREE;
			FT bd = FT(0);
			if (is_bc && !tfree) bd = A.nodev[3][base + (long long)(r0 + i) * stride];
			const FT Vh = V[i] * K.inv2h;
			const FT a = is_int ? -Vh - K.vis_T : ((r & R_END) && tfree ? FT(-1) : FT(0));
			const FT c = is_int ? Vh - K.vis_T : ((r & R_START) && tfree ? FT(-1) : FT(0));
			const FT b = is_int ? K.b_T : (is_bc && tfree ? FT(2) : FT(1));
			const FT d = is_int ? cT[i] * K.c3dt + K.t_phi * diss[i] : bd;
			if (i == M - 1) { lp[i] = a; cp[i] = c; b7 = b; dT[i] = d; }
			else if (i == 0) { const FT rr = rcp<FT>(b); cp[0] = c * rr; lp[0] = a * rr; dT[0] = d * rr; }
			else {
				const FT rr = rcp<FT>(b - a * cp[i - 1]);
				cp[i] = c * rr; lp[i] = -a * lp[i - 1] * rr; dT[i] = (d - a * dT[i - 1]) * rr;
			}
		}
	}
	{
		FT y0 = dT[M - 2], v0 = lp[M - 2], w0 = cp[M - 2];
#pragma unroll
		for (int i = M - 3; i >= 0; i--) { y0 = dT[i] - cp[i] * y0; v0 = lp[i] - cp[i] * v0; w0 = -cp[i] * w0; }
		head[0 * stride_s + e] = y0; head[3 * stride_s + e] = v0; head[4 * stride_s + e] = w0;
		__syncthreads();
		const bool has_next = g + 1 < GP;
		const FT ny0 = has_next ? head[0 * stride_s + e_next] : FT(0);
		const FT nv = has_next ? head[3 * stride_s + e_next] : FT(0), nw = has_next ? head[4 * stride_s + e_next] : FT(0);
		const FT a7 = lp[M - 1], c7 = cp[M - 1];
		const FT rr = rcp<FT>(b7 - a7 * cp[M - 2] - c7 * nv);
		FT Rd[1] = {(dT[M - 1] - a7 * dT[M - 2] - c7 * ny0) * rr}, ET[1];
		pcr_solve<FT, 1>(sys, GP, g, e, gstep, -a7 * lp[M - 2] * rr, -c7 * nw * rr, Rd, ET);
		sol[e] = ET[0];
		__syncthreads();
		const FT El = g > 0 ? sol[e_prev] : FT(0);
		FT x[M], tq[M];
		x[M - 1] = ET[0];
#pragma unroll
		for (int i = M - 2; i >= 0; i--) x[i] = dT[i] - lp[i] * El - cp[i] * x[i + 1];
		if (chunk_ok) {
			load8<FT, DIR>(A.temp[3], base, stride, r0, n, tq);
			bool all_seg = true, hole = false;
#pragma unroll
			for (int i = 0; i < M; i++) {
				const bool seg = role[i] & R_SEG, in = role[i] & R_IN;
				all_seg &= seg || (DIR == 2 && r0 + i >= n);
				hole |= in && !seg;
			}
			if (hole) {
#pragma unroll
				for (int i = 0; i < M; i++)
					if ((role[i] & R_IN) && !(role[i] & R_SEG)) x[i] = A.next[3][base + (long long)(r0 + i) * stride];
			}
#pragma unroll
			for (int i = 0; i < M; i++) tq[i] = (role[i] & R_IN) ? (tq[i] + x[i]) * FT(0.5) : tq[i];
			store8<FT, DIR>(A.temp_out[3], base, stride, r0, n, tq);
			if (all_seg) store8<FT, DIR>(A.next[3], base, stride, r0, n, x);
			else {
#pragma unroll
				for (int i = 0; i < M; i++)
					if (role[i] & R_SEG) A.next[3][base + (long long)(r0 + i) * stride] = x[i];
			}
		}
	}
}

template <typename FT>
static size_t fast_smem_bytes(int GP) { return sizeof(FT) * (size_t)(2 * 5 * GP * NL + 5 * GP * NL + 3 * GP * NL); }

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
	pcr_solve<FT, 1>(sys, GP, g, e, NL, -a7 * lp[M - 2] * rr, -c7 * nw * rr, Rd, E);
	sol[e] = E[0];
	__syncthreads();
	const FT El = g > 0 ? sol[e - NL] : FT(0);
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
