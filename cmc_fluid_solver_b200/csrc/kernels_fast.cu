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
//   * the separators form a reduced tridiagonal system of GP = n/8 unknowns per line, solved in shared
//     memory by cyclic reduction down to 8 rows + parallel cyclic reduction (normalised rows, one
//     reciprocal per step);
//   * back substitution of the interior rows from registers, results stored once.
//
// Thread mapping.  X and Y sweeps (strided lines): a CTA owns NL = 8 neighbouring k-columns of one
// (j or i) row-set; lanes run along k first (64-byte segments per row for fp64), chunks across warps.
// Z sweep (contiguous lines): a CTA owns NL = 8 neighbouring lines; lanes run along the chunk index, so a
// warp reads 32 consecutive 64-byte chunks = 2 KB of one line with 128-bit loads.
// HBM traffic per cell and sweep is the algorithmic 16 values + 1 descriptor byte: stencil neighbours and
// re-reads for the merge are served by L1/L2.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kernels.h"
#include "fast_core.cuh"

namespace cmc {

// MODE 0: the complete sweep of a slab whose lines do not leave the slab (Y, Z always; X on a single GPU).
// x-sweep of a slab-decomposed grid (partitioned solve across GPUs, see dist.h):
// MODE 1: "spike" pass - eliminate the slab's rows and emit, per line, how its first / last row depend on the
//         neighbouring slabs' adjacent rows (16 coefficients); nothing is stored into the layers;
// MODE 2: the complete sweep with the neighbours' (now known) adjacent-row solutions folded into the first /
//         last row of the slab.
template <typename FT, int DIR, int GP, int NL, int MODE>
__global__ void __launch_bounds__(GP * NL, (GP * NL <= 64 && DIR == 2 && GP == 64) ? 8 : (GP * NL <= 128 && DIR == 2 && GP == 64) ? 4 : (GP * NL <= 256) ? 2 : 1) k_fast_sweep(const SweepArgs<FT> A, const FastConst<FT> K, long long *trace, const int one)
{
	static_assert(MODE == 0 || DIR == 0, "slab coupling exists along x only");
#define CMC_MARK(k) do { if (trace) { __syncthreads(); if (threadIdx.x == 0) trace[(size_t)blockIdx.x * 16 + (k)] = clock64(); } } while (0)
	CMC_MARK(0);
	constexpr int STR = GP * NL;
	constexpr int GS = DIR == 2 ? 1 : NL;          // shared-memory distance of neighbouring chunks of a line
	extern __shared__ __align__(16) unsigned char smem_raw[];
	constexpr int NRMAX = MODE == 1 ? 7 : 5;       // matrix (2) + right-hand sides of the widest reduced solve
	FT *sys = reinterpret_cast<FT *>(smem_raw);    // reduced-solve scratch (CR publications + compacted PCR ping-pong)
	FT *sol = sys;                                 // separator solutions: alias the CR publications (see reduced_solve)
	FT *head = sys + reduced_scratch_elems<NRMAX - 2, GP, NL>();   // 5 arrays: y0[3], v0, w0 of every chunk

	const Layout &L = A.L;
	const int t = threadIdx.x;
	int g, l;                                      // chunk, line-in-CTA
	if (DIR == 2) { g = t % GP; l = t / GP; } else { l = t % NL; g = t / NL; }
	const int e = t;                               // == l * GP + g (Z) or g * NL + l (X, Y): lanes -> consecutive elements

	// ---- which line ----------------------------------------------------------------------------------------
	long long base;                 // element index of row 0 of this thread's line
	long long stride;               // along the line (y lines: inside one y-block, see rowoff)
	int jline = 0;                  // x / z lines: the j-row of the line (cross-line neighbours j +- 1)
	int n;                          // rows of the line
	bool line_ok;
	if (DIR == 0) {                 // lines along x: CTA = (j, k-tile)
		const int ktiles = (L.nz + NL - 1) / NL;
		const int j = blockIdx.x / ktiles, k = (blockIdx.x % ktiles) * NL + l;
		line_ok = k < L.nz; n = L.nx; stride = L.plane; base = L.idx(0, j, line_ok ? k : 0); jline = j;
	} else if (DIR == 1) {          // lines along y: CTA = (i, k-tile)
		const int ktiles = (L.nz + NL - 1) / NL;
		const int i = blockIdx.x / ktiles, k = (blockIdx.x % ktiles) * NL + l;
		line_ok = k < L.nz; n = L.ny; stride = L.nzp; base = L.idx(i, 0, line_ok ? k : 0);
	} else {                        // lines along z: CTA = (i, j-tile)
		const int jtiles = (L.ny + NL - 1) / NL;
		const int i = blockIdx.x / jtiles, j = (blockIdx.x % jtiles) * NL + l;
		line_ok = j < L.ny; n = L.nz; stride = 1; base = L.idx(i, line_ok ? j : 0, 0); jline = line_ok ? j : 0;
	}
	const int r0 = g * M;           // first row of this chunk
	int pi = 0;                     // y / z lines: the x-plane this CTA works in
	if (DIR == 1) pi = blockIdx.x / ((L.nz + NL - 1) / NL);
	if (DIR == 2) pi = blockIdx.x / ((L.ny + NL - 1) / NL);
	// slab-coupled x-lines: global line id, its owner rank for the interface solve, and the first / last chunk
	int xo = 0;                     // element offset of this line's interface values: owner * 8 * lpo + line_in_owner
	FT *cto = nullptr;              // MODE 1: this line's slot in its owner's coefficient table
	// chunks that hold real rows (every slab but the last holds a multiple of 8 planes; the ragged last chunk of the last
	// slab ends in identity rows and has no neighbour above, so its "last row" coupling is zero by construction)
	const int GL = (MODE != 0) ? (A.L.nx + M - 1) / M : GP;
	if (MODE != 0) {
		const int ktiles = (L.nz + NL - 1) / NL;
		const int j = blockIdx.x / ktiles, k = min((blockIdx.x % ktiles) * NL + l, L.nz - 1);
		const int line = j * L.nz + k, owner = line / A.lpo;
		xo = owner * A.lpo * 8 + (line - owner * A.lpo);
		if (MODE == 1) cto = A.xcoef_to[owner] + (line - owner * A.lpo);
	}
	// Row offsets, clamped into the line so that EVERY thread issues valid (if redundant) loads: threads of padding
	// chunks / lines outside the grid see role 0 everywhere, compute identity rows and store nothing.
	// offset of row r of a line from its row 0: a y-line jumps to the next y-block every 2^jbs rows
	auto rowoff = [&](int r) -> int {
		return DIR == 1 ? (r >> L.jbs) * (int)L.bstride + (r & L.jbm) * (int)L.nzp : r * (int)stride;
	};
	int off[M], off_lo, off_hi;      // 32-bit element offsets (launch_fast_sweep checks total < 2^31)
	if (DIR == 2) {
		const int rc = min(r0, L.nzp - M);
#pragma unroll
		for (int i = 0; i < M; i++) off[i] = (int)base + rc + i;
		off_lo = (int)base + max(min(r0 - 1, n - 1), 0);
		off_hi = (int)base + min(r0 + M, n - 1);
	} else {
		// (a chunk never straddles a y-block: one block offset per chunk, rows nzp apart inside it)
		const int rc = min(r0, n - 1), off0 = (int)base + rowoff(rc), rmax = n - 1 - rc;
#pragma unroll
		for (int i = 0; i < M; i++) off[i] = off0 + min(i, rmax) * (int)stride;
		// rows -1 and n exist for a slab inside a decomposed grid (halo planes): its first / last row can be interior
		// (padding chunks - r0 past the end of the line - must stay inside the buffer too: along x a row is a whole plane)
		off_lo = (int)base + rowoff(max(min(r0 - 1, n - 1), MODE != 0 ? -1 : 0));
		off_hi = (int)base + rowoff(min(r0 + M, MODE != 0 ? n : n - 1));
	}
	// ptxas schedules inside basic blocks only.  These never-taken branches (`one` is always 1) split the kernel where
	// its latency hiding wants it split - address setup | all loads of a phase | arithmetic; without them the generated
	// code for the 512-thread kernels is 8 % slower (measured, profiles/r01_variants.md).
#define CMC_SCHED_FENCE() do { if (one > 1) asm volatile("nanosleep.u32 1;"); } while (0)
	CMC_SCHED_FENCE();
	const int step_last = off_hi - off[M - 1];   // from the chunk's last row to the next row of the line (may cross a y-block)
	unsigned rowmask = 0;           // rows of this chunk that exist
#pragma unroll
	for (int i = 0; i < M; i++) rowmask |= (line_ok && r0 + i < n) ? (1u << i) : 0u;

	// roles of the chunk's rows
	unsigned rw0 = 0, rw1 = 0;      // role bytes of rows 0-3 / 4-7, packed
	if (DIR == 2) {
		const uint2 w = *reinterpret_cast<const uint2 *>(A.role + off[0]);
		rw0 = w.x; rw1 = w.y;
	} else {
#pragma unroll
		for (int i = 0; i < 4; i++) { rw0 |= (unsigned)A.role[off[i]] << (8 * i); rw1 |= (unsigned)A.role[off[4 + i]] << (8 * i); }
	}
#pragma unroll
	for (int i = 0; i < 4; i++) {
		if (!(rowmask & (1u << i))) rw0 &= ~(0xffu << (8 * i));
		if (!(rowmask & (1u << (4 + i)))) rw1 &= ~(0xffu << (8 * i));
	}
#define ROLE(i) (((i) < 4 ? rw0 >> (8 * (i)) : rw1 >> (8 * ((i) - 4))) & 0xffu)
	unsigned segmask = 0, inmask = 0;
#pragma unroll
	for (int i = 0; i < M; i++) {
		segmask |= (ROLE(i) & R_SEG) ? (1u << i) : 0u;
		inmask |= (ROLE(i) & R_IN) ? (1u << i) : 0u;
	}
	const bool any_int = ((rw0 | rw1) & (R_INT * 0x01010101u)) != 0;
	const unsigned holes = inmask & ~segmask;      // fluid cells outside every segment (dropped runs)
	// every row a plain interior row (no boundary row, no folded shared cell): the per-row special cases are skipped
	const bool plain = DIR == 2 && (((rw0 & 0x87878787u) ^ 0x01010101u) | ((rw1 & 0x87878787u) ^ 0x01010101u)) == 0u;
	// z-lines: the padded tail of the last chunk may be rewritten freely, which keeps the vector path
	const unsigned full = (DIR == 2 && line_ok && r0 < n) ? 0xffu : rowmask;
	const unsigned segfull = (segmask | (full & ~rowmask)) == 0xffu ? 0xffu : segmask;

	// ======================================= phase V: u, v, w ==============================================
	FT cp[M], lp[M], dp[3][M];
	FT b7 = FT(1), rr;
	{
		FT V[M], Tl[M], Tlo, Thi;
		// `one` is always 1: the real branch keeps every load of the phase in one basic block, ahead of the
		// arithmetic, so that they are all in flight together (ptxas does not schedule across the branch)
		if (one) {
			load8<FT, DIR>(A.temp[DIR], off, V);
			load8<FT, DIR>(A.cur[0], off, dp[0]);
			load8<FT, DIR>(A.cur[1], off, dp[1]);
			load8<FT, DIR>(A.cur[2], off, dp[2]);
			load8<FT, DIR>(A.temp[3], off, Tl);
			Tlo = A.temp[3][off_lo]; Thi = A.temp[3][off_hi];
			if (DIR != 2) CMC_SCHED_FENCE();      // (z: measured faster without this one)
		}
		// right-hand sides of interior rows, in place (rows that are not plain interior are patched below):
		//   d = cur * 3/dt  ( - v_T * dT/dD for the velocity component along the sweep )
#pragma unroll
		for (int i = 0; i < M; i++) {
			dp[0][i] *= K.c3dt; dp[1][i] *= K.c3dt; dp[2][i] *= K.c3dt;
			dp[DIR][i] -= K.v_T * cdiff<FT>(Tl, Tlo, Thi, i, K.inv2h);
		}
#pragma unroll
		for (int i = 0; i < M; i++) {
			const unsigned r = ROLE(i);
			const FT Vh = V[i] * K.inv2h;
			FT a = -Vh - K.vis_v, c = Vh - K.vis_v, b = K.b_v;
			FT d0 = dp[0][i], d1 = dp[1][i], d2 = dp[2][i];
			if (!plain && (r & (R_SEG | R_PRE)) != R_INT) {       // rare: boundary row, cell outside every segment, or shared-cell fold
				const bool vfree = r & R_VFREE;
				if (r & R_INT) {                        // R_PRE: the next cell ends this segment AND starts the next one -
					if (vfree) b += FT(0.5) * c;        // fold its ApplyBC1 row in:  -x[p-1] + 2 x[p] = 0
					else {                              //                            x[p] = node value
						const int idn = off[i] + (i == M - 1 ? step_last : (int)stride);     // the next cell along the line
						d0 -= c * A.nodev[0][idn]; d1 -= c * A.nodev[1][idn]; d2 -= c * A.nodev[2][idn];
					}
					c = FT(0);
				} else if (r & (R_START | R_END)) {     // ApplyBC0 / ApplyBC1 (a shared cell keeps its start row only)
					a = ((r & (R_END | R_START)) == R_END && vfree) ? FT(-1) : FT(0);
					c = ((r & R_START) && vfree) ? FT(-1) : FT(0);
					b = vfree ? FT(2) : FT(1);
					d0 = d1 = d2 = FT(0);
					if (!vfree) { d0 = A.nodev[0][off[i]]; d1 = A.nodev[1][off[i]]; d2 = A.nodev[2][off[i]]; }
				} else { a = FT(0); c = FT(0); b = FT(1); d0 = d1 = d2 = FT(0); }
			}
			if (MODE == 2) {        // neighbours' adjacent rows are known: move their terms to the right-hand side
				if (i == 0 && g == 0) {
					d0 -= a * A.xbnd[xo]; d1 -= a * A.xbnd[xo + A.lpo]; d2 -= a * A.xbnd[xo + 2 * A.lpo];
					a = FT(0);
				}
				if (i == M - 1 && g == GL - 1) {
					d0 -= c * A.xbnd[xo + 4 * A.lpo]; d1 -= c * A.xbnd[xo + 5 * A.lpo]; d2 -= c * A.xbnd[xo + 6 * A.lpo];
					c = FT(0);
				}
			}
			CMC_ELIM_ROW(i, a, b, c)
			if (i == M - 1) { dp[0][i] = d0; dp[1][i] = d1; dp[2][i] = d2; }
			else if (i == 0) { dp[0][0] = d0 * rr; dp[1][0] = d1 * rr; dp[2][0] = d2 * rr; }
			else {
				dp[0][i] = (d0 - a * dp[0][i - 1]) * rr;
				dp[1][i] = (d1 - a * dp[1][i - 1]) * rr;
				dp[2][i] = (d2 - a * dp[2][i - 1]) * rr;
			}
		}
	}
	CMC_MARK(1);   // phase V loads + elimination done
	FT E[3];
	{
		// coupling of the first interior row to the two separators: x_0 = y0 - v0*E(g-1) - w0*E(g)
		FT y0[3] = {dp[0][M - 2], dp[1][M - 2], dp[2][M - 2]}, v0 = lp[M - 2], w0 = cp[M - 2];
#pragma unroll
		for (int i = M - 3; i >= 0; i--) {
			y0[0] = dp[0][i] - cp[i] * y0[0]; y0[1] = dp[1][i] - cp[i] * y0[1]; y0[2] = dp[2][i] - cp[i] * y0[2];
			v0 = lp[i] - cp[i] * v0; w0 = -cp[i] * w0;
		}
		head[0 * STR + e] = y0[0]; head[1 * STR + e] = y0[1]; head[2 * STR + e] = y0[2];
		head[3 * STR + e] = v0; head[4 * STR + e] = w0;
		__syncthreads();
		// reduced row of this chunk's separator (row M-1): needs the head of chunk g+1
		const bool has_next = g + 1 < GP;
		const FT *hn = head + e + GS;
		const FT ny0 = has_next ? hn[0 * STR] : FT(0), ny1 = has_next ? hn[1 * STR] : FT(0), ny2 = has_next ? hn[2 * STR] : FT(0);
		const FT nv = has_next ? hn[3 * STR] : FT(0), nw = has_next ? hn[4 * STR] : FT(0);
		const FT a7 = lp[M - 1], c7 = cp[M - 1];
		rr = rcp<FT>(b7 - a7 * cp[M - 2] - c7 * nv);
		FT Rd[3];
		Rd[0] = (dp[0][M - 1] - a7 * dp[0][M - 2] - c7 * ny0) * rr;
		Rd[1] = (dp[1][M - 1] - a7 * dp[1][M - 2] - c7 * ny1) * rr;
		Rd[2] = (dp[2][M - 1] - a7 * dp[2][M - 2] - c7 * ny2) * rr;
		if (MODE == 1) {
			// open-ended slab: E = Y - P*x_left - Q*x_right.  P, Q are two more right-hand sides: the left spike
			// of chunk 0 and the last row's coupling c7 (its "next chunk" lives on the next slab).
			FT Re[5] = {Rd[0], Rd[1], Rd[2], FT(0), FT(0)}, Xe[5];
			FT Ra = -a7 * lp[M - 2] * rr;
			if (g == 0) { Re[3] = Ra; Ra = FT(0); }
			if (g == GL - 1) Re[4] = c7 * rr;        // nv == nw == 0 there: chunk GL is a padding (identity) chunk
			reduced_solve<FT, 5, GP, GS, NL>(sys, sol, g, e, Ra, -c7 * nw * rr, Re, Xe);
			if (line_ok && g == 0) {                 // first row: x = f - pf*x_left - qf*x_right
#pragma unroll
				for (int q = 0; q < 3; q++) cto[q * A.lpo] = y0[q] - w0 * Xe[q];
				cto[6 * A.lpo] = v0 - w0 * Xe[3];
				cto[7 * A.lpo] = -w0 * Xe[4];
			}
			if (line_ok && g == GL - 1) {            // last row (= last separator): x = l - pl*x_left - ql*x_right
#pragma unroll
				for (int q = 0; q < 3; q++) cto[(3 + q) * A.lpo] = Xe[q];
				cto[8 * A.lpo] = Xe[3];
				cto[9 * A.lpo] = Xe[4];
			}
			E[0] = E[1] = E[2] = FT(0);
		} else
			reduced_solve<FT, 3, GP, GS, NL>(sys, sol, g, e, -a7 * lp[M - 2] * rr, -c7 * nw * rr, Rd, E);    // every row's solution -> sol[]
	}
	CMC_MARK(2);   // phase V reduced solve done
	// back substitution in place (dp[q] <- x: retires cp / lp early), then store u, v, w and the relaxed
	// linearisation layer
#pragma unroll
	for (int q = 0; q < (MODE == 1 ? 0 : 3); q++) {
		const FT El = g > 0 ? sol[q * STR + e - GS] : FT(0);
		dp[q][M - 1] = E[q];
#pragma unroll
		for (int i = M - 2; i >= 0; i--) dp[q][i] = dp[q][i] - lp[i] * El - cp[i] * dp[q][i + 1];
	}
	if (DIR == 2) CMC_SCHED_FENCE();      // (measured: -0.1 ms along z, nothing along x / y)
#pragma unroll
	for (int q = 0; q < (MODE == 1 ? 0 : 3); q++) {
		FT (&x)[M] = dp[q];
		FT tq[M];
		if (one) load8<FT, DIR>(A.temp[q], off, tq);
		if (holes) {                      // merge those with the OLD value of `next` (MergeFieldTo reads whatever is there)
#pragma unroll
			for (int i = 0; i < M; i++)
				if (holes & (1u << i)) x[i] = A.next[q][off[i]];
		}
		relax8<FT, DIR>(tq, x, inmask, A.extra_merge);
		store8<FT, DIR>(A.temp_out[q], off, full, tq);
		store8<FT, DIR>(A.next[q], off, segfull, x);
		push_planes<FT, DIR, MODE>(A, q, pi, g, GL, off, full, segfull, tq, x);
	}
	CMC_MARK(3);   // phase V stores issued

	// ======================================= phase T ======================================================
	__syncthreads();                // head / sol / sys are reused
	FT (&dT)[M] = dp[0];
	{
		FT diss[M], V[M], cT[M];
		load8<FT, DIR>(A.cur[3], off, cT);
		{
			// dissipation function of the sweep direction (TimeLayer3D.h:554-588), accumulated component by component:
			//   X: 2 u_x^2 + v_x^2 + w_x^2 + v_x u_y + w_x u_z ; Y: u_y^2 + 2 v_y^2 + w_y^2 + u_y v_x + w_y v_z ;
			//   Z: u_z^2 + v_z^2 + 2 w_z^2 + u_z w_x + v_z w_y
			// c1, c2: the two cross-line derivatives of temp[DIR]; c1 pairs with component QA, c2 with QB.  Lines next
			// to an interior cell always exist, for everything else the (clamped, valid) addresses just deliver values
			// that are never selected
			constexpr int QA = DIR == 0 ? 1 : 0, QB = DIR == 2 ? 1 : 2;
			// (j +- 1 are different distances at the edges of a y-block)
			const long long s1p = DIR == 0 ? L.jup(jline) : L.plane, s1m = DIR == 0 ? L.jdn(jline) : L.plane;
			const long long s2p = DIR == 2 ? L.jup(jline) : 1, s2m = DIR == 2 ? L.jdn(jline) : 1;
			const FT *tp = A.temp[DIR];
			FT c1[M], c2[M];
			{
				FT p1[M], m1[M];
				load8<FT, DIR>(tp + s1p, off, p1); load8<FT, DIR>(tp - s1m, off, m1);
#pragma unroll
				for (int i = 0; i < M; i++) c1[i] = (p1[i] - m1[i]) * K.inv2h1;
			}
			{
				FT p2[M], m2[M];
				if (DIR == 2) { load8<FT, DIR>(tp + s2p, off, p2); load8<FT, DIR>(tp - s2m, off, m2); }
				else {          // +-1 element along k: unaligned, scalar
#pragma unroll
					for (int i = 0; i < M; i++) { p2[i] = tp[off[i] + 1]; m2[i] = tp[off[i] - 1]; }
				}
#pragma unroll
				for (int i = 0; i < M; i++) c2[i] = (p2[i] - m2[i]) * K.inv2h2;
			}
#pragma unroll
			for (int i = 0; i < M; i++) diss[i] = FT(0);
#pragma unroll
			for (int q = 0; q < 3; q++) {
				FT f[M];
				load8<FT, DIR>(A.temp[q], off, f);
				const FT lo = A.temp[q][off_lo], hi = A.temp[q][off_hi];
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
			if (!any_int) {
#pragma unroll
				for (int i = 0; i < M; i++) diss[i] = FT(0);
			}
		}
#pragma unroll
		for (int i = 0; i < M; i++) {
			const unsigned r = ROLE(i);
			const FT Vh = V[i] * K.inv2h;
			FT a = -Vh - K.vis_T, c = Vh - K.vis_T, b = K.b_T;
			FT d = cT[i] * K.c3dt + K.t_phi * diss[i];
			if (!plain && (r & (R_SEG | R_PRE)) != R_INT) {
				const bool tfree = r & R_TFREE;
				if (r & R_INT) {
					if (tfree) b += FT(0.5) * c;
					else d -= c * A.nodev[3][off[i] + (i == M - 1 ? step_last : (int)stride)];
					c = FT(0);
				} else if (r & (R_START | R_END)) {
					a = ((r & (R_END | R_START)) == R_END && tfree) ? FT(-1) : FT(0);
					c = ((r & R_START) && tfree) ? FT(-1) : FT(0);
					b = tfree ? FT(2) : FT(1);
					d = tfree ? FT(0) : A.nodev[3][off[i]];
				} else { a = FT(0); c = FT(0); b = FT(1); d = FT(0); }
			}
			if (MODE == 2) {
				if (i == 0 && g == 0) { d -= a * A.xbnd[xo + 3 * A.lpo]; a = FT(0); }
				if (i == M - 1 && g == GL - 1) { d -= c * A.xbnd[xo + 7 * A.lpo]; c = FT(0); }
			}
			CMC_ELIM_ROW(i, a, b, c)
			if (i == M - 1) dT[i] = d;
			else if (i == 0) dT[0] = d * rr;
			else dT[i] = (d - a * dT[i - 1]) * rr;
		}
	}
	CMC_MARK(4);   // phase T loads + elimination done
	{
		FT y0 = dT[M - 2], v0 = lp[M - 2], w0 = cp[M - 2];
#pragma unroll
		for (int i = M - 3; i >= 0; i--) { y0 = dT[i] - cp[i] * y0; v0 = lp[i] - cp[i] * v0; w0 = -cp[i] * w0; }
		head[0 * STR + e] = y0; head[3 * STR + e] = v0; head[4 * STR + e] = w0;
		__syncthreads();
		const bool has_next = g + 1 < GP;
		const FT *hn = head + e + GS;
		const FT ny0 = has_next ? hn[0 * STR] : FT(0), nv = has_next ? hn[3 * STR] : FT(0), nw = has_next ? hn[4 * STR] : FT(0);
		const FT a7 = lp[M - 1], c7 = cp[M - 1];
		rr = rcp<FT>(b7 - a7 * cp[M - 2] - c7 * nv);
		FT Rd[1] = {(dT[M - 1] - a7 * dT[M - 2] - c7 * ny0) * rr}, ET[1];
		if (MODE == 1) {
			FT Re[3] = {Rd[0], FT(0), FT(0)}, Xe[3];
			FT Ra = -a7 * lp[M - 2] * rr;
			if (g == 0) { Re[1] = Ra; Ra = FT(0); }
			if (g == GL - 1) Re[2] = c7 * rr;
			reduced_solve<FT, 3, GP, GS, NL>(sys, sol, g, e, Ra, -c7 * nw * rr, Re, Xe);
			if (line_ok && g == 0) {
				cto[10 * A.lpo] = y0 - w0 * Xe[0];
				cto[12 * A.lpo] = v0 - w0 * Xe[1];
				cto[13 * A.lpo] = -w0 * Xe[2];
			}
			if (line_ok && g == GL - 1) {
				cto[11 * A.lpo] = Xe[0];
				cto[14 * A.lpo] = Xe[1];
				cto[15 * A.lpo] = Xe[2];
			}
			return;
		}
		reduced_solve<FT, 1, GP, GS, NL>(sys, sol, g, e, -a7 * lp[M - 2] * rr, -c7 * nw * rr, Rd, ET);
		CMC_MARK(5);   // phase T reduced solve done
		const FT El = g > 0 ? sol[e - GS] : FT(0);
		FT x[M], tq[M];
		load8<FT, DIR>(A.temp[3], off, tq);
		x[M - 1] = ET[0];
#pragma unroll
		for (int i = M - 2; i >= 0; i--) x[i] = dT[i] - lp[i] * El - cp[i] * x[i + 1];
		if (holes) {
#pragma unroll
			for (int i = 0; i < M; i++)
				if (holes & (1u << i)) x[i] = A.next[3][off[i]];
		}
		relax8<FT, DIR>(tq, x, inmask, A.extra_merge);
		store8<FT, DIR>(A.temp_out[3], off, full, tq);
		store8<FT, DIR>(A.next[3], off, segfull, x);
		push_planes<FT, DIR, MODE>(A, 3, pi, g, GL, off, full, segfull, tq, x);
	}
	CMC_MARK(6);
#undef CMC_MARK
#undef ROLE
#undef CMC_SCHED_FENCE
}

template <typename FT, int GP, int NL, int NRHS>
static size_t fast_smem_bytes() { return sizeof(FT) * (size_t)(reduced_scratch_elems<NRHS, GP, NL>() + 5 * GP * NL); }

// x / y lines: short lines take more of them per CTA, which widens the row segments a warp touches (8 lines = 64 bytes
// in fp64, 32 lines = 256 bytes) at the same CTA size; z lines are contiguous anyway.
constexpr int lines_per_cta(int GP, int DIR = 2) { return DIR == 2 ? 8 : GP <= 16 ? 32 : GP == 32 ? 16 : 8; }

template <typename FT, int DIR, int GP, int MODE = 0, int NL = lines_per_cta(GP, DIR)>
static unsigned launch_one(const SweepArgs<FT> &A, cudaStream_t s, long long *trace, bool dry)
{
	const Layout &L = A.L;
	unsigned grid;
	if (DIR == 0) grid = (unsigned)L.ny * (unsigned)((L.nz + NL - 1) / NL);
	else if (DIR == 1) grid = (unsigned)L.nx * (unsigned)((L.nz + NL - 1) / NL);
	else grid = (unsigned)L.nx * (unsigned)((L.ny + NL - 1) / NL);
	if (dry) return grid;
	const size_t smem = fast_smem_bytes<FT, GP, NL, (MODE == 1 ? 5 : 3)>();
	{   // the attribute is per device: set it once on every device this process launches on
		static bool attr_set[64] = {};
		int dev = 0;
		cudaGetDevice(&dev);
		if (dev < 0 || dev >= 64 || !attr_set[dev]) {
			cudaFuncSetAttribute((const void *)k_fast_sweep<FT, DIR, GP, NL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			if (dev >= 0 && dev < 64) attr_set[dev] = true;
		}
	}
	FastConst<FT> K; K.init(A, DIR);
	k_fast_sweep<FT, DIR, GP, NL, MODE><<<grid, GP * NL, smem, s>>>(A, K, trace, 1);
	return grid;
}

template <typename FT, int DIR>
static unsigned launch_dir(int GP, const SweepArgs<FT> &A, cudaStream_t s, long long *trace, bool dry)
{
	switch (GP) {
	case 4: return launch_one<FT, DIR, 4>(A, s, trace, dry);
	case 8: return launch_one<FT, DIR, 8>(A, s, trace, dry);
	case 16: return launch_one<FT, DIR, 16>(A, s, trace, dry);
	case 32: return launch_one<FT, DIR, 32>(A, s, trace, dry);
	case 128: return launch_one<FT, 2, 128, 0, 2>(A, s, trace, dry);       // (z only: fast_sweep_supported)
	default: {
		// z lines of 512 rows: fewer lines per CTA, more independent CTAs per SM (CMC_NLZ = 1, 2, 4 or 8 lines per CTA)
		static const int nlz = getenv("CMC_NLZ") ? atoi(getenv("CMC_NLZ")) : 1;     // measured, 512^3 fp64: 3.85 / 3.92 / 4.17 / 4.65 ms for 1 / 2 / 4 / 8 lines
		if (DIR == 2 && nlz == 4) return launch_one<FT, DIR, 64, 0, 4>(A, s, trace, dry);
		if (DIR == 2 && nlz == 2) return launch_one<FT, DIR, 64, 0, 2>(A, s, trace, dry);
		if (DIR == 2 && nlz == 1) return launch_one<FT, DIR, 64, 0, 1>(A, s, trace, dry);
		return launch_one<FT, DIR, 64>(A, s, trace, dry);
	}
	}
}

bool fast_sweep_supported(const Layout &L, int dir)
{
	const int n = dir == 0 ? L.nx : dir == 1 ? L.ny : L.nz;
	if (n < M) return false;                              // clamped row offsets need at least one full chunk
	if (L.total >= (1ll << 31)) return false;             // 32-bit element offsets
	// lines of up to 512 rows; along z (contiguous lines, 2 lines x 128 chunks per CTA) up to 1024.  Longer x / y lines: the
	// CTA-pair form of the TMA kernel (kernels_tma.cu) or the caller's fallback
	return (n + M - 1) / M <= (dir == 2 ? 128 : 64);
}

template <typename FT>
bool launch_fast_sweep(int dir, const SweepArgs<FT> &A, cudaStream_t s, long long *launches)
{
	const Layout &L = A.L;
	if (!fast_sweep_supported(L, dir)) return false;
	const int n = dir == 0 ? L.nx : dir == 1 ? L.ny : L.nz;
	const int G = (n + M - 1) / M;
	int GP = 4;                                           // at least one warp per CTA
	while (GP < G) GP <<= 1;
	const unsigned grid = dir == 0 ? launch_dir<FT, 0>(GP, A, s, nullptr, true) : dir == 1 ? launch_dir<FT, 1>(GP, A, s, nullptr, true)
	                                                                                       : launch_dir<FT, 2>(GP, A, s, nullptr, true);
	// debug facility: CMC_TRACE=1 records clock64() at the phase boundaries of every CTA and prints the mean phase
	// durations (cycles) per direction.  Never enabled in tests or benchmarks.
	static const bool tracing = getenv("CMC_TRACE") != nullptr;
	long long *trace = nullptr;
	if (tracing) { cudaMalloc((void **)&trace, (size_t)grid * 16 * sizeof(long long)); cudaMemset(trace, 0, (size_t)grid * 16 * sizeof(long long)); }
	if (dir == 0) launch_dir<FT, 0>(GP, A, s, trace, false);
	else if (dir == 1) launch_dir<FT, 1>(GP, A, s, trace, false);
	else launch_dir<FT, 2>(GP, A, s, trace, false);
	if (tracing) {
		cudaStreamSynchronize(s);
		std::vector<long long> h((size_t)grid * 16);
		cudaMemcpy(h.data(), trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
		cudaFree(trace);
		double acc[7] = {};
		for (unsigned b = 0; b < grid; b++) for (int k = 1; k < 7; k++) acc[k] += (double)(h[(size_t)b * 16 + k] - h[(size_t)b * 16 + k - 1]);
		fprintf(stderr, "[cmc trace] dir %d grid %u threads %d: mean cycles per CTA  loadV+elim %.0f | pcrV %.0f | storeV %.0f | loadT+elim %.0f | pcrT %.0f | storeT %.0f | total %.0f\n",
		        dir, grid, GP * lines_per_cta(GP, dir), acc[1] / grid, acc[2] / grid, acc[3] / grid, acc[4] / grid, acc[5] / grid, acc[6] / grid,
		        (acc[1] + acc[2] + acc[3] + acc[4] + acc[5] + acc[6]) / grid);
	}
	if (launches) (*launches)++;
	return true;
}

// x-sweep passes of a slab (MODE 1 / MODE 2); nx must be a multiple of 8 and at most 512
template <typename FT, int MODE>
static bool launch_x_mode(const SweepArgs<FT> &A, cudaStream_t s)
{
	const Layout &L = A.L;
	const bool last_slab = L.x0 + L.nx == L.gx;
	if (!fast_sweep_supported(L, 0) || (L.nx % M != 0 && !last_slab)) return false;
	int GP = 4;
	while (GP < (L.nx + M - 1) / M) GP <<= 1;
	switch (GP) {
	case 4: launch_one<FT, 0, 4, MODE>(A, s, nullptr, false); break;
	case 8: launch_one<FT, 0, 8, MODE>(A, s, nullptr, false); break;
	case 16: launch_one<FT, 0, 16, MODE>(A, s, nullptr, false); break;
	case 32: launch_one<FT, 0, 32, MODE>(A, s, nullptr, false); break;
	default: launch_one<FT, 0, 64, MODE>(A, s, nullptr, false); break;
	}
	return true;
}
template <typename FT>
bool launch_x_spike(const SweepArgs<FT> &A, cudaStream_t s, long long *launches)
{
	if (!launch_x_mode<FT, 1>(A, s)) return false;
	if (launches) (*launches)++;
	return true;
}
template <typename FT>
bool launch_x_coupled(const SweepArgs<FT> &A, cudaStream_t s, long long *launches)
{
	if (!launch_x_mode<FT, 2>(A, s)) return false;
	if (launches) (*launches)++;
	return true;
}
template bool launch_x_spike<float>(const SweepArgs<float> &, cudaStream_t, long long *);
template bool launch_x_spike<double>(const SweepArgs<double> &, cudaStream_t, long long *);
template bool launch_x_coupled<float>(const SweepArgs<float> &, cudaStream_t, long long *);
template bool launch_x_coupled<double>(const SweepArgs<double> &, cudaStream_t, long long *);

// ---- interface system of the partitioned x-sweep -------------------------------------------------------------------
// One thread per OWNED line.  coef[(src * 16 + v) * lpo + line]: the 16 coefficients rank `src` emitted for the line
// (k_fast_sweep MODE 1).  Unknowns per rank r: F_r (first row of its slab), L_r (last row):
//     F_r + pf_r L_{r-1} + qf_r F_{r+1} = f_r ,   L_r + pl_r L_{r-1} + ql_r F_{r+1} = l_r .
// Block-tridiagonal in Z_r = (L_r, F_{r+1}), r = 0..P-2, solved by block Thomas (2x2 blocks).  Writes, for every rank
// r, the values of its neighbours' adjacent rows: bnd[(r * 8 + q) * lpo + line] = L_{r-1} (q < 4), F_{r+1} (q >= 4).
constexpr int MAXP = MAX_SLABS;
template <typename FT> struct BndTargets { FT *p[MAXP]; };    // per rank: the [8][lpo] block this owner fills in its table

template <typename FT>
__global__ void k_x_interface(int P, int lpo, int nlines, const FT *__restrict__ coef, const BndTargets<FT> bnd)
{
	const int line = blockIdx.x * blockDim.x + threadIdx.x;
	if (line >= nlines) return;
#pragma unroll 1
	for (int mat = 0; mat < 2; mat++) {
		const int nr = mat == 0 ? 3 : 1;                 // right-hand sides: u, v, w | T
		const int vf = mat == 0 ? 0 : 10, vl = mat == 0 ? 3 : 11, vp = mat == 0 ? 6 : 12;   // f, l, (pf, qf, pl, ql)
		FT m01[MAXP], i00[MAXP], i01[MAXP], i10[MAXP], i11[MAXP], r0[3][MAXP], r1[3][MAXP];
		auto C = [&](int src, int v) { return coef[((size_t)src * 16 + v) * lpo + line]; };
		FT p_i00 = 0, p_i01 = 0;                         // inverse of the previous (eliminated) block
		FT p_s[3] = {0, 0, 0};                           // i00 rhs0 + i01 rhs1 of the previous block
		for (int r = 0; r + 1 < P; r++) {
			const FT pl = C(r, vp + 2), ql = C(r, vp + 3), pf1 = C(r + 1, vp + 0), qf_r = C(r, vp + 1);
			m01[r] = ql - pl * p_i01 * qf_r;             // (r == 0: pl == 0)
			const FT det = FT(1) - m01[r] * pf1, id = FT(1) / det;
			i00[r] = id; i01[r] = -m01[r] * id; i10[r] = -pf1 * id; i11[r] = id;
			for (int q = 0; q < nr; q++) {
				r0[q][r] = C(r, vl + q) - pl * p_s[q];
				r1[q][r] = C(r + 1, vf + q);
				p_s[q] = i00[r] * r0[q][r] + i01[r] * r1[q][r];
			}
			p_i00 = i00[r]; p_i01 = i01[r];
		}
		(void)p_i00;
		for (int q = 0; q < nr; q++) {
			FT Lr[MAXP], Fr1[MAXP];                      // L_r, F_{r+1}
			FT nextF = FT(0);                            // F_{r+2} of the block above
			for (int r = P - 2; r >= 0; r--) {
				const FT qf1 = C(r + 1, vp + 1);
				const FT b0 = r0[q][r], b1 = r1[q][r] - qf1 * nextF;
				Lr[r] = i00[r] * b0 + i01[r] * b1;
				Fr1[r] = i10[r] * b0 + i11[r] * b1;
				nextF = Fr1[r];
			}
			const int qq = mat == 0 ? q : 3;
			for (int r = 0; r < P; r++) {
				bnd.p[r][(size_t)qq * lpo + line] = r > 0 ? Lr[r - 1] : FT(0);
				bnd.p[r][(size_t)(4 + qq) * lpo + line] = r + 1 < P ? Fr1[r] : FT(0);
			}
		}
	}
}

template <typename FT>
void launch_x_interface(int P, int lpo, int nlines, const FT *coef, FT *const *bnd_to, cudaStream_t s, long long *launches)
{
	if (nlines <= 0) return;
	BndTargets<FT> bt;
	for (int r = 0; r < MAXP; r++) bt.p[r] = r < P ? bnd_to[r] : nullptr;
	k_x_interface<FT><<<(nlines + 127) / 128, 128, 0, s>>>(P, lpo, nlines, coef, bt);
	if (launches) (*launches)++;
}
template void launch_x_interface<float>(int, int, int, const float *, float *const *, cudaStream_t, long long *);
template void launch_x_interface<double>(int, int, int, const double *, double *const *, cudaStream_t, long long *);

template bool launch_fast_sweep<float>(int, const SweepArgs<float> &, cudaStream_t, long long *);
template bool launch_fast_sweep<double>(int, const SweepArgs<double> &, cudaStream_t, long long *);

// ---- standalone batched solver with the same partition + CR/PCR machinery (unit tests) --------------------------
// One CTA per NLB systems; system-major host layout.
template <typename FT, int GP>
__global__ void __launch_bounds__(GP * NLB) k_pcr_batch(int nsys, int n, const FT *a, const FT *b, const FT *c, const FT *d, FT *x)
{
	constexpr int STR = GP * NLB;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	FT *sys = reinterpret_cast<FT *>(smem_raw);
	FT *head = sys + 3 * 5 * STR;
	FT *sol = head + 5 * STR;
	const int t = threadIdx.x;
	const int l = t % NLB, g = t / NLB, e = t;
	const int sidx = blockIdx.x * NLB + l;
	const bool ok = sidx < nsys;
	const size_t base = (size_t)(ok ? sidx : 0) * n;
	const int r0 = g * M;
	FT cp[M], lp[M], dp[M], b7 = FT(1), rr;
#pragma unroll
	for (int i = 0; i < M; i++) {
		const int r = r0 + i;
		FT ra = FT(0), rb = FT(1), rc = FT(0), rd = FT(0);
		if (ok && r < n) { ra = r == 0 ? FT(0) : a[base + r]; rb = b[base + r]; rc = r == n - 1 ? FT(0) : c[base + r]; rd = d[base + r]; }
		CMC_ELIM_ROW(i, ra, rb, rc)
		if (i == M - 1) dp[i] = rd;
		else if (i == 0) dp[0] = rd * rr;
		else dp[i] = (rd - ra * dp[i - 1]) * rr;
	}
	FT y0 = dp[M - 2], v0 = lp[M - 2], w0 = cp[M - 2];
#pragma unroll
	for (int i = M - 3; i >= 0; i--) { y0 = dp[i] - cp[i] * y0; v0 = lp[i] - cp[i] * v0; w0 = -cp[i] * w0; }
	head[0 * STR + e] = y0; head[3 * STR + e] = v0; head[4 * STR + e] = w0;
	__syncthreads();
	const bool has_next = g + 1 < GP;
	const FT *hn = head + e + NLB;
	const FT ny0 = has_next ? hn[0 * STR] : FT(0), nv = has_next ? hn[3 * STR] : FT(0), nw = has_next ? hn[4 * STR] : FT(0);
	const FT a7 = lp[M - 1], c7 = cp[M - 1];
	rr = rcp<FT>(b7 - a7 * cp[M - 2] - c7 * nv);
	FT Rd[1] = {(dp[M - 1] - a7 * dp[M - 2] - c7 * ny0) * rr}, E[1];
	reduced_solve<FT, 1, GP, NLB, NLB>(sys, sol, g, e, -a7 * lp[M - 2] * rr, -c7 * nw * rr, Rd, E);
	const FT El = g > 0 ? sol[e - NLB] : FT(0);
	FT xx[M];
	xx[M - 1] = E[0];
#pragma unroll
	for (int i = M - 2; i >= 0; i--) xx[i] = dp[i] - lp[i] * El - cp[i] * xx[i + 1];
#pragma unroll
	for (int i = 0; i < M; i++) if (ok && r0 + i < n) x[base + r0 + i] = xx[i];
}

template <typename FT, int GP>
static void launch_batch_one(int nsys, int n, const FT *a, const FT *b, const FT *c, const FT *d, FT *x, cudaStream_t s)
{
	const size_t smem = sizeof(FT) * (size_t)(3 * 5 + 5 + 3) * GP * NLB;
	cudaFuncSetAttribute((const void *)k_pcr_batch<FT, GP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	k_pcr_batch<FT, GP><<<(nsys + NLB - 1) / NLB, GP * NLB, smem, s>>>(nsys, n, a, b, c, d, x);
}

template <typename FT>
bool launch_pcr_batch(int nsys, int n, const FT *a, const FT *b, const FT *c, const FT *d, FT *x, cudaStream_t s)
{
	const int G = (n + M - 1) / M;
	int GP = 4;
	while (GP < G) GP <<= 1;
	switch (GP) {
	case 4: launch_batch_one<FT, 4>(nsys, n, a, b, c, d, x, s); break;
	case 8: launch_batch_one<FT, 8>(nsys, n, a, b, c, d, x, s); break;
	case 16: launch_batch_one<FT, 16>(nsys, n, a, b, c, d, x, s); break;
	case 32: launch_batch_one<FT, 32>(nsys, n, a, b, c, d, x, s); break;
	case 64: launch_batch_one<FT, 64>(nsys, n, a, b, c, d, x, s); break;
	default: return false;
	}
	return true;
}
template bool launch_pcr_batch<float>(int, int, const float *, const float *, const float *, const float *, float *, cudaStream_t);
template bool launch_pcr_batch<double>(int, int, const double *, const double *, const double *, const double *, double *, cudaStream_t);

} // namespace cmc
