// kernels_ring.cu - CMC_MODE_FAST directional sweeps for sm_100a, persistent CTAs fed by an asynchronous
// shared-memory ring.
//
// Same arithmetic as k_fast_sweep (kernels_fast.cu: partition method with 8-row chunks in registers, CR + PCR
// reduced solve in shared memory, fused coefficient build / boundary rows / mask / relaxation), different data
// movement.  The direct-load kernel is latency bound: one 512-thread CTA per SM owns the register file, its six
// load phases are separated by barriers, and every SM runs the same phase at the same time, so HBM idles while
// the SMs solve (ncu, profiles/r01_*: long-scoreboard stalls 45 %, DRAM 40 %).  Here
//   * one persistent CTA per SM walks over its tiles (tile = NL = 8 neighbouring lines, all rows);
//   * every input field of a tile is brought into one of five shared-memory slots with cp.async (LDGSTS, 16-byte
//     pieces, L1 bypassed), one whole phase ahead of its use and across tile boundaries:
//         group G1 (for the u,v,w phase)      : temp[DIR], temp.T, cur.u, cur.v, cur.w   + line descriptors
//         group G2 (relaxation + the T phase) : temp.u, temp.v, temp.w, cur.T, temp.T    + cross-line halo of temp[DIR]
//     G2 of a tile is issued as soon as G1 has been read into registers and lands during the u,v,w solve; G1 of
//     the NEXT tile is issued as soon as G2 has been consumed and lands during the T solve and the stores;
//   * registers only ever hold the chunk being eliminated, so no load waits for a register and no register waits
//     for a load: HBM streams while the SM computes.
// Slots use rotated (bank-conflict free) layouts, see slot_*; results are stored straight from registers.
// HBM traffic per cell and sweep stays the algorithmic 16 values + 1 descriptor byte (re-used fields of G2 hit L2).
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include "kernels.h"
#include "fast_core.cuh"
#include "ring_slots.cuh"

namespace cmc {

template <typename FT, int DIR, int GP, int NL>
__global__ void __launch_bounds__(GP * NL, 1) k_ring_sweep(const SweepArgs<FT> A, const FastConst<FT> K, const int ntiles)
{
	typedef Slot<FT, DIR, GP, NL> S;
	constexpr int STR = GP * NL;
	constexpr int GS = DIR == 2 ? 1 : NL;          // shared-memory distance of neighbouring chunks of a line
	constexpr int NR = 5;                          // matrix (2) + three right-hand sides
	extern __shared__ __align__(128) unsigned char smem_raw[];
	FT *slots = reinterpret_cast<FT *>(smem_raw);                 // 5 slots
	FT *halo = slots + 5 * S::ELEMS;                              // cross-line neighbours of temp[DIR]: 2 * 8 * GP elements
	FT *sys = halo + 2 * M * GP;                                  // reduced-solve scratch
	FT *sol = sys;                                                // aliases the CR publications (see reduced_solve)
	FT *head = sys + reduced_scratch_elems<3, GP, NL>();          // 5 arrays: y0[3], v0, w0 of every chunk
	uint8_t *roles = reinterpret_cast<uint8_t *>(head + 5 * STR); // descriptor bytes of the tile: 8 * STR
#define SLOT(k) (slots + (k) * S::ELEMS)

	const Layout &L = A.L;
	const int t = threadIdx.x;
	int g, l;                                      // chunk, line-in-tile
	if (DIR == 2) { g = t % GP; l = t / GP; } else { l = t % NL; g = t / NL; }
	const int e = t;                               // == l * GP + g (Z) or g * NL + l (X, Y)
	const int r0 = g * M;                          // first row of this chunk

	// ---- copy groups ------------------------------------------------------------------------------------------
	auto issue_g1 = [&](const Tile<DIR, NL> &T) {
		S::issue(SLOT(0), A.temp[DIR], T, L, t);
		S::issue(SLOT(1), A.temp[3], T, L, t);
		S::issue(SLOT(2), A.cur[0], T, L, t);
		S::issue(SLOT(3), A.cur[1], T, L, t);
		S::issue(SLOT(4), A.cur[2], T, L, t);
		// descriptors: 8 bytes per thread.  X, Y: thread <-> tile row (NL = 8 bytes); Z: this thread's own chunk
		if (DIR == 2) cp_async8(roles + (size_t)e * 8, A.role + T.tbase + (long long)min(l, T.lines - 1) * L.nzp + min(r0, L.nzp - M));
		else cp_async8(roles + (size_t)t * 8, A.role + T.tbase + (long long)min(t, T.n - 1) * T.stride);
		cp_async_commit();
	};
	auto issue_g2 = [&](const Tile<DIR, NL> &T) {
		S::issue(SLOT(0), A.temp[0], T, L, t);
		S::issue(SLOT(1), A.temp[1], T, L, t);
		S::issue(SLOT(2), A.temp[2], T, L, t);
		S::issue(SLOT(3), A.cur[3], T, L, t);
		S::issue(SLOT(4), A.temp[3], T, L, t);
		// halo of temp[DIR] across the second cross direction (the first one is loaded directly, see below)
		const FT *tp = A.temp[DIR] + T.tbase;
		if (DIR == 2) {
			// lines j0 - 1 and j0 + NL: halo[side][row], 16-byte pieces
			constexpr int PIECES = 2 * GP * S::PC;
			for (int p = t; p < PIECES; p += STR) {
				const int side = p / (GP * S::PC), w = p - side * (GP * S::PC);
				// lines j0 - 1 / j0 + lines: the neighbouring rows may live in the neighbouring y-block
				const long long lo = side ? (long long)(T.lines - 1) * L.nzp + L.jup(T.j0 + T.lines - 1) : -L.jdn(T.j0);
				const FT *src = tp + lo + min(w * S::EPP, L.nzp - S::EPP);
				cp_async16(halo + side * (M * GP) + w * S::EPP, src);
			}
		} else {
			// columns k0 - 1 and k0 + NL of every tile row: halo[row][side]
			const long long ro = (long long)min(t, T.n - 1) * T.stride;
			cp_async_elem<FT>(halo + 2 * t, tp + ro - 1);
			cp_async_elem<FT>(halo + 2 * t + 1, tp + ro + NL);
		}
		cp_async_commit();
	};

	Tile<DIR, NL> T;
	int tile = blockIdx.x;
	if (tile < ntiles) { T.set(L, tile); issue_g1(T); }

	for (; tile < ntiles; tile += gridDim.x) {
		T.set(L, tile);
		const int n = T.n;
		const bool line_ok = l < T.lines;
		const long long stride = T.stride;
		const long long base = T.tbase + (DIR == 2 ? (long long)min(l, T.lines - 1) * L.nzp : (long long)(line_ok ? l : 0));
		// global offsets of the chunk's rows (stores, rare direct loads), clamped into the line
		int off[M];
		if (DIR == 2) {
			const int rc = min(r0, L.nzp - M);
#pragma unroll
			for (int i = 0; i < M; i++) off[i] = (int)base + rc + i;
		} else {
#pragma unroll
			for (int i = 0; i < M; i++) off[i] = (int)base + min(r0 + i, n - 1) * (int)stride;
		}
		// neighbouring rows of the chunk along the line (clamped like the direct-load kernel)
		const int r_lo = max(r0 - 1, 0), r_hi = min(r0 + M, n - 1);
		unsigned rowmask = 0;
#pragma unroll
		for (int i = 0; i < M; i++) rowmask |= (line_ok && r0 + i < n) ? (1u << i) : 0u;

		cp_async_wait_all();
		__syncthreads();            // G1 of this tile has landed

		// roles of the chunk's rows
		unsigned rw0 = 0, rw1 = 0;
		if (DIR == 2) {
			const uint2 w = *reinterpret_cast<const uint2 *>(roles + (size_t)e * 8);
			rw0 = w.x; rw1 = w.y;
		} else {
#pragma unroll
			for (int i = 0; i < 4; i++) {
				rw0 |= (unsigned)roles[(r0 + i) * 8 + l] << (8 * i);
				rw1 |= (unsigned)roles[(r0 + 4 + i) * 8 + l] << (8 * i);
			}
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
		const unsigned full = (DIR == 2 && line_ok && r0 < n) ? 0xffu : rowmask;
		const unsigned segfull = (segmask | (full & ~rowmask)) == 0xffu ? 0xffu : segmask;

		// ======================================= phase V: u, v, w ==========================================
		FT cp[M], lp[M], dp[3][M];
		FT b7 = FT(1), rr;
		{
			FT V[M], Tl[M];
			S::read_chunk(SLOT(0), l, g, V);
			S::read_chunk(SLOT(1), l, g, Tl);
			S::read_chunk(SLOT(2), l, g, dp[0]);
			S::read_chunk(SLOT(3), l, g, dp[1]);
			S::read_chunk(SLOT(4), l, g, dp[2]);
			const FT Tlo = SLOT(1)[S::at(l, r_lo)], Thi = SLOT(1)[S::at(l, r_hi)];
			__syncthreads();        // every thread has its G1 data: the slots are free
			issue_g2(T);
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
					if (r & R_INT) {                        // R_PRE: the next cell ends this segment AND starts the next one
						if (vfree) b += FT(0.5) * c;
						else {
							const int idn = off[i] + (int)stride;
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
		FT E[3];
		{
			FT y0[3] = {dp[0][M - 2], dp[1][M - 2], dp[2][M - 2]}, v0 = lp[M - 2], w0 = cp[M - 2];
#pragma unroll
			for (int i = M - 3; i >= 0; i--) {
				y0[0] = dp[0][i] - cp[i] * y0[0]; y0[1] = dp[1][i] - cp[i] * y0[1]; y0[2] = dp[2][i] - cp[i] * y0[2];
				v0 = lp[i] - cp[i] * v0; w0 = -cp[i] * w0;
			}
			head[0 * STR + e] = y0[0]; head[1 * STR + e] = y0[1]; head[2 * STR + e] = y0[2];
			head[3 * STR + e] = v0; head[4 * STR + e] = w0;
			__syncthreads();
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
			reduced_solve<FT, 3, GP, GS, NL>(sys, sol, g, e, -a7 * lp[M - 2] * rr, -c7 * nw * rr, Rd, E);
		}
		cp_async_wait_all();
		__syncthreads();            // G2 has landed (and every separator solution is in sol[])

		// back substitution in place (dp[q] <- x), which retires cp / lp before anything else needs registers
#pragma unroll
		for (int q = 0; q < 3; q++) {
			const FT El = g > 0 ? sol[q * STR + e - GS] : FT(0);
			dp[q][M - 1] = E[q];
#pragma unroll
			for (int i = M - 2; i >= 0; i--) dp[q][i] = dp[q][i] - lp[i] * El - cp[i] * dp[q][i + 1];
		}
		// store u, v, w and the relaxed linearisation layer
#pragma unroll
		for (int q = 0; q < 3; q++) {
			FT (&x)[M] = dp[q];
			FT tq[M];
			S::read_chunk(SLOT(q), l, g, tq);
			if (holes) {
#pragma unroll
				for (int i = 0; i < M; i++)
					if (holes & (1u << i)) x[i] = A.next[q][off[i]];
			}
			relax8<FT, DIR>(tq, x, inmask, A.extra_merge);
			store8<FT, DIR>(A.temp_out[q], off, full, tq);
			store8<FT, DIR>(A.next[q], off, segfull, x);
			push_planes<FT, DIR, 0>(A, q, T.pi, g, GP, off, full, segfull, tq, x);
		}

		// ======================================= phase T ==================================================
		FT dT[M];
		FT tT[M];                   // temp.T of the chunk, for the relaxation at the end
		{
			FT diss[M], V[M], cT[M];
			{
				// dissipation function of the sweep direction (TimeLayer3D.h:554-588), accumulated component by component:
				//   X: 2 u_x^2 + v_x^2 + w_x^2 + v_x u_y + w_x u_z ; Y: u_y^2 + 2 v_y^2 + w_y^2 + u_y v_x + w_y v_z ;
				//   Z: u_z^2 + v_z^2 + 2 w_z^2 + u_z w_x + v_z w_y
				// c1, c2: the two cross-line derivatives of temp[DIR]; c1 pairs with component QA, c2 with QB
				constexpr int QA = DIR == 0 ? 1 : 0, QB = DIR == 2 ? 1 : 2;
				FT c1[M], c2[M];
				{
					// first cross direction straight from L2 / HBM (issuing these loads earlier costs more in spills than
					// it hides in latency: measured), second one from the tile itself and its halo
					FT p1[M], m1[M], p2[M], m2[M];
					const long long s1 = DIR == 0 ? L.nzp : L.plane;
					load8<FT, DIR>(A.temp[DIR] + s1, off, p1); load8<FT, DIR>(A.temp[DIR] - s1, off, m1);
					if (DIR == 2) {
						if (l == NL - 1) {
#pragma unroll
							for (int i = 0; i < M; i++) p2[i] = halo[M * GP + min(r0, L.nzp - M) + i];
						} else S::read_chunk(SLOT(DIR), l + 1, g, p2);
						if (l == 0) {
#pragma unroll
							for (int i = 0; i < M; i++) m2[i] = halo[min(r0, L.nzp - M) + i];
						} else S::read_chunk(SLOT(DIR), l - 1, g, m2);
					} else {
#pragma unroll
						for (int i = 0; i < M; i++) {
							const int pr = (g << 3) + ((i + g) & 7);
							p2[i] = l == NL - 1 ? halo[2 * (r0 + i) + 1] : SLOT(DIR)[pr * NL + l + 1];
							m2[i] = l == 0 ? halo[2 * (r0 + i)] : SLOT(DIR)[pr * NL + l - 1];
						}
					}
#pragma unroll
					for (int i = 0; i < M; i++) { c1[i] = (p1[i] - m1[i]) * K.inv2h1; c2[i] = (p2[i] - m2[i]) * K.inv2h2; }
				}
#pragma unroll
				for (int i = 0; i < M; i++) diss[i] = FT(0);
#pragma unroll
				for (int q = 0; q < 3; q++) {
					FT f[M];
					S::read_chunk(SLOT(q), l, g, f);
					const FT lo = SLOT(q)[S::at(l, r_lo)], hi = SLOT(q)[S::at(l, r_hi)];
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
			S::read_chunk(SLOT(3), l, g, cT);
			S::read_chunk(SLOT(4), l, g, tT);
			__syncthreads();        // G2 consumed: the slots are free for the next tile
			if (tile + (int)gridDim.x < ntiles) {
				Tile<DIR, NL> Tn;
				Tn.set(L, tile + gridDim.x);
				issue_g1(Tn);
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
						else d -= c * A.nodev[3][off[i] + (int)stride];
						c = FT(0);
					} else if (r & (R_START | R_END)) {
						a = ((r & (R_END | R_START)) == R_END && tfree) ? FT(-1) : FT(0);
						c = ((r & R_START) && tfree) ? FT(-1) : FT(0);
						b = tfree ? FT(2) : FT(1);
						d = tfree ? FT(0) : A.nodev[3][off[i]];
					} else { a = FT(0); c = FT(0); b = FT(1); d = FT(0); }
				}
				CMC_ELIM_ROW(i, a, b, c)
				if (i == M - 1) dT[i] = d;
				else if (i == 0) dT[0] = d * rr;
				else dT[i] = (d - a * dT[i - 1]) * rr;
			}
		}
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
			reduced_solve<FT, 1, GP, GS, NL>(sys, sol, g, e, -a7 * lp[M - 2] * rr, -c7 * nw * rr, Rd, ET);
			const FT El = g > 0 ? sol[e - GS] : FT(0);
			FT x[M];
			x[M - 1] = ET[0];
#pragma unroll
			for (int i = M - 2; i >= 0; i--) x[i] = dT[i] - lp[i] * El - cp[i] * x[i + 1];
			if (holes) {
#pragma unroll
				for (int i = 0; i < M; i++)
					if (holes & (1u << i)) x[i] = A.next[3][off[i]];
			}
			relax8<FT, DIR>(tT, x, inmask, A.extra_merge);
			store8<FT, DIR>(A.temp_out[3], off, full, tT);
			store8<FT, DIR>(A.next[3], off, segfull, x);
			push_planes<FT, DIR, 0>(A, 3, T.pi, g, GP, off, full, segfull, tT, x);
		}
		// (head / sol are next written two barriers into the next tile)
	}
	cp_async_wait_all();
#undef ROLE
#undef SLOT
}

template <typename FT, int GP, int NL>
static size_t ring_smem_bytes()
{
	const size_t STR = (size_t)GP * NL;
	return sizeof(FT) * (5 * STR * M + 2 * M * GP + reduced_scratch_elems<3, GP, NL>() + 5 * STR) + 8 * STR;
}

template <typename FT, int DIR, int GP>
static bool launch_ring_one(const SweepArgs<FT> &A, cudaStream_t s)
{
	constexpr int NL = 8;
	const Layout &L = A.L;
	int ntiles;
	if (DIR == 0) ntiles = L.ny * ((L.nz + NL - 1) / NL);
	else if (DIR == 1) ntiles = L.nx * ((L.nz + NL - 1) / NL);
	else ntiles = L.nx * ((L.ny + NL - 1) / NL);
	const size_t smem = ring_smem_bytes<FT, GP, NL>();
	static int ctas_of[64] = {};   // persistent grid: every CTA slot of the device (per device of this process)
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) return false;
	if (!ctas_of[dev]) {
		if (cudaFuncSetAttribute((const void *)k_ring_sweep<FT, DIR, GP, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
		int per_sm = 0, sms = 0;
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ring_sweep<FT, DIR, GP, NL>, GP * NL, smem) != cudaSuccess || per_sm < 1) return false;
		ctas_of[dev] = per_sm * sms;
	}
	const int ctas = ctas_of[dev];
	FastConst<FT> K; K.init(A, DIR);
	k_ring_sweep<FT, DIR, GP, NL><<<std::min(ctas, ntiles), GP * NL, smem, s>>>(A, K, ntiles);
	return true;
}

template <typename FT, int DIR>
static bool launch_ring_dir(int GP, const SweepArgs<FT> &A, cudaStream_t s)
{
	switch (GP) {
	case 4: return launch_ring_one<FT, DIR, 4>(A, s);
	case 8: return launch_ring_one<FT, DIR, 8>(A, s);
	case 16: return launch_ring_one<FT, DIR, 16>(A, s);
	case 32: return launch_ring_one<FT, DIR, 32>(A, s);
	default: return launch_ring_one<FT, DIR, 64>(A, s);
	}
}

bool ring_sweep_supported(const Layout &L, int dir)
{
	if (!fast_sweep_supported(L, dir)) return false;
	const int n = dir == 0 ? L.nx : dir == 1 ? L.ny : L.nz;
	if ((n + M - 1) / M > 64) return false;              // (512 rows at most; longer z lines: k_fast_sweep with 128 chunks)
	return !(L.nblk > 1 && dir != 2);                    // the x / y tiles of this kernel assume the one-block layout
}

template <typename FT>
bool launch_ring_sweep(int dir, const SweepArgs<FT> &A, cudaStream_t s, long long *launches)
{
	const Layout &L = A.L;
	if (!ring_sweep_supported(L, dir)) return false;
	const int n = dir == 0 ? L.nx : dir == 1 ? L.ny : L.nz;
	const int G = (n + M - 1) / M;
	int GP = 4;
	while (GP < G) GP <<= 1;
	const bool ok = dir == 0 ? launch_ring_dir<FT, 0>(GP, A, s) : dir == 1 ? launch_ring_dir<FT, 1>(GP, A, s) : launch_ring_dir<FT, 2>(GP, A, s);
	if (ok && launches) (*launches)++;
	return ok;
}
template bool launch_ring_sweep<float>(int, const SweepArgs<float> &, cudaStream_t, long long *);
template bool launch_ring_sweep<double>(int, const SweepArgs<double> &, cudaStream_t, long long *);

} // namespace cmc
