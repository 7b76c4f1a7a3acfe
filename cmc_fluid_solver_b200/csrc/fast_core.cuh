// fast_core.cuh - building blocks shared by the fast sweep kernels (kernels_fast.cu) and the distributed
// x-sweep kernels (kernels_dist.cu): reciprocal, sweep constants, the CR+PCR reduced solver, chunk I/O.
#pragma once
#include "kernels.h"

namespace cmc {

constexpr int M = 8;          // rows per chunk
constexpr int NLB = 8;        // systems per CTA of the standalone batch solver

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
	FT inv2h1, inv2h2;   // 1 / (2 h) of the two cross directions
	FT vis_v, vis_T, b_v, b_T;
	FT c3dt;             // 3 / dt
	FT v_T, t_phi;
	__host__ __device__ __forceinline__ void init(const SweepArgs<FT> &A, int dir)
	{
		const FT h = A.h[dir];
		inv2h = FT(1) / (2 * h);
		inv2h1 = FT(1) / (2 * A.h[dir == 0 ? 1 : 0]);
		inv2h2 = FT(1) / (2 * A.h[dir == 2 ? 1 : 2]);
		vis_v = A.v_vis / (h * h); vis_T = A.t_vis / (h * h);
		c3dt = 3 / A.dt;
		b_v = c3dt + 2 * vis_v; b_T = c3dt + 2 * vis_T;
		v_T = A.v_T; t_phi = A.t_phi;
	}
};

// ---- reduced systems of a CTA: cyclic reduction + PCR hybrid in shared memory ----------------------------------
// One thread per reduced row (chunk g of a line), rows normalised (B == 1).  GP = chunks per line rounded up to a
// power of two (rows >= G are identity rows).  e = shared-memory element of this row, GS = element distance of
// neighbouring chunks of the same line.
//   forward : L levels of cyclic reduction - at level l (stride s = 2^l) every second surviving row is eliminated:
//             it publishes (A, C, D) once and keeps them in registers; its two neighbours absorb it;
//   middle  : PCR over the GP / 2^L surviving rows (<= 8: three steps);
//   backward: eliminated rows recover x from their two (already solved) neighbours.
// The solutions of ALL rows end up in sol[q * STR + e].  Fully unrolled: GP, GS are compile-time.
// Precondition: the first row of a line enters with A == 0 and the last one with C == 0 (exact zeros: boundary and
// identity rows have a = 0 / c = 0, the slab-coupled modes zero them explicitly).  Elimination keeps them zero, which is
// why a row may read an arbitrary (finite) published row in place of a neighbour that does not exist.
// Shared-memory footprint (elements): crs = NR * STR (publications; `sol` may alias it - the publications are not
// read again once the PCR starts), pp = 2 * NR * (STR >> L) (PCR ping-pong, surviving rows only, compacted).
template <int GP> struct RedGeom { static constexpr int L = GP > 8 ? (GP == 16 ? 1 : GP == 32 ? 2 : 3) : 0; };
template <int NRHS, int GP, int NL> __host__ __device__ constexpr int reduced_scratch_elems() { return (2 + NRHS) * GP * NL + 2 * (2 + NRHS) * ((GP * NL) >> RedGeom<GP>::L); }

template <typename FT, int NRHS, int GP, int GS, int NL>
__device__ __forceinline__ void reduced_solve(FT *sys, FT *sol, int g, int e, FT Ain, FT Cin, const FT (&Din)[NRHS], FT (&X)[NRHS])
{
	constexpr int NR = 2 + NRHS;
	constexpr int STR = GP * NL;
	constexpr int L = RedGeom<GP>::L;
	constexpr int STRC = STR >> L;       // compacted stride of the PCR arrays
	FT A = Ain, Cc = Cin, D[NRHS];
#pragma unroll
	for (int q = 0; q < NRHS; q++) D[q] = Din[q];
	FT *crs = sys;                       // CR publications: NR arrays (each row publishes once, at its own element)
	FT *pp = sys + NR * STR;             // PCR ping-pong: 2 * NR compacted arrays (surviving rows only)
	// compacted element of a surviving row (g a multiple of 2^L): chunk index g >> L
	const int ec = GS == 1 ? ((e - g) >> L) + (g >> L) : (g >> L) * GS + (e - g * GS);
	int my_level = -1;                   // level at which this row was eliminated (-1: survives into the PCR)
#pragma unroll
	for (int lv = 0; lv < L; lv++) {
		const int s = 1 << lv;
		const bool alive = (g & (s - 1)) == 0 && my_level < 0;
		const bool odd = alive && ((g >> lv) & 1);
		if (odd) {
			my_level = lv;
			crs[0 * STR + e] = A; crs[1 * STR + e] = Cc;
#pragma unroll
			for (int q = 0; q < NRHS; q++) crs[(2 + q) * STR + e] = D[q];
		}
		__syncthreads();
		if (alive && !odd) {
			// a row without a lower (upper) neighbour has A == 0 (C == 0) exactly - see the header comment - so it may
			// read ANY published row in its place: no bounds predicates, no zero fills
			const bool lo = g - s >= 0, hi = g + s < GP;
			const FT *l = crs + (lo ? e - s * GS : e + s * GS), *h = crs + (hi ? e + s * GS : e - s * GS);
			const FT Al = l[0 * STR], Cl = l[1 * STR];
			const FT Ah = h[0 * STR], Ch = h[1 * STR];
			const FT r = rcp<FT>(FT(1) - A * Cl - Cc * Ah);
#pragma unroll
			for (int q = 0; q < NRHS; q++) {
				const FT Dl = l[(2 + q) * STR], Dh = h[(2 + q) * STR];
				D[q] = (D[q] - A * Dl - Cc * Dh) * r;
			}
			A = -A * Al * r;
			Cc = -Cc * Ch * r;
		}
	}
	const bool survivor = my_level < 0 && (g & ((1 << L) - 1)) == 0;
#pragma unroll
	for (int st = 0; (1 << (L + st)) < GP; st++) {
		const int s = 1 << (L + st), sc = 1 << st;      // stride in chunks / in surviving rows
		FT *w = pp + (st & 1) * NR * STRC;
		if (survivor) {
			w[0 * STRC + ec] = A; w[1 * STRC + ec] = Cc;
#pragma unroll
			for (int q = 0; q < NRHS; q++) w[(2 + q) * STRC + ec] = D[q];
		}
		__syncthreads();
		if (survivor) {
			const bool lo = g - s >= 0, hi = g + s < GP;
			const FT *l = w + (lo ? ec - sc * GS : ec + sc * GS), *h = w + (hi ? ec + sc * GS : ec - sc * GS);
			const FT Al = l[0 * STRC], Cl = l[1 * STRC];
			const FT Ah = h[0 * STRC], Ch = h[1 * STRC];
			const FT r = rcp<FT>(FT(1) - A * Cl - Cc * Ah);
#pragma unroll
			for (int q = 0; q < NRHS; q++) {
				const FT Dl = l[(2 + q) * STRC], Dh = h[(2 + q) * STRC];
				D[q] = (D[q] - A * Dl - Cc * Dh) * r;
			}
			A = -A * Al * r;
			Cc = -Cc * Ch * r;
		}
	}
	if (survivor) {
#pragma unroll
		for (int q = 0; q < NRHS; q++) sol[q * STR + e] = D[q];
	}
	__syncthreads();
#pragma unroll
	for (int lv = L - 1; lv >= 0; lv--) {
		const int s = 1 << lv;
		if (my_level == lv) {
			const bool lo = g - s >= 0, hi = g + s < GP;
			const int el = lo ? e - s * GS : e + s * GS, eh = hi ? e + s * GS : e - s * GS;
#pragma unroll
			for (int q = 0; q < NRHS; q++) {
				const FT xl = sol[q * STR + el], xh = sol[q * STR + eh];
				D[q] = D[q] - A * xl - Cc * xh;
				sol[q * STR + e] = D[q];
			}
		}
		__syncthreads();
	}
#pragma unroll
	for (int q = 0; q < NRHS; q++) X[q] = D[q];
}

// ---- line I/O: 8 consecutive rows of this thread's chunk ----------------------------------------------------
// X / Y sweeps: rows are `stride` apart, lanes of a warp sit on neighbouring k (coalesced 64-byte segments);
// off[i] = element offset of row i (clamped into the line, computed once and shared by all fields).
// Z sweep: the 8 rows are 8 contiguous elements (64 bytes in fp64) -> 128-bit vector accesses at off[0].
template <typename FT> struct Vec16;
template <> struct Vec16<double> { typedef double2 type; static constexpr int N = 2; };
template <> struct Vec16<float> { typedef float4 type; static constexpr int N = 4; };

// 256-bit global accesses (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256): a z-chunk of 8 fp64 values is two of them (one
// for fp32), and every access covers whole 32-byte sectors even though neighbouring lanes are a chunk apart
__device__ __forceinline__ void ldg256(const double *p, double (&o)[4])
{
	asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p));
}
__device__ __forceinline__ void stg256(double *p, const double (&v)[4])
{
	asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ void ldg256(const float *p, float (&o)[8])
{
	asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(float *p, const float (&v)[8])
{
	asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]),
	             "f"(v[6]), "f"(v[7]) : "memory");
}

template <typename FT, int DIR>
__device__ __forceinline__ void load8(const FT *__restrict__ p, const int (&off)[M], FT (&o)[M])
{
	if (DIR == 2) {
		constexpr int N = 32 / (int)sizeof(FT);                        // elements per 256-bit access
		const FT *q = p + off[0];                                       // 64-byte aligned: r0 % 8 == 0, lines 128-byte aligned
#pragma unroll
		for (int v = 0; v < M / N; v++) {
			FT t[N];
			ldg256(q + v * N, t);
#pragma unroll
			for (int k = 0; k < N; k++) o[v * N + k] = t[k];
		}
	} else {
#pragma unroll
		for (int i = 0; i < M; i++) o[i] = p[off[i]];
	}
}

// store rows whose bit is set in `mask` (Z: whole chunk with vector stores when all 8 bits are set)
template <typename FT, int DIR>
__device__ __forceinline__ void store8(FT *__restrict__ p, const int (&off)[M], unsigned mask, const FT (&v)[M])
{
	if (DIR == 2) {
		if (mask == 0xffu) {
			constexpr int N = 32 / (int)sizeof(FT);
			FT *q = p + off[0];
#pragma unroll
			for (int w = 0; w < M / N; w++) {
				FT t[N];
#pragma unroll
				for (int k = 0; k < N; k++) t[k] = v[w * N + k];
				stg256(q + w * N, t);
			}
		} else {
#pragma unroll
			for (int i = 0; i < M; i++)
				if (mask & (1u << i)) p[off[0] + i] = v[i];
		}
	} else {
#pragma unroll
		for (int i = 0; i < M; i++)
			if (mask & (1u << i)) p[off[i]] = v[i];
	}
}

// nonlinear-layer relaxation temp' = (temp + next) / 2 on the fluid cells of a chunk (MergeFieldTo, reference
// TimeLayer3D.h:415-436), twice when the post-X merge is folded in; all-fluid chunks skip the per-row selects
template <typename FT, int DIR>
__device__ __forceinline__ void relax8(FT (&tq)[M], const FT (&x)[M], unsigned inmask, int twice)
{
	// (the shortcut pays along z only: in the 512-thread x / y kernels the extra code costs more in register
	// allocation than the selects it saves - measured)
	if (DIR == 2 && inmask == 0xffu) {
#pragma unroll
		for (int i = 0; i < M; i++) tq[i] = (tq[i] + x[i]) * FT(0.5);
		if (twice) {
#pragma unroll
			for (int i = 0; i < M; i++) tq[i] = (tq[i] + x[i]) * FT(0.5);
		}
	} else {
#pragma unroll
		for (int i = 0; i < M; i++) tq[i] = (inmask & (1u << i)) ? (tq[i] + x[i]) * FT(0.5) : tq[i];
		if (twice) {
#pragma unroll
			for (int i = 0; i < M; i++) tq[i] = (inmask & (1u << i)) ? (tq[i] + x[i]) * FT(0.5) : tq[i];
		}
	}
}

// Slab-decomposed runs: the boundary x-planes of a sweep's outputs also go straight into the x-neighbours' guard
// planes (SweepArgs::push_*: the neighbour slab's buffer, in peer memory when it lives on another GPU), so that no
// separate halo exchange follows the sweep.  y / z lines: the CTAs of planes 0 and nx-1 repeat their stores; x lines
// (coupled sweep, MODE 2): the first row of chunk 0 and the last row of the last chunk.
template <typename FT, int DIR, int MODE>
__device__ __forceinline__ void push_planes(const SweepArgs<FT> &A, int q, int pi, int g, int GL, const int (&off)[M],
                                            unsigned full, unsigned segfull, const FT (&tq)[M], const FT (&x)[M])
{
	const Layout &L = A.L;
	if (DIR != 0) {
		if (MODE != 0) return;
		// `off` holds slab offsets (pi + 1) * plane + j * nzp + k: shift the plane pointers accordingly
		if (pi == 0 && A.push_lo[q]) store8<FT, DIR>(A.push_lo[q] - L.plane, off, full, tq);
		if (pi == L.nx - 1 && A.push_hi[q]) store8<FT, DIR>(A.push_hi[q] - (long long)L.nx * L.plane, off, full, tq);
	} else if (MODE == 2) {
		if (g == 0) {
			const long long o = off[0] - L.plane;
			if (A.push_lo[q] && (full & 1u)) A.push_lo[q][o] = tq[0];
			if (A.pushn_lo[q] && (segfull & 1u)) A.pushn_lo[q][o] = x[0];
		}
		if (g == GL - 1) {
			const long long o = off[M - 1] - (long long)L.nx * L.plane;
			if (A.push_hi[q] && (full & 0x80u)) A.push_hi[q][o] = tq[M - 1];
			if (A.pushn_hi[q] && (segfull & 0x80u)) A.pushn_hi[q][o] = x[M - 1];
		}
	}
}

// central difference along the line for the 8 rows of a chunk (lo / hi = rows r0-1 and r0+8)
template <typename FT>
__device__ __forceinline__ FT cdiff(const FT (&f)[M], FT lo, FT hi, int i, FT inv2h)
{
	const FT m = i == 0 ? lo : f[i == 0 ? 0 : i - 1], p = i == M - 1 ? hi : f[i == M - 1 ? M - 1 : i + 1];
	return (p - m) * inv2h;
}

// One chunk: eliminate the 7 interior rows of (a, b, c | d[NRHS]) given row by row, keep the separator raw.
// cp/lp/dp hold c', the left spike and d' of the interior rows; entry M-1 holds the raw separator (lp = a, cp = c).
#define CMC_ELIM_ROW(i, a, b, c)                                                   \
	if ((i) == M - 1) { lp[i] = (a); cp[i] = (c); b7 = (b); }                      \
	else if ((i) == 0) { rr = rcp<FT>(b); cp[0] = (c) * rr; lp[0] = (a) * rr; }    \
	else { rr = rcp<FT>((b) - (a) * cp[(i) - 1]); cp[i] = (c) * rr; lp[i] = -(a) * lp[(i) - 1] * rr; }


} // namespace cmc
