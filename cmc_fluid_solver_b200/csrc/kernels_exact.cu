// kernels_exact.cu - CMC_MODE_EXACT line sweeps: sequential Thomas in the reference's operation
// order (Common::SolveTridiagonal, reference src/Common/Algorithms.h:21-38) over whole grid lines.
// This translation unit is compiled with -fmad=false so no multiply-add is contracted: results are
// bit-identical with the reference CPU solver (x86-64 -O2, no FMA).
//
// Whole-line formulation: one thread walks one grid line.  Cells outside any segment are skipped;
// a segment's first/last cell carries the ApplyBC0/ApplyBC1 row, which decouples it from its
// neighbours exactly (a = 0 / c = 0), so solving per line equals solving per Segment3D.
//
// Two passes per sweep (forward elimination, back substitution).  d' is kept in the output layer
// `next` in place, c' (one per matrix: u,v,w share theirs) in two scratch fields.
#include "kernels.h"
#include "rows.cuh"

namespace cmc {

// the grid line of a thread: cell p of it is at line.at(L, p)
struct ExactLine {
	int n, i, j, k;
	template <int DIR>
	__device__ __forceinline__ long long at(const Layout &L, int p) const
	{
		return DIR == 0 ? L.idx(p, j, k) : DIR == 1 ? L.idx(i, p, k) : L.idx(i, j, p);
	}
};

template <int DIR>
__device__ __forceinline__ bool line_of_thread(const Layout &L, long long t, ExactLine &ln)
{
	ln.i = ln.j = ln.k = 0;
	if (DIR == 0) {            // lines along x: one per (j, k), lanes along k (coalesced)
		ln.k = (int)(t % L.nz); ln.j = (int)(t / L.nz);
		if (ln.j >= L.ny) return false;
		ln.n = L.nx;
	} else if (DIR == 1) {     // lines along y: one per (i, k), lanes along k (coalesced)
		ln.k = (int)(t % L.nz); ln.i = (int)(t / L.nz);
		if (ln.i >= L.nx) return false;
		ln.n = L.ny;
	} else {                   // lines along z: one per (i, j)
		ln.j = (int)(t % L.ny); ln.i = (int)(t / L.ny);
		if (ln.i >= L.nx) return false;
		ln.n = L.nz;
	}
	return true;
}

template <typename FT, int DIR>
__global__ void __launch_bounds__(128) k_exact_forward(const SweepArgs<FT> A)
{
	ExactLine ln;
	if (!line_of_thread<DIR>(A.L, (long long)blockIdx.x * blockDim.x + threadIdx.x, ln)) return;
	const int n = ln.n;
	RowConst<FT> K; K.init(A, DIR);
	const long long sx = A.L.plane, sz = 1;
	FT cpv = 0, cpT = 0, dp[4] = {0, 0, 0, 0};
	if (DIR == 0 && A.xs_me > 0) {
		// slab of a decomposed grid: the line continues from the lower slab - its last row's c', d' were copied into this slab's
		// guard plane, so the elimination goes on with exactly the operations of the undivided line
		const long long im = ln.at<DIR>(A.L, -1);
		cpv = A.cv[im]; cpT = A.cT[im];
		dp[0] = A.next[0][im]; dp[1] = A.next[1][im]; dp[2] = A.next[2][im]; dp[3] = A.next[3][im];
	}
	for (int p = 0; p < n; p++) {
		const long long id = ln.at<DIR>(A.L, p);
		const int jj = DIR == 1 ? p : ln.j;
		const long long syp = A.L.jup(jj), sym = A.L.jdn(jj);
		const unsigned r = A.role[id];
		if (!(r & R_SEG)) continue;
		if (r & R_END) {        // ApplyBC1 row closes the running segment: c[n-1] = 0
			FT a_v, b_v, a_T, b_T, d[4];
			boundary_row(A, r, id, a_v, b_v, a_T, b_T, d);
			const FT den_v = b_v - a_v * cpv, den_T = b_T - a_T * cpT;
			if (!(r & R_START)) {   // (a shared cell keeps the NEXT segment's start row; see k_exact_backward)
				A.next[0][id] = (d[0] - dp[0] * a_v) / den_v;
				A.next[1][id] = (d[1] - dp[1] * a_v) / den_v;
				A.next[2][id] = (d[2] - dp[2] * a_v) / den_v;
				A.next[3][id] = (d[3] - dp[3] * a_T) / den_T;   // c'[n-1] = 0/den is never read
			}
		}
		if (r & R_START) {      // ApplyBC0 row: c0 /= b0, d0 /= b0
			FT c_v, b_v, c_T, b_T, d[4];
			boundary_row(A, r, id, c_v, b_v, c_T, b_T, d);
			cpv = c_v / b_v; cpT = c_T / b_T;
			dp[0] = d[0] / b_v; dp[1] = d[1] / b_v; dp[2] = d[2] / b_v; dp[3] = d[3] / b_T;
		} else if (r & R_INT) {
			FT a_v, c_v, a_T, c_T, d[4];
			build_interior_row<FT, DIR>(A, K, id, sx, syp, sym, sz, a_v, c_v, a_T, c_T, d);
			const FT den_v = K.b_v - a_v * cpv, den_T = K.b_T - a_T * cpT;
			cpv = c_v / den_v; cpT = c_T / den_T;
			dp[0] = (d[0] - dp[0] * a_v) / den_v;
			dp[1] = (d[1] - dp[1] * a_v) / den_v;
			dp[2] = (d[2] - dp[2] * a_v) / den_v;
			dp[3] = (d[3] - dp[3] * a_T) / den_T;
		} else continue;
		A.cv[id] = cpv; A.cT[id] = cpT;
		A.next[0][id] = dp[0]; A.next[1][id] = dp[1]; A.next[2][id] = dp[2]; A.next[3][id] = dp[3];
	}
}

template <typename FT, int DIR>
__global__ void __launch_bounds__(128) k_exact_backward(const SweepArgs<FT> A)
{
	ExactLine ln;
	if (!line_of_thread<DIR>(A.L, (long long)blockIdx.x * blockDim.x + threadIdx.x, ln)) return;
	const int n = ln.n;
	FT x[4] = {0, 0, 0, 0};
	if (DIR == 0 && A.xs_me + 1 < A.xs_P) {     // the value the upper slab's back substitution carried across its first row
		const long long ip = ln.at<DIR>(A.L, n);
		x[0] = A.next[0][ip]; x[1] = A.next[1][ip]; x[2] = A.next[2][ip]; x[3] = A.next[3][ip];
	}
	for (int p = n - 1; p >= 0; p--) {
		const long long id = ln.at<DIR>(A.L, p);
		const unsigned r = A.role[id];
		if (!(r & R_SEG)) continue;
		if ((r & R_END) && !(r & R_START)) {   // x[n-1] = d[n-1], already in place
			x[0] = A.next[0][id]; x[1] = A.next[1][id]; x[2] = A.next[2][id]; x[3] = A.next[3][id];
			continue;
		}
		// interior or start row: x[i] = d[i] - c[i] * x[i+1]
		const FT cv = A.cv[id], cT = A.cT[id];
		x[0] = A.next[0][id] - cv * x[0];
		x[1] = A.next[1][id] - cv * x[1];
		x[2] = A.next[2][id] - cv * x[2];
		x[3] = A.next[3][id] - cT * x[3];
		A.next[0][id] = x[0]; A.next[1][id] = x[1]; A.next[2][id] = x[2]; A.next[3][id] = x[3];
		if ((r & R_END) && (r & R_START)) {
			// Shared cell: the later segment's value stays in `next` (the reference writes segments in list
			// order, AdiSolver3D.cpp:596-602); the earlier segment still needs ITS last unknown, which is its
			// ApplyBC1 row eliminated against row p-1 - recompute it with the forward pass's exact operations.
			const long long im = ln.at<DIR>(A.L, p - 1);
			FT a_v, b_v, a_T, b_T, d[4];
			boundary_row(A, r, id, a_v, b_v, a_T, b_T, d);
			const FT den_v = b_v - a_v * A.cv[im], den_T = b_T - a_T * A.cT[im];
			x[0] = (d[0] - A.next[0][im] * a_v) / den_v;
			x[1] = (d[1] - A.next[1][im] * a_v) / den_v;
			x[2] = (d[2] - A.next[2][im] * a_v) / den_v;
			x[3] = (d[3] - A.next[3][im] * a_T) / den_T;
		}
	}
	if (DIR == 0 && A.xs_me > 0) {              // (the guard plane's copy of the lower slab's d' has served its purpose above)
		const long long im = ln.at<DIR>(A.L, -1);
		A.next[0][im] = x[0]; A.next[1][im] = x[1]; A.next[2][im] = x[2]; A.next[3][im] = x[3];
	}
}

template <typename FT>
void launch_exact_sweep(int dir, const SweepArgs<FT> &A, cudaStream_t s, long long *launches)
{
	const Layout &L = A.L;
	const long long lines = dir == 0 ? (long long)L.ny * L.nz : dir == 1 ? (long long)L.nx * L.nz : (long long)L.nx * L.ny;
	const int bs = 128;
	const unsigned grid = (unsigned)((lines + bs - 1) / bs);
	switch (dir) {
	case 0: k_exact_forward<FT, 0><<<grid, bs, 0, s>>>(A); k_exact_backward<FT, 0><<<grid, bs, 0, s>>>(A); break;
	case 1: k_exact_forward<FT, 1><<<grid, bs, 0, s>>>(A); k_exact_backward<FT, 1><<<grid, bs, 0, s>>>(A); break;
	default: k_exact_forward<FT, 2><<<grid, bs, 0, s>>>(A); k_exact_backward<FT, 2><<<grid, bs, 0, s>>>(A); break;
	}
	if (launches) *launches += 2;
}

// the two passes on their own: slabs of a decomposed grid run them as a chain along x (forward from the first slab to the last,
// back substitution from the last to the first, cmc_adi.cu)
template <typename FT>
void launch_exact_x_pass(bool forward, const SweepArgs<FT> &A, cudaStream_t s, long long *launches)
{
	const Layout &L = A.L;
	const long long lines = (long long)L.ny * L.nz;
	const int bs = 128;
	const unsigned grid = (unsigned)((lines + bs - 1) / bs);
	if (forward) k_exact_forward<FT, 0><<<grid, bs, 0, s>>>(A);
	else k_exact_backward<FT, 0><<<grid, bs, 0, s>>>(A);
	if (launches) *launches += 1;
}
template void launch_exact_x_pass<float>(bool, const SweepArgs<float> &, cudaStream_t, long long *);
template void launch_exact_x_pass<double>(bool, const SweepArgs<double> &, cudaStream_t, long long *);

template void launch_exact_sweep<float>(int, const SweepArgs<float> &, cudaStream_t, long long *);
template void launch_exact_sweep<double>(int, const SweepArgs<double> &, cudaStream_t, long long *);

// ---- standalone batched Thomas (unit test of the operation order) ---------------------------------
template <typename FT>
__global__ void k_thomas_batch(int nsys, int n, FT *a, FT *b, FT *c, FT *d, FT *x)
{
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nsys) return;
	a += (size_t)s * n; b += (size_t)s * n; c += (size_t)s * n; d += (size_t)s * n; x += (size_t)s * n;
	c[n - 1] = FT(0.0);
	c[0] = c[0] / b[0];
	d[0] = d[0] / b[0];
	for (int i = 1; i < n; i++) {
		const FT den = b[i] - a[i] * c[i - 1];
		c[i] = c[i] / den;
		d[i] = (d[i] - d[i - 1] * a[i]) / den;
	}
	x[n - 1] = d[n - 1];
	for (int i = n - 2; i >= 0; i--) x[i] = d[i] - c[i] * x[i + 1];
}

template <typename FT>
void launch_thomas_batch(int nsys, int n, FT *a, FT *b, FT *c, FT *d, FT *x, cudaStream_t s)
{
	k_thomas_batch<FT><<<(nsys + 63) / 64, 64, 0, s>>>(nsys, n, a, b, c, d, x);
}
template void launch_thomas_batch<float>(int, int, float *, float *, float *, float *, float *, cudaStream_t);
template void launch_thomas_batch<double>(int, int, double *, double *, double *, double *, double *, cudaStream_t);

} // namespace cmc
