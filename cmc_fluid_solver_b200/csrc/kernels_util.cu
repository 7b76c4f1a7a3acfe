// kernels_util.cu - line-descriptor build, masked copy / merge, boundary refresh, divergence
// residual reduction and the GetLayer readback kernels.  All elementwise kernels are HBM-bound
// streaming kernels: grid-stride over z-lines with the k index on the lanes (coalesced).
#include "kernels.h"

namespace cmc {

static inline unsigned grid_for(long long work, int bs, int max_blocks = 148 * 16)
{
	long long g = (work + bs - 1) / bs;
	if (g > max_blocks) g = max_blocks;
	if (g < 1) g = 1;
	return (unsigned)g;
}

// ncode byte (dense global layout (i*ny + j)*nz + k): bits 0-1 NodeType, bit 2 bc_vel == FREE, bit 3 bc_temp == FREE
__device__ __forceinline__ unsigned code_type(unsigned c) { return c & 3u; }

// ---- Grid3D::GenerateListSegments (reference Grid3D.cpp:47-127) as a device scan ----------------------
// One thread walks one grid line: a segment starts at the cell BEFORE the first NODE_IN of a run and ends
// at the first non-IN cell after it; a run that reaches the end of the line unterminated is dropped.
template <int DIR>
__global__ void k_build_roles(const Layout G, const uint8_t *__restrict__ ncode, const Layout L, uint8_t *role,
                              unsigned long long *seg_count)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	int n;               // cells along the (global) line
	long long gbase, gstride;   // into ncode (dense global)
	int li = 0, lj = 0, lk = 0;   // the line in the slab layout: role cell p of the line = LIDX(p)
	const int gny = G.ny, gnz = G.nz;
	if (DIR == 0) {
		const int k = (int)(t % gnz), j = (int)(t / gnz);
		if (j >= gny) return;
		n = G.nx; gbase = (long long)j * gnz + k; gstride = (long long)gny * gnz;
		lj = j; lk = k;
	} else if (DIR == 1) {
		const int k = (int)(t % gnz), i = (int)(t / gnz);
		if (i >= L.nx) return;
		n = gny; gbase = ((long long)(i + L.x0) * gny) * gnz + k; gstride = gnz;
		li = i; lk = k;
	} else {
		const int j = (int)(t % gny), i = (int)(t / gny);
		if (i >= L.nx) return;
		n = gnz; gbase = ((long long)(i + L.x0) * gny + j) * gnz; gstride = 1;
		li = i; lj = j;
	}
#define LIDX(p) (DIR == 0 ? L.idx((p) - L.x0, lj, lk) : DIR == 1 ? L.idx(li, (p), lk) : L.idx(li, lj, (p)))
	int state = 0, start = 0, prev_end = -1;
	unsigned long long count = 0, shared_free = 0;
	for (int p = 0; p + 1 < n; p++) {
		if (code_type(ncode[gbase + (long long)(p + 1) * gstride]) == 0u /* NODE_IN */) {
			if (state == 0) start = p;
			state = 1;
		} else if (state == 1) {
			const int end = p + 1;
			for (int q = start; q <= end; q++) {
				if (DIR == 0 && (q < L.x0 || q >= L.x0 + L.nx)) continue;   // other slabs' cells
				const unsigned bits = q == start ? R_START : q == end ? R_END : R_INT;
				role[LIDX(q)] |= (uint8_t)bits;
			}
			// a cell that ends one segment and starts the next one holds TWO unknowns when its boundary row is
			// BC_FREE; the whole-line fast solver cannot represent that (exact mode can) - count such cells
			if (start == prev_end) {
				const unsigned sc = ncode[gbase + (long long)start * gstride];
				if (sc & 12u) shared_free++;
				// fast solver: one unknown per cell, so the earlier segment's last unknown (this cell) is eliminated
				// into its last interior row (start - 1), which also records the shared cell's boundary kinds
				const int q = start - 1;
				if (!(DIR == 0 && (q < L.x0 || q >= L.x0 + L.nx)))
					role[LIDX(q)] |= (uint8_t)(R_PRE | ((sc & 4u) ? R_VFREE : 0u) | ((sc & 8u) ? R_TFREE : 0u));
			}
			prev_end = end;
			count++;
			state = 0;
		}
	}
#undef LIDX
	if (count) atomicAdd(seg_count, count);
	if (shared_free) atomicAdd(seg_count + 4, shared_free);
}

void launch_build_roles(int dir, const Layout &G, const uint8_t *ncode, const Layout &L, uint8_t *role,
                        unsigned long long *seg_count, cudaStream_t s, long long *launches)
{
	const long long lines = dir == 0 ? (long long)G.ny * G.nz : dir == 1 ? (long long)L.nx * G.nz : (long long)L.nx * G.ny;
	const int bs = 128;
	const unsigned grid = (unsigned)((lines + bs - 1) / bs);
	if (dir == 0) k_build_roles<0><<<grid, bs, 0, s>>>(G, ncode, L, role, seg_count);
	else if (dir == 1) k_build_roles<1><<<grid, bs, 0, s>>>(G, ncode, L, role, seg_count);
	else k_build_roles<2><<<grid, bs, 0, s>>>(G, ncode, L, role, seg_count);
	if (launches) (*launches)++;
}

// The same descriptors for the x direction from a WINDOW of the node codes (cmc_adi3d_set_nodes_slab: a rank that only
// knows its own planes plus two on either side).  Needs no scan: with no NODE_IN cell on the two x-faces of the grid (the
// reference's loaders always leave a NODE_OUT rim; checked by the caller) every run of fluid cells is closed
// inside the grid, and a cell's role follows from its own type and those of the cells at -1, +1, +2:
//   interior  : IN                          start : not IN, next IN            end : not IN, previous IN
//   R_PRE     : IN, next not IN, next-but-one IN (the next cell is shared by two segments; its boundary kinds ride here)
// `ncode` is addressed by GLOBAL plane (the caller passes the window's base shifted by its first plane).
__global__ void k_build_roles_x_local(const Layout G, const uint8_t *__restrict__ ncode, const Layout L, uint8_t *role,
                                      unsigned long long *seg_count)
{
	const long long rows = (long long)L.nx * L.ny;
	unsigned long long count = 0, shared_free = 0;
	const long long gplane = (long long)G.ny * G.nz;
	for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
		const int i = (int)(row / L.ny), j = (int)(row % L.ny);
		const int p = i + L.x0;
		const uint8_t *c0 = ncode + ((long long)p * G.ny + j) * G.nz;
		const long long dst = L.idx(i, j, 0);
		for (int k = threadIdx.x; k < L.nz; k += blockDim.x) {
			auto code = [&](int d) -> unsigned { const int q = p + d; return (q < 0 || q >= G.nx) ? 1u /* NODE_OUT */ : (unsigned)c0[(long long)d * gplane + k]; };
			const unsigned cm = code(-1), c = code(0), cp = code(1), cpp = code(2);
			const bool in = code_type(c) == 0u, inm = code_type(cm) == 0u, inp = code_type(cp) == 0u, inpp = code_type(cpp) == 0u;
			unsigned bits = 0;
			if (in) {
				bits |= R_INT;
				if (!inp && inpp) bits |= R_PRE | ((cp & 4u) ? R_VFREE : 0u) | ((cp & 8u) ? R_TFREE : 0u);
			} else {
				if (inp) bits |= R_START;
				if (inm) { bits |= R_END; count++; }
				if (inp && inm && (c & 12u)) shared_free++;
			}
			if (bits) role[dst + k] |= (uint8_t)bits;
		}
	}
	for (int o = 16; o > 0; o >>= 1) {
		count += __shfl_down_sync(0xffffffffu, count, o);
		shared_free += __shfl_down_sync(0xffffffffu, shared_free, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if (count) atomicAdd(seg_count, count);
		if (shared_free) atomicAdd(seg_count + 4, shared_free);
	}
}

void launch_build_roles_x_local(const Layout &G, const uint8_t *ncode_by_global_plane, const Layout &L, uint8_t *role,
                                unsigned long long *seg_count, cudaStream_t s, long long *launches)
{
	k_build_roles_x_local<<<grid_for((long long)L.nx * L.ny, 1), 128, 0, s>>>(G, ncode_by_global_plane, L, role, seg_count);
	if (launches) (*launches)++;
}

// number of NODE_IN cells in one global x-plane of the (windowed) node codes
__global__ void k_count_in_plane(const Layout G, const uint8_t *__restrict__ ncode, int gplane_index, unsigned long long *out)
{
	const long long n = (long long)G.ny * G.nz;
	const uint8_t *src = ncode + (long long)gplane_index * n;
	unsigned long long c = 0;
	for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) c += code_type(src[t]) == 0u;
	for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

void launch_count_in_plane(const Layout &G, const uint8_t *ncode_by_global_plane, int gplane_index, unsigned long long *out, cudaStream_t s)
{
	k_count_in_plane<<<64, 256, 0, s>>>(G, ncode_by_global_plane, gplane_index, out);
}

__global__ void k_role_type_bits(const Layout G, const uint8_t *__restrict__ ncode, const Layout L,
                                 uint8_t *rx, uint8_t *ry, uint8_t *rz)
{
	// the guard planes that hold a neighbouring slab's boundary plane get their type bits too (in the z array only):
	// k_update_boundaries refreshes them together with the slab's own cells
	const int ilo = L.x0 > 0 ? -1 : 0, ihi = L.x0 + L.nx < G.nx ? L.nx + 1 : L.nx;
	const long long rows = (long long)(ihi - ilo) * L.ny;
	for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
		const int i = (int)(row / L.ny) + ilo, j = (int)(row % L.ny);
		const bool guard = i < 0 || i >= L.nx;
		const uint8_t *src = ncode + ((long long)(i + L.x0) * G.ny + j) * G.nz;
		const long long dst = L.idx(i, j, 0);
		for (int k = threadIdx.x; k < L.nz; k += blockDim.x) {
			const unsigned c = src[k];
			const unsigned ty = code_type(c);
			// the boundary kinds of a NODE_IN cell are never read by the reference (its rows are interior rows); on such
			// a cell R_VFREE / R_TFREE are reserved for the folded shared cell that follows it (R_PRE, k_build_roles)
			unsigned bits = (ty == 0u ? R_IN : 0u) | ((ty == 2u || ty == 3u) ? R_BV : 0u)
			              | ((ty != 0u && (c & 4u)) ? R_VFREE : 0u) | ((ty != 0u && (c & 8u)) ? R_TFREE : 0u);
			if (!guard) { rx[dst + k] = (uint8_t)bits; ry[dst + k] = (uint8_t)bits; }
			rz[dst + k] = (uint8_t)bits;
		}
	}
}

void launch_role_type_bits(const Layout &G, const uint8_t *ncode, const Layout &L, uint8_t *rx, uint8_t *ry, uint8_t *rz,
                           cudaStream_t s, long long *launches)
{
	k_role_type_bits<<<grid_for((long long)(L.nx + 2) * L.ny, 1), 128, 0, s>>>(G, ncode, L, rx, ry, rz);
	if (launches) (*launches)++;
}

// ---- elementwise layer kernels --------------------------------------------------------------------------
// rows = interior z-lines of the slab; threads stride over k.  4 fields per kernel.
template <typename FT, typename F>
__device__ __forceinline__ void for_each_cell(const Layout &L, F f)
{
	const long long rows = (long long)L.nx * L.ny;
	const int lanes_per_row = 128;                  // one CTA = 128 threads = one row at a time
	(void)lanes_per_row;
	for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
		const long long base = L.idx((int)(row / L.ny), (int)(row % L.ny), 0);
		for (int k = threadIdx.x; k < L.nz; k += blockDim.x) f(base + k);
	}
}

template <typename FT>
__global__ void k_copy_masked(const Layout L, const uint8_t *__restrict__ role, unsigned mask, ConstLayerPtrs<FT> src, LayerPtrs<FT> dst)
{
	for_each_cell<FT>(L, [&](long long id) {
		if (role[id] & mask) {
			dst.f[0][id] = src.f[0][id]; dst.f[1][id] = src.f[1][id];
			dst.f[2][id] = src.f[2][id]; dst.f[3][id] = src.f[3][id];
		}
	});
}

template <typename FT>
__global__ void k_merge(const Layout L, const uint8_t *__restrict__ role, ConstLayerPtrs<FT> src, LayerPtrs<FT> dst)
{
	for_each_cell<FT>(L, [&](long long id) {
		if (role[id] & R_IN) {
			dst.f[0][id] = (dst.f[0][id] + src.f[0][id]) / 2;
			dst.f[1][id] = (dst.f[1][id] + src.f[1][id]) / 2;
			dst.f[2][id] = (dst.f[2][id] + src.f[2][id]) / 2;
			dst.f[3][id] = (dst.f[3][id] + src.f[3][id]) / 2;
		}
	});
}

template <typename FT>
__global__ void k_merge_to(const Layout L, const uint8_t *__restrict__ role, ConstLayerPtrs<FT> tmp, ConstLayerPtrs<FT> nxt, LayerPtrs<FT> out)
{
	for_each_cell<FT>(L, [&](long long id) {
		const bool in = role[id] & R_IN;
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const FT t = tmp.f[q][id];
			out.f[q][id] = in ? (t + nxt.f[q][id]) / 2 : t;
		}
	});
}

// (planes -1 and nx included: the guard planes mirror the neighbouring slabs' boundary planes, whose BOUND / VALVE cells
// the neighbours refresh at the same moment; guard planes at the ends of the grid carry no type bits)
template <typename FT>
__global__ void k_update_boundaries(const Layout L, const uint8_t *__restrict__ role, ConstLayerPtrs<FT> nodev, LayerPtrs<FT> cur)
{
	const long long rows = (long long)(L.nx + 2) * L.ny;
	for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
		const long long base = L.idx((int)(row / L.ny) - 1, (int)(row % L.ny), 0);
		for (int k = threadIdx.x; k < L.nz; k += blockDim.x) {
			const long long id = base + k;
			if (role[id] & R_BV) {
				cur.f[0][id] = nodev.f[0][id]; cur.f[1][id] = nodev.f[1][id];
				cur.f[2][id] = nodev.f[2][id]; cur.f[3][id] = nodev.f[3][id];
			}
		}
	}
}

// dense [nx][ny][nz] arrays -> the padded / y-blocked layer (asynchronous layer upload, cmc_adi3d_write_layer_commit)
template <typename FT>
__global__ void k_scatter_dense(const Layout L, ConstLayerPtrs<FT> src, LayerPtrs<FT> dst)
{
	const long long rows = (long long)L.nx * L.ny;
	for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
		const long long d0 = L.idx((int)(row / L.ny), (int)(row % L.ny), 0), s0 = row * L.nz;
		for (int k = threadIdx.x; k < L.nz; k += blockDim.x) {
			dst.f[0][d0 + k] = src.f[0][s0 + k]; dst.f[1][d0 + k] = src.f[1][s0 + k];
			dst.f[2][d0 + k] = src.f[2][s0 + k]; dst.f[3][d0 + k] = src.f[3][s0 + k];
		}
	}
}
template <typename FT>
void launch_scatter_dense(const Layout &L, ConstLayerPtrs<FT> src, LayerPtrs<FT> dst, cudaStream_t s, long long *launches)
{
	k_scatter_dense<FT><<<grid_for((long long)L.nx * L.ny, 1), 128, 0, s>>>(L, src, dst);
	if (launches) (*launches)++;
}

template <typename FT>
__global__ void k_clear_out(const Layout L, const uint8_t *__restrict__ role, LayerPtrs<FT> layer, FT value)
{
	for_each_cell<FT>(L, [&](long long id) {
		if (!(role[id] & (R_IN | R_BV))) { layer.f[0][id] = value; layer.f[1][id] = value; layer.f[2][id] = value; layer.f[3][id] = value; }
	});
}

template <typename FT>
__global__ void k_fill(FT *dst, long long n, FT value)
{
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = value;
}

template <typename FT>
void launch_copy_full(const Layout &L, ConstLayerPtrs<FT> src, LayerPtrs<FT> dst, cudaStream_t s, long long *launches)
{
	// whole buffers including guard planes: plain device-to-device copies (DMA/SM copy kernels of the runtime)
	for (int q = 0; q < 4; q++) cudaMemcpyAsync(dst.f[q], src.f[q], sizeof(FT) * (size_t)L.total, cudaMemcpyDeviceToDevice, s);
	if (launches) *launches += 4;
}

template <typename FT>
void launch_copy_masked(const Layout &L, const uint8_t *role, unsigned mask, ConstLayerPtrs<FT> src, LayerPtrs<FT> dst, cudaStream_t s, long long *launches)
{
	k_copy_masked<FT><<<grid_for((long long)L.nx * L.ny, 1), 128, 0, s>>>(L, role, mask, src, dst);
	if (launches) (*launches)++;
}
template <typename FT>
void launch_merge(const Layout &L, const uint8_t *role, ConstLayerPtrs<FT> src, LayerPtrs<FT> dst, cudaStream_t s, long long *launches)
{
	k_merge<FT><<<grid_for((long long)L.nx * L.ny, 1), 128, 0, s>>>(L, role, src, dst);
	if (launches) (*launches)++;
}
template <typename FT>
void launch_merge_to(const Layout &L, const uint8_t *role, ConstLayerPtrs<FT> tmp, ConstLayerPtrs<FT> nxt, LayerPtrs<FT> out, cudaStream_t s, long long *launches)
{
	k_merge_to<FT><<<grid_for((long long)L.nx * L.ny, 1), 128, 0, s>>>(L, role, tmp, nxt, out);
	if (launches) (*launches)++;
}
template <typename FT>
void launch_update_boundaries(const Layout &L, const uint8_t *role, ConstLayerPtrs<FT> nodev, LayerPtrs<FT> cur, cudaStream_t s, long long *launches)
{
	k_update_boundaries<FT><<<grid_for((long long)(L.nx + 2) * L.ny, 1), 128, 0, s>>>(L, role, nodev, cur);
	if (launches) (*launches)++;
}
template <typename FT>
void launch_clear_out(const Layout &L, const uint8_t *role, LayerPtrs<FT> layer, FT value, cudaStream_t s, long long *launches)
{
	k_clear_out<FT><<<grid_for((long long)L.nx * L.ny, 1), 128, 0, s>>>(L, role, layer, value);
	if (launches) (*launches)++;
}
template <typename FT>
void launch_fill(FT *dst, long long n, FT value, cudaStream_t s, long long *launches)
{
	k_fill<FT><<<grid_for(n, 256), 256, 0, s>>>(dst, n, value);
	if (launches) (*launches)++;
}

// ---- TimeLayer3D::EvalDivError (reference TimeLayer3D.h:595-641) ----------------------------------------
// mean over NODE_IN cells (global i <= dimx-2, j <= dimy-2, k <= dimz-2) of |face-averaged flux divergence|.
// The per-cell expression keeps the reference's types: 8-point sums and the two spacing products in FTYPE,
// the division by 4.0 and the accumulation in double.  Cells on a low face (i, j or k == 0) would read out of
// bounds in the reference (undefined) and are skipped, as in oracle/adi3d_oracle.c.
// Deterministic two-stage reduction: warp shuffle -> block -> fixed-order final pass.
template <typename FT>
__global__ void __launch_bounds__(256) k_div_error(const Layout L, const uint8_t *__restrict__ role,
                                                    const FT *__restrict__ U, const FT *__restrict__ V, const FT *__restrict__ W,
                                                    FT dx, FT dy, FT dz, double *partials)
{
	double err = 0.0, cnt = 0.0;
	const long long sx = L.plane;
	const long long rows = (long long)L.nx * L.ny;
	for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
		const int i = (int)(row / L.ny), j = (int)(row % L.ny);
		const int gi = i + L.x0;
		if (gi == 0 || gi > L.gx - 2 || j == 0 || j > L.ny - 2) continue;
		const long long sy = L.jdn(j);
		const long long base = L.idx(i, j, 0);
		for (int k = 1 + threadIdx.x; k <= L.nz - 2; k += blockDim.x) {
			const long long id = base + k;
			if (!(role[id] & R_IN)) continue;
			const double err_x = (U[id] + U[id - sy] + U[id - sy - 1] + U[id - 1] -
				U[id - sx] - U[id - sx - sy] - U[id - sx - sy - 1] - U[id - sx - 1]) * dz * dy / 4.0;
			const double err_y = (V[id] + V[id - sx] + V[id - sx - 1] + V[id - 1] -
				V[id - sy] - V[id - sx - sy] - V[id - sx - sy - 1] - V[id - sy - 1]) * dx * dz / 4.0;
			const double err_z = (W[id] + W[id - sy] + W[id - sx - sy] + W[id - sx] -
				W[id - 1] - W[id - sy - 1] - W[id - sx - sy - 1] - W[id - sx - 1]) * dx * dy / 4.0;
			err += fabs(err_x + err_y + err_z);
			cnt += 1.0;
		}
	}
	__shared__ double s_err[8], s_cnt[8];
	for (int o = 16; o > 0; o >>= 1) {
		err += __shfl_down_sync(0xffffffffu, err, o);
		cnt += __shfl_down_sync(0xffffffffu, cnt, o);
	}
	const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	if (l == 0) { s_err[w] = err; s_cnt[w] = cnt; }
	__syncthreads();
	if (threadIdx.x == 0) {
		double e = 0.0, c = 0.0;
		for (int q = 0; q < 8; q++) { e += s_err[q]; c += s_cnt[q]; }
		partials[2 * blockIdx.x] = e; partials[2 * blockIdx.x + 1] = c;
	}
}

__global__ void k_div_error_final(const double *partials, int nblocks, double *result2)
{
	__shared__ double s_e[256], s_c[256];
	double e = 0.0, c = 0.0;
	for (int b = threadIdx.x; b < nblocks; b += 256) { e += partials[2 * b]; c += partials[2 * b + 1]; }
	s_e[threadIdx.x] = e; s_c[threadIdx.x] = c;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o) { s_e[threadIdx.x] += s_e[threadIdx.x + o]; s_c[threadIdx.x] += s_c[threadIdx.x + o]; }
		__syncthreads();
	}
	if (threadIdx.x == 0) { result2[0] = s_e[0]; result2[1] = s_c[0]; }
}

template <typename FT>
void launch_div_error(const Layout &L, const uint8_t *role, const FT *U, const FT *V, const FT *W, FT dx, FT dy, FT dz,
                      double *block_partials, int max_blocks, double *result2, cudaStream_t s, long long *launches)
{
	unsigned g = grid_for((long long)L.nx * L.ny, 1, max_blocks);
	k_div_error<FT><<<g, 256, 0, s>>>(L, role, U, V, W, dx, dy, dz, block_partials);
	k_div_error_final<<<1, 256, 0, s>>>(block_partials, (int)g, result2);
	if (launches) *launches += 2;
}

// ---- per-field checksums of one layer: sum and sum of squares over the cells of the slab that are not NODE_OUT ------
// (the role of the reference's sum_layer debug hook, AdiSolver3D.cpp:30-58).  partials: [8] per block, result8:
// (sum u, v, w, T, sum of squares u, v, w, T).  Fixed reduction tree: the result does not depend on scheduling.
template <typename FT>
__global__ void __launch_bounds__(256) k_field_sums(const Layout L, const uint8_t *__restrict__ role, ConstLayerPtrs<FT> f, double *partials)
{
	double acc[8] = {};
	for_each_cell<FT>(L, [&](long long id) {
		if (role[id] & (R_IN | R_BV)) {
#pragma unroll
			for (int q = 0; q < 4; q++) { const double v = (double)f.f[q][id]; acc[q] += v; acc[4 + q] += v * v; }
		}
	});
	__shared__ double sh[8][8];
	const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
	for (int q = 0; q < 8; q++) {
		double v = acc[q];
		for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
		if (l == 0) sh[q][w] = v;
	}
	__syncthreads();
	if (threadIdx.x < 8) {
		double v = 0.0;
		for (int k = 0; k < 8; k++) v += sh[threadIdx.x][k];
		partials[8 * blockIdx.x + threadIdx.x] = v;
	}
}

__global__ void k_field_sums_final(const double *partials, int nblocks, double *result8)
{
	const int q = threadIdx.x;
	if (q >= 8) return;
	double v = 0.0;
	for (int b = 0; b < nblocks; b++) v += partials[8 * b + q];
	result8[q] = v;
}

template <typename FT>
void launch_field_sums(const Layout &L, const uint8_t *role, ConstLayerPtrs<FT> f, double *block_partials, int max_blocks, double *result8,
                       cudaStream_t s, long long *launches)
{
	unsigned g = grid_for((long long)L.nx * L.ny, 1, max_blocks);
	k_field_sums<FT><<<g, 256, 0, s>>>(L, role, f, block_partials);
	k_field_sums_final<<<1, 32, 0, s>>>(block_partials, (int)g, result8);
	if (launches) *launches += 2;
}

// ---- TimeLayer3D::FilterToArrays (reference TimeLayer3D.h:842-854): nearest-lower downsample ---------
template <typename FT>
__global__ void k_filter(const Layout L, ConstLayerPtrs<FT> layer, int ox, int oy, int oz, int oi0, int oi1, FT *vel, double *T)
{
	const long long total = (long long)(oi1 - oi0) * oy * oz;
	for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
		const int k = (int)(t % oz);
		const int j = (int)((t / oz) % oy);
		const int i = (int)(t / ((long long)oz * oy)) + oi0;
		const int x = (int)((long long)i * L.gx / ox), y = (int)((long long)j * L.ny / oy), z = (int)((long long)k * L.nz / oz);
		const long long id = L.idx(x - L.x0, y, z);
		const long long ind = ((long long)i * oy + j) * oz + k;
		vel[3 * ind + 0] = layer.f[0][id];
		vel[3 * ind + 1] = layer.f[1][id];
		vel[3 * ind + 2] = layer.f[2][id];
		T[ind] = (double)layer.f[3][id];
	}
}

template <typename FT>
void launch_filter(const Layout &L, ConstLayerPtrs<FT> layer, int ox, int oy, int oz, int oi0, int oi1, FT *vel, double *T,
                   cudaStream_t s, long long *launches)
{
	if (oi1 <= oi0) return;
	k_filter<FT><<<grid_for((long long)(oi1 - oi0) * oy * oz, 256), 256, 0, s>>>(L, layer, ox, oy, oz, oi0, oi1, vel, T);
	if (launches) (*launches)++;
}

#define CMC_INST(FT) \
	template void launch_copy_full<FT>(const Layout &, ConstLayerPtrs<FT>, LayerPtrs<FT>, cudaStream_t, long long *); \
	template void launch_copy_masked<FT>(const Layout &, const uint8_t *, unsigned, ConstLayerPtrs<FT>, LayerPtrs<FT>, cudaStream_t, long long *); \
	template void launch_merge<FT>(const Layout &, const uint8_t *, ConstLayerPtrs<FT>, LayerPtrs<FT>, cudaStream_t, long long *); \
	template void launch_merge_to<FT>(const Layout &, const uint8_t *, ConstLayerPtrs<FT>, ConstLayerPtrs<FT>, LayerPtrs<FT>, cudaStream_t, long long *); \
	template void launch_update_boundaries<FT>(const Layout &, const uint8_t *, ConstLayerPtrs<FT>, LayerPtrs<FT>, cudaStream_t, long long *); \
	template void launch_clear_out<FT>(const Layout &, const uint8_t *, LayerPtrs<FT>, FT, cudaStream_t, long long *); \
	template void launch_fill<FT>(FT *, long long, FT, cudaStream_t, long long *); \
	template void launch_div_error<FT>(const Layout &, const uint8_t *, const FT *, const FT *, const FT *, FT, FT, FT, double *, int, double *, cudaStream_t, long long *); \
	template void launch_filter<FT>(const Layout &, ConstLayerPtrs<FT>, int, int, int, int, int, FT *, double *, cudaStream_t, long long *); \
	template void launch_scatter_dense<FT>(const Layout &, ConstLayerPtrs<FT>, LayerPtrs<FT>, cudaStream_t, long long *); \
	template void launch_field_sums<FT>(const Layout &, const uint8_t *, ConstLayerPtrs<FT>, double *, int, double *, cudaStream_t, long long *);
CMC_INST(float)
CMC_INST(double)

} // namespace cmc
