// common.cuh - device data model shared by all kernels of the ADI hot path.
//
// Time-layer storage (replaces ScalarField3D / TimeLayer3D, reference
// src/FluidSolver3D/TimeLayer3D.h:249-260, 536-552): structure of arrays, one device buffer per
// field, lines along z padded to a multiple of 16 elements (128 B in fp64) so every z-line
// starts on a 128-byte boundary, plus one guard plane before and after the slab.  The guard
// planes double as the halo planes of the x-slab decomposition (reference haloSize = dimy*dimz,
// AdiSolver3D.cpp:251-258).
//
//   idx(i, j, k) = (i + 1) * plane + j * nzp + k,   i in [-1, nx],  plane = ny * nzp
// (optionally blocked along y, see Layout)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cmc {

constexpr int MAX_SLABS = 16;   // slabs (GPUs) of one grid

struct Layout {
	int nx, ny, nz;       // cells of this slab (nx = local x-planes), reference dimx/dimy/dimz
	int nzp;              // padded z-line length
	int x0;               // global index of local plane 0
	int gx;               // global dimx
	// y-blocking: the grid is stored as blocks of 2^jbs consecutive j-rows, each block holding ALL x-planes of its rows
	// ([j / jb][i][j % jb][k]).  One block (jbs = 30, the default) is the plain [i][j][k] order.  Blocking shortens the
	// distance between consecutive rows of an x-line from ny * nzp to jb * nzp elements (2 MB -> 128 KB at 512^3 fp64,
	// jb = 32), which is what the x-sweep's address translation needs (profiles/r01_variants.md), at the price of
	// y-lines that jump from block to block every jb rows.
	int jbs, jbm;         // log2(jb), jb - 1
	int nblk;             // number of blocks
	long long plane;      // distance of consecutive x-planes: rows per block * nzp
	long long bstride;    // distance of consecutive blocks: (nx + 2) * plane
	long long total;      // nblk * bstride
	__host__ __device__ __forceinline__ long long idx(int i, int j, int k) const
	{
		return (long long)(j >> jbs) * bstride + (long long)(i + 1) * plane + (long long)(j & jbm) * nzp + k;
	}
	// element distance to the next / previous j-row (valid memory for every j in [0, ny): rows outside the grid are
	// only ever touched for values that are not used)
	__host__ __device__ __forceinline__ long long jup(int j) const
	{
		return (((j + 1) & jbm) == 0 && j + 1 < ny) ? bstride - (long long)jbm * nzp : (long long)nzp;
	}
	__host__ __device__ __forceinline__ long long jdn(int j) const
	{
		return ((j & jbm) == 0 && j > 0) ? bstride - (long long)jbm * nzp : (long long)nzp;
	}
	// geometry of a slab of `lnx` planes of a grid ny x nz (block height 2^jbs_, or one block when jbs_ >= 30).  alloc_nx
	// (>= lnx): planes a y-block has room for - the slabs of one grid may hold different numbers of planes but share the
	// block stride of the largest one, so that an element has the same offset in every slab's buffers (stores into a
	// neighbouring slab's guard planes, SweepArgs::push_*, need nothing but the neighbour's plane count)
	__host__ void shape(int lnx, int ny_, int nz_, int nzp_, int jbs_, int alloc_nx = 0)
	{
		nx = lnx; ny = ny_; nz = nz_; nzp = nzp_;
		if (jbs_ >= 30 || (1 << jbs_) >= ny_) { jbs = 30; jbm = (1 << 30) - 1; nblk = 1; plane = (long long)ny_ * nzp_; }
		else { jbs = jbs_; jbm = (1 << jbs_) - 1; nblk = (ny_ + jbm) >> jbs_; plane = (long long)(1 << jbs_) * nzp_; }
		bstride = (long long)((alloc_nx > lnx ? alloc_nx : lnx) + 2) * plane;
		total = (long long)nblk * bstride;
	}
};

// Per-cell, per-direction line descriptor byte (replaces Segment3D + NodesBoundary3D,
// reference src/FluidSolver3D/Grid3D.h:63-93: 40 B + 2 Nodes per segment).  Built once on the
// device from the node types by k_build_roles (GenerateListSegments, Grid3D.cpp:47-127).
enum : unsigned {
	R_INT   = 1u,    // interior row of a segment (cell is NODE_IN inside a terminated run)
	R_START = 2u,    // first cell of a segment  -> ApplyBC0 row
	R_END   = 4u,    // last cell of a segment   -> ApplyBC1 row
	R_VFREE = 8u,    // node.bc_vel  == BC_FREE (selects the boundary row of u, v, w)
	R_TFREE = 16u,   // node.bc_temp == BC_FREE (selects the boundary row of T)
	R_IN    = 32u,   // node.type == NODE_IN    (merge mask)
	R_BV    = 64u,   // node.type is NODE_BOUND or NODE_VALVE (boundary refresh / copy mask)
	R_PRE   = 128u,  // interior row whose NEXT cell is shared by two segments (ends this one, starts the next): the
	                 // shared cell's ApplyBC1 row is folded into this row and R_VFREE / R_TFREE describe THAT cell
	                 // (NODE_OUT = neither R_IN nor R_BV: GetLayer writes 99999 there)
	R_SEG   = R_INT | R_START | R_END
};

template <typename FT>
struct SweepArgs {
	Layout L;
	FT dt;
	FT h[3];                 // dx, dy, dz as FTYPE (TimeLayer3D ctor casts, AdiSolver3D.cpp:256)
	FT v_T, v_vis, t_vis, t_phi;
	const uint8_t *role;     // descriptor bytes of the sweep direction
	const FT *cur[4];        // u, v, w, T of the sweep's "cur" layer
	const FT *temp[4];       // linearisation layer (read)
	FT *next[4];             // sweep output layer
	FT *temp_out[4];         // merged linearisation layer (fast mode, double-buffered)
	const FT *nodev[4];      // Node.v.x, v.y, v.z, Node.T (boundary row values)
	FT *cv, *cT;             // exact mode: Thomas c' scratch (velocity matrix, temperature matrix)
	// partitioned x-sweep of a slab-decomposed grid (kernels_fast.cu MODE 1 / 2, dist.h)
	FT *xcoef;               // MODE 1 output: [owner][16][lpo] coefficients of the slab's first / last row
	const FT *xbnd;          // MODE 2 input:  [owner][8][lpo] solutions of the neighbours' adjacent rows
	int lpo;                 // lines per owner rank of the interface solve
	int extra_merge;         // fast mode: apply the relaxation twice (folds the post-X MergeLayerTo, AdiSolver3D.cpp:354)
	int *tile_counter;       // persistent-CTA kernels (kernels_tma.cu): next tile to hand out; zeroed before every launch
	int tma_shape;           // kernels_tma.cu: 0 = automatic tile shape, else lines per tile + 256 * CTAs per tile
	// fused slab-coupled x-sweep (kernels_tma.cu XS): every slab holds a table [slab][tile][16 coefficients][words][lines per
	// tile] of self-validating 8-byte words (payload + epoch); a slab's kernel stores its part into all other slabs' tables
	int xs_P, xs_me;         // number of slabs, this slab's index
	int xs_epoch;            // the epoch of this sweep (grows from sweep to sweep; the tables are never reset)
	int xs_share;            // slabs that share this slab's device (1 unless a test runs several slabs on one GPU)
	unsigned long long *xs_tab;               // this slab's table
	unsigned long long *xs_tab_to[MAX_SLABS]; // every slab's table (peer memory)
	// ---- slab-decomposed runs: exchanges fused into the sweeps as stores into the other slabs' buffers (peer memory
	// over NVLink when the slabs live on different GPUs, see dist.h) --------------------------------------------------
	// boundary x-planes of the sweep's outputs -> the x-neighbours' guard planes.  Each pointer addresses the target
	// PLANE (element (j, k) at j * nzp + k); null = no neighbour on that side / nothing to push.
	FT *push_lo[4], *push_hi[4];     // temp_out: plane 0 -> lower neighbour's plane nx, plane nx-1 -> upper neighbour's plane -1
	FT *pushn_lo[4], *pushn_hi[4];   // the same for `next` (coupled x-sweep only: its output is the next sweep's / step's input)
	// MODE 1: where the 16 coefficients of a line go: xcoef_to[owner] addresses the [16][lpo] block this slab owns in
	// the owner's coefficient table
	FT *xcoef_to[MAX_SLABS];
};

// elementwise helpers -----------------------------------------------------------------------
template <typename FT>
struct LayerPtrs { FT *f[4]; };
template <typename FT>
struct ConstLayerPtrs { const FT *f[4]; };

} // namespace cmc
