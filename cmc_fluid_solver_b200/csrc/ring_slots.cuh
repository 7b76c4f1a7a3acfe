// ring_slots.cuh - cp.async helpers and the shared-memory slot layouts of the staged sweeps (kernels_ring.cu: all
// inputs of a tile ride through slots; kernels_fast.cu: the inputs of the second half of an x / y tile do).
#pragma once
#include "fast_core.cuh"

namespace cmc {

__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
	const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src)
{
	const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
	const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- geometry of one tile -------------------------------------------------------------------------------------
template <int DIR, int NL>
struct Tile {
	long long tbase;     // element index of (row 0, first line of the tile)
	long long stride;    // along the line
	int n;               // rows of a line
	int lines;           // lines of the tile that exist (1..NL)
	int pi;              // y / z lines: the x-plane of the tile
	int j0;              // z lines: j-row of the first line (the 8 lines of a tile never straddle a y-block)
	__device__ __forceinline__ void set(const Layout &L, int tile)
	{
		if (DIR == 0) {
			const int kt = (L.nz + NL - 1) / NL, j = tile / kt, k0 = (tile - j * kt) * NL;
			tbase = L.idx(0, j, k0); stride = L.plane; n = L.nx; lines = min(NL, L.nz - k0); pi = 0; j0 = j;
		} else if (DIR == 1) {
			const int kt = (L.nz + NL - 1) / NL, i = tile / kt, k0 = (tile - i * kt) * NL;
			tbase = L.idx(i, 0, k0); stride = L.nzp; n = L.ny; lines = min(NL, L.nz - k0); pi = i; j0 = 0;
		} else {
			const int jt = (L.ny + NL - 1) / NL, i = tile / jt, j0 = (tile - i * jt) * NL;
			tbase = L.idx(i, j0, 0); stride = 1; n = L.nz; lines = min(NL, L.ny - j0); pi = i; this->j0 = j0;
		}
	}
};

// ---- slot layouts -----------------------------------------------------------------------------------------------
// A slot holds one field of one tile: GP chunks x 8 rows x NL lines.
//   X, Y (lines strided in memory, a row of the tile = NL contiguous elements): row-major rows of NL elements, the
//     8 rows of chunk g rotated by g (physical row 8g + ((i + g) & 7)) so that the chunks read by one warp fall
//     into different banks;
//   Z (lines contiguous): chunk-major, chunk (l, g) = 8 contiguous elements = PC 16-byte pieces, the pieces of a
//     chunk rotated by g / (8 / PC) for the same reason.
template <typename FT, int DIR, int GP, int NL>
struct Slot {
	static constexpr int STR = GP * NL;
	static constexpr int EPP = 16 / (int)sizeof(FT);    // elements per 16-byte piece
	static constexpr int PC = M / EPP;                  // pieces per chunk
	static constexpr int PR = NL / EPP;                 // pieces per tile row (X, Y)
	static constexpr int ELEMS = STR * M;
	static_assert(NL == M, "tile rows and chunks are both 8 wide");

	// element index of (line l, row r)
	static __device__ __forceinline__ int at(int l, int r)
	{
		const int g = r >> 3, i = r & 7;
		if (DIR == 2) return ((l * GP + g) * PC + (((i / EPP) + g / (8 / PC)) & (PC - 1))) * EPP + (i % EPP);
		return ((g << 3) + ((i + g) & 7)) * NL + l;
	}

	// this thread's chunk (line l, chunk g) -> registers
	static __device__ __forceinline__ void read_chunk(const FT *slot, int l, int g, FT (&o)[M])
	{
		if (DIR == 2) {
			typedef typename Vec16<FT>::type V;
			const V *q = reinterpret_cast<const V *>(slot) + (l * GP + g) * PC;
			const int rot = g / (8 / PC);
#pragma unroll
			for (int v = 0; v < PC; v++) {
				const V t = q[(v + rot) & (PC - 1)];
				const FT *e = reinterpret_cast<const FT *>(&t);
#pragma unroll
				for (int k = 0; k < EPP; k++) o[v * EPP + k] = e[k];
			}
		} else {
#pragma unroll
			for (int i = 0; i < M; i++) o[i] = slot[((g << 3) + ((i + g) & 7)) * NL + l];
		}
	}

	// all threads: copy one field of the tile into `slot` (rows / lines outside the grid are clamped into it, like the
	// clamped offsets of the direct loads: every element of the slot holds valid data)
	static __device__ __forceinline__ void issue(FT *slot, const FT *__restrict__ field, const Tile<DIR, NL> &T, const Layout &L, int t)
	{
#pragma unroll
		for (int c = 0; c < PC; c++) {
			const int p = t + c * STR;
			if (DIR == 2) {
				const int l = p / (GP * PC), w = p - l * (GP * PC), g = w / PC, v = w - g * PC;
				const FT *src = field + T.tbase + (long long)min(l, T.lines - 1) * L.nzp + min(g << 3, L.nzp - M) + v * EPP;
				cp_async16(slot + ((l * GP + g) * PC + ((v + g / (8 / PC)) & (PC - 1))) * EPP, src);
			} else {
				const int r = p / PR, qt = p - r * PR, g = r >> 3, i = r & 7;
				const FT *src = field + T.tbase + (long long)min(r, T.n - 1) * T.stride + qt * EPP;
				cp_async16(slot + ((g << 3) + ((i + g) & 7)) * NL + qt * EPP, src);
			}
		}
	}
};

template <typename FT> __device__ __forceinline__ void cp_async_elem(FT *dst, const FT *src);
template <> __device__ __forceinline__ void cp_async_elem<double>(double *dst, const double *src) { cp_async8(dst, src); }
template <> __device__ __forceinline__ void cp_async_elem<float>(float *dst, const float *src) { cp_async4(dst, src); }

} // namespace cmc
