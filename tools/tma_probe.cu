// tma_probe.cu - micro-benchmark: how fast can persistent CTAs pull x / y sweep tiles out of HBM with bulk-tensor copies?
// Same tensor maps and box shapes as kernels_tma.cu (5-D boxes that gather 64-byte row segments), no arithmetic: every
// CTA keeps NSLOT copies in flight round-robin over 8 "fields" and hands tiles out with an atomic counter.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tma_probe tools/tma_probe.cu
//   build/tma_probe            # prints GB/s for (x | y) x (64-byte | 128-byte rows) x (slots in flight)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Maps { CUtensorMap m[8]; };

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int NSLOT>
__global__ void __launch_bounds__(128, 1) k_probe(const __grid_constant__ Maps TM, int ntiles, int ktiles, int dir, int kw, int *counter, int slot_bytes,
                                                 int jbm, int jbs, long long *sink)
{
	extern __shared__ __align__(128) unsigned char smem[];
	unsigned long long *full = reinterpret_cast<unsigned long long *>(smem + (size_t)NSLOT * slot_bytes);
	__shared__ int next_box;
	const int t = threadIdx.x;
	if (t == 0) {
		for (int s = 0; s < NSLOT; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(full + s)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();
	if (t != 0) return;          // one thread drives the copies: this measures the copy engine + memory system only
	unsigned ph = 0;
	int issued = 0, waited = 0;
	long long acc = 0;
	int tile = blockIdx.x;
	while (tile < ntiles) {
		const int a = tile / ktiles, k0 = (tile - a * ktiles) * kw;
		for (int f = 0; f < 11; f++) {                  // 11 copies per tile like the sweep kernel (8 fields + 3 repeats)
			const int s = issued % NSLOT;
			if (issued >= NSLOT) {                       // slot busy: wait for its previous copy
				const unsigned par = (ph >> s) & 1u;
				unsigned ok = 0;
				while (!ok)
					asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x100000; selp.u32 %0, 1, 0, P1; }"
					             : "=r"(ok) : "r"(smem_u32(full + s)), "r"(par) : "memory");
				ph ^= 1u << s;
				acc += *reinterpret_cast<const long long *>(smem + (size_t)s * slot_bytes);
				waited++;
			}
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full + s)), "r"(slot_bytes) : "memory");
			int c3, c4;
			if (dir == 0) { c3 = a & jbm; c4 = a >> jbs; } else { c3 = 0; c4 = a + 1; }
			asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
			             ::"r"(smem_u32(smem + (size_t)s * slot_bytes)), "l"(reinterpret_cast<unsigned long long>(&TM.m[f % 8])), "r"(smem_u32(full + s)),
			             "r"(k0), "r"(0), "r"(0), "r"(c3), "r"(c4) : "memory");
			issued++;
		}
		tile = (int)gridDim.x + atomicAdd(counter, 1);
	}
	while (waited < issued) {
		const int s = waited % NSLOT;
		const unsigned par = (ph >> s) & 1u;
		unsigned ok = 0;
		while (!ok)
			asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x100000; selp.u32 %0, 1, 0, P1; }"
			             : "=r"(ok) : "r"(smem_u32(full + s)), "r"(par) : "memory");
		ph ^= 1u << s;
		waited++;
	}
	if (acc == 0x7fffffffffffffffll) *sink = acc;
	(void)next_box;
}


// The same tiles pulled by 1-D bulk copies (cp.async.bulk, UBLKCP): every thread issues the 64-byte row segments of its
// own rows, one instruction per row.  Does the copy engine take many small independent requests faster than the rows of
// one tensor box?
template <int NSLOT>
__global__ void __launch_bounds__(128, 1) k_probe_bulk(const double *const *fields, int ntiles, int ktiles, int *counter, int slot_bytes,
                                                      long long plane, long long bstride, int nzp, int jbm, int jbs, int nrows, int row_bytes, long long *sink)
{
	extern __shared__ __align__(128) unsigned char smem[];
	unsigned long long *full = reinterpret_cast<unsigned long long *>(smem + (size_t)NSLOT * slot_bytes);
	__shared__ int next_tile;
	const int t = threadIdx.x;
	if (t == 0) {
		for (int s = 0; s < NSLOT; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(full + s)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();
	unsigned ph = 0;
	int issued = 0;
	int tile = blockIdx.x;
	while (tile < ntiles) {
		const int jj = tile / ktiles, k0 = (tile - jj * ktiles) * (row_bytes / 8);
		const int j = jj & 511, half = jj >> 9;          // (128-byte rows: two half-height tiles per line position)
		const long long tbase = (long long)(j >> jbs) * bstride + (1 + (long long)half * nrows) * plane + (long long)(j & jbm) * nzp + k0;
		for (int f = 0; f < 11; f++) {
			const int s = issued % NSLOT;
			if (t == 0) {
				if (issued >= NSLOT) {
					const unsigned par = (ph >> s) & 1u;
					unsigned ok = 0;
					while (!ok)
						asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x100000; selp.u32 %0, 1, 0, P1; }"
						             : "=r"(ok) : "r"(smem_u32(full + s)), "r"(par) : "memory");
					ph ^= 1u << s;
				}
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full + s)), "r"(slot_bytes) : "memory");
			}
			__syncthreads();
			const double *src = fields[f % 8] + tbase;
			for (int r = t; r < nrows; r += 128)
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				             ::"r"(smem_u32(smem + (size_t)s * slot_bytes + (size_t)r * row_bytes)), "l"(src + (long long)r * plane), "r"(row_bytes),
				             "r"(smem_u32(full + s)) : "memory");
			issued++;
		}
		if (t == 0) next_tile = (int)gridDim.x + atomicAdd(counter, 1);
		__syncthreads();
		tile = next_tile;
		__syncthreads();
	}
	if (t == 0) {
		for (int w = issued > NSLOT ? issued - NSLOT : 0; w < issued; w++) {
			const int s = w % NSLOT;
			const unsigned par = (ph >> s) & 1u;
			unsigned ok = 0;
			while (!ok)
				asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x100000; selp.u32 %0, 1, 0, P1; }"
				             : "=r"(ok) : "r"(smem_u32(full + s)), "r"(par) : "memory");
			ph ^= 1u << s;
		}
		if (issued == -1) *sink = 1;
	}
}

// Copy-through probe: the sweep's memory traffic without its arithmetic.  Per tile 11 box loads (8 distinct fields + 3
// repeats, like the sweep kernel) and 8 box stores (cp.async.bulk.tensor shared -> global) into 8 other fields, tiles
// handed out by the same atomic counter.  What the memory system delivers for THIS access pattern (64- or 128-byte row
// segments a plane apart, reads and writes mixed) is the practical ceiling of the x / y sweep kernels.
struct Maps16 { CUtensorMap in[8], out[8]; };
template <int NSLOT>
__global__ void __launch_bounds__(128, 1) k_probe_copy(const __grid_constant__ Maps16 TM, int ntiles, int ktiles, int dir, int kw, int *counter, int slot_bytes,
                                                      int jbm, int jbs, int halves, int c1h, int c2h)
{
	extern __shared__ __align__(128) unsigned char smem[];
	unsigned long long *full = reinterpret_cast<unsigned long long *>(smem + (size_t)NSLOT * slot_bytes);
	const int t = threadIdx.x;
	if (t == 0) {
		for (int s = 0; s < NSLOT; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(full + s)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();
	if (t != 0) return;
	unsigned ph = 0;
	int tile = blockIdx.x;
	// software pipeline over (tile, field): loads run NSLOT - 1 ahead of the stores
	struct Item { int c0, c1, c2, c3, c4, f; };
	Item ring[NSLOT];
	int head = 0, count = 0;     // items in flight (loaded, not yet stored)
	auto drain_one = [&]() {
		const int s = head % NSLOT;
		const unsigned par = (ph >> s) & 1u;
		unsigned ok = 0;
		while (!ok)
			asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x100000; selp.u32 %0, 1, 0, P1; }"
			             : "=r"(ok) : "r"(smem_u32(full + s)), "r"(par) : "memory");
		ph ^= 1u << s;
		const Item it = ring[s];
		if (it.f < 8) {           // the 3 repeats are read-only
			asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
			             ::"l"(reinterpret_cast<unsigned long long>(&TM.out[it.f])), "r"(smem_u32(smem + (size_t)s * slot_bytes)),
			             "r"(it.c0), "r"(it.c1), "r"(it.c2), "r"(it.c3), "r"(it.c4) : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		}
		head++; count--;
	};
	int issued = 0;
	while (tile < ntiles) {
		// (128-byte rows: a box holds half of every line; two boxes - `halves` - per line position)
		const int pos = tile / halves, half = tile - pos * halves;
		const int a = pos / ktiles, k0 = (pos - a * ktiles) * kw;
		const int c1 = half * c1h, c2 = half * c2h;
		int c3, c4;
		if (dir == 0) { c3 = a & jbm; c4 = a >> jbs; } else { c3 = 0; c4 = a + 1; }
		for (int f = 0; f < 11; f++) {
			if (count == NSLOT - 1) drain_one();
			const int s = issued % NSLOT;
			// the slot's previous store must have finished reading shared memory
			asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NSLOT - 2) : "memory");
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full + s)), "r"(slot_bytes) : "memory");
			asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
			             ::"r"(smem_u32(smem + (size_t)s * slot_bytes)), "l"(reinterpret_cast<unsigned long long>(&TM.in[f % 8])), "r"(smem_u32(full + s)),
			             "r"(k0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
			ring[s] = Item{k0, c1, c2, c3, c4, f};
			issued++; count++;
		}
		tile = (int)gridDim.x + atomicAdd(counter, 1);
	}
	while (count > 0) drain_one();
	asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
	const int nx = 512, ny = 512, nz = 512, nzp = 512, jb = 64, jbs = 6, jbm = 63;
	const long long plane = (long long)jb * nzp, bstride = (long long)(nx + 2) * plane, total = (ny / jb) * bstride;
	void *p = nullptr;
	cudaDriverEntryPointQueryResult q;
	CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
	EncodeTiledFn enc = (EncodeTiledFn)p;
	double *fields[8];
	for (int f = 0; f < 8; f++) { CK(cudaMalloc(&fields[f], sizeof(double) * total)); CK(cudaMemset(fields[f], 0, sizeof(double) * total)); }
	int *counter; long long *sink;
	CK(cudaMalloc(&counter, 4)); CK(cudaMalloc(&sink, 8));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int dir = 0; dir < 2; dir++)
		for (int kw = 8; kw <= 16; kw *= 2) {               // 64-byte rows (the sweep tile) / 128-byte rows (half as many chunks per box)
			const int GP = 64 * 8 / kw;                     // chunks per box so that a box stays 32 KB
			Maps TM;
			for (int f = 0; f < 8; f++) {
				cuuint64_t dims[5], strides[4]; cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
				void *base;
				if (dir == 0) {
					dims[0] = nzp; dims[1] = nx / 8; dims[2] = 8; dims[3] = jb; dims[4] = ny / jb;
					strides[0] = 8 * plane * 8; strides[1] = plane * 8; strides[2] = nzp * 8; strides[3] = bstride * 8;
					box[0] = kw; box[1] = GP; box[2] = 8; box[3] = 1; box[4] = 1;
					base = fields[f] + plane;
				} else {
					dims[0] = nzp; dims[1] = jb / 8; dims[2] = ny / jb; dims[3] = 8; dims[4] = nx + 2;
					strides[0] = 8 * nzp * 8; strides[1] = bstride * 8; strides[2] = nzp * 8; strides[3] = plane * 8;
					box[0] = kw; box[1] = jb / 8; box[2] = GP / (jb / 8); box[3] = 8; box[4] = 1;
					base = fields[f];
				}
				CUresult r = enc(&TM.m[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
				                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
				if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
			}
			const int ktiles = nz / kw;
			// 64 chunks of the line per tile: a 128-byte-row tile is two boxes high; count bytes, not tiles
			const int ntiles = (dir == 0 ? ny : nx) * ktiles;
			const int slot_bytes = kw * GP * 8 * 8;
			auto run = [&](auto kern, int nslot) {
				const size_t smem = (size_t)nslot * slot_bytes + 64;
				CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
				float best = 1e9f;
				for (int rep = 0; rep < 3; rep++) {
					CK(cudaMemset(counter, 0, 4));
					cudaEventRecord(e0);
					kern<<<148, 128, smem>>>(TM, ntiles, ktiles, dir, kw, counter, slot_bytes, jbm, jbs, sink);
					cudaEventRecord(e1);
					CK(cudaDeviceSynchronize());
					float ms; cudaEventElapsedTime(&ms, e0, e1);
					if (ms < best) best = ms;
				}
				const double bytes = (double)ntiles * 11 * slot_bytes;
				printf("dir %c  row %3d B  box %d x %d chunks  %d slots in flight: %.3f ms  %.0f GB/s delivered to shared memory (11 copies per tile, 8 distinct fields)\n",
				       "xy"[dir], kw * 8, kw, GP, nslot, best, bytes / best / 1e6);
			};
			run(k_probe<2>, 2); run(k_probe<4>, 4); run(k_probe<6>, 6);
		}
	// 1-D bulk copies, x tiles only: 512 rows of 64 / 128 bytes per field per tile
	{
		const double **dfields;
		CK(cudaMalloc(&dfields, 8 * sizeof(double *)));
		CK(cudaMemcpy(dfields, fields, 8 * sizeof(double *), cudaMemcpyHostToDevice));
		for (int kw = 8; kw <= 16; kw *= 2) {
			const int nrows = 512 * 8 / kw;                 // rows per slot so that a slot stays 32 KB
			const int slot_bytes = nrows * kw * 8, ktiles = nz / kw, ntiles = ny * ktiles * (kw / 8);
			auto run = [&](auto kern, int nslot) {
				const size_t smem = (size_t)nslot * slot_bytes + 64;
				CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
				float best = 1e9f;
				for (int rep = 0; rep < 3; rep++) {
					CK(cudaMemset(counter, 0, 4));
					cudaEventRecord(e0);
					kern<<<148, 128, smem>>>(dfields, ntiles, ktiles, counter, slot_bytes, plane, bstride, nzp, jbm, jbs, nrows, kw * 8, sink);
					cudaEventRecord(e1);
					CK(cudaDeviceSynchronize());
					float ms; cudaEventElapsedTime(&ms, e0, e1);
					if (ms < best) best = ms;
				}
				const double bytes = (double)ntiles * 11 * slot_bytes;
				printf("1-D bulk copies, x rows of %3d B, %d rows per slot, %d slots in flight: %.3f ms  %.0f GB/s delivered to shared memory\n",
				       kw * 8, nrows, nslot, best, bytes / best / 1e6);
			};
			run(k_probe_bulk<4>, 4); run(k_probe_bulk<6>, 6);
		}
	}
	// copy-through: loads + stores
	{
		double *outs[8];
		for (int f = 0; f < 8; f++) { CK(cudaMalloc(&outs[f], sizeof(double) * total)); CK(cudaMemset(outs[f], 0, sizeof(double) * total)); }
		for (int dir = 0; dir < 2; dir++)
			for (int kw = 8; kw <= 16; kw *= 2) {
				const int GP = 64 * 8 / kw;
				Maps16 TM;
				for (int f = 0; f < 16; f++) {
					double *fld = f < 8 ? fields[f] : outs[f - 8];
					cuuint64_t dims[5], strides[4]; cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
					void *base;
					if (dir == 0) {
						dims[0] = nzp; dims[1] = nx / 8; dims[2] = 8; dims[3] = jb; dims[4] = ny / jb;
						strides[0] = 8 * plane * 8; strides[1] = plane * 8; strides[2] = nzp * 8; strides[3] = bstride * 8;
						box[0] = kw; box[1] = GP; box[2] = 8; box[3] = 1; box[4] = 1;
						base = fld + plane;
					} else {
						dims[0] = nzp; dims[1] = jb / 8; dims[2] = ny / jb; dims[3] = 8; dims[4] = nx + 2;
						strides[0] = 8 * nzp * 8; strides[1] = bstride * 8; strides[2] = nzp * 8; strides[3] = plane * 8;
						box[0] = kw; box[1] = jb / 8; box[2] = GP / (jb / 8); box[3] = 8; box[4] = 1;
						base = fld;
					}
					CUresult r = enc(f < 8 ? &TM.in[f] : &TM.out[f - 8], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
					                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
					if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
				}
				const int halves = 64 / GP;                      // boxes per line position
				const int ktiles = nz / kw, ntiles = (dir == 0 ? ny : nx) * ktiles * halves, slot_bytes = kw * GP * 8 * 8;
				const int c1h = dir == 0 ? GP : 0, c2h = dir == 0 ? 0 : GP / (jb / 8);
				const size_t smem = (size_t)6 * slot_bytes + 64;
				CK(cudaFuncSetAttribute(k_probe_copy<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
				float best = 1e9f;
				for (int rep = 0; rep < 3; rep++) {
					CK(cudaMemset(counter, 0, 4));
					cudaEventRecord(e0);
					k_probe_copy<6><<<148, 128, smem>>>(TM, ntiles, ktiles, dir, kw, counter, slot_bytes, jbm, jbs, halves, c1h, c2h);
					cudaEventRecord(e1);
					CK(cudaDeviceSynchronize());
					float ms; cudaEventElapsedTime(&ms, e0, e1);
					if (ms < best) best = ms;
				}
				// algorithmic bytes of a sweep: 8 fields read once + 8 fields written (here every cell, the sweep skips masked ones)
				const double alg = 16.0 * nx * ny * nz * 8;
				printf("copy-through dir %c  row %3d B: %.3f ms per sweep-equivalent  %.0f GB/s algorithmic (8 fields in, 8 out; 11 loads + 8 stores per tile)\n",
				       "xy"[dir], kw * 8, best, alg / best / 1e6);
			}
	}
	return 0;
}
