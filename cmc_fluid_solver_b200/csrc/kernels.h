// kernels.h - launchers implemented by the kernel translation units.
#pragma once
#include "common.cuh"

namespace cmc {

// ---- kernels_exact.cu (-fmad=false) --------------------------------------------------------------
template <typename FT>
void launch_exact_sweep(int dir, const SweepArgs<FT> &A, cudaStream_t s, long long *launches);
template <typename FT>
void launch_exact_x_pass(bool forward, const SweepArgs<FT> &A, cudaStream_t s, long long *launches);
template <typename FT>
void launch_thomas_batch(int nsys, int n, FT *a, FT *b, FT *c, FT *d, FT *x, cudaStream_t s);

// ---- kernels_fast.cu -----------------------------------------------------------------------------
// Returns false when the fast path does not support this line length (caller falls back to exact sweeps
// + merge kernel - still on the GPU; there is no CPU path).
bool fast_sweep_supported(const Layout &L, int dir);
template <typename FT>
bool launch_fast_sweep(int dir, const SweepArgs<FT> &A, cudaStream_t s, long long *launches);
// kernels_ring.cu: the same sweep as persistent CTAs fed by a cp.async shared-memory ring (MODE 0 lines only)
template <typename FT>
bool launch_ring_sweep(int dir, const SweepArgs<FT> &A, cudaStream_t s, long long *launches);
bool ring_sweep_supported(const Layout &L, int dir);
// kernels_tma.cu: x / y sweeps as persistent CTAs fed by bulk-tensor copies (TMA); false = not applicable, use the above
template <typename FT>
bool launch_tma_sweep(int dir, const SweepArgs<FT> &A, cudaStream_t s, long long *launches);
// fused slab-coupled x-sweep (one pass): lines per tile of the slab shape, 0 = not supported; the launch
int tma_xs_lines(const Layout &L);
template <typename FT>
bool launch_tma_xs(const SweepArgs<FT> &A, cudaStream_t s, long long *launches);
bool tma_sweep_supported(const Layout &L, int dir);
// partitioned x-sweep (slab-decomposed grid): spike pass, interface solve, coupled sweep
template <typename FT>
bool launch_x_spike(const SweepArgs<FT> &A, cudaStream_t s, long long *launches);
template <typename FT>
bool launch_x_coupled(const SweepArgs<FT> &A, cudaStream_t s, long long *launches);
template <typename FT>
// bnd_to[r]: the [8][lpo] block of rank r's interface table that this owner fills
void launch_x_interface(int P, int lpo, int nlines, const FT *coef, FT *const *bnd_to, cudaStream_t s, long long *launches);
template <typename FT>
bool launch_pcr_batch(int nsys, int n, const FT *a, const FT *b, const FT *c, const FT *d, FT *x, cudaStream_t s);

// ---- kernels_util.cu -----------------------------------------------------------------------------
// line descriptors (Grid3D::GenerateListSegments, reference Grid3D.cpp:47-127)
void launch_build_roles(int dir, const Layout &G, const uint8_t *ncode_global, const Layout &L, uint8_t *role,
                        unsigned long long *seg_count, cudaStream_t s, long long *launches);
// x-direction descriptors from a window of the node codes (no NODE_IN on the x-faces of the grid): purely local rule
void launch_build_roles_x_local(const Layout &G, const uint8_t *ncode_by_global_plane, const Layout &L, uint8_t *role,
                                unsigned long long *seg_count, cudaStream_t s, long long *launches);
void launch_count_in_plane(const Layout &G, const uint8_t *ncode_by_global_plane, int gplane_index, unsigned long long *out, cudaStream_t s);
// adds the type bits (R_IN / R_BV / R_OUT / R_VFREE / R_TFREE) to all three role arrays
void launch_role_type_bits(const Layout &G, const uint8_t *ncode_global, const Layout &L,
                           uint8_t *rx, uint8_t *ry, uint8_t *rz, cudaStream_t s, long long *launches);

template <typename FT>
void launch_copy_full(const Layout &L, ConstLayerPtrs<FT> src, LayerPtrs<FT> dst, cudaStream_t s, long long *launches);
// dst <- src where role has `mask` bits (CopyFieldTo, reference TimeLayer3D.h:394-413)
template <typename FT>
void launch_copy_masked(const Layout &L, const uint8_t *role, unsigned mask, ConstLayerPtrs<FT> src, LayerPtrs<FT> dst,
                        cudaStream_t s, long long *launches);
// dest = (dest + src) / 2 on NODE_IN (MergeFieldTo, reference TimeLayer3D.h:415-436)
template <typename FT>
void launch_merge(const Layout &L, const uint8_t *role, ConstLayerPtrs<FT> src, LayerPtrs<FT> dst, cudaStream_t s, long long *launches);
// out = (a + b) / 2 on NODE_IN, out = a elsewhere (merge into the other temp buffer)
template <typename FT>
void launch_merge_to(const Layout &L, const uint8_t *role, ConstLayerPtrs<FT> tmp, ConstLayerPtrs<FT> nxt, LayerPtrs<FT> out,
                     cudaStream_t s, long long *launches);
// cur <- Node.v / Node.T on BOUND and VALVE cells (CopyFromGrid, reference TimeLayer3D.h:926-944);
// optionally the same values into `also` (next <- cur on BOUND/VALVE, AdiSolver3D.cpp:310-311)
template <typename FT>
void launch_update_boundaries(const Layout &L, const uint8_t *role, ConstLayerPtrs<FT> nodev, LayerPtrs<FT> cur,
                              cudaStream_t s, long long *launches);
// fills all cells of `dst` with value (guard planes included)
template <typename FT>
void launch_fill(FT *dst, long long n, FT value, cudaStream_t s, long long *launches);
// TimeLayer3D::EvalDivError (reference TimeLayer3D.h:595-641): partial[0] = sum |div|, partial[1] = count
template <typename FT>
void launch_div_error(const Layout &L, const uint8_t *role, const FT *U, const FT *V, const FT *W,
                      FT dx, FT dy, FT dz, double *block_partials, int max_blocks, double *result2,
                      cudaStream_t s, long long *launches);
// dense [nx][ny][nz] staging arrays -> the padded / y-blocked layer
template <typename FT>
void launch_scatter_dense(const Layout &L, ConstLayerPtrs<FT> src, LayerPtrs<FT> dst, cudaStream_t s, long long *launches);
// per-field sums and sums of squares over the non-OUT cells of the slab (checksums; block_partials: 8 per block)
template <typename FT>
void launch_field_sums(const Layout &L, const uint8_t *role, ConstLayerPtrs<FT> f, double *block_partials, int max_blocks, double *result8,
                       cudaStream_t s, long long *launches);
// Solver3D::GetLayer: OUT cells of `layer` <- 99999 (Clear, reference TimeLayer3D.h:974-998) ...
template <typename FT>
void launch_clear_out(const Layout &L, const uint8_t *role, LayerPtrs<FT> layer, FT value, cudaStream_t s, long long *launches);
// ... then nearest-lower downsample (FilterToArrays, reference TimeLayer3D.h:842-854) of the rows
// [oi0, oi1) of the output grid that fall into this slab.
template <typename FT>
void launch_filter(const Layout &L, ConstLayerPtrs<FT> layer, int ox, int oy, int oz, int oi0, int oi1,
                   FT *vel_xyz, double *T, cudaStream_t s, long long *launches);

} // namespace cmc
