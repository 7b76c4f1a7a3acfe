/*
 * cmc_adi.h - C ABI of the B200-native implicit (ADI) time-stepping path of cmc-fluid-solver.
 *
 * The reference has no FFI layer; its seam for this path is the C++ abstract class
 * FluidSolver3D::Solver3D (src/FluidSolver3D/Solver3D.h:24-49) plus AdiSolver3D::CreateSegments
 * (src/FluidSolver3D/AdiSolver3D.h:52-61).  Every entry point below replaces one of those
 * members; the citation on each declaration names it.  INTEGRATION.md shows the adapter
 * (class B200AdiSolver3D : public Solver3D) a maintainer adds on the reference side.
 *
 * Conventions: plain pointers and sizes only; every function returns an int status
 * (CMC_OK == 0); cmc_last_error() returns a thread-local message for the last failure.
 * Host arrays use the reference's dense layout  idx = (i*dimy + j)*dimz + k  (k fastest,
 * src/FluidSolver3D/TimeLayer3D.h:256-259).  `fp_bytes` (4 or 8) plays the role of the
 * reference's compile-time FTYPE (src/Common/Geometry.h:21): all `void*` field arrays are
 * float[] when fp_bytes == 4 and double[] when fp_bytes == 8.
 *
 * There is no CPU fallback: every call needs a CUDA device (sm_100a) and fails loudly otherwise.
 */
#ifndef CMC_ADI_H
#define CMC_ADI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMC_ADI_ABI_VERSION 1

/* status codes */
enum {
	CMC_OK = 0,
	CMC_ERR_DIVERGED = 1,     /* residual > 0.01: the reference prints "Error is too big!" and throws (AdiSolver3D.cpp:371-374) */
	CMC_ERR_INVALID = -1,     /* bad argument / call order */
	CMC_ERR_CUDA = -2,        /* CUDA runtime or driver error (message carries device + code, cf. GPUplan.cpp:173-193) */
	CMC_ERR_NO_DEVICE = -3,   /* no usable sm_100 device: there is no CPU fallback */
	CMC_ERR_UNSUPPORTED = -4,
	CMC_ERR_COMM = -5         /* NCCL error */
};

/* enum values mirror src/Common/Geometry.h:29-43 and are part of the contract */
enum { CMC_NODE_IN = 0, CMC_NODE_OUT = 1, CMC_NODE_BOUND = 2, CMC_NODE_VALVE = 3 };
enum { CMC_BC_NOSLIP = 0, CMC_BC_FREE = 1 };
enum { CMC_DIR_X = 0, CMC_DIR_Y = 1, CMC_DIR_Z = 2 };
/* src/FluidSolver3D/AdiSolver3D.h:38 */
enum { CMC_VAR_U = 0, CMC_VAR_V = 1, CMC_VAR_W = 2, CMC_VAR_T = 3 };
/* the four time layers AdiSolver3D keeps (AdiSolver3D.cpp:254-258) */
enum { CMC_LAYER_CUR = 0, CMC_LAYER_HALF = 1, CMC_LAYER_NEXT = 2, CMC_LAYER_TEMP = 3 };

/* solve modes (cmc_adi3d_set_option "mode") */
enum {
	CMC_MODE_FAST = 0,   /* partitioned (SPIKE/PCR) line solves, FMA contraction; matches the reference within tolerance */
	CMC_MODE_EXACT = 1   /* sequential Thomas in the reference's operation order, no FMA; bit-identical with the reference CPU solver */
};

#define CMC_MISSING_VALUE 99999.0f   /* src/Common/Geometry.h:25 */
#define CMC_ERR_THRESHOLD 0.01       /* src/FluidSolver3D/AdiSolver3D.h:32 */

/* Grid3D public members read by the solver (src/FluidSolver3D/Grid3D.h:97-100) */
typedef struct cmc_grid_desc {
	int32_t dimx, dimy, dimz;
	double dx, dy, dz;
} cmc_grid_desc;

/* Common::FluidParams (src/Common/Geometry.h:538-562) */
typedef struct cmc_fluid_params {
	double v_T, v_vis, t_vis, t_phi;
} cmc_fluid_params;

typedef struct cmc_adi3d cmc_adi3d;   /* opaque solver handle (one per GPU / rank) */

const char *cmc_last_error(void);
int cmc_abi_version(void);
/* number of visible CUDA devices, or CMC_ERR_NO_DEVICE */
int cmc_device_count(void);

/* ---- lifetime: replaces `new AdiSolver3D` + Solver3D::Init (Solver3D.h:27, AdiSolver3D.cpp:166-268) ---- */
int cmc_adi3d_create(const cmc_grid_desc *grid, const cmc_fluid_params *params,
                     int fp_bytes, int device, cmc_adi3d **out);
/* one slab of a multi-GPU run (one process per GPU).  The grid is split into contiguous slabs
 * of whole x-planes (the slowest axis; GPUplan::splitEven1D, src/Common/GPUplan.cpp:122-141).
 * `nccl_unique_id` is the 128-byte ncclUniqueId produced by cmc_nccl_unique_id() on rank 0 and
 * distributed by the host (e.g. torch.distributed broadcast). */
int cmc_nccl_unique_id(void *id128);
int cmc_adi3d_create_dist(const cmc_grid_desc *grid, const cmc_fluid_params *params,
                          int fp_bytes, int device, int rank, int nranks,
                          const void *nccl_unique_id, cmc_adi3d **out);
/* the same slab decomposition with all `n_slabs` slabs held by ONE handle on ONE device, exchanging halos and
 * the reduced systems with device-to-device copies: exercises the multi-GPU code path on a single GPU (the
 * counterpart of the reference's MGPU_EMU switch, src/Common/GPUplan.h:10-15). */
int cmc_adi3d_create_emulated(const cmc_grid_desc *grid, const cmc_fluid_params *params,
                              int fp_bytes, int device, int n_slabs, cmc_adi3d **out);
/* ONE process, ONE host thread, n devices - the reference's own multi-GPU mode ("GPU <n>" on its command line:
 * FluidSolver3D.cpp:88-95, GPUplan::init enables peer access between the devices, GPUplan.cpp:35-77).  The handle holds
 * all n x-slabs, slab i on devices[i]; every slab has its own stream; the sweeps store their boundary planes and
 * interface systems straight into the neighbouring devices' buffers through peer access (NVLink), ordering is by CUDA
 * events.  All other entry points work as on a single-GPU handle (whole-grid arrays in, whole layers out). */
int cmc_adi3d_create_multi(const cmc_grid_desc *grid, const cmc_fluid_params *params, int fp_bytes,
                           const int *devices, int n_devices, cmc_adi3d **out);
/* ---- how the grid is cut into x-slabs: Grid3D::Split / SplitSegments_X (Grid3D.cpp:148-235), GPUplan::splitEven1D
 * (GPUplan.cpp:122-141), PARAplan::split1D (PARAplan.cpp:89-126).  planes_out[r] = x-planes of slab r.  The cut positions of
 * every policy are moved to multiples of 8 planes (the partitioned x-sweep works on 8-row chunks), the last slab takes the
 * rest; every slab holds at least 8 planes.
 *   CMC_SPLIT_EVEN_X        equal numbers of planes (the default of every cmc_adi3d_create_* without a split)
 *   CMC_SPLIT_EVEN_SEGMENTS equal numbers of line segments: a plane weighs the y- and z-segments that start in it plus its
 *                           share 1/size of every x-segment that crosses it (Grid3D.cpp:167-187)
 *   CMC_SPLIT_EVEN_VOLUME   equal numbers of NODE_IN cells (Grid3D.cpp:189-202)
 * `type` = the node types of the whole grid (dimx*dimy*dimz, CMC_NODE_*); not read for CMC_SPLIT_EVEN_X. */
enum { CMC_SPLIT_EVEN_X = 0, CMC_SPLIT_EVEN_SEGMENTS = 1, CMC_SPLIT_EVEN_VOLUME = 2 };
int cmc_split_planes(int policy, const cmc_grid_desc *grid, const int32_t *type, int n_slabs, int32_t *planes_out);

/* the general constructor: any of the decompositions above with a caller-supplied split (planes[n_slabs], e.g. from
 * cmc_split_planes; NULL = CMC_SPLIT_EVEN_X).  kind: 0 one slab on `device`; 1 n_slabs emulated on `device`; 2 slab `rank`
 * of n_slabs in this process on `device`, the others in other processes (nccl_unique_id); 3 all n_slabs in this process,
 * slab i on devices[i]. */
typedef struct cmc_decomp {
	int32_t kind, device, n_slabs, rank;
	const void *nccl_unique_id;
	const int32_t *devices;
	const int32_t *planes;
} cmc_decomp;
int cmc_adi3d_create_ex(const cmc_grid_desc *grid, const cmc_fluid_params *params, int fp_bytes, const cmc_decomp *decomp, cmc_adi3d **out);
int cmc_adi3d_destroy(cmc_adi3d *h);           /* ~AdiSolver3D (AdiSolver3D.cpp:153-158) */

/* planes held by this handle: global x range [x0, x0+nx) */
int cmc_adi3d_slab(const cmc_adi3d *h, int *x0, int *nx);

/* ---- grid nodes: replaces Grid3D::GetNodesCPU(), GetType, GetBC_vel, GetBC_temp, GetVel, GetT (Grid3D.h:114-124) ----
 * Arrays cover the WHOLE grid (dimx*dimy*dimz); a distributed handle takes its own slab from them.
 * Also initialises the time layers like `new TimeLayer3D(grid)` does (TimeLayer3D.h:734-751,1077-1090):
 * cur = (Node.v, Node.T) in every cell; half/next/temp (uninitialised in the reference) start as copies. */
int cmc_adi3d_set_nodes(cmc_adi3d *h, const int32_t *type, const int32_t *bc_vel, const int32_t *bc_temp,
                        const void *vx, const void *vy, const void *vz, const void *T);
/* same, straight from the reference's `Node*` (Grid3D.h:73-88): {int type, bc_vel, bc_temp; FTYPE vx,vy,vz,T}
 * (28 bytes in the fp32 build; 48 in fp64, where the FTYPE members start at offset 16) */
int cmc_adi3d_set_nodes_aos(cmc_adi3d *h, const void *nodes, size_t node_stride_bytes);

/* the same for ONE slab of a distributed run, from node arrays that cover only the planes [x0 - hlo, x0 + nx + hhi) of
 * the grid (x0, nx: cmc_adi3d_slab; hlo = min(halo, x0), hhi = min(halo, dimx - x0 - nx); halo >= 2) - no rank has to hold
 * the whole grid.  This is what Grid3D::Init_GPU does per device (node slices with their halo, Grid3D.cpp:567-596).  Needs
 * a grid without NODE_IN cells on its two x-faces (every grid the reference's loaders produce; CMC_ERR_UNSUPPORTED
 * otherwise). */
int cmc_adi3d_set_nodes_slab(cmc_adi3d *h, const int32_t *type, const int32_t *bc_vel, const int32_t *bc_temp,
                             const void *vx, const void *vy, const void *vz, const void *T, int halo_planes);

/* ---- moving boundaries: Grid3D::Prepare(t) (Grid3D.cpp:900-945, ComputeSubframeInfo; the driver's hook is
 * FluidSolver3D.cpp:237) rewrites the Node[] array between steps - types, boundary kinds and boundary values - and the
 * solver keeps its time layers.  Same arrays as cmc_adi3d_set_nodes / _aos; the layers are NOT touched.  Call
 * cmc_adi3d_build_lines afterwards: the line descriptors are rebuilt on the device (a scan per grid line, < 1 % of a step). */
int cmc_adi3d_update_nodes(cmc_adi3d *h, const int32_t *type, const int32_t *bc_vel, const int32_t *bc_temp,
                           const void *vx, const void *vy, const void *vz, const void *T);
int cmc_adi3d_update_nodes_aos(cmc_adi3d *h, const void *nodes, size_t node_stride_bytes);

/* ---- AdiSolver3D::CreateSegments (AdiSolver3D.cpp:553-562; Grid3D::GenerateListSegments, Grid3D.cpp:47-127):
 * builds the per-direction line descriptors (segment roles + boundary rows) on the device. */
int cmc_adi3d_build_lines(cmc_adi3d *h);
int cmc_adi3d_num_segments(const cmc_adi3d *h, int dir, int64_t *n);

/* ---- Solver3D::UpdateBoundaries (Solver3D.h:35, AdiSolver3D.cpp:286-304) ---- */
int cmc_adi3d_update_boundaries(cmc_adi3d *h);

/* ---- Solver3D::TimeStep (Solver3D.h:28, AdiSolver3D.cpp:306-391).  `dt` is rounded to FTYPE like the
 * caller's (FTYPE)dt cast (FluidSolver3D.cpp:242).  *err_out receives diffError (refreshed only when
 * compute_error != 0).  Returns CMC_ERR_DIVERGED (layers NOT swapped) when diffError > 0.01. */
int cmc_adi3d_time_step(cmc_adi3d *h, double dt, int num_global, int num_local,
                        int compute_error, double *err_out);

/* ---- Solver3D::GetLayer (Solver3D.h:33, Solver3D.cpp:21-25; TimeLayer3D::FilterToArrays, TimeLayer3D.h:819-924):
 * writes 99999 into the NODE_OUT cells of `next` (the PREVIOUS time layer - and keeps that mutation, like the
 * reference), then nearest-lower downsamples it.  vel_xyz = Vec3D[ox*oy*oz] (3 x FTYPE interleaved), T = double[].
 * Output dims of 0 mean "grid dims".  On a distributed handle every rank must call; rank 0 receives the result. */
int cmc_adi3d_get_layer(cmc_adi3d *h, void *vel_xyz, double *T, int outdimx, int outdimy, int outdimz);

/* ---- overlapped I/O (fused readback + output staging; reference: TimeLayer3D::FilterToArrays +
 * OutputNetCDF3D_layer between two time steps, TimeLayer3D.h:819-924, IO.h:350-388).
 * cmc_adi3d_get_layer_async = cmc_adi3d_get_layer without the wait: Clear + FilterToArrays (+ the gather of a distributed
 * handle) are enqueued behind the time steps already issued, the copy to vel_xyz / T runs on a separate copy stream out of
 * one of two staging sets, and the caller goes on stepping; cmc_adi3d_get_layer_wait returns when every readback started
 * so far has landed.  The host arrays must stay valid until then (pinned memory keeps the copy asynchronous).
 * cmc_adi3d_write_layer_async uploads a whole layer (u, v, w, T: dense arrays of this handle's planes) on the copy stream
 * into a staging buffer while the solver keeps computing; cmc_adi3d_write_layer_commit(layer) orders the solver's stream
 * behind that upload and scatters the staging buffer into the layer (the asynchronous form of cmc_adi3d_write_field x 4). */
/* Distributed output (one process per GPU): with option "local_output" = 1 (cmc_adi3d_set_option) GetLayer gathers nothing -
 * every rank receives the output rows whose source planes it holds, rows [row_lo, row_hi) of the whole output, stored from
 * the start of the arrays IT passed ((row_hi - row_lo) * outdimy * outdimz entries) - so N host links carry the result instead
 * of rank 0's alone (parallel output files, like the per-node slices of Grid3D::Init_GPU in the other direction). */
int cmc_adi3d_output_rows(const cmc_adi3d *h, int outdimx, int *row_lo, int *row_hi);
int cmc_adi3d_get_layer_async(cmc_adi3d *h, void *vel_xyz, double *T, int outdimx, int outdimy, int outdimz);
int cmc_adi3d_get_layer_wait(cmc_adi3d *h);
int cmc_adi3d_write_layer_async(cmc_adi3d *h, const void *u, const void *v, const void *w, const void *T);
int cmc_adi3d_write_layer_commit(cmc_adi3d *h, int layer);

/* ---- options ----  "mode": CMC_MODE_FAST | CMC_MODE_EXACT;  "tma": bit 0 / bit 1 = run the x / y sweeps as TMA-staged
 * persistent tiles where the grid allows (kernels_tma.cu; default from the environment variable CMC_TMA);  "profile": 0|1|2 (see cmc_adi3d_get_timing);
 * "tma_shape": 0 = measured default, else lines per tile (8 | 16) + 256 * CTAs per tile (1 | 2) of the TMA kernel;  "local_output": 0|1 (see
 * cmc_adi3d_output_rows);  "xs": 0|1 = one-pass slab-coupled x-sweep (off by default - it measured slower than the two-pass form;
 * its tables exist only when the environment variable CMC_XS=1 was set at creation);
 * read-only: "kernel_x" / "kernel_y" / "kernel_z" (which kernel a sweep along that axis runs: 0 exact Thomas kernels,
 * 1 direct-load partition kernel, 2 cp.async ring kernel, 3 TMA-staged tile kernel, 4 slab-coupled x-sweep in two passes, 5 in one pass),
 * "nzp" (padded z-line length), "jb" (rows per y-block of the field storage, 0 = one block), "exchange" (how slabs exchange planes and interface systems: 0 single slab,
 * 1 NCCL send/recv groups, 2 stores fused into the sweep kernels (slabs on one device), 3 the same into peer memory
 * over NVLink - every rank maps the other ranks' exchange arena with CUDA IPC, 4 the same between the devices of one
 * process - cmc_adi3d_create_multi) */
int cmc_adi3d_set_option(cmc_adi3d *h, const char *key, int64_t value);
int cmc_adi3d_get_option(const cmc_adi3d *h, const char *key, int64_t *value);

/* ---- test / debug hooks (the reference's TimeLayer3D dumps and sum_layer, AdiSolver3D.cpp:30-58) ---- */
/* dense host copy (this handle's slab: nx*dimy*dimz values) of one field of one layer */
int cmc_adi3d_read_field(cmc_adi3d *h, int layer, int var, void *dst);
int cmc_adi3d_write_field(cmc_adi3d *h, int layer, int var, const void *src);
/* TimeStep prologue only: next<-cur on BOUND/VALVE, temp<-cur (AdiSolver3D.cpp:310-320) */
int cmc_adi3d_step_prologue(cmc_adi3d *h);
/* one AdiSolver3D::SolveDirection (AdiSolver3D.cpp:564-666): num_local x (solve all lines, merge into temp) */
int cmc_adi3d_solve_direction(cmc_adi3d *h, int dir, double dt, int num_local, int cur_layer, int next_layer);
/* TimeLayer3D::EvalDivError of one layer (TimeLayer3D.h:595-641) */
int cmc_adi3d_eval_div_error(cmc_adi3d *h, int layer, double *err_out);

/* per-field checksums of one layer over all non-OUT cells of the WHOLE grid (all ranks of a distributed handle must
 * call; every rank receives the result): sums8 = { sum u, v, w, T, sum of squares u, v, w, T }.  The role of the
 * reference's sum_layer debug hook (AdiSolver3D.cpp:30-58). */
int cmc_adi3d_field_sums(cmc_adi3d *h, int layer, double *sums8);

/* ---- device-resident control for benchmarking (inputs already in HBM) ---- */
/* enqueue a step without any host synchronisation (compute_error results are fetched by cmc_adi3d_sync) */
int cmc_adi3d_time_step_async(cmc_adi3d *h, double dt, int num_global, int num_local, int compute_error);
int cmc_adi3d_sync(cmc_adi3d *h, double *err_out);
/* the CUDA stream all work of this handle is ordered on (a cudaStream_t), for event timing */
int cmc_adi3d_stream(const cmc_adi3d *h, void **stream_out);
/* kernels launched by this handle since creation / since the last reset */
int cmc_adi3d_launch_count(const cmc_adi3d *h, int64_t *n, int reset);
/* device time per kernel family, measured with CUDA event pairs on the handle's stream while option
 * "profile" is 1 (the reference's Profiler event names, src/FluidSolver3D/AdiSolver3D.cpp:297-367:
 * SolveSegments_X/Y/Z, MergeLayer, CopyLayer, UpdateBoundaries, EvalDivError; plus GetLayer's kernels).
 * set_option("profile", 2) also clears the accumulators. */
enum {
	CMC_TIMING_SWEEP_X = 0, CMC_TIMING_SWEEP_Y = 1, CMC_TIMING_SWEEP_Z = 2, CMC_TIMING_MERGE = 3,
	CMC_TIMING_COPY = 4, CMC_TIMING_BOUNDARY = 5, CMC_TIMING_RESIDUAL = 6, CMC_TIMING_READBACK = 7,
	CMC_TIMING_COMM = 8,
	/* slab-decomposed x-sweep: the spike pass and the interface solve are timed apart from the coupled pass
	 * (CMC_TIMING_SWEEP_X), so that each launch can be set against the bytes it moves */
	CMC_TIMING_X_SPIKE = 9, CMC_TIMING_X_INTERFACE = 10, CMC_TIMING_KINDS = 11
};
int cmc_adi3d_get_timing(cmc_adi3d *h, int kind, double *total_ms, int64_t *calls);
/* bytes of device memory held by this handle */
int cmc_adi3d_device_bytes(const cmc_adi3d *h, int64_t *n);

/* ---- standalone batched line solver (unit tests of the GPU tridiagonal kernels against
 * Common::SolveTridiagonal, src/Common/Algorithms.h:21-38).  Host arrays a,b,c,d,x: nsys*n, system-major. */
int cmc_solve_tridiagonal_batch(int fp_bytes, int mode, int nsys, int n,
                                const void *a, const void *b, const void *c, const void *d, void *x);

/* =====================================================================================================================
 * 2D ADI solver: replaces FluidSolver2D::AdiSolver2D behind FluidSolver2D::Solver2D (src/FluidSolver2D/Solver2D.h:24-45,
 * AdiSolver2D.h:40-64).  Host arrays use the reference's layout idx = i * dimy + j (TimeLayer2D.h:27-40).  The arithmetic
 * is the reference's, operation by operation (bit-identical results, including the data-dependent number of outer
 * iterations).  Layers: CMC_LAYER_CUR / _HALF / _NEXT / _TEMP; variables: 0 = u, 1 = v, 2 = T.
 * ===================================================================================================================== */
typedef struct cmc_adi2d cmc_adi2d;

/* `new AdiSolver2D` + the allocations of AdiSolver2D::Init (AdiSolver2D.cpp:21-35); dx, dy, startT = Grid2D::dx, dy, startT */
int cmc_adi2d_create(int dimx, int dimy, double dx, double dy, const cmc_fluid_params *params, double startT,
                     int fp_bytes, int device, cmc_adi2d **out);
int cmc_adi2d_destroy(cmc_adi2d *h);
/* what the solver reads through Grid2D::GetType / GetData (src/FluidSolver2D/Grid2D.h:52-54): node type, CondData2D::type
 * (CMC_BC_*), CondData2D::vel.x / .y / .T per cell.  The reference driver refreshes the grid before every step
 * (grid.Prepare(t), FluidSolver2D.cpp:129), so call this before every cmc_adi2d_update_boundaries / _time_step. */
int cmc_adi2d_set_grid(cmc_adi2d *h, const int32_t *type, const int32_t *bc_type, const void *vx, const void *vy, const void *T);
/* cur <- grid data in every cell: the second half of AdiSolver2D::Init (AdiSolver2D.cpp:36-50) */
int cmc_adi2d_init_layer(cmc_adi2d *h);
/* Solver2D::UpdateBoundaries (Solver2D.cpp:48-62) */
int cmc_adi2d_update_boundaries(cmc_adi2d *h);
/* AdiSolver2D::TimeStep (AdiSolver2D.cpp:279-323): iterates until it >= num_global AND the residual <= 0.1.  *err_out = the
 * residual the reference prints, *iters_out = outer iterations done.  CMC_ERR_DIVERGED where the reference exits
 * ("Exceeded max number of iterations", "Error is too big!"). */
int cmc_adi2d_time_step(cmc_adi2d *h, double dt, int num_global, int num_local, double *err_out, int *iters_out);
/* One TimeStep of a host-driven loop in one round trip - what B200AdiSolver2D::TimeStep needs, because Solver2D's non-virtual
 * UpdateBoundaries / SetGridBoundaries edit the layers on the host between steps (Solver2D.cpp:48-71): the grid arrays (as in
 * cmc_adi2d_set_grid) and the host's cur / next layers (3 dense arrays each: u, v, T) go up in ONE copy from a pinned staging
 * block, the step runs, both layers come back in ONE copy and are updated in place.  Same result as cmc_adi2d_set_grid +
 * 6 x write_field + cmc_adi2d_time_step + 6 x read_field (17 synchronised copies). */
int cmc_adi2d_step_host(cmc_adi2d *h, const int32_t *type, const int32_t *bc_type, const void *vx, const void *vy, const void *T,
                        void *const cur_uvT[3], void *const next_uvT[3], double dt, int num_global, int num_local,
                        double *err_out, int *iters_out);
/* many independent 2D cases in ONE launch, one thread block (one SM) per case - the throughput form of the 2D solver
 * (a single ~16 k-cell case cannot fill a GPU).  All handles: same precision, grid dimensions and
 * device.  update_boundaries != 0 runs Solver2D::UpdateBoundaries of every case first.  Per case: residual, outer
 * iterations and status (CMC_OK / CMC_ERR_DIVERGED); returns the first non-OK status.  Results are bit-identical with n
 * separate cmc_adi2d_time_step calls. */
int cmc_adi2d_time_step_batch(cmc_adi2d *const *handles, int n, double dt, int num_global, int num_local, int update_boundaries,
                              double *err_out, int *iters_out, int *status_out);
/* Solver2D::GetLayer (Solver2D.cpp:20-34): nearest-lower downsample of `next`; vel_xy = Vec2D[ox*oy] (2 x FTYPE), T = double[] */
int cmc_adi2d_get_layer(cmc_adi2d *h, void *vel_xy, double *T, int outdimx, int outdimy);
/* dense host copy of one field of one layer (Solver2D::SetGridBoundaries, Solver2D.cpp:64-71, reads cur.u / cur.v this
 * way; Solver2D::SetLayer, :36-46, writes cur) */
int cmc_adi2d_read_field(cmc_adi2d *h, int layer, int var, void *dst);
int cmc_adi2d_write_field(cmc_adi2d *h, int layer, int var, const void *src);
int cmc_adi2d_launch_count(const cmc_adi2d *h, int64_t *n);

#ifdef __cplusplus
}
#endif
#endif /* CMC_ADI_H */
