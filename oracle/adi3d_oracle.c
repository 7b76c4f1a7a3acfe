/* TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.
 *
 * CPU restatement ("port") of the reference's 3D ADI time step, CPU semantics
 * (SURVEY.md 8 note N1), used ONLY as the checker by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  The product path
 * (cmc_fluid_solver_b200/csrc) never links, loads or calls this file.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py compares this restatement
 * bit-for-bit (fp32 and fp64) against the reference's own compiled CPU solver
 * (oracle/_ref/ref_probe3d_*, built from /root/reference/src by oracle/build_ref.sh) and
 * against golden vectors that binary produced (tests/golden/, generator committed).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src).  Expression order and operand types follow the source exactly so
 * that a -O2 -ffp-contract=off build is bit-identical with the reference's -O2 build.
 *
 * Compiled twice (-DOR_FT=float -DOR_SUF=_f32 and -DOR_FT=double -DOR_SUF=_f64) into
 * oracle/_build/liboracle_adi.so.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef OR_FT
#define OR_FT double
#define OR_SUF _f64
#endif
#define OR_CAT2(a, b) a##b
#define OR_CAT(a, b) OR_CAT2(a, b)
#define FN(name) OR_CAT(name, OR_SUF)

typedef OR_FT FT;

/* Common/Geometry.h:29-43 */
enum { NODE_IN = 0, NODE_OUT = 1, NODE_BOUND = 2, NODE_VALVE = 3 };
enum { BC_NOSLIP = 0, BC_FREE = 1 };
enum { DIR_X = 0, DIR_Y = 1, DIR_Z = 2 };
enum { VAR_U = 0, VAR_V = 1, VAR_W = 2, VAR_T = 3 };          /* AdiSolver3D.h:38 */
enum { LAYER_CUR = 0, LAYER_HALF = 1, LAYER_NEXT = 2, LAYER_TEMP = 3 };

#define OR_MISSING_VALUE 99999.0f                              /* Geometry.h:25 */
#define OR_ERR_THRESHOLD 0.01                                  /* AdiSolver3D.h:32 */

typedef struct { int posx, posy, posz, endx, endy, endz, size, dir; } Seg; /* Grid3D.h:63-71 */

typedef struct {
	int dimx, dimy, dimz;
	FT dx, dy, dz;                         /* TimeLayer3D ctor casts grid->dx to FTYPE, AdiSolver3D.cpp:256 */
	FT v_T, v_vis, t_vis, t_phi;           /* Geometry.h:538-562 */
	/* Node[] as structure of arrays (Grid3D.h:73-88) */
	int *type, *bc_vel, *bc_temp;
	FT *nvx, *nvy, *nvz, *nT;
	/* layers: slot -> 4 fields; cur/next are swapped by pointer like AdiSolver3D.cpp:388-390 */
	FT *layer[4][4];
	Seg *segs[3];
	int numSegs[3];
	int max_n;
	double diffError;
} Oracle;

static inline size_t IDX(const Oracle *o, int i, int j, int k)
{
	return ((size_t)i * o->dimy + j) * (size_t)o->dimz + k;   /* TimeLayer3D.h:256-259 (haloSize 0 on CPU) */
}

/* ---- Grid3D::GenerateListSegments, Grid3D.cpp:47-127 (nblockZ == 1) --------------------- */
static int gen_segments(const Oracle *o, int dir, Seg *out)
{
	int dim2, dim3, n = 0;
	const int dimx = o->dimx, dimy = o->dimy, dimz = o->dimz;
	switch (dir) {                                              /* AdiSolver3D.cpp:557-559 */
	case DIR_X: dim2 = dimy; dim3 = dimz; break;
	case DIR_Y: dim2 = dimx; dim3 = dimz; break;
	default:    dim2 = dimx; dim3 = dimy; break;
	}
	for (int i = 0; i < dim2; i++)
		for (int j = 0; j < dim3; j++) {
			Seg seg, new_seg;
			int state = 0, incx = 0, incy = 0, incz = 0;
			memset(&new_seg, 0, sizeof(new_seg));
			switch (dir) {
			case DIR_X: seg.posx = 0; seg.posy = i; seg.posz = j; incx = 1; break;
			case DIR_Y: seg.posx = i; seg.posy = 0; seg.posz = j; incy = 1; break;
			default:    seg.posx = i; seg.posy = j; seg.posz = 0; incz = 1; break;
			}
			seg.dir = dir;
			while ((seg.posx + incx < dimx) && (seg.posy + incy < dimy) && (seg.posz + incz < dimz)) {
				if (o->type[IDX(o, seg.posx + incx, seg.posy + incy, seg.posz + incz)] == NODE_IN) {
					if (state == 0) new_seg = seg;
					state = 1;
				} else if (state == 1) {
					new_seg.endx = seg.posx + incx;
					new_seg.endy = seg.posy + incy;
					new_seg.endz = seg.posz + incz;
					new_seg.size = (new_seg.endx - new_seg.posx) + (new_seg.endy - new_seg.posy) + (new_seg.endz - new_seg.posz) + 1;
					if (out) out[n] = new_seg;
					n++;
					state = 0;
				}
				seg.posx += incx; seg.posy += incy; seg.posz += incz;
			}
			/* an unterminated run at the domain edge is dropped, exactly like the reference */
		}
	return n;
}

/* ---- ScalarField3D::d_x/d_y/d_z, TimeLayer3D.h:338-340 ----------------------------------- */
static inline FT d_x(const Oracle *o, const FT *f, int i, int j, int k) { return (f[IDX(o, i + 1, j, k)] - f[IDX(o, i - 1, j, k)]) / (2 * o->dx); }
static inline FT d_y(const Oracle *o, const FT *f, int i, int j, int k) { return (f[IDX(o, i, j + 1, k)] - f[IDX(o, i, j - 1, k)]) / (2 * o->dy); }
static inline FT d_z(const Oracle *o, const FT *f, int i, int j, int k) { return (f[IDX(o, i, j, k + 1)] - f[IDX(o, i, j, k - 1)]) / (2 * o->dz); }

/* ---- TimeLayer3D::DissFuncX/Y/Z, TimeLayer3D.h:554-588 ----------------------------------- */
static inline FT DissFuncX(const Oracle *o, FT *const *L, int i, int j, int k)
{
	FT u_x = d_x(o, L[VAR_U], i, j, k), v_x = d_x(o, L[VAR_V], i, j, k), w_x = d_x(o, L[VAR_W], i, j, k);
	FT u_y = d_y(o, L[VAR_U], i, j, k), u_z = d_z(o, L[VAR_U], i, j, k);
	return 2 * u_x * u_x + v_x * v_x + w_x * w_x + v_x * u_y + w_x * u_z;
}
static inline FT DissFuncY(const Oracle *o, FT *const *L, int i, int j, int k)
{
	FT u_y = d_y(o, L[VAR_U], i, j, k), v_y = d_y(o, L[VAR_V], i, j, k), w_y = d_y(o, L[VAR_W], i, j, k);
	FT v_x = d_x(o, L[VAR_V], i, j, k), v_z = d_z(o, L[VAR_V], i, j, k);
	return u_y * u_y + 2 * v_y * v_y + w_y * w_y + u_y * v_x + w_y * v_z;
}
static inline FT DissFuncZ(const Oracle *o, FT *const *L, int i, int j, int k)
{
	FT u_z = d_z(o, L[VAR_U], i, j, k), v_z = d_z(o, L[VAR_V], i, j, k), w_z = d_z(o, L[VAR_W], i, j, k);
	FT w_x = d_x(o, L[VAR_W], i, j, k), w_y = d_y(o, L[VAR_W], i, j, k);
	return u_z * u_z + v_z * v_z + 2 * w_z * w_z + u_z * w_x + v_z * w_y;
}

/* ---- Common::SolveTridiagonal, Common/Algorithms.h:21-38 --------------------------------- */
static void SolveTridiagonal(FT *a, FT *b, FT *c, FT *d, FT *x, int num)
{
	c[num - 1] = 0.0;
	c[0] = c[0] / b[0];
	d[0] = d[0] / b[0];
	for (int i = 1; i < num; i++) {
		c[i] = c[i] / (b[i] - a[i] * c[i - 1]);
		d[i] = (d[i] - d[i - 1] * a[i]) / (b[i] - a[i] * c[i - 1]);
	}
	x[num - 1] = d[num - 1];
	for (int i = num - 2; i >= 0; i--)
		x[i] = d[i] - c[i] * x[i + 1];
}

/* ---- AdiSolver3D::ApplyBC0 / ApplyBC1, AdiSolver3D.cpp:804-852 --------------------------- */
static FT node_value(const Oracle *o, size_t id, int var)
{
	switch (var) {
	case VAR_U: return o->nvx[id];
	case VAR_V: return o->nvy[id];
	case VAR_W: return o->nvz[id];
	default:    return o->nT[id];
	}
}
static void ApplyBC0(const Oracle *o, int i, int j, int k, int var, FT *b0, FT *c0, FT *d0)
{
	size_t id = IDX(o, i, j, k);
	if ((var == VAR_T && o->bc_temp[id] == BC_FREE) || (var != VAR_T && o->bc_vel[id] == BC_FREE)) {
		*b0 = 2.0; *c0 = -1.0; *d0 = 0.0;
	} else {
		*b0 = 1.0; *c0 = 0.0; *d0 = node_value(o, id, var);
	}
}
static void ApplyBC1(const Oracle *o, int i, int j, int k, int var, FT *a1, FT *b1, FT *d1)
{
	size_t id = IDX(o, i, j, k);
	if ((var == VAR_T && o->bc_temp[id] == BC_FREE) || (var != VAR_T && o->bc_vel[id] == BC_FREE)) {
		*a1 = -1.0; *b1 = 2.0; *d1 = 0.0;
	} else {
		*a1 = 0.0; *b1 = 1.0; *d1 = node_value(o, id, var);
	}
}

/* ---- AdiSolver3D::BuildMatrix, AdiSolver3D.cpp:732-802 ----------------------------------- */
static void BuildMatrix(const Oracle *o, FT dt, int i, int j, int k, int var, int dir,
                        FT *a, FT *b, FT *c, FT *d, int n, FT *const *cur, FT *const *temp)
{
	FT vis_dx2, vis_dy2, vis_dz2;
	const FT dx = o->dx, dy = o->dy, dz = o->dz;
	if (var == VAR_T) {
		vis_dx2 = o->t_vis / (dx * dx); vis_dy2 = o->t_vis / (dy * dy); vis_dz2 = o->t_vis / (dz * dz);
	} else {
		vis_dx2 = o->v_vis / (dx * dx); vis_dy2 = o->v_vis / (dy * dy); vis_dz2 = o->v_vis / (dz * dz);
	}
	for (int p = 1; p < n - 1; p++) {
		switch (dir) {
		case DIR_X: {
			size_t id = IDX(o, i + p, j, k);
			a[p] = -temp[VAR_U][id] / (2 * dx) - vis_dx2;
			b[p] = 3 / dt + 2 * vis_dx2;
			c[p] = temp[VAR_U][id] / (2 * dx) - vis_dx2;
			switch (var) {
			case VAR_U: d[p] = cur[VAR_U][id] * 3 / dt - o->v_T * d_x(o, temp[VAR_T], i + p, j, k); break;
			case VAR_V: d[p] = cur[VAR_V][id] * 3 / dt; break;
			case VAR_W: d[p] = cur[VAR_W][id] * 3 / dt; break;
			default:    d[p] = cur[VAR_T][id] * 3 / dt + o->t_phi * DissFuncX(o, temp, i + p, j, k); break;
			}
			break;
		}
		case DIR_Y: {
			size_t id = IDX(o, i, j + p, k);
			a[p] = -temp[VAR_V][id] / (2 * dy) - vis_dy2;
			b[p] = 3 / dt + 2 * vis_dy2;
			c[p] = temp[VAR_V][id] / (2 * dy) - vis_dy2;
			switch (var) {
			case VAR_U: d[p] = cur[VAR_U][id] * 3 / dt; break;
			case VAR_V: d[p] = cur[VAR_V][id] * 3 / dt - o->v_T * d_y(o, temp[VAR_T], i, j + p, k); break;
			case VAR_W: d[p] = cur[VAR_W][id] * 3 / dt; break;
			default:    d[p] = cur[VAR_T][id] * 3 / dt + o->t_phi * DissFuncY(o, temp, i, j + p, k); break;
			}
			break;
		}
		default: {
			size_t id = IDX(o, i, j, k + p);
			a[p] = -temp[VAR_W][id] / (2 * dz) - vis_dz2;
			b[p] = 3 / dt + 2 * vis_dz2;
			c[p] = temp[VAR_W][id] / (2 * dz) - vis_dz2;
			switch (var) {
			case VAR_U: d[p] = cur[VAR_U][id] * 3 / dt; break;
			case VAR_V: d[p] = cur[VAR_V][id] * 3 / dt; break;
			case VAR_W: d[p] = cur[VAR_W][id] * 3 / dt - o->v_T * d_z(o, temp[VAR_T], i, j, k + p); break;
			default:    d[p] = cur[VAR_T][id] * 3 / dt + o->t_phi * DissFuncZ(o, temp, i, j, k + p); break;
			}
			break;
		}
		}
	}
}

/* ---- AdiSolver3D::SolveSegment + UpdateSegment, AdiSolver3D.cpp:687-730 ------------------ */
static void SolveSegment(const Oracle *o, FT dt, const Seg *seg, int var, int dir,
                         FT *const *cur, FT *const *temp, FT *const *next, FT *scratch)
{
	const int n = seg->size;
	FT *a = scratch, *b = a + o->max_n, *c = b + o->max_n, *d = c + o->max_n, *x = d + o->max_n;
	ApplyBC0(o, seg->posx, seg->posy, seg->posz, var, &b[0], &c[0], &d[0]);
	ApplyBC1(o, seg->endx, seg->endy, seg->endz, var, &a[n - 1], &b[n - 1], &d[n - 1]);
	BuildMatrix(o, dt, seg->posx, seg->posy, seg->posz, var, dir, a, b, c, d, n, cur, temp);
	SolveTridiagonal(a, b, c, d, x, n);
	int i = seg->posx, j = seg->posy, k = seg->posz;
	for (int t = 0; t < n; t++) {
		next[var][IDX(o, i, j, k)] = x[t];
		switch (seg->dir) { case DIR_X: i++; break; case DIR_Y: j++; break; default: k++; break; }
	}
}

/* ---- ScalarField3D::MergeFieldTo / TimeLayer3D::MergeLayerTo, TimeLayer3D.h:415-436,664-683 */
static void MergeLayerTo(const Oracle *o, FT *const *src, FT *const *dest, int type)
{
	const size_t N = (size_t)o->dimx * o->dimy * o->dimz;
	for (int q = 0; q < 4; q++) {
		const FT *s = src[q];
		FT *dd = dest[q];
		#pragma omp parallel for schedule(static)
		for (size_t id = 0; id < N; id++)
			if (o->type[id] == type)
				dd[id] = (dd[id] + s[id]) / 2;
	}
}

/* ---- ScalarField3D::CopyFieldTo (masked), TimeLayer3D.h:394-413, 726-732 ----------------- */
static void CopyLayerMasked(const Oracle *o, FT *const *src, FT *const *dest, int type)
{
	const size_t N = (size_t)o->dimx * o->dimy * o->dimz;
	for (int q = 0; q < 4; q++)
		for (size_t id = 0; id < N; id++)
			if (o->type[id] == type) dest[q][id] = src[q][id];
}

/* ---- TimeLayer3D::CopyLayerTo (full), TimeLayer3D.h:685-697 ------------------------------ */
static void CopyLayerFull(const Oracle *o, FT *const *src, FT *const *dest)
{
	const size_t N = (size_t)o->dimx * o->dimy * o->dimz;
	for (int q = 0; q < 4; q++) memcpy(dest[q], src[q], N * sizeof(FT));
}

/* ---- AdiSolver3D::SolveDirection (CPU branch), AdiSolver3D.cpp:564-666 ------------------- */
static void SolveDirection(Oracle *o, int dir, FT dt, int num_local, FT *const *cur, FT *const *temp, FT *const *next)
{
	const Seg *list = o->segs[dir];
	const int ns = o->numSegs[dir];
	for (int it = 0; it < num_local; it++) {
		#pragma omp parallel
		{
			FT *scratch = (FT *)malloc(sizeof(FT) * 5 * (size_t)o->max_n);
			#pragma omp for schedule(static)
			for (int s = 0; s < ns; s++) {
				SolveSegment(o, dt, &list[s], VAR_U, dir, cur, temp, next, scratch);
				SolveSegment(o, dt, &list[s], VAR_V, dir, cur, temp, next, scratch);
				SolveSegment(o, dt, &list[s], VAR_W, dir, cur, temp, next, scratch);
				SolveSegment(o, dt, &list[s], VAR_T, dir, cur, temp, next, scratch);
			}
			free(scratch);
		}
		MergeLayerTo(o, next, temp, NODE_IN);                 /* :647-653 */
	}
}

/* ---- TimeLayer3D::EvalDivError, TimeLayer3D.h:595-641 (single rank: ndimx = dimx-1) ------
 * The reference reads (i-1, j-1, k-1) neighbours without a bounds check; an IN cell on a
 * low face is undefined behaviour there.  Such cells are skipped here (and in the CUDA path). */
static double EvalDivError(const Oracle *o, FT *const *L)
{
	const FT *U = L[VAR_U], *V = L[VAR_V], *W = L[VAR_W];
	const FT dx = o->dx, dy = o->dy, dz = o->dz;
	double err = 0.0;
	long count = 0;
	for (int i = 0; i < o->dimx - 1; i++)
		for (int j = 0; j < o->dimy - 1; j++)
			for (int k = 0; k < o->dimz - 1; k++)
				if (o->type[IDX(o, i, j, k)] == NODE_IN) {
					if (i == 0 || j == 0 || k == 0) continue;
					double err_x = (U[IDX(o, i, j, k)] + U[IDX(o, i, j - 1, k)] + U[IDX(o, i, j - 1, k - 1)] + U[IDX(o, i, j, k - 1)] -
						U[IDX(o, i - 1, j, k)] - U[IDX(o, i - 1, j - 1, k)] - U[IDX(o, i - 1, j - 1, k - 1)] - U[IDX(o, i - 1, j, k - 1)]) * dz * dy / 4.0;
					double err_y = (V[IDX(o, i, j, k)] + V[IDX(o, i - 1, j, k)] + V[IDX(o, i - 1, j, k - 1)] + V[IDX(o, i, j, k - 1)] -
						V[IDX(o, i, j - 1, k)] - V[IDX(o, i - 1, j - 1, k)] - V[IDX(o, i - 1, j - 1, k - 1)] - V[IDX(o, i, j - 1, k - 1)]) * dx * dz / 4.0;
					double err_z = (W[IDX(o, i, j, k)] + W[IDX(o, i, j - 1, k)] + W[IDX(o, i - 1, j - 1, k)] + W[IDX(o, i - 1, j, k)] -
						W[IDX(o, i, j, k - 1)] - W[IDX(o, i, j - 1, k - 1)] - W[IDX(o, i - 1, j - 1, k - 1)] - W[IDX(o, i - 1, j, k - 1)]) * dx * dy / 4.0;
					err += fabs(err_x + err_y + err_z);
					count++;
				}
	return err / count;
}

/* ======================================= public C API ==================================== */

void *FN(oracle3d_create)(int dimx, int dimy, int dimz, double dx, double dy, double dz,
                          double v_T, double v_vis, double t_vis, double t_phi,
                          const int *type, const int *bc_vel, const int *bc_temp,
                          const FT *vx, const FT *vy, const FT *vz, const FT *T)
{
	Oracle *o = (Oracle *)calloc(1, sizeof(Oracle));
	const size_t N = (size_t)dimx * dimy * dimz;
	o->dimx = dimx; o->dimy = dimy; o->dimz = dimz;
	o->dx = (FT)dx; o->dy = (FT)dy; o->dz = (FT)dz;
	o->v_T = (FT)v_T; o->v_vis = (FT)v_vis; o->t_vis = (FT)t_vis; o->t_phi = (FT)t_phi;
	o->type = (int *)malloc(N * 4); o->bc_vel = (int *)malloc(N * 4); o->bc_temp = (int *)malloc(N * 4);
	o->nvx = (FT *)malloc(N * sizeof(FT)); o->nvy = (FT *)malloc(N * sizeof(FT));
	o->nvz = (FT *)malloc(N * sizeof(FT)); o->nT = (FT *)malloc(N * sizeof(FT));
	memcpy(o->type, type, N * 4); memcpy(o->bc_vel, bc_vel, N * 4); memcpy(o->bc_temp, bc_temp, N * 4);
	memcpy(o->nvx, vx, N * sizeof(FT)); memcpy(o->nvy, vy, N * sizeof(FT));
	memcpy(o->nvz, vz, N * sizeof(FT)); memcpy(o->nT, T, N * sizeof(FT));
	o->max_n = dimx > dimy ? (dimx > dimz ? dimx : dimz) : (dimy > dimz ? dimy : dimz);
	/* cur = TimeLayer3D(grid): every cell takes Node.v / Node.T (TimeLayer3D.h:734-751, 1077-1090).
	 * half/next/temp are uninitialised in the reference (TimeLayer3D.h:353); defined here as
	 * copies of cur - the probe driver does the same to the real reference. */
	for (int l = 0; l < 4; l++)
		for (int q = 0; q < 4; q++) {
			o->layer[l][q] = (FT *)malloc(N * sizeof(FT));
			const FT *src = q == VAR_U ? vx : q == VAR_V ? vy : q == VAR_W ? vz : T;
			memcpy(o->layer[l][q], src, N * sizeof(FT));
		}
	o->diffError = 0.0;
	return o;
}

/* Moving boundaries: Grid3D::Prepare(t) (Grid3D.cpp:900-945, ComputeSubframeInfo) rewrites the Node[] array between steps
 * and the solver reads it through its grid pointer; the time layers are left as they are.  The segment lists must be
 * rebuilt afterwards (create_segments): the 2D solver does so inside every TimeStep (AdiSolver2D.cpp:279-283), the 3D
 * driver has the Prepare call commented out (FluidSolver3D.cpp:237). */
void FN(oracle3d_update_nodes)(void *h, const int *type, const int *bc_vel, const int *bc_temp,
                               const FT *vx, const FT *vy, const FT *vz, const FT *T)
{
	Oracle *o = (Oracle *)h;
	const size_t N = (size_t)o->dimx * o->dimy * o->dimz;
	memcpy(o->type, type, N * 4); memcpy(o->bc_vel, bc_vel, N * 4); memcpy(o->bc_temp, bc_temp, N * 4);
	memcpy(o->nvx, vx, N * sizeof(FT)); memcpy(o->nvy, vy, N * sizeof(FT));
	memcpy(o->nvz, vz, N * sizeof(FT)); memcpy(o->nT, T, N * sizeof(FT));
}

void FN(oracle3d_destroy)(void *h)
{
	Oracle *o = (Oracle *)h;
	if (!o) return;
	free(o->type); free(o->bc_vel); free(o->bc_temp);
	free(o->nvx); free(o->nvy); free(o->nvz); free(o->nT);
	for (int l = 0; l < 4; l++) for (int q = 0; q < 4; q++) free(o->layer[l][q]);
	for (int d = 0; d < 3; d++) free(o->segs[d]);
	free(o);
}

/* AdiSolver3D::CreateSegments, AdiSolver3D.cpp:553-562 */
void FN(oracle3d_create_segments)(void *h)
{
	Oracle *o = (Oracle *)h;
	for (int d = 0; d < 3; d++) {
		free(o->segs[d]);
		int n = gen_segments(o, d, NULL);
		o->segs[d] = (Seg *)malloc(sizeof(Seg) * (size_t)(n > 0 ? n : 1));
		o->numSegs[d] = gen_segments(o, d, o->segs[d]);
	}
}

int FN(oracle3d_num_segments)(void *h, int dir) { return ((Oracle *)h)->numSegs[dir]; }

/* copies segments as 8 ints each: posx,posy,posz,endx,endy,endz,size,dir */
void FN(oracle3d_get_segments)(void *h, int dir, int *out)
{
	Oracle *o = (Oracle *)h;
	memcpy(out, o->segs[dir], sizeof(Seg) * (size_t)o->numSegs[dir]);
}

/* AdiSolver3D::UpdateBoundaries (CPU), AdiSolver3D.cpp:286-294 -> CopyFromGrid(grid,type), TimeLayer3D.h:926-944 */
void FN(oracle3d_update_boundaries)(void *h)
{
	Oracle *o = (Oracle *)h;
	const size_t N = (size_t)o->dimx * o->dimy * o->dimz;
	FT **cur = o->layer[LAYER_CUR];
	for (int pass = 0; pass < 2; pass++) {
		const int target = pass == 0 ? NODE_BOUND : NODE_VALVE;
		for (size_t id = 0; id < N; id++)
			if (o->type[id] == target) {
				cur[VAR_U][id] = o->nvx[id]; cur[VAR_V][id] = o->nvy[id];
				cur[VAR_W][id] = o->nvz[id]; cur[VAR_T][id] = o->nT[id];
			}
	}
}

/* AdiSolver3D::TimeStep, AdiSolver3D.cpp:306-391.  Returns 0, or 1 when the divergence guard
 * fires (the reference prints "Error is too big!" and throws std::runtime_error("")). */
int FN(oracle3d_time_step)(void *h, double dt_in, int num_global, int num_local, int computeError, double *err_out)
{
	Oracle *o = (Oracle *)h;
	const FT dt = (FT)dt_in;                                   /* FluidSolver3D.cpp:242 casts to FTYPE */
	FT **cur = o->layer[LAYER_CUR], **half = o->layer[LAYER_HALF], **next = o->layer[LAYER_NEXT], **temp = o->layer[LAYER_TEMP];
	CopyLayerMasked(o, cur, next, NODE_BOUND);                 /* :310 */
	CopyLayerMasked(o, cur, next, NODE_VALVE);                 /* :311 */
	CopyLayerFull(o, cur, temp);                               /* :320 */
	for (int it = 0; it < num_global; it++) {                  /* :335-358 */
		SolveDirection(o, DIR_Z, dt, num_local, cur, temp, next);
		SolveDirection(o, DIR_Y, dt, num_local, next, temp, half);
		SolveDirection(o, DIR_X, dt, num_local, half, temp, next);
		MergeLayerTo(o, next, temp, NODE_IN);                  /* :354 */
	}
	if (computeError) o->diffError = EvalDivError(o, next);    /* :363-368 */
	if (err_out) *err_out = o->diffError;
	if (o->diffError > OR_ERR_THRESHOLD) return 1;             /* :371-374 */
	for (int q = 0; q < 4; q++) {                              /* :388-390 swap(cur,next) */
		FT *t = o->layer[LAYER_NEXT][q];
		o->layer[LAYER_NEXT][q] = o->layer[LAYER_CUR][q];
		o->layer[LAYER_CUR][q] = t;
	}
	return 0;
}

/* component hook: one SolveDirection between arbitrary layer slots */
void FN(oracle3d_solve_direction)(void *h, int dir, double dt, int num_local, int cur_slot, int temp_slot, int next_slot)
{
	Oracle *o = (Oracle *)h;
	SolveDirection(o, dir, (FT)dt, num_local, o->layer[cur_slot], o->layer[temp_slot], o->layer[next_slot]);
}

/* component hook: the TimeStep prologue only (AdiSolver3D.cpp:310-320) */
void FN(oracle3d_step_prologue)(void *h)
{
	Oracle *o = (Oracle *)h;
	CopyLayerMasked(o, o->layer[LAYER_CUR], o->layer[LAYER_NEXT], NODE_BOUND);
	CopyLayerMasked(o, o->layer[LAYER_CUR], o->layer[LAYER_NEXT], NODE_VALVE);
	CopyLayerFull(o, o->layer[LAYER_CUR], o->layer[LAYER_TEMP]);
}

double FN(oracle3d_eval_div_error)(void *h, int slot) { Oracle *o = (Oracle *)h; return EvalDivError(o, o->layer[slot]); }

/* raw pointer to a field (numpy view in tests) */
void *FN(oracle3d_field)(void *h, int slot, int var) { return ((Oracle *)h)->layer[slot][var]; }

/* Solver3D::GetLayer, Solver3D.cpp:21-25: next->Clear(OUT -> 99999) (TimeLayer3D.h:974-998) then
 * FilterToArrays (TimeLayer3D.h:819-856, 916-920).  vel is Vec3D[] = 3 x FTYPE interleaved. */
void FN(oracle3d_get_layer)(void *h, FT *vel, double *T, int outdimx, int outdimy, int outdimz)
{
	Oracle *o = (Oracle *)h;
	const size_t N = (size_t)o->dimx * o->dimy * o->dimz;
	FT **next = o->layer[LAYER_NEXT];
	for (size_t id = 0; id < N; id++)
		if (o->type[id] == NODE_OUT) {
			next[VAR_U][id] = OR_MISSING_VALUE; next[VAR_V][id] = OR_MISSING_VALUE;
			next[VAR_W][id] = OR_MISSING_VALUE; next[VAR_T][id] = OR_MISSING_VALUE;
		}
	if (outdimx == 0) outdimx = o->dimx;
	if (outdimy == 0) outdimy = o->dimy;
	if (outdimz == 0) outdimz = o->dimz;
	for (int i = 0; i < outdimx; i++)
		for (int j = 0; j < outdimy; j++)
			for (int k = 0; k < outdimz; k++) {
				int x = (i * o->dimx / outdimx), y = (j * o->dimy / outdimy), z = (k * o->dimz / outdimz);
				size_t ind = ((size_t)i * outdimy + j) * outdimz + k, id = IDX(o, x, y, z);
				vel[3 * ind + 0] = next[VAR_U][id];
				vel[3 * ind + 1] = next[VAR_V][id];
				vel[3 * ind + 2] = next[VAR_W][id];
				T[ind] = next[VAR_T][id];
			}
}

/* OpenMP thread count of the segment loop.  The reference's `omp for` over segments (AdiSolver3D.cpp:593-603)
 * races when two segments share an end cell whose boundary row is BC_FREE; with one thread the later segment
 * of the list wins, which is the order the CUDA path implements. */
void FN(oracle_set_threads)(int n)
{
#ifdef _OPENMP
	omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
	(void)n;
#endif
}

/* standalone Thomas (unit tests of the GPU line solvers) */
void FN(oracle_solve_tridiagonal)(FT *a, FT *b, FT *c, FT *d, FT *x, int num) { SolveTridiagonal(a, b, c, d, x, num); }
