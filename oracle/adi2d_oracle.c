/* TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.
 *
 * CPU restatement ("port") of the reference's 2D ADI time step, used ONLY as the checker by
 * tests/.  The product path (cmc_fluid_solver_b200/csrc) never links, loads or calls this file.
 *
 * Parity status: PINNED.  tests/test_oracle2d.py compares this restatement bit-for-bit (fp32) against golden vectors
 * produced by the reference's own compiled 2D solver (oracle/_ref/ref_probe2d_f32, built from /root/reference/src by
 * oracle/build_ref.sh) on the reference's data/2D/box_pipe case, and against that binary itself when it is present.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference/src/FluidSolver2D).
 * Expression order and operand types follow the source so that a -O2 -ffp-contract=off build is bit-identical with
 * the reference's -O2 build.  The grid (Grid2D::GetType / GetData) is an INPUT that the caller refreshes before every
 * step: the reference driver calls grid.Prepare(t) per step (FluidSolver2D.cpp:129).
 *
 * Compiled twice (-DOR_FT=float -DOR_SUF=_f32 and -DOR_FT=double -DOR_SUF=_f64) into oracle/_build/liboracle_adi.so.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>

#ifndef OR_FT
#define OR_FT float
#define OR_SUF _f32
#endif
#define OR_CAT2(a, b) a##b
#define OR_CAT(a, b) OR_CAT2(a, b)
#define FN(name) OR_CAT(name, OR_SUF)

typedef OR_FT FT;

enum { NODE_IN = 0, NODE_OUT = 1, NODE_BOUND = 2, NODE_VALVE = 3 };   /* Common/Geometry.h:29-43 */
enum { BC_NOSLIP = 0, BC_FREE = 1 };
enum { D_X = 0, D_Y = 1 };                                             /* AdiSolver2D.h:30 */
enum { T_U = 0, T_V = 1, T_T = 2 };                                    /* AdiSolver2D.h:29 */
enum { L_CUR = 0, L_HALF = 1, L_NEXT = 2, L_TEMP = 3, L_NEXT_LOCAL = 4, L_TEMP_LOCAL = 5 };
#define ERR_THRESHOLD 0.1                                              /* AdiSolver2D.h:24 */
#define MAX_GLOBAL_ITERS 100                                           /* AdiSolver2D.h:25 */

typedef struct { int posx, posy, endx, endy, size, dir; } Seg2;        /* AdiSolver2D.h:32-38 */

typedef struct {
	int dimx, dimy;
	FT dx, dy;                       /* TimeLayer2D members: (FTYPE)grid->dx (AdiSolver2D.cpp:30-35) */
	double gdx, gdy;                 /* Grid2D::dx, dy (double) */
	FT v_T, v_vis, t_vis, t_phi;     /* Common/Geometry.h:538-562 */
	double startT;                   /* Grid2D::startT */
	int *type, *bc;                  /* Grid2D::GetType, GetData().type */
	FT *gvx, *gvy, *gT;              /* GetData().vel, .T */
	FT *f[6][3];                     /* layers x (u, v, t), index i * dimy + j (TimeLayer2D.h:27-40) */
	Seg2 *listX, *listY;             /* listX: one segment per i, direction Y; listY: per j, direction X (AdiSolver2D.cpp:228-277) */
	int nX, nY;
	int iters;
} O2;

#define ID(o, i, j) ((i) * (o)->dimy + (j))
#define U_(l) o->f[l][0]
#define V_(l) o->f[l][1]
#define T_(l) o->f[l][2]

void *FN(oracle2d_create)(int dimx, int dimy, double dx, double dy, double v_T, double v_vis, double t_vis, double t_phi, double startT)
{
	O2 *o = (O2 *)calloc(1, sizeof(O2));
	const size_t N = (size_t)dimx * dimy;
	o->dimx = dimx; o->dimy = dimy; o->gdx = dx; o->gdy = dy; o->dx = (FT)dx; o->dy = (FT)dy;
	o->v_T = (FT)v_T; o->v_vis = (FT)v_vis; o->t_vis = (FT)t_vis; o->t_phi = (FT)t_phi; o->startT = startT;
	o->type = (int *)calloc(N, sizeof(int)); o->bc = (int *)calloc(N, sizeof(int));
	o->gvx = (FT *)calloc(N, sizeof(FT)); o->gvy = (FT *)calloc(N, sizeof(FT)); o->gT = (FT *)calloc(N, sizeof(FT));
	/* the reference allocates half / next / temp / next_local uninitialised (TimeLayer2D.h:176-181); zero here, like
	 * the probe driver does for the cells the copy loops never touch */
	for (int l = 0; l < 6; l++) for (int q = 0; q < 3; q++) o->f[l][q] = (FT *)calloc(N, sizeof(FT));
	o->listX = (Seg2 *)calloc((size_t)dimx, sizeof(Seg2)); o->listY = (Seg2 *)calloc((size_t)dimy, sizeof(Seg2));
	return o;
}

void FN(oracle2d_destroy)(void *h)
{
	O2 *o = (O2 *)h;
	if (!o) return;
	free(o->type); free(o->bc); free(o->gvx); free(o->gvy); free(o->gT);
	for (int l = 0; l < 6; l++) for (int q = 0; q < 3; q++) free(o->f[l][q]);
	free(o->listX); free(o->listY); free(o);
}

/* what the solver reads through Grid2D::GetType / GetData (Grid2D.h:52-54) */
void FN(oracle2d_set_grid)(void *h, const int *type, const int *bc, const FT *vx, const FT *vy, const FT *T)
{
	O2 *o = (O2 *)h;
	const size_t N = (size_t)o->dimx * o->dimy;
	memcpy(o->type, type, N * sizeof(int)); memcpy(o->bc, bc, N * sizeof(int));
	memcpy(o->gvx, vx, N * sizeof(FT)); memcpy(o->gvy, vy, N * sizeof(FT)); memcpy(o->gT, T, N * sizeof(FT));
}

/* AdiSolver2D::Init, AdiSolver2D.cpp:36-50: cur <- grid data in every cell */
void FN(oracle2d_init_layer)(void *h)
{
	O2 *o = (O2 *)h;
	for (int i = 0; i < o->dimx; i++)
		for (int j = 0; j < o->dimy; j++) {
			U_(L_CUR)[ID(o, i, j)] = o->gvx[ID(o, i, j)];
			V_(L_CUR)[ID(o, i, j)] = o->gvy[ID(o, i, j)];
			T_(L_CUR)[ID(o, i, j)] = o->gT[ID(o, i, j)];
		}
}

FT *FN(oracle2d_field)(void *h, int layer, int var) { return ((O2 *)h)->f[layer][var]; }
int FN(oracle2d_iters)(void *h) { return ((O2 *)h)->iters; }

/* TimeLayer2D::CopyUto / CopyVto / CopyTto (TimeLayer2D.h:104-158): i < dimx-1, j < dimy-1 only */
static void copy_type(O2 *o, int src, int dst, int type)
{
	for (int q = 0; q < 3; q++)
		for (int i = 0; i < o->dimx - 1; i++)
			for (int j = 0; j < o->dimy - 1; j++)
				if (o->type[ID(o, i, j)] == type) o->f[dst][q][ID(o, i, j)] = o->f[src][q][ID(o, i, j)];
}
/* MergeAllto (TimeLayer2D.h:112-166): dest = (dest + src) / 2 */
static void merge_type(O2 *o, int src, int dst, int type)
{
	for (int q = 0; q < 3; q++)
		for (int i = 0; i < o->dimx - 1; i++)
			for (int j = 0; j < o->dimy - 1; j++)
				if (o->type[ID(o, i, j)] == type) o->f[dst][q][ID(o, i, j)] = (o->f[dst][q][ID(o, i, j)] + o->f[src][q][ID(o, i, j)]) / 2;
}
/* CopyAllto(grid, dest) (TimeLayer2D.h:168-174): IN, OUT, BOUND, VALVE in this order */
static void copy_all(O2 *o, int src, int dst)
{
	copy_type(o, src, dst, NODE_IN); copy_type(o, src, dst, NODE_OUT); copy_type(o, src, dst, NODE_BOUND); copy_type(o, src, dst, NODE_VALVE);
}

/* TimeLayer2D::EvalDivError (TimeLayer2D.h:88-102): FTYPE accumulation, double quotient */
static double eval_div_error(O2 *o, int l)
{
	FT err = 0.0;
	int count = 0;
	const FT dx = o->dx, dy = o->dy;
	const FT *U = U_(l), *V = V_(l);
	for (int i = 0; i < o->dimx - 1; i++)
		for (int j = 0; j < o->dimy - 1; j++)
			if (o->type[ID(o, i, j)] == NODE_IN && o->type[ID(o, i + 1, j)] == NODE_IN && o->type[ID(o, i, j + 1)] == NODE_IN && o->type[ID(o, i + 1, j + 1)] == NODE_IN) {
				FT tx = dy * (U[ID(o, i + 1, j)] - U[ID(o, i, j)]) + (U[ID(o, i + 1, j + 1)] - U[ID(o, i, j + 1)]) / 2;
				FT ty = dx * (V[ID(o, i, j + 1)] - V[ID(o, i, j)]) + (V[ID(o, i + 1, j + 1)] - V[ID(o, i + 1, j)]) / 2;
				FT s = tx + ty;
				err += s < 0 ? -s : s;        /* abs(FTYPE): the <cmath> overload */
				count++;
			}
	return err / count;
}
double FN(oracle2d_eval_div_error)(void *h, int layer) { return eval_div_error((O2 *)h, layer); }

/* AdiSolver2D::CreateSegments (AdiSolver2D.cpp:228-277): one segment per row / column, first..last non-OUT run */
static void create_segments(O2 *o)
{
	const int dimx = o->dimx, dimy = o->dimy;
	o->nX = 0;
	for (int i = 0; i < dimx; i++) {
		Seg2 s; s.posx = i; s.dir = D_Y;
		int j = 0;
		while (j < dimy && o->type[ID(o, i, j)] == NODE_OUT) j++;
		while (j + 1 < dimy && o->type[ID(o, i, j + 1)] != NODE_IN) j++;
		if (j + 1 >= dimy) continue;
		s.posy = j;
		j = dimy - 1;
		while (j >= 0 && o->type[ID(o, i, j)] == NODE_OUT) j--;
		while (j - 1 >= 0 && o->type[ID(o, i, j - 1)] != NODE_IN) j--;
		s.size = j - s.posy + 1; s.endx = i; s.endy = j;
		o->listX[o->nX++] = s;
	}
	o->nY = 0;
	for (int j = 0; j < dimy; j++) {
		Seg2 s; s.posy = j; s.dir = D_X;
		int i = 0;
		while (i < dimx && o->type[ID(o, i, j)] == NODE_OUT) i++;
		while (i + 1 < dimx && o->type[ID(o, i + 1, j)] != NODE_IN) i++;
		if (i + 1 >= dimx) continue;
		s.posx = i;
		i = dimx - 1;
		while (i >= 0 && o->type[ID(o, i, j)] == NODE_OUT) i--;
		while (i - 1 >= 0 && o->type[ID(o, i - 1, j)] != NODE_IN) i--;
		s.size = i - s.posx + 1; s.endx = i; s.endy = j;
		o->listY[o->nY++] = s;
	}
}
int FN(oracle2d_num_segments)(void *h, int which) { O2 *o = (O2 *)h; create_segments(o); return which == 0 ? o->nX : o->nY; }

/* Common::SolveTridiagonal, Common/Algorithms.h:21-38 */
static void solve_tridiagonal(FT *a, FT *b, FT *c, FT *d, FT *x, int num)
{
	c[num - 1] = 0.0;
	c[0] = c[0] / b[0];
	d[0] = d[0] / b[0];
	for (int i = 1; i < num; i++) {
		c[i] = c[i] / (b[i] - a[i] * c[i - 1]);
		d[i] = (d[i] - d[i - 1] * a[i]) / (b[i] - a[i] * c[i - 1]);
	}
	x[num - 1] = d[num - 1];
	for (int i = num - 2; i >= 0; i--) x[i] = d[i] - c[i] * x[i + 1];
}

/* central differences of TimeLayer2D (TimeLayer2D.h:43-60) and the dissipation functions (:64-84) */
#define DXF(F, i, j) ((F[ID(o, (i) + 1, j)] - F[ID(o, (i) - 1, j)]) / (2 * o->dx))
#define DYF(F, i, j) ((F[ID(o, i, (j) + 1)] - F[ID(o, i, (j) - 1)]) / (2 * o->dy))
static FT diss_x(O2 *o, int l, int i, int j)
{
	FT ux = DXF(U_(l), i, j), vx = DXF(V_(l), i, j), uy = DYF(U_(l), i, j);
	return 2 * ux * ux + vx * vx + uy * vx;
}
static FT diss_y(O2 *o, int l, int i, int j)
{
	FT vx = DXF(V_(l), i, j), uy = DYF(U_(l), i, j), vy = DYF(V_(l), i, j);
	return uy * uy + 2 * vy * vy + vx * uy;
}

/* AdiSolver2D::SolveSegment (AdiSolver2D.cpp:180-203) = ApplyBC0 (:74-95), BuildMatrix (:118-178), ApplyBC1 (:97-116),
 * SolveTridiagonal, UpdateSegment (:52-72).  temp_local is the linearisation layer, next_local receives the solution. */
static void solve_segment(O2 *o, FT dt, const Seg2 *seg, int var, int dir, int cur, FT *a, FT *b, FT *c, FT *d, FT *x)
{
	const int n = seg->size, i = seg->posx, j = seg->posy;
	const int tl = L_TEMP_LOCAL;
	const FT gdx = (FT)o->gdx, gdy = (FT)o->gdy;
	/* ApplyBC0 */
	{
		const int id = ID(o, seg->posx, seg->posy);
		if (o->bc[id] == BC_NOSLIP) { b[0] = 1.0; c[0] = 0.0; d[0] = var == T_U ? o->gvx[id] : var == T_V ? o->gvy[id] : o->gT[id]; }
		else if (o->bc[id] == BC_FREE) { b[0] = 1.0; c[0] = -1.0; d[0] = 0.0; }
	}
	/* BuildMatrix */
	{
		const FT v_vis_dx2 = (FT)o->v_vis / (gdx * gdx), t_vis_dx2 = (FT)o->t_vis / (gdx * gdx);
		const FT v_vis_dy2 = (FT)o->v_vis / (gdy * gdy), t_vis_dy2 = (FT)o->t_vis / (gdy * gdy);
		for (int p = 1; p < n - 1; p++) {
			if (dir == D_X) {
				const FT vel = U_(tl)[ID(o, i + p, j)];
				const FT vis = var == T_T ? t_vis_dx2 : v_vis_dx2;
				a[p] = -vel / (2 * gdx) - vis;
				b[p] = 1 / dt + 2 * vis;
				c[p] = vel / (2 * gdx) - vis;
				if (var == T_U) d[p] = U_(cur)[ID(o, i + p, j)] / dt - o->v_T * DXF(T_(tl), i + p, j);
				else if (var == T_V) d[p] = V_(cur)[ID(o, i + p, j)] / dt;
				else d[p] = T_(cur)[ID(o, i + p, j)] / dt + o->t_phi * diss_x(o, tl, i + p, j);
			} else {
				const FT vel = V_(tl)[ID(o, i, j + p)];
				const FT vis = var == T_T ? t_vis_dy2 : v_vis_dy2;
				a[p] = -vel / (2 * gdy) - vis;
				b[p] = 1 / dt + 2 * vis;
				c[p] = vel / (2 * gdy) - vis;
				if (var == T_U) d[p] = U_(cur)[ID(o, i, j + p)] / dt;
				else if (var == T_V) d[p] = V_(cur)[ID(o, i, j + p)] / dt - o->v_T * DYF(T_(tl), i, j + p);
				else d[p] = T_(cur)[ID(o, i, j + p)] / dt + o->t_phi * diss_y(o, tl, i, j + p);
			}
		}
	}
	/* ApplyBC1 */
	{
		const int id = ID(o, seg->endx, seg->endy);
		if (o->bc[id] == BC_NOSLIP) { a[n - 1] = 0.0; b[n - 1] = 1.0; d[n - 1] = var == T_U ? o->gvx[id] : var == T_V ? o->gvy[id] : o->gT[id]; }
		else if (o->bc[id] == BC_FREE) { a[n - 1] = 1.0; b[n - 1] = -1.0; d[n - 1] = 0.0; }
	}
	solve_tridiagonal(a, b, c, d, x, n);
	/* UpdateSegment: all n cells, boundary cells included */
	{
		int ii = i, jj = j;
		for (int t = 0; t < n; t++) {
			o->f[L_NEXT_LOCAL][var][ID(o, ii, jj)] = x[t];
			if (dir == D_X) ii++; else jj++;
		}
	}
}

/* AdiSolver2D::SolveDirection (AdiSolver2D.cpp:205-226) */
static void solve_direction(O2 *o, FT dt, int num_local, const Seg2 *list, int nlist, int cur, int temp, int next, FT *w)
{
	const size_t N = (size_t)o->dimx * o->dimy;
	const int maxn = o->dimx > o->dimy ? o->dimx : o->dimy;
	FT *a = w, *b = w + maxn, *c = w + 2 * maxn, *d = w + 3 * maxn, *x = w + 4 * maxn;
	/* temp_local = new TimeLayer2D (uninitialised in the reference; zero here) ; temp->CopyAllto(grid, temp_local) */
	for (int q = 0; q < 3; q++) memset(o->f[L_TEMP_LOCAL][q], 0, N * sizeof(FT));
	copy_all(o, temp, L_TEMP_LOCAL);
	if (nlist == 0) return;
	const int dir = list[0].dir;
	for (int it = 0; it < num_local; it++) {
		for (int s = 0; s < nlist; s++) {
			solve_segment(o, dt, &list[s], T_U, dir, cur, a, b, c, d, x);
			solve_segment(o, dt, &list[s], T_V, dir, cur, a, b, c, d, x);
			solve_segment(o, dt, &list[s], T_T, dir, cur, a, b, c, d, x);
		}
		if (it == 0) copy_type(o, L_NEXT_LOCAL, L_TEMP_LOCAL, NODE_IN);
		else merge_type(o, L_NEXT_LOCAL, L_TEMP_LOCAL, NODE_IN);
	}
	copy_type(o, L_TEMP_LOCAL, temp, NODE_IN);
	copy_type(o, L_NEXT_LOCAL, next, NODE_IN);
}

/* Solver2D::UpdateBoundaries (Solver2D.cpp:48-62) */
void FN(oracle2d_update_boundaries)(void *h)
{
	O2 *o = (O2 *)h;
	for (int i = 0; i < o->dimx; i++)
		for (int j = 0; j < o->dimy; j++) {
			const int id = ID(o, i, j);
			if (o->type[id] == NODE_BOUND || o->type[id] == NODE_VALVE) {
				U_(L_CUR)[id] = o->gvx[id]; V_(L_CUR)[id] = o->gvy[id]; T_(L_CUR)[id] = o->gT[id];
			}
		}
	copy_type(o, L_CUR, L_NEXT, NODE_BOUND);
	copy_type(o, L_CUR, L_NEXT, NODE_VALVE);
}

/* AdiSolver2D::TimeStep (AdiSolver2D.cpp:279-323).  Returns 0, 1 (exceeded MAX_GLOBAL_ITERS: the reference exits) or
 * 2 ("Error is too big!": the reference exits). */
int FN(oracle2d_time_step)(void *h, double dt_in, int num_global, int num_local, double *err_out)
{
	O2 *o = (O2 *)h;
	const FT dt = (FT)dt_in;                              /* FluidSolver2D.cpp:131 */
	const int maxn = o->dimx > o->dimy ? o->dimx : o->dimy;
	FT *w = (FT *)malloc(sizeof(FT) * 5 * (size_t)maxn);
	int rc = 0;
	create_segments(o);
	copy_all(o, L_CUR, L_NEXT);
	copy_all(o, L_CUR, L_HALF);
	copy_all(o, L_CUR, L_TEMP);
	int it;
	double err = eval_div_error(o, L_NEXT);
	for (it = 0; (it < num_global) || (err > ERR_THRESHOLD); it++) {
		solve_direction(o, dt, num_local, o->listY, o->nY, L_CUR, L_TEMP, L_HALF, w);
		solve_direction(o, dt, num_local, o->listX, o->nX, L_HALF, L_TEMP, L_NEXT, w);
		err = eval_div_error(o, L_NEXT);
		if (it == 0) copy_type(o, L_NEXT, L_TEMP, NODE_IN);
		else merge_type(o, L_NEXT, L_TEMP, NODE_IN);
		if (it > MAX_GLOBAL_ITERS) { rc = 1; break; }
		if (err > ERR_THRESHOLD * 10) { rc = 2; break; }
	}
	o->iters = it;
	/* Solver2D::ClearOutterCells (Solver2D.cpp:73-84) */
	for (int i = 0; i < o->dimx; i++)
		for (int j = 0; j < o->dimy; j++)
			if (o->type[ID(o, i, j)] == NODE_OUT) {
				U_(L_NEXT)[ID(o, i, j)] = 0.0; V_(L_NEXT)[ID(o, i, j)] = 0.0; T_(L_NEXT)[ID(o, i, j)] = (FT)o->startT;
			}
	copy_all(o, L_NEXT, L_CUR);
	if (err_out) *err_out = err;
	free(w);
	return rc;
}

/* Solver2D::GetLayer (Solver2D.cpp:20-34): nearest-lower downsample of `next` */
void FN(oracle2d_get_layer)(void *h, FT *vel_xy, double *T, int outdimx, int outdimy)
{
	O2 *o = (O2 *)h;
	if (outdimx == 0) outdimx = o->dimx;
	if (outdimy == 0) outdimy = o->dimy;
	for (int i = 0; i < outdimx; i++)
		for (int j = 0; j < outdimy; j++) {
			const int x = i * o->dimx / outdimx, y = j * o->dimy / outdimy;
			vel_xy[2 * (i * outdimy + j)] = U_(L_NEXT)[ID(o, x, y)];
			vel_xy[2 * (i * outdimy + j) + 1] = V_(L_NEXT)[ID(o, x, y)];
			T[i * outdimy + j] = T_(L_NEXT)[ID(o, x, y)];
		}
}
