// TEST INFRASTRUCTURE (oracle/): probe driver for the UNMODIFIED reference 2D ADI solver (BASELINE config 1: "2D case
// from data/2D on the reference CPU solver").  Compiled by oracle/build_ref.sh against a scratch copy of
// /root/reference/src whose only edits are the Windows include paths of the 2D sources.
//
// Follows the reference driver's time loop (src/FluidSolver2D/FluidSolver2D.cpp:53-155: Config, Grid2D load, Prepare,
// per step Prepare(t) / UpdateBoundaries / TimeStep / SetGridBoundaries, GetLayer every out_time_steps) with the
// AdiSolver2D backend, and dumps what crosses the Solver2D interface (src/FluidSolver2D/Solver2D.h:24-45):
//   per step: the grid arrays the solver reads through Grid2D::GetType / GetData (after Prepare(t)), then the
//   solver's current layer and the divergence residual after TimeStep, and every GetLayer output.
// Nothing here is part of the product; only tests/ and bench.py's cpu_baseline leg may execute the binary.
//
// usage: ref_probe2d <data> <config> <out.bin|-> <nsteps|0=all> [dump=every|last|none] [solver=cpu|b200]
//
// Built with -DWITH_B200 (oracle/_ref/dropin2d_f32) the same driver puts the reference's loader and Grid2D in front of
// the B200 solver through the Solver2D adapter (cmc_fluid_solver_b200/host/B200AdiSolver2D.*): the 2D drop-in demonstration.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <sstream>
#include <iostream>
#include <fstream>
#include <algorithm>
#include <map>
#include <list>
#include <stdexcept>
#include <omp.h>

#define private public
#define protected public
#include "AdiSolver2D.h"
#undef private
#undef protected
#include "../Common/Config.h"
#ifdef WITH_B200
#include "B200AdiSolver2D.h"
#endif

using namespace FluidSolver2D;
using namespace Common;

static FILE *g_out = NULL;
static void put(const void *p, size_t n) { if (g_out && fwrite(p, 1, n, g_out) != n) { perror("fwrite"); exit(2); } }
static void put_i32(int v) { put(&v, 4); }
static void put_f64(double v) { put(&v, 8); }

static void dump_grid(Grid2D &grid, int step)
{
	const int N = grid.dimx * grid.dimy;
	std::vector<int> ty(N), bc(N);
	std::vector<FTYPE> vx(N), vy(N), T(N);
	for (int i = 0; i < grid.dimx; i++)
		for (int j = 0; j < grid.dimy; j++) {
			const int id = i * grid.dimy + j;
			CondData2D d = grid.GetData(i, j);
			ty[id] = (int)grid.GetType(i, j); bc[id] = (int)d.type;
			vx[id] = d.vel.x; vy[id] = d.vel.y; T[id] = d.T;
		}
	put_i32(step); put_i32(2);
	put(ty.data(), N * 4); put(bc.data(), N * 4);
	put(vx.data(), N * sizeof(FTYPE)); put(vy.data(), N * sizeof(FTYPE)); put(T.data(), N * sizeof(FTYPE));
}

int main(int argc, char **argv)
{
	if (argc < 5) { fprintf(stderr, "usage: %s <data> <config> <out.bin|-> <nsteps|0> [dump=every|last|none]\n", argv[0]); return 1; }
	std::string dump = "every", which = "cpu";
	int nsteps = atoi(argv[4]);
	for (int a = 5; a < argc; a++)
		if (!strncmp(argv[a], "dump=", 5)) dump = argv[a] + 5;
		else if (!strncmp(argv[a], "solver=", 7)) which = argv[a] + 7;

	Config();
	Config::LoadFromFile(argv[2]);
	if (Config::solverID != ADI) { fprintf(stderr, "probe2d: the config must say `solver ADI` (the hot path)\n"); return 1; }

	Grid2D grid(Config::dx, Config::dy, Config::baseT, Config::bc_noslip, Config::bc_strength);
	char field[4] = "";
	if (!grid.LoadFromFile(argv[1], field)) return 1;
	grid.Prepare(0, 0);
	FluidParams params(Config::viscosity, Config::density, Config::R_specific, Config::k, Config::cv);

	Solver2D *solver = NULL;
	if (which == "cpu") {
		AdiSolver2D *adi = new AdiSolver2D();
		adi->Init(&grid, params);
		// half / next / temp / next_local are allocated uninitialised by the reference (TimeLayer2D.h:176-181); define the
		// cells its copy loops never touch (last row / column) so that dumps are deterministic
		TimeLayer2D *ls[4] = {adi->half, adi->next, adi->temp, adi->next_local};
		for (int l = 0; l < 4; l++)
			for (int i = 0; i < grid.dimx; i++)
				for (int j = 0; j < grid.dimy; j++) { ls[l]->U(i, j) = 0; ls[l]->V(i, j) = 0; ls[l]->T(i, j) = 0; }
		solver = adi;
	} else {
#ifdef WITH_B200
		try {
			solver = new B200AdiSolver2D(0);       // driven through the Solver2D interface only from here on
			solver->Init(&grid, params);
		} catch (const std::exception &e) { fprintf(stderr, "probe2d: exception: %s\n", e.what()); return 3; }
#else
		fprintf(stderr, "probe2d: built without the B200 adapter\n");
		return 1;
#endif
	}

	const int frames = grid.GetFramesNum();
	const double length = grid.GetCycleLenght();
	const double dt = length / (frames * Config::time_steps);
	const double finaltime = length * Config::cycles;
	const int N = grid.dimx * grid.dimy;
	const int outN = Config::outdimx * Config::outdimy;
	printf("probe2d: grid %d x %d, dx %g dy %g, dt %.9g, frames %d, time_steps %d, num_global %d, num_local %d, FTYPE %d bytes\n",
	       grid.dimx, grid.dimy, Config::dx, Config::dy, dt, frames, Config::time_steps, Config::num_global, Config::num_local, (int)sizeof(FTYPE));

	if (strcmp(argv[3], "-")) { g_out = fopen(argv[3], "wb"); if (!g_out) { perror(argv[3]); return 1; } }
	put("CMCPRB2D", 8);
	int hdr[10] = {1, (int)sizeof(FTYPE), grid.dimx, grid.dimy, Config::num_global, Config::num_local, nsteps, Config::outdimx, Config::outdimy, Config::out_time_steps};
	put(hdr, sizeof hdr);
	double dh[9] = {grid.dx, grid.dy, dt, (double)params.v_T, (double)params.v_vis, (double)params.t_vis, (double)params.t_phi, grid.startT, 0.0};
	put(dh, sizeof dh);
	// initial layer (TimeLayer2D filled from the grid by AdiSolver2D::Init, AdiSolver2D.cpp:36-50)
	{
		std::vector<FTYPE> f(N);
		put_i32(-1); put_i32(0); put_f64(0.0); put_i32(0);
		for (int q = 0; q < 3; q++) {
			for (int i = 0; i < grid.dimx; i++)
				for (int j = 0; j < grid.dimy; j++) f[i * grid.dimy + j] = q == 0 ? solver->cur->U(i, j) : q == 1 ? solver->cur->V(i, j) : solver->cur->T(i, j);
			put(f.data(), N * sizeof(FTYPE));
		}
	}

	Vec2D *resVel = new Vec2D[outN];
	double *resT = new double[outN];
	std::vector<FTYPE> f(N);
	int lastframe = -1, done = 0;
	double t = dt, t_steps = 0.0;
	for (int i = 0; t < finaltime && (nsteps == 0 || done < nsteps); t += dt, i++, done++) {
		int currentframe = grid.GetFrame(t);
		if (currentframe != lastframe) { lastframe = currentframe; i = 0; }      // FluidSolver2D.cpp:100-127
		grid.Prepare(t);
		const bool rec = dump == "every" || (dump == "last" && (nsteps ? done == nsteps - 1 : t + dt >= finaltime));
		if (rec) dump_grid(grid, done);
		const double t0 = omp_get_wtime();
		solver->UpdateBoundaries();
		solver->TimeStep((FTYPE)dt, Config::num_global, Config::num_local);
		t_steps += omp_get_wtime() - t0;
		const double err = solver->next->EvalDivError(&grid);     // what TimeStep printed (NODE_IN cells only: unchanged by ClearOutterCells)
		if (rec) {
			put_i32(done); put_i32(0); put_f64(err); put_i32(0);
			for (int q = 0; q < 3; q++) {
				for (int a = 0; a < grid.dimx; a++)
					for (int b = 0; b < grid.dimy; b++) f[a * grid.dimy + b] = q == 0 ? solver->cur->U(a, b) : q == 1 ? solver->cur->V(a, b) : solver->cur->T(a, b);
				put(f.data(), N * sizeof(FTYPE));
			}
		}
		solver->SetGridBoundaries();
		if ((i % Config::out_time_steps) == 0) {
			solver->GetLayer(resVel, resT, Config::outdimx, Config::outdimy);
			if (rec) { put_i32(done); put_i32(1); put_f64(err); put_i32(0); put(resVel, outN * sizeof(Vec2D)); put(resT, outN * sizeof(double)); }
		}
	}
	printf("\nprobe2d: steps %d, seconds %.6f, sec_per_step %.6g, mcells_per_s %.6g\n", done, t_steps, t_steps / (done ? done : 1),
	       done ? (double)N * done / t_steps / 1e6 : 0.0);
	if (g_out) fclose(g_out);
	fflush(stdout);
	_Exit(0);
}
