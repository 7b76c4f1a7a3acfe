// B200AdiSolver2D.cpp - see B200AdiSolver2D.h.  Error behaviour mirrors AdiSolver2D::TimeStep
// (src/FluidSolver2D/AdiSolver2D.cpp:303-313): where the reference prints and exits, so does the adapter.
#include "B200AdiSolver2D.h"

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace FluidSolver2D
{
	B200AdiSolver2D::B200AdiSolver2D(int _device) : h(NULL), device(_device), iters(0), err(0.0), type(NULL), bc(NULL), gvx(NULL), gvy(NULL), gT(NULL)
	{
		grid = NULL;
		cur = NULL;
		next = NULL;
	}

	B200AdiSolver2D::~B200AdiSolver2D()
	{
		if (h) cmc_adi2d_destroy(h);
		delete cur; delete next;
		delete [] type; delete [] bc; delete [] gvx; delete [] gvy; delete [] gT;
	}

	void B200AdiSolver2D::check(int rc, const char *what)
	{
		if (rc == CMC_OK) return;
		if (rc == CMC_ERR_DIVERGED) {                 // AdiSolver2D.cpp:303-313: message + exit(1)
			printf("\n%s\n", cmc_last_error());
			exit(1);
		}
		throw std::runtime_error(std::string(what) + ": " + cmc_last_error());
	}

	void B200AdiSolver2D::Init(Grid2D *_grid, FluidParams &_params)
	{
		grid = _grid;
		dimx = grid->dimx; dimy = grid->dimy;
		params = _params;
		const int N = dimx * dimy;
		type = new int[N]; bc = new int[N];
		gvx = new FTYPE[N]; gvy = new FTYPE[N]; gT = new FTYPE[N];
		// host mirrors of the two layers the non-virtual Solver2D members work on (AdiSolver2D.cpp:30-50)
		cur = new TimeLayer2D(dimx, dimy, (FTYPE)grid->dx, (FTYPE)grid->dy);
		next = new TimeLayer2D(dimx, dimy, (FTYPE)grid->dx, (FTYPE)grid->dy);
		for (int i = 0; i < dimx; i++)
			for (int j = 0; j < dimy; j++) {
				cur->U(i, j) = grid->GetData(i, j).vel.x; cur->V(i, j) = grid->GetData(i, j).vel.y; cur->T(i, j) = grid->GetData(i, j).T;
				next->U(i, j) = 0; next->V(i, j) = 0; next->T(i, j) = 0;      // uninitialised in the reference
			}
		cmc_fluid_params p = { params.v_T, params.v_vis, params.t_vis, params.t_phi };
		check(cmc_adi2d_create(dimx, dimy, grid->dx, grid->dy, &p, grid->startT, (int)sizeof(FTYPE), device, &h), "cmc_adi2d_create");
	}

	void B200AdiSolver2D::TimeStep(FTYPE dt, int num_global, int num_local)
	{
		// what the solver may read of the grid (the driver refreshed it: grid.Prepare(t), FluidSolver2D.cpp:129)
		for (int i = 0; i < dimx; i++)
			for (int j = 0; j < dimy; j++) {
				const int id = i * dimy + j;
				CondData2D d = grid->GetData(i, j);
				type[id] = (int)grid->GetType(i, j); bc[id] = (int)d.type;
				gvx[id] = d.vel.x; gvy[id] = d.vel.y; gT[id] = d.T;
			}
		// one round trip: grid arrays + both host layers up (Solver2D::UpdateBoundaries has just edited them on the host), the
		// step on the device (half / temp / next_local stay resident), both layers down
		void *c[3] = { &cur->U(0, 0), &cur->V(0, 0), &cur->T(0, 0) }, *n[3] = { &next->U(0, 0), &next->V(0, 0), &next->T(0, 0) };
		check(cmc_adi2d_step_host(h, type, bc, gvx, gvy, gT, c, n, (double)dt, num_global, num_local, &err, &iters), "cmc_adi2d_step_host");
		printf("\rerr = %.4f,", err);              // AdiSolver2D.cpp:318
	}
}
