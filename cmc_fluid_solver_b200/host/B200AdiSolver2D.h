// B200AdiSolver2D.h - the adapter a maintainer of the reference adds next to AdiSolver2D: a Solver2D subclass
// (reference src/FluidSolver2D/Solver2D.h:24-45) whose TimeStep runs on the GPU through the C ABI of include/cmc_adi.h
// (cmc_adi2d_*).  Compiled AGAINST the reference's own headers; contains no solver arithmetic.
//
//   Solver2D *solver = new B200AdiSolver2D();      // instead of new AdiSolver2D()  (FluidSolver2D.cpp:71)
//   solver->Init(&grid, params);                   // everything else of the driver loop is unchanged:
//   loop: grid.Prepare(t); solver->UpdateBoundaries(); solver->TimeStep(dt, num_global, num_local);
//         solver->SetGridBoundaries(); solver->GetLayer(...)
//
// Solver2D's UpdateBoundaries / SetGridBoundaries / GetLayer are NOT virtual and work on the host layers `cur` and
// `next`, so the adapter keeps those two as host mirrors: TimeStep uploads them (UpdateBoundaries has just edited them),
// steps on the device, and downloads them again.  half / temp / next_local live on the device only.  No edit of the
// reference is needed.
#pragma once
#include "Solver2D.h"
#include "cmc_adi.h"

namespace FluidSolver2D
{
	class B200AdiSolver2D : public Solver2D
	{
	public:
		explicit B200AdiSolver2D(int device = 0);
		~B200AdiSolver2D();

		void Init(Grid2D *grid, FluidParams &params);
		void TimeStep(FTYPE dt, int num_global, int num_local);

		double GetError() const { return err; }
		int GetIterations() const { return iters; }

	private:
		cmc_adi2d *h;
		int device, iters;
		double err;
		int *type, *bc;
		FTYPE *gvx, *gvy, *gT;
		void check(int rc, const char *what);
	};
}
