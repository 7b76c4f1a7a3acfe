// B200AdiSolver3D.h - the adapter a maintainer of the reference adds next to AdiSolver3D: a Solver3D subclass
// (reference src/FluidSolver3D/Solver3D.h:24-49) that forwards every call to the C ABI of include/cmc_adi.h.
// This file is compiled AGAINST the reference's own headers (it includes Solver3D.h); it contains no solver logic.
//
//   Solver3D *solver = new B200AdiSolver3D(CMC_MODE_FAST, 0, pplan->gpuNum());   // instead of new AdiSolver3D()  (FluidSolver3D.cpp:184)
//   solver->Init(GPU, csv, grid, *params, false, 1);   // unchanged
//   static_cast<B200AdiSolver3D*>(solver)->CreateSegments();   // instead of dynamic_cast<AdiSolver3D*> (:224)
//   loop: solver->UpdateBoundaries(); solver->TimeStep(dt, num_global, num_local, computeError); ... GetLayer
//
// Solver3D::GetLayer is not virtual in the reference; a call through a Solver3D* needs the one-word change
// `virtual void GetLayer(...)` in Solver3D.h:33.  Calls on a B200AdiSolver3D* work as is.
#pragma once
#include "Solver3D.h"
#include "cmc_adi.h"

namespace FluidSolver3D
{
	class B200AdiSolver3D : public Solver3D
	{
	public:
		// num_gpus: what the reference's command line says after "GPU" (FluidSolver3D.cpp:88-95 -> PARAplan::setGPUnum):
		// one host thread drives devices [device, device + num_gpus) through cmc_adi3d_create_multi
		explicit B200AdiSolver3D(int mode = CMC_MODE_FAST, int device = 0, int num_gpus = 1);
		~B200AdiSolver3D();

		void Init(BackendType backend, bool csv, Grid3D *grid, FluidParams &params, bool useBlocking, int nblockZ);
		void CreateSegments();
		void UpdateBoundaries();
		void TimeStep(FTYPE dt, int num_global, int num_local, bool computeError);
		void GetLayer(Vec3D *v, double *T, int outdimx = 0, int outdimy = 0, int outdimz = 0);
		double sum_layer(char ch);
		void debug(bool ifdebug);

		// raw field of the current time layer (dense dimx*dimy*dimz FTYPE), for dumps and tests
		void ReadField(int layer, int var, FTYPE *dst);
		double GetError() const { return diffError; }

	private:
		cmc_adi3d *h;
		int mode, device, num_gpus;
		double diffError;
		void check(int rc, const char *what);
	};
}
