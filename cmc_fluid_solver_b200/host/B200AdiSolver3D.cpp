// B200AdiSolver3D.cpp - see B200AdiSolver3D.h.  Error behaviour mirrors AdiSolver3D: failures become
// std::runtime_error (the reference's gpuSafeCall does the same, src/Common/GPUplan.cpp:173-193); a residual above
// ERR_THRESHOLD prints "Error is too big!" and throws std::runtime_error("") (AdiSolver3D.cpp:371-374).
#include "B200AdiSolver3D.h"

#include <cstdio>
#include <stdexcept>
#include <string>

namespace FluidSolver3D
{
	B200AdiSolver3D::B200AdiSolver3D(int _mode, int _device, int _num_gpus) : h(NULL), mode(_mode), device(_device), num_gpus(_num_gpus), diffError(0.0)
	{
		grid = NULL;
		cur = NULL;      // the time layers live on the device behind the C ABI
		next = NULL;
	}

	B200AdiSolver3D::~B200AdiSolver3D()
	{
		if (h) cmc_adi3d_destroy(h);
	}

	void B200AdiSolver3D::check(int rc, const char *what)
	{
		if (rc == CMC_OK) return;
		if (rc == CMC_ERR_DIVERGED) {
			printf("\nError is too big! %f\n", diffError);
			throw std::runtime_error("");
		}
		throw std::runtime_error(std::string(what) + ": " + cmc_last_error());
	}

	void B200AdiSolver3D::Init(BackendType, bool, Grid3D *_grid, FluidParams &_params, bool, int)
	{
		grid = _grid;
		dimx = grid->dimx; dimy = grid->dimy; dimz = grid->dimz;
		params = _params;
		cmc_grid_desc g = { dimx, dimy, dimz, grid->dx, grid->dy, grid->dz };
		cmc_fluid_params p = { params.v_T, params.v_vis, params.t_vis, params.t_phi };
		if (num_gpus > 1) {
			// "GPU <n>": one host thread, n devices (GPUplan::init, src/Common/GPUplan.cpp:35-77)
			int devices[16];
			const int n = num_gpus > 16 ? 16 : num_gpus;
			for (int i = 0; i < n; i++) devices[i] = device + i;
			check(cmc_adi3d_create_multi(&g, &p, (int)sizeof(FTYPE), devices, n, &h), "cmc_adi3d_create_multi");
			printf("B200AdiSolver3D: %d devices, x-slabs of %d planes\n", n, dimx / n);
		} else
			check(cmc_adi3d_create(&g, &p, (int)sizeof(FTYPE), device, &h), "cmc_adi3d_create");
		check(cmc_adi3d_set_option(h, "mode", mode), "cmc_adi3d_set_option");
		// the reference's Node[] goes across the ABI as it is (Grid3D.h:73-88)
		check(cmc_adi3d_set_nodes_aos(h, grid->GetNodesCPU(), sizeof(Node)), "cmc_adi3d_set_nodes_aos");
	}

	void B200AdiSolver3D::CreateSegments() { check(cmc_adi3d_build_lines(h), "cmc_adi3d_build_lines"); }

	void B200AdiSolver3D::UpdateBoundaries() { check(cmc_adi3d_update_boundaries(h), "cmc_adi3d_update_boundaries"); }

	void B200AdiSolver3D::TimeStep(FTYPE dt, int num_global, int num_local, bool computeError)
	{
		int rc = cmc_adi3d_time_step(h, (double)dt, num_global, num_local, computeError ? 1 : 0, &diffError);
		check(rc, "cmc_adi3d_time_step");
		printf("\rerr = %.8f,", diffError);      // AdiSolver3D.cpp:378
		fflush(stdout);
	}

	void B200AdiSolver3D::GetLayer(Vec3D *v, double *T, int outdimx, int outdimy, int outdimz)
	{
		check(cmc_adi3d_get_layer(h, v, T, outdimx, outdimy, outdimz), "cmc_adi3d_get_layer");   // Vec3D = 3 x FTYPE
	}

	void B200AdiSolver3D::ReadField(int layer, int var, FTYPE *dst) { check(cmc_adi3d_read_field(h, layer, var, dst), "cmc_adi3d_read_field"); }

	double B200AdiSolver3D::sum_layer(char) { return 0.0; }   // the reference's body is commented out too (AdiSolver3D.cpp:51-57)
	void B200AdiSolver3D::debug(bool) {}
}
