// TEST INFRASTRUCTURE (oracle/): probe driver around the UNMODIFIED reference CPU solver.
//
// This file is ours; it is compiled together with the reference's own sources
// (read in place from /root/reference/src by oracle/build_ref.sh, never copied into
// the repo) into oracle/_ref/ref_probe3d_{f32,f64}.  It mirrors the reference driver
// src/FluidSolver3D/FluidSolver3D.cpp:53-286 (construct grid -> Init -> CreateSegments ->
// loop {UpdateBoundaries; TimeStep; GetLayer}) but, instead of writing NetCDF, dumps the
// raw Node[] array and the raw time layers so tests can compare them with the oracle
// restatement (oracle/adi3d_oracle.c) and with the CUDA path.  Only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs may execute the resulting binary.
//
// usage: ref_probe3d <data> <config> <out.bin|-> <nsteps> [align] [dump=every|last|none|list:<s0,s1,..>]
//                    [stats=<stride>] [getlayer] [dt=<v>] [sweep=<Z|Y|X>] [threads=<n>] [solver=cpu|refgpu|b200|b200exact]
//                    [gpus=<n>]
// stats=<stride>: the steps selected by dump= are written as compact records (kind 5: per-field sum, sum of squares,
// sum of |.|, and the strided subsample [::stride, ::stride, ::stride]) instead of whole layers - what the golden
// vectors of the BASELINE-sized configs (256^3, 512^3) hold.
//
// Built with -DWITH_B200 (oracle/_ref/dropin3d_*) the same driver can put the reference's loader and Node[] in
// front of the B200 solver through the Solver3D adapter (cmc_fluid_solver_b200/host/B200AdiSolver3D.*): that binary
// is the drop-in demonstration - reference Grid3D/Grid2D/Config + our C ABI - and is exercised by the GPU tests.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <stdexcept>
#include <typeinfo>
#include <sstream>
#include <fstream>
#include <iostream>
#include <algorithm>
#include <map>
#include <omp.h>
#include <cuda_runtime.h>

// the probe needs the solver's private layers (temp/half) and SolveDirection for
// component-level vectors; widening access does not change any layout or code.
#define private public
#define protected public
#include "FluidSolver3D.h"
#undef private
#undef protected

#ifdef WITH_B200
#include "B200AdiSolver3D.h"
#endif

using namespace FluidSolver3D;
using namespace Common;

static FILE *g_out = NULL;

static void put(const void *p, size_t bytes)
{
	if (g_out && fwrite(p, 1, bytes, g_out) != bytes) { perror("fwrite"); exit(2); }
}
static void put_i32(int v) { put(&v, 4); }
static void put_f64(double v) { put(&v, 8); }

static void put_field(ScalarField3D *f, size_t n)
{
	// the initial `cur` layer carries a dimy*dimz halo even on CPU (AdiSolver3D.cpp:254) while
	// half/next/temp have none (:251-258): always address cell (0,0,0) through elem()
	put(&f->elem(0, 0, 0), n * sizeof(FTYPE));
}

// kind 5 record of one layer given as four dense arrays (u, v, w, T)
static void put_stats(int step, double err, const FTYPE *const f[4], int dimx, int dimy, int dimz, int stride)
{
	put_i32(step); put_i32(5); put_f64(err);
	const int sx = (dimx + stride - 1) / stride, sy = (dimy + stride - 1) / stride, sz = (dimz + stride - 1) / stride;
	put_i32(stride); put_i32(sx); put_i32(sy); put_i32(sz);
	std::vector<FTYPE> smp((size_t)sx * sy * sz);
	for (int q = 0; q < 4; q++) {
		double s1 = 0.0, s2 = 0.0, sa = 0.0;
		const size_t n = (size_t)dimx * dimy * dimz;
		for (size_t id = 0; id < n; id++) { const double v = (double)f[q][id]; s1 += v; s2 += v * v; sa += fabs(v); }
		put_f64(s1); put_f64(s2); put_f64(sa);
		size_t o = 0;
		for (int i = 0; i < dimx; i += stride)
			for (int j = 0; j < dimy; j += stride)
				for (int k = 0; k < dimz; k += stride) smp[o++] = f[q][((size_t)i * dimy + j) * dimz + k];
		put(smp.data(), smp.size() * sizeof(FTYPE));
	}
}

static bool step_selected(const std::string &dump, const std::vector<int> &list, int i, int nsteps)
{
	if (dump == "every") return true;
	if (dump == "last") return i == nsteps - 1;
	if (dump == "list") return std::find(list.begin(), list.end(), i) != list.end();
	return false;
}

static void put_layer(int step, int kind, double err, TimeLayer3D *L, size_t n)
{
	put_i32(step); put_i32(kind); put_f64(err);
	put_field(L->U, n); put_field(L->V, n); put_field(L->W, n); put_field(L->T, n);
}

int main(int argc, char **argv)
{
	if (argc < 5) {
		fprintf(stderr, "usage: %s <data> <config> <out.bin|-> <nsteps> [align] [dump=every|last|none] [getlayer] [dt=v] [sweep=Z|Y|X] [threads=n]\n", argv[0]);
		return 1;
	}
	bool align = false, getlayer = false;
	std::string dump = "last", sweep = "", which = "cpu";
	double dt_override = -1;
	int nsteps = atoi(argv[4]);
	int stats_stride = 0, ngpus = 1;
	std::vector<int> dump_list;
	for (int a = 5; a < argc; a++) {
		if (!strcmp(argv[a], "align")) align = true;
		else if (!strcmp(argv[a], "getlayer")) getlayer = true;
		else if (!strncmp(argv[a], "dump=list:", 10)) {
			dump = "list";
			std::stringstream ss(argv[a] + 10);
			std::string tok;
			while (std::getline(ss, tok, ',')) dump_list.push_back(atoi(tok.c_str()));
		}
		else if (!strncmp(argv[a], "dump=", 5)) dump = argv[a] + 5;
		else if (!strncmp(argv[a], "stats=", 6)) stats_stride = atoi(argv[a] + 6);
		else if (!strncmp(argv[a], "gpus=", 5)) ngpus = atoi(argv[a] + 5);
		else if (!strncmp(argv[a], "dt=", 3)) dt_override = atof(argv[a] + 3);
		else if (!strncmp(argv[a], "sweep=", 6)) sweep = argv[a] + 6;
		else if (!strncmp(argv[a], "threads=", 8)) omp_set_num_threads(atoi(argv[a] + 8));
		else if (!strncmp(argv[a], "solver=", 7)) which = argv[a] + 7;
	}
	try {
		// solver=refgpu: the reference's OWN CUDA backend (AdiSolver3D.cu / TimeLayer3D.cu, compiled unmodified for the GPU at
		// hand) - timing and residual only; the second baseline SURVEY.md offers.  gpus=<n> = the reference CLI's "GPU <n>"
		const BackendType be = (which == "refgpu") ? GPU : CPU;
		PARAplan *pplan = PARAplan::Instance();
		pplan->init(be);
		if (be == GPU) pplan->setGPUnum(ngpus);
		Config();
		Config::LoadFromFile(argv[2]);

		Grid3D *grid = NULL;
		if (Config::in_fmt == Shape3D)
			grid = new Grid3D(Config::dx, Config::dy, Config::dz, Config::baseT, be, false, EVEN_X);
		else if (Config::in_fmt == Shape2D)
			grid = new Grid3D(Config::dx, Config::dy, Config::dz, Config::depth, Config::depth_var, Config::baseT, be, false, EVEN_X);
		else
			throw std::runtime_error("probe: SeaNetCDF input needs libnetcdf (not available)");
		grid->SetFrameTime(Config::frame_time);
		grid->SetBoundParams(Config::bc_inV, Config::bc_inT);
		if (!grid->LoadFromFile(argv[1], align)) throw std::runtime_error("probe: cannot load grid");
		grid->Prepare_CPU(0.0);
		grid->Split();
		grid->Init_GPU();

		const int dimx = grid->dimx, dimy = grid->dimy, dimz = grid->dimz;
		const size_t N = (size_t)dimx * dimy * dimz;
		size_t n_in = 0;
		for (size_t id = 0; id < N; id++) if (grid->GetNodesCPU()[id].type == NODE_IN) n_in++;

		FluidParams *params;
		if (Config::useNormalizedParams) params = new FluidParams(Config::Re, Config::Pr, Config::lambda);
		else params = new FluidParams(Config::viscosity, Config::density, Config::R_specific, Config::k, Config::cv);

		AdiSolver3D *solver = NULL;
#ifdef WITH_B200
		B200AdiSolver3D *b200 = NULL;
		if (which != "cpu" && which != "refgpu") {
			b200 = new B200AdiSolver3D(which == "b200exact" ? CMC_MODE_EXACT : CMC_MODE_FAST, 0, ngpus);      // gpus=<n>: the reference CLI's "GPU <n>"
			b200->Init(GPU, false, grid, *params, false, 1);
		}
#else
		if (which != "cpu" && which != "refgpu") throw std::runtime_error("probe: built without the B200 adapter");
#endif
		if (which == "cpu" || which == "refgpu") {
			solver = new AdiSolver3D();
			if (be == GPU) solver->SetOptionsGPU(false, false);       // the driver's defaults (FluidSolver3D.cpp:63-64,187)
			solver->Init(be, false, grid, *params, false, 1);
		}

		int frames = grid->GetFramesNum();
		double length = grid->GetCycleLength();
		double dt = length / (frames * Config::time_steps);
		if (dt_override > 0) dt = dt_override;

		printf("probe: %s precision, grid %d x %d x %d, NODE_IN %zu, dt %.17g, num_global %d, num_local %d, threads %d\n",
			(typeid(FTYPE) == typeid(float)) ? "single" : "double", dimx, dimy, dimz, n_in, dt,
			Config::num_global, Config::num_local, omp_get_max_threads());

		if (strcmp(argv[3], "-")) {
			g_out = fopen(argv[3], "wb");
			if (!g_out) throw std::runtime_error("probe: cannot open output");
		}
		// ---- header + nodes -------------------------------------------------------------
		put("CMCPROBE", 8);
		put_i32(1); put_i32((int)sizeof(FTYPE));
		put_i32(dimx); put_i32(dimy); put_i32(dimz);
		put_i32(Config::num_global); put_i32(Config::num_local); put_i32(nsteps);
		put_i32(Config::outdimx); put_i32(Config::outdimy); put_i32(Config::outdimz);
		put_i32(0);
		put_f64(grid->dx); put_f64(grid->dy); put_f64(grid->dz); put_f64((double)(FTYPE)dt);
		put_f64(params->v_T); put_f64(params->v_vis); put_f64(params->t_vis); put_f64(params->t_phi);
		put_f64(grid->baseT);
		{
			Node *nodes = grid->GetNodesCPU();
			std::vector<int> ti(N);
			std::vector<FTYPE> tf(N);
			for (size_t id = 0; id < N; id++) ti[id] = (int)nodes[id].type;    put(ti.data(), N * 4);
			for (size_t id = 0; id < N; id++) ti[id] = (int)nodes[id].bc_vel;  put(ti.data(), N * 4);
			for (size_t id = 0; id < N; id++) ti[id] = (int)nodes[id].bc_temp; put(ti.data(), N * 4);
			for (size_t id = 0; id < N; id++) tf[id] = nodes[id].v.x; put(tf.data(), N * sizeof(FTYPE));
			for (size_t id = 0; id < N; id++) tf[id] = nodes[id].v.y; put(tf.data(), N * sizeof(FTYPE));
			for (size_t id = 0; id < N; id++) tf[id] = nodes[id].v.z; put(tf.data(), N * sizeof(FTYPE));
			for (size_t id = 0; id < N; id++) tf[id] = nodes[id].T;   put(tf.data(), N * sizeof(FTYPE));
		}

#ifdef WITH_B200
		if (b200) {
			// the drop-in path: reference loader + Node[] -> Solver3D adapter -> C ABI -> sm_100a kernels
			b200->CreateSegments();
			grid->Prepare(0);
			size_t outN = (size_t)Config::outdimx * Config::outdimy * Config::outdimz;
			Vec3D *resVel = getlayer ? new Vec3D[outN] : NULL;
			double *resT = getlayer ? new double[outN] : NULL;
			std::vector<FTYPE> buf(N);
			double t_steps = 0.0;
			for (int i = 0; i < nsteps; i++) {
				bool computeError = (i % 10 == 0) || (i == nsteps - 1);
				double t0 = omp_get_wtime();
				b200->UpdateBoundaries();
				b200->TimeStep((FTYPE)dt, Config::num_global, Config::num_local, computeError);
				t_steps += omp_get_wtime() - t0;
				if (getlayer && (i % Config::out_time_steps) == 0) {
					b200->GetLayer(resVel, resT, Config::outdimx, Config::outdimy, Config::outdimz);
					put_i32(i); put_i32(1); put_f64(b200->GetError());
					put(resVel, outN * sizeof(Vec3D));
					put(resT, outN * sizeof(double));
				}
				if (step_selected(dump, dump_list, i, nsteps)) {
					if (stats_stride > 0) {
						std::vector<FTYPE> all(4 * N);
						const FTYPE *f[4];
						for (int q = 0; q < 4; q++) { b200->ReadField(CMC_LAYER_CUR, q, all.data() + q * N); f[q] = all.data() + q * N; }
						put_stats(i, b200->GetError(), f, dimx, dimy, dimz, stats_stride);
					} else {
						put_i32(i); put_i32(0); put_f64(b200->GetError());
						for (int q = 0; q < 4; q++) { b200->ReadField(CMC_LAYER_CUR, q, buf.data()); put(buf.data(), N * sizeof(FTYPE)); }
					}
				}
			}
			printf("\nprobe: b200 steps %d, seconds %.6f, err %.10g\n", nsteps, t_steps, b200->GetError());
			if (g_out) fclose(g_out);
			delete b200;
			fflush(stdout);
			_Exit(0);
		}
#endif
		// half/next/temp are allocated uninitialised by the reference (TimeLayer3D.h:353).
		// Define their initial contents (= cur) so dumps are deterministic, OUT cells included.
		if (be == CPU) {
			solver->cur->CopyLayerTo(solver->next);
			solver->cur->CopyLayerTo(solver->half);
			solver->cur->CopyLayerTo(solver->temp);
		}

		solver->CreateSegments();
		grid->Prepare(0);
		printf("probe: segments X %d Y %d Z %d\n", solver->numSegs[X], solver->numSegs[Y], solver->numSegs[Z]);
		if (be == GPU) {
			// timing + residual only (the layers live on the device)
			double t_steps = 0.0;
			for (int i = 0; i < nsteps; i++) {
				bool computeError = (i % 10 == 0) || (i == nsteps - 1);
				cudaDeviceSynchronize();
				double t0 = omp_get_wtime();
				solver->UpdateBoundaries();
				solver->TimeStep((FTYPE)dt, Config::num_global, Config::num_local, computeError);
				cudaDeviceSynchronize();
				const double t1 = omp_get_wtime();
				t_steps += t1 - t0;
				printf("\nprobe: refgpu step %d seconds %.6f err %.10g\n", i, t1 - t0, solver->diffError);
			}
			printf("\nprobe: refgpu steps %d, seconds %.6f, sec_per_step %.6f, mcells_per_s %.6f, err %.10g\n",
				nsteps, t_steps, nsteps ? t_steps / nsteps : 0.0,
				(nsteps && t_steps > 0) ? (double)N * nsteps / t_steps / 1e6 : 0.0, solver->diffError);
			if (g_out) fclose(g_out);
			fflush(stdout);
			_Exit(0);
		}

		if (!sweep.empty()) {
			// component vector: the TimeStep prologue (AdiSolver3D.cpp:310-320) followed by ONE
			// SolveDirection call; dumps cur (kind 2), next (kind 3) and temp (kind 4).
			solver->UpdateBoundaries();
			solver->cur->CopyLayerTo(grid, solver->next, NODE_BOUND);
			solver->cur->CopyLayerTo(grid, solver->next, NODE_VALVE);
			solver->cur->CopyLayerTo(solver->temp);
			DirType d = (sweep == "X") ? X : (sweep == "Y") ? Y : Z;
			Segment3D *list = (d == X) ? solver->h_listX : (d == Y) ? solver->h_listY : solver->h_listZ;
			solver->SolveDirection(d, (FTYPE)dt, Config::num_local, list, NULL, NULL, solver->cur, solver->temp, solver->next);
			put_layer(0, 2, 0.0, solver->cur, N);
			put_layer(0, 3, 0.0, solver->next, N);
			put_layer(0, 4, 0.0, solver->temp, N);
			if (g_out) fclose(g_out);
			return 0;
		}

		size_t outN = (size_t)Config::outdimx * Config::outdimy * Config::outdimz;
		Vec3D *resVel = getlayer ? new Vec3D[outN] : NULL;
		double *resT = getlayer ? new double[outN] : NULL;

		double t_steps = 0.0;
		for (int i = 0; i < nsteps; i++) {
			bool computeError = (i % 10 == 0) || (i == nsteps - 1);
			double t0 = omp_get_wtime();
			solver->UpdateBoundaries();
			solver->TimeStep((FTYPE)dt, Config::num_global, Config::num_local, computeError);
			const double t1 = omp_get_wtime();
			t_steps += t1 - t0;
			printf("\nprobe: step %d seconds %.6f\n", i, t1 - t0);
			if (getlayer && (i % Config::out_time_steps) == 0) {
				solver->GetLayer(resVel, resT, Config::outdimx, Config::outdimy, Config::outdimz);
				put_i32(i); put_i32(1); put_f64(solver->diffError);
				put(resVel, outN * sizeof(Vec3D));
				put(resT, outN * sizeof(double));
			}
			if (step_selected(dump, dump_list, i, nsteps)) {
				if (stats_stride > 0) {
					const FTYPE *f[4] = {&solver->cur->U->elem(0, 0, 0), &solver->cur->V->elem(0, 0, 0), &solver->cur->W->elem(0, 0, 0), &solver->cur->T->elem(0, 0, 0)};
					put_stats(i, solver->diffError, f, dimx, dimy, dimz, stats_stride);
				} else
					put_layer(i, 0, solver->diffError, solver->cur, N);
			}
		}
		printf("\nprobe: steps %d, seconds %.6f, sec_per_step %.6f, mcells_per_s %.6f, err %.10g\n",
			nsteps, t_steps, nsteps ? t_steps / nsteps : 0.0,
			(nsteps && t_steps > 0) ? (double)N * nsteps / t_steps / 1e6 : 0.0, solver->diffError);
		if (g_out) fclose(g_out);
		fflush(stdout);
		_Exit(0);   // skip the reference's profiler table / destructors
	}
	catch (std::exception &e) {
		fprintf(stderr, "probe: exception: %s\n", e.what());
		return 3;
	}
	return 0;
}
