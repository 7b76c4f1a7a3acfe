// rows.cuh - coefficient build of one tridiagonal row, boundary conditions and obstacle mask
// folded in.  Restates AdiSolver3D::BuildMatrix / ApplyBC0 / ApplyBC1 (reference
// src/FluidSolver3D/AdiSolver3D.cpp:732-852) and the finite-difference helpers
// ScalarField3D::d_x/d_y/d_z, TimeLayer3D::DissFuncX/Y/Z (TimeLayer3D.h:338-340, 554-588).
// Expression order follows the source; a translation unit compiled with -fmad=false therefore
// reproduces the reference CPU arithmetic bit for bit.
#pragma once
#include "common.cuh"

namespace cmc {

template <typename FT>
struct RowConst {
	FT two_h;        // 2 * h_D
	FT vis_v, vis_T; // v_vis / (h*h), t_vis / (h*h)
	FT b_v, b_T;     // 3/dt + 2*vis
	FT dt;
	FT two_hx, two_hy, two_hz;
	FT v_T, t_phi;
	__device__ __forceinline__ void init(const SweepArgs<FT> &A, int dir)
	{
		const FT h = A.h[dir];
		two_h = 2 * h;
		vis_v = A.v_vis / (h * h);
		vis_T = A.t_vis / (h * h);
		dt = A.dt;
		b_v = 3 / dt + 2 * vis_v;
		b_T = 3 / dt + 2 * vis_T;
		two_hx = 2 * A.h[0]; two_hy = 2 * A.h[1]; two_hz = 2 * A.h[2];
		v_T = A.v_T; t_phi = A.t_phi;
	}
};

// Interior row at cell `id` of a sweep in direction DIR.
//   a = -V/(2h) - vis ; c = V/(2h) - vis ; b = 3/dt + 2 vis ; d = cur*3/dt (+ gradient / dissipation terms)
// sx, sz: element strides of +1 in x, z; syp / sym: distance to the next / previous j-row (they differ at the edges
// of a y-block, see Layout).
template <typename FT, int DIR>
__device__ __forceinline__ void build_interior_row(const SweepArgs<FT> &A, const RowConst<FT> &K, long long id,
                                                   long long sx, long long syp, long long sym, long long sz,
                                                   FT &a_v, FT &c_v, FT &a_T, FT &c_T, FT d[4])
{
	const FT *tu = A.temp[0], *tv = A.temp[1], *tw = A.temp[2], *tT = A.temp[3];
	const FT V = A.temp[DIR][id];
	a_v = -V / K.two_h - K.vis_v;
	c_v = V / K.two_h - K.vis_v;
	a_T = -V / K.two_h - K.vis_T;
	c_T = V / K.two_h - K.vis_T;
	const FT du = A.cur[0][id] * 3 / K.dt;
	const FT dv = A.cur[1][id] * 3 / K.dt;
	const FT dw = A.cur[2][id] * 3 / K.dt;
	const FT dT = A.cur[3][id] * 3 / K.dt;
	if (DIR == 0) {
		const FT T_x = (tT[id + sx] - tT[id - sx]) / K.two_hx;
		const FT u_x = (tu[id + sx] - tu[id - sx]) / K.two_hx;
		const FT v_x = (tv[id + sx] - tv[id - sx]) / K.two_hx;
		const FT w_x = (tw[id + sx] - tw[id - sx]) / K.two_hx;
		const FT u_y = (tu[id + syp] - tu[id - sym]) / K.two_hy;
		const FT u_z = (tu[id + sz] - tu[id - sz]) / K.two_hz;
		d[0] = du - K.v_T * T_x;
		d[1] = dv;
		d[2] = dw;
		d[3] = dT + K.t_phi * (2 * u_x * u_x + v_x * v_x + w_x * w_x + v_x * u_y + w_x * u_z);
	} else if (DIR == 1) {
		const FT T_y = (tT[id + syp] - tT[id - sym]) / K.two_hy;
		const FT u_y = (tu[id + syp] - tu[id - sym]) / K.two_hy;
		const FT v_y = (tv[id + syp] - tv[id - sym]) / K.two_hy;
		const FT w_y = (tw[id + syp] - tw[id - sym]) / K.two_hy;
		const FT v_x = (tv[id + sx] - tv[id - sx]) / K.two_hx;
		const FT v_z = (tv[id + sz] - tv[id - sz]) / K.two_hz;
		d[0] = du;
		d[1] = dv - K.v_T * T_y;
		d[2] = dw;
		d[3] = dT + K.t_phi * (u_y * u_y + 2 * v_y * v_y + w_y * w_y + u_y * v_x + w_y * v_z);
	} else {
		const FT T_z = (tT[id + sz] - tT[id - sz]) / K.two_hz;
		const FT u_z = (tu[id + sz] - tu[id - sz]) / K.two_hz;
		const FT v_z = (tv[id + sz] - tv[id - sz]) / K.two_hz;
		const FT w_z = (tw[id + sz] - tw[id - sz]) / K.two_hz;
		const FT w_x = (tw[id + sx] - tw[id - sx]) / K.two_hx;
		const FT w_y = (tw[id + syp] - tw[id - sym]) / K.two_hy;
		d[0] = du;
		d[1] = dv;
		d[2] = dw - K.v_T * T_z;
		d[3] = dT + K.t_phi * (u_z * u_z + v_z * v_z + 2 * w_z * w_z + u_z * w_x + v_z * w_y);
	}
}

// Boundary rows.  ApplyBC0 (first cell): FREE -> (b0, c0, d0) = (2, -1, 0); NOSLIP -> (1, 0, value).
//                 ApplyBC1 (last cell):  FREE -> (a1, b1, d1) = (-1, 2, 0); NOSLIP -> (0, 1, value).
// `off` is the off-diagonal entry (c0 for a start row, a1 for an end row).
template <typename FT>
__device__ __forceinline__ void boundary_row(const SweepArgs<FT> &A, unsigned role, long long id,
                                             FT &off_v, FT &b_v, FT &off_T, FT &b_T, FT d[4])
{
	if (role & R_VFREE) {
		off_v = FT(-1.0); b_v = FT(2.0);
		d[0] = FT(0.0); d[1] = FT(0.0); d[2] = FT(0.0);
	} else {
		off_v = FT(0.0); b_v = FT(1.0);
		d[0] = A.nodev[0][id]; d[1] = A.nodev[1][id]; d[2] = A.nodev[2][id];
	}
	if (role & R_TFREE) {
		off_T = FT(-1.0); b_T = FT(2.0); d[3] = FT(0.0);
	} else {
		off_T = FT(0.0); b_T = FT(1.0); d[3] = A.nodev[3][id];
	}
}

} // namespace cmc
