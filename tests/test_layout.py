"""Host-side check of the y-blocked storage layout (cmc_fluid_solver_b200/csrc/common.cuh, struct Layout): compiles a
small host program against the header and verifies that idx() is a bijection onto the buffer, that one block is the
reference's plain [i][j][k] order (TimeLayer3D.h:256-259, plus guard planes) and that jup / jdn are the distances to the
neighbouring j-rows, for single-block and blocked shapes including a ragged last block.  CPU only (nvcc, no GPU)."""
import shutil
import subprocess

import pytest

from conftest import ROOT

SRC = r'''
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace cmc;
static int check(int nx, int ny, int nz, int jbs)
{
	Layout L{};
	const int nzp = (nz + 15) / 16 * 16;
	L.x0 = 0; L.gx = nx;
	L.shape(nx, ny, nz, nzp, jbs);
	std::vector<char> seen((size_t)L.total, 0);
	for (int i = -1; i <= nx; i++)
		for (int j = 0; j < ny; j++)
			for (int k = 0; k < nzp; k++) {
				const long long id = L.idx(i, j, k);
				if (id < 0 || id >= L.total || seen[(size_t)id]) { printf("collision nx %d ny %d nz %d jbs %d at %d %d %d\n", nx, ny, nz, jbs, i, j, k); return 1; }
				seen[(size_t)id] = 1;
				if (L.nblk == 1 && id != (long long)(i + 1) * ny * nzp + (long long)j * nzp + k) { printf("one block is not [i][j][k]\n"); return 1; }
			}
	for (int j = 0; j < ny; j++) {
		if (j + 1 < ny && L.idx(3, j + 1, 5) - L.idx(3, j, 5) != L.jup(j)) { printf("jup wrong at %d (jbs %d)\n", j, jbs); return 1; }
		if (j > 0 && L.idx(3, j, 5) - L.idx(3, j - 1, 5) != L.jdn(j)) { printf("jdn wrong at %d (jbs %d)\n", j, jbs); return 1; }
		// rows outside the grid: the distance must still land inside the buffer for every plane of the slab
		if (L.idx(0, j, 0) - L.jdn(j) < 0 || L.idx(nx - 1, j, nzp - 1) + L.jup(j) >= L.total + L.plane) { printf("neighbour row leaves the buffer at %d\n", j); return 1; }
	}
	if (L.idx(1, 0, 0) - L.idx(0, 0, 0) != L.plane) { printf("plane stride\n"); return 1; }
	return 0;
}
int main()
{
	int bad = 0;
	const int shapes[][3] = {{8, 37, 19}, {16, 64, 32}, {5, 130, 17}, {24, 40, 48}};
	for (auto &s : shapes)
		for (int jbs : {30, 3, 4, 5, 6}) bad += check(s[0], s[1], s[2], jbs);
	printf(bad ? "FAILED\n" : "OK\n");
	return bad;
}
'''


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_layout_indexing(tmp_path):
    (tmp_path / "layout_check.cu").write_text(SRC)
    exe = tmp_path / "layout_check"
    r = subprocess.run(["nvcc", "-std=c++17", "-O1", f"-I{ROOT / 'cmc_fluid_solver_b200' / 'csrc'}", "-o", str(exe), str(tmp_path / "layout_check.cu")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-2000:]
