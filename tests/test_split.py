"""x-split policies of the slab decomposition (cmc_split_planes: Grid3D::Split / SplitSegments_X, reference
src/FluidSolver3D/Grid3D.cpp:148-235) - host logic, no GPU."""
import numpy as np
import pytest

from cmc_fluid_solver_b200.cases import channel_case
from cmc_fluid_solver_b200.solver import CmcError, SPLIT_EVEN_SEGMENTS, SPLIT_EVEN_VOLUME, SPLIT_EVEN_X, split_planes


@pytest.mark.parametrize("policy", [SPLIT_EVEN_X, SPLIT_EVEN_SEGMENTS, SPLIT_EVEN_VOLUME])
@pytest.mark.parametrize("dims,n", [((96, 40, 48), 4), ((69, 40, 44), 3), ((53, 53, 52), 3), ((512, 16, 16), 8), ((105, 24, 24), 5)])
def test_split_is_a_partition_on_multiples_of_eight(dims, n, policy):
    case = channel_case(*dims)
    p = split_planes(case, n, policy)
    assert len(p) == n and sum(p) == dims[0]
    assert all(v >= 8 for v in p) and all(v % 8 == 0 for v in p[:-1])


def test_even_volume_follows_the_fluid():
    """A grid whose fluid sits in the low-x half: EVEN_VOLUME gives the empty half fewer slabs' worth of planes than EVEN_X."""
    case = channel_case(128, 24, 24, baffle=False)
    t = case.type.reshape(case.shape)
    t[64:, :, :] = 1                                   # NODE_OUT
    even = split_planes(case, 4, SPLIT_EVEN_X)
    vol = split_planes(case, 4, SPLIT_EVEN_VOLUME)
    seg = split_planes(case, 4, SPLIT_EVEN_SEGMENTS)
    assert even == [32, 32, 32, 32]
    assert vol[-1] > 64 and sum(vol[:3]) <= 64 and seg[-1] > 64          # three slabs share the fluid half, the last takes the rest
    fluid = (t == 0).sum(axis=(1, 2)).astype(float)
    cuts = np.cumsum([0] + vol)
    loads = [fluid[a:b].sum() for a, b in zip(cuts[:-1], cuts[1:])]
    assert max(loads[:3]) <= 1.5 * fluid.sum() / 4 + fluid.max() * 8     # balanced up to the 8-plane rounding


def test_too_small_grids_are_refused():
    with pytest.raises(CmcError):
        split_planes(channel_case(20, 12, 12, baffle=False), 4, SPLIT_EVEN_VOLUME)
