"""The density-map oracle (oracle/density_oracle.c) against the arithmetic the reference uses: SciPy's uniform_filter in
float32 (`CityModel._update_density_map`, city_model.py:1764-1778), bit for bit; and against the live method itself."""
import numpy as np
import pytest

from oracle import oracle as O


def scipy_density(occupancy_map, is_road_map, r=10):
    """The body of CityModel._update_density_map, with the reference's own third-party call."""
    from scipy.ndimage import uniform_filter
    occ = occupancy_map.astype(np.float32)
    sum_occ = uniform_filter(occ, size=(2 * r + 1, 2 * r + 1), mode='constant', cval=0.0) * ((2 * r + 1) ** 2)
    road = is_road_map.astype(np.float32)
    sum_road = uniform_filter(road, size=(2 * r + 1, 2 * r + 1), mode='constant', cval=0.0) * ((2 * r + 1) ** 2)
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.where(sum_road > 0, sum_occ / sum_road, 0.0)


@pytest.mark.parametrize("seed,shape,p_road,p_occ", [(1, (200, 200), 0.3, 0.1), (2, (64, 150), 0.6, 0.5), (3, (21, 21), 1.0, 0.0),
                                                     (4, (5, 90), 0.2, 1.0), (5, (130, 7), 0.05, 0.3), (6, (300, 257), 0.32, 0.02)])
def test_density_oracle_is_bit_exact_with_scipy(seed, shape, p_road, p_occ):
    rng = np.random.default_rng(seed)
    road = (rng.random(shape) < p_road).astype(np.int8)
    occ = ((rng.random(shape) < p_occ) & (road == 1)).astype(np.int8)
    want = scipy_density(occ, road)
    got = O.density_map(occ, road)
    assert want.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), want.astype(np.float32).view(np.uint32))
    assert (got >= 0).all() and (got <= 1.0001).all()


@pytest.mark.reference
def test_density_oracle_matches_the_live_method():
    from oracle.refharness import stubs
    if hasattr(stubs, "install"):
        stubs.install()
    from Simulation.city_model import CityModel

    class Model:
        pass
    rng = np.random.default_rng(9)
    m = Model()
    m.is_road_map = (rng.random((120, 160)) < 0.3).astype(np.int8)
    m.occupancy_map = ((rng.random((120, 160)) < 0.2) & (m.is_road_map == 1)).astype(np.int8)
    CityModel._update_density_map(m)
    got = O.density_map(m.occupancy_map, m.is_road_map)
    assert np.array_equal(got.view(np.uint32), np.asarray(m.density_map, np.float32).view(np.uint32))
    assert np.array_equal(np.asarray(m.density_map, np.float64), got.astype(np.float64))   # what the planner reads
