"""Rain map on the device (tsim_rain_discs through GpuRain) vs the restatement of RainAgent / RainManager (oracle/rain_oracle.py):
clouds drifting across the grid edge, overlapping clouds, a cloud that leaves the map, radius 0."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_rain_map_follows_the_clouds():
    from oracle.rain_oracle import RainOracle
    from trafficsimulation_b200.rain import GpuRain
    W, H = 200, 160
    gpu, ora = GpuRain(W, H), RainOracle(W, H)
    rng = np.random.default_rng(5)
    clouds = [[-3.0, 20.5, 12, 1.0, 0.4], [150.2, 158.0, 25, -0.7, -0.7], [100.0, 80.0, 0, 0.3, 0.9], [60.5, 70.5, 18, 0.9, -0.1], [70.0, 75.0, 18, 0.8, 0.0]]
    for t in range(120):
        for c in clouds:
            c[0] += c[3]; c[1] += c[4]
        live = [(c[0], c[1], c[2]) for c in clouds if t < 90 or c[2] != 25]   # one cloud is removed on the way
        got = gpu.step(live).cpu().numpy()
        want = ora.step(live)
        assert np.array_equal(got, want), t
    assert want.sum() > 0


def test_rain_offsets_are_the_reference_offsets():
    """The disc a kernel row spans is exactly the reference's offset list (agents/rain.py:44-50)."""
    from oracle.rain_oracle import offsets
    from trafficsimulation_b200.rain import GpuRain
    for r in (0, 1, 2, 5, 13, 40):
        gpu = GpuRain(2 * r + 9, 2 * r + 7)
        got = gpu.step([(r + 4.9, r + 3.2, r)]).cpu().numpy()
        want = np.zeros_like(got)
        for dx, dy in offsets(r):
            want[r + 3 + dy, r + 4 + dx] = 1
        assert np.array_equal(got, want), r
