"""oracle/rain_oracle.py against the LIVE reference: RainAgent's own offsets and covered cells (agents/rain.py:44-72)."""
import numpy as np
import pytest

pytestmark = pytest.mark.reference


def test_rain_oracle_equals_reference_rain_agent():
    import types
    from oracle.refharness import harness as Hn
    from oracle.rain_oracle import RainOracle, offsets
    Hn.load_reference()
    from Simulation.agents.rain import RainAgent
    W, H = 60, 50
    lookup = {(x, y): (x, y) for x in range(W) for y in range(H)}
    model = types.SimpleNamespace(cell_lookup=lookup, get_width=lambda: W, get_height=lambda: H, rains=[], schedule=types.SimpleNamespace(remove=lambda a: None),
                                  next_id=lambda: 0, random=None)
    ora = RainOracle(W, H)
    for pos, direction in (((3.5, 4.25), (1.0, 0.5)), ((55.0, 45.0), (-1.0, -0.2))):
        ag = RainAgent("Rain_0", model, pos, direction)
        assert sorted(ag._offsets) == sorted(offsets(ag.radius))
        for _ in range(30):
            ag.step()
            want = np.zeros((H, W), np.uint8)
            for (x, y) in ag.covered_cells:
                want[y, x] = 1
            got = ora.step([(ag.x, ag.y, ag.radius)])
            assert np.array_equal(got, want)
