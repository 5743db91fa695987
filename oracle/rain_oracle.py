"""CPU restatement of the rain map (TEST INFRASTRUCTURE ONLY): RainAgent's precomputed offsets (agents/rain.py:44-50) and
step (:61-72: covered cells = offsets around (int(x), int(y)) that exist in cell_lookup, i.e. lie on the grid), and
RainManager.step (:154-185: clear the previous tick's cells, set the union of the clouds' covered cells)."""
import numpy as np


def offsets(radius):
    return [(dx, dy) for dx in range(-radius, radius + 1) for dy in range(-radius, radius + 1) if dx * dx + dy * dy <= radius * radius]


class RainOracle:
    def __init__(self, W, H):
        self.W, self.H = W, H
        self.rain_map = np.zeros((H, W), np.uint8)
        self._prev = set()

    def step(self, clouds):
        for x, y in self._prev:
            self.rain_map[y, x] = 0
        new = set()
        for cx, cy, r in clouds:
            cx, cy = int(cx), int(cy)
            for dx, dy in offsets(int(r)):
                xi, yi = cx + dx, cy + dy
                if 0 <= xi < self.W and 0 <= yi < self.H:
                    new.add((xi, yi))
        for x, y in new:
            self.rain_map[y, x] = 1
        self._prev = new
        return self.rain_map
