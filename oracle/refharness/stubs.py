"""Import stubs that let the UNMODIFIED reference (`/root/reference/Simulation`) be imported here.

TEST INFRASTRUCTURE ONLY.  Nothing in `trafficsimulation_b200/` may import this.

The reference needs `mesa`, `tensorflow` and `matplotlib` only to *import*
(SURVEY.md F6/F7, §8c); none of them is installed and there is no network.  These
stubs give the container semantics the hot path relies on:

* `mesa.Model`      – `self.random = random.Random(seed)`, `running`
* `mesa.Agent`      – `unique_id`, `model`, `pos`
* `mesa.space.MultiGrid` – `grid[x, y]` is a live list; place/remove/move; `coord_iter`
  yields `(list, (x, y))` x-outer / y-inner (what `city_model.py:1761` unpacks)
* `mesa.time.RandomActivation` – dict of agents keyed by `unique_id`; `step()` shuffles the
  keys with `model.random` and steps those still scheduled (Mesa 2.x behaviour).  The
  activation order is what the rank tape replaces (see `harness.TapeRandom`).
* `tensorflow`, `matplotlib` – permissive dummies, never executed on the default config.

Mesa's version is unpinned upstream (no requirements file), so the activation
shuffle is defined by this stub + the tape: "parity unpinned" for that one point, as
SURVEY.md §8c records.
"""
from __future__ import annotations

import random
import sys
import types

REFERENCE_ROOT = "/root/reference"


class _Model:
    def __init__(self, *args, seed=None, **kwargs):
        self._seed = seed
        self.random = random.Random(seed)
        self.running = True


class _Agent:
    def __init__(self, unique_id, model):
        self.unique_id = unique_id
        self.model = model
        self.pos = None

    def step(self):  # pragma: no cover - overridden
        pass


class _MultiGrid:
    def __init__(self, width, height, torus=False):
        self.width = width
        self.height = height
        self.torus = torus
        self._cells = [[[] for _ in range(height)] for _ in range(width)]

    def __getitem__(self, key):
        x, y = key
        return self._cells[x][y]

    def place_agent(self, agent, pos):
        x, y = pos
        self._cells[x][y].append(agent)
        agent.pos = pos

    def remove_agent(self, agent):
        pos = agent.pos
        if pos is not None:
            x, y = pos
            lst = self._cells[x][y]
            if agent in lst:
                lst.remove(agent)
        agent.pos = None

    def move_agent(self, agent, pos):
        self.remove_agent(agent)
        self.place_agent(agent, pos)

    def coord_iter(self):
        for x in range(self.width):
            for y in range(self.height):
                yield self._cells[x][y], (x, y)


class _RandomActivation:
    def __init__(self, model):
        self.model = model
        self.steps = 0
        self.time = 0
        self._agents = {}

    @property
    def agents(self):
        return list(self._agents.values())

    def add(self, agent):
        self._agents[agent.unique_id] = agent

    def remove(self, agent):
        self._agents.pop(agent.unique_id, None)

    def step(self):
        keys = list(self._agents.keys())
        self.model.random.shuffle(keys)
        for k in keys:
            ag = self._agents.get(k)
            if ag is not None:
                ag.step()
        self.steps += 1
        self.time += 1


class _Permissive(types.ModuleType):
    """A module whose every attribute is another permissive object."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        mod = sys.modules.get(full)
        if mod is None:
            mod = _Permissive(full)
            sys.modules[full] = mod
        setattr(self, name, mod)
        return mod

    def __call__(self, *a, **k):
        return self


def install():
    """Install the stubs into `sys.modules` and put the reference on `sys.path` (idempotent)."""
    if "mesa" not in sys.modules or not hasattr(sys.modules["mesa"], "_tsim_stub"):
        mesa = types.ModuleType("mesa")
        mesa.Model = _Model
        mesa.Agent = _Agent
        mesa._tsim_stub = True
        space = types.ModuleType("mesa.space")
        space.MultiGrid = _MultiGrid
        time_ = types.ModuleType("mesa.time")
        time_.RandomActivation = _RandomActivation
        mesa.space = space
        mesa.time = time_
        sys.modules["mesa"] = mesa
        sys.modules["mesa.space"] = space
        sys.modules["mesa.time"] = time_
    for name in ("tensorflow", "tensorflow.keras", "tensorflow.keras.layers",
                 "tensorflow.keras.optimizers", "matplotlib", "matplotlib.colors"):
        if name not in sys.modules:
            sys.modules[name] = _Permissive(name)
    # wire parents → children so `from tensorflow.keras import layers` works
    sys.modules["tensorflow"].keras = sys.modules["tensorflow.keras"]
    sys.modules["tensorflow.keras"].layers = sys.modules["tensorflow.keras.layers"]
    sys.modules["tensorflow.keras"].optimizers = sys.modules["tensorflow.keras.optimizers"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
