"""A duck-typed stand-in for the reference's ``CityModel`` / ``CellAgent`` / ``Defaults`` with exactly the surface the adaptor
touches (city_model.py:96-115 trackers and maps, place_cell :1864-1870, get_cell_contents :1965-1972, grid.place_agent /
move_agent / remove_agent; cell.py:22-50 attributes).  It lets the seam of INTEGRATION.md run where the reference is absent
(the GPU box) and its result be read back into planes with the harness's own extraction rules."""
import types

import numpy as np

from trafficsimulation_b200 import encoding as E


def fake_defaults():
    roads = {"R1", "R2", "R3"}
    return types.SimpleNamespace(
        ZONES=list(E.ZONES), AVAILABLE_CITY_BLOCKS=list(E.AVAILABLE_CITY_BLOCKS),
        CITY_BLOCK_CHANCE={"Residential": 0.25, "Office": 0.25, "Market": 0.2, "Leisure": 0.2, "Other": 0.1, "Empty": 0.0},
        ROAD_LIKE_TYPES=roads | {"Intersection", "HighwayEntrance", "HighwayExit", "BlockEntrance"}, ROADS=roads,
        ZONE_COLORS={}, SUBBLOCK_CHANGE=0.3, BLOCK_ENTRANCE_ROAD_LEVEL=0)


class FakeCell:
    def __init__(self, cid, model, position, cell_type):
        self.id, self.position, self.cell_type = cid, position, cell_type
        self.road_type = cell_type if cell_type in ("R1", "R2", "R3") else None
        self.directions = []
        self.base_color = None
        self.block_id = self.block_type = self.highway_id = self.highway_orientation = None
        self.light = None
        self.assigned_incoming_road_blocks, self.assigned_outgoing_road_blocks, self.controlled_blocks = [], [], []


class FakeGrid:
    def __init__(self):
        self.where = {}

    def place_agent(self, ag, pos):
        self.where[id(ag)] = tuple(pos)
        ag.pos = tuple(pos)

    def move_agent(self, ag, pos):
        self.where[id(ag)] = tuple(pos)

    def remove_agent(self, ag):
        self.where.pop(id(ag), None)


class FakeModel:
    """Constructor kwargs = the reference's (city_model.py:27-53), defaults from a fixture's ``meta["cfg"]``."""

    def __init__(self, **cfg):
        d = dict(width=200, height=200, wall_thickness=15, sidewalk_ring_width=2, ring_road_type="R2", optimized_intersections=True,
                 carve_subblock_roads=False, subblock_roads_have_intersections=True, subblock_road_type="R3", min_subblock_spacing=5,
                 traffic_light_range=10, forward_traffic_light_range=False, forward_traffic_light_range_intersections="Skip")
        d.update({k: v for k, v in cfg.items() if k in d})
        for k, v in d.items():
            setattr(self, k, v)
        W, H = self.width, self.height
        self.cells = {}
        self.grid = FakeGrid()
        self.block_entrances, self.highway_entrances, self.highway_exits = [], [], []
        self.controlled_roads, self.traffic_lights, self._blocks_data = [], [], []
        self._intersection_cells, self._ring_road_cells, self._road_cells = set(), set(), set()
        self.occupancy_map = np.zeros((H, W), np.int8)
        self.stop_map = np.ones((H, W), np.int8)
        self.stuck_map = np.zeros((H, W), np.int8)
        self.active_vehicle_agents = []
        self.step_count = 0
        self.place_calls = 0

    def place_cell(self, x, y, new_type, new_id):
        self.cells[(x, y)] = FakeCell(new_id, self, (x, y), new_type)
        self.place_calls += 1

    def get_cell_contents(self, x, y):
        c = self.cells.get((x, y))
        return [c] if c is not None else []


def extract_planes(model):
    """Same rules as oracle/refharness/harness.py::extract_planes, on the fake model."""
    W, H = model.width, model.height
    T, D = np.zeros((H, W), np.uint8), np.zeros((H, W), np.uint16)
    A, B = np.zeros((H, W), np.uint8), np.zeros((H, W), np.int32)
    for (x, y), c in model.cells.items():
        T[y, x] = E.TYPE_CODE[c.cell_type]
        if c.directions:
            D[y, x] = E.encode_dirs(c.directions)
        a = E.TYPE_CODE[c.road_type] if c.cell_type == "ControlledRoad" else 0
        if c.light is not None:
            a |= 0x80
        A[y, x] = a
        if c.cell_type == "BlockEntrance" and c.block_id is not None:
            B[y, x] = c.block_id
    for (x, y) in model._ring_road_cells:
        A[y, x] |= 0x20
    for (x, y) in model._intersection_cells:
        A[y, x] |= 0x40
    for info in model._blocks_data:
        for (x, y) in info["region"]:
            B[y, x] = info["block_id"]
    return {"cell_type": T, "dirs": D, "aux": A, "block_id": B}


def extract_links(model):
    W = model.width
    idx = lambda c: c.position[1] * W + c.position[0]
    lights = np.array(sorted(idx(t) for t in model.traffic_lights), np.int32)
    ctrl = np.array(sorted((idx(t), idx(c)) for t in model.traffic_lights for c in t.controlled_blocks), np.int32).reshape(-1, 2)
    inc = np.array(sorted((idx(t), idx(c)) for t in model.traffic_lights for c in t.assigned_incoming_road_blocks), np.int32).reshape(-1, 2)
    out = np.array(sorted((idx(t), idx(c)) for t in model.traffic_lights for c in t.assigned_outgoing_road_blocks), np.int32).reshape(-1, 2)
    return {"lights": lights, "ctrl": ctrl, "incoming": inc, "outgoing": out}
