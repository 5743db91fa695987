"""Cell-type / direction encoding of the packed planes (must round-trip to the reference's strings).

cell_type u8  = index into ``Defaults.ZONES`` (reference Simulation/config.py:74-95)
dirs      u16 = bits0-3 mask (N=1, E=2, S=4, W=8: the ``allowed_dirs_map`` bits, city_model.py:2191-2196),
                bits 4+2i..5+2i = i-th entry of the ordered ``CellAgent.directions`` list, bits12-14 = len
aux       u8  = bits0-4 original type of a ControlledRoad (``CellAgent.road_type``), bit5 member of
                ``_ring_road_cells``, bit6 member of ``_intersection_cells``, bit7 ``light is not None``
"""
ZONES = [
    "Residential", "Office", "Market", "Leisure", "Other", "Empty", "Nothing", "Sidewalk", "Wall",
    "R1", "R2", "R3", "Intersection", "HighwayEntrance", "HighwayExit", "TrafficLight",
    "TrafficLightStop", "ControlledRoad", "ControlledRoadStop", "BlockEntrance",
]
TYPE_CODE = {z: i for i, z in enumerate(ZONES)}
DIR_NAMES = ["N", "E", "S", "W"]
DIR_INDEX = {d: i for i, d in enumerate(DIR_NAMES)}
ROAD_CODE = {None: 0, "R1": 1, "R2": 2, "R3": 3}
AVAILABLE_CITY_BLOCKS = ["Residential", "Office", "Market", "Leisure", "Other"]
FORWARD_MODES = ["Skip", "Include in Range", "Include as Extra"]

AUX_ORIG_MASK, AUX_RING, AUX_EVER_INT, AUX_HAS_LIGHT = 0x1F, 0x20, 0x40, 0x80


def encode_dirs(dirs) -> int:
    code = 0
    for i, d in enumerate(dirs):
        k = DIR_INDEX[d]
        code |= (1 << k) | (k << (4 + 2 * i))
    return code | (len(dirs) << 12)


def decode_dirs(code: int):
    return [DIR_NAMES[(code >> (4 + 2 * i)) & 3] for i in range((code >> 12) & 7)]
