"""cokrig_b200: ctypes binding + host orchestration of the B200-native cokriging hot path."""
from . import _lib  # noqa: F401  (raises ImportError if libcokrig_b200.so is missing)
from ._lib import METRIC_EUCLID, METRIC_HAVERSINE, CokrigError  # noqa: F401
from . import ops  # noqa: F401
