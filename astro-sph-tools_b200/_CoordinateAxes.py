"""Axis selector of the projection API (mirrors the reference type, _CoordinateAxes.py:3-32: same
member names, values, string forms and ValueError behaviour, so user code that holds the reference's
enum semantics keeps working)."""
from enum import Enum

_NAMES = ("x", "y", "z")


class CoordinateAxes(Enum):
    """The three axes of a 3-D cartesian grid; the projection axis is the one integrated out."""
    X = 0
    Y = 1
    Z = 2

    def __str__(self) -> str:
        return _NAMES[self.value]

    @staticmethod
    def from_string(value: str) -> "CoordinateAxes":
        key = value.strip().lower()
        if key not in _NAMES:
            raise ValueError()
        return CoordinateAxes(_NAMES.index(key))

    @property
    def plane_columns(self):
        """Position columns spanning the image plane: X->(1,2), Y->(0,2), Z->(0,1)
        (reference: _pixel_calculations.pyx:20-28, _projector.py:38-46)."""
        return ((1, 2), (0, 2), (0, 1))[self.value]
