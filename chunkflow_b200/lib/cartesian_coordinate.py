"""zyx coordinate triple used in the ``Inferencer`` constructor signature.

Boundary type only: mirrors the arithmetic and the
all-axes ("partial order") comparisons of the reference's ``Cartesian``
(chunkflow/lib/cartesian_coordinate.py:34-188).  Bounding boxes and task grids are
out of scope for the inference hot path.
"""
from __future__ import annotations

import math
import operator
from numbers import Number
from typing import NamedTuple, Optional, Sequence, Union


class _ZYX(NamedTuple):
    z: Union[int, float]
    y: Union[int, float]
    x: Union[int, float]


def _elementwise(op):
    def method(self, other):
        if isinstance(other, Number):
            other = (other, other, other)
        if len(other) != 3:
            raise TypeError("expected a number or a zyx triple")
        return Cartesian(*(op(a, b) for a, b in zip(self, other)))
    return method


def _all_axes(op):
    def method(self, other) -> bool:
        if isinstance(other, Number):
            other = (other, other, other)
        return all(op(a, b) for a, b in zip(self, other))
    return method


class Cartesian(_ZYX):
    """Immutable (z, y, x) with element-wise arithmetic.

    Comparisons hold only if they hold on EVERY axis, like the reference
    (cartesian_coordinate.py:131-165) -- so ``a >= b`` being False does not imply
    ``a < b``.
    """
    __slots__ = ()

    @classmethod
    def from_collection(cls, col: Sequence) -> "Cartesian":
        if len(col) != 3:
            raise ValueError("a Cartesian needs exactly 3 components (z, y, x)")
        return cls(*col)

    __add__ = _elementwise(operator.add)
    __radd__ = __add__
    __sub__ = _elementwise(operator.sub)
    __mul__ = _elementwise(operator.mul)
    __rmul__ = __mul__
    __floordiv__ = _elementwise(operator.floordiv)
    __truediv__ = _elementwise(operator.truediv)
    __mod__ = _elementwise(operator.mod)

    __lt__ = _all_axes(operator.lt)
    __le__ = _all_axes(operator.le)
    __gt__ = _all_axes(operator.gt)
    __ge__ = _all_axes(operator.ge)
    __eq__ = _all_axes(operator.eq)
    __ne__ = _all_axes(operator.ne)

    def __hash__(self):
        return hash((self.z, self.y, self.x))

    def __neg__(self):
        return Cartesian(-self.z, -self.y, -self.x)

    @property
    def ceil(self):
        return Cartesian(*(math.ceil(v) for v in self))

    @property
    def floor(self):
        return Cartesian(*(math.floor(v) for v in self))

    @property
    def tuple(self):
        return (self.z, self.y, self.x)

    @property
    def list(self):
        return [self.z, self.y, self.x]

    @property
    def inverse(self):
        return Cartesian(self.x, self.y, self.z)


def to_cartesian(x: Optional[Sequence]) -> Optional[Cartesian]:
    """None passes through; anything else must be a 3-sequence (reference :26-31)."""
    if x is None:
        return None
    return Cartesian.from_collection(tuple(x))
