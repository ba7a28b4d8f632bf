"""Unit-transparent quantity shim — TEST INFRASTRUCTURE ONLY (oracle/).

Lets the UNMODIFIED reference modules splib/spcpl.py and splib/sputils.py import
and run without omuse/amuse.  Valid because every hot-path unit in the reference
is SI-coherent (Pa, K, m, s, kg/kg, m^2/s^2), so stripping units changes no number.
"""
import numpy as np


class Unit(object):
    __array_ufunc__ = None  # `ndarray | unit` must defer to Unit.__ror__

    def __init__(self, name="1"):
        self.name = name

    def __mul__(self, other):
        return Unit()

    __rmul__ = __truediv__ = __rtruediv__ = __pow__ = __mul__

    def __ror__(self, x):  # value | unit
        return Q(x, self)


class Q(np.ndarray):
    __array_priority__ = 1000

    def __new__(cls, x, unit=None):
        o = np.asarray(x, dtype=float).view(cls)
        o.unit = unit or Unit()
        return o

    def __array_finalize__(self, obj):
        self.unit = getattr(obj, "unit", Unit())

    def __getitem__(self, i):  # scalars stay Q so `Ph[-1].value_in(...)` works
        r = np.ndarray.__getitem__(self, i)
        return r if isinstance(r, Q) else Q(r, self.unit)

    def __bool__(self):
        return bool(np.asarray(self).any())

    @property
    def number(self):
        return np.asarray(self)

    def value_in(self, u):
        return np.asarray(self)


class _Units(object):
    def __getattr__(self, name):
        return Unit(name)


units = _Units()


def to_quantity(x):
    return x if isinstance(x, Q) else Q(x)
