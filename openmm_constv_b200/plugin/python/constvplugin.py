"""`constvplugin` for hosts without SWIG/OpenMM: the reference's Python surface
(python/constvplugin.i:65-80 -- class ConstVLangevinIntegrator, constructor
``(temperature, frictionCoeff, stepSize, numCells=2, zmax=-1)``, the ten accessors and ``step``; :82-100 --
class ConstVDrudeLangevinIntegrator)
served by the ctypes host mirror in ``openmm_constv_b200.integrator``.

The reference decorates two getters with units (constvplugin.i:35-41): getTemperature -> kelvin,
getFriction -> 1/picosecond.  When an OpenMM unit package is importable the same Quantity objects
are returned; otherwise the bare numbers in those units are.
"""
import os
import sys

_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from openmm_constv_b200.integrator import (  # noqa: E402
    ConstVDrudeLangevinIntegrator as _HostDrudeIntegrator, ConstVLangevinIntegrator as _HostIntegrator, DrudeForce,
    OpenMMException, SlabContext, SlabSystem)

try:  # OpenMM >= 7.6, then the simtk namespace the reference imports
    from openmm import unit  # type: ignore
except ImportError:  # pragma: no cover - neither exists in the build container
    try:
        from simtk import unit  # type: ignore
    except ImportError:
        unit = None


def _strip(value, in_unit):
    if unit is not None and isinstance(value, unit.Quantity):
        return value.value_in_unit(in_unit)
    return value


class ConstVLangevinIntegrator(_HostIntegrator):
    def __init__(self, temperature, frictionCoeff, stepSize, numCells=2, zmax=-1):
        if unit is not None:
            temperature = _strip(temperature, unit.kelvin)
            frictionCoeff = _strip(frictionCoeff, 1 / unit.picosecond)
            stepSize = _strip(stepSize, unit.picosecond)
            zmax = _strip(zmax, unit.nanometer)
        super().__init__(temperature, frictionCoeff, stepSize, numCells, zmax)

    def getTemperature(self):
        val = super().getTemperature()
        return unit.Quantity(val, unit.kelvin) if unit is not None else val

    def getFriction(self):
        val = super().getFriction()
        return unit.Quantity(val, 1 / unit.picosecond) if unit is not None else val


class ConstVDrudeLangevinIntegrator(_HostDrudeIntegrator):
    """python/constvplugin.i:82-100; unit decoration of the getters as :43-61."""

    def __init__(self, temperature, friction, drudeTemperature, drudeFriction, stepSize, numCells=2, zmax=-1):
        if unit is not None:
            temperature = _strip(temperature, unit.kelvin)
            friction = _strip(friction, 1 / unit.picosecond)
            drudeTemperature = _strip(drudeTemperature, unit.kelvin)
            drudeFriction = _strip(drudeFriction, 1 / unit.picosecond)
            stepSize = _strip(stepSize, unit.picosecond)
            zmax = _strip(zmax, unit.nanometer)
        super().__init__(temperature, friction, drudeTemperature, drudeFriction, stepSize, numCells, zmax)

    def getTemperature(self):
        val = super().getTemperature()
        return unit.Quantity(val, unit.kelvin) if unit is not None else val

    def getFriction(self):
        val = super().getFriction()
        return unit.Quantity(val, 1 / unit.picosecond) if unit is not None else val

    def getDrudeTemperature(self):
        val = super().getDrudeTemperature()
        return unit.Quantity(val, unit.kelvin) if unit is not None else val

    def getDrudeFriction(self):
        val = super().getDrudeFriction()
        return unit.Quantity(val, 1 / unit.picosecond) if unit is not None else val

    def getMaxDrudeDistance(self):
        val = super().getMaxDrudeDistance()
        return unit.Quantity(val, unit.nanometer) if unit is not None else val


__all__ = ["ConstVLangevinIntegrator", "ConstVDrudeLangevinIntegrator", "DrudeForce", "OpenMMException", "SlabContext",
           "SlabSystem"]
