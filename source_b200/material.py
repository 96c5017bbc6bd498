"""Materials of the hot path, with Raysect's constructor signatures.

Only the description lives here (spectral functions, flags, importance); evaluation happens on the
device (``csrc/rsb_path.h``).  Reference: raysect/optical/material/{material,lambert,absorber,
dielectric}.pyx and emitter/uniform.pyx.
"""
from .spectral import ConstantSF, InterpolatedSF, Sellmeier, SpectralFunction


class Material:
    """raysect/core/material.pyx:32-50 + optical/material/material.pyx: base (importance 0)"""

    def __init__(self):
        self._importance = 0.0

    @property
    def importance(self):
        return self._importance

    @importance.setter
    def importance(self, value):
        if value < 0:
            raise ValueError("Material sampling importance cannot be less than zero.")
        self._importance = float(value)


class AbsorbingSurface(Material):
    """absorber.pyx:50-55: returns an empty spectrum, terminates the path"""


class Lambert(Material):
    """lambert.pyx:43-112: ideal diffuse reflector"""

    def __init__(self, reflectivity=None):
        super().__init__()
        if reflectivity is None:
            reflectivity = ConstantSF(0.5)
        if not isinstance(reflectivity, SpectralFunction):
            raise TypeError("reflectivity must be a SpectralFunction")
        self.reflectivity = reflectivity


class UniformSurfaceEmitter(Material):
    """emitter/uniform.pyx:37-88: emission_spectrum * scale, importance 1"""

    def __init__(self, emission_spectrum, scale=1.0):
        super().__init__()
        self.emission_spectrum = emission_spectrum
        self.scale = float(scale)
        self.importance = 1.0


class Checkerboard(Material):
    """emitter/checkerboard.pyx:38-146: alternating squares of two emission spectra, picked by the local hit point"""

    def __init__(self, width=1.0, emission_spectrum1=None, emission_spectrum2=None, scale1=0.25, scale2=0.5):
        super().__init__()
        if emission_spectrum1 is None or emission_spectrum2 is None:
            raise TypeError("the stand-alone mirror has no d65_white: pass both emission spectra")
        if width == 0:
            raise ZeroDivisionError("float division")
        self.width = float(width)
        self.emission_spectrum1 = emission_spectrum1
        self.emission_spectrum2 = emission_spectrum2
        self.scale1 = float(scale1)
        self.scale2 = float(scale2)
        self.importance = 1.0


class Conductor(Material):
    """conductor.pyx:39-147: specular reflection off a metal with complex refractive index n + ik (Fresnel)"""

    def __init__(self, index, extinction):
        super().__init__()
        if not isinstance(index, SpectralFunction) or not isinstance(extinction, SpectralFunction):
            raise TypeError("index and extinction must be SpectralFunction objects")
        self.index = index
        self.extinction = extinction


class RoughConductor(Material):
    """conductor.pyx:157-344: Cook-Torrance metal -- GGX facet distribution, Smith shadowing, conductor Fresnel;
    roughness in (0, 1]"""

    def __init__(self, index, extinction, roughness):
        super().__init__()
        if not isinstance(index, SpectralFunction) or not isinstance(extinction, SpectralFunction):
            raise TypeError("index and extinction must be SpectralFunction objects")
        self.index = index
        self.extinction = extinction
        self.roughness = roughness

    @property
    def roughness(self):
        return self._roughness

    @roughness.setter
    def roughness(self, value):
        if value <= 0 or value > 1:
            raise ValueError("Surface roughness must lie in the range (0, 1].")
        self._roughness = float(value)


class UniformVolumeEmitter(Material):
    """emitter/uniform.pyx:91-133 on emitter/homogeneous.pyx:40-93: transparent surface, emission_spectrum * scale
    (W/m^3/str/nm) integrated along the path inside the primitive; importance 1 (homogeneous.pyx:48)"""

    def __init__(self, emission_spectrum, scale=1.0):
        super().__init__()
        self.emission_spectrum = emission_spectrum
        self.scale = float(scale)
        self.importance = 1.0


class UnityVolumeEmitter(Material):
    """emitter/unity.pyx:79-99: 1 W/m^3/str/nm in every bin"""

    def __init__(self):
        super().__init__()
        self.importance = 1.0


class UnitySurfaceEmitter(Material):
    """emitter/unity.pyx:37-76: 1 W/m^2/str/nm in every bin, whatever the direction (importance stays 0)"""


class Dielectric(Material):
    """dielectric.pyx:120-330: Fresnel reflect/transmit with Beer-Lambert volume attenuation, importance 1"""

    def __init__(self, index, transmission, external_index=None, transmission_only=False):
        super().__init__()
        self.index = index
        self.transmission = transmission
        self.transmission_only = bool(transmission_only)
        self.external_index = external_index if external_index is not None else ConstantSF(1.0)
        self.importance = 1.0


# ---- Schott catalogue entries used by the named configurations -------------------------------------
# Published catalogue data (Sellmeier B1..C3 and internal transmittance at 25 mm), the same rows the
# reference reads from raysect/optical/library/glass/data/schott_catalog_2000.csv; processing follows
# raysect/optical/library/glass/schott.py:44-94 (transmission per metre = tau25 ** 40, zero entries dropped).
_TAUI25_WAVELENGTHS_UM = [2.500, 2.325, 1.970, 1.530, 1.060, 0.700, 0.660, 0.620, 0.580, 0.546, 0.500, 0.460, 0.436,
                          0.420, 0.405, 0.400, 0.390, 0.380, 0.370, 0.365, 0.350, 0.334, 0.320, 0.310, 0.300, 0.290,
                          0.280, 0.270, 0.260, 0.250]
_SCHOTT = {
    "N-BK7": ((1.03961212, 0.231792344, 1.01046945, 0.0060006987, 0.0200179144, 103.560653),
              [0.36, 0.56, 0.84, 0.98, 0.997, 0.996, 0.994, 0.994, 0.995, 0.996, 0.994, 0.993, 0.992, 0.993, 0.993,
               0.992, 0.989, 0.983, 0.977, 0.971, 0.92, 0.78, 0.52, 0.25, 0.05]),
    "SF11": ((1.73848403, 0.311168974, 1.17490871, 0.0136068604, 0.0615960463, 121.922711),
             [0.61, 0.7, 0.93, 0.982, 0.997, 0.993, 0.991, 0.991, 0.991, 0.989, 0.976, 0.94, 0.86, 0.7, 0.34, 0.2,
              0.01]),
}


def schott(glass_name):
    """``raysect.optical.library.schott(name)`` for the glasses the named configurations use."""
    import numpy as np
    try:
        sellmeier, taui25 = _SCHOTT[glass_name]
    except KeyError:
        raise ValueError("This glass could not be found in the available Schott catalog.")
    wavelengths = np.array(_TAUI25_WAVELENGTHS_UM) * 1000
    pairs = [(w, t ** 40) for w, t in zip(wavelengths, taui25) if t]
    w = np.array([p[0] for p in pairs])
    t = np.array([p[1] for p in pairs])
    return Dielectric(index=Sellmeier(*sellmeier), transmission=InterpolatedSF(w, t))
