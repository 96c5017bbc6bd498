"""Spectral functions: the host-side resampling that feeds the device's per-slice material tables.

Mirrors raysect/optical/spectralfunction.pyx (SpectralFunction.sample/average, ConstantSF,
InterpolatedSF, NumericallyIntegratedSF) and the Sellmeier dispersion of
raysect/optical/material/dielectric.pyx:40-117.  In the reference these are evaluated once per
(material, spectral slice) and cached (spectralfunction.pyx:196-215); the trace loop only ever reads
the cached table, which is exactly what is uploaded to the GPU.
"""
import math

import numpy as np


class SpectralFunction:
    """spectralfunction.pyx:45-327: bin-averaging resampler over an integrate() primitive."""

    def __init__(self):
        self._average_cache = (None, None, 0.0)
        self._sample_cache = (None, None, None, None)

    def integrate(self, min_wavelength, max_wavelength):
        raise NotImplementedError("Virtual method integrate() not implemented.")

    def average(self, min_wavelength, max_wavelength):
        """spectralfunction.pyx:140-168"""
        if self._average_cache[0] == min_wavelength and self._average_cache[1] == max_wavelength:
            return self._average_cache[2]
        average = self.integrate(min_wavelength, max_wavelength) / (max_wavelength - min_wavelength)
        self._average_cache = (min_wavelength, max_wavelength, average)
        return average

    def sample(self, min_wavelength, max_wavelength, bins):
        """spectralfunction.pyx:171-216: sample[i] = (1/delta) * integrate(lower_i, upper_i)"""
        key = (min_wavelength, max_wavelength, bins)
        if self._sample_cache[:3] == key:
            return self._sample_cache[3]
        samples = np.zeros(bins, dtype=np.float64)
        delta = (max_wavelength - min_wavelength) / bins
        lower = min_wavelength
        reciprocal = 1.0 / delta
        for index in range(bins):
            upper = min_wavelength + (index + 1) * delta
            samples[index] = reciprocal * self.integrate(lower, upper)
            lower = upper
        self._sample_cache = key + (samples,)
        return samples


class ConstantSF(SpectralFunction):
    """spectralfunction.pyx:509-595"""

    def __init__(self, value):
        super().__init__()
        self.value = float(value)

    def evaluate(self, wavelength):
        return self.value

    def integrate(self, min_wavelength, max_wavelength):
        return self.value * (max_wavelength - min_wavelength)

    def average(self, min_wavelength, max_wavelength):
        return self.value

    def sample(self, min_wavelength, max_wavelength, bins):
        return np.full(bins, self.value, dtype=np.float64)


def _find_index(x, v):
    """raysect/core/math/cython/utility.pyx:40-94"""
    if v < x[0]:
        return -1
    top = len(x) - 1
    if v >= x[top]:
        return top
    bottom = 0
    bis = top // 2
    while (top - bottom) != 1:
        if v >= x[bis]:
            bottom = bis
        else:
            top = bis
        bis = (top + bottom) // 2
    return bottom


def _lerp(x0, x1, y0, y1, x):
    """utility.pxd:95-96"""
    return ((y1 - y0) / (x1 - x0)) * (x - x0) + y0


def _integrate(x, y, x0, x1):
    """raysect/core/math/cython/utility.pyx:137-240: trapezium integral of the piecewise-linear
    function with nearest-neighbour extrapolation."""
    if x1 <= x0:
        return 0.0
    lower_index = _find_index(x, x0) + 1
    upper_index = _find_index(x, x1)
    if upper_index == -1:
        return y[0] * (x1 - x0)
    top_index = len(x) - 1
    if lower_index > top_index:
        return y[top_index] * (x1 - x0)
    if lower_index > upper_index:
        m = (y[lower_index] - y[upper_index]) / (x[lower_index] - x[upper_index])
        y0 = m * (x0 - x[upper_index]) + y[upper_index]
        y1 = m * (x1 - x[upper_index]) + y[upper_index]
        return 0.5 * (y0 + y1) * (x1 - x0)
    integral_sum = 0.0
    if lower_index == 0:
        integral_sum += y[0] * (x[0] - x0)
    else:
        y0 = _lerp(x[lower_index - 1], x[lower_index], y[lower_index - 1], y[lower_index], x0)
        integral_sum += 0.5 * (y0 + y[lower_index]) * (x[lower_index] - x0)
    for index in range(lower_index, upper_index):
        integral_sum += 0.5 * (y[index] + y[index + 1]) * (x[index + 1] - x[index])
    if upper_index == top_index:
        integral_sum += y[top_index] * (x1 - x[top_index])
    else:
        y1 = _lerp(x[upper_index], x[upper_index + 1], y[upper_index], y[upper_index + 1], x1)
        integral_sum += 0.5 * (y[upper_index] + y1) * (x1 - x[upper_index])
    return integral_sum


class InterpolatedSF(SpectralFunction):
    """spectralfunction.pyx:403-506: linearly interpolated samples, ends extrapolated flat."""

    def __init__(self, wavelengths, samples, normalise=False):
        super().__init__()
        w = np.array(wavelengths, dtype=np.float64)
        s = np.array(samples, dtype=np.float64)
        if w.ndim != 1:
            raise ValueError("Wavelength array must be 1D.")
        if s.shape[0] != w.shape[0]:
            raise ValueError("Wavelength and sample arrays must be the same length.")
        indices = np.argsort(w)
        self.wavelengths = w[indices]
        self.samples = s[indices]
        self._w = [float(v) for v in self.wavelengths]
        self._s = [float(v) for v in self.samples]
        if normalise:
            self.samples /= self.integrate(self.wavelengths.min(), self.wavelengths.max())
            self._s = [float(v) for v in self.samples]

    def integrate(self, min_wavelength, max_wavelength):
        return _integrate(self._w, self._s, float(min_wavelength), float(max_wavelength))


class NumericallyIntegratedSF(SpectralFunction):
    """spectralfunction.pyx:330-400: midpoint-rule integration of function()."""

    def __init__(self, sample_resolution=1.0):
        super().__init__()
        if sample_resolution <= 0:
            raise ValueError("Sampling resolution must be greater than zero.")
        self.sample_resolution = float(sample_resolution)

    def function(self, wavelength):
        raise NotImplementedError("Virtual method function() not implemented.")

    def evaluate(self, wavelength):
        return self.function(wavelength)

    def integrate(self, min_wavelength, max_wavelength):
        samples = int(math.ceil((max_wavelength - min_wavelength) / self.sample_resolution))
        samples = max(samples, 1)
        total = 0.0
        delta = (max_wavelength - min_wavelength) / samples
        for i in range(samples):
            centre = min_wavelength + (0.5 + i) * delta
            total += self.function(centre) * delta
        return total


class Sellmeier(NumericallyIntegratedSF):
    """raysect/optical/material/dielectric.pyx:40-117: three-term Sellmeier refractive index."""

    def __init__(self, b1, b2, b3, c1, c2, c3, sample_resolution=10):
        super().__init__(sample_resolution)
        self.b1, self.b2, self.b3 = float(b1), float(b2), float(b3)
        self.c1, self.c2, self.c3 = float(c1), float(c2), float(c3)

    def function(self, wavelength):
        w2 = wavelength * wavelength * 1e-6
        return math.sqrt(1 + (self.b1 * w2) / (w2 - self.c1)
                         + (self.b2 * w2) / (w2 - self.c2)
                         + (self.b3 * w2) / (w2 - self.c3))
