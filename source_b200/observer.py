"""Observer side of the hot path: PinholeCamera + FullFrameSampler2D + SpectralPowerPipeline2D.

Mirrors raysect/optical/observer/base/observer.pyx (observe, _slice_spectrum, _generate_templates),
base/slice.pyx, imaging/pinhole.pyx, sampler2d.pyx:60-102 and pipeline/spectral/power.pyx:335-486.
The per-pixel inner loop (_render_pixel: ray generation, Ray.trace, add_sample) runs on the device;
this module keeps the reference's orchestration: one render pass per spectral slice, results merged
into an accumulating StatsArray3D frame with the reference's combine rule.
"""
import math

import numpy as np

from . import _cabi as cabi
from .engine import camera_desc, ray_config
from .math3d import AffineMatrix3D
from .scenegraph import Node


class Observer(Node):
    """raysect/core/scenegraph/observer.pyx: marker base so World can track observers"""


class SpectralSlice:
    """raysect/optical/observer/base/slice.pyx:31-69"""

    def __init__(self, min_wavelength, max_wavelength, bins, slice_bins, slice_offset):
        if bins <= 0:
            raise ValueError("The bin count must be greater than 0.")
        if min_wavelength <= 0:
            raise ValueError("The minimum wavelength must be greater than 0.")
        if max_wavelength <= 0:
            raise ValueError("The maximum wavelength must be greater than 0.")
        if min_wavelength >= max_wavelength:
            raise ValueError("The minimum wavelength must be less than the maximum wavelength.")
        if slice_bins <= 0:
            raise ValueError("The slice bin count must be greater than 0.")
        if slice_offset < 0:
            raise ValueError("The slice offset cannot be less that 0.")
        if (slice_offset + slice_bins) > bins:
            raise ValueError("The slice offset plus the bin count extends beyond the full bin count.")
        delta_wavelength = (max_wavelength - min_wavelength) / bins
        self.min_wavelength = min_wavelength + delta_wavelength * slice_offset
        self.max_wavelength = min_wavelength + delta_wavelength * (slice_offset + slice_bins)
        self.offset = slice_offset
        self.bins = slice_bins
        self.total_bins = bins
        self.total_min_wavelength = min_wavelength
        self.total_max_wavelength = max_wavelength


def combine_samples(mx, vx, nx, my, vy, ny):
    """raysect/core/math/statsarray.pyx:780-857 (_combine_samples), vectorised over arrays.
    (mx, vx, nx) = stored frame, (my, vy, ny) = new results; returns (mt, vt, nt)."""
    mx, vx, my, vy = (np.asarray(a, dtype=np.float64) for a in (mx, vx, my, vy))
    nx = np.asarray(nx, dtype=np.int32)
    ny = np.broadcast_to(np.asarray(ny, dtype=np.int32), nx.shape)
    swap = nx < ny
    nx, ny = np.where(swap, ny, nx), np.where(swap, nx, ny)
    mx, my = np.where(swap, my, mx), np.where(swap, mx, my)
    vx, vy = np.where(swap, vy, vx), np.where(swap, vx, vy)
    mt = np.zeros_like(mx)
    vt = np.zeros_like(mx)
    nt = np.zeros_like(nx)
    with np.errstate(divide="ignore", invalid="ignore"):
        # common case
        c = (nx > 1) & (ny > 1)
        n = (nx + ny).astype(np.float64)
        m = (nx * mx + ny * my) / n
        bx = (nx - 1) * vx / nx.astype(np.float64)
        by = (ny - 1) * vy / ny.astype(np.float64)
        v = (nx * (mx * mx + bx) + ny * (my * my + by)) / n - m * m
        v = (nx + ny) * v / (n - 1)
        mt[c], vt[c], nt[c] = m[c], v[c], (nx + ny)[c]
        # nx == 1: (ny == 0) single sample, (ny == 1) two samples
        c = (nx == 1) & (ny == 0)
        mt[c], vt[c], nt[c] = mx[c], 0.0, 1
        c = (nx == 1) & (ny == 1)
        m2 = 0.5 * (mx + my)
        t = mx - m2
        mt[c], vt[c], nt[c] = m2[c], (2 * t * t)[c], 2
        # nx > 1 with ny in {0, 1}
        c = (nx > 1) & (ny == 0)
        mt[c], vt[c], nt[c] = mx[c], vx[c], nx[c]
        c = (nx > 1) & (ny == 1)
        # _add_sample(my, mt, vt, nt), statsarray.pyx:743-777
        n1 = nx + 1
        ma = mx + (my - mx) / n1
        va = (vx * (nx - 1) + (my - mx) * (my - ma)) / (n1 - 1)
        mt[c], vt[c], nt[c] = ma[c], va[c], n1[c]
    return mt, vt, nt


class StatsArray3D:
    """raysect/core/math/statsarray.pyx:513-725: mean / variance / samples arrays of shape (nx, ny, nz)"""

    def __init__(self, nx, ny, nz):
        if nx < 1 or ny < 1 or nz < 1:
            raise ValueError("Number of elements must be >= 1.")
        self.nx, self.ny, self.nz = nx, ny, nz
        self.mean = np.zeros((nx, ny, nz), dtype=np.float64)
        self.variance = np.zeros((nx, ny, nz), dtype=np.float64)
        self.samples = np.zeros((nx, ny, nz), dtype=np.int32)

    @property
    def shape(self):
        return (self.nx, self.ny, self.nz)

    def clear(self):
        self.mean[...] = 0
        self.variance[...] = 0
        self.samples[...] = 0

    def errors(self):
        """statsarray.pyx:725-739 (_std_error): sqrt(v / n), 0 where n <= 0 or v <= 0"""
        out = np.zeros_like(self.mean)
        ok = (self.samples > 0) & (self.variance > 0)
        out[ok] = np.sqrt(self.variance[ok] / self.samples[ok])
        return out

    def combine_slice(self, pixels, offset, mean, variance, sample_count):
        """StatsArray3D.combine_samples (statsarray.pyx:623-667) for bins [offset, offset+mean.shape[-1])
        of the listed pixels (all pixels when ``pixels`` is None)."""
        if sample_count < 1:
            raise ValueError('Number of samples must not be less than 1.')
        nb = mean.shape[-1]
        sl = slice(offset, offset + nb)
        if pixels is None:
            idx = (slice(None), slice(None))
        else:
            pixels = np.asarray(pixels).reshape(-1, 2)
            idx = (pixels[:, 0], pixels[:, 1])
        new_m = mean[idx]
        new_v = np.maximum(variance[idx], 0.0)   # clamp, statsarray.pyx:647-650
        mt, vt, nt = combine_samples(self.mean[idx][..., sl], self.variance[idx][..., sl], self.samples[idx][..., sl],
                                     new_m, new_v, sample_count)
        if pixels is None:
            self.mean[:, :, sl], self.variance[:, :, sl], self.samples[:, :, sl] = mt, vt, nt
        else:
            self.mean[pixels[:, 0], pixels[:, 1], sl] = mt
            self.variance[pixels[:, 0], pixels[:, 1], sl] = vt
            self.samples[pixels[:, 0], pixels[:, 1], sl] = nt


class SpectralPowerPipeline2D:
    """raysect/optical/observer/pipeline/spectral/power.pyx:335-443"""

    def __init__(self, accumulate=True, name=None):
        self.name = name or "SpectralPowerPipeline2D"
        self.accumulate = accumulate
        self.frame = None
        self.min_wavelength = self.max_wavelength = self.delta_wavelength = 0
        self.bins = 0
        self.wavelengths = None
        self._samples = 0
        self._spectral_slices = None

    def initialise(self, pixels, pixel_samples, min_wavelength, max_wavelength, spectral_bins, spectral_slices, quiet=True):
        nx, ny = pixels
        self._pixels = pixels
        self._samples = pixel_samples
        self._spectral_slices = spectral_slices
        self.min_wavelength = min_wavelength
        self.max_wavelength = max_wavelength
        self.delta_wavelength = (max_wavelength - min_wavelength) / spectral_bins
        self.bins = spectral_bins
        self.wavelengths = np.array([min_wavelength + (0.5 + i) * self.delta_wavelength for i in range(spectral_bins)])
        if not self.accumulate or self.frame is None or self.frame.shape != (nx, ny, spectral_bins):
            self.frame = StatsArray3D(nx, ny, spectral_bins)

    def update_slice(self, pixels, slice_id, mean, variance):
        """power.pyx:424-437 (update) for all listed pixels at once"""
        s = self._spectral_slices[slice_id]
        self.frame.combine_slice(pixels, s.offset, mean, variance, self._samples)

    def finalise(self):
        pass


class SpectralRadiancePipeline2D(SpectralPowerPipeline2D):
    """raysect/optical/observer/pipeline/spectral/radiance.pyx: the same frame statistics, of the spectral RADIANCE:
    its pixel processor does not apply the pixel sensitivity (radiance.pyx:256-260)"""
    radiance = True

    def __init__(self, accumulate=True, name=None):
        super().__init__(accumulate, name or "SpectralRadiancePipeline2D")


class FullFrameSampler2D:
    """raysect/optical/observer/sampler2d.pyx:38-102.  The reference shuffles the task list so the image
    assembles randomly on screen; pixel streams here are keyed on the pixel, so order does not matter and
    tasks are returned in scan order."""

    def __init__(self, mask=None):
        self.mask = None if mask is None else np.asarray(mask).astype(bool)
        if self.mask is not None and self.mask.ndim != 2:
            raise ValueError("Mask must be a 2D array.")

    def generate_tasks(self, pixels):
        nx, ny = pixels
        if self.mask is None:
            return None   # every pixel
        if self.mask.shape != (nx, ny):
            if np.all(self.mask):
                return None
            raise ValueError('The pixel geometry passed to the frame sampler is inconsistent with the mask shape.')
        ix, iy = np.nonzero(self.mask.T)[1], np.nonzero(self.mask.T)[0]
        return np.stack([ix, iy], axis=1).astype(np.int32)


class SpectralAdaptiveSampler2D:
    """raysect/optical/observer/sampler2d.pyx:325-700: frame sampler that re-samples the noisiest fraction of the
    frame.  Reads the pipeline's frame statistics (mean, variance, samples -- the arrays the device render fills) and
    returns the pixels whose normalised standard error exceeds the (1 - fraction) percentile (or ``cutoff``), plus
    every pixel that has fewer than max(min_samples, max_samples / ratio) samples.  Task order does not matter here
    (pixel streams are keyed on the pixel), so tasks come back in scan order instead of shuffled."""

    def __init__(self, pipeline, fraction=0.2, ratio=10.0, min_samples=1000, cutoff=0.0, reduction_method='percentile',
                 percentile=100., mask=None):
        if not isinstance(pipeline, SpectralPowerPipeline2D):
            raise TypeError('Sampler only compatible with SpectralPowerPipeline2D or SpectralRadiancePipeline2D pipelines.')
        if fraction <= 0 or fraction > 1.:
            raise ValueError("Attribute 'fraction' must be in the range (0, 1].")
        if ratio < 1.:
            raise ValueError("Attribute 'ratio' must be >= 1.")
        if min_samples < 1:
            raise ValueError("Attribute 'min_samples' must be >= 1.")
        if cutoff < 0 or cutoff > 1.:
            raise ValueError("Attribute 'cutoff' must be in the range [0, 1].")
        if reduction_method not in {'weighted', 'mean', 'percentile', 'power_percentile'}:
            raise ValueError("Attribute 'reduction_method' must be 'weighted', 'mean', 'percentile' or 'power_percentile'.")
        if percentile < 0 or percentile > 100.:
            raise ValueError("Percentiles must be in the range [0, 100].")
        self.pipeline, self.fraction, self.ratio, self.min_samples = pipeline, fraction, ratio, int(min_samples)
        self.cutoff, self.reduction_method, self.percentile = cutoff, reduction_method, percentile
        self.mask = None if mask is None else np.asarray(mask).astype(bool)
        if self.mask is not None and self.mask.ndim != 2:
            raise ValueError("Mask must be a 2D array.")

    def _normalised(self, frame, mask):
        """_reduce_weighted / _reduce_mean / _reduce_percentile / _reduce_power_percentile (sampler2d.pyx:545-668);
        per-pixel sums run over the bins in order, as the reference's loops do"""
        mean = frame.mean
        samples = frame.samples
        # StatsArray3D.errors -> _std_error (statsarray.pyx:728-739)
        ok = (samples > 0) & (frame.variance > 0)
        error = np.zeros_like(mean)
        error[ok] = np.sqrt(frame.variance[ok] / samples[ok])
        normalised = np.zeros((frame.nx, frame.ny))
        for x in range(frame.nx):
            for y in range(frame.ny):
                if not mask[x, y]:
                    continue
                pos = mean[x, y] > 0
                if not pos.any():
                    continue
                e, m = error[x, y][pos], mean[x, y][pos]
                if self.reduction_method == 'weighted':
                    acc, power = 0.0, 0.0
                    for ei, mi in zip(e, m):
                        acc += ei
                        power += mi
                    normalised[x, y] = acc / power if power else acc
                elif self.reduction_method == 'mean':
                    acc = 0.0
                    for ei, mi in zip(e, m):
                        acc += ei / mi
                    normalised[x, y] = acc / len(e)
                elif self.reduction_method == 'percentile':
                    normalised[x, y] = np.percentile(e / m, self.percentile)
                else:
                    threshold = np.percentile(m, 100. - self.percentile)
                    sel = mean[x, y] >= threshold
                    normalised[x, y] = max(0.0, np.max(error[x, y][sel] / mean[x, y][sel]))
        return normalised

    def generate_tasks(self, pixels):
        nx, ny = pixels
        if self.mask is None or (self.mask.shape != (nx, ny) and np.all(self.mask)):
            self.mask = np.ones((nx, ny), dtype=bool)
        if self.mask.shape != (nx, ny):
            raise ValueError('The pixel geometry passed to the frame sampler is inconsistent with the mask shape.')
        frame = self.pipeline.frame
        full = np.argwhere(self.mask).astype(np.int32)
        if frame is None:
            return full
        if (nx, ny) != (frame.nx, frame.ny):
            raise ValueError('The pixel geometry passed to the frame sampler is inconsistent with the pipeline frame size.')
        min_samples = max(self.min_samples, int(frame.samples.max(2)[self.mask].max() / self.ratio))
        normalised = self._normalised(frame, self.mask)
        percentile_error = np.percentile(normalised[self.mask], (1 - self.fraction) * 100)
        cutoff = max(self.cutoff, percentile_error)
        frame_min_samples = frame.samples.min(2)
        todo = self.mask & ((frame_min_samples < min_samples) | (normalised > cutoff))
        return np.argwhere(todo).astype(np.int32)


class PinholeCamera(Observer):
    """raysect/optical/observer/imaging/pinhole.pyx + base/observer.pyx (Observer2D, _ObserverBase).

    Defaults follow observer.pyx:114-122, 916: 15 bins, 1 spectral ray, 375-740 nm, extinction 0.01,
    min depth 3, max depth 500, importance sampling on with path weight 0.2, 100 samples per pixel.
    ``rng_mode``/``seed`` select the per-pixel random streams (see include/raysect_b200.h).
    """

    def __init__(self, pixels, fov=None, sensitivity=None, frame_sampler=None, pipelines=None, parent=None,
                 transform=None, name=None):
        self._pixels = tuple(pixels)
        if len(self._pixels) != 2:
            raise ValueError("Pixels must be a 2 element tuple defining the x and y resolution.")
        if self._pixels[0] <= 0:
            raise ValueError("Number of x pixels must be greater than 0.")
        if self._pixels[1] <= 0:
            raise ValueError("Number of y pixels must be greater than 0.")
        self.fov = fov or 45
        self.sensitivity = sensitivity or 1.0
        self.frame_sampler = frame_sampler or FullFrameSampler2D()
        self.pipelines = pipelines or [SpectralPowerPipeline2D()]
        for p in self.pipelines:
            if not isinstance(p, SpectralPowerPipeline2D):
                raise NotImplementedError("only SpectralPowerPipeline2D runs on the B200 path; other pipelines are "
                                          "host-side post-processing of its spectral frame")
        self.pixel_samples = 100
        self.spectral_bins = 15
        self.spectral_rays = 1
        self.min_wavelength = 375.0
        self.max_wavelength = 740.0
        self.ray_extinction_prob = 0.01
        self.ray_extinction_min_depth = 3
        self.ray_max_depth = 500
        self.ray_importance_sampling = True
        self.ray_important_path_weight = 0.2
        self.quiet = True
        self.rng_mode = cabi.RNG_MT19937_64
        self.seed = 1
        self.render_complete = False
        self.ray_count = 0
        super().__init__(parent, transform, name)

    @property
    def pixels(self):
        return self._pixels

    @property
    def fov(self):
        return self._fov

    @fov.setter
    def fov(self, value):
        if value <= 0 or value >= 180:
            raise ValueError("The field-of-view angle must lie in the range (0, 180).")
        self._fov = value

    def _camera_desc(self, pixel_samples):
        """the RsbCamera of this observer for one render call (``pixel_samples`` samples per pixel)"""
        nx, ny = self._pixels
        return camera_desc(nx, ny, pixel_samples, self._fov, self.sensitivity, self.to_root())

    def _slice_spectrum(self):
        """observer.pyx:311-340"""
        if self.spectral_rays < 1 or self.spectral_rays > self.spectral_bins:
            raise ValueError("The number of spectral rays must be in the range [1, spectral_bins].")
        current = 0
        start = 0
        ranges = []
        while start < self.spectral_bins:
            current += self.spectral_bins / self.spectral_rays
            end = round(current)
            ranges.append((start, end))
            start = end
        return [SpectralSlice(self.min_wavelength, self.max_wavelength, self.spectral_bins, end - start, start)
                for start, end in ranges]

    def observe(self, passes=1):
        """observer.pyx:265-309.  ``passes`` > 1 is ``passes`` consecutive observe() calls into accumulating
        pipelines (the progressive-render loop of demos/cornell_box.py) rendered CONCURRENTLY: pass p draws from
        the pixel streams seeded ``seed + (p*n_slices + slice)*nx*ny + y*nx + x`` and the passes are merged in
        order with StatsArray3D.combine_samples, starting from an empty frame (include/raysect_b200.h,
        rsb_render_passes)."""
        from .scenegraph import World
        passes = int(passes)
        if passes < 1:
            raise ValueError("The number of passes must be at least 1.")
        self.render_complete = False
        world = self.root
        if not isinstance(world, World):
            raise TypeError("Observer is not connected to a scene graph containing a World object.")
        slices = self._slice_spectrum()
        for p in self.pipelines:
            if passes > 1 and p.accumulate and p.frame is not None and np.any(p.frame.samples):
                raise NotImplementedError("concurrent passes merge into an empty frame: clear the pipeline's frame "
                                          "or call observe() once per pass")
            p.initialise(self._pixels, self.pixel_samples * passes, self.min_wavelength, self.max_wavelength,
                         self.spectral_bins, slices, self.quiet)
        tasks = self.frame_sampler.generate_tasks(self._pixels)
        if tasks is not None and len(tasks) == 0:
            self.render_complete = True
            return
        accel = world.build_accelerator()
        nx, ny = self._pixels
        # power pipelines see samples scaled by the pixel sensitivity, radiance pipelines unscaled ones (= sensitivity
        # exactly 1.0): one render per distinct sensitivity, over the very same pixel streams
        cams = {}
        for p in self.pipelines:
            sens = 1.0 if getattr(p, "radiance", False) else float(self.sensitivity)
            if sens not in cams:
                cam = self._camera_desc(self.pixel_samples)
                cam.sensitivity = sens
                cams[sens] = cam
        self.ray_count = 0

        def sens_of(p):
            return 1.0 if getattr(p, "radiance", False) else float(self.sensitivity)

        def config_of(s):
            return ray_config(s.bins, s.min_wavelength, s.max_wavelength, self.ray_extinction_prob, self.ray_extinction_min_depth,
                              self.ray_max_depth, self.ray_importance_sampling, self.ray_important_path_weight)
        if len(slices) > 1 and len({s.bins for s in slices}) == 1 and hasattr(accel, "render_slices"):
            # Every spectral slice is an independent set of pixel streams (the reference renders them one after the other,
            # observer.pyx:299-305): all of them in ONE device render (rsb_render_slices), slice k of pass p drawing from
            # the streams seeded seed + (p*n_slices + k)*nx*ny + y*nx + x -- the seeds the per-slice loop below uses.
            cfg = config_of(slices[0])
            spectrals = [accel.flat.spectral(s.min_wavelength, s.max_wavelength, s.bins) for s in slices]
            for sens, cam in cams.items():
                rays = accel.render_slices(cam, cfg, spectrals, self.rng_mode, self.seed, tasks, passes=passes)
                for p in self.pipelines:
                    if sens_of(p) == sens:
                        f = p.frame
                        p._samples = self.pixel_samples * passes
                        accel.update_frame(f.mean, f.variance, f.samples, 0, frame_is_empty=not f.samples.any())
            self.ray_count += rays
            slices_done = True
        else:
            slices_done = False
        for slice_id, s in enumerate(slices if not slices_done else []):
            cfg = config_of(s)
            spectral = accel.flat.spectral(s.min_wavelength, s.max_wavelength, s.bins)
            # each slice is an independent pass with its own streams (the reference's single global stream
            # simply keeps running): offset the seed by the slice so passes are not correlated
            frames, rays = {}, 0
            for sens, cam in cams.items():
                mean, variance, rays = accel.render(cam, cfg, spectral, self.rng_mode,
                                                    self.seed + slice_id * nx * ny, tasks, passes=passes,
                                                    seed_stride=len(slices) * nx * ny)
                frames[sens] = (mean, variance)
            self.ray_count += rays
            for p in self.pipelines:
                mean, variance = frames[sens_of(p)]
                p.update_slice(tasks, slice_id, mean, variance)
        for p in self.pipelines:
            p.finalise()
        # the next observe() draws from fresh streams (the reference's single global stream simply keeps running):
        # a progressive loop -- observe() again and again into accumulating pipelines, demos/cornell_box.py:160-174 --
        # must not redraw the samples it already has.  Call c of a loop started at seed s uses the streams
        # s + ((c*passes + p)*n_slices + slice)*nx*ny + y*nx + x; assign ``seed`` to restart a sequence.
        self.seed += passes * len(slices) * nx * ny
        self.render_complete = True


class OrthographicCamera(PinholeCamera):
    """raysect/optical/observer/imaging/orthographic.pyx:37-170: parallel rays along the camera's +z axis from a
    ``width`` metres wide image plane (height follows the pixel aspect ratio); samples radiance directly
    (projection weight 1).  Everything else -- spectral slicing, pipelines, ray settings -- is Observer2D's."""

    def __init__(self, pixels, width, sensitivity=None, frame_sampler=None, pipelines=None, parent=None, transform=None,
                 name=None):
        super().__init__(pixels, None, sensitivity, frame_sampler, pipelines, parent, transform, name)
        self.width = width

    @property
    def width(self):
        return self._width

    @width.setter
    def width(self, width):
        if width <= 0:
            raise ValueError("width can not be less than or equal to 0 meters.")
        self._width = float(width)

    def _camera_desc(self, pixel_samples):
        nx, ny = self._pixels
        return camera_desc(nx, ny, pixel_samples, None, self.sensitivity, self.to_root(), width=self._width)


class CCDArray(OrthographicCamera):
    """raysect/optical/observer/imaging/ccd.pyx:40-151: a bare ``width`` metres wide sensor of square pixels; every sample
    leaves a uniformly drawn point of its pixel in a cosine-weighted direction over the hemisphere in front of the sensor
    (projection weight 0.5); a pixel's sensitivity is its area times 2 pi (ccd.pyx:150-151)."""

    def __init__(self, pixels=(720, 480), width=0.035, frame_sampler=None, pipelines=None, parent=None, transform=None, name=None):
        super().__init__(pixels, width, None, frame_sampler, pipelines, parent, transform, name)

    @property
    def sensitivity(self):
        nx = self._pixels[0]
        return math.pow(self._width / nx, 2.0) * 2 * math.pi

    @sensitivity.setter
    def sensitivity(self, value):
        if value not in (None, 1.0):
            raise AttributeError("a CCDArray's sensitivity follows from its pixel area (ccd.pyx:150-151)")

    def _camera_desc(self, pixel_samples):
        nx, ny = self._pixels
        return camera_desc(nx, ny, pixel_samples, None, self.sensitivity, self.to_root(), width=self._width, ccd=True)


class VectorCamera(PinholeCamera):
    """raysect/optical/observer/imaging/vector.pyx:44-156: every pixel has its own origin and viewing direction
    (``pixel_origins`` / ``pixel_directions``: (nx, ny, 3) arrays, or (nx, ny) object arrays of Point3D / Vector3D, in the
    observer's space); pixels off the edge of the image are sub-sampled by interpolating between the diagonal neighbours'
    directions, edge pixels trace their own direction."""

    def __init__(self, pixel_origins, pixel_directions, frame_sampler=None, pipelines=None, sensitivity=None, parent=None,
                 transform=None, name=None):
        def as_array(a):
            a = np.asarray(a)
            if a.dtype == object:
                a = np.array([[[v.x, v.y, v.z] for v in row] for row in a], dtype=np.float64)
            return np.ascontiguousarray(a, dtype=np.float64)
        origins, directions = as_array(pixel_origins), as_array(pixel_directions)
        if origins.ndim != 3 or origins.shape[2] != 3:
            raise ValueError("Pixel arrays must have 2 dimensions.")
        if origins.shape != directions.shape:
            raise ValueError("Pixel arrays must have equal shapes.")
        super().__init__(origins.shape[:2], None, sensitivity, frame_sampler, pipelines, parent, transform, name)
        self.pixel_origins, self.pixel_directions = origins, directions

    def _camera_desc(self, pixel_samples):
        nx, ny = self._pixels
        return camera_desc(nx, ny, pixel_samples, None, self.sensitivity, self.to_root(),
                           vector=(self.pixel_origins, self.pixel_directions))

