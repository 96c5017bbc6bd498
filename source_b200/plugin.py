"""Drop-in plugins for a real Raysect installation.

    from source_b200.plugin import CudaAccelerator, CudaRenderEngine
    world.accelerator = CudaAccelerator()          # World.hit / World.contains on the GPU
    camera.render_engine = CudaRenderEngine()      # camera.observe() renders on the GPU
    camera.frame_sampler = WholeFrameSampler2D()   # optional: "every pixel once" as ONE task instead of nx*ny tuples

``CudaAccelerator`` subclasses ``raysect.core.acceleration.Accelerator`` (accelerator.pxd:37-41; installed
through the ``World.accelerator`` setter, world.pyx:67-70) and ``CudaRenderEngine`` subclasses
``raysect.core.workflow.RenderEngine`` (workflow.py:35-97; installed through ``observer.render_engine``,
observer.pyx:106).  Both flatten the live Raysect scenegraph with ``source_b200.flatten`` -- every AABB,
bounding sphere, matrix and spectral table is obtained from Raysect's own objects -- and call
``libraysect_b200.so``.  Anything outside the supported set raises ``NotImplementedError``: there is no
CPU fallback.

Importing this module requires ``raysect``; the rest of ``source_b200`` does not.
"""
import numpy as np

from raysect.core import Normal3D, Point3D
from raysect.core.acceleration.accelerator import Accelerator
from raysect.core.intersection import Intersection
from raysect.core.workflow import RenderEngine
from raysect.optical.observer.base import FrameSampler2D

from . import _cabi as cabi
from .engine import Accelerator as _DeviceAccelerator
from .engine import camera_desc, default_device, ray_config
from .flatten import flatten_world


class WholeFrameTask(tuple):
    """(nx, ny): the one task WholeFrameSampler2D hands out -- every pixel of the frame, once"""


class WholeFrameSampler2D(FrameSampler2D):
    """The task set of an unmasked ``FullFrameSampler2D`` (every pixel of the frame exactly once, sampler2d.pyx:75-102)
    said in ONE task instead of a shuffled list of nx*ny tuples.  Building and shuffling that list is 0.55 s of host
    time per ``observe()`` of a 1024 x 1024 frame -- a quarter of what the whole call takes once the render runs on the
    device -- and its order means nothing to an engine whose pixel streams are keyed on the pixel.  Only
    ``CudaRenderEngine`` understands the task; the reference's own engines need a per-pixel sampler."""

    def generate_tasks(self, pixels):
        return [WholeFrameTask(pixels)]


class _PrimitiveList:
    """what flatten_world needs from a World: the ordered primitive list"""

    def __init__(self, primitives):
        self.primitives = list(primitives)


class CudaAccelerator(Accelerator):
    """Scene accelerator backed by the device kd-tree traversal + primitive kernels.

    ``hit(ray)`` / ``contains(point)`` keep the reference's scalar contract (1-element batches; meant for
    API parity and tests); ``hit_batch`` / ``contains_batch`` are the forms that use the GPU properly.
    """

    def __init__(self, device=None, backend=None):
        self._device = device
        self._backend_factory = backend     # test hook: callable(FlatScene) -> object with the Accelerator surface
        self._accel = None
        self._primitives = []

    # -- Accelerator interface (accelerator.pyx:32-41) ---------------------------------------------------
    def build(self, primitives):
        if self._accel is not None:
            self._accel.close()
            self._accel = None
        self._primitives = list(primitives)
        if not self._primitives:
            return
        flat = flatten_world(_PrimitiveList(self._primitives))
        if self._backend_factory is not None:
            self._accel = self._backend_factory(flat)
        else:
            self._accel = _DeviceAccelerator(self._device or default_device(), flat)

    def hit(self, ray):
        if self._accel is None:
            return None
        o, d = ray.origin, ray.direction
        r = self._accel.hit_batch([[o.x, o.y, o.z]], [[d.x, d.y, d.z]], [ray.max_distance], geometry=True)
        i = int(r.primitive[0])
        if i < 0:
            return None
        prim = self._primitives[i]
        g = r.geometry[0]
        # (an EncapsulatedPrimitive -- a lens -- re-labels the Intersection of the primitive it hides: the matrices it carries
        # are that primitive's, utility.pyx:74-86)
        from .flatten import _encapsulated
        frame = _encapsulated(prim) or prim
        return Intersection(ray, float(r.distance[0]), prim, Point3D(g[0], g[1], g[2]), Point3D(g[3], g[4], g[5]),
                            Point3D(g[6], g[7], g[8]), Normal3D(g[9], g[10], g[11]), bool(r.exiting[0]),
                            frame.to_local(), frame.to_root())

    def contains(self, point):
        if self._accel is None:
            return []
        cap = 8
        while True:
            count, prims = self._accel.contains_batch([[point.x, point.y, point.z]], cap)
            if count[0] <= cap:
                return [self._primitives[int(i)] for i in prims[0, :count[0]]]
            cap = int(count[0])

    # -- batched forms -----------------------------------------------------------------------------------
    def hit_batch(self, origins, directions, max_distance=None, geometry=False):
        """-> source_b200.engine.HitBatch (primitive = index into the list passed to build())"""
        return self._accel.hit_batch(origins, directions, max_distance, geometry=geometry)

    def contains_batch(self, points, cap=8):
        return self._accel.contains_batch(points, cap)

    @property
    def device_accelerator(self):
        return self._accel


class CudaRenderEngine(RenderEngine):
    """RenderEngine that renders every task of a spectral slice on the GPU(s).

    Contract (workflow.py:78-91, observer.pyx:299-305): ``run(tasks, render, update, render_args=(slice_id,
    template_ray), update_args=(slice_id,))`` is called once per spectral slice; ``render`` is the bound
    ``observer._render_pixel``, so the observer, its world and its pipelines are reachable from it.
    Supported: ``PinholeCamera`` and ``OrthographicCamera`` observers feeding ``SpectralPowerPipeline2D`` / ``SpectralRadiancePipeline2D`` pipelines (the spectral frame
    every other 2-D pipeline is a post-processing of), worlds built from Sphere/Box/Cylinder/Cone/CSG/Mesh
    with Lambert / UniformSurfaceEmitter / UnitySurfaceEmitter / Dielectric / Conductor / RoughConductor / AbsorbingSurface /
    UniformVolumeEmitter / UnityVolumeEmitter materials.

    Random streams: pixel (x, y) of slice k draws from the reference generator seeded with
    ``seed + k*nx*ny + y*nx + x`` (``rng="mt"``), i.e. exactly what a SerialEngine would produce if
    ``raysect.core.math.random.seed`` were called with that value before each pixel task; ``rng="philox"``
    uses counter-based streams instead (faster, statistically equivalent).  After the last slice of an
    ``observe()`` the engine's ``seed`` moves past every stream that call used (``+= passes*n_slices*nx*ny``), so
    the progressive / adaptive loops of the demos (``while not camera.render_complete: camera.observe()`` into an
    accumulating pipeline) draw NEW samples on every pass, as the reference's free-running global stream does;
    assign ``engine.seed`` to restart a sequence.

    ``bulk_update=True`` writes the slice straight into each pipeline's ``frame`` arrays (StatsArray3D)
    with the reference's combine rule instead of calling ``update`` once per pixel.

    ``passes`` > 1 renders the observer's ``pixel_samples`` as that many CONCURRENT passes of
    ``pixel_samples / passes`` samples (rsb_render_passes): pass p of slice k draws from the streams seeded
    ``seed + (p*n_slices + k)*nx*ny + y*nx + x`` and the passes are merged in order with
    StatsArray3D.combine_samples -- bit for bit what ``passes`` observe() calls of ``pixel_samples / passes``
    samples into an empty accumulating pipeline give on the reference (its progressive-render loop,
    demos/cornell_box.py:160-174), with ``passes`` times as many independent pixel streams in flight.

    ``devices=[0, 1, ...]`` (CUDA device indices or ``source_b200.Device`` objects) spreads the tasks of every
    ``run`` over several GPUs from this one process: 16 x 16 pixel tiles are dealt round-robin to the devices, one
    host thread per device drives its own context (ctypes releases the GIL), every device renders into an otherwise
    zero frame and the frames are summed on the host -- exact, because the sum only ever adds zeros, and independent
    of the number of devices, because pixel streams are keyed on the pixel.  (One process per GPU with a single NCCL
    reduce -- ``source_b200.distributed.FrameRenderer`` under torchrun -- is the faster route; this one needs no
    launcher.)
    """

    def __init__(self, seed=1, rng="mt", device=None, bulk_update=True, backend=None, passes=1, devices=None):
        if rng not in ("mt", "philox"):
            raise ValueError("rng must be 'mt' or 'philox'")
        if seed < 1:
            raise ValueError("seed must be >= 1")
        self.seed = int(seed)
        self.rng_mode = cabi.RNG_MT19937_64 if rng == "mt" else cabi.RNG_PHILOX
        self.bulk_update = bulk_update
        self.passes = int(passes)
        if self.passes < 1:
            raise ValueError("passes must be >= 1")
        self._device = device
        self._devices = list(devices) if devices else None
        if self._devices is not None and device is not None:
            raise ValueError("give either device or devices")
        self._backend_factory = backend
        self._accel = None
        self._accel_world = None
        self.ray_count = 0
        self.timing = {}      # seconds spent in the last run(): flatten + upload, device render, frame update (diagnostics)

    def worker_count(self):
        return len(self._devices) if self._devices else 1

    # -- the pipelines' frame buffers are made ready WHILE the device renders -----------------------------------------
    # pipeline.initialise() hands every observe() a freshly allocated StatsArray3D when the pipeline does not accumulate:
    # 1.3 GB of never-touched pageable memory at 1024^2 x 64 bins, into which a device->host copy crawls at ~4 GB/s
    # (0.30 s of a 2.56 s observe()).  While the render call runs the host has nothing else to do, so helper threads
    #   * fault the pages of an EMPTY frame in by storing the zeros it already holds (no driver involved: page-locking
    #     them instead made the copy 10x faster but held a driver lock that cost the concurrent render 0.2 s), or
    #   * page-lock the buffers of a frame that carries samples -- the accumulating frame of a progressive loop, which
    #     lives across observe() calls -- once, and keep them locked until another frame shows up.
    _READY_MIN_BYTES = 64 << 20

    def _frames_ready_start(self, accel, pipelines, empty, frame_of):
        import threading
        arrays = []
        for p in pipelines:
            f = frame_of(p)
            arrays.append((np.asarray(f.mean), empty[id(p)]))
            arrays.append((np.asarray(f.variance), empty[id(p)]))
            arrays.append((np.asarray(f.samples), empty[id(p)]))
        if sum(a.nbytes for a, _ in arrays) < self._READY_MIN_BYTES:
            return []
        pinned = getattr(self, "_pinned", {})
        live = {a.ctypes.data for a, _ in arrays}
        for ptr in [k for k in pinned if k not in live]:       # frames that are gone: release their locks
            pinned.pop(ptr)[0]()
        self._pinned = pinned
        threads = []
        for a, is_empty in arrays:
            if is_empty:
                th = threading.Thread(target=a.fill, args=(0,))
            elif hasattr(accel, "pin") and a.ctypes.data not in pinned:
                def lock(a=a):
                    try:
                        pinned[a.ctypes.data] = (accel.pin(a), a)      # (the array stays alive while it is page-locked)
                    except Exception:        # an optimisation only: pageable copies still work
                        pass
                th = threading.Thread(target=lock)
            else:
                continue
            th.start()
            threads.append(th)
        return threads

    @staticmethod
    def _frames_ready_wait(threads):
        for th in threads:
            th.join()

    def _run_pixel(self, observer, tasks, update, render_args, update_args, update_kwargs):
        """``Pixel`` (optical/observer/nonimaging/pixel.pyx) and ``SightLine`` (nonimaging/sightline.pyx: every ray leaves the
        origin along +z with weight 1 and nothing is drawn -- the edge pixel of a VectorCamera whose arrays say (0, 0, 0) and
        (0, 0, 1)), 0-D observers: their tasks are (samples,) tuples
        (Observer0D._generate_tasks, base/observer.pyx:634-649) all sampling the same rectangle.  The device renders them as the
        pixels of an (n_tasks, 1) frame -- task k of slice s draws from the stream seeded ``seed + s*n_tasks + k`` -- and every
        task's packed result goes through the observer's own ``update`` (0-D pipelines merge task by task, spectral/power.pyx:
        137-150, mono/power.pyx:125-133): spectral pipelines get (mean[bins], variance[bins]), the mono ones the statistics of
        their filtered total (a projection channel of the accumulate kernel)."""
        from raysect.optical.observer import (PowerPipeline0D, RadiancePipeline0D, SightLine, SpectralPowerPipeline0D,
                                              SpectralRadiancePipeline0D)
        slice_id, template = render_args[0], render_args[1]
        pipelines = list(observer.pipelines)
        for p in pipelines:
            if not isinstance(p, (SpectralPowerPipeline0D, SpectralRadiancePipeline0D, PowerPipeline0D, RadiancePipeline0D)):
                raise NotImplementedError("CudaRenderEngine feeds the 0D pipelines SpectralPower, SpectralRadiance, Power, Radiance; "
                                          "got %r (no CPU fallback)" % type(p).__name__)
        if self.passes != 1:
            raise ValueError("a 0-D observer splits its samples into tasks itself (samples_per_task): use passes=1")
        counts = {int(t[0]) for t in tasks}
        if len(counts) != 1:
            raise NotImplementedError("%s with pixel_samples (%d) not a multiple of samples_per_task (%d): the device renders "
                                      "tasks of equal size" % (type(observer).__name__, observer.pixel_samples, observer.samples_per_task))
        spp, n_tasks = counts.pop(), len(tasks)
        accel = self._accelerator_for(observer.root, slice_id)
        if isinstance(accel, list) or not hasattr(accel, "read_slice"):
            raise NotImplementedError("0-D observers render on one device")
        n_slices = len(slice_offsets(observer.spectral_bins, observer.spectral_rays))
        cfg = ray_config(template.bins, template.min_wavelength, template.max_wavelength, template.extinction_prob,
                         template.extinction_min_depth, template.max_depth, template.importance_sampling,
                         template.important_path_weight, template.max_distance)
        spectral = accel.flat.spectral(template.min_wavelength, template.max_wavelength, template.bins)
        delta = (template.max_wavelength - template.min_wavelength) / template.bins
        pixel_sensitivity = float(observer._pixel_sensitivity())

        def is_mono(p):
            return isinstance(p, PowerPipeline0D)                  # (RadiancePipeline0D subclasses it, mono/radiance.pyx:40)

        def sens_of(p):
            radiance = isinstance(p, (SpectralRadiancePipeline0D, RadiancePipeline0D))
            if isinstance(p, RadiancePipeline0D):       # its processor ignores the sensitivity: rides with whichever render there is
                return pixel_sensitivity if any(not isinstance(q, (SpectralRadiancePipeline0D, RadiancePipeline0D)) for q in pipelines) else 1.0
            return 1.0 if radiance else pixel_sensitivity
        packed, rays = {}, 0
        for sensitivity in dict.fromkeys(sens_of(p) for p in pipelines):
            group = [p for p in pipelines if sens_of(p) == sensitivity]
            if isinstance(observer, SightLine):
                directions = np.zeros((n_tasks, 1, 3))
                directions[:, :, 2] = 1.0
                cam = camera_desc(n_tasks, 1, spp, None, sensitivity, observer.to_root(), vector=(np.zeros((n_tasks, 1, 3)), directions))
            else:
                cam = camera_desc(n_tasks, 1, spp, None, sensitivity, observer.to_root(), pixel=(observer.x_width, observer.y_width))
            mono = [p for p in group if is_mono(p)]
            xyz = None
            if mono:
                curves = np.stack([np.asarray(p.filter.sample(template.min_wavelength, template.max_wavelength, template.bins)) for p in mono],
                                  axis=1)[None]
                modes = [cabi.PROJ_RADIANCE if isinstance(p, RadiancePipeline0D) else cabi.PROJ_POWER for p in mono]
                xyz = (curves, np.array([delta]), modes)
            keep = any(not is_mono(p) for p in group)
            kw = dict(xyz=xyz, keep_spectral=keep) if xyz is not None else {}
            rays = accel.render_slices(cam, cfg, [spectral], self.rng_mode, self.seed + slice_id * n_tasks, None, passes=1,
                                       seed_stride=n_tasks, **kw)
            if keep:
                mean, variance = accel.read_slice()                # (n_tasks, 1, bins)
                for p in group:
                    if not is_mono(p):
                        packed[id(p)] = [(np.ascontiguousarray(mean[k, 0]), np.ascontiguousarray(variance[k, 0])) for k in range(n_tasks)]
            for ch, p in enumerate(mono):
                # the per-task statistics of the channel: "merged" into an empty (n_tasks, 1) frame they come back as they are
                m, v, s = np.zeros((n_tasks, 1)), np.zeros((n_tasks, 1)), np.zeros((n_tasks, 1), dtype=np.int32)
                accel.update_proj_frame(ch, m, v, s, frame_is_empty=True)
                packed[id(p)] = [(float(m[k, 0]), float(v[k, 0])) for k in range(n_tasks)]
        self.ray_count += rays
        if slice_id == n_slices - 1:
            self.seed += n_slices * n_tasks
        share, extra = divmod(rays, n_tasks)
        for k, task in enumerate(tasks):
            update((task, [packed[id(p)][k] for p in pipelines], share + (extra if k == 0 else 0)), *update_args, **update_kwargs)

    def _vector_pixels(self, observer):
        """(nx, ny, 3) float64 arrays of a VectorCamera's per-pixel Point3D / Vector3D objects (imaging/vector.pyx:84-86; the
        object arrays are read-only attributes, so the conversion is done once per array pair)"""
        key = (id(observer.pixel_origins), id(observer.pixel_directions))
        cached = getattr(self, "_vector_cache", None)
        if cached is None or cached[0] != key:
            nx, ny = observer.pixel_origins.shape
            o = np.array([[c for pt in row for c in (pt.x, pt.y, pt.z)] for row in observer.pixel_origins], dtype=np.float64).reshape(nx, ny, 3)
            d = np.array([[c for v in row for c in (v.x, v.y, v.z)] for row in observer.pixel_directions], dtype=np.float64).reshape(nx, ny, 3)
            self._vector_cache = (key, o, d, observer.pixel_origins, observer.pixel_directions)    # (the arrays stay alive with their ids)
        return self._vector_cache[1], self._vector_cache[2]

    def _accelerator_for(self, world, slice_id):
        # a new observe() starts at slice 0: re-flatten there so scene edits between renders are picked up
        if self._accel is None or self._accel_world is not world or slice_id == 0:
            for a in (self._accel if isinstance(self._accel, list) else [self._accel] if self._accel else []):
                a.close()
            flat = flatten_world(world)
            if self._devices:
                from .engine import Device
                self._devices = [d if hasattr(d, "ctx") or self._backend_factory else Device(int(d)) for d in self._devices]
                self._accel = [self._backend_factory(flat) if self._backend_factory is not None else _DeviceAccelerator(d, flat)
                               for d in self._devices]
                if self.bulk_update and all(hasattr(a, "render_slices") for a in self._accel) and hasattr(self._accel[0], "gather_from"):
                    # whole-slice device path on every GPU; the first one gathers the others' rows over peer memory
                    from .engine import DeviceGroup
                    self._accel = DeviceGroup(self._accel)
            elif self._backend_factory is not None:
                self._accel = self._backend_factory(flat)
            else:
                self._accel = _DeviceAccelerator(self._device or default_device(), flat)
            self._accel_world = world
        return self._accel

    @staticmethod
    def _render_on_devices(accels, pix, tile, render_one):
        """tiles of 16 x 16 pixels dealt round-robin to the devices; one thread per device; frames summed on the host"""
        from concurrent.futures import ThreadPoolExecutor
        n = len(accels)
        tile_id = (pix[:, 0] // tile) * 65536 + (pix[:, 1] // tile)
        order = {t: k for k, t in enumerate(sorted(set(tile_id.tolist())))}
        owner = np.array([order[t] % n for t in tile_id.tolist()], dtype=np.int64)
        jobs = [(a, np.ascontiguousarray(pix[owner == k])) for k, a in enumerate(accels)]
        jobs = [(a, p) for a, p in jobs if len(p)]
        with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as pool:
            parts = list(pool.map(lambda job: render_one(job[0], job[1]), jobs))
        mean, variance, rays = parts[0]
        for m, v, r in parts[1:]:
            mean += m           # exact: every frame is zero outside its own tiles
            variance += v
            rays += r
        return mean, variance, rays

    def run(self, tasks, render, update, render_args=(), render_kwargs={}, update_args=(), update_kwargs={}):
        from raysect.optical.observer import (BayerPipeline2D, CCDArray, OrthographicCamera, PinholeCamera, PowerPipeline2D,
                                              RadiancePipeline2D, RGBPipeline2D, SpectralPowerPipeline2D, SpectralRadiancePipeline2D,
                                              VectorCamera)
        observer = getattr(render, "__self__", None)
        from raysect.optical.observer import Pixel, SightLine
        if isinstance(observer, (Pixel, SightLine)):
            return self._run_pixel(observer, tasks, update, render_args, update_args, update_kwargs)
        if not isinstance(observer, (PinholeCamera, OrthographicCamera, CCDArray, VectorCamera)):
            raise NotImplementedError("CudaRenderEngine renders PinholeCamera, OrthographicCamera, CCDArray and VectorCamera observers; "
                                      "got %r (no CPU fallback)" % type(observer).__name__)
        pipelines = list(observer.pipelines)
        for p in pipelines:
            if not isinstance(p, (SpectralPowerPipeline2D, SpectralRadiancePipeline2D, RGBPipeline2D, PowerPipeline2D, RadiancePipeline2D,
                                  BayerPipeline2D)):
                raise NotImplementedError("CudaRenderEngine feeds the 2D pipelines of raysect.optical.observer (SpectralPower, "
                                          "SpectralRadiance, RGB, Power, Radiance, Bayer); got %r (no CPU fallback)" % type(p).__name__)
        # pipelines whose pixel processors PROJECT every sample's spectrum on curves (CIE XYZ; filters): done on the device
        rgb = [p for p in pipelines if isinstance(p, (RGBPipeline2D, PowerPipeline2D, RadiancePipeline2D, BayerPipeline2D))]

        def frame_of(p):
            return p.xyz_frame if isinstance(p, RGBPipeline2D) else p.frame
        import time
        slice_id, template = render_args[0], render_args[1]
        world = observer.root
        t0 = time.perf_counter()
        accel = self._accelerator_for(world, slice_id)
        self.timing = {"flatten_upload_s": time.perf_counter() - t0, "render_s": 0.0, "update_s": 0.0}
        nx, ny = observer.pixels
        if observer.pixel_samples % self.passes:
            raise ValueError("the observer's pixel_samples (%d) must be a multiple of the engine's passes (%d)"
                             % (observer.pixel_samples, self.passes))
        def camera_for(sensitivity):
            if isinstance(observer, VectorCamera):
                return camera_desc(nx, ny, observer.pixel_samples // self.passes, None, sensitivity, observer.to_root(),
                                   vector=self._vector_pixels(observer))
            if isinstance(observer, CCDArray):
                return camera_desc(nx, ny, observer.pixel_samples // self.passes, None, sensitivity, observer.to_root(),
                                   width=observer.width, ccd=True)
            if isinstance(observer, OrthographicCamera):
                return camera_desc(nx, ny, observer.pixel_samples // self.passes, None, sensitivity,
                                   observer.to_root(), width=observer.width)
            return camera_desc(nx, ny, observer.pixel_samples // self.passes, observer.fov, sensitivity,
                               observer.to_root())
        cfg = ray_config(template.bins, template.min_wavelength, template.max_wavelength, template.extinction_prob,
                         template.extinction_min_depth, template.max_depth, template.importance_sampling,
                         template.important_path_weight, template.max_distance)
        spectral = (accel[0] if isinstance(accel, list) else accel).flat.spectral(
            template.min_wavelength, template.max_wavelength, template.bins)
        # The power pipeline's pixel processor scales every sample by the pixel sensitivity (power.pyx:478-481), the
        # radiance pipeline's does not (radiance.pyx:256-260) -- the same as a sensitivity of exactly 1.0.  One render
        # per distinct sensitivity; the pixel streams are keyed on the pixel, so both see the very same paths.
        # (pinhole / orthographic: the observer's `sensitivity`; CCD: pixel area x 2 pi, ccd.pyx:150-151)
        pixel_sensitivity = float(observer._pixel_sensitivity(0, 0))

        def sens_of(p):
            if isinstance(p, SpectralRadiancePipeline2D):
                return 1.0
            if isinstance(p, RadiancePipeline2D):
                # (its processor ignores the sensitivity, mono/radiance.pyx:184-195: it rides with whichever render there is)
                return pixel_sensitivity if any(not isinstance(q, (SpectralRadiancePipeline2D, RadiancePipeline2D)) for q in pipelines) else 1.0
            return pixel_sensitivity
        kw = dict(passes=self.passes, seed_stride=observer.spectral_rays * nx * ny) if self.passes > 1 else {}
        n_slices = len(slice_offsets(observer.spectral_bins, observer.spectral_rays))
        offset = slice_offsets(observer.spectral_bins, observer.spectral_rays)[slice_id]
        seed = self.seed + slice_id * nx * ny
        fast = self.bulk_update and not isinstance(accel, list) and hasattr(accel, "render_slice")
        # a stock FullFrameSampler2D without a mask lists every pixel exactly once (sampler2d.pyx:75-102), in shuffled
        # order; pixel streams are keyed on the pixel, so "the whole frame" says the same in zero bytes
        from itertools import chain
        from raysect.optical.observer import FullFrameSampler2D
        if len(tasks) == 1 and isinstance(tasks[0], WholeFrameTask):
            if tuple(tasks[0]) != (nx, ny):
                raise ValueError("the whole-frame task was generated for another frame size")
            pix = None if fast else np.stack(np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij"), axis=-1).reshape(-1, 2).astype(np.int32)
        elif fast and type(observer.frame_sampler) is FullFrameSampler2D and len(tasks) == nx * ny:
            pix = None
        else:
            pix = np.fromiter(chain.from_iterable(tasks), dtype=np.int32, count=2 * len(tasks)).reshape(-1, 2)
        # Every spectral slice at once (observer.pyx:299-305 calls run() once per slice, one after the other): slices are
        # independent pixel streams, so the first run() of an observe() renders ALL of them concurrently
        # (rsb_render_slices; slice k draws from the seeds the per-slice path uses) and merges the whole frame into the
        # pipelines; the run() calls for slices 1.. find their work done.
        all_slices = None
        if fast and (n_slices > 1 or rgb) and hasattr(accel, "render_slices"):
            spec = observer._slice_spectrum()
            if len({sl.bins for sl in spec}) == 1:
                all_slices = spec
        if rgb and all_slices is None:
            # XYZPixelProcessor's per-sample projection (rgb.pyx:550-558) runs inside the device's accumulate kernel and the
            # slices of a pixel are summed before they enter the frame (rgb.pyx:259-265): the all-slices device path only
            raise NotImplementedError("RGBPipeline2D / PowerPipeline2D / RadiancePipeline2D need CudaRenderEngine's whole-slice device "
                                      "path: bulk_update=True, spectral slices of equal size (spectral_bins divisible by spectral_rays)")
        if all_slices is not None and slice_id > 0:
            if slice_id == n_slices - 1:
                self.seed += self.passes * n_slices * nx * ny
            return
        frames, rays = {}, 0
        for sensitivity in dict.fromkeys(sens_of(p) for p in pipelines):
            cam = camera_for(sensitivity)
            if all_slices is not None:
                cfg0 = ray_config(all_slices[0].bins, all_slices[0].min_wavelength, all_slices[0].max_wavelength, template.extinction_prob,
                                  template.extinction_min_depth, template.max_depth, template.importance_sampling,
                                  template.important_path_weight, template.max_distance)
                spectrals = [accel.flat.spectral(sl.min_wavelength, sl.max_wavelength, sl.bins) for sl in all_slices]
                t0 = time.perf_counter()
                group = [p for p in pipelines if sens_of(p) == sensitivity]
                empty = {id(p): not np.asarray(frame_of(p).samples).any() for p in group}
                ready = self._frames_ready_start(accel, group, empty, frame_of)
                xyz, channel_of = None, {}
                if any(p in rgb for p in group):
                    # the curves the pipelines' initialise() hands their pixel processors (rgb.pyx:232-233: resample_ciexyz;
                    # mono/power.pyx:501: filter.sample_mv), one set per slice, and the delta_wavelength of the slice's
                    # Spectrum (spectrum.pyx:132: (max - min) / bins)
                    from raysect.optical.colour import resample_ciexyz
                    curves, modes = [], []
                    for p in group:
                        if p not in rgb:
                            continue
                        channel_of[id(p)] = len(modes)
                        if isinstance(p, RGBPipeline2D):
                            curves.append(np.stack([np.asarray(resample_ciexyz(sl.min_wavelength, sl.max_wavelength, sl.bins)) for sl in all_slices]))
                            modes += [cabi.PROJ_XYZ] * 3
                        elif isinstance(p, BayerPipeline2D):
                            # one PowerPixelProcessor per filter (bayer.pyx:322-330); the mosaic picks per pixel at update time
                            curves.append(np.stack([np.stack([np.asarray(f.sample(sl.min_wavelength, sl.max_wavelength, sl.bins))
                                                              for f in (p.red_filter, p.green_filter, p.blue_filter)], axis=1)
                                                    for sl in all_slices]))
                            modes += [cabi.PROJ_POWER] * 3
                        else:
                            curves.append(np.stack([np.asarray(p.filter.sample(sl.min_wavelength, sl.max_wavelength, sl.bins))[:, None]
                                                    for sl in all_slices]))
                            # (RadiancePipeline2D subclasses PowerPipeline2D, mono/radiance.pyx:121)
                            modes.append(cabi.PROJ_RADIANCE if isinstance(p, RadiancePipeline2D) else cabi.PROJ_POWER)
                    if len(modes) > cabi.PROJ_MAX:
                        raise NotImplementedError("more than %d projection channels in one observe() (an RGB pipeline takes 3, a mono "
                                                  "pipeline 1)" % cabi.PROJ_MAX)
                    xyz = (np.concatenate(curves, axis=2), np.array([(sl.max_wavelength - sl.min_wavelength) / sl.bins for sl in all_slices]),
                           modes)
                xkw = dict(xyz=xyz, keep_spectral=any(p not in rgb for p in group)) if xyz is not None else {}
                rays = accel.render_slices(cam, cfg0, spectrals, self.rng_mode, self.seed, pix, passes=self.passes, **xkw)
                self._frames_ready_wait(ready)
                t1 = time.perf_counter()
                for p in sorted(group, key=lambda q: q not in rgb):       # projected frames first (DeviceGroup: before the rows are gathered)
                    f = frame_of(p)
                    fm, fv, fs = np.asarray(f.mean), np.asarray(f.variance), np.asarray(f.samples)
                    if isinstance(p, BayerPipeline2D):
                        accel.update_bayer_frame(channel_of[id(p)], fm, fv, fs, frame_is_empty=empty[id(p)])
                    elif p in rgb:
                        accel.update_proj_frame(channel_of[id(p)], fm, fv, fs, frame_is_empty=empty[id(p)])
                    else:
                        accel.update_frame(fm, fv, fs, 0, frame_is_empty=empty[id(p)])
                self.timing["render_s"] += t1 - t0
                self.timing["update_s"] += time.perf_counter() - t1
            elif fast:
                # one device render per distinct sensitivity, kept on the device and merged into every pipeline frame that
                # wants it with the reference's combine rule (power.pyx:424-437 -> statsarray.pyx:780-857) -- no per-pixel
                # Python, no host-side gather / scatter
                t0 = time.perf_counter()
                empty = {id(p): not np.asarray(p.frame.samples)[:, :, offset:offset + template.bins].any() for p in pipelines}
                # zeros may only be stored over a frame that is empty in EVERY slice
                whole = {id(p): slice_id == 0 and not np.asarray(p.frame.samples).any() for p in pipelines}
                ready = self._frames_ready_start(accel, [p for p in pipelines if sens_of(p) == sensitivity], whole, frame_of)
                rays = accel.render_slice(cam, cfg, spectral, self.rng_mode, seed, pix, **kw)
                self._frames_ready_wait(ready)
                t1 = time.perf_counter()
                for p in pipelines:
                    if sens_of(p) == sensitivity:
                        fm, fv, fs = np.asarray(p.frame.mean), np.asarray(p.frame.variance), np.asarray(p.frame.samples)
                        accel.update_frame(fm, fv, fs, offset, frame_is_empty=empty[id(p)])
                self.timing["render_s"] += t1 - t0
                self.timing["update_s"] += time.perf_counter() - t1
            elif isinstance(accel, list):
                mean, variance, rays = self._render_on_devices(
                    accel, pix, 16, lambda a, p: a.render(cam, cfg, a.flat.spectral(
                        template.min_wavelength, template.max_wavelength, template.bins), self.rng_mode, seed, p, **kw))
                frames[sensitivity] = (mean, variance)
            else:
                mean, variance, rays = accel.render(cam, cfg, spectral, self.rng_mode, seed, pix, **kw)
                frames[sensitivity] = (mean, variance)
        self.ray_count += rays
        if slice_id == n_slices - 1:
            self.seed += self.passes * n_slices * nx * ny     # the next observe() draws from fresh streams
        if fast:
            observer._update_statistics(rays)
            return
        per_pipeline = [frames[sens_of(p)] for p in pipelines]
        if self.bulk_update:
            for p, (mean, variance) in zip(pipelines, per_pipeline):
                self._bulk_update([p], pix, offset, mean, variance, observer.pixel_samples)
            # statistics hook of the observer: one update carrying the whole ray count
            observer._update_statistics(rays)
            return
        share, extra = divmod(rays, len(pix))
        for k, (x, y) in enumerate(pix):
            result = ((int(x), int(y)), [(m[x, y].copy(), v[x, y].copy()) for m, v in per_pipeline],
                      share + (extra if k == 0 else 0))
            update(result, *update_args, **update_kwargs)

    @staticmethod
    def _bulk_update(pipelines, pix, offset, mean, variance, samples):
        """SpectralPowerPipeline2D.update for every pixel at once (power.pyx:424-437 ->
        StatsArray3D.combine_samples, statsarray.pyx:623-667)."""
        from .observer import combine_samples
        xs, ys = pix[:, 0], pix[:, 1]
        sl = slice(offset, offset + mean.shape[2])
        for p in pipelines:
            frame = p.frame
            fm, fv, fs = np.asarray(frame.mean), np.asarray(frame.variance), np.asarray(frame.samples)
            mt, vt, nt = combine_samples(fm[xs, ys, sl], fv[xs, ys, sl], fs[xs, ys, sl],
                                         mean[xs, ys], np.maximum(variance[xs, ys], 0.0), samples)
            fm[xs, ys, sl] = mt
            fv[xs, ys, sl] = vt
            fs[xs, ys, sl] = nt


def slice_offsets(spectral_bins, spectral_rays):
    """Bin offset of every spectral slice, cut as observer._slice_spectrum does (observer.pyx:311-340)."""
    current, start, offsets = 0, 0, []
    while start < spectral_bins:
        current += spectral_bins / spectral_rays
        end = round(current)
        offsets.append(start)
        start = end
    return offsets
