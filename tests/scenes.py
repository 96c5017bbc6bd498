"""Scene definitions shared by the golden-vector generator (run against the real reference) and the
parity tests (run against source_b200's mirror object model).  Each builder takes an ``api`` namespace
exposing Raysect's names (World, Node, Sphere, Box, ..., translate, rotate, Point3D, materials), so the
SAME construction code produces the reference scene and the device scene.
"""
import numpy as np

# ---- spectra of the physical Cornell box, as tabulated in the reference's demos/cornell_box.py:28-57 ----
CB_WAVELENGTHS = list(range(400, 704, 4))
CB_WHITE = [0.343, 0.445, 0.551, 0.624, 0.665, 0.687, 0.708, 0.723, 0.715, 0.71, 0.745, 0.758, 0.739, 0.767, 0.777, 0.765,
            0.751, 0.745, 0.748, 0.729, 0.745, 0.757, 0.753, 0.75, 0.746, 0.747, 0.735, 0.732, 0.739, 0.734, 0.725, 0.721,
            0.733, 0.725, 0.732, 0.743, 0.744, 0.748, 0.728, 0.716, 0.733, 0.726, 0.713, 0.74, 0.754, 0.764, 0.752, 0.736,
            0.734, 0.741, 0.74, 0.732, 0.745, 0.755, 0.751, 0.744, 0.731, 0.733, 0.744, 0.731, 0.712, 0.708, 0.729, 0.73,
            0.727, 0.707, 0.703, 0.729, 0.75, 0.76, 0.751, 0.739, 0.724, 0.73, 0.74, 0.737]
CB_GREEN = [0.092, 0.096, 0.098, 0.097, 0.098, 0.095, 0.095, 0.097, 0.095, 0.094, 0.097, 0.098, 0.096, 0.101, 0.103, 0.104,
            0.107, 0.109, 0.112, 0.115, 0.125, 0.14, 0.16, 0.187, 0.229, 0.285, 0.343, 0.39, 0.435, 0.464, 0.472, 0.476, 0.481,
            0.462, 0.447, 0.441, 0.426, 0.406, 0.373, 0.347, 0.337, 0.314, 0.285, 0.277, 0.266, 0.25, 0.23, 0.207, 0.186,
            0.171, 0.16, 0.148, 0.141, 0.136, 0.13, 0.126, 0.123, 0.121, 0.122, 0.119, 0.114, 0.115, 0.117, 0.117, 0.118, 0.12,
            0.122, 0.128, 0.132, 0.139, 0.144, 0.146, 0.15, 0.152, 0.157, 0.159]
CB_RED = [0.04, 0.046, 0.048, 0.053, 0.049, 0.05, 0.053, 0.055, 0.057, 0.056, 0.059, 0.057, 0.061, 0.061, 0.06, 0.062, 0.062,
          0.062, 0.061, 0.062, 0.06, 0.059, 0.057, 0.058, 0.058, 0.058, 0.056, 0.055, 0.056, 0.059, 0.057, 0.055, 0.059,
          0.059, 0.058, 0.059, 0.061, 0.061, 0.063, 0.063, 0.067, 0.068, 0.072, 0.08, 0.09, 0.099, 0.124, 0.154, 0.192,
          0.255, 0.287, 0.349, 0.402, 0.443, 0.487, 0.513, 0.558, 0.584, 0.62, 0.606, 0.609, 0.651, 0.612, 0.61, 0.65, 0.638,
          0.627, 0.62, 0.63, 0.628, 0.642, 0.639, 0.657, 0.639, 0.635, 0.642]
CB_LIGHT = ([400, 500, 600, 700], [0.0, 8.0, 15.6, 18.4])


def cornell_box(api, glass=True, extra=None):
    """World of demos/cornell_box.py:63-121: 5 zero-thickness Lambert walls, 1 emitter box, a glass box
    and a glass sphere (N-BK7).  ``extra(api, world)`` may add more primitives (e.g. the bunny)."""
    a = api
    white = a.InterpolatedSF(CB_WAVELENGTHS, CB_WHITE)
    red = a.InterpolatedSF(CB_WAVELENGTHS, CB_RED)
    green = a.InterpolatedSF(CB_WAVELENGTHS, CB_GREEN)
    light_spectrum = a.InterpolatedSF(*CB_LIGHT)
    world = a.World()
    enclosure = a.Node(world)
    lo, hi = a.Point3D(-1, -1, 0), a.Point3D(1, 1, 0)
    a.Box(lo, hi, parent=enclosure, transform=a.translate(0, 0, 1) * a.rotate(0, 0, 0), material=a.Lambert(white))
    a.Box(lo, hi, parent=enclosure, transform=a.translate(0, -1, 0) * a.rotate(0, -90, 0), material=a.Lambert(white))
    a.Box(lo, hi, parent=enclosure, transform=a.translate(0, 1, 0) * a.rotate(0, 90, 0), material=a.Lambert(white))
    a.Box(lo, hi, parent=enclosure, transform=a.translate(1, 0, 0) * a.rotate(-90, 0, 0), material=a.Lambert(red))
    a.Box(lo, hi, parent=enclosure, transform=a.translate(-1, 0, 0) * a.rotate(90, 0, 0), material=a.Lambert(green))
    a.Box(a.Point3D(-0.4, -0.4, -0.01), a.Point3D(0.4, 0.4, 0.0), parent=enclosure,
          transform=a.translate(0, 1, 0) * a.rotate(0, 90, 0), material=a.UniformSurfaceEmitter(light_spectrum, 2))
    if glass:
        a.Box(a.Point3D(-0.4, 0, -0.4), a.Point3D(0.3, 1.4, 0.3), parent=world,
              transform=a.translate(0.4, -1 + 1e-6, 0.4) * a.rotate(30, 0, 0), material=a.schott("N-BK7"))
        a.Sphere(0.4, parent=world, transform=a.translate(-0.4, -0.6 + 1e-6, -0.4) * a.rotate(0, 0, 0),
                 material=a.schott("N-BK7"))
    if extra is not None:
        extra(a, world)
    return world


GOLD_WAVELENGTHS = [300.0, 400.0, 450.0, 500.0, 550.0, 600.0, 700.0, 800.0]
GOLD_N = [1.53, 1.47, 1.38, 0.97, 0.43, 0.25, 0.16, 0.15]      # gold-like complex index n + ik
GOLD_K = [1.89, 1.95, 1.92, 1.87, 2.46, 2.99, 3.95, 4.85]


def metal_scene(api):
    """Cornell box with metal: a Conductor sphere and a smooth-shaded Conductor mesh (interpolated normals: the side
    of the reflection is chosen from n.d, conductor.pyx:108-118), the glass box kept, and a UnitySurfaceEmitter."""
    a = api
    gold = a.Conductor(a.InterpolatedSF(GOLD_WAVELENGTHS, GOLD_N), a.InterpolatedSF(GOLD_WAVELENGTHS, GOLD_K))
    verts, tris, normals = icosphere(2, radius=0.3, bumps=0.2)

    def extra(a, w):
        a.Sphere(0.35, parent=w, transform=a.translate(-0.45, -0.65, -0.3), material=gold)
        a.Mesh(verts, tris, normals, smoothing=True, closed=True, parent=w,
               transform=a.translate(0.3, 0.35, -0.2) * a.rotate(15, 25, 5), material=gold)
        a.Sphere(0.12, parent=w, transform=a.translate(-0.6, 0.6, 0.2), material=a.UnitySurfaceEmitter())
    world = cornell_box(a, glass=False, extra=extra)
    a.Box(a.Point3D(-0.4, 0, -0.4), a.Point3D(0.3, 1.0, 0.3), parent=world,
          transform=a.translate(0.45, -1 + 1e-6, 0.45) * a.rotate(30, 0, 0), material=a.schott("N-BK7"))
    return world


def rough_metal_scene(api):
    """Cornell box with RoughConductor objects (GGX microfacet metal) of three roughnesses: sphere, box, smooth-shaded mesh"""
    a = api
    n, k = a.InterpolatedSF(GOLD_WAVELENGTHS, GOLD_N), a.InterpolatedSF(GOLD_WAVELENGTHS, GOLD_K)
    verts, tris, normals = icosphere(2, radius=0.3, bumps=0.2)

    def extra(a, w):
        a.Sphere(0.35, parent=w, transform=a.translate(-0.45, -0.65, -0.3), material=a.RoughConductor(n, k, 0.25))
        a.Box(a.Point3D(-0.3, 0, -0.3), a.Point3D(0.3, 0.9, 0.3), parent=w,
              transform=a.translate(0.45, -1 + 1e-6, 0.4) * a.rotate(25, 0, 0), material=a.RoughConductor(n, k, 0.8))
        a.Mesh(verts, tris, normals, smoothing=True, closed=True, parent=w,
               transform=a.translate(0.2, 0.4, -0.3) * a.rotate(15, 25, 5), material=a.RoughConductor(n, k, 0.1))
    return cornell_box(a, glass=False, extra=extra)


def volume_scene(api, fog=True):
    """Cornell box with emitting volumes (HomogeneousVolumeEmitter on a NullSurface): a glowing sphere overlapped by
    a glass sphere, a unity-emission cylinder, and (fog) a faint emitting box that contains the whole room AND the
    camera, so that every segment of every path integrates emission and paths that end dark still carry light."""
    a = api
    glow = a.UniformVolumeEmitter(a.InterpolatedSF(*CB_LIGHT), 0.8)

    def extra(a, w):
        a.Sphere(0.3, parent=w, transform=a.translate(-0.4, -0.5, -0.2), material=glow)
        a.Sphere(0.25, parent=w, transform=a.translate(-0.15, -0.6, -0.35), material=a.schott("N-BK7"))
        a.Cylinder(0.15, 0.7, parent=w, transform=a.translate(0.45, -0.2, 0.1) * a.rotate(20, 70, 0), material=a.UnityVolumeEmitter())
        if fog:
            a.Box(a.Point3D(-1.5, -1.5, -4.0), a.Point3D(1.5, 1.5, 1.5), parent=w,
                  material=a.UniformVolumeEmitter(a.ConstantSF(1.0), 0.01))
    return cornell_box(a, glass=False, extra=extra)


def cornell_camera(api, world, pixels=(128, 128), samples=1, bins=15, spectral_rays=1, min_depth=3, max_depth=500,
                   extinction=0.01, path_weight=0.25, importance=True, fov=None, sensitivity=None):
    """Camera of demos/cornell_box.py:147-156 with a SpectralPowerPipeline2D and a full-frame sampler."""
    a = api
    pipeline = a.SpectralPowerPipeline2D()
    camera = a.PinholeCamera(pixels, parent=world, transform=a.translate(0, 0, -3.3) * a.rotate(0, 0, 0),
                             pipelines=[pipeline], frame_sampler=a.FullFrameSampler2D())
    camera.spectral_rays = 1             # (rays <= bins is checked by both setters: settle the bins first)
    camera.spectral_bins = bins
    camera.spectral_rays = spectral_rays
    camera.pixel_samples = samples
    camera.ray_importance_sampling = importance
    camera.ray_important_path_weight = path_weight
    camera.ray_max_depth = max_depth
    camera.ray_extinction_min_depth = min_depth
    camera.ray_extinction_prob = extinction
    camera.quiet = True
    if fov is not None:
        camera.fov = fov
    if sensitivity is not None:
        camera.sensitivity = sensitivity
    return camera, pipeline


def orthographic_camera(api, world, pixels=(20, 16), width=2.4, samples=3, bins=8, spectral_rays=1, transform=None):
    """OrthographicCamera (imaging/orthographic.pyx) looking into the Cornell box: parallel rays, weight 1"""
    a = api
    pipeline = a.SpectralPowerPipeline2D()
    camera = a.OrthographicCamera(pixels, width, parent=world, pipelines=[pipeline], frame_sampler=a.FullFrameSampler2D(),
                                  transform=transform if transform is not None else a.translate(0.05, -0.1, -3.3) * a.rotate(4, -3, 2))
    camera.spectral_rays = 1             # (rays <= bins is checked by both setters: settle the bins first)
    camera.spectral_bins = bins
    camera.spectral_rays = spectral_rays
    camera.pixel_samples = samples
    camera.ray_importance_sampling = True
    camera.ray_important_path_weight = 0.25
    camera.ray_max_depth = 500
    camera.ray_extinction_min_depth = 3
    camera.ray_extinction_prob = 0.01
    camera.quiet = True
    return camera, pipeline


def random_spheres(api, n, seed=7, uniform=None):
    """BASELINE config 5: n spheres, centres uniform in [-1,1]^3, radii uniform in [0.01,0.03]; the draws
    come from ``uniform`` (the reference RNG after seed(7) when generating goldens) or numpy."""
    a = api
    if uniform is None:
        rng = np.random.default_rng(seed)
        uniform = lambda: float(rng.random())
    world = a.World()
    mat = a.AbsorbingSurface()
    for _ in range(n):
        c = [2 * uniform() - 1 for _ in range(3)]
        r = 0.01 + 0.02 * uniform()
        a.Sphere(r, world, a.translate(*c), mat)
    return world


def primitive_zoo(api):
    """One of every analytic primitive, rotated/translated, plus CSG combinations of every operator
    including the nested prism/screen shapes of demos/prism.py:21-41,77-98."""
    a = api
    world = a.World()
    m = a.AbsorbingSurface()
    a.Sphere(0.5, world, a.translate(-2.0, 0.3, 0.2) * a.rotate(10, 20, 30), m)
    a.Box(a.Point3D(-0.4, -0.3, -0.2), a.Point3D(0.5, 0.6, 0.7), world, a.translate(-0.7, -0.2, 0.1) * a.rotate(25, -15, 40), m)
    a.Cylinder(0.35, 1.1, world, a.translate(0.6, -0.5, 0.0) * a.rotate(-35, 60, 5), m)
    a.Cone(0.45, 0.9, world, a.translate(1.8, -0.4, 0.3) * a.rotate(15, -70, 0), m)
    # every CSG operator on a sphere/box pair
    a.Union(a.Sphere(0.4, transform=a.translate(0.2, 0, 0)), a.Box(a.Point3D(-0.3, -0.3, -0.3), a.Point3D(0.3, 0.3, 0.3)),
            world, a.translate(-2.0, 1.5, 0.0) * a.rotate(20, 10, 0), m)
    a.Intersect(a.Sphere(0.45), a.Box(a.Point3D(-0.35, -0.35, -0.35), a.Point3D(0.35, 0.35, 0.35)),
                world, a.translate(-0.7, 1.5, 0.0) * a.rotate(-20, 30, 10), m)
    a.Subtract(a.Box(a.Point3D(-0.4, -0.4, -0.4), a.Point3D(0.4, 0.4, 0.4)), a.Sphere(0.5),
               world, a.translate(0.6, 1.5, 0.0) * a.rotate(40, 20, -10), m)
    # cylinder/cone operands
    a.Intersect(a.Cylinder(0.3, 1.0, transform=a.rotate(0, 90, 0) * a.translate(0, 0, -0.5)),
                a.Cone(0.5, 0.8, transform=a.translate(0, 0, -0.3)), world, a.translate(1.9, 1.5, 0.0) * a.rotate(10, 10, 10), m)
    # nested: prism of demos/prism.py (Subtract(Subtract(Box, Box), Box))
    prism = a.Subtract(
        a.Subtract(a.Box(a.Point3D(-0.5, 0, -0.5), a.Point3D(0.5, 0.8, 0.5)),
                   a.Box(a.Point3D(-0.5, -0.1, -0.1), a.Point3D(1.5, 1.0, 0.9), transform=a.translate(0.5, 0, 0) * a.rotate(30, 0, 0))),
        a.Box(a.Point3D(-1.5, -0.1, -0.1), a.Point3D(0.5, 1.0, 0.9), transform=a.translate(-0.5, 0, 0) * a.rotate(-30, 0, 0)),
        world, a.translate(-1.2, -1.9, 0.2) * a.rotate(5, 0, 0), m)
    # nested: slotted screen (Intersect(Box, Subtract(Cylinder, Cylinder)))
    a.Intersect(a.Box(a.Point3D(-0.5, -0.4, -1), a.Point3D(0.5, 0.4, 1)),
                a.Subtract(a.Cylinder(0.6, 0.5, transform=a.rotate(0, 90, 0) * a.translate(0, 0, -0.25)),
                           a.Cylinder(0.45, 0.7, transform=a.rotate(0, 90, 0) * a.translate(0, 0, -0.35))),
                world, a.translate(0.9, -1.8, 0.0) * a.rotate(-15, 25, 0), m)
    # Union of unions (depth 2, 4 leaves)
    a.Union(a.Union(a.Sphere(0.25, transform=a.translate(-0.3, 0, 0)), a.Sphere(0.25, transform=a.translate(0.3, 0, 0))),
            a.Union(a.Cylinder(0.1, 1.2, transform=a.translate(0, 0, -0.6)), a.Cone(0.3, 0.5, transform=a.translate(0, 0.2, 0))),
            world, a.translate(2.4, -1.8, 0.4) * a.rotate(30, 30, 30), m)
    return world


def edge_scene(api):
    """UNROTATED primitives at integer positions, so that rays can be aimed exactly at faces, edges, corners, tips,
    tangent lines and shared faces: every `==`, `<=` and zero-discriminant branch of the hit code sees exact input."""
    a = api
    world = a.World()
    m = a.AbsorbingSurface()
    a.Box(a.Point3D(-1, -1, -1), a.Point3D(1, 1, 1), world, a.translate(0, 0, 0), m)
    a.Sphere(0.5, world, a.translate(3, 0, 0), m)
    a.Cylinder(0.5, 1.0, world, a.translate(0, 3, 0), m)
    a.Cone(0.5, 1.0, world, a.translate(0, -3, 0), m)
    a.Box(a.Point3D(5, -0.5, -0.5), a.Point3D(6, 0.5, 0.5), world, a.translate(0, 0, 0), m)      # shares the face x = 6 ...
    a.Box(a.Point3D(6, -0.5, -0.5), a.Point3D(7, 0.5, 0.5), world, a.translate(0, 0, 0), m)      # ... with this one
    a.Box(a.Point3D(-4, -1, 0), a.Point3D(-3, 1, 0), world, a.translate(0, 0, 0), m)              # zero thickness
    a.Subtract(a.Box(a.Point3D(-0.5, -0.5, -0.5), a.Point3D(0.5, 0.5, 0.5)), a.Sphere(0.5), world, a.translate(0, 0, 4), m)
    a.Union(a.Box(a.Point3D(-0.5, -0.5, -0.5), a.Point3D(0.5, 0.5, 0.5)),
            a.Box(a.Point3D(0.5, -0.5, -0.5), a.Point3D(1.5, 0.5, 0.5)), world, a.translate(3, 0, 4), m)     # coplanar operands
    return world


def scaled_scene(api):
    """Non-rigid transforms (anisotropic scale and shear): ray directions are transformed without renormalisation
    (core/ray.pxd), normals through the inverse transpose (normal.pyx:222-248), bounding boxes from the 8 transformed
    corners -- every primitive type, a CSG tree and a mesh under AffineMatrix3D matrices that are not rotations."""
    a = api
    world = a.World()
    lam = a.Lambert(a.ConstantSF(0.6))

    def M(sx, sy, sz, shear=0.0):
        return a.AffineMatrix3D([[sx, shear, 0, 0], [0, sy, 0.5 * shear, 0], [0, 0, sz, 0], [0, 0, 0, 1]])
    a.Sphere(0.5, world, a.translate(-1.6, 0.4, 0.2) * a.rotate(10, 20, 30) * M(1.6, 0.5, 1.0, 0.3), lam)
    a.Box(a.Point3D(-0.4, -0.3, -0.2), a.Point3D(0.5, 0.6, 0.7), world, a.translate(-0.4, -0.3, 0.1) * M(0.7, 1.4, 0.9, -0.4) * a.rotate(25, -15, 40), lam)
    a.Cylinder(0.35, 1.1, world, a.translate(0.8, -0.6, 0.0) * a.rotate(-35, 60, 5) * M(1.3, 0.6, 0.8), lam)
    a.Cone(0.45, 0.9, world, a.translate(1.9, -0.4, 0.3) * M(0.5, 1.5, 1.2, 0.2) * a.rotate(15, -70, 0), lam)
    a.Subtract(a.Box(a.Point3D(-0.4, -0.4, -0.4), a.Point3D(0.4, 0.4, 0.4), transform=M(1.2, 0.8, 1.0, 0.1)),
               a.Sphere(0.5, transform=a.translate(0.1, 0, 0) * M(0.9, 1.1, 0.7)),
               world, a.translate(0.5, 1.3, 0.0) * a.rotate(40, 20, -10) * M(1.4, 0.7, 1.1, 0.25), lam)
    verts, tris, normals = icosphere(2, radius=0.4, bumps=0.15)
    a.Mesh(verts, tris, normals, smoothing=True, closed=True, parent=world,
           transform=a.translate(-1.2, 1.4, 0.1) * M(1.5, 0.6, 1.0, -0.3) * a.rotate(20, 35, 10), material=lam)
    a.Box(a.Point3D(-3, -0.05, -3), a.Point3D(3, 0, 3), world, a.translate(0, -1.4, 0), lam)
    a.Sphere(0.3, world, a.translate(0.2, 2.6, -1.0) * M(2.0, 0.4, 2.0), a.UniformSurfaceEmitter(a.ConstantSF(1.0), 3.0))
    return world


def parabola_scene(api):
    """Parabola primitives: rotated, non-rigidly scaled, as CSG operands, next to a lit Lambert floor"""
    a = api
    world = a.World()
    lam = a.Lambert(a.ConstantSF(0.7))
    a.Parabola(0.5, 1.0, world, a.translate(-1.2, -0.4, 0.1) * a.rotate(20, -60, 10), lam)
    a.Parabola(0.3, 0.6, world, a.translate(0.0, -0.5, 0.0) * a.rotate(0, -90, 0)
               * a.AffineMatrix3D([[1.25, 0.25, 0, 0], [0, 0.75, 0, 0], [0, 0, 1.5, 0], [0, 0, 0, 1]]), lam)
    a.Subtract(a.Parabola(0.5, 0.9), a.Parabola(0.4, 0.7, transform=a.translate(0, 0, -0.05)),
               world, a.translate(1.2, -0.5, 0.2) * a.rotate(-30, -75, 0), lam)                      # a parabolic dish
    a.Intersect(a.Parabola(0.6, 1.2), a.Box(a.Point3D(-0.3, -1, 0.1), a.Point3D(0.3, 1, 1.0)),
                world, a.translate(0.0, 0.9, 0.0) * a.rotate(40, -100, 20), lam)
    a.Parabola(0.5, 1.0, world, a.translate(-1.2, 1.0, 0.0), lam)                                    # unrotated: exact axis rays
    a.Box(a.Point3D(-3, -0.05, -3), a.Point3D(3, 0, 3), world, a.translate(0, -1.2, 0), lam)
    a.Sphere(0.3, world, a.translate(0.3, 2.4, -1.2), a.UniformSurfaceEmitter(a.ConstantSF(1.0), 4.0))
    return world


def parabola_rays(n=3000):
    o, d = zoo_rays(n, seed=33)
    # exact rays for the unrotated parabola at (-1.2, 1.0, 0): along the axis through the tip (t0 == t1 branch), in the
    # base plane, tangent to the rim, from inside
    extra_o = np.array([[-1.2, 1.0, -3], [-1.2, 1.0, 3], [-1.2, 1.0, 0.5], [-3, 1.0, 0.0], [-3, 1.5, 0.0], [-1.2, 1.0, 1.0],
                        [-0.95, 1.0, -2], [-1.2, 1.0, 0.0], [-3, 1.0, 1.0], [-1.2, 1.25, 0.75]], dtype=float)
    extra_d = np.array([[0, 0, 1], [0, 0, -1], [0, 0, 1], [1, 0, 0], [1, 0, 0], [0, 0, 1],
                        [0, 0, 1], [1, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=float)
    return np.ascontiguousarray(np.r_[o, extra_o]), np.ascontiguousarray(np.r_[d, extra_d])


def edge_rays():
    """(origins, directions, max_distance) of hand-placed rays: see edge_scene"""
    inf = np.inf
    rows = []

    def add(o, d, md=inf):
        rows.append((o, d, md))
    for s in (1.0, 1e-3, 1e3, 7.0):                          # directions need not be unit length (core/ray.pxd)
        add((-3, 0, 0), (s, 0, 0)); add((0, 0, -3), (0, 0, s)); add((3, 0, -3), (0, 0, s)); add((0, 3, -3), (0, 0, s))
        add((0, -3, -3), (0, 0, s)); add((-3, 0.25, 0.125), (s, 0, 0)); add((4, 0.1, 0.2), (s * 0.6, 0, s * 0.8))
    # origins exactly ON a box face, pointing out / in / along it; edges and corners
    for d in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, 0, 1), (0.6, 0.8, 0), (-0.6, 0.8, 0)):
        add((1, 0, 0), d); add((1, 1, 0), d); add((1, 1, 1), d); add((-1, 0.5, -1), d)
    add((-3, 1, 1), (1, 0, 0)); add((-3, 1, 0), (1, 0, 0)); add((-3, -1, -1), (1, 0, 0)); add((-2, -2, -2), (1, 1, 1))
    add((2, 2, 2), (-1, -1, -1)); add((-3, 1, 0.5), (1, 0, 0)); add((0, 0, 0), (1, 0, 0)); add((0, 0, 0), (0, -1, 0))
    # sphere: tangent lines (zero discriminant), centre, surface origins
    add((3, 0.5, -3), (0, 0, 1)); add((3.5, 0, -3), (0, 0, 1)); add((3, -0.5, 3), (0, 0, -1)); add((3, 0, 0), (0, 1, 0))
    add((3.5, 0, 0), (1, 0, 0)); add((3.5, 0, 0), (-1, 0, 0)); add((3, 0, 0.5), (0, 0, 1)); add((3, 0.5, 0), (1, 0, 0))
    # cylinder: parallel to the axis at r, inside, outside; in the cap planes; tangent to the barrel
    for x in (0.5, 0.25, 0.0, 0.75, -0.5):
        add((x, 3, -2), (0, 0, 1)); add((x, 3, 3), (0, 0, -1)); add((x, 3, 0.5), (0, 0, 1))
    add((-2, 3, 0), (1, 0, 0)); add((-2, 3, 1), (1, 0, 0)); add((-2, 3.5, 0.5), (1, 0, 0)); add((-2, 2.5, 0.25), (1, 0, 0))
    add((0, 3, 0), (0, 0, 1)); add((0, 3, 1), (0, 0, -1)); add((0.5, 3, 0), (1, 0, 0)); add((0, 3, 0.5), (1, 0, 0))
    # cone: through the tip, along the axis, in the base plane, along the slant
    add((0, -3, -2), (0, 0, 1)); add((0, -3, 3), (0, 0, -1)); add((-2, -3, 1), (1, 0, 0)); add((-2, -3, 0), (1, 0, 0))
    add((0.5, -3, 0), (-0.5, 0, 1)); add((-0.5, -3, 0), (0.5, 0, 1)); add((0.25, -3, -2), (0, 0, 1)); add((0, -3, 0.5), (1, 0, 0))
    add((0, -3, 1), (0, 0, 1)); add((0, -3, 1), (0, 0, -1)); add((0, -3, 1), (1, 0, 0))
    # shared face of two boxes: ties between the exit of one and the entry of the other
    add((4, 0, 0), (1, 0, 0)); add((8, 0, 0), (-1, 0, 0)); add((6, 0, 0), (1, 0, 0)); add((6, 0, 0), (-1, 0, 0))
    add((6, 0, -2), (0, 0, 1)); add((6, 0.5, -2), (0, 0, 1)); add((5.5, 0, 0), (1, 0, 0)); add((6, 0, 0), (0, 1, 0))
    # zero-thickness box: through it, in its plane, from a point on it
    add((-3.5, 0, -1), (0, 0, 1)); add((-3.5, 0, 1), (0, 0, -1)); add((-5, 0, 0), (1, 0, 0)); add((-3.5, 0, 0), (0, 0, 1))
    add((-3.5, 0, 0), (1, 0, 0)); add((-3.5, -2, 0), (0, 1, 0))
    # CSG: tangent operands, coplanar faces of a union
    add((0, 0, 2), (0, 0, 1)); add((0.5, 0.5, 2), (0, 0, 1)); add((-2, 0, 4), (1, 0, 0)); add((0, 0, 4), (1, 0, 0))
    add((0.45, 0.45, 4), (-1, 0, 0)); add((1, 0, 4), (1, 0, 0)); add((3.5, 0, 2), (0, 0, 1)); add((3.5, 0, 4), (1, 0, 0))
    add((3.5, 0, 4), (-1, 0, 0)); add((6, 0, 4), (-1, 0, 0)); add((3.5, -2, 4), (0, 1, 0))
    # max_distance: zero, exactly the hit distance, one ulp below and above it, huge
    for md in (0.0, 2.0, np.nextafter(2.0, 0.0), np.nextafter(2.0, 3.0), 1e300, 5e-324):
        add((-3, 0, 0), (1, 0, 0), md); add((3, 0, -2.5), (0, 0, 1), md); add((0, 0, 0), (1, 0, 0), md)
    o = np.array([r[0] for r in rows], dtype=np.float64)
    d = np.array([r[1] for r in rows], dtype=np.float64)
    md = np.array([r[2] for r in rows], dtype=np.float64)
    return np.ascontiguousarray(o), np.ascontiguousarray(d), np.ascontiguousarray(md)


def edge_points():
    """points exactly on surfaces, edges, corners, tips and shared faces (Primitive.contains boundary rules)"""
    pts = [(1, 0, 0), (1, 1, 1), (-1, -1, -1), (0, 0, 0), (1.0000000001, 0, 0), (3.5, 0, 0), (3, 0.5, 0), (3, 0, 0),
           (0.5, 3, 0.5), (0, 3, 0), (0, 3, 1), (0, 3, 1.0000001), (0.5, 3, 0), (0, -3, 1), (0, -3, 0), (0.5, -3, 0),
           (0.25, -3, 0.5), (6, 0, 0), (5, 0, 0), (7, 0.5, 0.5), (-3.5, 0, 0), (-3.5, 0, 1e-12), (0, 0, 4), (0.5, 0, 4),
           (0.45, 0.45, 4.45), (3.5, 0, 4), (4.5, 0, 4), (2.5, 0, 4), (100, 100, 100), (0, 0, 4.5)]
    return np.ascontiguousarray(np.array(pts, dtype=np.float64))


def zoo_rays(n, seed=3):
    """Rays aimed from a shell of random origins at random points inside the zoo's extent, plus rays
    started INSIDE primitives (so exiting hits and t0 < 0 branches are exercised)."""
    rng = np.random.default_rng(seed)
    n_out = (2 * n) // 3
    theta = rng.uniform(0, 2 * np.pi, n_out)
    z = rng.uniform(-1, 1, n_out)
    r = np.sqrt(1 - z * z)
    o_out = 6.0 * np.c_[r * np.cos(theta), r * np.sin(theta), z]
    tgt = np.c_[rng.uniform(-2.8, 2.8, n_out), rng.uniform(-2.6, 2.2, n_out), rng.uniform(-0.6, 0.8, n_out)]
    d_out = tgt - o_out
    d_out /= np.linalg.norm(d_out, axis=1)[:, None]
    n_in = n - n_out
    o_in = np.c_[rng.uniform(-2.6, 2.6, n_in), rng.uniform(-2.4, 2.0, n_in), rng.uniform(-0.4, 0.6, n_in)]
    d_in = rng.normal(size=(n_in, 3))
    d_in /= np.linalg.norm(d_in, axis=1)[:, None]
    # some axis-aligned directions (zero components exercise the parallel-ray branches)
    k = min(60, n_in)
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=float)
    d_in[:k] = axes[np.arange(k) % 6]
    return np.ascontiguousarray(np.r_[o_out, o_in]), np.ascontiguousarray(np.r_[d_out, d_in])


def icosphere(subdivisions=3, radius=0.5, bumps=0.08, seed=5):
    """Deterministic closed triangle mesh: subdivided icosahedron with a smooth radial perturbation.
    Returns (vertices f32 [n,3], triangles i32 [m,6], normals f32 [n,3]) with per-vertex normals."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    v = np.array(v)
    rng = np.random.default_rng(seed)
    k = rng.normal(size=(4, 3))
    scale = 1.0 + bumps * sum(np.sin(3.0 * v @ k[i] + i) for i in range(4)) / 4.0
    verts = (radius * v * scale[:, None]).astype(np.float32)
    tris = np.array(f, dtype=np.int32)
    # area-weighted vertex normals
    p = verts.astype(np.float64)
    fn = np.cross(p[tris[:, 1]] - p[tris[:, 0]], p[tris[:, 2]] - p[tris[:, 0]])
    vn = np.zeros_like(p)
    for i in range(3):
        np.add.at(vn, tris[:, i], fn)
    vn /= np.linalg.norm(vn, axis=1)[:, None]
    return verts, np.ascontiguousarray(np.c_[tris, tris]), vn.astype(np.float32)


def refine_mesh(vertices, triangles, target):
    """Deterministic refinement of a closed triangle mesh to EXACTLY ``target`` triangles (the
    144,046-triangle Stanford bunny -> 1,000,000): 1 -> 4 midpoint subdivision (shared edge midpoints, so the surface
    stays closed) while 4 x count <= target, then longest-edge bisection in triangle index order -- an edge split
    replaces BOTH triangles on that edge by two each (+2, no T-junctions) -- until the count is reached; if one
    triangle is missing at the end (odd remainder) the target is not reachable this way and ValueError is raised.
    Positions are float32, like MeshData's.  Returns (vertices f32 [n,3], triangles i32 [m,3])."""
    v = [tuple(float(c) for c in p) for p in np.asarray(vertices, dtype=np.float32)]
    t = [tuple(int(i) for i in tri[:3]) for tri in np.asarray(triangles)]

    def midpoint(a, b, cache):
        key = (a, b) if a < b else (b, a)
        if key not in cache:
            pa, pb = v[a], v[b]
            m = tuple(float(np.float32(0.5) * (np.float32(pa[k]) + np.float32(pb[k]))) for k in range(3))
            cache[key] = len(v)
            v.append(m)
        return cache[key]
    while 4 * len(t) <= target:
        cache, out = {}, []
        for a, b, c in t:
            ab, bc, ca = midpoint(a, b, cache), midpoint(b, c, cache), midpoint(c, a, cache)
            out += [(a, ab, ca), (ab, b, bc), (ca, bc, c), (ab, bc, ca)]
        t = out
    if (target - len(t)) % 2:
        raise ValueError("target not reachable: longest-edge bisection of a closed mesh adds two triangles per split")
    # edge -> the (up to two) triangles using it, kept current while triangles are replaced
    edge_tris = {}

    def edges_of(tri):
        a, b, c = tri
        return [tuple(sorted(e)) for e in ((a, b), (b, c), (c, a))]

    def link(i):
        for e in edges_of(t[i]):
            edge_tris.setdefault(e, set()).add(i)

    def unlink(i):
        for e in edges_of(t[i]):
            edge_tris[e].discard(i)
    for i in range(len(t)):
        link(i)

    def length2(e):
        pa, pb = np.array(v[e[0]], dtype=np.float64), np.array(v[e[1]], dtype=np.float64)
        return float(((pa - pb) ** 2).sum())

    def split(i, e, m):
        """replace triangle i (which has edge e = (p, q)) by (.., p, m) and (.., m, q), keeping its orientation"""
        tri = t[i]
        k = [j for j in range(3) if tuple(sorted((tri[j], tri[(j + 1) % 3]))) == e][0]
        p, q, r = tri[k], tri[(k + 1) % 3], tri[(k + 2) % 3]
        unlink(i)
        t[i] = (p, m, r)
        t.append((m, q, r))
        link(i)
        link(len(t) - 1)
    i = 0
    cache = {}
    while len(t) < target:
        e = max(edges_of(t[i]), key=lambda ed: (length2(ed), ed))      # ties broken by the vertex ids: deterministic
        m = midpoint(e[0], e[1], cache)
        for j in sorted(edge_tris[e]):
            split(j, e, m)
        i += 1
    return np.array(v, dtype=np.float32), np.array(t, dtype=np.int32)


def mesh_scene(api, smoothing=True):
    """Two instances of the bumpy icosphere (one via Mesh.instance) plus a floor box."""
    a = api
    verts, tris, normals = icosphere()
    world = a.World()
    m = a.AbsorbingSurface()
    mesh = a.Mesh(verts, tris, normals, smoothing=smoothing, closed=True, parent=world,
                  transform=a.translate(-0.45, 0.1, 0.0) * a.rotate(20, 35, 10), material=m)
    mesh.instance(parent=world, transform=a.translate(0.55, -0.05, 0.2) * a.rotate(-40, 10, 70), material=m)
    a.Box(a.Point3D(-2, -0.05, -2), a.Point3D(2, 0, 2), world, a.translate(0, -0.7, 0), m)
    return world


def mesh_rays(n, seed=9):
    rng = np.random.default_rng(seed)
    o = np.c_[rng.uniform(-1.5, 1.5, n), rng.uniform(-0.5, 1.5, n), np.full(n, -3.0)]
    tgt = np.c_[rng.uniform(-1.1, 1.2, n), rng.uniform(-0.7, 0.7, n), rng.uniform(-0.4, 0.6, n)]
    d = tgt - o
    d /= np.linalg.norm(d, axis=1)[:, None]
    # a third of the rays start inside the meshes' bounding region
    k = n // 3
    o[:k] = np.c_[rng.uniform(-0.9, 1.0, k), rng.uniform(-0.5, 0.6, k), rng.uniform(-0.5, 0.7, k)]
    d[:k] = rng.normal(size=(k, 3))
    d[:k] /= np.linalg.norm(d[:k], axis=1)[:, None]
    return np.ascontiguousarray(o), np.ascontiguousarray(d)


def prism_scene(api):
    """Dispersive CSG scene in the spirit of demos/prism.py:16-107: an SF11 glass prism
    (Subtract(Subtract(Box, Box), Box)), a Lambert screen cut by Intersect(Box, Subtract(Cylinder,
    Cylinder)), a small box emitter and an absorbing backdrop."""
    a = api
    world = a.World()
    glass = a.schott("SF11")
    glass.importance = 9
    a.Subtract(
        a.Subtract(a.Box(a.Point3D(-0.3, 0, -0.3), a.Point3D(0.3, 0.5, 0.3)),
                   a.Box(a.Point3D(-0.3, -0.1, -0.1), a.Point3D(1.0, 0.7, 0.9), transform=a.translate(0.3, 0, 0) * a.rotate(30, 0, 0))),
        a.Box(a.Point3D(-1.0, -0.1, -0.1), a.Point3D(0.3, 0.7, 0.9), transform=a.translate(-0.3, 0, 0) * a.rotate(-30, 0, 0)),
        world, a.translate(0, -0.25, 0.3) * a.rotate(15, 0, 0), glass)
    a.Intersect(a.Box(a.Point3D(-1.2, -0.6, -0.05), a.Point3D(1.2, 0.6, 0.05)),
                a.Subtract(a.Cylinder(1.3, 0.2, transform=a.translate(0, 0, -0.1)),
                           a.Cylinder(0.15, 0.4, transform=a.translate(0.4, 0.1, -0.2))),
                world, a.translate(0, 0, 1.6), a.Lambert(a.ConstantSF(0.8)))
    a.Box(a.Point3D(-0.2, -0.2, -0.02), a.Point3D(0.2, 0.2, 0.0), world, a.translate(0.1, 0.9, 0.2) * a.rotate(0, 90, 0),
          a.UniformSurfaceEmitter(a.InterpolatedSF(*CB_LIGHT), 5.0))
    a.Box(a.Point3D(-3, -3, 0), a.Point3D(3, 3, 0.1), world, a.translate(0, 0, 3), a.AbsorbingSurface())
    return world


# ---- rays of rsb_hit_sweep_dev, restated in numpy (csrc/rsb_trav.cuh k_rq_sweep_gen) -----------------------------------
def philox_pair(seed, index):
    """Philox4x32-10 words 0 and 1 of block 0 of stream `index` (csrc/rsb_rng.h Philox4x32: key = seed,
    counter = (0, 0, index lo, index hi)) -> two uniform() values, as the sweep draws them."""
    index = np.asarray(index, dtype=np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    c0 = np.zeros_like(index); c1 = np.zeros_like(index)
    c2 = index & m32; c3 = index >> np.uint64(32)
    k0 = np.uint64(seed & 0xFFFFFFFF); k1 = np.uint64(seed >> 32)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & m32, p1 >> np.uint64(32), p1 & m32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    w0 = (c1 << np.uint64(32)) | c0
    w1 = (c3 << np.uint64(32)) | c2
    scale = 1.0 / 9007199254740992.0
    return (w0 >> np.uint64(11)).astype(np.float64) * scale, (w1 >> np.uint64(11)).astype(np.float64) * scale


def _morton_compact(x):
    x = x & np.uint64(0x5555555555555555)
    for shift, mask in ((1, 0x3333333333333333), (2, 0x0F0F0F0F0F0F0F0F), (4, 0x00FF00FF00FF00FF), (8, 0x0000FFFF0000FFFF),
                        (16, 0x00000000FFFFFFFF)):
        x = (x | (x >> np.uint64(shift))) & np.uint64(mask)
    return x


def sweep_rays(seed, first, n, origin, target, half_window, order_log2=0):
    """origins, directions [n][3] of rays first .. first + n - 1 of a sweep: toward target + (jx, jy, 0) * half_window;
    order_log2 = 0: (jx, jy) uniform over the window per ray; g > 0: jittered inside cell (index mod 4^g) of a
    2^g x 2^g grid walked along the Morton curve."""
    idx = np.arange(first, first + n, dtype=np.uint64)
    u1, u2 = philox_pair(seed, idx)
    if order_log2 > 0:
        cell = idx & np.uint64((1 << (2 * order_log2)) - 1)
        inv = 1.0 / float(1 << order_log2)
        u1 = (_morton_compact(cell).astype(np.float64) + u1) * inv
        u2 = (_morton_compact(cell >> np.uint64(1)).astype(np.float64) + u2) * inv
    px = target[0] + (2.0 * u1 - 1.0) * half_window
    py = target[1] + (2.0 * u2 - 1.0) * half_window
    pz = np.full(n, float(target[2]))
    dx, dy, dz = px - origin[0], py - origin[1], pz - origin[2]
    t = dx * dx + dy * dy + dz * dz
    t = 1.0 / np.sqrt(t)
    d = np.stack([dx * t, dy * t, dz * t], axis=1)
    o = np.tile(np.array(origin, dtype=np.float64), (n, 1))
    return np.ascontiguousarray(o), np.ascontiguousarray(d), idx


SWEEP_ORIGIN, SWEEP_TARGET, SWEEP_HALF = (0.0, 0.0, -4.0), (0.0, 0.0, 0.0), 0.9    # BASELINE config 5 (tools_sweep.py, bench.py)


def sweep_spheres(api, uniform, n=10000):
    """BASELINE config 5 scene: n spheres drawn from the REFERENCE generator after seed(7); `uniform` is a
    callable yielding that stream: raysect.core.math.random.uniform on the reference, the device's / the host build's
    restatement of MT19937-64 elsewhere (rng_uniform(7, 4 n))."""
    return random_spheres(api, n, seed=7, uniform=uniform)


# ---- the reference's own mesh fixture and BASELINE config 4 (Cornell box + Stanford bunny refined to 1,000,000 triangles) ---
import os as _os

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
BUNNY_RSM = _os.path.join(_ROOT, "oracle", "_ref", "resources", "stanford_bunny.rsm")   # demos/resources/stanford_bunny.rsm
BUNNY_OBJ = _os.path.join(_ROOT, "oracle", "_ref", "resources", "stanford_bunny.obj")   # demos/resources/stanford_bunny.obj
MESH_CACHE = _os.path.join(_ROOT, "build", "cache")


def bunny_rsm_scene(api, path=BUNNY_RSM):
    """The 144,046-triangle bunny with the kd-tree the reference shipped inside the file (Mesh.from_file: nothing is
    rebuilt), as demos/bunny.py places it, plus a ground box."""
    a = api
    world = a.World()
    m = a.AbsorbingSurface()
    a.Mesh.from_file(path, parent=world, transform=a.translate(0.0, 0.0, 0.0) * a.rotate(165, 0, 0), material=m)
    a.Box(a.Point3D(-1, -0.05, -1), a.Point3D(1, 0.033, 1), world, a.translate(0, 0, 0), m)
    return world


def bunny_rsm_rays(n=4000, seed=12):
    """rays from a shell around the bunny into its bounding region; a tenth start inside the region"""
    rng = np.random.default_rng(seed)
    centre = np.array([-0.017, 0.11, 0.0])
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    u[:, 1] = np.abs(u[:, 1]) * 0.8 + 0.05
    o = centre + 0.45 * u
    tgt = centre + rng.uniform([-0.09, -0.08, -0.07], [0.09, 0.08, 0.07], (n, 3))
    k = n // 10
    o[:k] = centre + rng.uniform([-0.08, -0.07, -0.06], [0.08, 0.07, 0.06], (k, 3))
    d = tgt - o
    d[:k] = rng.normal(size=(k, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    return np.ascontiguousarray(o), np.ascontiguousarray(d)


def bunny_rsm_points(n=1500, seed=13):
    rng = np.random.default_rng(seed)
    return np.array([-0.017, 0.11, 0.0]) + rng.uniform([-0.1, -0.09, -0.08], [0.1, 0.09, 0.08], (n, 3))


def refined_bunny_rsm(triangles=1000000, obj=BUNNY_OBJ, cache_dir=MESH_CACHE):
    """BASELINE config 4 mesh as SURVEY 8(d) specifies it: demos/resources/stanford_bunny.obj refined deterministically to
    exactly `triangles` triangles (refine_mesh), scaled to stand 1 m tall on the floor of the Cornell box, its SAH
    kd-tree built by this package (byte-identical to the reference builder's, tests/test_cabi_symbols.py), written as a
    Raysect mesh file with the mirror's byte-identical .rsm writer and cached: both the reference (Mesh.from_file) and
    this package load mesh AND tree from the very same bytes.  Returns the file name."""
    path = _os.path.join(cache_dir, "bunny_%d.rsm" % triangles)
    if _os.path.exists(path):
        return path
    import source_b200 as mirror
    base = mirror.import_obj(obj)
    v, t = refine_mesh(base.data.vertices, base.data.triangles, triangles)
    v = v.astype(np.float64)
    lo, hi = v.min(0), v.max(0)
    s = 1.0 / (hi[1] - lo[1])
    v = (v - [0.5 * (lo[0] + hi[0]), lo[1], 0.5 * (lo[2] + hi[2])]) * s + [0.0, -1.0 + 1e-6 + 0.5, 0.0]
    mesh = mirror.Mesh(v.astype(np.float32), t, None, smoothing=False, closed=True)
    _os.makedirs(cache_dir, exist_ok=True)
    tmp = path + ".tmp.%d" % _os.getpid()
    mesh.save(tmp)
    _os.replace(tmp, path)
    return path


def cornell_mesh_scene(api, rsm_path, glass=False):
    """BASELINE config 4: the Cornell box with a Lambert mesh (from a .rsm file) standing in it"""
    def extra(a, w):
        a.Mesh.from_file(rsm_path, parent=w, transform=a.translate(0.1, -0.5, 0.1) * a.rotate(20, 0, 0),
                         material=a.Lambert(a.ConstantSF(0.7)))
    return cornell_box(api, glass=glass, extra=extra)


def cornell_mesh_rays(n=4000, seed=14):
    """camera-like rays toward the mesh plus rays scattered from points inside the box (bounce-like)"""
    rng = np.random.default_rng(seed)
    o = np.tile(np.array([0.0, 0.0, -3.3]), (n, 1))
    tgt = np.array([0.1, -0.5, 0.1]) + rng.uniform([-0.6, -0.5, -0.5], [0.6, 0.55, 0.5], (n, 3))
    k = n // 2
    o[:k] = rng.uniform([-0.95, -0.95, -0.95], [0.95, 0.95, 0.95], (k, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1)[:, None]
    return np.ascontiguousarray(o), np.ascontiguousarray(d)


# ---- transforms whose inverse is not exactly affine: AffineMatrix3D.inverse() of a non-rigid chain can leave m33 = 1 - 1 ulp,
# ---- and Point3D.transform divides by it (raysect/core/math/point.pyx:272-281)
def w_transforms(api, count, seed=0):
    """`count` non-rigid chains translate * rotate * (scale + shear), drawn from a fixed stream and kept only if the
    bottom-right element of their inverse differs from 1.0 (about one random chain in eight)"""
    a = api
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        s = rng.uniform(0.5, 1.8, 3)
        sh = rng.uniform(-0.4, 0.4, 3)
        m = a.AffineMatrix3D([[s[0], sh[0], sh[1], 0], [0, s[1], sh[2], 0], [0, 0, s[2], 0], [0, 0, 0, 1]])
        t = a.translate(*rng.uniform(-1, 1, 3)) * a.rotate(*rng.uniform(-90, 90, 3)) * m
        if t.inverse()[3, 3] != 1.0:
            out.append(t)
    return out


def w_scene(api):
    """every primitive type, a CSG tree (operands included) and a mesh under transforms whose to_local() has m33 != 1"""
    a = api
    world = a.World()
    lam = a.Lambert(a.ConstantSF(0.6))
    t = w_transforms(a, 9)
    place = [(-1.7, 0.3, 0.1), (-0.5, -0.4, 0.2), (0.8, -0.5, 0.0), (1.9, -0.3, 0.2), (0.4, 1.2, 0.0), (-1.2, 1.4, 0.1)]

    def at(k):
        return a.translate(*place[k]) * t[k]
    a.Sphere(0.45, world, at(0), lam)
    a.Box(a.Point3D(-0.35, -0.3, -0.25), a.Point3D(0.4, 0.45, 0.5), world, at(1), lam)
    a.Cylinder(0.3, 0.9, world, at(2), lam)
    a.Cone(0.4, 0.8, world, at(3), lam)
    a.Subtract(a.Box(a.Point3D(-0.4, -0.4, -0.4), a.Point3D(0.4, 0.4, 0.4), transform=t[6]),
               a.Sphere(0.45, transform=a.translate(0.1, 0, 0) * t[7]), world, at(4), lam)
    verts, tris, normals = icosphere(2, radius=0.4, bumps=0.15)
    a.Mesh(verts, tris, normals, smoothing=True, closed=True, parent=world, transform=at(5), material=lam)
    a.Box(a.Point3D(-3, -0.05, -3), a.Point3D(3, 0, 3), world, a.translate(0, -1.6, 0), lam)
    a.Sphere(0.3, world, a.translate(0.2, 2.8, -1.0) * t[8], a.UniformSurfaceEmitter(a.ConstantSF(1.0), 3.0))
    return world
