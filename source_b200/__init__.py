"""source_b200 -- B200-native ray/scene intersection and spectral trace path behind Raysect's
World / Primitive / Material / Observer plugin API.

Two ways in:
  * with Raysect installed: ``source_b200.plugin.CudaAccelerator`` (``world.accelerator = ...``) and
    ``source_b200.plugin.CudaRenderEngine`` (``camera.render_engine = ...``) on the real Raysect objects;
  * stand-alone: the mirror object model exported below (same names and argument meaning).
Both flatten the scenegraph (``source_b200.flatten``) and call ``libraysect_b200.so`` through ctypes.
See DESIGN.md and INTEGRATION.md.
"""
__version__ = "0.1.0"

from .math3d import (AffineMatrix3D, BoundingBox3D, BoundingSphere3D, Normal3D, Point3D, Vector3D, rotate, rotate_x,
                     rotate_y, rotate_z, translate)
from .spectral import ConstantSF, InterpolatedSF, NumericallyIntegratedSF, Sellmeier, SpectralFunction
from .material import (AbsorbingSurface, Checkerboard, Conductor, Dielectric, Lambert, Material, RoughConductor, UniformSurfaceEmitter,
                       UniformVolumeEmitter, UnitySurfaceEmitter, UnityVolumeEmitter, schott)
from .scenegraph import (Box, Cone, Cylinder, Intersect, Intersection, Mesh, MeshData, Node, Parabola, Primitive, Ray,
                         Sphere, Subtract, Union, World)
from .observer import (CCDArray, FullFrameSampler2D, Observer, OrthographicCamera, PinholeCamera, SpectralAdaptiveSampler2D,
                       SpectralPowerPipeline2D, SpectralRadiancePipeline2D, SpectralSlice, StatsArray3D, VectorCamera)
from .meshio import import_obj, import_ply, import_stl, import_vtk
from .engine import Accelerator, Device, default_device
from ._cabi import RNG_MT19937_64, RNG_PHILOX, RsbError
