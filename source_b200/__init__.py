"""source_b200 -- B200-native ray/scene intersection and spectral trace path behind Raysect's
World / Primitive / Material / Observer plugin API.  See DESIGN.md and INTEGRATION.md."""
__version__ = "0.1.0"
