"""Setup builders mirroring the reference's scripts/*/setup/*Setup.jl files."""
from .ks import KSSetup  # noqa: F401
from .kseg import KellerSegelSetup, KellerSegel2DSetup  # noqa: F401
from .ns import FluidSetup  # noqa: F401
