"""B200-native (sm_100a) hot path of GAN_SR_wind_field: Conv3d fwd/dgrad/wgrad of the RRDB generator and the
3-D discriminator, fused epilogues, nearest upsample and the wind-field loss stencils, behind the reference's
own module constructors.  See DESIGN.md / INTEGRATION.md."""
from . import _lib, ops  # noqa: F401
from .ops import get_precision, precision, set_precision  # noqa: F401

__all__ = ["ops", "set_precision", "get_precision", "precision"]
