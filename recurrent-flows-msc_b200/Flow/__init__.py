from .glow import GlowStep, ListGlow  # noqa: F401
from .glow_modules import (ActNorm, AffineCoupling, BatchNormFlow, Conv2dNorm, Conv2dZeros, InvConv,  # noqa: F401
                           Split2d, Squeeze2d)
