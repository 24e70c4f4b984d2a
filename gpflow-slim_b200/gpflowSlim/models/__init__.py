from .model import Model, GPModel
from .gpr import GPR
from .sgpr import SGPR, GPRFITC, SGPRUpperMixin
from .svgp import SVGP
from .gpmc import GPMC
from .sgpmc import SGPMC
