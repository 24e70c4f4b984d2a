from .model import Model, GPModel
from .gpr import GPR
from .sgpr import SGPR
from .svgp import SVGP
