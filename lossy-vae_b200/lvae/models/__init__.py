from . import qarv
from . import rd
