from . import qarv
from . import rd
from . import qresvae
