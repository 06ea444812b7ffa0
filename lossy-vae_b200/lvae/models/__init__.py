from . import qarv
