from . import zoo
