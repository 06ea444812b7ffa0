from .mlp import Mlp
