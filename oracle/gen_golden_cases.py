"""Case table and probe of the training-gradient fixtures, shared by oracle/gen_golden.py (which needs the reference) and
the tests (which do not).  TEST INFRASTRUCTURE ONLY."""
import torch

import lvae_oracle as O

# family -> (nB, H, W, lambdas, image seed, noise seed)
GRAD_CASES = {'qarv': (2, 64, 64, [64.0, 1024.0], 5, 17), 'qres': (1, 64, 64, None, 9, 23)}


def grad_probe(key, shape):
    """fixed pseudo-random direction per parameter tensor: <grad, probe> pins the gradient's direction, not only its norm"""
    return torch.randn(shape, generator=O._key_generator(7, key))
