import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / 'lossy-vae_b200', ROOT / 'oracle', ROOT):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu under gpurun)')


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a box without a CUDA device."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='needs a CUDA device (B200)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def _build():
    import importlib.util
    spec = importlib.util.spec_from_file_location('lvae_b200_build', ROOT / 'lossy-vae_b200' / 'build.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build(verbose=False)


@pytest.fixture(scope='session')
def native_lib():
    """The C-ABI shared library (built on demand; loading it needs no GPU)."""
    from lvae import _native
    if not _native.lib_path().is_file():
        _build()
    return _native.lib()


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    def load(name):
        return np.load(GOLDEN / f'{name}.npz')
    return load


@pytest.fixture(scope='session')
def sensitised_sd():
    import lvae_oracle as O
    return O.sensitised_state_dict(O.qarv_param_shapes(), seed=0)


@pytest.fixture(scope='session')
def gpu_model(native_lib, sensitised_sd):
    """qarv_base on cuda:0 with the seeded sensitised weights, in compress mode."""
    import torch
    import lvae
    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    torch.manual_seed(0)
    model = lvae.get_model('qarv_base')
    missing, unexpected = model.load_state_dict(sensitised_sd, strict=False)
    assert not unexpected and all('discrete_gaussian' in k for k in missing), (missing, unexpected)
    model = model.to('cuda:0').eval()
    model.compress_mode()
    return model


def parity_log(**record):
    """Append one parity measurement (case, precision, symbol / index flips, d bpp, d PSNR ...) to
    gpurun_out/r2_parity_records.jsonl -- the raw material of profiles/r2_parity.md (scripts/parity_report.py).  The
    directory is what gpurun brings back from the GPU box; without it (plain CPU run) this is a no-op."""
    import json
    out = ROOT / 'gpurun_out'
    if not out.is_dir():
        return
    with open(out / 'r2_parity_records.jsonl', 'a') as f:
        f.write(json.dumps(record) + '\n')
