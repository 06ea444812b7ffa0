"""The C-ABI library loads without a GPU and exports exactly what include/lvae_b200.h declares."""
import re
import subprocess

from conftest import ROOT


def _header_symbols():
    text = (ROOT / 'include' / 'lvae_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return set(re.findall(r'\b(lvae_[a-z0-9_]+)\s*\(', text))


def test_header_matches_exports(native_lib):
    from lvae import _native
    declared = _header_symbols()
    out = subprocess.run(['nm', '-D', '--defined-only', str(_native.lib_path())], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ' T ' in ln and ln.split()[-1].startswith('lvae_')}
    assert declared == exported, (sorted(declared - exported), sorted(exported - declared))
    assert declared == set(_native.EXPORTS), sorted(declared ^ set(_native.EXPORTS))


def test_version_and_error_string(native_lib):
    assert native_lib.lvae_version() >= 100
    assert isinstance(native_lib.lvae_last_error(), bytes)


def test_bad_arguments_are_rejected_without_a_gpu(native_lib):
    # argument validation happens before any CUDA call
    assert native_lib.lvae_gemm(None, None) == -1
    assert b'bad argument' in native_lib.lvae_last_error()
    assert native_lib.lvae_latent_eval(None, None, None, 0, None, None, 0, None, None, None, 1, 1, 1, 0, None) == -1
    assert native_lib.lvae_rans_decode(None, 0, None, 0, None, 0, None, None, 0, None) == -1


def test_only_sm100a_code_is_embedded(native_lib):
    from lvae import _native
    out = subprocess.run(['cuobjdump', '-lelf', str(_native.lib_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs
