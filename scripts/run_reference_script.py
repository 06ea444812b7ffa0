"""Run one of the reference's own scripts, unmodified, against this package:

    python scripts/run_reference_script.py /path/to/lossy-vae/eval-var-rate.py -m qarv_base -a "pretrained='ckpt.pt'" -n kodak
    python scripts/run_reference_script.py /path/to/lossy-vae/scripts/speedtest-lvae.py -a "pretrained='ckpt.pt'"

    python scripts/run_reference_script.py /path/to/lossy-vae/train-var-rate.py --model qarv_base --batch_size 16

`python /path/to/lossy-vae/<script>` would put the reference's own `lvae` first on sys.path; this runner
puts ours there instead and hands over with runpy.  It also sets LVAE_REFERENCE_ROOT to the checkout the script lives in, so
that the sub-modules this package leaves to the reference (lvae.trainer, lvae.datasets, lvae.utils.general: callers of the
path) are loaded from there, unmodified.  Datasets: set LVAE_DATASETS (lvae/paths.py)."""
import os
import runpy
import sys
from pathlib import Path

if __name__ == '__main__':
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = sys.argv[1]
    for root in Path(script).resolve().parents:          # the checkout: first ancestor that holds an `lvae` package
        if (root / 'lvae' / '__init__.py').is_file():
            os.environ.setdefault('LVAE_REFERENCE_ROOT', str(root))
            break
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / 'lossy-vae_b200'))
    import lvae  # noqa: F401,E402 -- first: without timm installed it registers the `timm.utils` restatement (lossy-vae_b200/compat), which train-var-rate.py:5 imports before it imports lvae
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name='__main__')
