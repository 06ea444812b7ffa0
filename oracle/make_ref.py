"""Stage the UNMODIFIED reference under oracle/_ref/reference (git-ignored, NOT gpurun-ignored: it travels to the GPU box).

    python oracle/make_ref.py            # no-op when /root/reference is absent (the GPU box: it uses the staged copy)

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python that needs two absent third-party packages
(timm, compressai), so it cannot be pip-installed offline; what CAN travel is a byte-identical copy of its
Python sources, made by this committed recipe and kept out of the repository's history.  Consumers:
  * bench.py --impl reference and the `cpu_baseline` leg: the reference's own `lvae` package on the host CPU
    (`cpu_baseline.kind = "reference"`), imported through oracle/ref_loader.py with oracle/shims for timm / compressai;
  * scripts/boundary_proof.py: runs the reference's eval-var-rate.py, scripts/speedtest-lvae.py and train-var-rate.py
    unmodified against the B200 package (they are loaded from this copy; `lvae` resolves to lossy-vae_b200/lvae).
Nothing under lossy-vae_b200/ (the product) reads this directory, except in overlay mode when the USER points
LVAE_REFERENCE_ROOT at a reference checkout.  A MANIFEST with the sha256 of every staged file is
written next to the copy so that "unmodified" can be checked on the box."""
import hashlib
import shutil
import sys
from pathlib import Path

SRC = Path('/root/reference')
DST = Path(__file__).resolve().parent / '_ref' / 'reference'
WHAT = ['lvae', 'eval-var-rate.py', 'eval-fix-rate.py', 'train-var-rate.py', 'train-fix-rate.py', 'scripts/speedtest-lvae.py',
        'results/kodak']      # results/kodak: the VTM anchor json train-var-rate.py's evaluate() reads relative to the cwd


def stage(verbose=True):
    if not (SRC / 'lvae' / '__init__.py').is_file():
        if verbose:
            print(f'{SRC} not present: keeping {DST} as it is ({"staged" if DST.is_dir() else "absent"})')
        return DST.is_dir()
    if DST.is_dir():
        shutil.rmtree(DST)
    manifest = []
    for rel in WHAT:
        s, d = SRC / rel, DST / rel
        d.parent.mkdir(parents=True, exist_ok=True)
        if s.is_dir():
            shutil.copytree(s, d, ignore=shutil.ignore_patterns('__pycache__', '*.pyc', '*.ipynb'))
        else:
            shutil.copy2(s, d)
    for f in sorted(p for p in DST.rglob('*') if p.is_file()):
        manifest.append(f'{hashlib.sha256(f.read_bytes()).hexdigest()}  {f.relative_to(DST)}')
    (DST.parent / 'MANIFEST.sha256').write_text('\n'.join(manifest) + '\n')
    if verbose:
        print(f'staged {len(manifest)} reference files under {DST}')
    return True


if __name__ == '__main__':
    sys.exit(0 if stage() else 1)
