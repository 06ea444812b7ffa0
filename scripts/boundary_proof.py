"""Boundary proof: run the reference's own three scripts -- eval-var-rate.py,
scripts/speedtest-lvae.py and train-var-rate.py -- UNMODIFIED against the B200 package on a GPU box.

    python scripts/boundary_proof.py [--out gpurun_out] [--iterations 20]

The scripts are taken from a reference checkout: /root/reference in the build container, else the byte-identical staged
copy oracle/_ref/reference (oracle/make_ref.py; its MANIFEST.sha256 is checked first).  They are started through
scripts/run_reference_script.py, which puts THIS repo's `lvae` package first on sys.path and points LVAE_REFERENCE_ROOT
at the checkout, so that `lvae.trainer`, `lvae.datasets` and `lvae.utils.general` (callers of the path: out of scope,
SURVEY 2) come from the reference, unmodified, while `lvae.get_model`, `lvae.models.*`, `lvae.evaluation`,
`lvae.utils.coding` and `lvae.paths` are this repo's.  `timm.utils` resolves to lossy-vae_b200/compat (timm is not in the
image).  There is no data set and no checkpoint offline, so this script first writes
  * datasets/kodak            6 seeded image-like PNGs, 768 x 512 (Kodak shape)
  * datasets/coco/train2017   48 seeded image-like PNGs, 320 x 320
  * ckpt.pt                   {'model': state_dict} with the seeded sensitised weights the parity tests use
under a scratch directory and exports LVAE_DATASETS.  Each script's stdout + stderr goes to <out>/r2_boundary_<name>.log;
a summary line per script (exit code, seconds, key results parsed from the log) to <out>/r2_boundary_summary.json.
TEST INFRASTRUCTURE: nothing here is on the product path."""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'oracle'))


def reference_root():
    live, staged = Path('/root/reference'), ROOT / 'oracle' / '_ref' / 'reference'
    if (live / 'lvae' / '__init__.py').is_file():
        return live, 'live checkout'
    man = staged.parent / 'MANIFEST.sha256'
    assert staged.is_dir() and man.is_file(), 'no reference copy: run oracle/make_ref.py where /root/reference exists'
    bad = [l.split('  ', 1)[1] for l in man.read_text().splitlines()
           if hashlib.sha256((staged / l.split('  ', 1)[1]).read_bytes()).hexdigest() != l.split('  ', 1)[0]]
    assert not bad, f'staged reference files differ from the manifest: {bad}'
    return staged, f'staged copy, {len(man.read_text().splitlines())} files, sha256 manifest verified'


def write_inputs(work):
    import numpy as np
    import torch
    from PIL import Image
    import lvae_oracle as O
    from oracle_inputs import make_input
    kodak, coco = work / 'datasets' / 'kodak', work / 'datasets' / 'coco' / 'train2017'
    kodak.mkdir(parents=True, exist_ok=True)
    coco.mkdir(parents=True, exist_ok=True)
    for i in range(6):
        arr = (make_input('synth', 1, 512, 768, 700 + i)[0].permute(1, 2, 0).numpy() * 255).round().astype(np.uint8)
        Image.fromarray(arr).save(kodak / f'kodim{i + 1:02d}.png')
    for i in range(48):
        arr = (make_input('synth', 1, 320, 320, 800 + i)[0].permute(1, 2, 0).numpy() * 255).round().astype(np.uint8)
        Image.fromarray(arr).save(coco / f'{i:012d}.png')
    # a checkpoint as the reference's trainer writes it: {'model': model.state_dict()} incl. the entropy-model buffers
    sys.path.insert(0, str(ROOT / 'lossy-vae_b200'))
    import lvae
    torch.manual_seed(0)
    m = lvae.get_model('qarv_base')
    m.load_state_dict(O.sensitised_state_dict(O.qarv_param_shapes(), seed=0), strict=False)
    torch.save({'model': m.state_dict()}, work / 'ckpt.pt')
    return work / 'datasets', work / 'ckpt.pt'


def run(name, cmd, cwd, env, out, timeout):
    log = out / f'r2_boundary_{name}.log'
    t0 = time.time()
    with open(log, 'w') as f:
        f.write('$ ' + ' '.join(cmd) + f'\n# cwd {cwd}\n')
        f.flush()
        try:
            rc = subprocess.run(cmd, cwd=cwd, env=env, stdout=f, stderr=subprocess.STDOUT, timeout=timeout).returncode
        except subprocess.TimeoutExpired:
            rc = 'timeout'
    txt = log.read_text()
    return dict(script=name, rc=rc, seconds=round(time.time() - t0, 1), log=str(log.relative_to(ROOT) if log.is_relative_to(ROOT) else log), tail=txt[-1500:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=str(ROOT / 'gpurun_out'))
    ap.add_argument('--work', default='/tmp/lvae_boundary')
    ap.add_argument('--iterations', type=int, default=20)
    args = ap.parse_args()
    out, work = Path(args.out).resolve(), Path(args.work)
    out.mkdir(parents=True, exist_ok=True)
    work.mkdir(parents=True, exist_ok=True)
    ref, how = reference_root()
    datasets, ckpt = write_inputs(work)
    env = dict(os.environ, LVAE_DATASETS=str(datasets), LVAE_REFERENCE_ROOT=str(ref), WANDB_MODE='disabled', WANDB_SILENT='true')
    runner = [sys.executable, str(ROOT / 'scripts' / 'run_reference_script.py')]
    # the scripts write runs/... relative to the cwd and train-var-rate.py reads results/kodak/... relative to it: run
    # them in a scratch copy of the checkout's top level (symlinks), never inside the checkout itself
    cwd = work / 'cwd'
    cwd.mkdir(exist_ok=True)
    if not (cwd / 'results').exists():
        os.symlink(ref / 'results', cwd / 'results')
    res = [dict(reference=str(ref), how=how)]
    res.append(run('eval_var_rate', runner + [str(ref / 'eval-var-rate.py'), '-m', 'qarv_base', '-a', f"pretrained='{ckpt}'",
                                            '-n', 'kodak', '-s', '3', '-l', '64', '1024'], cwd, env, out, 900))
    res.append(run('speedtest_lvae', runner + [str(ref / 'scripts' / 'speedtest-lvae.py'), '-a', f"pretrained='{ckpt}'"], cwd, env, out, 900))
    res.append(run('train_var_rate', runner + [str(ref / 'train-var-rate.py'), '--model', 'qarv_base', '--batch_size', '8',
                                             '--iterations', str(args.iterations), '--trainset', 'coco-train2017', '--valset', 'kodak',
                                             '--val_steps', '2', '--workers', '2', '--wbmode', 'disabled', '--ema_warmup', '10'],
                   cwd, env, out, 1500))
    js = cwd / 'runs' / 'results' / 'kodak-qarv_base.json'
    if js.is_file():
        res[1]['results_json'] = json.loads(js.read_text())
    (out / 'r2_boundary_summary.json').write_text(json.dumps(res, indent=1))
    for r in res[1:]:
        print(f"{r['script']}: rc={r['rc']} in {r['seconds']} s -> {r['log']}")
        print('   ' + r['tail'].strip().splitlines()[-1][:200] if r['tail'].strip() else '')
    return 0 if all(r['rc'] == 0 for r in res[1:]) else 1


if __name__ == '__main__':
    sys.exit(main())
