"""Evaluation helpers the reference's scripts import next to the model surface (reference: lvae/evaluation.py:15-115;
`eval-var-rate.py:10,46`, `train-var-rate.py:134-148`): real-bit-stream evaluation through `compress_file` /
`decompress_file`, and rate-estimate evaluation through `forward()`.  Host-side orchestration only -- every number comes
from the model's B200 path.  `batch_size > 1` (an extension) groups same-shape images into one call: one
forward() for the estimate-based evaluation, one compress_batch() / decompress_batch() pair for the bit-stream
evaluation -- every image still gets its own standard single-image bit stream (byte-identical to compress_file's: the
kernels are batch-invariant), whose size incl. the 4-byte file header is what is counted."""
import math
from collections import defaultdict
from pathlib import Path
from tempfile import gettempdir

import torch

from .paths import known_datasets
from .utils.coding import crop_divisible_by


class _Mean:
    def __init__(self):
        self.sum, self.count = 0.0, 0

    def update(self, v, n=1):
        self.sum += float(v) * n
        self.count += n

    @property
    def avg(self):
        return self.sum / max(1, self.count)


def _image_paths(dataset):
    root = known_datasets.get(dataset, Path(dataset))
    return sorted(Path(root).rglob('*.*'))


@torch.no_grad()
def imcoding_evaluate(model, dataset, progress=False, batch_size=1):
    """Average bpp / mse / psnr over a dataset with real entropy coding (reference lvae/evaluation.py:15-44): the size of
    the file `compress_file` writes and the reconstruction `decompress_file` returns.  batch_size > 1: same-size images
    are coded `batch_size` at a time through compress_batch / decompress_batch (models that have them); the bits counted
    per image are those of the file compress_file would have written (4-byte size header + the image's own stream)."""
    import torchvision.transforms.functional as tvf
    from PIL import Image
    assert hasattr(model, 'compress_file') and hasattr(model, 'decompress_file')
    tmp = Path(gettempdir())
    stats = defaultdict(_Mean)

    def account(impath, real, fake, num_bits):
        mse = (real - fake).square().mean().item()
        cur = dict(bpp=num_bits / float(real.shape[1] * real.shape[2]), mse=mse, psnr=-10 * math.log10(mse))
        for k, v in cur.items():
            stats[k].update(v)
        if progress:
            print(f'image {impath.stem}: ' + ', '.join(f'{k}={v:.3f}' for k, v in cur.items()))

    batched = batch_size > 1 and hasattr(model, 'compress_batch') and hasattr(model, 'decompress_batch')
    if not batched:
        for impath in _image_paths(dataset):
            bits = tmp / f'{impath.stem}.bits'
            model.compress_file(impath, bits)
            num_bits = bits.stat().st_size * 8
            fake = model.decompress_file(bits).squeeze(0).cpu()
            bits.unlink()
            account(impath, tvf.to_tensor(Image.open(impath)), fake, num_bits)
        return {k: m.avg for k, m in stats.items()}

    from .utils.coding import pad_divisible_by
    device = next(model.parameters()).device
    pending = []          # (path, real [3,h,w], padded [1,3,H,W]) of one padded shape

    def flush():
        if not pending:
            return
        blobs = model.compress_batch(torch.cat([p[2] for p in pending], dim=0).to(device))
        rec = model.decompress_batch(blobs).cpu()
        for (impath, real, _), blob, fake in zip(pending, blobs, rec):
            h, w = real.shape[1:]
            account(impath, real, fake[:, :h, :w], (4 + len(blob)) * 8)     # '2H' header of compress_file + the stream
        pending.clear()

    for impath in _image_paths(dataset):
        img = Image.open(impath)
        padded = tvf.to_tensor(pad_divisible_by(img, div=model.max_stride)).unsqueeze_(0)
        if pending and (pending[0][2].shape != padded.shape or len(pending) >= batch_size):
            flush()
        pending.append((impath, tvf.to_tensor(img), padded))
        if len(pending) >= batch_size:
            flush()
    flush()
    return {k: m.avg for k, m in stats.items()}


@torch.no_grad()
def image_self_evaluate(model, dataset, progress=False, batch_size=1):
    """Average of the model's own `forward()` statistics (estimated rate, no entropy coding) over a dataset, images
    centre-cropped to a multiple of the model stride.  batch_size > 1 runs same-shape images together."""
    import torchvision.transforms.functional as tvf
    from PIL import Image
    device = next(model.parameters()).device
    stats = defaultdict(_Mean)
    pending = []          # images of one shape waiting for a batch

    def flush():
        if not pending:
            return
        out = model(torch.cat(pending, dim=0))
        assert isinstance(out, dict), f'{type(out)=}. expected a dict.'
        for k, v in out.items():
            if isinstance(v, torch.Tensor) and v.dim() > 0:
                continue
            stats[k].update(float(v), n=len(pending))
        pending.clear()

    for impath in _image_paths(dataset):
        img = Image.open(impath)
        if hasattr(model, 'max_stride'):
            img = crop_divisible_by(img, div=model.max_stride)
        im = tvf.to_tensor(img).unsqueeze_(0).to(device=device)
        if pending and (pending[0].shape != im.shape or len(pending) >= batch_size):
            flush()
        pending.append(im)
        if len(pending) >= batch_size:
            flush()
    flush()
    return {k: m.avg for k, m in stats.items()}
