"""Bit-stream container and image padding helpers on the compress path.

Format contract (reference lvae/utils/coding.py:26-110): a packed blob is
    uint8  n                      number of byte strings (one per latent layer)
    uint32 length[n]              native byte order (`struct` default), one per string
    bytes  payload                the strings back to back
and images are edge-padded on the right / bottom to a multiple of `div`.
"""
import math
import struct

import numpy as np


def pack_byte_strings(list_of_strings):
    n = len(list_of_strings)
    if n > 255:
        raise ValueError(f'at most 255 strings fit the uint8 count field, got {n}')
    head = struct.pack('B', n) + struct.pack(f'{n}I', *(len(s) for s in list_of_strings))
    return head + b''.join(list_of_strings)


def unpack_byte_string(string):
    n = struct.unpack_from('B', string, 0)[0]
    lengths = struct.unpack_from(f'{n}I', string, 1)
    body = memoryview(string)[1 + 4 * n:]
    assert sum(lengths) == len(body), f'{sum(lengths)=} should equal to {len(body)=}'
    out, pos = [], 0
    for ln in lengths:
        out.append(bytes(body[pos:pos + ln]))
        pos += ln
    return out


def pad_divisible_by(img, div=64):
    """Edge-pad a PIL image (right/bottom) so both sides divide by `div` (coding.py:73-91)."""
    import torchvision.transforms.functional as tvf
    h_old, w_old = img.height, img.width
    if h_old % div == 0 and w_old % div == 0:
        return img
    h_tgt = div * math.ceil(h_old / div)
    w_tgt = div * math.ceil(w_old / div)
    return tvf.pad(img, padding=(0, 0, w_tgt - w_old, h_tgt - h_old), padding_mode='edge')


def crop_divisible_by(img, div=64):
    """Center-crop a PIL image so both sides divide by `div` (coding.py:94-110)."""
    import torchvision.transforms.functional as tvf
    h_old, w_old = img.height, img.width
    if h_old % div == 0 and w_old % div == 0:
        return img
    return tvf.center_crop(img, output_size=(div * (h_old // div), div * (w_old // div)))


def bd_rate(r1, psnr1, r2, psnr2):
    """Bjontegaard delta rate of curve 2 relative to curve 1 (coding.py:113-164): cubic fit of
    log-rate over PSNR, integrated over the common PSNR interval; returns percent."""
    lr1, lr2 = np.log(np.asarray(r1, dtype=np.float64)), np.log(np.asarray(r2, dtype=np.float64))
    p1 = np.polyfit(psnr1, lr1, 3)
    p2 = np.polyfit(psnr2, lr2, 3)
    lo = max(min(psnr1), min(psnr2))
    hi = min(max(psnr1), max(psnr2))
    i1, i2 = np.polyint(p1), np.polyint(p2)
    int1 = np.polyval(i1, hi) - np.polyval(i1, lo)
    int2 = np.polyval(i2, hi) - np.polyval(i2, lo)
    avg_diff = (int2 - int1) / (hi - lo)
    return float((np.exp(avg_diff) - 1) * 100)
