"""Seeded synthetic inputs shared by oracle/gen_golden.py (fixture generation), tests/ and bench.py.

TEST INFRASTRUCTURE ONLY.  Fixtures hold outputs only; inputs are regenerated from these seeds.
"""
import torch


def synth_image(nB, H, W, seed):
    """Seeded smooth-ish RGB in [0,1]: random low-frequency waves + noise (more image-like than iid)."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing='ij')
    im = torch.zeros(nB, 3, H, W)
    for b in range(nB):
        for c in range(3):
            acc = torch.zeros(H, W)
            for _ in range(6):
                fx, fy, ph, amp = (torch.rand(4, generator=g) * torch.tensor([9.0, 9.0, 6.283, 0.5])).tolist()
                acc += amp * torch.sin(6.283 * (fx * xx + fy * yy) + ph)
            im[b, c] = acc
    im = im * 0.25 + 0.5 + 0.08 * torch.randn(nB, 3, H, W, generator=g)
    return im.clamp_(0, 1).contiguous()


CASES = {
    # name: (kind, nB, H, W, lambdas, seed)
    'qarv_rand_1x64x64': ('rand', 1, 64, 64, [2048.0], 0),          # BASELINE config 1
    'qarv_rand_2x128x192': ('rand', 2, 128, 192, [2048.0, 64.0], 1),
    'qarv_synth_1x256x256': ('synth', 1, 256, 256, [256.0], 2),
    'qarv_synth_3x64x128': ('synth', 3, 64, 128, [16.0, 700.0, 2048.0], 3),
}


def make_input(kind, nB, H, W, seed):
    if kind == 'rand':
        return torch.rand(nB, 3, H, W, generator=torch.Generator().manual_seed(seed))
    return synth_image(nB, H, W, seed)


# rd model fixtures: name -> (kind, nB, H, W, lambdas, image seed, noise seed)
RD_CASES = {
    'rd_rand_1x64x64': ('rand', 1, 64, 64, [256.0], 10, 5),
    'rd_synth_2x128x128': ('synth', 2, 128, 128, [16.0, 1024.0], 11, 6),
}

# qres34m fixtures (lambda fixed at 2048): name -> (kind, nB, H, W, image seed, train-noise seed)
QRES_LMB = 2048
QRES_CASES = {
    'qres_rand_2x64x128': ('rand', 2, 64, 128, 3, 21),
    'qres_synth_1x192x256': ('synth', 1, 192, 256, 12, 22),
}

# qres34m_lossless fixtures: name -> (kind, nB, H, W, image seed, train-noise seed); images are rounded to 8 bits
LOSSLESS_CASES = {
    'qresll_synth_2x64x128': ('synth', 2, 64, 128, 9, 31),
}


def make_input_8bit(kind, nB, H, W, seed):
    return (make_input(kind, nB, H, W, seed) * 255).round() / 255
