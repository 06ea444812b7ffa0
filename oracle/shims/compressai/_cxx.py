"""pmf_to_quantized_cdf restated from CompressAI cpp_exts/ops/ops.cpp (ryg_rans derived)."""
import numpy as np


def pmf_to_quantized_cdf(pmf, precision=16):
    pmf = [float(np.float32(p)) for p in pmf]
    for p in pmf:
        if p < 0 or not np.isfinite(p):
            raise ValueError(f"Invalid `pmf`, non-finite or negative element found: {p}")
    n = len(pmf)
    cdf = [0] * (n + 1)
    for i, p in enumerate(pmf):
        # std::round(float * int) -> float arithmetic, round half away from zero
        v = np.float32(np.float32(p) * np.float32(1 << precision))
        cdf[i + 1] = int(np.floor(float(v) + 0.5))
    total = sum(cdf)
    if total == 0:
        raise ValueError("Invalid `pmf`: at least one element must have a non-zero probability.")
    cdf = [((1 << precision) * c) // total for c in cdf]
    for i in range(1, n + 1):
        cdf[i] += cdf[i - 1]
    cdf[-1] = 1 << precision
    for i in range(n):
        if cdf[i] == cdf[i + 1]:
            best_freq, best_steal = 1 << 62, -1
            for j in range(n):
                freq = cdf[j + 1] - cdf[j]
                if 1 < freq < best_freq:
                    best_freq, best_steal = freq, j
            assert best_steal != -1
            if best_steal < i:
                for j in range(best_steal + 1, i + 1):
                    cdf[j] -= 1
            else:
                assert best_steal > i
                for j in range(i + 1, best_steal + 1):
                    cdf[j] += 1
    assert cdf[0] == 0 and cdf[-1] == (1 << precision)
    for i in range(n):
        assert cdf[i + 1] > cdf[i]
    return cdf
