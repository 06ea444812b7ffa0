"""rANS coder restated from CompressAI cpp_exts/rans/rans_interface.cpp + ryg_rans rans64.h.
Pure Python (slow): used only to pin the C++ host coder of the product on small cases."""
import struct

PRECISION = 16
BYPASS_PRECISION = 4
MAX_BYPASS_VAL = (1 << BYPASS_PRECISION) - 1
RANS64_L = 1 << 31
MASK64 = (1 << 64) - 1


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        syms = []  # (start, range, bypass)
        for s, ci in zip(symbols, indexes):
            cdf = cdfs[ci]
            max_value = cdfs_sizes[ci] - 2
            value = s - offsets[ci]
            raw_val = 0
            if value < 0:
                raw_val = -2 * value - 1
                value = max_value
            elif value >= max_value:
                raw_val = 2 * (value - max_value)
                value = max_value
            syms.append((cdf[value], cdf[value + 1] - cdf[value], False))
            if value == max_value:
                n_bypass = 0
                while (raw_val >> (n_bypass * BYPASS_PRECISION)) != 0:
                    n_bypass += 1
                val = n_bypass
                while val >= MAX_BYPASS_VAL:
                    syms.append((MAX_BYPASS_VAL, MAX_BYPASS_VAL + 1, True))
                    val -= MAX_BYPASS_VAL
                syms.append((val, val + 1, True))
                for j in range(n_bypass):
                    v = (raw_val >> (j * BYPASS_PRECISION)) & MAX_BYPASS_VAL
                    syms.append((v, v + 1, True))
        x = RANS64_L
        out = []  # 32-bit words, emitted back to front
        for start, rng, bypass in reversed(syms):
            if not bypass:
                freq = rng
                x_max = ((RANS64_L >> PRECISION) << 32) * freq
                if x >= x_max:
                    out.append(x & 0xFFFFFFFF)
                    x >>= 32
                x = ((x // freq) << PRECISION) + (x % freq) + start
            else:
                freq = 1 << (16 - BYPASS_PRECISION)
                x_max = ((RANS64_L >> 16) << 32) * freq
                if x >= x_max:
                    out.append(x & 0xFFFFFFFF)
                    x >>= 32
                x = ((x << BYPASS_PRECISION) | start) & MASK64
        out.append((x >> 32) & 0xFFFFFFFF)
        out.append(x & 0xFFFFFFFF)
        out.reverse()
        return struct.pack(f'<{len(out)}I', *out)


class RansDecoder:
    def decode_with_indexes(self, encoded, indexes, cdfs, cdfs_sizes, offsets):
        words = struct.unpack(f'<{len(encoded)//4}I', encoded)
        pos = 2
        x = words[0] | (words[1] << 32)

        def get_bits(nbits):
            nonlocal x, pos
            val = x & ((1 << nbits) - 1)
            x >>= nbits
            if x < RANS64_L:
                x = (x << 32) | words[pos]
                pos += 1
            return val

        output = []
        mask = (1 << PRECISION) - 1
        for ci in indexes:
            cdf = cdfs[ci]
            size = cdfs_sizes[ci]
            max_value = size - 2
            cum = x & mask
            s = 0
            while s + 1 < size and cdf[s + 1] <= cum:
                s += 1
            start, freq = cdf[s], cdf[s + 1] - cdf[s]
            x = freq * (x >> PRECISION) + (x & mask) - start
            if x < RANS64_L:
                x = (x << 32) | words[pos]
                pos += 1
            value = s
            if value == max_value:
                val = get_bits(BYPASS_PRECISION)
                n_bypass = val
                while val == MAX_BYPASS_VAL:
                    val = get_bits(BYPASS_PRECISION)
                    n_bypass += val
                raw_val = 0
                for j in range(n_bypass):
                    val = get_bits(BYPASS_PRECISION)
                    raw_val |= val << (j * BYPASS_PRECISION)
                value = raw_val >> 1
                if raw_val & 1:
                    value = -value - 1
                else:
                    value += max_value
            output.append(value + offsets[ci])
        return output
