"""The arithmetic identities the CUDA path rests on, checked exhaustively on the CPU:

* DIVIDE_ROUNDED(c, quant*q) (mdec.c:438: round((double)n / d), half away from zero) equals
  floor((floor(2|c| / quant) + q) / 2q) for every coefficient magnitude, quantiser entry and scale;
* the multiply-high reciprocals used on the device reproduce both floor divisions exactly
  (bs_encode.cu: ymagic_at, c_qmagic) — including the variant that tolerates 6 junk low bits;
* the byte-budget rule derived from flush_bits (mdec.c:321-333).
"""
import numpy as np

QUANT = [2, 16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37, 19, 22, 26, 27, 29, 34, 34, 38,
         22, 22, 26, 27, 29, 34, 37, 40, 22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32, 35, 40, 48, 58,
         26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83]      # mdec.c:189-198


def test_nested_floor_equals_divide_rounded():
    c = np.arange(0, 8300, dtype=np.int64)
    for quant in sorted(set(QUANT[1:])):
        y = (2 * c) // quant
        for q in range(1, 64):
            d = quant * q
            ref = np.floor(c.astype(np.float64) / d + 0.5).astype(np.int64)      # round half away, c >= 0
            assert np.array_equal((y + q) // (2 * q), ref), (quant, q)


def test_device_reciprocals_are_exact():
    mag = np.arange(0, 1 << 15, dtype=np.uint64)
    for quant in sorted(set(QUANT[1:])):
        magic = (1 << 33) // quant + 1
        assert magic < (1 << 32)
        assert np.array_equal((mag * magic) >> 32, (2 * mag) // quant), quant
    y = np.arange(0, 1024, dtype=np.uint64)
    for q in range(1, 64):
        m_hi = (1 << 32) // (2 * q) + 1
        m_lo = (1 << 32) // (128 * q) + 1
        want = (y + q) // (2 * q)
        assert np.array_equal(((y + q) * m_hi) >> 32, want)
        for junk in (0, 1, 31, 63):      # the position bits below y in a list entry
            t = (y << 6) + junk + (q << 6)
            assert np.array_equal((t * m_lo) >> 32, want), (q, junk)


def test_budget_rule_matches_byte_writer():
    """A stream of `bits` bits (blocks + 10-bit end code) fits iff 8 + 2*ceil(bits/16) <= max_size;
    simulated with the reference's writer: 16-bit words, low byte stored first, failure when the
    second byte would not fit (mdec.c:323-325)."""
    def writer_fits(bits, max_size):
        used = 8
        for _ in range((bits + 15) // 16):       # every started word is flushed (final flush pads it)
            used += 1
            if used >= max_size:
                return False
            used += 1
        return True
    for max_size in (8, 9, 10, 11, 12, 64, 65, 2016, 2017):
        for bits in range(1, 16 * 40):
            assert writer_fits(bits, max_size) == (8 + 2 * ((bits + 15) // 16) <= max_size), (bits, max_size)
