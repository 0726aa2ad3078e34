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


def test_list_entry_fields_cannot_overflow(restated):
    """A list entry packs y = floor(2|c| / quant) into 10 bits beside the 6-bit position
    (bs_encode.cu: (y << 6) | i): y must stay below 1024 for any 8-bit input. The coefficient of
    a basis function is largest for its own 0/255 sign pattern, so those 128 blocks (and their
    row/column-only variants) bound |c| for both FDCT variants; also the census' run-0 code
    lengths must be the shortest of each level and grow with the level."""
    import re
    n = np.arange(8)
    blocks = []
    for u in range(8):
        for v in range(8):
            basis = np.outer(np.cos((2 * n + 1) * u * np.pi / 16), np.cos((2 * n + 1) * v * np.pi / 16))
            for pat in (basis > 0, basis < 0, basis >= 0):
                blocks.append(np.where(pat, 127, -128))
    rng = np.random.default_rng(0)
    blocks += [rng.choice([-128, 127], size=(8, 8)) for _ in range(2000)]
    blocks = np.stack(blocks).astype(np.int16).reshape(-1, 64)
    zigzag = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
              35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]
    quant_zz = np.array([QUANT[r] for r in zigzag], dtype=np.int64)
    for variant in (0, 1):
        co = np.abs(restated.fdct(variant, blocks).astype(np.int64))[:, zigzag]
        y = (2 * co[:, 1:]) // quant_zz[None, 1:]
        assert co[:, 1:].max() < 8192 and y.max() < 1024, (variant, int(co[:, 1:].max()), int(y.max()))
    # VLC lengths: the table the kernels use (generated from the MPEG-1 code strings)
    import os
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "psxavenc_b200", "csrc", "bs_tables.h")).read()
    m = re.search(r"BS_AC_VLC\[[^\]]*\]\s*=\s*\{(.*?)\};", text, re.S)
    vlc = np.array([int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", m.group(1))]).reshape(32, 41)
    lens = np.full((64, 64), 22)
    for lv in range(1, 41):
        for run in range(32):
            if vlc[run, lv]:
                lens[lv, run] = vlc[run, lv] >> 24
    assert all(lens[lv].min() == lens[lv, 0] for lv in range(1, 64))
    assert all(lens[lv, 0] <= lens[lv + 1, 0] for lv in range(1, 63))
