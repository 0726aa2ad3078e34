#!/usr/bin/env python3
"""Generate the constant tables of the MDEC/BS encoder as C headers.

Writes two byte-identical-in-content headers (different symbol prefixes):
  psxavenc_b200/csrc/bs_tables.h   (product; prefix BS_)
  oracle/orc_tables.h              (test oracle; prefix ORC_)

The master data below is the MPEG-1 (ISO 11172-2) Table B.14 run/level VLC that the MDEC
bitstream uses, written as bit strings; it carries the same codes as the reference's
`ac_huffman_tree` (psxavenc/mdec.c:39-157), which stores them as (length, value) pairs.
DC size codes are MPEG-1 Tables B.12/B.13 == `dc_y_huffman_tree`/`dc_c_huffman_tree`
(mdec.c:159-187). The quantiser matrix is `quant_dec` (mdec.c:189-198) and the scan is the
classic zig-zag (== `dct_zagzig_table`, mdec.c:213-222), generated here algorithmically.

Packed entry format (both AC and DC): (nbits << 24) | code, MSB-first, nbits includes the
sign / magnitude bits. AC entries are stored for the POSITIVE level with the sign bit
(LSB) clear; OR in 1 for negative levels (mdec.c:282-283).
"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

AC_VLC = """
0 1 11|0 2 0100|0 3 00101|0 4 0000110|0 5 00100110|0 6 00100001|0 7 0000001010
0 8 000000011101|0 9 000000011000|0 10 000000010011|0 11 000000010000
0 12 0000000011010|0 13 0000000011001|0 14 0000000011000|0 15 0000000010111
0 16 00000000011111|0 17 00000000011110|0 18 00000000011101|0 19 00000000011100
0 20 00000000011011|0 21 00000000011010|0 22 00000000011001|0 23 00000000011000
0 24 00000000010111|0 25 00000000010110|0 26 00000000010101|0 27 00000000010100
0 28 00000000010011|0 29 00000000010010|0 30 00000000010001|0 31 00000000010000
0 32 000000000011000|0 33 000000000010111|0 34 000000000010110|0 35 000000000010101
0 36 000000000010100|0 37 000000000010011|0 38 000000000010010|0 39 000000000010001
0 40 000000000010000
1 1 011|1 2 000110|1 3 00100101|1 4 0000001100|1 5 000000011011|1 6 0000000010110
1 7 0000000010101|1 8 000000000011111|1 9 000000000011110|1 10 000000000011101
1 11 000000000011100|1 12 000000000011011|1 13 000000000011010|1 14 000000000011001
1 15 0000000000010011|1 16 0000000000010010|1 17 0000000000010001|1 18 0000000000010000
2 1 0101|2 2 0000100|2 3 0000001011|2 4 000000010100|2 5 0000000010100
3 1 00111|3 2 00100100|3 3 000000011100|3 4 0000000010011
4 1 00110|4 2 0000001111|4 3 000000010010
5 1 000111|5 2 0000001001|5 3 0000000010010
6 1 000101|6 2 000000011110|6 3 0000000000010100
7 1 000100|7 2 000000010101
8 1 0000111|8 2 000000010001
9 1 0000101|9 2 0000000010001
10 1 00100111|10 2 0000000010000
11 1 00100011|11 2 0000000000011010
12 1 00100010|12 2 0000000000011001
13 1 00100000|13 2 0000000000011000
14 1 0000001110|14 2 0000000000010111
15 1 0000001101|15 2 0000000000010110
16 1 0000001000|16 2 0000000000010101
17 1 000000011111|18 1 000000011010|19 1 000000011001|20 1 000000010111|21 1 000000010110
22 1 0000000011111|23 1 0000000011110|24 1 0000000011101|25 1 0000000011100|26 1 0000000011011
27 1 0000000000011111|28 1 0000000000011110|29 1 0000000000011101|30 1 0000000000011100
31 1 0000000000011011
"""

# dct_dc_size codes, index = number of magnitude bits (0..8)
DC_SIZE_LUMA = ["100", "00", "01", "101", "110", "1110", "11110", "111110", "1111110"]
DC_SIZE_CHROMA = ["00", "01", "10", "110", "1110", "11110", "111110", "1111110", "11111110"]

QUANT = [
    2, 16, 19, 22, 26, 27, 29, 34,
    16, 16, 22, 24, 27, 29, 34, 37,
    19, 22, 26, 27, 29, 34, 34, 38,
    22, 22, 26, 27, 29, 34, 37, 40,
    22, 26, 27, 29, 32, 35, 40, 48,
    26, 27, 29, 32, 35, 40, 48, 58,
    26, 27, 29, 34, 38, 46, 56, 69,
    27, 29, 35, 38, 46, 56, 69, 83,
]

AC_RUNS, AC_LEVELS = 32, 41  # dense LUT [run][|level|], level 0 unused


def zigzag_scan():
    """scan[i] = raster index of the i-th coefficient in zig-zag order."""
    order = []
    for s in range(15):
        diag = [(y, s - y) for y in range(8) if 0 <= s - y < 8]
        if s % 2 == 0:
            diag.reverse()  # even diagonals run bottom-left -> top-right
        order += [y * 8 + x for y, x in diag]
    return order


def ac_table():
    tab = [[0] * AC_LEVELS for _ in range(AC_RUNS)]
    n = 0
    for item in AC_VLC.replace("\n", "|").split("|"):
        item = item.strip()
        if not item:
            continue
        run, level, bits = item.split()
        run, level = int(run), int(level)
        nbits = len(bits) + 1
        tab[run][level] = (nbits << 24) | (int(bits, 2) << 1)
        n += 1
    assert n == 111, n
    return tab


def dc_table(size_codes):
    """Entry for delta & 0x1FF, delta in -255..255 (mdec.c:285-318); 0 where no code exists."""
    tab = [0] * 512
    tab[0] = (len(size_codes[0]) << 24) | int(size_codes[0], 2)
    for delta in range(-255, 256):
        if delta == 0:
            continue
        mag = abs(delta).bit_length()
        extra = delta if delta > 0 else delta + (1 << mag) - 1
        prefix = size_codes[mag]
        tab[delta & 0x1FF] = ((len(prefix) + mag) << 24) | (int(prefix, 2) << mag) | extra
    return tab


def fmt_array(vals, per_line, fmt):
    lines = []
    for i in range(0, len(vals), per_line):
        lines.append("\t" + ", ".join(fmt % v for v in vals[i:i + per_line]) + ",")
    return "\n".join(lines)


def emit(path, prefix, guard):
    zz = zigzag_scan()
    ac = ac_table()
    out = []
    out.append("/* GENERATED by tools/gen_bs_tables.py -- do not edit. */")
    out.append("#ifndef %s\n#define %s\n#include <stdint.h>\n" % (guard, guard))
    out.append("#define %sAC_RUNS %d\n#define %sAC_LEVELS %d" % (prefix, AC_RUNS, prefix, AC_LEVELS))
    out.append("/* escape: 000001 + 6-bit run + 10-bit level = 22 bits (mdec.c:258) */")
    out.append("#define %sAC_ESCAPE_BITS 22\n" % prefix)
    out.append("/* raster index of the i-th zig-zag coefficient (mdec.c:213-222) */")
    out.append("static const uint8_t %sZIGZAG[64] = {\n%s\n};\n" % (prefix, fmt_array(zz, 8, "%2d")))
    out.append("#define %sZIGZAG_LIST %s\n" % (prefix, ", ".join(str(z) for z in zz)))
    out.append("/* quantiser matrix, raster order (mdec.c:189-198) */")
    out.append("static const uint8_t %sQUANT[64] = {\n%s\n};\n" % (prefix, fmt_array(QUANT, 8, "%2d")))
    out.append("#define %sQUANT_ZZ_LIST %s\n" % (prefix, ", ".join(str(QUANT[z]) for z in zz)))
    out.append("/* quantiser matrix in zig-zag order */")
    out.append("static const uint8_t %sQUANT_ZZ[64] = {\n%s\n};\n" % (prefix, fmt_array([QUANT[z] for z in zz], 8, "%2d")))
    out.append("/* [run][|level|] -> (nbits<<24)|code with sign bit clear; 0 = escape (mdec.c:39-157, 278-284) */")
    flat = [v for row in ac for v in row]
    out.append("static const uint32_t %sAC_VLC[%d * %d] = {\n%s\n};\n" % (prefix, AC_RUNS, AC_LEVELS, fmt_array(flat, 8, "0x%08x")))
    out.append("/* v3 DC delta codes, index delta & 0x1FF (mdec.c:159-187, 270-272, 285-318) */")
    out.append("static const uint32_t %sDC_VLC_CHROMA[512] = {\n%s\n};\n" % (prefix, fmt_array(dc_table(DC_SIZE_CHROMA), 8, "0x%08x")))
    out.append("static const uint32_t %sDC_VLC_LUMA[512] = {\n%s\n};\n" % (prefix, fmt_array(dc_table(DC_SIZE_LUMA), 8, "0x%08x")))
    out.append("#endif")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    emit(os.path.join(ROOT, "psxavenc_b200", "csrc", "bs_tables.h"), "BS_", "PSXB200_BS_TABLES_H")
    emit(os.path.join(ROOT, "oracle", "orc_tables.h"), "ORC_", "PSX_ORACLE_TABLES_H")
    print("tables written")
