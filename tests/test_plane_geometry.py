"""Host-side check of the coefficient plane's block mapping (psxavenc_b200/csrc/bs_encode.h):
every bitstream-order block has exactly one plane slot, chroma and luma never share a group of
32, and the group counts match what the kernels are launched with. Compiles the header with g++."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r"""
#include <cstdio>
#include <vector>
#include "bs_encode.h"
using namespace psxb200;
int main() {
	const int sizes[][2] = {{16, 16}, {32, 16}, {16, 32}, {256, 16}, {272, 16}, {64, 48}, {176, 144}, {320, 240}, {640, 480}, {640, 512}};
	for (auto &wh : sizes) {
		BsGeometry g(wh[0], wh[1]);
		const int cpad = g.cgroups * 32;
		if (g.nblk != 6 * g.nmb || g.nsgroups != (g.nblk + 31) / 32) return 1;
		if (g.ngroups != g.cgroups + (4 * g.nmb + 31) / 32 || g.frame_stride_u4 != (size_t)g.ngroups * BS_U4_PER_BLOCK * 32) return 2;
		std::vector<int> seen(g.nblk, 0);
		for (int pi = 0; pi < g.ngroups * 32; pi++) {
			int b = bs_plane_to_block(pi, cpad, g.nmb);
			if (b < 0) continue;
			if (b >= g.nblk) return 3;
			seen[b]++;
			const bool chroma_block = b % 6 < 2;
			if (chroma_block != (pi < cpad)) return 4;      /* groups are of one kind */
			if (b / 6 != (pi < cpad ? pi / 2 : (pi - cpad) / 4)) return 5;   /* macroblock order kept */
		}
		for (int b = 0; b < g.nblk; b++)
			if (seen[b] != 1) return 6;
	}
	std::puts("ok");
	return 0;
}
"""


def test_plane_mapping_is_a_bijection(tmp_path):
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isdir(cuda_inc):
        pytest.skip("CUDA headers not installed")
    src = tmp_path / "geo.cpp"
    src.write_text(SRC)
    exe = tmp_path / "geo"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "psxavenc_b200", "csrc"),
                    "-I", os.path.join(ROOT, "include"), "-I", cuda_inc, "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout, out.stderr)
