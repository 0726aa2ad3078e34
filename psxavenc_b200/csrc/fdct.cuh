// Integer 8x8 forward DCTs, bit-exact models of the two FFmpeg AVDCT.fdct variants the
// reference can reach through `state->dct_context->fdct(block)` (psxavenc/mdec.c:640):
//
//   FDCT_ISLOW  ff_jpeg_fdct_islow_8  (FFmpeg built without x86 SIMD — the reference's
//               official release binaries, .github/scripts/build.sh:55 — and all non-x86)
//   FDCT_SSE2   ff_fdct_sse2          (default on a SIMD-enabled x86-64 FFmpeg)
//
// Both are specified in SURVEY.md Appendix A and pinned against the libavcodec 62.11.100
// binary by tests/test_oracle_vs_reference.py (via the oracle) and tests/test_gpu_bs.py.
// One thread transforms one block held entirely in registers (64 x int32): for 8-bit input
// every intermediate stays far inside int16, so the saturating/wrapping 16-bit steps of the
// SSE2 variant reduce to plain integer arithmetic.
#pragma once

#include <stdint.h>

enum { FDCT_ISLOW = 0, FDCT_SSE2 = 1 };

__device__ __forceinline__ int rshr(int v, int n) { return (v + (1 << (n - 1))) >> n; }

// ---- islow: 13-bit constants, 4 bits of extra precision after the row pass -------------
template <bool SECOND>
__device__ __forceinline__ void islow_pass(int &d0, int &d1, int &d2, int &d3, int &d4, int &d5, int &d6, int &d7) {
	constexpr int SH = SECOND ? 17 : 9;
	int e0 = d0 + d7, o0 = d0 - d7;
	int e1 = d1 + d6, o1 = d1 - d6;
	int e2 = d2 + d5, o2 = d2 - d5;
	int e3 = d3 + d4, o3 = d3 - d4;
	int ee0 = e0 + e3, eo0 = e0 - e3;
	int ee1 = e1 + e2, eo1 = e1 - e2;

	if (SECOND) {
		d0 = rshr(ee0 + ee1, 4);
		d4 = rshr(ee0 - ee1, 4);
	} else {
		d0 = (ee0 + ee1) << 4;
		d4 = (ee0 - ee1) << 4;
	}
	int z = (eo1 + eo0) * 4433;
	d2 = rshr(z + eo0 * 6270, SH);
	d6 = rshr(z - eo1 * 15137, SH);

	int z1 = o3 + o0, z2 = o2 + o1, z3 = o3 + o1, z4 = o2 + o0;
	int z5 = (z3 + z4) * 9633;
	z3 = z3 * -16069 + z5;
	z4 = z4 * -3196 + z5;
	z1 *= -7373;
	z2 *= -20995;
	d7 = rshr(o3 * 2446 + z1 + z3, SH);
	d5 = rshr(o2 * 16819 + z2 + z4, SH);
	d3 = rshr(o1 * 25172 + z2 + z3, SH);
	d1 = rshr(o0 * 12299 + z1 + z4, SH);
}

// ---- sse2: 16-bit column butterflies with >>16 multiplies, then a MAC row pass ---------
__device__ __forceinline__ int mulhi16(int a, int b) { return (a * b) >> 16; }

__device__ __forceinline__ void sse2_column(int &x0, int &x1, int &x2, int &x3, int &x4, int &x5, int &x6, int &x7) {
	constexpr int TAN1 = 13036, TAN2 = 27146, TAN3 = -21746, COS4 = 23170;
	int s16 = (x1 + x6) << 3, s25 = (x2 + x5) << 3, s07 = (x0 + x7) << 3, s34 = (x3 + x4) << 3;
	int m12 = s16 - s25, p12 = s16 + s25, m03 = s07 - s34, p03 = s07 + s34;
	int d16 = (x1 - x6) << 4, d25 = (x2 - x5) << 4, d34 = (x3 - x4) << 3, d07 = (x0 - x7) << 3;
	int p65 = mulhi16(d16 + d25, COS4) | 1;
	int m65 = mulhi16(d16 - d25, COS4);
	int p465 = d34 + m65, m465 = d34 - m65, m765 = d07 - p65, p765 = d07 + p65;

	x0 = p03 + p12;
	x4 = p03 - p12;
	x2 = (mulhi16(m12, TAN2) + m03) | 1;
	x6 = (mulhi16(m03, TAN2) - m12) | 1;
	x1 = (mulhi16(p465, TAN1) + p765) | 1;
	x3 = m765 - (mulhi16(m465, TAN3) + m465);
	x5 = mulhi16(m765, TAN3) + m765 + m465;
	x7 = mulhi16(p765, TAN1) - p465;
}

template <int C1, int C2, int C3, int C4, int C5, int C6, int C7>
__device__ __forceinline__ void sse2_row(int &a0, int &a1, int &a2, int &a3, int &a4, int &a5, int &a6, int &a7) {
	int s0 = a0 + a7, s1 = a1 + a6, s2 = a2 + a5, s3 = a3 + a4;
	int d0 = a0 - a7, d1 = a1 - a6, d2 = a2 - a5, d3 = a3 - a4;
	int e03 = s0 - s3, e12 = s1 - s2;
	a0 = ((s0 + s1 + s2 + s3) * C4 + 65536) >> 17;
	a4 = ((s0 - s1 - s2 + s3) * C4 + 65536) >> 17;
	a2 = (e03 * C2 + e12 * C6 + 65536) >> 17;
	a6 = (e03 * C6 - e12 * C2 + 65536) >> 17;
	a1 = (d0 * C1 + d1 * C3 + d2 * C5 + d3 * C7 + 65536) >> 17;
	a3 = (d0 * C3 - d1 * C7 - d2 * C1 - d3 * C5 + 65536) >> 17;
	a5 = (d0 * C5 - d1 * C1 + d2 * C7 + d3 * C3 + 65536) >> 17;
	a7 = (d0 * C7 - d1 * C5 + d2 * C3 - d3 * C1 + 65536) >> 17;
}

#define FDCT_ROW(v, r) v[8 * (r) + 0], v[8 * (r) + 1], v[8 * (r) + 2], v[8 * (r) + 3], v[8 * (r) + 4], v[8 * (r) + 5], v[8 * (r) + 6], v[8 * (r) + 7]
#define FDCT_COL(v, c) v[(c)], v[8 + (c)], v[16 + (c)], v[24 + (c)], v[32 + (c)], v[40 + (c)], v[48 + (c)], v[56 + (c)]

// In-place transform of v[8*row + col] (level-shifted samples in, x8-scaled DCT out).
template <int VARIANT>
__device__ __forceinline__ void fdct8x8(int (&v)[64]) {
	if (VARIANT == FDCT_ISLOW) {
#pragma unroll
		for (int r = 0; r < 8; r++) islow_pass<false>(FDCT_ROW(v, r));
#pragma unroll
		for (int c = 0; c < 8; c++) islow_pass<true>(FDCT_COL(v, c));
	} else {
#pragma unroll
		for (int c = 0; c < 8; c++) sse2_column(FDCT_COL(v, c));
		sse2_row<22725, 21407, 19266, 16384, 12873, 8867, 4520>(FDCT_ROW(v, 0));
		sse2_row<31521, 29692, 26722, 22725, 17855, 12299, 6270>(FDCT_ROW(v, 1));
		sse2_row<29692, 27969, 25172, 21407, 16819, 11585, 5906>(FDCT_ROW(v, 2));
		sse2_row<26722, 25172, 22654, 19266, 15137, 10426, 5315>(FDCT_ROW(v, 3));
		sse2_row<22725, 21407, 19266, 16384, 12873, 8867, 4520>(FDCT_ROW(v, 4));
		sse2_row<26722, 25172, 22654, 19266, 15137, 10426, 5315>(FDCT_ROW(v, 5));
		sse2_row<29692, 27969, 25172, 21407, 16819, 11585, 5906>(FDCT_ROW(v, 6));
		sse2_row<31521, 29692, 26722, 22725, 17855, 12299, 6270>(FDCT_ROW(v, 7));
	}
}
