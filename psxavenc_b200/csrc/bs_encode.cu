// MDEC "BS" v2/v3 frame encoder for sm_100a — the GPU side of encode_frame_bs
// (reference psxavenc/mdec.c:580-755).
//
// Data flow for a batch of n NV21 frames resident in HBM:
//
//   bs_dct_kernel      one thread per 8x8 block: gathers the block from the NV21 frame
//                      (mdec.c:605-634), level-shifts, runs the bit-exact integer FDCT in
//                      registers (mdec.c:640) and stores |coef| (u16) in zig-zag order plus a
//                      64-bit sign mask into a coefficient plane laid out so that the 32
//                      lanes of a warp (32 consecutive blocks in bitstream order) read and
//                      write 512 contiguous bytes per uint4 access.
//   bs_pack_kernel     one CTA per frame, one thread per block and quant scale:
//                      (1) first-fit quant-scale search q = 1,2,... (mdec.c:663-722): each
//                          thread prices its blocks' run/level codes from a shared-memory
//                          length LUT, the CTA sums and applies the byte-budget rule
//                          8 + 2*ceil(bits/16) <= frame_max_size (flush_bits, mdec.c:321-333);
//                      (2) exclusive scan of the per-block bit lengths in bitstream order;
//                      (3) every thread re-quantises its blocks at the winning q and ORs its
//                          codes into a shared-memory image of the bitstream at its bit
//                          offset (encode_dct_block / encode_bits, mdec.c:441-510, 335-385);
//                      (4) header (mdec.c:725-754) and coalesced copy-out with the 16-bit
//                          little-endian word order of the format, zero padded.
//
// Division by the quantiser step uses an exact 32-bit reciprocal (DIVIDE_ROUNDED,
// mdec.c:438, is round-half-away-from-zero == (|n| + d/2) / d in integers).
#include <cuda_runtime.h>
#include <stdint.h>

#include "bs_encode.h"
#include "bs_tables.h"
#include "fdct.cuh"

namespace psxb200 {

// ---- constant tables -------------------------------------------------------------------
// [q][i] -> (magic, half): level = umulhi(|coef| + half, magic), d = QUANT_ZZ[i] * q.
__constant__ uint2 c_qparam[64 * 64];
// [min(level,63)][run] -> code length in bits incl. sign (22 = escape), 0 for level 0.
__constant__ uint8_t c_lenlut[64 * 64];
// [min(level,63)][run] -> (len << 24) | code with the sign bit (LSB) clear; level 0 -> 0;
// escape -> (22 << 24) with code 0. Copied to shared memory by every CTA.
__device__ uint32_t g_vlc[64 * 64];
// v3 DC delta codes: [0] chroma, [1] luma.
__constant__ uint32_t c_dcvlc[2 * 512];

__host__ __device__ constexpr int zigzag_at(int i) {
	constexpr int t[64] = {BS_ZIGZAG_LIST};
	return t[i];
}

__host__ __device__ constexpr int quant_zz_at(int i) {
	constexpr int t[64] = {BS_QUANT_ZZ_LIST};
	return t[i];
}

// Smallest quantiser entry among the AC coefficients of plane row j (zig-zag 8j..8j+7).
__host__ __device__ constexpr int row_min_quant(int j) {
	int m = 255;
	for (int i = 8 * j; i < 8 * j + 8; i++)
		if (i > 0 && quant_zz_at(i) < m) m = quant_zz_at(i);
	return m;
}

void bs_upload_tables() {
	static uint2 qparam[64 * 64];
	static uint8_t lenlut[64 * 64];
	static uint32_t vlc[64 * 64];
	static uint32_t dcvlc[2 * 512];
	for (int q = 0; q < 64; q++) {
		for (int i = 0; i < 64; i++) {
			uint32_t d = (uint32_t)BS_QUANT_ZZ[i] * (q ? q : 1);
			if (i == 0) d = 16;
			qparam[q * 64 + i].x = (uint32_t)(0x100000000ull / d) + 1;
			qparam[q * 64 + i].y = d / 2;
		}
	}
	for (int lv = 0; lv < 64; lv++) {
		for (int run = 0; run < 64; run++) {
			int len = 0;
			if (lv > 0) {
				uint32_t e = (run < BS_AC_RUNS && lv < BS_AC_LEVELS) ? BS_AC_VLC[run * BS_AC_LEVELS + lv] : 0;
				len = e ? (int)(e >> 24) : BS_AC_ESCAPE_BITS;
			}
			lenlut[(lv << 6) | run] = (uint8_t)len;
			uint32_t e = (lv > 0 && run < BS_AC_RUNS && lv < BS_AC_LEVELS) ? BS_AC_VLC[run * BS_AC_LEVELS + lv] : 0;
			vlc[(lv << 6) | run] = lv == 0 ? 0u : (e ? e : (uint32_t)BS_AC_ESCAPE_BITS << 24);
		}
	}
	for (int i = 0; i < 512; i++) {
		dcvlc[i] = BS_DC_VLC_CHROMA[i];
		dcvlc[512 + i] = BS_DC_VLC_LUMA[i];
	}
	cudaMemcpyToSymbol(c_qparam, qparam, sizeof(qparam));
	cudaMemcpyToSymbol(c_lenlut, lenlut, sizeof(lenlut));
	cudaMemcpyToSymbol(g_vlc, vlc, sizeof(vlc));
	cudaMemcpyToSymbol(c_dcvlc, dcvlc, sizeof(dcvlc));
}

// ---- kernel 1: gather + FDCT -----------------------------------------------------------

__device__ __forceinline__ int byte_of(uint32_t w, int k) { return (int)((w >> (8 * k)) & 0xFF); }

template <int VARIANT>
__global__ void __launch_bounds__(BS_DCT_THREADS, BS_DCT_MIN_CTAS)
bs_dct_kernel(const uint8_t *__restrict__ frames, size_t frame_bytes, int n_frames, int width, int height,
              int mbh, uint32_t mbh_magic, int nmb, int cpad, int ngroups, uint4 *__restrict__ coefs,
              size_t frame_stride_u4) {
	// grid: x = chunk of 128 plane lanes within the frame, y = frame
	const int f = blockIdx.y;
	const int b = blockIdx.x * BS_DCT_THREADS + threadIdx.x;   // plane index (type-major, see BsGeometry)
	if (b >= ngroups * 32) return;                             // whole warps: ngroups * 32 lanes per frame
	// whole warps map to one group of 32 blocks of one kind; padding lanes idle but stay for
	// the warp reductions below
	const bool chroma = b < cpad;   // warp-uniform: cpad is a multiple of 32
	const bool active = chroma ? b < 2 * nmb : b - cpad < 4 * nmb;

	uint32_t sign_lo = 0, sign_hi = 0;
	uint32_t rowq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	uint4 *dst = coefs + (size_t)f * frame_stride_u4 + (size_t)(b >> 5) * (BS_U4_PER_BLOCK * 32) + (b & 31);

	if (active) {
		// macroblocks in bitstream order: columns outermost, rows next (mdec.c:689-704)
		int mb = chroma ? b >> 1 : (b - cpad) >> 2;
		int k = chroma ? b & 1 : 2 + ((b - cpad) & 3);
		int mx = mbh_magic ? (int)__umulhi((uint32_t)mb, mbh_magic) : mb, my = mb - mx * mbh;   // mb / mbh (magic 0: mbh == 1)
		const uint8_t *fr = frames + (size_t)f * frame_bytes;

		int v[64];
		if (chroma) {
			// interleaved CrCb plane: Cr at even bytes, Cb at odd (mdec.c:627-628)
			const uint8_t *p = fr + (size_t)width * height + (size_t)width * (my * 8) + mx * 16;
#pragma unroll
			for (int y = 0; y < 8; y++) {
				uint4 r = __ldg(reinterpret_cast<const uint4 *>(p + (size_t)y * width));
				uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
				for (int x = 0; x < 8; x++) {
					uint32_t pair = w[x >> 1] >> (16 * (x & 1));
					v[8 * y + x] = (int)((k ? (pair >> 8) : pair) & 0xFF);
				}
			}
		} else {
			int ox = ((k - 2) & 1) * 8, oy = ((k - 2) >> 1) * 8;
			const uint8_t *p = fr + (size_t)width * (my * 16 + oy) + mx * 16 + ox;
#pragma unroll
			for (int y = 0; y < 8; y++) {
				uint2 r = __ldg(reinterpret_cast<const uint2 *>(p + (size_t)y * width));
#pragma unroll
				for (int x = 0; x < 4; x++) {
					v[8 * y + x] = byte_of(r.x, x);
					v[8 * y + 4 + x] = byte_of(r.y, x);
				}
			}
		}

		// The reference level-shifts every sample by -128 first (mdec.c:627-632). Both FDCT
		// variants only ever take differences of samples except in the DC term, where the 64
		// offsets add up to exactly 8192 through both passes' exact scalings, so the shift is
		// applied once here.
		fdct8x8<VARIANT>(v);
		v[0] -= 8192;

		// Coefficients are visited in descending zig-zag order so that each sign can be shifted
		// into its mask with one funnel shift and coefficient i ends up at bit i of its half.
#pragma unroll
		for (int j = 7; j >= 0; j--) {
			uint32_t w[4];
			uint32_t rowmax = 0;
#pragma unroll
			for (int t = 3; t >= 0; t--) {
				int i0 = 8 * j + 2 * t;
				int c0 = v[zigzag_at(i0)], c1 = v[zigzag_at(i0 + 1)];
				uint32_t m0 = (uint32_t)abs(c0), m1 = (uint32_t)abs(c1);
				w[t] = m0 | (m1 << 16);
				rowmax = max(rowmax, i0 == 0 ? m1 : max(m0, m1));   // the DC term has its own fixed step
				if (i0 < 32) {
					sign_lo = __funnelshift_l((uint32_t)c1, sign_lo, 1);
					sign_lo = __funnelshift_l((uint32_t)c0, sign_lo, 1);
				} else {
					sign_hi = __funnelshift_l((uint32_t)c1, sign_hi, 1);
					sign_hi = __funnelshift_l((uint32_t)c0, sign_hi, 1);
				}
			}
			dst[j * 32] = make_uint4(w[0], w[1], w[2], w[3]);
			// A coefficient quantises to nonzero at scale q iff 2*|c| >= quant*q, so nothing in
			// this row survives beyond q = 2*rowmax / (smallest quant of the row).
			rowq[j] = min(63u, 2u * rowmax / (uint32_t)row_min_quant(j));
		}
	}

	// per group and plane row: the largest quant scale at which any of the 32 blocks still has a
	// nonzero coefficient there; the pack kernel skips rows that are dead at its q
	uint32_t packed_lo = 0, packed_hi = 0;
#pragma unroll
	for (int j = 6; j >= 0; j--) rowq[j] = max(rowq[j], rowq[j + 1]);   // a live row keeps its predecessors live
#pragma unroll
	for (int j = 0; j < 8; j++) {
		uint32_t m = __reduce_max_sync(0xFFFFFFFFu, rowq[j]);
		if (j < 4) packed_lo |= m << (8 * j); else packed_hi |= m << (8 * (j - 4));
	}
	if (active) {
		dst[8 * 32] = make_uint4(sign_lo, sign_hi, packed_lo, packed_hi);
	} else {
		// padding lanes of the last group: keep the plane fully defined (the pack kernel lets
		// them run along so that its warps stay convergent)
#pragma unroll
		for (int j = 0; j < 9; j++) dst[j * 32] = make_uint4(0, 0, 0, 0);
	}
}

// ---- kernel 2: quant-scale search + bit packing ------------------------------------------

struct PackSmem {
	uint32_t *stream;   // bitstream image, 32-bit words, first stream bit = bit 31 of word 0
	uint32_t *dctab;    // v3: DC delta codes, [0..511] chroma, [512..1023] luma (len<<24 | code)
	uint32_t *gtot;     // per group bit totals -> exclusive group bases
	uint2 *qpar;        // [2][64] copies of c_qparam rows: [q & 1] holds the current q's
	uint2 *rowq;        // per group: last live quant scale of each plane row (8 bytes, from bs_dct_kernel)
	uint32_t *misc;     // [0..2] rotating frame totals, [3] nonzero AC count, [8..] scan scratch
	uint32_t *vlc;      // [min(level,63)][run] -> (len<<24)|code, see g_vlc
	uint16_t *lens;     // per block bit length at the current q; after the scan: exclusive offset in its group
	int16_t *dcval;     // v3: per block quantised DC, replaced in place by its coded delta
	uint8_t *lenlut;
	uint8_t *lev;       // emit: min(level,63) of the thread's current block, [coef][thread] with stride lev_stride
};

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
	return v;
}

// Integer multiply-add pinned to the FMA pipe (IMAD). The pricing loop is bound by the ALU
// pipe (adds, min/max, compares, shifts-and-adds) while the FMA pipe idles; routing the index,
// run-length and accumulate arithmetic through IMAD balances the two.
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
	uint32_t d;
	asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}

// bs_dct_kernel records, per group of 32 blocks and per plane row (8 zig-zag positions), the
// largest quant scale at which any block of the group still has a nonzero level there. Rows
// die from the high-frequency end, so a warp only prices / stages the live prefix of rows of
// its group at the current q; every prefix length has its own straight-line instantiation
// (rows are fetched up to four at a time to bound the register footprint).

// Leading plane rows to process at quant scale q, rounded up to an instantiated prefix length
// (0, 2, 4 or 8); rowq bytes are non-increasing in the row index. The same in every lane.
__device__ __forceinline__ int live_prefix(uint2 rowq, int q) {
	if ((int)(rowq.y & 0xFFu) >= q) return 8;            // row 4 live
	if ((int)((rowq.x >> 16) & 0xFFu) >= q) return 4;    // row 2 live
	return (int)(rowq.x & 0xFFu) >= q ? 2 : 0;           // row 0 live
}

// Loads N (<= 4) magnitude rows starting at row R0 of block (group, lane).
template <int R0, int N>
__device__ __forceinline__ void load_rows(const uint4 *__restrict__ gp, uint32_t (&w)[4 * N]) {
#pragma unroll
	for (int j = 0; j < N; j++) {
		uint4 r = gp[(R0 + j) * 32];
		w[4 * j + 0] = r.x; w[4 * j + 1] = r.y; w[4 * j + 2] = r.z; w[4 * j + 3] = r.w;
	}
}

// |coef| number i of the loaded rows
template <int N>
__device__ __forceinline__ uint32_t mag_at(const uint32_t (&w)[4 * N], int i) {
	return (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xFFFFu);
}

// Prices the AC coefficients of rows R0..R0+N-1 (the pricing half of encode_dct_block).
// qpar: the current q's (reciprocal, half step) pairs in shared memory.
template <int R0, int N>
__device__ __forceinline__ void price_rows(const uint4 *__restrict__ gp, const uint2 *qpar, const uint8_t *lenlut,
                                           uint32_t &bits, uint32_t &run) {
	uint32_t w[4 * N];
	load_rows<R0, N>(gp, w);
	const uint2 *qp = qpar + 8 * R0;
#pragma unroll
	for (int i = (R0 ? 0 : 1); i < 8 * N; i++) {
		uint2 p = qp[i];
		uint32_t lv = __umulhi(mag_at<N>(w, i) + p.y, p.x);
		uint32_t m = min(lv, 63u);
		bits = imad(lenlut[imad(m, 64u, run)], 1u, bits);
		uint32_t z = imad(lv, 1u, 0xFFFFFFFFu) >> 31;      // 1 when the level is zero (lv < 2^31)
		run = imad(run, z, z);                              // (run + 1) * z
	}
}

// AC bit cost of one block whose rows >= NROWS are known to quantise to zero.
template <int NROWS>
__device__ __forceinline__ int ac_bits_prefix(const uint4 *__restrict__ gp, const uint2 *qpar, const uint8_t *lenlut) {
	uint32_t bits = 0, run = 0;
	if (NROWS > 0) price_rows<0, (NROWS < 4 ? NROWS : 4)>(gp, qpar, lenlut, bits, run);
	if (NROWS > 4) price_rows<4, (NROWS > 4 ? NROWS - 4 : 1)>(gp, qpar, lenlut, bits, run);
	return (int)bits;
}

__device__ __forceinline__ int ac_bits(const uint4 *__restrict__ gp, const uint2 *qpar, const uint2 *cpar, const uint8_t *lenlut, int prefix) {
	if (prefix == 8) return ac_bits_prefix<8>(gp, cpar, lenlut);   // full blocks: reciprocals via the constant bank
	if (prefix == 4) return ac_bits_prefix<4>(gp, qpar, lenlut);
	if (prefix == 2) return ac_bits_prefix<2>(gp, qpar, lenlut);
	return 0;
}

// Emit, convergent part: quantise rows R0..R0+N-1, park min(level,63) in the thread's column of
// the level staging area and return their nonzero mask (bit i = coefficient 8*R0 + i).
template <int R0, int N>
__device__ __forceinline__ uint32_t stage_rows(const uint4 *__restrict__ gp, const uint2 *qpar, uint8_t *lev, int lev_stride) {
	uint32_t w[4 * N];
	load_rows<R0, N>(gp, w);
	const uint2 *qp = qpar + 8 * R0;
	uint32_t nz = 0;
#pragma unroll
	for (int i = (R0 ? 0 : 1); i < 8 * N; i++) {
		uint2 p = qp[i];
		uint32_t m = min(__umulhi(mag_at<N>(w, i) + p.y, p.x), 63u);
		lev[(8 * R0 + i) * lev_stride] = (uint8_t)m;
		nz = imad(min(m, 1u), 1u << i, nz);
	}
	return nz;
}

template <int NROWS>
__device__ __forceinline__ void stage_prefix(const uint4 *__restrict__ gp, const uint2 *qpar, uint8_t *lev, int lev_stride,
                                             uint32_t &nz_lo, uint32_t &nz_hi) {
	nz_lo = NROWS > 0 ? stage_rows<0, (NROWS < 4 ? (NROWS > 0 ? NROWS : 1) : 4)>(gp, qpar, lev, lev_stride) : 0u;
	nz_hi = NROWS > 4 ? stage_rows<4, (NROWS > 4 ? NROWS - 4 : 1)>(gp, qpar, lev, lev_stride) : 0u;
}

__device__ __forceinline__ void stage_levels(const uint4 *__restrict__ gp, const uint2 *qpar, const uint2 *cpar, uint8_t *lev, int lev_stride,
                                             int prefix, uint32_t &nz_lo, uint32_t &nz_hi) {
	nz_lo = nz_hi = 0;
	if (prefix == 8) stage_prefix<8>(gp, cpar, lev, lev_stride, nz_lo, nz_hi);
	else if (prefix == 4) stage_prefix<4>(gp, qpar, lev, lev_stride, nz_lo, nz_hi);
	else if (prefix == 2) stage_prefix<2>(gp, qpar, lev, lev_stride, nz_lo, nz_hi);
}

// Appends MSB-first codes at an arbitrary bit position of the 32-bit-word stream image. Words
// are shared with neighbouring blocks at both ends, hence the atomic OR on flush.
struct BitWriter {
	uint32_t *words;   // next word to flush
	uint32_t cur;      // pending bits, left-aligned; the top `fill` bits are valid
	int fill;
	__device__ __forceinline__ void begin(uint32_t *w, uint32_t bitpos) {
		words = w + (bitpos >> 5); fill = (int)(bitpos & 31); cur = 0;
	}
	__device__ __forceinline__ void put(int len, uint32_t code) {   // 1 <= len <= 22, fill < 32 on entry
		int room = 32 - fill;
		if (len < room) {
			cur |= code << (room - len);
			fill += len;
		} else {
			int rem = len - room;
			atomicOr(words, cur | (code >> rem));
			words++;
			cur = rem ? code << (32 - rem) : 0u;
			fill = rem;
		}
	}
	__device__ __forceinline__ void finish() {
		if (fill) atomicOr(words, cur);
	}
};

// v3 DC prediction chain (mdec.c:455-461): last += 4*round((dc-last)/4) per plane. `last`
// stays a multiple of 4, so with L = last/4, dc = 4a + r: L' = a + (r==3) for r != 2 and
// L' = a + (L <= a) on exact ties. Each element is therefore a two-valued step function of
// the incoming L; such functions compose in closed form, which turns the chain into a scan.
struct DcFn { int t, lo, hi, valid; };
__device__ __forceinline__ int dc_apply(const DcFn &f, int L) { return !f.valid ? L : (L <= f.t ? f.lo : f.hi); }
__device__ __forceinline__ DcFn dc_then(const DcFn &f, const DcFn &g) {
	if (!g.valid) return f;
	if (!f.valid) return g;
	return DcFn{f.t, dc_apply(g, f.lo), dc_apply(g, f.hi), 1};
}
__device__ __forceinline__ DcFn dc_elem(int dc) {
	int a = dc >> 2, r = dc & 3;
	if (r == 2) return DcFn{a, a + 1, a, 1};
	int v = a + (r == 3);
	return DcFn{0, v, v, 1};
}
__device__ __forceinline__ DcFn dc_shfl_up(const DcFn &f, int d) {
	return DcFn{__shfl_up_sync(0xFFFFFFFFu, f.t, d), __shfl_up_sync(0xFFFFFFFFu, f.lo, d),
	            __shfl_up_sync(0xFFFFFFFFu, f.hi, d), __shfl_up_sync(0xFFFFFFFFu, f.valid, d)};
}

// All threads of the CTA call this; replaces dcval[] (quantised DC per block) by the delta
// that gets coded, for every block of the frame.
__device__ void dc_delta_codes(int codec, int nmb, int16_t *dcval, int *scratch /* 4*32 ints */) {
	const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = (T + 31) >> 5;
	for (int plane = 0; plane < 3; plane++) {
		int n = plane < 2 ? nmb : 4 * nmb;
		int chunk = (n + T - 1) / T;
		int lo = min(n, tid * chunk), hi = min(n, lo + chunk);
		auto block_of = [&](int i) { return plane < 2 ? 6 * i + plane : 6 * (i >> 2) + 2 + (i & 3); };

		DcFn acc{0, 0, 0, 0};
		for (int i = lo; i < hi; i++) acc = dc_then(acc, dc_elem(dcval[block_of(i)]));

		// inclusive scan across the CTA
		DcFn inc = acc;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			DcFn up = dc_shfl_up(inc, d);
			if (lane >= d) inc = dc_then(up, inc);
		}
		__syncthreads();
		if (lane == 31) {
			scratch[4 * wid + 0] = inc.t; scratch[4 * wid + 1] = inc.lo;
			scratch[4 * wid + 2] = inc.hi; scratch[4 * wid + 3] = inc.valid;
		}
		__syncthreads();
		DcFn pre{0, 0, 0, 0};   // everything before this warp
		for (int w = 0; w < wid && w < nw; w++)
			pre = dc_then(pre, DcFn{scratch[4 * w], scratch[4 * w + 1], scratch[4 * w + 2], scratch[4 * w + 3]});
		DcFn excl = dc_shfl_up(inc, 1);
		if (lane == 0) excl = DcFn{0, 0, 0, 0};
		excl = dc_then(pre, excl);

		int L = dc_apply(excl, 0);
		for (int i = lo; i < hi; i++) {
			int b = block_of(i);
			int Ln = dc_apply(dc_elem(dcval[b]), L);
			int delta = Ln - L;
			L = Ln;
			if (codec == 2) {           // v3dc wrap-around (mdec.c:469-474)
				if (delta < -0x80) delta += 0x100;
				else if (delta > 0x80) delta -= 0x100;
			}
			dcval[b] = (int16_t)delta;
		}
	}
	__syncthreads();
}

__device__ __forceinline__ int quant_dc(uint32_t mag, uint32_t negative) {
	// round(c/16) half away from zero, clamp to [-512, 510] (mdec.c:447-449, 262-265)
	int d = (int)((mag + 8) >> 4);
	d = negative ? -d : d;
	return max(-0x200, min(0x1FE, d));
}

template <bool V3, bool SMEM_STREAM, int MAX_THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(MAX_THREADS, MIN_CTAS)
bs_pack_kernel(const uint4 *__restrict__ coefs, size_t frame_stride_u4, int nblk, int ngroups, int nsgroups, int cpad,
               int nmb, int codec,
               const int *__restrict__ max_sizes, int max_size_bound, uint8_t *__restrict__ out, size_t out_stride,
               psxb200_bs_result_t *__restrict__ results, uint32_t *__restrict__ gstream, size_t gstream_stride,
               const BsStrLayout str) {
	extern __shared__ __align__(16) uint8_t smem_raw[];
	const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = T >> 5;
	const int f = blockIdx.x;
	const int padded = nsgroups * 32;   // blocks in bitstream order, rounded up to whole scan groups
	const int stream_words = (max_size_bound + 3) / 4 + 2;

	// fixed-size tables first so that their shared-memory addresses are compile-time offsets
	PackSmem s;
	{
		uint8_t *p = smem_raw;
		auto take = [&](size_t bytes) { uint8_t *at = p; p += (bytes + 15) & ~(size_t)15; return at; };
		s.lenlut = take(64 * 64);
		s.vlc = reinterpret_cast<uint32_t *>(take(4 * 64 * 64));
		s.misc = reinterpret_cast<uint32_t *>(take(4 * (8 + 4 * 32)));
		s.qpar = reinterpret_cast<uint2 *>(take(8 * 2 * 64));
		s.stream = reinterpret_cast<uint32_t *>(SMEM_STREAM ? take(4 * (size_t)stream_words) : p);
		s.dctab = reinterpret_cast<uint32_t *>(V3 ? take(4 * 1024) : p);
		s.gtot = reinterpret_cast<uint32_t *>(take(4 * (size_t)(nsgroups + 1)));
		s.rowq = reinterpret_cast<uint2 *>(take(8 * (size_t)ngroups));
		s.lens = reinterpret_cast<uint16_t *>(take(2 * (size_t)padded));
		s.dcval = reinterpret_cast<int16_t *>(V3 ? take(2 * (size_t)padded) : p);
		s.lev = p;
	}
	const int lev_stride = bs_lev_stride(T);
	uint32_t *stream = SMEM_STREAM ? s.stream : gstream + (size_t)f * gstream_stride;

	const uint4 *fc = coefs + (size_t)f * frame_stride_u4;
	// STR mode: budget and sector position follow from the frame index alone
	const long long str_k = (long long)str.frame_index0 + f;
	const long long str_before = str.sector_size ? (str_k - 1) * str.sectors_num / str.sectors_den : 0;
	int max_size = str.sector_size ? (int)(str_k * str.sectors_num / str.sectors_den - str_before) * 2016 : max_sizes[f];
	if (max_size > max_size_bound) max_size = 0;   // contract violation -> frame fails
	if (str.sector_size) out += (size_t)(str_before - str.sector0) * str.sector_size - (size_t)f * out_stride;
	const int words = max_size > 0 ? (max_size + 3) / 4 + 2 : 0;

	for (int i = tid; i < 64 * 64 / 4; i += T)
		reinterpret_cast<uint32_t *>(s.lenlut)[i] = reinterpret_cast<const uint32_t *>(c_lenlut)[i];
	for (int i = tid; i < 64 * 64; i += T) s.vlc[i] = g_vlc[i];
	if (V3) for (int i = tid; i < 1024; i += T) s.dctab[i] = c_dcvlc[i];
	for (int i = tid; i < words; i += T) stream[i] = 0;
	if (tid < 8) s.misc[tid] = 0;
	for (int b = nblk + tid; b < padded; b += T) s.lens[b] = 0;   // scan padding
	for (int i = tid; i < 64; i += T) s.qpar[64 + i] = c_qparam[64 + i];   // q = 1
	for (int g = tid; g < ngroups; g += T) {
		uint4 r8 = fc[(size_t)g * (BS_U4_PER_BLOCK * 32) + 8 * 32];   // lane 0's sign row carries the group's rowq
		s.rowq[g] = make_uint2(r8.z, r8.w);
	}
	if (V3) {
		for (int pi = tid; pi < ngroups * 32; pi += T) {
			int b = bs_plane_to_block(pi, cpad, nmb);
			if (b < 0) continue;
			const uint4 *gp = fc + (size_t)(pi >> 5) * (BS_U4_PER_BLOCK * 32) + (pi & 31);
			uint32_t w0 = reinterpret_cast<const uint32_t *>(gp)[0];
			uint32_t sg = reinterpret_cast<const uint32_t *>(gp + 8 * 32)[0];
			s.dcval[b] = (int16_t)quant_dc(w0 & 0xFFFFu, sg & 1u);
		}
	}
	__syncthreads();
	if (V3) dc_delta_codes(codec, nmb, s.dcval, reinterpret_cast<int *>(s.misc + 8));
	// code of block b's DC delta (chroma table for Cr/Cb, luma for Y1..Y4)
	auto dc_code = [&](int b) { return s.dctab[((b % 6) < 2 ? 0 : 512) + (s.dcval[b] & 0x1FF)]; };

	// ---- (1) first-fit quant scale search ------------------------------------------------
	// Largest bit total (blocks only) that still fits: 8 + 2*ceil((bits + 10)/16) <= max_size. A
	// pass whose running total passes it has failed (the reference's writer overflows at that
	// point too, mdec.c:323-325) and is abandoned early.
	const int limit_bits = max_size >= 8 ? 16 * ((max_size - 8) >> 1) - 10 : -1;
	int q = 1;
	uint32_t total_bits = 0;
	for (; q < 64; q++) {
		// next pass's reciprocals; readers only touch them after this pass's closing barrier
		if (q < 63)
			for (int i = tid; i < 64; i += T) s.qpar[((q + 1) & 1) * 64 + i] = c_qparam[(q + 1) * 64 + i];
		const uint2 *qpar = s.qpar + (q & 1) * 64;
		// the (usually busier) luma groups at the high end of the plane go first; drawing groups
		// from a shared ticket counter instead of this static round-robin measured no better
		uint32_t *total = &s.misc[q % 3];
		for (int g = ngroups - 1 - wid; g >= 0; g -= nw) {
			const int b = bs_plane_to_block(g * 32 + lane, cpad, nmb);
			const int prefix = live_prefix(s.rowq[g], q);
			int bits = 0;
			if (b >= 0) {
				bits = ac_bits(fc + (size_t)g * (BS_U4_PER_BLOCK * 32) + lane, qpar, c_qparam + q * 64, s.lenlut, prefix);
				bits += 2 + (V3 ? (int)(dc_code(b) >> 24) : 10);
				s.lens[b] = (uint16_t)bits;
			}
			const uint32_t sum = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)bits);
			uint32_t before = 0;
			if (lane == 0) before = atomicAdd(total, sum);
			before = __shfl_sync(0xFFFFFFFFu, before, 0);
			if ((int)(before + sum) > limit_bits) break;
		}
		__syncthreads();
		total_bits = *total;
		if (tid == 0) s.misc[(q + 2) % 3] = 0;
		// stream = blocks + 10-bit end-of-frame code; byte budget rule of flush_bits
		int units = (int)((total_bits + 10 + 15) >> 4);
		if (8 + 2 * units <= max_size) break;
	}

	uint32_t *out32 = reinterpret_cast<uint32_t *>(out + (size_t)f * out_stride);
	// destination of 32-bit word i of the frame's bitstream buffer: contiguous, or sliced into
	// 2016-byte sector payloads behind their 32-byte headers (mdec.c:831-832)
	auto word_at = [&](int i) -> uint32_t * {
		if (!str.sector_size) return out32 + i;
		int j = i / 504;
		return reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(out32) + (size_t)j * str.sector_size +
		                                    str.header_offset + 32) + (i - 504 * j);
	};
	// STR sector headers (mdec.c:782-820)
	auto write_str_headers = [&](uint32_t bytes_used, uint32_t bs0, uint32_t bs1) {
		const int chunks = max_size / 2016;
		for (int j = tid; j < chunks; j += T) {
			uint32_t *h = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(out32) + (size_t)j * str.sector_size +
			                                           str.header_offset);
			h[0] = 0x0160u | ((uint32_t)(str.video_id & 0xFFFF) << 16);
			h[1] = (uint32_t)j | ((uint32_t)chunks << 16);
			h[2] = (uint32_t)str_k;
			h[3] = bytes_used;
			h[4] = (uint32_t)(str.width & 0xFFFF) | ((uint32_t)(str.height & 0xFFFF) << 16);
			h[5] = bs0;
			h[6] = bs1;
			h[7] = 0;
		}
	};
	if (q >= 64) {
		for (int i = tid; i < (max_size >> 2); i += T) *word_at(i) = 0;
		if (!str.sector_size)
			for (int i = (max_size & ~3) + tid; i < max_size; i += T) out[(size_t)f * out_stride + i] = 0;
		else
			write_str_headers(0, 0, 0);
		if (tid == 0) results[f] = psxb200_bs_result_t{0, 0, 64, 0};
		return;
	}

	// ---- (2) exclusive scan of block bit lengths in bitstream order ---------------------
	for (int g = wid; g < nsgroups; g += nw) {
		uint32_t v = s.lens[g * 32 + lane], inc = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, d);
			if (lane >= d) inc += u;
		}
		s.lens[g * 32 + lane] = (uint16_t)(inc - v);
		if (lane == 31) s.gtot[g] = inc;
	}
	__syncthreads();
	if (wid == 0) {
		uint32_t carry = 0;
		for (int g0 = 0; g0 < nsgroups; g0 += 32) {
			int g = g0 + lane;
			uint32_t v = g < nsgroups ? s.gtot[g] : 0, inc = v;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, d);
				if (lane >= d) inc += u;
			}
			if (g < nsgroups) s.gtot[g] = carry + inc - v;
			carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
		}
	}
	__syncthreads();

	// ---- (3) emit --------------------------------------------------------------------------
	// Per block: a convergent pass re-quantises all 63 AC coefficients, parks min(level,63) in
	// the thread's shared-memory column and builds a 64-bit nonzero mask; the codes are then
	// produced by walking the set bits only (runs fall out of the bit positions).
	{
		const uint2 *qpar = s.qpar + (q & 1) * 64;
		uint8_t *lev = s.lev + tid;
		uint32_t nnz = 0;
		for (int g = ngroups - 1 - wid; g >= 0; g -= nw) {
			const int b = bs_plane_to_block(g * 32 + lane, cpad, nmb);
			if (b < 0) continue;
			const uint4 *gp = fc + (size_t)g * (BS_U4_PER_BLOCK * 32) + lane;
			uint32_t nz_lo, nz_hi;
			const uint32_t dc_mag = reinterpret_cast<const uint32_t *>(gp)[0] & 0xFFFFu;
			stage_levels(gp, qpar, c_qparam + q * 64, lev, lev_stride, live_prefix(s.rowq[g], q), nz_lo, nz_hi);
			uint4 sg = gp[8 * 32];
			BitWriter bw;
			bw.begin(stream, s.gtot[b >> 5] + s.lens[b]);
			if (V3) {
				uint32_t e = dc_code(b);
				bw.put((int)(e >> 24), e & 0xFFFFFFu);
			} else {
				bw.put(10, (uint32_t)quant_dc(dc_mag, sg.x & 1u) & 0x3FFu);
			}
			nnz += __popc(nz_lo) + __popc(nz_hi);
			int prev = 0;
#pragma unroll
			for (int half = 0; half < 2; half++) {
				uint32_t nz = half ? nz_hi : nz_lo;
				const uint32_t signs = half ? sg.y : sg.x;
				const uint8_t *lv = lev + 32 * half * lev_stride;
				while (nz) {
					int i = __ffs((int)nz) - 1;
					nz &= nz - 1;
					int pos = 32 * half + i;
					int run = pos - prev - 1;
					prev = pos;
					uint32_t m = lv[i * lev_stride];
					uint32_t neg = (signs >> i) & 1u;
					uint32_t e = s.vlc[(m << 6) | run];
					if (e & 0xFFFFFFu) {
						bw.put((int)(e >> 24), (e & 0xFFFFFFu) | neg);
					} else {
						// escape: exact level from the coefficient plane, clamped to [-512, 510]
						// (mdec.c:262-265), as 10-bit two's complement
						uint32_t word = reinterpret_cast<const uint32_t *>(gp + (pos >> 3) * 32)[(pos >> 1) & 3];
						uint32_t mag = (pos & 1) ? (word >> 16) : (word & 0xFFFFu);
						uint2 pq = qpar[pos];
						uint32_t lvl = __umulhi(mag + pq.y, pq.x);
						int level = neg ? -(int)min(lvl, 0x200u) : (int)min(lvl, 0x1FEu);
						bw.put(BS_AC_ESCAPE_BITS, (1u << 16) | ((uint32_t)run << 10) | ((uint32_t)level & 0x3FFu));
					}
				}
			}
			bw.put(2, 2u);   // end of block (mdec.c:502)
			if (b == nblk - 1) bw.put(10, V3 ? 0x3FFu : 0x1FFu);   // end of frame (mdec.c:645-652, 710)
			bw.finish();
		}
		nnz = warp_sum(nnz);
		if (lane == 0) atomicAdd(&s.misc[3], nnz);
	}
	__syncthreads();

	// ---- (4) header, results, copy-out -----------------------------------------------------
	int units = (int)((total_bits + 10 + 15) >> 4);
	int hwords = ((int)s.misc[3] + 2 * nblk + 2 + 0x3F) & ~0x3F;   // mdec.c:497,507,719,726
	int blocks_used = (hwords + 1) >> 1;
	if (tid == 0)
		results[f] = psxb200_bs_result_t{(8 + 2 * units + 3) & ~3, blocks_used, q, hwords};
	uint32_t hdr0 = (uint32_t)(blocks_used & 0xFFFF) | 0x38000000u;
	uint32_t hdr1 = (uint32_t)q | ((V3 ? 3u : 2u) << 16);
	for (int i = tid; i < (max_size >> 2); i += T) {
		uint32_t v;
		if (i == 0) v = hdr0;
		else if (i == 1) v = hdr1;
		else { uint32_t x = stream[i - 2]; v = (x >> 16) | (x << 16); }
		*word_at(i) = v;
	}
	if (str.sector_size) write_str_headers((uint32_t)((8 + 2 * units + 3) & ~3), hdr0, hdr1);
	if (!str.sector_size && tid < (max_size & 3)) {
		int i = (max_size & ~3) + tid;   // >= 8 here, since the frame fitted
		uint32_t x = stream[(i >> 2) - 2];
		uint32_t v = (x >> 16) | (x << 16);
		out[(size_t)f * out_stride + i] = (uint8_t)(v >> (8 * (i & 3)));
	}
}

// ---- launchers ---------------------------------------------------------------------------

size_t bs_pack_smem_bytes(bool v3, bool smem_stream, const BsGeometry &geo, int max_size_bound, int threads) {
	const int ngroups = geo.ngroups;
	size_t padded = (size_t)geo.nsgroups * 32;
	size_t n = 0;
	auto take = [&](size_t bytes) { n += (bytes + 15) & ~(size_t)15; };
	take(64 * 64);                        // lenlut
	take(4 * 64 * 64);                    // vlc
	take(4 * (8 + 4 * 32));               // misc
	take(8 * 2 * 64);                     // qpar
	if (smem_stream) take(4 * (size_t)((max_size_bound + 3) / 4 + 2));
	if (v3) take(4 * 1024);               // dctab
	take(4 * (size_t)(geo.nsgroups + 1)); // gtot
	take(8 * (size_t)ngroups);            // rowq
	take(2 * padded);                     // lens
	if (v3) take(2 * padded);             // dcval
	take(64 * (size_t)bs_lev_stride(threads));   // lev
	return n;
}

cudaError_t bs_launch_dct(int fdct_variant, const uint8_t *d_frames, size_t frame_bytes, int n, int width, int height,
                          const BsGeometry &geo, uint4 *d_coefs, cudaStream_t stream) {
	const unsigned per_frame = (unsigned)((geo.ngroups * 32 + BS_DCT_THREADS - 1) / BS_DCT_THREADS);
	// exact mb / mbh for mb < 2^16; 0 stands for mbh == 1, whose reciprocal does not fit
	const uint32_t mbh_magic = geo.mbh > 1 ? (uint32_t)(0x100000000ull / (unsigned)geo.mbh) + 1 : 0;
	for (int first = 0; first < n; first += 65535) {   // gridDim.y limit
		const int m = n - first < 65535 ? n - first : 65535;
		const dim3 grid(per_frame, (unsigned)m);
		const uint8_t *src = d_frames + (size_t)first * frame_bytes;
		uint4 *dst = d_coefs + (size_t)first * geo.frame_stride_u4;
		if (fdct_variant == FDCT_SSE2)
			bs_dct_kernel<FDCT_SSE2><<<grid, BS_DCT_THREADS, 0, stream>>>(src, frame_bytes, m, width, height, geo.mbh, mbh_magic,
			                                                              geo.nmb, geo.cgroups * 32, geo.ngroups, dst,
			                                                              geo.frame_stride_u4);
		else
			bs_dct_kernel<FDCT_ISLOW><<<grid, BS_DCT_THREADS, 0, stream>>>(src, frame_bytes, m, width, height, geo.mbh, mbh_magic,
			                                                               geo.nmb, geo.cgroups * 32, geo.ngroups, dst,
			                                                               geo.frame_stride_u4);
	}
	return cudaGetLastError();
}

template <bool V3, bool SMEM_STREAM, int MAX_THREADS, int MIN_CTAS>
static cudaError_t launch_pack_t(int threads, size_t smem, int n, const uint4 *d_coefs, const BsGeometry &geo, int codec,
                                 const int *d_max_sizes, int max_size_bound, uint8_t *d_out, size_t out_stride,
                                 psxb200_bs_result_t *d_results, uint32_t *d_gstream, size_t gstream_stride,
                                 const BsStrLayout &str, cudaStream_t stream) {
	auto kern = bs_pack_kernel<V3, SMEM_STREAM, MAX_THREADS, MIN_CTAS>;
	static size_t configured = 0;
	if (smem > configured) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return e;
		configured = smem;
	}
	kern<<<n, threads, smem, stream>>>(d_coefs, geo.frame_stride_u4, geo.nblk, geo.ngroups, geo.nsgroups,
	                                   geo.cgroups * 32, geo.nmb, codec,
	                                   d_max_sizes, max_size_bound, d_out, out_stride, d_results, d_gstream,
	                                   gstream_stride, str);
	return cudaGetLastError();
}

template <int MAX_THREADS, int MIN_CTAS>
static cudaError_t launch_pack_cfg(bool v3, bool smem_stream, int threads, size_t smem, int n, const uint4 *d_coefs,
                                   const BsGeometry &geo, int codec, const int *d_max_sizes, int max_size_bound,
                                   uint8_t *d_out, size_t out_stride, psxb200_bs_result_t *d_results,
                                   uint32_t *d_gstream, size_t gstream_stride, const BsStrLayout &str,
                                   cudaStream_t stream) {
#define PSXB200_PACK_ARGS threads, smem, n, d_coefs, geo, codec, d_max_sizes, max_size_bound, d_out, out_stride, \
	d_results, d_gstream, gstream_stride, str, stream
	if (v3)
		return smem_stream ? launch_pack_t<true, true, MAX_THREADS, MIN_CTAS>(PSXB200_PACK_ARGS)
		                   : launch_pack_t<true, false, MAX_THREADS, MIN_CTAS>(PSXB200_PACK_ARGS);
	return smem_stream ? launch_pack_t<false, true, MAX_THREADS, MIN_CTAS>(PSXB200_PACK_ARGS)
	                   : launch_pack_t<false, false, MAX_THREADS, MIN_CTAS>(PSXB200_PACK_ARGS);
#undef PSXB200_PACK_ARGS
}

cudaError_t bs_launch_pack(int codec, int threads, int min_ctas, int n, const uint4 *d_coefs, const BsGeometry &geo,
                           const int *d_max_sizes, int max_size_bound, uint8_t *d_out, size_t out_stride,
                           psxb200_bs_result_t *d_results, uint32_t *d_gstream, size_t gstream_stride,
                           const BsStrLayout &str, cudaStream_t stream) {
	bool v3 = codec != 0;
	bool smem_stream = d_gstream == nullptr;
	size_t smem = bs_pack_smem_bytes(v3, smem_stream, geo, max_size_bound, threads);
#define PSXB200_CFG_ARGS v3, smem_stream, threads, smem, n, d_coefs, geo, codec, d_max_sizes, max_size_bound, d_out, \
	out_stride, d_results, d_gstream, gstream_stride, str, stream
	// register budget variants: (max threads per CTA, CTAs per SM the register file must hold)
	if (threads <= 320) {
		if (min_ctas >= 4) return launch_pack_cfg<320, 4>(PSXB200_CFG_ARGS);
		if (min_ctas == 3) return launch_pack_cfg<320, 3>(PSXB200_CFG_ARGS);
		return launch_pack_cfg<320, 2>(PSXB200_CFG_ARGS);
	}
	if (min_ctas >= 2) return launch_pack_cfg<BS_PACK_MAX_THREADS, 2>(PSXB200_CFG_ARGS);
	return launch_pack_cfg<BS_PACK_MAX_THREADS, 1>(PSXB200_CFG_ARGS);
#undef PSXB200_CFG_ARGS
}

}  // namespace psxb200
