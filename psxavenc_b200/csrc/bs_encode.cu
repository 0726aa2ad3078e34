// MDEC "BS" v2/v3 frame encoder for sm_100a — the GPU side of encode_frame_bs
// (reference psxavenc/mdec.c:580-755).
//
// Data flow for a batch of n NV21 frames resident in HBM:
//
//   bs_dct_kernel      one thread per 8x8 block: gathers the block from the NV21 frame
//                      (mdec.c:605-634), level-shifts, runs the bit-exact integer FDCT in
//                      registers (mdec.c:640) and reduces every AC coefficient to
//                      y = floor(2|c| / quant[i]), from which the level at ANY quant scale q is
//                      (y + q) / (2q) exactly (see below). Only the coefficients with y >= 1 —
//                      the ones that can be nonzero at some q — are kept, as a list of 16-bit
//                      (y, zig-zag position) entries per block, plus a 64-bit sign mask and the
//                      DC term. Lists are laid out so that the 32 lanes of a warp (32 blocks of
//                      one kind) read and write 512 contiguous bytes per uint4 access.
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
// Quantisation: DIVIDE_ROUNDED (mdec.c:438) is round-half-away-from-zero, i.e. for d = quant*q
//   |level| = floor((2|c| + d) / 2d) = floor((floor(2|c| / quant) + q) / 2q)
// (nested floor divisions; quant*q / quant = q is an integer). The inner division is by a
// compile-time constant and independent of q, so it is done once in the FDCT kernel; the outer
// one has the same divisor for all 63 AC coefficients of a pass, and a coefficient is nonzero at
// q iff y >= q. For 8-bit input |c| <= 8192 and quant >= 16, so y < 1024 (10 bits).
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

#include <cstdlib>
#include <mutex>

#include "bs_encode.h"
#include "bs_tables.h"
#include "edc.cuh"
#include "fdct.cuh"

namespace psxb200 {

// ---- constant tables -------------------------------------------------------------------
// [q] -> reciprocals of the pass's common divisor: .x = floor(2^32 / 2q) + 1 for t = y + q,
// .y = floor(2^32 / 128q) + 1 for t = 64 (y + q) + (any 6 low bits); level = umulhi(t, magic).
__constant__ uint2 c_qmagic[64];
// One zero guard byte, then [min(level,63)][run] -> code length in bits incl. sign (22 = escape),
// 0 for level 0 (the list walks index it with run + 1, see price_entry); padded to whole words.
// (Global, not constant memory: every CTA copies the tables to shared memory with one word per
// thread, and constant-bank reads at 32 different addresses per warp are served one by one.)
constexpr int LENLUT1_BYTES = 4 + 64 * 64;
__device__ __align__(16) uint8_t g_lenlut1[LENLUT1_BYTES];
// [min(level,41)][min(run,32)] -> (len << 24) | code with the sign bit (LSB) clear; level 0 -> 0;
// escape (including all of row 41 and column 32) -> (22 << 24) with code 0. Copied to shared
// memory by every CTA.
__device__ uint32_t g_vlc[BS_VLC_ROWS * BS_VLC_COLS];
// v3 DC delta codes: [0] chroma, [1] luma.
__device__ uint32_t g_dcvlc[2 * 512];

__host__ __device__ constexpr int zigzag_at(int i) {
	constexpr int t[64] = {BS_ZIGZAG_LIST};
	return t[i];
}

__host__ __device__ constexpr int quant_zz_at(int i) {
	constexpr int t[64] = {BS_QUANT_ZZ_LIST};
	return t[i];
}

// umulhi(|c|, ymagic_at(i)) == floor(2|c| / quant of zig-zag position i), exact for |c| < 2^15.
__host__ __device__ constexpr uint32_t ymagic_at(int i) {
	return (uint32_t)(0x200000000ull / (unsigned long long)quant_zz_at(i)) + 1u;
}

// EDC byte table + advance tables (edc.cuh), see EDC_TABLE_WORDS
__device__ uint32_t g_edc[EDC_TABLE_WORDS];

constexpr int MAX_DEVICES = 64;
static std::mutex g_device_lock;          // guards the per-device one-time setup below
static bool g_tables_uploaded[MAX_DEVICES];

cudaError_t bs_upload_tables() {
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (dev < 0 || dev >= MAX_DEVICES) return cudaErrorInvalidDevice;
	std::lock_guard<std::mutex> guard(g_device_lock);
	if (g_tables_uploaded[dev]) return cudaSuccess;

	static uint2 qmagic[64];
	static uint8_t lenlut1[LENLUT1_BYTES];
	static uint32_t vlc[BS_VLC_ROWS * BS_VLC_COLS];
	static uint32_t dcvlc[2 * 512];
	static uint32_t edc[EDC_TABLE_WORDS];
	for (int q = 0; q < 64; q++) {
		uint32_t d = 2u * (q ? q : 1);
		qmagic[q].x = (uint32_t)(0x100000000ull / d) + 1;
		qmagic[q].y = (uint32_t)(0x100000000ull / (64ull * d)) + 1;
	}
	for (int lv = 0; lv < 64; lv++) {
		for (int run = 0; run < 64; run++) {
			int len = 0;
			if (lv > 0) {
				uint32_t e = (run < BS_AC_RUNS && lv < BS_AC_LEVELS) ? BS_AC_VLC[run * BS_AC_LEVELS + lv] : 0;
				len = e ? (int)(e >> 24) : BS_AC_ESCAPE_BITS;
			}
			lenlut1[1 + ((lv << 6) | run)] = (uint8_t)len;
			uint32_t e = (lv > 0 && run < BS_AC_RUNS && lv < BS_AC_LEVELS) ? BS_AC_VLC[run * BS_AC_LEVELS + lv] : 0;
			if (lv < BS_VLC_ROWS && run < BS_VLC_COLS)
				vlc[lv * BS_VLC_COLS + run] = lv == 0 ? 0u : (e ? e : (uint32_t)BS_AC_ESCAPE_BITS << 24);
		}
	}
	for (int i = 0; i < 512; i++) {
		dcvlc[i] = BS_DC_VLC_CHROMA[i];
		dcvlc[512 + i] = BS_DC_VLC_LUMA[i];
	}
	for (uint32_t i = 0; i < 256; i++) edc[i] = edc_byte_table_entry(i);
	edc_build_advance_table(edc, EDC_PIECE_FORM1, edc + 256);
	edc_build_advance_table(edc, EDC_PIECE_FORM2, edc + 1280);
	// Synchronous copies on the legacy stream, bracketed by device synchronisation: they happen
	// once per device, before any kernel of this library has been launched on it.
	if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;
	if ((e = cudaMemcpyToSymbol(c_qmagic, qmagic, sizeof(qmagic))) != cudaSuccess) return e;
	if ((e = cudaMemcpyToSymbol(g_lenlut1, lenlut1, sizeof(lenlut1))) != cudaSuccess) return e;
	if ((e = cudaMemcpyToSymbol(g_vlc, vlc, sizeof(vlc))) != cudaSuccess) return e;
	if ((e = cudaMemcpyToSymbol(g_dcvlc, dcvlc, sizeof(dcvlc))) != cudaSuccess) return e;
	if ((e = cudaMemcpyToSymbol(g_edc, edc, sizeof(edc))) != cudaSuccess) return e;
	if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;
	g_tables_uploaded[dev] = true;
	return cudaSuccess;
}

const uint32_t *edc_tables_device() {
	void *p = nullptr;
	if (bs_upload_tables() != cudaSuccess || cudaGetSymbolAddress(&p, g_edc) != cudaSuccess) return nullptr;
	return static_cast<const uint32_t *>(p);
}

// ---- kernel 1: gather + FDCT -----------------------------------------------------------

// byte k of w, zero extended: one PRMT (a shift and a mask measured 1.6 % slower over the kernel)
__device__ __forceinline__ int byte_of(uint32_t w, int k) { return (int)__byte_perm(w, 0u, 0x4440u + (uint32_t)k); }

template <int VARIANT>
__global__ void __launch_bounds__(BS_DCT_THREADS, BS_DCT_MIN_CTAS)
bs_dct_kernel(const uint8_t *__restrict__ frames, size_t frame_bytes, int n_frames, int width, int height,
              int mbh, uint32_t mbh_magic, int nmb, int cpad, int ngroups, uint4 *__restrict__ coefs,
              size_t frame_stride_u4) {
	// Per-thread list columns: entry k of thread t at s_list[k * BS_DCT_THREADS + t] (a warp's
	// appends fall into distinct banks whatever the lanes' fill levels). Zeroed first so that the
	// tail of every list reads as "no coefficient" (masking the tail on read-back instead, without
	// this block-wide zeroing and its barrier, measured 3 % slower).
	__shared__ __align__(16) uint16_t s_list[64 * BS_DCT_THREADS];
	// A pack kernel launched behind this one with programmatic stream serialisation (cluster mode
	// with PSXB200_PDL=1) may start its set-up now; it waits for this grid's completion before it
	// touches the plane. Without such a launch behind it this does nothing.
	asm volatile("griddepcontrol.launch_dependents;");
#pragma unroll
	for (int j = 0; j < (64 * 2) / 16; j++)
		reinterpret_cast<uint4 *>(s_list)[threadIdx.x + BS_DCT_THREADS * j] = make_uint4(0, 0, 0, 0);
	__syncthreads();

	// grid: x = chunk of 128 plane lanes within the frame, y = frame
	const int f = blockIdx.y;
	const int b = blockIdx.x * BS_DCT_THREADS + threadIdx.x;   // plane index (type-major, see BsGeometry)
	if (b >= ngroups * 32) return;                             // whole warps: ngroups * 32 lanes per frame
	// whole warps map to one group of 32 blocks of one kind; padding lanes idle but stay for
	// the warp reductions below
	const bool chroma = b < cpad;   // warp-uniform: cpad is a multiple of 32
	const bool active = chroma ? b < 2 * nmb : b - cpad < 4 * nmb;

	uint4 *dst = coefs + (size_t)f * frame_stride_u4 + (size_t)(b >> 5) * (BS_U4_PER_BLOCK * 32) + (b & 31);

	int v[64];   // samples, then coefficients, then their y values (raster order)
	if (!active) {
#pragma unroll
		for (int i = 0; i < 64; i++) v[i] = 0;
	} else {
		// macroblocks in bitstream order: columns outermost, rows next (mdec.c:689-704)
		int mb = chroma ? b >> 1 : (b - cpad) >> 2;
		int k = chroma ? b & 1 : 2 + ((b - cpad) & 3);
		int mx = mbh_magic ? (int)__umulhi((uint32_t)mb, mbh_magic) : mb, my = mb - mx * mbh;   // mb / mbh (magic 0: mbh == 1)
		const uint8_t *fr = frames + (size_t)f * frame_bytes;

		if (chroma) {
			// interleaved CrCb plane: Cr at even bytes, Cb at odd (mdec.c:627-628)
			const uint8_t *p = fr + (size_t)width * height + (size_t)width * (my * 8) + mx * 16;
			// byte k / k + 2 of each word, zero extended: one PRMT per sample, selectors made once
			const uint32_t sel_lo = 0x4440u + (uint32_t)k, sel_hi = 0x4442u + (uint32_t)k;
#pragma unroll
			for (int y = 0; y < 8; y++) {
				uint4 r = __ldg(reinterpret_cast<const uint4 *>(p + (size_t)y * width));
				uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
				for (int x = 0; x < 8; x++) {
					v[8 * y + x] = (int)__byte_perm(w[x >> 1], 0u, (x & 1) ? sel_hi : sel_lo);
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
	}

	// Coefficients are visited in descending zig-zag order: each sign is shifted into its mask
	// with one funnel shift (coefficient i ends up at bit i of its half), and the list comes out
	// with the highest position first — the pack kernel walks it backwards. The y values replace
	// the coefficients in v[] for the dense format below.
	uint32_t sign_lo = 0, sign_hi = 0, dc_mag = 0;
	uint16_t *const col = s_list + threadIdx.x;
	// the append cursor as a 32-bit shared-space address: one add per bump, no generic pointers
	const uint32_t col_addr = (uint32_t)__cvta_generic_to_shared(col);
	uint32_t tail = col_addr;
#pragma unroll
	for (int i = 63; i >= 0; i--) {
		const int c = v[zigzag_at(i)];
		if (i < 32) sign_lo = __funnelshift_l((uint32_t)c, sign_lo, 1);
		else sign_hi = __funnelshift_l((uint32_t)c, sign_hi, 1);
		const uint32_t mag = (uint32_t)abs(c);
		if (i == 0) {
			dc_mag = mag;   // the DC term has its own fixed step (mdec.c:447, 671)
			v[0] = 0;
		} else {
			const uint32_t y = __umulhi(mag, ymagic_at(i));   // floor(2|c| / quant[i])
			v[zigzag_at(i)] = (int)y;
			if (y) {
				asm volatile("st.shared.u16 [%0], %1;" ::"r"(tail), "h"((uint16_t)((y << 6) | (uint32_t)i)) : "memory");
				tail += 2 * BS_DCT_THREADS;
			}
		}
	}

	const int count = (int)(tail - col_addr) / (2 * BS_DCT_THREADS);
	const int longest = (int)__reduce_max_sync(0xFFFFFFFFu, (uint32_t)count);
	// Lists pay off while the group's longest one stays well below the 63 possible entries. A
	// group whose longest list would fill (nearly) all eight rows anyway is stored dense instead:
	// y of all 64 positions in zig-zag order (position implicit, DC slot 0), which the pack
	// kernel handles with straight-line code that needs no positions (cheaper per coefficient).
	const bool dense = longest >= BS_DENSE_MIN;
	if (dense) {
#pragma unroll
		for (int j = 0; j < 8; j++) {
			uint32_t w[4];
#pragma unroll
			for (int t = 0; t < 4; t++)
				w[t] = (uint32_t)v[zigzag_at(8 * j + 2 * t)] | ((uint32_t)v[zigzag_at(8 * j + 2 * t + 1)] << 16);
			dst[j * 32] = make_uint4(w[0], w[1], w[2], w[3]);
		}
	} else {
		// rows of 8 entries per lane, as many as the longest list needs; shorter lists are zero
		// padded (y = 0: never a coefficient)
		const int nrows = (longest + 7) >> 3;
		for (int r = 0; r < nrows; r++) {
			const uint16_t *e = col + 8 * r * BS_DCT_THREADS;
			uint32_t w[4];
#pragma unroll
			for (int t = 0; t < 4; t++)
				w[t] = (uint32_t)e[(2 * t) * BS_DCT_THREADS] | ((uint32_t)e[(2 * t + 1) * BS_DCT_THREADS] << 16);
			dst[r * 32] = make_uint4(w[0], w[1], w[2], w[3]);
		}
	}
	// meta row: signs by zig-zag position, |DC| and this block's list length, the group's longest
	// list with BS_DENSE_FLAG when the rows are dense
	dst[BS_META_ROW * 32] = make_uint4(sign_lo, sign_hi, dc_mag | ((uint32_t)count << 16),
	                                   (uint32_t)longest | (dense ? BS_DENSE_FLAG : 0u));
}

// ---- kernel 2: quant-scale search + bit packing ------------------------------------------

struct PackSmem {
	uint32_t *stream;   // bitstream image, 32-bit words, first stream bit = bit 31 of word 0
	uint32_t *dctab;    // v3: DC delta codes, [0..511] chroma, [512..1023] luma (len<<24 | code)
	uint32_t *gtot;     // per group bit totals -> exclusive group bases
	uint8_t *grows;     // per plane group: list rows in use (longest list of its 32 blocks), or 0x80: dense rows
	uint32_t *misc;     // [0..2] rotating frame totals, [3] nonzero AC count, [8..] scan scratch
	uint32_t *vlc;      // [min(level,41)][min(run,32)] -> (len<<24)|code, see g_vlc
	uint16_t *lens;     // per block bit length at the current q; after the scan: exclusive offset in its group
	int16_t *dcval;     // v3: per block quantised DC, replaced in place by its coded delta
	uint8_t *lenlut;    // [min(level,63)][run] -> code length; preceded by a zero guard byte (see price_entry)
	uint32_t lut1;      // shared-space address of that guard byte (lenlut - 1), a multiple of 64
	uint32_t *stage;    // emit: up to four list rows of the thread's current block, word j at stage[j * T + tid]
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

// ---- pricing: the pricing half of encode_dct_block (mdec.c:482-499) over a block's list ---
//
// List entries hold (y << 6) | position, two per 32-bit word; a block's list is stored highest
// position first, so it is walked from its last entry down: zero padding first, then the
// coefficients in ascending zig-zag order. An entry is a coefficient at quant scale q iff its
// level (y + q) / 2q is nonzero; its run is the distance to the previous such entry.

struct QuantScale {
	uint32_t q, q64;    // q and q << 6
	uint32_t m_hi;      // floor(2^32 / 2q) + 1: level of t = y + q
	uint32_t m_lo;      // floor(2^32 / 128q) + 1: level of t = 64 (y + q) + junk < 64
	uint32_t q22;       // q << 22: a word's upper entry is a coefficient iff word >= q22
	__device__ __forceinline__ explicit QuantScale(int qs)
		: q((uint32_t)qs), q64((uint32_t)qs << 6), m_hi(c_qmagic[qs].x), m_lo(c_qmagic[qs].y), q22((uint32_t)qs << 22) {}
};

template <bool UPPER>
__device__ __forceinline__ uint32_t entry_level(uint32_t word, const QuantScale &k) {
	return UPPER ? __umulhi((word >> 22) + k.q, k.m_hi) : __umulhi((word & 0xFFFFu) + k.q64, k.m_lo);
}

template <bool UPPER>
__device__ __forceinline__ uint32_t entry_pos(uint32_t word) {
	return UPPER ? (word >> 16) & 63u : word & 63u;
}

__device__ __forceinline__ uint32_t lds_u8(uint32_t shared_addr) {
	uint32_t v;
	asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(shared_addr));
	return v;
}

// lut1 = the shared-space address of the byte in front of the length table: run = pos - prev - 1,
// and a padding entry (level 0, position 0, met only while prev is still 0) reads that zero guard
// byte. The table's address goes into the same three-input add as the run (the compiler, left with
// a pointer, spent one more add per entry on it).
template <bool UPPER>
__device__ __forceinline__ void price_entry(uint32_t word, const QuantScale &k, uint32_t lut1, uint32_t &bits,
                                            uint32_t &prev) {
	const uint32_t lv = entry_level<UPPER>(word, k);
	const uint32_t pos = entry_pos<UPPER>(word);
	bits += lds_u8(imad(min(lv, 63u), 64u, pos - prev + lut1));
	// prev = lv ? pos : prev as a predicated move (the select the compiler makes of it sits on the
	// ALU pipe, which bounds this loop; measured -0.7 % of the kernel)
	asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p mov.u32 %0, %2;\n\t}" : "+r"(prev) : "r"(lv), "r"(pos));
}

// AC bit cost of the block whose list occupies rows 0..nrows-1 of (group, lane).
__device__ __forceinline__ int ac_bits(const uint4 *__restrict__ gp, int nrows, const QuantScale &k, uint32_t lut1) {
	uint32_t bits = 0, prev = 0;
	uint4 next = gp[(nrows > 0 ? nrows - 1 : 0) * 32];   // the row after this one is in flight while this one is priced
	for (int r = nrows - 1; r >= 0; r--) {
		const uint4 w = next;
		next = gp[(r > 0 ? r - 1 : 0) * 32];
		price_entry<true>(w.w, k, lut1, bits, prev);
		price_entry<false>(w.w, k, lut1, bits, prev);
		price_entry<true>(w.z, k, lut1, bits, prev);
		price_entry<false>(w.z, k, lut1, bits, prev);
		price_entry<true>(w.y, k, lut1, bits, prev);
		price_entry<false>(w.y, k, lut1, bits, prev);
		price_entry<true>(w.x, k, lut1, bits, prev);
		price_entry<false>(w.x, k, lut1, bits, prev);
	}
	return (int)bits;
}

// The first pass (q = 1): every list entry is a coefficient (y >= 1), so its level is (y + 1) >> 1
// without a division and its run is the distance to the entry walked before it — padding entries
// (all zero, walked first) cost the guard byte and leave the position at 0.
// lut1x2 = 64 + twice the table's address, lut1max = the address of its last row: with the address a
// multiple of 64 it rides through the shift and the mask, ((e + 64 + 2 lut1) >> 1) & ~63 =
// 64 * level + lut1.
template <bool UPPER>
__device__ __forceinline__ void price_entry_q1(uint32_t word, uint32_t lut1x2, uint32_t lut1max, uint32_t &bits, uint32_t &prev) {
	const uint32_t e = UPPER ? word >> 16 : word & 0xFFFFu;
	const uint32_t pos = e & 63u;
	const uint32_t row = min(((e + lut1x2) >> 1) & ~63u, lut1max);   // lut1 + 64 * min(level, 63)
	bits += lds_u8(row + pos - prev);
	prev = pos;
}

__device__ __forceinline__ int ac_bits_q1(const uint4 *__restrict__ gp, int nrows, uint32_t lut1) {
	uint32_t bits = 0, prev = 0;
	const uint32_t lut1x2 = 64u + 2u * lut1, lut1max = lut1 + (63u << 6);
	uint4 next = gp[(nrows > 0 ? nrows - 1 : 0) * 32];
	for (int r = nrows - 1; r >= 0; r--) {
		const uint4 w = next;
		next = gp[(r > 0 ? r - 1 : 0) * 32];
		price_entry_q1<true>(w.w, lut1x2, lut1max, bits, prev);
		price_entry_q1<false>(w.w, lut1x2, lut1max, bits, prev);
		price_entry_q1<true>(w.z, lut1x2, lut1max, bits, prev);
		price_entry_q1<false>(w.z, lut1x2, lut1max, bits, prev);
		price_entry_q1<true>(w.y, lut1x2, lut1max, bits, prev);
		price_entry_q1<false>(w.y, lut1x2, lut1max, bits, prev);
		price_entry_q1<true>(w.x, lut1x2, lut1max, bits, prev);
		price_entry_q1<false>(w.x, lut1x2, lut1max, bits, prev);
	}
	return (int)bits;
}

// The same for a dense group (all 64 y values in zig-zag order, position implicit), one half
// (four rows, 32 positions) at a time to bound the register footprint.
template <int HALF>
__device__ __forceinline__ void price_dense_half(const uint4 *__restrict__ gp, const QuantScale &k, const uint8_t *lenlut,
                                                 uint32_t &bits, uint32_t &run) {
	uint32_t w[16];
#pragma unroll
	for (int j = 0; j < 4; j++) {
		const uint4 r = gp[(4 * HALF + j) * 32];
		w[4 * j + 0] = r.x; w[4 * j + 1] = r.y; w[4 * j + 2] = r.z; w[4 * j + 3] = r.w;
	}
#pragma unroll
	for (int i = (HALF ? 0 : 1); i < 32; i++) {
		const uint32_t y = (i & 1) ? w[i >> 1] >> 16 : w[i >> 1] & 0xFFFFu;
		const uint32_t lv = __umulhi(y + k.q, k.m_hi);
		bits = imad(lenlut[imad(min(lv, 63u), 64u, run)], 1u, bits);
		const uint32_t z = imad(lv, 1u, 0xFFFFFFFFu) >> 31;   // 1 when the level is zero
		run = imad(run, z, z);                                 // (run + 1) * z
	}
}

__device__ __forceinline__ int ac_bits_dense(const uint4 *__restrict__ gp, const QuantScale &k, const uint8_t *lenlut) {
	uint32_t bits = 0, run = 0;
	price_dense_half<0>(gp, k, lenlut, bits, run);
	price_dense_half<1>(gp, k, lenlut, bits, run);
	return (int)bits;
}

// Emit, convergent part for a dense group: parks rows 4*half..4*half+3 in the thread's column and
// returns the mask of positions 32*half + e that are coefficients at this quant scale (bit e).
template <bool PRED>
__device__ __forceinline__ uint32_t stage_dense(const uint4 *__restrict__ gp, int half, const QuantScale &k,
                                                uint32_t *stage, int stride) {
	uint32_t live = 0;
	const uint32_t q16 = k.q << 16;
#pragma unroll
	for (int j = 0; j < 4; j++) {
		const uint4 r = gp[(4 * half + j) * 32];
		const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
		for (int t = 0; t < 4; t++) {
			stage[(4 * j + t) * stride] = w[t];
			if (PRED) {
				// compare + predicated add as in stage_rows — in the kernel for busy frames only: in
				// the common kernel, where dense groups are rare, it cost the list path 1 % (measured)
				asm("{\n\t.reg .pred p;\n\tsetp.ge.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, 1, %0;\n\t}"
				    : "+r"(live) : "r"(w[t] & 0xFFFFu), "r"(k.q), "r"(1u << (8 * j + 2 * t)));
				asm("{\n\t.reg .pred p;\n\tsetp.ge.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, 1, %0;\n\t}"
				    : "+r"(live) : "r"(w[t]), "r"(q16), "r"(1u << (8 * j + 2 * t + 1)));
			} else {
				live |= ((w[t] & 0xFFFFu) >= k.q ? 1u : 0u) << (8 * j + 2 * t);
				live |= (w[t] >= q16 ? 1u : 0u) << (8 * j + 2 * t + 1);
			}
		}
	}
	return live;
}

// Emit, convergent part: parks up to four list rows (rows r0..r0+n-1) in the thread's column
// of the staging area (word j of the thread at stage[j * stride]) and returns the mask of
// entries that are coefficients at this quant scale: bit 31 - e for local entry e = 8 * row + k,
// so that walking the set bits upwards visits the coefficients in ascending zig-zag order. (Bit e
// and a walk from the top saves the bit reversal per step but measured 1.4 % slower: clearing the
// bit then waits for the find.)
__device__ __forceinline__ uint32_t stage_rows(const uint4 *__restrict__ gp, int r0, int n, const QuantScale &k,
                                               uint32_t *stage, int stride) {
	uint32_t live = 0;
	for (int j = n - 1; j >= 0; j--) {
		const uint4 w = gp[(r0 + j) * 32];
		stage[(4 * j + 0) * stride] = w.x;
		stage[(4 * j + 1) * stride] = w.y;
		stage[(4 * j + 2) * stride] = w.z;
		stage[(4 * j + 3) * stride] = w.w;
		uint32_t row = 0;   // bit 7 - e for entry e of this row
		// row += bit where the entry is a coefficient, as a compare and a predicated IMAD (FMA pipe)
		// instead of compare, select and add on the ALU pipe, which bounds the kernel (-1.9 % of it)
#define PSXB200_ROWBIT(value, bound, bit) \
		asm("{\n\t.reg .pred p;\n\tsetp.ge.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, 1, %0;\n\t}" : "+r"(row) : "r"(value), "r"(bound), "r"(bit))
		PSXB200_ROWBIT(w.w, k.q22, 0x01u);
		PSXB200_ROWBIT(w.w & 0xFFFFu, k.q64, 0x02u);
		PSXB200_ROWBIT(w.z, k.q22, 0x04u);
		PSXB200_ROWBIT(w.z & 0xFFFFu, k.q64, 0x08u);
		PSXB200_ROWBIT(w.y, k.q22, 0x10u);
		PSXB200_ROWBIT(w.y & 0xFFFFu, k.q64, 0x20u);
		PSXB200_ROWBIT(w.x, k.q22, 0x40u);
		PSXB200_ROWBIT(w.x & 0xFFFFu, k.q64, 0x80u);
#undef PSXB200_ROWBIT
		live |= row << (24 - 8 * j);
	}
	return live;
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

// ---- census: skipping quant scales that provably cannot fit ------------------------------
//
// The reference takes the first quant scale whose stream fits, trying q = 1, 2, ... in turn
// (mdec.c:663-722); skipping a q is only allowed when it is certain to fail. Every coefficient
// that is nonzero at q costs at least the run-0 code of its level (the shortest code of a level,
// and lengths grow with the level), so  fixed + sum over entries of len(level(y, q), run 0)  is a
// lower bound of the frame's bit total at q; capping y at 15 only lowers it further. One walk over
// the lists builds a histogram of min(y, 15): each thread adds, per PAIR of entries (one 32-bit
// word of a row), a 128-bit word from a 256-entry table that holds the two ones in the bytes of
// the entries' bins (a block has at most 64 entries, so the byte counters cannot overflow within a
// block), and unpacks the bytes into its 15 counters once per block. (One table read per entry from
// a 16-entry table was bound by the shared-memory bandwidth of the 128-bit reads.) From the
// CTA-wide sums the bound follows for all q at once.
// Called by all threads of the CTA; returns the smallest q >= 2 the bound cannot exclude (64: none).
constexpr int CENSUS_BINS = 15;

__device__ __noinline__ int census_first_candidate(const uint4 *__restrict__ fc, int ngroups, int cpad, int nmb,
                                                   const uint8_t *grows, const uint8_t *lenlut,
                                                   uint32_t *scratch /* >= 64 words */, uint4 *pairtab /* 256 entries */,
                                                   int fixed_bits, int limit_bits) {
	const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = T >> 5;
	// [min(y_lo, 15) | min(y_hi, 15) << 4] -> a one in byte (bin - 1) for each of the two, none for y = 0
	for (int i = tid; i < 256; i += T) {
		uint32_t w[4] = {0, 0, 0, 0};
		const int a = i & 15, b = i >> 4;
		if (a) w[(a - 1) >> 2] += 1u << (8 * ((a - 1) & 3));
		if (b) w[(b - 1) >> 2] += 1u << (8 * ((b - 1) & 3));
		pairtab[i] = make_uint4(w[0], w[1], w[2], w[3]);
	}
	if (tid < 64) scratch[tid] = tid == CENSUS_BINS ? 64u : 0u;   // [0..14] bin totals, [15] the answer
	__syncthreads();
	uint32_t bins[CENSUS_BINS];
#pragma unroll
	for (int b = 0; b < CENSUS_BINS; b++) bins[b] = 0;
	for (int g = ngroups - 1 - wid; g >= 0; g -= nw) {
		if (bs_plane_to_block(g * 32 + lane, cpad, nmb) < 0) continue;
		const uint4 *gp = fc + (size_t)g * (BS_U4_PER_BLOCK * 32) + lane;
		const int rows = grows[g];
		const bool dense = rows & 0x80;
		const int nrows = dense ? 8 : rows;
		uint4 acc = make_uint4(0, 0, 0, 0);
		for (int r = 0; r < nrows; r++) {
			const uint4 w = gp[r * 32];
			const uint32_t v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
			for (int t = 0; t < 4; t++) {
				// dense rows hold bare y values, list rows (y << 6) | position: both y of the word,
				// capped at 15 in one packed minimum, then merged into lo | hi << 4
				const uint32_t yy = dense ? v[t] : (v[t] >> 6) & 0x03FF03FFu;
				const uint32_t m = __vminu2(yy, 0x000F000Fu);
				const uint4 two = pairtab[(m | (m >> 12)) & 0xFFu];
				acc.x += two.x; acc.y += two.y; acc.z += two.z; acc.w += two.w;
			}
		}
		const uint32_t a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
		for (int b = 0; b < CENSUS_BINS; b++) bins[b] += (a[b >> 2] >> (8 * (b & 3))) & 0xFFu;
	}
#pragma unroll
	for (int b = 0; b < CENSUS_BINS; b++) {
		const uint32_t sum = warp_sum(bins[b]);
		if (lane == 0 && sum) atomicAdd(&scratch[b], sum);
	}
	__syncthreads();
	if (tid >= 2 && tid < 64) {
		const uint32_t q = (uint32_t)tid;
		long long bound = fixed_bits;
		for (int b = 0; b < CENSUS_BINS; b++) {
			const uint32_t level = ((uint32_t)(b + 1) + q) / (2 * q);
			if (level) bound += (long long)scratch[b] * lenlut[min(level, 63u) << 6];
		}
		if (bound <= (long long)limit_bits) atomicMin(&scratch[CENSUS_BINS], q);
	}
	__syncthreads();
	const int first = (int)scratch[CENSUS_BINS];
	__syncthreads();
	return first;
}

// STR mode: where frame f of the launch sits in its file (see BsStrLayout)
struct StrFrame {
	long long k;        // frame_index (1-based)
	long long before;   // video sectors of the file before this frame
	int file;
};
__device__ __forceinline__ StrFrame str_frame(const BsStrLayout &str, int f) {
	const int g = str.frame_base + f;
	StrFrame sf;
	sf.file = str.frames_per_file > 0 ? g / str.frames_per_file : 0;
	sf.k = (long long)str.frame_index0 + (g - sf.file * str.frames_per_file);
	sf.before = (sf.k - 1) * str.sectors_num / str.sectors_den;
	return sf;
}

// BUSY = false: the kernel every frame goes through. A frame whose first pass overruns its budget
// before half of the frame has been priced is not finished here: its result row is marked
// (quant_scale 0) and the CTA exits. BUSY = true: launched right behind it on the same stream,
// takes only the marked frames, rules out the hopeless quant scales with the census and then
// carries on with the first-fit search. (Two kernels rather than one branch: the extra code in the
// common kernel cost it a third of its speed, measured.)
//
// CL > 1: a thread-block CLUSTER of CL CTAs (one per SM) shares a frame — for calls with only a
// few frames (the drop-in symbols encode one at a time), where the latency of a frame is what
// counts and a single SM needs 31 us for its ~100 k warp instructions. The CTAs of the cluster
// split the plane groups between them in the search and in the emit phase; every CTA keeps the
// whole frame's block lengths (each length is stored into all CL shared memories through
// distributed shared memory) and its own image of the bitstream holding only its blocks' codes;
// pass totals and the coefficient count are summed over the cluster after a cluster barrier,
// and the copy-out ORs the CL images together, each CTA writing a share of the words.
template <bool V3, bool SMEM_STREAM, bool STR, bool BUSY, int MAX_THREADS, int MIN_CTAS, int CL = 1>
__global__ void __launch_bounds__(MAX_THREADS, MIN_CTAS)
bs_pack_kernel(const uint4 *__restrict__ coefs, size_t frame_stride_u4, int nblk, int ngroups, int nsgroups, int cpad,
               int nmb, int codec,
               const int *__restrict__ max_sizes, int max_size_bound, uint8_t *__restrict__ out, size_t out_stride,
               psxb200_bs_result_t *__restrict__ results, uint32_t *__restrict__ gstream, size_t gstream_stride,
               const BsStrLayout str) {
	extern __shared__ __align__(128) uint8_t smem_raw[];
	const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = T >> 5;
	static_assert(CL == 1 || (SMEM_STREAM && !BUSY), "cluster mode: shared-memory image, common kernel only");
	namespace cg = cooperative_groups;
	const int f = CL > 1 ? (int)blockIdx.x / CL : (int)blockIdx.x;
	const int rank = CL > 1 ? (int)cg::this_cluster().block_rank() : 0;   // == blockIdx.x % CL
	if (BUSY && results[f].quant_scale != 0) return;   // finished by the first kernel
	const int padded = nsgroups * 32;   // blocks in bitstream order, rounded up to whole scan groups
	const int stream_words = (max_size_bound + 3) / 4 + 2;

	// fixed-size tables first so that their shared-memory addresses are compile-time offsets
	PackSmem s;
	{
		uint8_t *p = smem_raw;
		auto take = [&](size_t bytes) { uint8_t *at = p; p += (bytes + 15) & ~(size_t)15; return at; };
		s.lenlut = take(LENLUT1_BYTES) + 1;
		s.lut1 = (uint32_t)__cvta_generic_to_shared(s.lenlut - 1);
		if (s.lut1 & 63u) __trap();   // the q = 1 walk relies on it (ac_bits_q1)
		s.vlc = reinterpret_cast<uint32_t *>(take(4 * BS_VLC_ROWS * BS_VLC_COLS));
		s.misc = reinterpret_cast<uint32_t *>(take(4 * (8 + 4 * 32)));
		s.stream = reinterpret_cast<uint32_t *>(SMEM_STREAM ? take(4 * (size_t)stream_words) : p);
		s.dctab = reinterpret_cast<uint32_t *>(V3 ? take(4 * 1024) : p);
		s.gtot = reinterpret_cast<uint32_t *>(take(4 * (size_t)(nsgroups + 1)));
		s.grows = take((size_t)ngroups);
		s.lens = reinterpret_cast<uint16_t *>(take(2 * (size_t)padded));
		s.dcval = reinterpret_cast<int16_t *>(V3 ? take(2 * (size_t)padded) : p);
		s.stage = reinterpret_cast<uint32_t *>(p);
	}
	uint32_t *stream = SMEM_STREAM ? s.stream : gstream + (size_t)f * gstream_stride;

	const uint4 *fc = coefs + (size_t)f * frame_stride_u4;
	// STR mode: the budget follows from the frame index alone (the sector positions are worked out
	// again after the search: nothing of this stays live across it)
	int max_size;
	if (STR) {
		const StrFrame sf = str_frame(str, f);
		max_size = (int)(sf.k * str.sectors_num / str.sectors_den - sf.before) * 2016;
	} else {
		max_size = max_sizes ? max_sizes[f] : max_size_bound;   // no per-frame budgets: all frames get the bound
	}
	if (max_size > max_size_bound) max_size = 0;   // contract violation -> frame fails
	const int words = max_size > 0 ? (max_size + 3) / 4 + 2 : 0;

	for (int i = tid; i < LENLUT1_BYTES / 4; i += T)
		reinterpret_cast<uint32_t *>(s.lenlut - 1)[i] = reinterpret_cast<const uint32_t *>(g_lenlut1)[i];
	for (int i = tid; i < BS_VLC_ROWS * BS_VLC_COLS; i += T) s.vlc[i] = g_vlc[i];
	if (V3) for (int i = tid; i < 1024; i += T) s.dctab[i] = g_dcvlc[i];
	for (int i = tid; i < words; i += T) stream[i] = 0;
	if (tid < 8) s.misc[tid] = 0;
	for (int b = nblk + tid; b < padded; b += T) s.lens[b] = 0;   // scan padding
	// cluster mode may be launched with programmatic stream serialisation (PSXB200_PDL=1): everything
	// up to here then ran beside the FDCT kernel, and the plane may only be read once that grid has
	// completed (a no-op otherwise)
	if (CL > 1) asm volatile("griddepcontrol.wait;" ::: "memory");
	for (int g = tid; g < ngroups; g += T) {
		// every lane's meta row carries the group's longest list and the dense flag
		uint32_t longest = reinterpret_cast<const uint32_t *>(fc + (size_t)g * (BS_U4_PER_BLOCK * 32) + BS_META_ROW * 32)[3];
		s.grows[g] = (longest & BS_DENSE_FLAG) ? (uint8_t)0x80 : (uint8_t)((min(longest, 63u) + 7) >> 3);
	}
	if (V3) {
		for (int pi = tid; pi < ngroups * 32; pi += T) {
			int b = bs_plane_to_block(pi, cpad, nmb);
			if (b < 0) continue;
			const uint4 *gp = fc + (size_t)(pi >> 5) * (BS_U4_PER_BLOCK * 32) + (pi & 31);
			const uint4 meta = gp[BS_META_ROW * 32];
			s.dcval[b] = (int16_t)quant_dc(meta.z & 0xFFFFu, meta.x & 1u);
		}
	}
	__syncthreads();
	if (CL > 1) {
		// latency mode: pull the rows of this warp's groups towards L1 now — the walks below fetch a
		// group's rows one after the other, each a round trip to L2 otherwise
		for (int g = ngroups - 1 - (wid * CL + rank); g >= 0; g -= nw * CL) {
			const int rows = s.grows[g];
			const uint4 *gp = fc + (size_t)g * (BS_U4_PER_BLOCK * 32) + lane;
			for (int r = 0; r < ((rows & 0x80) ? 8 : rows); r++) asm volatile("prefetch.global.L1 [%0];" ::"l"(gp + r * 32));
		}
		cg::this_cluster().sync();   // every CTA of the cluster runs and has its tables before any remote store
	}
	if (V3) dc_delta_codes(codec, nmb, s.dcval, reinterpret_cast<int *>(s.misc + 8));
	// code of block b's DC delta (chroma table for Cr/Cb, luma for Y1..Y4)
	auto dc_code = [&](int b) { return s.dctab[((b % 6) < 2 ? 0 : 512) + (s.dcval[b] & 0x1FF)]; };

	// ---- (1) first-fit quant scale search ------------------------------------------------
	// Largest bit total (blocks only) that still fits: 8 + 2*ceil((bits + 10)/16) <= max_size. A
	// pass whose running total passes it has failed (the reference's writer overflows at that
	// point too, mdec.c:323-325) and is abandoned early.
	const int limit_bits = max_size >= 8 ? 16 * ((max_size - 8) >> 1) - 10 : -1;
	// frames with a tiny budget are not worth a second kernel
	const bool census_possible = CL == 1 && SMEM_STREAM && max_size >= 2016;   // (a cluster finishes its frame itself)
	int q = 1;
	if (BUSY) {
		// v2: 10-bit DC + 2-bit end of block per block; v3: DC codes are at least 2 bits long
		// (the pair table borrows the emit phase's staging area)
		q = census_first_candidate(fc, ngroups, cpad, nmb, s.grows, s.lenlut, s.misc + 8, reinterpret_cast<uint4 *>(s.stage),
		                           nblk * (V3 ? 4 : 12), limit_bits);
		q = max(q, 2);   // q = 1 failed in the first kernel
	}
	uint32_t total_bits = 0;
	for (; q < 64; q++) {
		const QuantScale qs(q);
		// the (usually busier) luma groups at the high end of the plane go first; drawing groups
		// from a shared ticket counter instead of this static round-robin measured no better
		uint32_t *total = &s.misc[q % 3];
		int visited = 0;
		for (int g = ngroups - 1 - (wid * CL + rank); g >= 0; g -= nw * CL) {
			const int b = bs_plane_to_block(g * 32 + lane, cpad, nmb);
			int bits = 0;
			if (b >= 0) {
				const uint4 *gp = fc + (size_t)g * (BS_U4_PER_BLOCK * 32) + lane;
				const int rows = s.grows[g];
				// the first pass has its own list walk (measured: -2.9 % of the kernel on q = 2 content)
				if (!BUSY && q == 1 && !(rows & 0x80)) bits = ac_bits_q1(gp, rows, s.lut1);
				else bits = (rows & 0x80) ? ac_bits_dense(gp, qs, s.lenlut) : ac_bits(gp, rows, qs, s.lut1);
				bits += 2 + (V3 ? (int)(dc_code(b) >> 24) : 10);
				if (CL > 1) {
#pragma unroll
					for (int r = 0; r < CL; r++) cg::this_cluster().map_shared_rank(s.lens, r)[b] = (uint16_t)bits;
				} else {
					s.lens[b] = (uint16_t)bits;
				}
			}
			visited++;
			const uint32_t sum = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)bits);
			uint32_t before = 0;
			if (lane == 0) before = atomicAdd(total, sum);
			before = __shfl_sync(0xFFFFFFFFu, before, 0);
			if ((int)(before + sum) > limit_bits) break;
		}
		if (!BUSY && q == 1 && lane == 0) atomicAdd(&s.misc[4], (uint32_t)visited);
		__syncthreads();
		if (CL > 1) {
			// the pass is over in every CTA (and all remote length stores have landed): sum the totals
			cg::this_cluster().sync();
			total_bits = 0;
#pragma unroll
			for (int r = 0; r < CL; r++) total_bits += *cg::this_cluster().map_shared_rank(total, r);
		} else {
			total_bits = *total;
		}
		if (tid == 0) s.misc[(q + 2) % 3] = 0;
		// stream = blocks + 10-bit end-of-frame code; byte budget rule of flush_bits
		int units = (int)((total_bits + 10 + 15) >> 4);
		if (8 + 2 * units <= max_size) break;
		// The first pass ran over its budget before it had seen half of the frame: the content is
		// far too busy for the small quant scales. Leave the frame to the BUSY kernel.
		if (!BUSY && q == 1 && census_possible && 2 * (int)s.misc[4] < ngroups) {
			if (tid == 0) results[f] = psxb200_bs_result_t{0, 0, 0, 0};
			return;
		}
	}

	uint32_t *out32 = reinterpret_cast<uint32_t *>(out + (size_t)f * out_stride);
	const StrFrame sf = STR ? str_frame(str, f) : StrFrame{0, 0, 0};
	// sector j of this frame: its slot in the file's output region
	auto str_sector = [&](int j) -> uint8_t * {
		return out + (size_t)sf.file * (size_t)str.file_stride + (size_t)(bs_str_slot(str, sf.before + j) - str.slot0) * str.sector_size;
	};
	// destination of 32-bit word i of the frame's bitstream buffer: contiguous, or sliced into
	// 2016-byte sector payloads behind their 32-byte headers (mdec.c:831-832)
	auto word_at = [&](int i) -> uint32_t * {
		if (!STR) return out32 + i;
		int j = i / 504;
		return reinterpret_cast<uint32_t *>(str_sector(j) + str.header_offset + 32) + (i - 504 * j);
	};
	// STR sector headers (mdec.c:782-820)
	auto write_str_headers = [&](uint32_t bytes_used, uint32_t bs0, uint32_t bs1) {
		const int chunks = max_size / 2016;
		for (int j = tid; j < chunks; j += T) {
			uint32_t *h = reinterpret_cast<uint32_t *>(str_sector(j) + str.header_offset);
			h[0] = 0x0160u | ((uint32_t)(str.video_id & 0xFFFF) << 16);
			h[1] = (uint32_t)j | ((uint32_t)chunks << 16);
			h[2] = (uint32_t)sf.k;
			h[3] = bytes_used;
			h[4] = (uint32_t)(str.width & 0xFFFF) | ((uint32_t)(str.height & 0xFFFF) << 16);
			h[5] = bs0;
			h[6] = bs1;
			h[7] = 0;
		}
	};
	if (q >= 64) {
		if (CL > 1 && rank != 0) return;   // (nobody reads a peer's shared memory on this path)
		for (int i = tid; i < (max_size >> 2); i += T) *word_at(i) = 0;
		if (!STR)
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
	// Per block, for each half of its list (rows 4.. first: they hold the lower positions): a
	// convergent pass parks the rows in the thread's shared-memory column and marks the entries
	// that are coefficients at q; the codes are then produced by walking the marks only.
	{
		const QuantScale qs(q);
		uint32_t *stage = s.stage + tid;
		uint32_t nnz = 0;
		for (int g = ngroups - 1 - (wid * CL + rank); g >= 0; g -= nw * CL) {
			// Lanes walk different numbers of coefficients, so the warp is brought back together
			// (__syncwarp) before every convergent stretch — left to itself it ran the staging code
			// below in up to seven separate lane groups per group of blocks (ncu: executed 395
			// instead of 57 times per frame). Padding lanes (b < 0) come along with empty lists.
			const int b = bs_plane_to_block(g * 32 + lane, cpad, nmb);
			const bool act = b >= 0;
			const uint4 *gp = fc + (size_t)g * (BS_U4_PER_BLOCK * 32) + lane;
			const uint4 meta = gp[BS_META_ROW * 32];
			const int nrows = s.grows[g];
			BitWriter bw;
			bw.begin(stream, act ? s.gtot[b >> 5] + s.lens[b] : 0u);
			if (act) {
				if (V3) {
					uint32_t e = dc_code(b);
					bw.put((int)(e >> 24), e & 0xFFFFFFu);
				} else {
					bw.put(10, (uint32_t)quant_dc(meta.z & 0xFFFFu, meta.x & 1u) & 0x3FFu);
				}
			}
			int prev = 0;
			const unsigned long long signs = ((unsigned long long)meta.y << 32) | meta.x;   // by zig-zag position
			// one coefficient: y at zig-zag position pos (mdec.c:484-499)
			auto put_coef = [&](uint32_t y, int pos) {
				const uint32_t lvl = __umulhi(y + qs.q, qs.m_hi);
				const int run = pos - prev - 1;
				prev = pos;
				const uint32_t neg = (uint32_t)(signs >> pos) & 1u;
				const uint32_t code = s.vlc[min(lvl, (uint32_t)(BS_VLC_ROWS - 1)) * BS_VLC_COLS + min(run, BS_VLC_COLS - 1)];
				if (code & 0xFFFFFFu) {
					bw.put((int)(code >> 24), (code & 0xFFFFFFu) | neg);
				} else {
					// escape: the level clamped to [-512, 510] (mdec.c:262-265) as 10-bit two's complement
					int level = neg ? -(int)min(lvl, 0x200u) : (int)min(lvl, 0x1FEu);
					bw.put(BS_AC_ESCAPE_BITS, (1u << 16) | ((uint32_t)run << 10) | ((uint32_t)level & 0x3FFu));
				}
			};
			if (nrows & 0x80) {
				for (int half = 0; half < 2; half++) {
					__syncwarp();
					uint32_t live = stage_dense<BUSY>(gp, half, qs, stage, T);
					if (!act) live = 0;
					nnz += __popc(live);
					while (live) {
						const int e = __ffs((int)live) - 1;
						live &= live - 1;
						const uint32_t word = stage[(e >> 1) * T];
						put_coef((e & 1) ? word >> 16 : word & 0xFFFFu, 32 * half + e);
					}
				}
			} else {
				for (int r0 = nrows > 4 ? 4 : 0; r0 >= 0; r0 -= 4) {
					__syncwarp();
					uint32_t live = stage_rows(gp, r0, min(nrows - r0, 4), qs, stage, T);
					if (!act) live = 0;
					nnz += __popc(live);
					while (live) {
						const int e = 32 - __ffs((int)live);   // local entry, highest (= lowest position) first
						live &= live - 1;
						const uint32_t word = stage[(e >> 1) * T];
						const uint32_t ent = (e & 1) ? word >> 16 : word & 0xFFFFu;
						put_coef(ent >> 6, (int)(ent & 63u));
					}
				}
			}
			__syncwarp();
			if (act) {
				bw.put(2, 2u);   // end of block (mdec.c:502)
				if (b == nblk - 1) bw.put(10, V3 ? 0x3FFu : 0x1FFu);   // end of frame (mdec.c:645-652, 710)
				bw.finish();
			}
		}
		__syncwarp();
		nnz = warp_sum(nnz);
		if (lane == 0) atomicAdd(&s.misc[3], nnz);
	}
	__syncthreads();

	// ---- (4) header, results, copy-out -----------------------------------------------------
	uint32_t coefficients = s.misc[3];
	// word i of the bitstream image: in a cluster the OR of the CTAs' images
	auto image_word = [&](int i) -> uint32_t {
		if (CL == 1) return stream[i];
		uint32_t x = 0;
#pragma unroll
		for (int r = 0; r < CL; r++) x |= cg::this_cluster().map_shared_rank(s.stream, r)[i];
		return x;
	};
	if (CL > 1) {
		cg::this_cluster().sync();   // every CTA's image and coefficient count are complete
		coefficients = 0;
#pragma unroll
		for (int r = 0; r < CL; r++) coefficients += cg::this_cluster().map_shared_rank(s.misc, r)[3];
	}
	int units = (int)((total_bits + 10 + 15) >> 4);
	int hwords = ((int)coefficients + 2 * nblk + 2 + 0x3F) & ~0x3F;   // mdec.c:497,507,719,726
	int blocks_used = (hwords + 1) >> 1;
	if (tid == 0 && rank == 0)
		results[f] = psxb200_bs_result_t{(8 + 2 * units + 3) & ~3, blocks_used, q, hwords};
	uint32_t hdr0 = (uint32_t)(blocks_used & 0xFFFF) | 0x38000000u;
	uint32_t hdr1 = (uint32_t)q | ((V3 ? 3u : 2u) << 16);
	// (a cluster's CTAs take contiguous shares of the words)
	const int nwords = max_size >> 2, share = (nwords + CL - 1) / CL;
	for (int i = rank * share + tid; i < min(nwords, (rank + 1) * share); i += T) {
		uint32_t v;
		if (i == 0) v = hdr0;
		else if (i == 1) v = hdr1;
		else { uint32_t x = image_word(i - 2); v = (x >> 16) | (x << 16); }
		*word_at(i) = v;
	}
	if (STR && rank == 0) write_str_headers((uint32_t)((8 + 2 * units + 3) & ~3), hdr0, hdr1);
	if (!STR && rank == 0 && tid < (max_size & 3)) {
		int i = (max_size & ~3) + tid;   // >= 8 here, since the frame fitted
		uint32_t x = image_word((i >> 2) - 2);
		uint32_t v = (x >> 16) | (x << 16);
		out[(size_t)f * out_stride + i] = (uint8_t)(v >> (8 * (i & 3)));
	}
	if (CL > 1) cg::this_cluster().sync();   // nobody leaves while a peer still reads its image
}

// ---- launchers ---------------------------------------------------------------------------

size_t bs_pack_smem_bytes(bool v3, bool smem_stream, const BsGeometry &geo, int max_size_bound, int threads) {
	const int ngroups = geo.ngroups;
	size_t padded = (size_t)geo.nsgroups * 32;
	size_t n = 0;
	auto take = [&](size_t bytes) { n += (bytes + 15) & ~(size_t)15; };
	take(LENLUT1_BYTES);                  // guard + lenlut
	take(4 * BS_VLC_ROWS * BS_VLC_COLS);  // vlc
	take(4 * (8 + 4 * 32));               // misc
	if (smem_stream) take(4 * (size_t)((max_size_bound + 3) / 4 + 2));
	if (v3) take(4 * 1024);               // dctab
	take(4 * (size_t)(geo.nsgroups + 1)); // gtot
	take((size_t)ngroups);                // grows
	take(2 * padded);                     // lens
	if (v3) take(2 * padded);             // dcval
	// stage: 16 words per thread; the census of the BUSY kernel borrows 4 KB of it for its pair table
	take(64 * (size_t)threads > 4096 ? 64 * (size_t)threads : 4096);
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

template <bool V3, bool SMEM_STREAM, bool STR, bool BUSY, int MAX_THREADS, int MIN_CTAS>
static cudaError_t launch_pack_t(int threads, size_t smem, int n, const uint4 *d_coefs, const BsGeometry &geo, int codec,
                                 const int *d_max_sizes, int max_size_bound, uint8_t *d_out, size_t out_stride,
                                 psxb200_bs_result_t *d_results, uint32_t *d_gstream, size_t gstream_stride,
                                 const BsStrLayout &str, cudaStream_t stream) {
	auto kern = bs_pack_kernel<V3, SMEM_STREAM, STR, BUSY, MAX_THREADS, MIN_CTAS>;
	// The opt-in to more than 48 KB of dynamic shared memory is a per-device attribute of this
	// instantiation: raised to the hardware maximum once per device.
	static bool configured[MAX_DEVICES];
	{
		int dev = 0;
		cudaError_t e = cudaGetDevice(&dev);
		if (e != cudaSuccess) return e;
		if (dev < 0 || dev >= MAX_DEVICES) return cudaErrorInvalidDevice;
		std::lock_guard<std::mutex> guard(g_device_lock);
		if (!configured[dev]) {
			int optin = 0;
			e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
			if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
			if (e != cudaSuccess) return e;
			configured[dev] = true;
		}
	}
	kern<<<n, threads, smem, stream>>>(d_coefs, geo.frame_stride_u4, geo.nblk, geo.ngroups, geo.nsgroups,
	                                   geo.cgroups * 32, geo.nmb, codec,
	                                   d_max_sizes, max_size_bound, d_out, out_stride, d_results, d_gstream,
	                                   gstream_stride, str);
	return cudaGetLastError();
}

// Cluster mode (see bs_pack_kernel): n * BS_PACK_CLUSTER CTAs, BS_PACK_CLUSTER of them per frame.
template <bool V3, bool STR>
static cudaError_t launch_pack_cluster_t(int threads, size_t smem, int n, const uint4 *d_coefs, const BsGeometry &geo, int codec,
                                         const int *d_max_sizes, int max_size_bound, uint8_t *d_out, size_t out_stride,
                                         psxb200_bs_result_t *d_results, const BsStrLayout &str, cudaStream_t stream) {
	auto kern = bs_pack_kernel<V3, true, STR, false, BS_PACK_MAX_THREADS, 1, BS_PACK_CLUSTER>;
	// Programmatic stream serialisation (the kernel's set-up beside the FDCT kernel in front of it) is opt-in:
	// measured -2 us per encode_frame_bs call, nothing inside the captured look-ahead graph, and one
	// pathological run of the latter (263 instead of 51 us per frame) — PSXB200_PDL=1 turns it on.
	static const bool pdl = getenv("PSXB200_PDL") != nullptr;
	static bool configured[MAX_DEVICES];
	{
		int dev = 0;
		cudaError_t e = cudaGetDevice(&dev);
		if (e != cudaSuccess) return e;
		if (dev < 0 || dev >= MAX_DEVICES) return cudaErrorInvalidDevice;
		std::lock_guard<std::mutex> guard(g_device_lock);
		if (!configured[dev]) {
			int optin = 0;
			e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
			if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
			if (e != cudaSuccess) return e;
			configured[dev] = true;
		}
	}
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)n * BS_PACK_CLUSTER);
	cfg.blockDim = dim3((unsigned)threads);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = stream;
	cudaLaunchAttribute attr[2];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = BS_PACK_CLUSTER;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	// (opt-in) the kernel's set-up overlaps the FDCT kernel in front of it (griddepcontrol.wait in the kernel)
	attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[1].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = pdl ? 2 : 1;
	uint32_t *no_gstream = nullptr;
	return cudaLaunchKernelEx(&cfg, kern, d_coefs, geo.frame_stride_u4, geo.nblk, geo.ngroups, geo.nsgroups, geo.cgroups * 32,
	                          geo.nmb, codec, d_max_sizes, max_size_bound, d_out, out_stride, d_results, no_gstream, (size_t)0, str);
}

cudaError_t bs_launch_pack_cluster(int codec, int threads, int n, const uint4 *d_coefs, const BsGeometry &geo,
                                   const int *d_max_sizes, int max_size_bound, uint8_t *d_out, size_t out_stride,
                                   psxb200_bs_result_t *d_results, const BsStrLayout &str, cudaStream_t stream) {
	const bool v3 = codec != 0;
	const size_t smem = bs_pack_smem_bytes(v3, true, geo, max_size_bound, threads);
#define PSXB200_CL_ARGS threads, smem, n, d_coefs, geo, codec, d_max_sizes, max_size_bound, d_out, out_stride, d_results, str, stream
	if (str.sector_size) return v3 ? launch_pack_cluster_t<true, true>(PSXB200_CL_ARGS) : launch_pack_cluster_t<false, true>(PSXB200_CL_ARGS);
	return v3 ? launch_pack_cluster_t<true, false>(PSXB200_CL_ARGS) : launch_pack_cluster_t<false, false>(PSXB200_CL_ARGS);
#undef PSXB200_CL_ARGS
}

template <bool BUSY, int MAX_THREADS, int MIN_CTAS>
static cudaError_t launch_pack_cfg(bool v3, bool smem_stream, int threads, size_t smem, int n, const uint4 *d_coefs,
                                   const BsGeometry &geo, int codec, const int *d_max_sizes, int max_size_bound,
                                   uint8_t *d_out, size_t out_stride, psxb200_bs_result_t *d_results,
                                   uint32_t *d_gstream, size_t gstream_stride, const BsStrLayout &str,
                                   cudaStream_t stream) {
#define PSXB200_PACK_ARGS threads, smem, n, d_coefs, geo, codec, d_max_sizes, max_size_bound, d_out, out_stride, \
	d_results, d_gstream, gstream_stride, str, stream
	if (BUSY || smem_stream) {   // the BUSY kernel is instantiated for the shared-memory image only
		if (str.sector_size)
			return v3 ? launch_pack_t<true, true, true, BUSY, MAX_THREADS, MIN_CTAS>(PSXB200_PACK_ARGS)
			          : launch_pack_t<false, true, true, BUSY, MAX_THREADS, MIN_CTAS>(PSXB200_PACK_ARGS);
		return v3 ? launch_pack_t<true, true, false, BUSY, MAX_THREADS, MIN_CTAS>(PSXB200_PACK_ARGS)
		          : launch_pack_t<false, true, false, BUSY, MAX_THREADS, MIN_CTAS>(PSXB200_PACK_ARGS);
	}
	if (str.sector_size)
		return v3 ? launch_pack_t<true, false, true, false, MAX_THREADS, MIN_CTAS>(PSXB200_PACK_ARGS)
		          : launch_pack_t<false, false, true, false, MAX_THREADS, MIN_CTAS>(PSXB200_PACK_ARGS);
	return v3 ? launch_pack_t<true, false, false, false, MAX_THREADS, MIN_CTAS>(PSXB200_PACK_ARGS)
	          : launch_pack_t<false, false, false, false, MAX_THREADS, MIN_CTAS>(PSXB200_PACK_ARGS);
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
	// register budget variants: (max threads per CTA, CTAs per SM the register file must hold).
	// The common kernel first, then — shared-memory image only — the kernel for the frames it deferred.
	cudaError_t e;
	if (threads <= 320) {
		if (min_ctas >= 4) {
			e = launch_pack_cfg<false, 320, 4>(PSXB200_CFG_ARGS);
			return e != cudaSuccess || !smem_stream ? e : launch_pack_cfg<true, 320, 4>(PSXB200_CFG_ARGS);
		}
		e = launch_pack_cfg<false, 320, 3>(PSXB200_CFG_ARGS);
		return e != cudaSuccess || !smem_stream ? e : launch_pack_cfg<true, 320, 3>(PSXB200_CFG_ARGS);
	}
	e = launch_pack_cfg<false, BS_PACK_MAX_THREADS, 1>(PSXB200_CFG_ARGS);
	return e != cudaSuccess || !smem_stream ? e : launch_pack_cfg<true, BS_PACK_MAX_THREADS, 1>(PSXB200_CFG_ARGS);
#undef PSXB200_CFG_ARGS
}

// ---- STR mode: sector framing + FORM1 EDC ------------------------------------------------
//
// One warp per (frame, sector of the frame). Writes what init_sector_buffer_video puts into a
// video sector before encode_sector_str runs (filefmt.c:73-92: for FORMAT_STRCD sync, BCD
// timecode of the sector's LBA, mode 2 and the doubled subheader, cdrom.c:55-74; for FORMAT_STR
// the doubled subheader at offset 0) and then what psx_cdrom_calculate_checksums(.., FORM1)
// does to the buffer it is handed (filefmt.c:474, cdrom.c:92-100): the EDC of bytes
// [0x10, 0x818) stored at 0x818. For FORMAT_STR that buffer is the 2336-byte sector itself, so the
// range is shifted by 16 bytes against the sector's real layout and covers 16 bytes the encoder
// never writes — reproduced as is (those bytes are whatever the output buffer held).
__global__ void __launch_bounds__(256)
str_frame_kernel(int n_frames, int max_chunks, uint8_t *__restrict__ out, const uint32_t *__restrict__ edc_tab, const BsStrLayout str) {
	__shared__ uint32_t tab[256 + 1024];
	for (int i = threadIdx.x; i < 256 + 1024; i += blockDim.x) tab[i] = edc_tab[i];   // byte table + FORM1 advance table
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (w >= (long long)n_frames * max_chunks) return;
	const int f = (int)(w / max_chunks), j = (int)(w - (long long)f * max_chunks);
	const int g = str.frame_base + f;
	const int file = str.frames_per_file > 0 ? g / str.frames_per_file : 0;
	const long long k = (long long)str.frame_index0 + (g - file * str.frames_per_file);
	const long long before = (k - 1) * str.sectors_num / str.sectors_den;
	const int chunks = (int)(k * str.sectors_num / str.sectors_den - before);
	if (j >= chunks) return;
	const long long v = before + j;
	uint8_t *sec = out + (size_t)file * (size_t)str.file_stride + (size_t)(bs_str_slot(str, v) - str.slot0) * str.sector_size;
	uint32_t *sec32 = reinterpret_cast<uint32_t *>(sec);

	// subheader: file, channel & 0x1F, DATA | RT, coding 0 (filefmt.c:84-87)
	const uint32_t sub = (uint32_t)(str.xa_file & 0xFF) | ((uint32_t)(str.xa_channel & 0x1F) << 8) | (0x48u << 16);
	if (str.format == FORMAT_STRCD) {
		const int t = (int)bs_str_lba(str, v) + 150;
		auto bcd = [](int x) { return (uint32_t)(x + (x / 10) * 6) & 0xFFu; };
		if (lane == 0) sec32[0] = 0xFFFFFF00u;
		if (lane == 1) sec32[1] = 0xFFFFFFFFu;
		if (lane == 2) sec32[2] = 0x00FFFFFFu;
		if (lane == 3) sec32[3] = bcd(t / 4500) | (bcd((t / 75) % 60) << 8) | (bcd(t % 75) << 16) | (2u << 24);
		if (lane == 4 || lane == 5) sec32[lane] = sub;
	} else {
		if (lane < 2) sec32[lane] = sub;
	}
	__threadfence_block();   // the warp reads its own stores back below
	__syncwarp();
	const uint32_t edc = warp_edc<EDC_PIECE_FORM1>(sec32 + 4, 0x808 / 4, tab, tab + 256);
	if (lane == 0) sec32[0x818 / 4] = edc;
}

cudaError_t bs_launch_str_framing(int n, int max_chunks, uint8_t *d_out, const BsStrLayout &str, cudaStream_t stream) {
	if (n <= 0 || max_chunks <= 0 || str.format == FORMAT_STRV) return cudaSuccess;
	const uint32_t *tab = edc_tables_device();
	if (!tab) return cudaErrorInitializationError;
	const long long warps = (long long)n * max_chunks;
	const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
	str_frame_kernel<<<grid, 256, 0, stream>>>(n, max_chunks, d_out, tab, str);
	return cudaGetLastError();
}

}  // namespace psxb200
