// SPU / XA ADPCM predictor search for sm_100a — the GPU side of psx_audio_spu_encode and
// psx_audio_xa_encode (reference libpsxav/adpcm.c:39-233, 293-376).
//
// A channel is a strictly sequential chain: every 28-sample unit starts from the decoded
// last two samples of the winning candidate of the previous unit (adpcm.c:135-136,186-190).
// Parallelism therefore comes from (a) the candidates of one unit — filters x {m-1,m,m+1}
// shifts, <= 15 — which run in the 16 lanes of a half-warp and are reduced with a shuffle
// arg-min, and (b) independent chains (channels / streams), two per warp.
//
//   adpcm_spu_kernel   half-warp per SPU stream; 16 output bytes per unit (adpcm.c:356-376)
//   adpcm_xa_kernel    warp per stereo XA stream (half-warp 0 = left, 1 = right) or per two
//                      mono streams (one per half-warp); sound groups are assembled in
//                      shared memory and stored as one 128-byte row
//                      (encode_block_xa, adpcm.c:193-233; header duplication :321-322)
//   xa_frame_kernel    sector sync/header/subheader (adpcm.c:266-291, cdrom.c:55-74) and the
//                      EDC CRC (cdrom.c:30-41, 102-109), one thread per sector
#include <cuda_runtime.h>

#include <cstdlib>
#include <stdint.h>

#include "adpcm_encode.h"
#include "edc.cuh"

namespace psxb200 {

constexpr int UNIT = 28;

static_assert(sizeof(ChannelState) == 24, "state layout");

__device__ __forceinline__ int filter_k1(int f) { return f == 0 ? 0 : f == 1 ? 60 : f == 2 ? 115 : f == 3 ? 98 : 122; }
__device__ __forceinline__ int filter_k2(int f) { return f < 2 ? 0 : f == 2 ? -52 : f == 3 ? -55 : -60; }

__device__ __forceinline__ int bit_length(uint32_t v) { return 32 - __clz(v); }

// One unit, one candidate per lane of a half-warp. `sub` = lane & 15.
// On return every lane of the half-warp holds the winner's decoder state in (p1, p2), its
// error in `mse` and the header byte; `mine` tells whether this lane is the winner, whose
// `codes` (28 nibbles or bytes, sample 0 in the low bits of codes[0]) are the unit's data.
template <int FILTERS, int RANGE>
__device__ __forceinline__ void encode_unit(const int (&s)[UNIT], int qerr, int &p1, int &p2, int sub,
                                            uint32_t (&codes)[RANGE == 12 ? 4 : 7], unsigned long long &mse,
                                            int &header, bool &mine) {
	constexpr int LO = -0x8000 >> RANGE, HI = 0x7FFF >> RANGE;
	constexpr int KEEP = 16 - RANGE - 1;         // magnitude bits that fit without shifting
	constexpr int BITS = 16 - RANGE;             // code width
	constexpr uint32_t MASK = 0xFFFFu >> RANGE;

	const int filter = sub / 3, which = sub - 3 * filter;
	const bool has_candidate = filter < FILTERS;
	const int k1 = filter_k1(has_candidate ? filter : 0), k2 = filter_k2(has_candidate ? filter : 0);

	// find_min_shift (adpcm.c:39-79): open-loop residual range over the raw samples. The range
	// does not depend on the shift, so the three shift candidates of a filter scan a third of
	// the unit each (samples 0-9, 10-18, 19-27) and combine the bit lengths of their extremes.
	int lo = 0, hi = 0;
	{
		// branch-free three-way selects (masks), so that the compiler keeps one copy of the code
		const int m0 = -(which == 0), m1 = -(which == 1), m2 = -(which == 2);
		int q1 = (p1 & m0) | (s[9] & m1) | (s[18] & m2);
		int q2 = (p2 & m0) | (s[8] & m1) | (s[17] & m2);
#pragma unroll
		for (int j = 0; j < 10; j++) {
			const int x = (s[j] & m0) | (s[j < 9 ? j + 10 : 18] & m1) | (s[j < 9 ? j + 19 : 27] & m2);
			int r = x - ((k1 * q1 + k2 * q2 + 32) >> 6);
			if (j == 9) r &= m0;   // the second and third parts are 9 samples long
			lo = min(lo, r);
			hi = max(hi, r);
			q2 = q1;
			q1 = x;
		}
	}
	int bl = max(bit_length((uint32_t)hi), bit_length((uint32_t)max(~lo, 0)));
	{
		const int first = sub - which;   // lanes first..first+2 hold this filter
		const int b1 = __shfl_sync(0xFFFFFFFFu, bl, first + which + 1 - 3 * (which == 2), 16);   // (which + 1) % 3
		const int b2 = __shfl_sync(0xFFFFFFFFu, bl, first + which + 2 - 3 * (which != 0), 16);   // (which + 2) % 3
		bl = max(bl, max(b1, b2));
	}
	int rs = bl - KEEP;
	rs = min(max(rs, 0), RANGE);
	const int shift = RANGE - rs + which - 1;    // candidates m-1, m, m+1 (adpcm.c:161-167)
	const bool valid = has_candidate && shift >= 0 && shift <= RANGE;
	const int sh = min(max(shift, 0), RANGE);

	// attempt_to_encode (adpcm.c:81-140): closed-loop trial
	int t1 = p1, t2 = p2;
	unsigned long long err2 = 0;
#pragma unroll
	for (int j = 0; j < (RANGE == 12 ? 4 : 7); j++) codes[j] = 0;
	// ((x << sh) + 2^(RANGE-1)) >> RANGE == (x + (2^(RANGE-1) >> sh)) >> (RANGE - sh) for 0 <= sh <= RANGE
	// (floor division by a power of two; nothing overflows: |x| < 2^17), which takes the left
	// shift off the serial chain; qerr rides along in the rounding term.
	const int down = RANGE - sh;
	const int scale = 1 << down;   // decoding multiplies (IMAD, FMA pipe) instead of shifting: the loop is ALU-pipe bound
	const int round_q = ((1 << (RANGE - 1)) >> sh) + qerr;
#pragma unroll
	for (int i = 0; i < UNIT; i++) {
		int pred = (k1 * t1 + k2 * t2 + 32) >> 6;
		int e = (s[i] - pred + round_q) >> down;
		e = min(max(e, LO), HI);
		int dec;
		asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(dec) : "r"(e), "r"(scale), "r"(pred));   // (e << down) + pred
		dec = min(max(dec, -0x8000), 0x7FFF);
		int d = dec - s[i] - qerr;
		err2 += (unsigned long long)((long long)d * d);
		codes[(i * BITS) >> 5] |= ((uint32_t)e & MASK) << ((i * BITS) & 31);
		t2 = t1;
		t1 = dec;
	}

	// arg-min with first-wins ties in (filter, shift) order == lane order (adpcm.c:177-181)
	unsigned long long key = valid ? ((err2 << 4) | (unsigned)sub) : ~0ull;
	unsigned long long best = key;
#pragma unroll
	for (int o = 8; o > 0; o >>= 1) {
		unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, best, o, 16);
		best = other < best ? other : best;
	}
	const int winner = (int)(best & 15);
	mine = sub == winner;
	mse = best >> 4;
	p1 = __shfl_sync(0xFFFFFFFFu, t1, winner, 16);
	p2 = __shfl_sync(0xFFFFFFFFu, t2, winner, 16);
	header = __shfl_sync(0xFFFFFFFFu, (sh & 0x0F) | (filter << 4), winner, 16);
}

// Samples of one 28-sample unit, spread over the 16 lanes of a half-warp: lane `sub` holds
// samples sub and sub+16. Fetching the NEXT unit this way (2 coalesced loads per lane) while the
// current one is being searched takes the global-memory latency off the serial chain; the
// values are then broadcast to every lane with shuffles (each candidate needs all 28).
struct UnitFetch { int a, b; };

__device__ __forceinline__ UnitFetch fetch_unit(const int16_t *__restrict__ src, long pitch, int limit, int sub) {
	UnitFetch f;
	f.a = sub < limit ? (int)__ldg(src + sub * pitch) : 0;
	f.b = (sub + 16 < UNIT && sub + 16 < limit) ? (int)__ldg(src + (sub + 16) * pitch) : 0;
	return f;
}

__device__ __forceinline__ void spread_unit(const UnitFetch &f, int (&s)[UNIT]) {
#pragma unroll
	for (int i = 0; i < UNIT; i++) s[i] = __shfl_sync(0xFFFFFFFFu, i < 16 ? f.a : f.b, i & 15, 16);
}

// ---- SPU ---------------------------------------------------------------------------------

__global__ void __launch_bounds__(ADPCM_THREADS, 8)
adpcm_spu_kernel(int n_streams, int stream_first, int stream_step, const int16_t *__restrict__ samples, int pitch,
                 long group_stride, int sample_count, const int *__restrict__ counts, ChannelState *__restrict__ states,
                 uint8_t *__restrict__ out, long out_stride) {
	const int lane = threadIdx.x & 31, sub = lane & 15;
	// the launch's i-th chain is stream stream_first + i * stream_step of the caller's arrays
	const int index = (int)(((long)blockIdx.x * ADPCM_THREADS + threadIdx.x) >> 4);
	const bool live = index < n_streams;
	const int stream = stream_first + index * stream_step;
	const int count = live ? (counts ? counts[stream] : sample_count) : 0;
	const int units = (count + UNIT - 1) / UNIT;
	// both half-warps of a warp iterate together
	const int warp_units = max(units, __shfl_xor_sync(0xFFFFFFFFu, units, 16));

	const int16_t *src = samples;
	ChannelState st{0, 0, 0, 0, 0};
	if (live) {
		src += (long)(stream / pitch) * group_stride + (stream % pitch);
		st = states[stream];
	}
	int p1 = st.prev1, p2 = st.prev2;
	unsigned long long mse = st.mse;
	uint8_t *dst = out + (long)(live ? stream : 0) * out_stride;

	UnitFetch next = fetch_unit(src, pitch, count, sub);
	for (int u = 0; u < warp_units; u++) {
		const bool active = u < units;
		int s[UNIT];
		spread_unit(next, s);
		next = fetch_unit(src + (long)(u + 1) * UNIT * pitch, pitch, u + 1 < units ? count - (u + 1) * UNIT : 0, sub);
		uint32_t codes[4];
		unsigned long long m;
		int header;
		bool mine;
		int n1 = p1, n2 = p2;
		encode_unit<5, 12>(s, st.qerr, n1, n2, sub, codes, m, header, mine);
		if (active) {
			p1 = n1; p2 = n2; mse = m;
			if (mine) {
				// header, flags = 0, 14 bytes of nibble pairs, low nibble first (adpcm.c:367-372)
				uint4 blk;
				blk.x = (uint32_t)header | (codes[0] << 16);
				blk.y = (codes[0] >> 16) | (codes[1] << 16);
				blk.z = (codes[1] >> 16) | (codes[2] << 16);
				blk.w = (codes[2] >> 16) | (codes[3] << 16);
				*reinterpret_cast<uint4 *>(dst + 16L * u) = blk;
			}
		}
	}
	if (live && sub == 0 && units > 0) {
		st.prev1 = p1; st.prev2 = p2; st.mse = mse;
		states[stream] = st;
	}
}

// One short chain whose samples and state travel in the kernel's parameters and whose blocks,
// state and completion flag go to page-locked host memory mapped into the device address space:
// the drop-in psx_audio_spu_encode is called with one 28-sample block at a time by the
// reference's own encode_file_spu (filefmt.c:243), so what counts there is the latency of a
// call — no read over the link, no copy engine, and the host sees the flag as soon as the
// results have landed instead of waiting for the stream.
__global__ void __launch_bounds__(32)
adpcm_spu_small_kernel(const __grid_constant__ SpuSmallCall call) {
	const int lane = threadIdx.x & 31, sub = lane & 15;
	const bool live = lane < 16;   // one chain: the second half-warp only keeps the shuffles of encode_unit company
	const int units = (call.count + UNIT - 1) / UNIT;
	int p1 = call.state.prev1, p2 = call.state.prev2;
	unsigned long long mse = call.state.mse;
	for (int u = 0; u < units; u++) {
		int s[UNIT];
#pragma unroll
		for (int i = 0; i < UNIT; i++) s[i] = u * UNIT + i < call.count ? (int)call.samples[u * UNIT + i] : 0;
		uint32_t codes[4];
		unsigned long long m;
		int header;
		bool mine;
		int n1 = p1, n2 = p2;
		encode_unit<5, 12>(s, call.state.qerr, n1, n2, sub, codes, m, header, mine);
		p1 = n1; p2 = n2; mse = m;
		if (live && mine) {
			uint4 blk;   // header, flags = 0, 14 bytes of nibble pairs, low nibble first (adpcm.c:367-372)
			blk.x = (uint32_t)header | (codes[0] << 16);
			blk.y = (codes[0] >> 16) | (codes[1] << 16);
			blk.z = (codes[1] >> 16) | (codes[2] << 16);
			blk.w = (codes[2] >> 16) | (codes[3] << 16);
			*reinterpret_cast<uint4 *>(call.out + 16 * u) = blk;
		}
	}
	if (lane == 0) {
		ChannelState st = call.state;
		if (units > 0) { st.prev1 = p1; st.prev2 = p2; st.mse = mse; }
		*call.state_out = st;
	}
	__threadfence_system();   // every lane's stores are out before the flag
	__syncwarp();
	if (lane == 0) *call.flag = call.seq;
}

cudaError_t adpcm_launch_spu_small(const SpuSmallCall &call, cudaStream_t stream) {
	adpcm_spu_small_kernel<<<1, 32, 0, stream>>>(call);
	return cudaGetLastError();
}

// ---- XA ----------------------------------------------------------------------------------

// 4 or 8 bits per sample; mono and stereo are compiled separately; WARPS per CTA: 4 spreads the
// chains over the machine, 16 packs them onto few SMs (see adpcm_launch_xa)
template <int BITS, bool STEREO, int WARPS, int MIN_CTAS>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS)
adpcm_xa_kernel(int n_streams, int sector_size, long sector_stride, const int16_t *__restrict__ samples, long in_stride,
                int sample_count, ChannelState *__restrict__ states, uint8_t *__restrict__ out, long out_stride) {
	constexpr int RANGE = BITS == 4 ? 12 : 8;
	constexpr int UNITS = BITS == 4 ? 8 : 4;        // units per 128-byte sound group
	constexpr int JUMP = BITS == 4 ? 224 : 112;     // interleaved samples per sound group
	// unit codes, one byte per sample, and unit headers: one set per stream of the warp
	constexpr bool stereo = STEREO;
	constexpr int SETS = STEREO ? 1 : 2;
	__shared__ uint8_t stage[WARPS][SETS][UNITS][32];
	__shared__ uint8_t hdrs[WARPS][SETS][UNITS];

	const int lane = threadIdx.x & 31, sub = lane & 15, half = lane >> 4, wslot = threadIdx.x >> 5;
	// stereo: the warp's stream, half-warp 0 drives the left channel, 1 the right one;
	// mono: each half-warp drives a stream of its own
	const int warp_id = blockIdx.x * WARPS + wslot;
	const int stream = stereo ? warp_id : 2 * warp_id + half;
	const int set = stereo ? 0 : half;              // which stage/hdrs set this half-warp fills
	if ((stereo ? warp_id : 2 * warp_id) >= n_streams) return;
	const bool chain = stream < n_streams;          // false: the odd stream out of a mono pair

	const int16_t *base = samples + (long)(chain ? stream : 0) * in_stride;
	const int total = stereo ? sample_count * 2 : sample_count;
	const int groups = ((total + JUMP - 1) / JUMP + 17) / 18 * 18;   // padded to whole sectors (adpcm.c:310)
	ChannelState st = states[(long)(chain ? stream : 0) * 2 + (stereo ? half : 0)];
	int p1 = st.prev1, p2 = st.prev2;
	unsigned long long mse = st.mse;
	const int steps = stereo ? UNITS / 2 : UNITS;   // sequential units per chain per group

	// unit (j, step) of this half-warp's chain: source pointer and sample limit. Stereo: the
	// pointer advances 56 interleaved samples per L/R pair but the limit only drops by 28
	// (adpcm.c:204-211); mono: 28 and 28.
	auto fetch = [&](int j, int step) {
		const int16_t *gsrc = base + (long)j * JUMP;
		const int16_t *src = stereo ? gsrc + 56 * step + half : gsrc + 28 * step;
		int limit = (chain && j < groups) ? total - j * JUMP - 28 * step : 0;
		return fetch_unit(src, stereo ? 2 : 1, limit, sub);
	};
	UnitFetch next = fetch(0, 0);
	for (int j = 0; j < groups; j++) {
		for (int step = 0; step < steps; step++) {
			const int unit = stereo ? 2 * step + half : step;
			int s[UNIT];
			spread_unit(next, s);
			next = step + 1 < steps ? fetch(j, step + 1) : fetch(j + 1, 0);
			uint32_t codes[BITS == 4 ? 4 : 7];
			unsigned long long m;
			int header;
			bool mine;
			int n1 = p1, n2 = p2;
			encode_unit<4, RANGE>(s, st.qerr, n1, n2, sub, codes, m, header, mine);
			if (chain) {
				p1 = n1; p2 = n2; mse = m;
				if (mine) {
					hdrs[wslot][set][unit] = (uint8_t)header;
#pragma unroll
					for (int i = 0; i < UNIT; i++)
						stage[wslot][set][unit][i] = (uint8_t)((codes[(i * BITS) >> 5] >> ((i * BITS) & 31)) & (BITS == 4 ? 0xF : 0xFF));
				}
			}
		}
		__syncwarp();
		// assemble the 128-byte sound group(s): words 0-3 headers (with duplicates), 4-31 data rows;
		// the whole warp stores one group per round — one round when stereo, one per stream when mono
#pragma unroll
		for (int g = 0; g < SETS; g++) {
			const int gstream = stereo ? warp_id : 2 * warp_id + g;
			if (gstream >= n_streams) break;
			const uint8_t(*stg)[32] = stage[wslot][g];
			const uint8_t *hdr = hdrs[wslot][g];
			uint32_t word;
			if (lane < 4) {
				if (BITS == 4) {
					int h = (lane >> 1) * 4;   // words 0,1 <- units 0-3; words 2,3 <- units 4-7
					word = hdr[h] | (hdr[h + 1] << 8) | (hdr[h + 2] << 16) | (hdr[h + 3] << 24);
				} else {
					// 8-bit: bytes 0-3 = unit headers, copied to 4-7; bytes 8-15 keep the caller's
					// content in the reference (copied 8-11 -> 12-15, adpcm.c:322) and are
					// written as zero here (the batch API requires zeroed output buffers).
					word = lane < 2 ? (hdr[0] | (hdr[1] << 8) | (hdr[2] << 16) | (hdr[3] << 24)) : 0u;
				}
			} else {
				int i = lane - 4;
				if (BITS == 4) {
					word = 0;
#pragma unroll
					for (int k = 0; k < 4; k++)
						word |= (uint32_t)(stg[2 * k][i] | (stg[2 * k + 1][i] << 4)) << (8 * k);
				} else {
					word = stg[0][i] | (stg[1][i] << 8) | (stg[2][i] << 16) | (stg[3][i] << 24);
				}
			}
			uint8_t *sec = out + (long)gstream * out_stride + (long)(j / 18) * sector_stride - (2352 - sector_size);
			uint8_t *grp = sec + 24 + (j % 18) * 128;
			if (BITS == 8 && lane >= 2 && lane < 4) {
				// leave bytes 8-15 of 8-bit groups exactly as the reference does: 12-15 := 8-11
				if (lane == 3) {
					uint32_t keep = *reinterpret_cast<const uint32_t *>(grp + 8);
					*reinterpret_cast<uint32_t *>(grp + 12) = keep;
				}
			} else {
				*reinterpret_cast<uint32_t *>(grp + 4 * lane) = word;
			}
		}
		__syncwarp();
	}
	if (chain && sub == 0 && groups > 0) {
		st.prev1 = p1; st.prev2 = p2; st.mse = mse;
		states[(long)stream * 2 + (stereo ? half : 0)] = st;
	}
}

// ---- XA sector framing + EDC -------------------------------------------------------------

__device__ __forceinline__ uint32_t to_bcd(int v) { return (uint32_t)(v + (v / 10) * 6) & 0xFFu; }

// One warp per sector: sync/header (XACD only; psx_cdrom_init_sector, cdrom.c:55-74), the
// doubled subheader (adpcm.c:266-291) and the FORM2 EDC over bytes 0x10..0x92B (cdrom.c:102-109).
__global__ void __launch_bounds__(128)
xa_frame_kernel(int n_streams, int sectors_per_stream, int format, int sector_size, long sector_stride, int file_number,
                int channel_number, int coding, int lba, int lba_step, uint8_t *__restrict__ out, long out_stride,
                const uint32_t *__restrict__ edc_tab) {
	__shared__ uint32_t tab[256 + 1024];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = edc_tab[i];
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[256 + i] = edc_tab[1280 + i];   // FORM2 advance table
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const long id = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (id >= (long)n_streams * sectors_per_stream) return;
	const int stream = (int)(id / sectors_per_stream), k = (int)(id - (long)stream * sectors_per_stream);
	uint8_t *sec = out + (long)stream * out_stride + (long)k * sector_stride - (2352 - sector_size);
	uint32_t *sec32 = reinterpret_cast<uint32_t *>(sec);

	if (format == 1) {
		const int t = lba + k * lba_step + 150;
		if (lane == 0) sec32[0] = 0xFFFFFF00u;
		if (lane == 1) sec32[1] = 0xFFFFFFFFu;
		if (lane == 2) sec32[2] = 0x00FFFFFFu;
		if (lane == 3) sec32[3] = to_bcd(t / 4500) | (to_bcd((t / 75) % 60) << 8) | (to_bcd(t % 75) << 16) | (2u << 24);
	}
	if (lane == 4) {
		// file, channel & 0x1F, AUDIO | FORM2 | RT (adpcm.c:272-275); the coding byte is OR-ed into
		// what the buffer holds in the 2336-byte format (adpcm.c:277-288; psx_cdrom_init_sector has
		// zeroed it in the 2352-byte one)
		const uint32_t prior = format == 1 ? 0u : (uint32_t)sec[19];
		const uint32_t sub = (uint32_t)(file_number & 0xFF) | ((uint32_t)(channel_number & 0x1F) << 8) | (0x64u << 16) |
		                     (((prior | (uint32_t)coding) & 0xFFu) << 24);
		sec32[4] = sub;
		sec32[5] = sub;
	}
	__threadfence_block();   // the warp reads its own stores back below
	__syncwarp();
	const uint32_t edc = warp_edc<EDC_PIECE_FORM2>(sec32 + 4, 0x91C / 4, tab, tab + 256);
	if (lane == 0) sec32[0x92C / 4] = edc;
}

// ---- launchers ---------------------------------------------------------------------------

cudaError_t adpcm_launch_spu(int n_streams, const int16_t *d_samples, int pitch, long group_stride, int sample_count,
                             const int *d_counts, void *d_states, uint8_t *d_out, long out_stride, cudaStream_t stream,
                             int stream_first, int stream_step) {
	if (n_streams <= 0) return cudaSuccess;
	long threads = (long)n_streams * 16;
	unsigned grid = (unsigned)((threads + ADPCM_THREADS - 1) / ADPCM_THREADS);
	adpcm_spu_kernel<<<grid, ADPCM_THREADS, 0, stream>>>(n_streams, stream_first, stream_step, d_samples, pitch, group_stride,
	                                                     sample_count, d_counts, static_cast<ChannelState *>(d_states), d_out,
	                                                     out_stride);
	return cudaGetLastError();
}

// Number of int16 elements of `samples` that psx_audio_xa_encode reads (adpcm.c:193-233,
// 310-319): in stereo the per-unit limit shrinks by 28 while the pointer advances by 56, so
// the tail group may be read past sample_count*2 (never past its own 224/112 samples).
long adpcm_xa_input_extent(int stereo, int bits, int sample_count) {
	const int jump = bits == 8 ? 112 : 224;
	const int units = bits == 8 ? 4 : 8;
	const long total = stereo ? 2L * sample_count : sample_count;
	if (!stereo || total <= 0) return total > 0 ? total : 0;
	const long j = (total - 1) / jump;   // only the last group holding samples can over-read
	const long remaining = total - j * jump;
	long furthest = 0;
	for (int step = 0; step < units / 2; step++) {
		long lim = remaining - 28L * step < 28 ? remaining - 28L * step : 28;
		if (lim > 0 && 56L * step + 2 * lim > furthest) furthest = 56L * step + 2 * lim;
	}
	return j * jump + furthest;   // all earlier groups are read completely
}

int adpcm_xa_sectors(int stereo, int bits_per_sample, int sample_count) {
	int jump = bits_per_sample == 8 ? 112 : 224;
	int total = stereo ? sample_count * 2 : sample_count;
	return ((total + jump - 1) / jump + 17) / 18;
}

cudaError_t adpcm_launch_xa(int n_streams, int format, int stereo, int frequency, int bits_per_sample, int file_number,
                            int channel_number, const int16_t *d_samples, long in_stride, int sample_count, int lba,
                            int lba_step, void *d_states, uint8_t *d_out, long out_stride, long sector_stride,
                            const uint32_t *d_edc_tables, cudaStream_t stream) {
	int sectors = adpcm_xa_sectors(stereo, bits_per_sample, sample_count);
	if (n_streams <= 0 || sectors == 0) return cudaSuccess;
	int sector_size = format == 0 ? 2336 : 2352;
	if (sector_stride <= 0) sector_stride = sector_size;
	const int warps = stereo ? n_streams : (n_streams + 1) / 2;   // a warp takes two mono streams
	// Two shapes (profiles/r2_xa_shape_ab.txt). Spread: 4 warps per CTA, one chain-pair per scheduler on
	// as many SMs as there are CTAs — lowest latency when the chains have the machine to themselves.
	// Packed: 8 warps per CTA at 127 registers, used when the sectors go into a muxed image (a sector
	// stride that leaves room for other sectors), i.e. when the video kernels of the same files run
	// beside the chains: they then crowd half as many SMs and the two kernels overlap far better
	// (512 files x 8 frames + 10 XA sectors: 1.93 -> 1.59 ms per step) at 15 % more latency alone.
	const bool packed = sector_stride > sector_size;
	cudaError_t e = cudaSuccess;
	auto launch = [&](auto kern, int W) {
		unsigned grid = (unsigned)((warps + W - 1) / W);
		kern<<<grid, W * 32, 0, stream>>>(n_streams, sector_size, sector_stride, d_samples, in_stride, sample_count,
		                                  static_cast<ChannelState *>(d_states), d_out, out_stride);
	};
#define PSXB200_XA_LAUNCH(W, M)                                                                       \
	do {                                                                                              \
		if (bits_per_sample == 8) { if (stereo) launch(adpcm_xa_kernel<8, true, W, M>, W); else launch(adpcm_xa_kernel<8, false, W, M>, W); } \
		else { if (stereo) launch(adpcm_xa_kernel<4, true, W, M>, W); else launch(adpcm_xa_kernel<4, false, W, M>, W); }                     \
	} while (0)
	if (packed) PSXB200_XA_LAUNCH(8, 1);
	else PSXB200_XA_LAUNCH(4, 4);
#undef PSXB200_XA_LAUNCH
	if (e != cudaSuccess) return e;
	e = cudaGetLastError();
	if (e != cudaSuccess || !d_edc_tables) return e;
	int coding = (stereo ? 1 : 0) | (frequency == 37800 ? 0 : 4) | (bits_per_sample == 8 ? 16 : 0);
	long n = (long)n_streams * sectors;
	xa_frame_kernel<<<(unsigned)((n * 32 + 127) / 128), 128, 0, stream>>>(n_streams, sectors, format, sector_size, sector_stride,
	                                                                       file_number, channel_number, coding, lba, lba_step,
	                                                                       d_out, out_stride, d_edc_tables);
	return cudaGetLastError();
}

}  // namespace psxb200
