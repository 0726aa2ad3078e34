/*
 * psx_oracle.c — CPU restatement of the psxavenc encode core (see psx_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: never linked into the product. Written from the behaviour of
 * the reference (citations are file:line under /root/reference) in plain scalar C, with
 * its own structure: coefficients are transformed once into scan order, the bit writer is
 * a 16-bit-word MSB-first packer with the reference's byte-budget rule, and the ADPCM
 * search works on a zero-padded local copy of each 28-sample unit.
 */
#include "psx_oracle.h"
#include "orc_tables.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* FNV-1a 64                                                      */

uint64_t orc_fnv1a64(const uint8_t *data, long length) {
	uint64_t h = 1469598103934665603ull;
	for (long i = 0; i < length; i++) {
		h ^= data[i];
		h *= 1099511628211ull;
	}
	return h;
}

/* ------------------------------------------------------------------------------------ */
/* FDCT model 1: FFmpeg ff_jpeg_fdct_islow_8 (reached via AVDCT.fdct, mdec.c:640, when    */
/* FFmpeg has no x86 SIMD or dct_algo=FF_DCT_INT). SURVEY.md Appendix A.1.                */

#define ISLOW_CONST_BITS 13
#define ISLOW_PASS1_BITS 4

static inline int32_t rshift_round(int32_t v, int n) {
	return (v + (1 << (n - 1))) >> n;
}

/* One 8-point pass over v[0], v[stride], ... ; `second` selects the column-pass scaling. */
static void islow_1d(int16_t *v, int stride, int second) {
	int32_t in[8];
	for (int i = 0; i < 8; i++)
		in[i] = v[i * stride];

	int32_t e0 = in[0] + in[7], o0 = in[0] - in[7];
	int32_t e1 = in[1] + in[6], o1 = in[1] - in[6];
	int32_t e2 = in[2] + in[5], o2 = in[2] - in[5];
	int32_t e3 = in[3] + in[4], o3 = in[3] - in[4];

	int32_t ee0 = e0 + e3, eo0 = e0 - e3;
	int32_t ee1 = e1 + e2, eo1 = e1 - e2;

	int sh = second ? ISLOW_CONST_BITS + ISLOW_PASS1_BITS : ISLOW_CONST_BITS - ISLOW_PASS1_BITS;
	int32_t r[8];

	if (second) {
		r[0] = rshift_round(ee0 + ee1, ISLOW_PASS1_BITS);
		r[4] = rshift_round(ee0 - ee1, ISLOW_PASS1_BITS);
	} else {
		r[0] = (ee0 + ee1) * (1 << ISLOW_PASS1_BITS);
		r[4] = (ee0 - ee1) * (1 << ISLOW_PASS1_BITS);
	}

	int32_t z = (eo1 + eo0) * 4433;
	r[2] = rshift_round(z + eo0 * 6270, sh);
	r[6] = rshift_round(z - eo1 * 15137, sh);

	/* odd part; o3..o0 are tmp4..tmp7 of the classic formulation */
	int32_t z1 = o3 + o0, z2 = o2 + o1, z3 = o3 + o1, z4 = o2 + o0;
	int32_t z5 = (z3 + z4) * 9633;
	int32_t t4 = o3 * 2446, t5 = o2 * 16819, t6 = o1 * 25172, t7 = o0 * 12299;
	z1 *= -7373;
	z2 *= -20995;
	z3 = z3 * -16069 + z5;
	z4 = z4 * -3196 + z5;
	r[7] = rshift_round(t4 + z1 + z3, sh);
	r[5] = rshift_round(t5 + z2 + z4, sh);
	r[3] = rshift_round(t6 + z2 + z3, sh);
	r[1] = rshift_round(t7 + z1 + z4, sh);

	for (int i = 0; i < 8; i++)
		v[i * stride] = (int16_t)r[i];
}

void orc_fdct_islow(int16_t *b) {
	for (int r = 0; r < 8; r++)
		islow_1d(b + 8 * r, 1, 0);
	for (int c = 0; c < 8; c++)
		islow_1d(b + c, 8, 1);
}

/* ------------------------------------------------------------------------------------ */
/* FDCT model 2: FFmpeg ff_fdct_sse2 (default AVDCT.fdct on SIMD-enabled x86-64 FFmpeg,   */
/* including the libavcodec in this image). 16-bit saturating column butterflies, then a  */
/* multiply-accumulate row pass. SURVEY.md Appendix A.2.                                  */

static inline int16_t sat16(int32_t v) {
	return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
}
static inline int16_t addsat(int16_t a, int16_t b) { return sat16((int32_t)a + b); }
static inline int16_t subsat(int16_t a, int16_t b) { return sat16((int32_t)a - b); }
static inline int16_t shl16(int16_t a, int n) { return (int16_t)(uint16_t)((uint16_t)a << n); }
static inline int16_t mulhi16(int16_t a, int16_t b) { return (int16_t)(((int32_t)a * b) >> 16); }

#define SSE2_TAN1 13036
#define SSE2_TAN2 27146
#define SSE2_TAN3 (-21746)
#define SSE2_COS4 23170

static void sse2_column(int16_t *b, int c) {
	int16_t x[8], y[8];
	for (int i = 0; i < 8; i++)
		x[i] = b[8 * i + c];

	int16_t s16 = shl16(addsat(x[1], x[6]), 3), s25 = shl16(addsat(x[2], x[5]), 3);
	int16_t s07 = shl16(addsat(x[0], x[7]), 3), s34 = shl16(addsat(x[3], x[4]), 3);
	int16_t m12 = subsat(s16, s25), p12 = addsat(s16, s25);
	int16_t m03 = subsat(s07, s34), p03 = addsat(s07, s34);

	y[0] = addsat(p03, p12);
	y[4] = subsat(p03, p12);
	y[2] = addsat(mulhi16(m12, SSE2_TAN2), m03) | 1;
	y[6] = subsat(mulhi16(m03, SSE2_TAN2), m12) | 1;

	int16_t d16 = shl16(subsat(x[1], x[6]), 4), d25 = shl16(subsat(x[2], x[5]), 4);
	int16_t d34 = shl16(subsat(x[3], x[4]), 3), d07 = shl16(subsat(x[0], x[7]), 3);
	int16_t p65 = mulhi16(addsat(d16, d25), SSE2_COS4) | 1;
	int16_t m65 = mulhi16(subsat(d16, d25), SSE2_COS4);
	int16_t p465 = addsat(d34, m65), m465 = subsat(d34, m65);
	int16_t m765 = subsat(d07, p65), p765 = addsat(d07, p65);

	y[1] = addsat(mulhi16(p465, SSE2_TAN1), p765) | 1;
	y[3] = subsat(m765, addsat(mulhi16(m465, SSE2_TAN3), m465));
	y[5] = addsat(addsat(mulhi16(m765, SSE2_TAN3), m765), m465);
	y[7] = subsat(mulhi16(p765, SSE2_TAN1), p465);

	for (int i = 0; i < 8; i++)
		b[8 * i + c] = y[i];
}

/* cosine sets C1..C7 of the four row tables (rows 0/4, 1/7, 2/6, 3/5) */
static const int16_t sse2_row_cos[4][7] = {
	{22725, 21407, 19266, 16384, 12873,  8867, 4520},
	{31521, 29692, 26722, 22725, 17855, 12299, 6270},
	{29692, 27969, 25172, 21407, 16819, 11585, 5906},
	{26722, 25172, 22654, 19266, 15137, 10426, 5315},
};
static const int sse2_row_table_of[8] = {0, 1, 2, 3, 0, 3, 2, 1};

static void sse2_row(int16_t *a, const int16_t *C /* C[k-1] = Ck */) {
	int32_t c1 = C[0], c2 = C[1], c3 = C[2], c4 = C[3], c5 = C[4], c6 = C[5], c7 = C[6];
	int32_t s0 = addsat(a[0], a[7]), s1 = addsat(a[1], a[6]);
	int32_t s2 = addsat(a[2], a[5]), s3 = addsat(a[3], a[4]);
	int32_t d0 = subsat(a[0], a[7]), d1 = subsat(a[1], a[6]);
	int32_t d2 = subsat(a[2], a[5]), d3 = subsat(a[3], a[4]);
	int32_t o[8];
	o[0] = (s0 + s1 + s2 + s3) * c4;
	o[1] = d0 * c1 + d1 * c3 + d2 * c5 + d3 * c7;
	o[2] = (s0 - s3) * c2 + (s1 - s2) * c6;
	o[3] = d0 * c3 - d1 * c7 - d2 * c1 - d3 * c5;
	o[4] = (s0 - s1 - s2 + s3) * c4;
	o[5] = d0 * c5 - d1 * c1 + d2 * c7 + d3 * c3;
	o[6] = (s0 - s3) * c6 - (s1 - s2) * c2;
	o[7] = d0 * c7 - d1 * c5 + d2 * c3 - d3 * c1;
	for (int k = 0; k < 8; k++)
		a[k] = sat16((o[k] + 65536) >> 17);
}

void orc_fdct_sse2(int16_t *b) {
	for (int c = 0; c < 8; c++)
		sse2_column(b, c);
	for (int r = 0; r < 8; r++)
		sse2_row(b + 8 * r, sse2_row_cos[sse2_row_table_of[r]]);
}

void orc_fdct_batch(int variant, int16_t *blocks, int count) {
	for (int i = 0; i < count; i++) {
		if (variant == ORC_FDCT_SSE2)
			orc_fdct_sse2(blocks + 64 * i);
		else
			orc_fdct_islow(blocks + 64 * i);
	}
}

/* ------------------------------------------------------------------------------------ */
/* MDEC / BS frame encoder (psxavenc/mdec.c:580-755)                                      */

/* round(n/d), halves away from zero == DIVIDE_ROUNDED's round((double)n/(double)d)
 * (mdec.c:438); exact in integers for d > 0 of either parity. */
static inline int div_round(int n, int d) {
	int a = n < 0 ? -n : n;
	int r = (a + d / 2) / d;
	return n < 0 ? -r : r;
}

/* coeff_clamp_map (mdec.c:260-267): 0x1FF is reserved for the v2 end-of-frame code. */
static inline int clamp_coeff(int v) {
	return v < -0x200 ? -0x200 : (v > 0x1FE ? 0x1FE : v);
}

typedef struct {
	uint8_t *out;
	int limit;      /* frame_max_size */
	int pos;        /* bytes_used */
	uint32_t word;  /* pending 16-bit word, MSB-first */
	int filled;     /* bits in `word` */
	int failed;
} bitpacker_t;

/* flush_bits (mdec.c:321-333): the low byte is stored before the budget check. */
static void pack_flush(bitpacker_t *p) {
	if (p->failed || p->filled == 0)
		return;
	p->out[p->pos++] = (uint8_t)p->word;
	if (p->pos >= p->limit) {
		p->failed = 1;
		return;
	}
	p->out[p->pos++] = (uint8_t)(p->word >> 8);
	p->word = 0;
	p->filled = 0;
}

/* encode_bits (mdec.c:335-385): a full word is only stored when more bits arrive or at the
 * final flush, so `filled` may sit at 16 between calls. */
static void pack_bits(bitpacker_t *p, int nbits, uint32_t value) {
	for (int i = nbits - 1; i >= 0 && !p->failed; i--) {
		if (p->filled == 16)
			pack_flush(p);
		if (p->failed)
			return;
		p->word |= ((value >> i) & 1u) << (15 - p->filled);
		p->filled++;
	}
}

/* Gathers the six 8x8 blocks of every macroblock from the NV21 frame, level-shifts by 128
 * (mdec.c:605-634), runs the FDCT (mdec.c:640) and stores coefficients in scan order, in
 * the order the bitstream visits them: macroblock columns outermost, then rows, then
 * Cr, Cb, Y1..Y4 (mdec.c:689-704). */
static void transform_frame(int variant, int width, int height, const uint8_t *nv21, int16_t *coefs) {
	const uint8_t *luma = nv21;
	const uint8_t *chroma = nv21 + width * height;
	int mbw = width / 16, mbh = height / 16;
	int16_t blk[64];
	int16_t *dst = coefs;

	for (int mx = 0; mx < mbw; mx++) {
		for (int my = 0; my < mbh; my++) {
			for (int k = 0; k < 6; k++) {
				for (int y = 0; y < 8; y++) {
					for (int x = 0; x < 8; x++) {
						int v;
						if (k < 2) {
							/* interleaved plane: Cr at even, Cb at odd bytes (mdec.c:627-628) */
							v = chroma[width * (my * 8 + y) + 2 * (mx * 8 + x) + k];
						} else {
							int ox = ((k - 2) & 1) * 8, oy = ((k - 2) >> 1) * 8;
							v = luma[width * (my * 16 + oy + y) + mx * 16 + ox + x];
						}
						blk[y * 8 + x] = (int16_t)(v - 128);
					}
				}
				if (variant == ORC_FDCT_SSE2)
					orc_fdct_sse2(blk);
				else
					orc_fdct_islow(blk);
				for (int i = 0; i < 64; i++)
					dst[i] = blk[ORC_ZIGZAG[i]];
				dst += 64;
			}
		}
	}
}

/* One attempt at quant scale q (body of the loop at mdec.c:663-722 + encode_dct_block
 * mdec.c:441-510). Returns 1 when the stream fits. */
static int try_quant_scale(int codec, int q, int nblocks, const int16_t *coefs,
                           uint8_t *out, int frame_max_size, int *bytes_used, int *hwords) {
	bitpacker_t bp = {out, frame_max_size, 8, 0, 0, 0};
	int last_dc[3] = {0, 0, 0};
	int uncomp = 0;

	memset(out, 0, frame_max_size);

	for (int b = 0; b < nblocks && !bp.failed; b++) {
		const int16_t *c = coefs + 64 * b;
		int dc = clamp_coeff(div_round(c[0], ORC_QUANT[0] * 8));

		if (codec == ORC_BS_V2) {
			pack_bits(&bp, 10, dc & 0x3FF);
		} else {
			int plane = (b % 6) < 2 ? (b % 6) : 2;   /* Cr, Cb, Y (mdec.c:455-458) */
			int delta = div_round(dc - last_dc[plane], 4);
			last_dc[plane] = (int16_t)(last_dc[plane] + delta * 4);
			if (codec == ORC_BS_V3DC) {           /* wrap-around trick, mdec.c:469-474 */
				if (delta < -0x80)
					delta += 0x100;
				else if (delta > 0x80)
					delta -= 0x100;
			}
			uint32_t e = (plane == 2 ? ORC_DC_VLC_LUMA : ORC_DC_VLC_CHROMA)[delta & 0x1FF];
			pack_bits(&bp, e >> 24, e & 0xFFFFFF);
		}

		int run = 0;
		for (int i = 1; i < 64; i++) {
			int level = clamp_coeff(div_round(c[i], ORC_QUANT_ZZ[i] * q));
			if (level == 0) {
				run++;
				continue;
			}
			int mag = level < 0 ? -level : level;
			uint32_t e = (run < ORC_AC_RUNS && mag < ORC_AC_LEVELS) ? ORC_AC_VLC[run * ORC_AC_LEVELS + mag] : 0;
			if (e)
				pack_bits(&bp, e >> 24, (e & 0xFFFFFF) | (level < 0));
			else
				pack_bits(&bp, ORC_AC_ESCAPE_BITS, (1u << 16) | (run << 10) | (level & 0x3FF));
			run = 0;
			uncomp++;
		}
		pack_bits(&bp, 2, 2);   /* end of block (mdec.c:502) */
		uncomp += 2;
	}

	/* end-of-frame code (mdec.c:645-652, 710) and final flush (mdec.c:716) */
	pack_bits(&bp, 10, codec == ORC_BS_V2 ? 0x1FF : 0x3FF);
	pack_flush(&bp);
	if (bp.failed)
		return 0;

	*bytes_used = bp.pos;
	*hwords = uncomp + 2;
	return 1;
}

int orc_bs_encode_frame(int codec, int fdct_variant, int width, int height,
                        const uint8_t *nv21, int frame_max_size,
                        uint8_t *out, orc_bs_result_t *result) {
	int nblocks = (width / 16) * (height / 16) * 6;
	int16_t *coefs = (int16_t *)malloc((size_t)nblocks * 64 * sizeof(int16_t));
	int bytes = 0, hwords = 0, q;

	transform_frame(fdct_variant, width, height, nv21, coefs);

	for (q = 1; q < 64; q++) {
		if (try_quant_scale(codec, q, nblocks, coefs, out, frame_max_size, &bytes, &hwords))
			break;
	}
	free(coefs);

	result->quant_scale = q;
	if (q >= 64) {
		result->bytes_used = result->blocks_used = result->uncomp_hwords_used = 0;
		return -1;
	}

	/* trailer + header, mdec.c:725-754 */
	hwords = (hwords + 0x3F) & ~0x3F;
	result->uncomp_hwords_used = hwords;
	result->blocks_used = (hwords + 1) >> 1;
	result->bytes_used = (bytes + 3) & ~3;
	out[0] = (uint8_t)result->blocks_used;
	out[1] = (uint8_t)(result->blocks_used >> 8);
	out[2] = 0x00;
	out[3] = 0x38;
	out[4] = (uint8_t)q;
	out[5] = (uint8_t)(q >> 8);
	out[6] = codec == ORC_BS_V2 ? 0x02 : 0x03;
	out[7] = 0x00;
	return 0;
}

int orc_bs_encode_batch(int codec, int fdct_variant, int width, int height, int n,
                        const uint8_t *frames, const int *frame_max_sizes,
                        uint8_t *out, long out_stride, orc_bs_result_t *results) {
	long frame_bytes = (long)width * height * 3 / 2;
	int rc = 0;
	for (int i = 0; i < n; i++) {
		if (orc_bs_encode_frame(codec, fdct_variant, width, height, frames + i * frame_bytes,
		                        frame_max_sizes[i], out + i * out_stride, results + i))
			rc = -1;
	}
	return rc;
}

/* ------------------------------------------------------------------------------------ */
/* SPU / XA ADPCM (libpsxav/adpcm.c)                                                      */

#define UNIT 28

static const int adpcm_k1[5] = {0, 60, 115, 98, 122};   /* adpcm.c:36 */
static const int adpcm_k2[5] = {0, 0, -52, -55, -60};   /* adpcm.c:37 */

static inline int predict(int k1, int k2, int p1, int p2) {
	return (k1 * p1 + k2 * p2 + 32) >> 6;
}

/* find_min_shift (adpcm.c:39-79): open-loop residual range on the RAW samples. */
static int min_shift_for(const orc_adpcm_state_t *st, const int32_t *s, int filter, int range) {
	int p1 = st->prev1, p2 = st->prev2;
	int32_t lo = 0, hi = 0;
	for (int i = 0; i < UNIT; i++) {
		int32_t r = s[i] - predict(adpcm_k1[filter], adpcm_k2[filter], p1, p2);
		if (r < lo) lo = r;
		if (r > hi) hi = r;
		p2 = p1;
		p1 = s[i];
	}
	int rs = 0;
	while (rs < range && (hi >> rs) > (0x7FFF >> range)) rs++;
	while (rs < range && (lo >> rs) < (-0x8000 >> range)) rs++;
	return range - rs;
}

/* attempt_to_encode (adpcm.c:81-140): closed-loop encode of one unit with a fixed
 * filter/shift. Produces the masked codes, the squared error and the decoder state. */
static uint64_t trial_encode(const orc_adpcm_state_t *st, const int32_t *s, int filter, int shift,
                             int range, uint8_t *codes, int *out_p1, int *out_p2) {
	int p1 = st->prev1, p2 = st->prev2;
	int lo = -0x8000 >> range, hi = 0x7FFF >> range;
	uint32_t mask = 0xFFFFu >> range;
	uint64_t err2 = 0;
	for (int i = 0; i < UNIT; i++) {
		int32_t want = s[i] + st->qerr;
		int32_t pred = predict(adpcm_k1[filter], adpcm_k2[filter], p1, p2);
		int32_t e = (int32_t)((uint32_t)(want - pred) << shift);
		e = (e + (1 << (range - 1))) >> range;
		if (e < lo) e = lo;
		if (e > hi) e = hi;
		uint32_t code = (uint32_t)e & mask;
		int32_t dec = (int16_t)(code << range);
		dec = (dec >> shift) + pred;
		if (dec > 0x7FFF) dec = 0x7FFF;
		if (dec < -0x8000) dec = -0x8000;
		int64_t d = (int64_t)dec - want;
		err2 += (uint64_t)d * (uint64_t)d;
		codes[i] = (uint8_t)code;
		p2 = p1;
		p1 = dec;
	}
	*out_p1 = p1;
	*out_p2 = p2;
	return err2;
}

/* encode (adpcm.c:142-191): candidates = filters x {m-1, m, m+1}; strict-less keeps the
 * first minimum (lowest filter, then lowest shift). Returns the header byte and leaves the
 * winner's codes in `codes` and its decoder state in *st. */
static uint8_t encode_unit(orc_adpcm_state_t *st, const int16_t *samples, int limit, int pitch,
                           int filters, int range, uint8_t *codes) {
	int32_t s[UNIT];
	for (int i = 0; i < UNIT; i++)
		s[i] = i < limit ? samples[i * pitch] : 0;

	uint64_t best = (uint64_t)1 << 50;
	int best_f = 0, best_sh = 0, p1, p2;
	uint8_t scratch[UNIT];

	for (int f = 0; f < filters; f++) {
		int m = min_shift_for(st, s, f, range);
		int a = m - 1 < 0 ? 0 : m - 1;
		int b = m + 1 > range ? range : m + 1;
		for (int sh = a; sh <= b; sh++) {
			uint64_t e = trial_encode(st, s, f, sh, range, scratch, &p1, &p2);
			if (e < best) {
				best = e;
				best_f = f;
				best_sh = sh;
			}
		}
	}
	st->mse = trial_encode(st, s, best_f, best_sh, range, codes, &p1, &p2);
	st->prev1 = p1;
	st->prev2 = p2;
	return (uint8_t)((best_sh & 0x0F) | (best_f << 4));
}

int orc_spu_encode(orc_adpcm_state_t *state, const int16_t *samples, int sample_count,
                   int pitch, uint8_t *output) {
	uint8_t codes[UNIT];
	uint8_t *o = output;
	for (int i = 0; i < sample_count; i += UNIT, o += 16) {
		o[0] = encode_unit(state, samples + (long)i * pitch, sample_count - i, pitch, 5, 12, codes);
		o[1] = 0;
		for (int j = 0; j < UNIT; j += 2)
			o[2 + j / 2] = (uint8_t)((codes[j] & 0x0F) | (codes[j + 1] << 4));
	}
	return (int)(o - output);
}

/* encode_block_xa (adpcm.c:193-233): one 128-byte sound group. */
static void xa_sound_group(const int16_t *samples, int limit, uint8_t *group,
                           const orc_xa_settings_t *cfg, orc_adpcm_state_t state[2]) {
	int four = cfg->bits_per_sample == 4;
	int units = four ? 8 : 4;
	int range = four ? 12 : 8;
	uint8_t codes[UNIT];

	for (int u = 0; u < units; u++) {
		int chan = cfg->stereo ? (u & 1) : 0;
		/* stereo: units alternate L/R, pointer advances 56 interleaved samples per pair but
		 * the limit only drops by 28 (adpcm.c:204-211) */
		int step = cfg->stereo ? (u >> 1) : u;
		const int16_t *src = samples + (cfg->stereo ? 56 * step + chan : 28 * step);
		int pitch = cfg->stereo ? 2 : 1;
		uint8_t hdr = encode_unit(&state[chan], src, limit - 28 * step, pitch, 4, range, codes);

		group[four && u >= 4 ? u + 4 : u] = hdr;
		if (four) {
			uint8_t *d = group + 0x10 + (u >> 1);
			int sh = (u & 1) * 4;
			for (int i = 0; i < UNIT; i++)
				d[4 * i] = (uint8_t)((d[4 * i] & ~(0x0F << sh)) | (codes[i] << sh));
		} else {
			uint8_t *d = group + 0x10 + u;
			for (int i = 0; i < UNIT; i++)
				d[4 * i] = codes[i];
		}
	}
}

/* edc_crc32 (cdrom.c:30-41): reflected CRC-32, polynomial 0xD8018001, zero seed. */
uint32_t orc_edc_crc32(const uint8_t *data, int length) {
	uint32_t crc = 0;
	for (int i = 0; i < length; i++) {
		crc ^= data[i];
		for (int b = 0; b < 8; b++)
			crc = (crc >> 1) ^ ((crc & 1) ? 0xD8018001u : 0);
	}
	return crc;
}

static inline uint8_t to_bcd(int v) { return (uint8_t)(v + (v / 10) * 6); }

/* psx_audio_xa_encode (adpcm.c:293-332). `sec` points at the notional start of a 2352-byte
 * sector; in the 2336-byte format the first 16 bytes lie before the caller's buffer and
 * are never touched. */
int orc_xa_encode(const orc_xa_settings_t *cfg, orc_adpcm_state_t state[2],
                  const int16_t *samples, int sample_count, int lba, uint8_t *output) {
	int jump = cfg->bits_per_sample == 8 ? 112 : 224;
	int size = cfg->format == 0 ? 2336 : 2352;
	int total = cfg->stereo ? sample_count * 2 : sample_count;
	int i = 0, j = 0;

	for (; i < total || (j % 18) != 0; i += jump, j++) {
		uint8_t *sec = output + (long)(j / 18) * size - (2352 - size);
		uint8_t *group = sec + 24 + (j % 18) * 128;

		if (j % 18 == 0) {
			/* psx_audio_xa_encode_init_sector (adpcm.c:266-291) */
			uint8_t coding = (uint8_t)((cfg->stereo ? 1 : 0) | (cfg->frequency == 37800 ? 0 : 4) |
			                           (cfg->bits_per_sample == 8 ? 16 : 0));
			if (cfg->format == 1) {
				/* psx_cdrom_init_sector, MODE2_FORM2 (cdrom.c:55-74) */
				int t = lba + 150;
				sec[0] = 0;
				memset(sec + 1, 0xFF, 10);
				sec[11] = 0;
				sec[12] = to_bcd(t / 4500);
				sec[13] = to_bcd((t / 75) % 60);
				sec[14] = to_bcd(t % 75);
				sec[15] = 2;
				sec[19] = 0;
			}
			sec[16] = (uint8_t)cfg->file_number;
			sec[17] = (uint8_t)(cfg->channel_number & 0x1F);
			sec[18] = 0x04 | 0x20 | 0x40;   /* AUDIO | FORM2 | RT */
			sec[19] |= coding;              /* OR into the caller's byte for 2336-byte sectors */
			memcpy(sec + 20, sec + 16, 4);
		}

		xa_sound_group(samples + i, total - i, group, cfg, state);
		memcpy(group + 4, group, 4);
		memcpy(group + 12, group + 8, 4);

		if ((j + 1) % 18 == 0) {
			/* psx_cdrom_calculate_checksums, MODE2_FORM2 (cdrom.c:102-109) */
			uint32_t edc = orc_edc_crc32(sec + 0x10, 0x91C);
			sec[0x92C] = (uint8_t)edc;
			sec[0x92D] = (uint8_t)(edc >> 8);
			sec[0x92E] = (uint8_t)(edc >> 16);
			sec[0x92F] = (uint8_t)(edc >> 24);
			lba++;
		}
	}
	return ((j + 17) / 18) * size;
}

void orc_xa_finalize(const orc_xa_settings_t *cfg, uint8_t *output, int output_length) {
	(void)cfg;
	if (output_length >= 2336) {
		uint8_t *sec = output + output_length - 2352;
		sec[18] |= 0x80;   /* EOF submode bit */
		memcpy(sec + 20, sec + 16, 4);
	}
}
