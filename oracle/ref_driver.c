/*
 * ref_driver.c — ctypes-friendly glue compiled TOGETHER WITH the unmodified reference
 * sources (/root/reference/psxavenc/mdec.c, libpsxav/adpcm.c, libpsxav/cdrom.c) into
 * oracle/_ref/libpsxav_ref.so. TEST INFRASTRUCTURE ONLY (checker + CPU baseline).
 *
 * The reference's mdec_encoder_t is caller-owned and its FDCT is whatever FFmpeg's
 * avcodec_dct_init() selects (mdec.c:524,548). `fdct_mode` lets a test pick:
 *   0  libavcodec default (dct_algo AUTO -> ff_fdct_sse2 on this x86-64 build)
 *   1  libavcodec FF_DCT_INT (dct_algo=2 -> ff_jpeg_fdct_islow_8; official release builds)
 *   2  oracle restatement of islow   (orc_fdct_islow)
 *   3  oracle restatement of sse2    (orc_fdct_sse2)
 * With -DREF_NO_LIBAVCODEC (no libavcodec binary available) modes 0/1 map to 2.
 */
#include <stdlib.h>
#include <string.h>
#include "mdec.h"
#include "libpsxav.h"
#include "psx_oracle.h"

#ifdef REF_NO_LIBAVCODEC
AVDCT *avcodec_dct_alloc(void) { return calloc(1, sizeof(AVDCT)); }
int avcodec_dct_init(AVDCT *d) { d->fdct = orc_fdct_islow; return 0; }
void av_free(void *p) { free(p); }
int ref_has_libavcodec(void) { return 0; }
#else
int ref_has_libavcodec(void) { return 1; }
#endif

static void select_fdct(AVDCT *d, int fdct_mode) {
	switch (fdct_mode) {
	case 1:
#ifndef REF_NO_LIBAVCODEC
		d->dct_algo = 2; /* FF_DCT_INT */
		avcodec_dct_init(d);
#endif
		break;
	case 2: d->fdct = orc_fdct_islow; break;
	case 3: d->fdct = orc_fdct_sse2; break;
	default: break;
	}
}

void *ref_bs_open(int codec, int width, int height, int fdct_mode) {
	mdec_encoder_t *enc = calloc(1, sizeof(mdec_encoder_t));
	if (!enc || !init_mdec_encoder(enc, (bs_codec_t)codec, width, height))
		return NULL;
	select_fdct(enc->state.dct_context, fdct_mode);
	return enc;
}

void ref_bs_close(void *h) {
	if (h) {
		destroy_mdec_encoder((mdec_encoder_t *)h);
		free(h);
	}
}

/* res = {bytes_used, blocks_used, quant_scale, uncomp_hwords_used}. The reference aborts
 * (assert, mdec.c:723) when no quant scale fits; callers must avoid such frames. */
void ref_bs_encode(void *h, const uint8_t *frame, int frame_max_size, uint8_t *out, int *res) {
	mdec_encoder_t *enc = (mdec_encoder_t *)h;
	enc->state.frame_output = out;
	enc->state.frame_max_size = frame_max_size;
	encode_frame_bs(enc, frame);
	res[0] = enc->state.bytes_used;
	res[1] = enc->state.blocks_used;
	res[2] = enc->state.quant_scale;
	res[3] = enc->state.uncomp_hwords_used;
}

void ref_bs_encode_batch(void *h, int n, const uint8_t *frames, const int *frame_max_sizes,
                         uint8_t *out, long out_stride, int *res) {
	mdec_encoder_t *enc = (mdec_encoder_t *)h;
	long frame_bytes = (long)enc->video_width * enc->video_height * 3 / 2;
	for (int i = 0; i < n; i++)
		ref_bs_encode(h, frames + i * frame_bytes, frame_max_sizes[i], out + i * out_stride, res + 4 * i);
}

int ref_bs_quant_scale_sum(void *h) { return ((mdec_encoder_t *)h)->state.quant_scale_sum; }

/* Runs the selected FDCT over n blocks (for differential tests of the two models). */
void ref_fdct_blocks(int fdct_mode, int16_t *blocks, int n) {
	AVDCT *d = avcodec_dct_alloc();
	avcodec_dct_init(d);
	select_fdct(d, fdct_mode);
	int16_t *tmp = aligned_alloc(64, 128);
	for (int i = 0; i < n; i++) {
		memcpy(tmp, blocks + 64 * i, 128);
		d->fdct(tmp);
		memcpy(blocks + 64 * i, tmp, 128);
	}
	free(tmp);
	av_free(d);
}

/* encode_sector_str driver (mdec.c:757-836) set up the way encode_file_strspu does for a
 * video-only stream (filefmt.c:546-562): `sectors_per_frame_num/den` is the Bresenham
 * ratio of 2016-byte sectors per frame. Emits `n_sectors` sectors of `sector_size` bytes
 * (2048 for strv with offset 0; 2352 layouts use format STR/STRCD offsets) into out.
 * Returns frames consumed. */
int ref_str_encode(void *h, int format, int video_id, const uint8_t *frames, int n_sectors,
                   int overflow_base, int overflow_den, uint8_t *frame_buf, int frame_buf_size,
                   uint8_t *out, int sector_size, int *bytes_used_per_sector) {
	mdec_encoder_t *enc = (mdec_encoder_t *)h;
	long frame_bytes = (long)enc->video_width * enc->video_height * 3 / 2;
	int used = 0;
	(void)frame_buf_size;
	enc->state.frame_output = frame_buf;
	enc->state.frame_index = 0;
	enc->state.frame_data_offset = 0;
	enc->state.frame_max_size = 0;
	enc->state.frame_block_base_overflow = overflow_base;
	enc->state.frame_block_overflow_num = 0;
	enc->state.frame_block_overflow_den = overflow_den;
	enc->state.quant_scale_sum = 0;
	for (int s = 0; s < n_sectors; s++) {
		used += encode_sector_str(enc, (format_t)format, (uint16_t)video_id,
		                          frames + used * frame_bytes, out + (long)s * sector_size);
		if (bytes_used_per_sector)
			bytes_used_per_sector[s] = enc->state.bytes_used;
	}
	return used;
}

#include <stddef.h>
size_t ref_sizeof_mdec_encoder(void) { return sizeof(mdec_encoder_t); }
