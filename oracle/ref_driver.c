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

/* The sector loop of encode_file_str (filefmt.c:391-520) on in-memory inputs, calling the
 * reference's own encode_sector_str, psx_audio_xa_encode, psx_cdrom_init_sector and
 * psx_cdrom_calculate_checksums. init_sector_buffer_video is static in filefmt.c (which needs
 * FFmpeg headers to compile), so its dozen lines (filefmt.c:73-92) are restated here. The
 * reference's sector buffer is uninitialised stack memory (filefmt.c:453); here every sector
 * starts zeroed so that the output is deterministic. Runs until all n_frames frames have been
 * emitted completely; returns the number of sectors written (-1: out of room). audio: n_samples
 * sample frames per channel (0: no audio stream -> interleave 1, filefmt.c:416-421). */
int ref_str_mux(void *h, int format, int video_id, int xa_file, int xa_channel, const uint8_t *frames, int n_frames,
                const int16_t *pcm, int n_samples, int xa_freq, int xa_bits, int xa_stereo, int cd_speed, int fps_num,
                int fps_den, int trailing_audio, uint8_t *frame_buf, uint8_t *out, int max_sectors, int *quant_scale_sum) {
	mdec_encoder_t *enc = (mdec_encoder_t *)h;
	long frame_bytes = (long)enc->video_width * enc->video_height * 3 / 2;
	psx_audio_xa_settings_t xa;
	xa.bits_per_sample = xa_bits;
	xa.frequency = xa_freq;
	xa.stereo = xa_stereo != 0;
	xa.file_number = xa_file;
	xa.channel_number = xa_channel;
	xa.format = format == FORMAT_STRCD ? PSX_AUDIO_XA_FORMAT_XACD : PSX_AUDIO_XA_FORMAT_XA;
	int sector_size = psx_audio_xa_get_buffer_size_per_sector(xa);
	int interleave, audio_samples_per_sector, video_sectors_per_block;
	if (n_samples > 0) {                                             /* filefmt.c:399-415 */
		interleave = psx_audio_xa_get_sector_interleave(xa) * cd_speed;
		audio_samples_per_sector = psx_audio_xa_get_samples_per_sector(xa);
		video_sectors_per_block = interleave - 1;
	} else {
		interleave = 1;
		audio_samples_per_sector = 0;
		video_sectors_per_block = 1;
	}
	psx_audio_encoder_state_t audio_state;
	memset(&audio_state, 0, sizeof(audio_state));
	enc->state.frame_block_base_overflow = (75 * cd_speed) * video_sectors_per_block * fps_den;   /* filefmt.c:428-429 */
	enc->state.frame_block_overflow_den = interleave * fps_num;
	enc->state.frame_output = frame_buf;
	enc->state.frame_index = 0;
	enc->state.frame_data_offset = 0;
	enc->state.frame_max_size = 0;
	enc->state.frame_block_overflow_num = 0;
	enc->state.quant_scale_sum = 0;
	int channels = xa_stereo ? 2 : 1;
	int frames_used = 0, samples_used = 0, sector_count = 0;
	for (; frames_used < n_frames || enc->state.frame_data_offset < enc->state.frame_max_size; sector_count++) {
		if (sector_count >= max_sectors) return -1;
		uint8_t sector[2352];
		memset(sector, 0, sizeof(sector));
		int is_video;
		if (audio_samples_per_sector == 0) is_video = 1;                              /* filefmt.c:456-461 */
		else if (trailing_audio) is_video = (sector_count % interleave) < video_sectors_per_block;
		else is_video = (sector_count % interleave) > 0;
		if (is_video) {
			psx_cdrom_sector_xa_subheader_t *subheader = NULL;                         /* filefmt.c:73-92 */
			if (format == FORMAT_STRCD) {
				psx_cdrom_init_sector((psx_cdrom_sector_t *)sector, sector_count, PSX_CDROM_SECTOR_TYPE_MODE2_FORM1);
				subheader = ((psx_cdrom_sector_t *)sector)->mode2.subheader;
			} else if (format == FORMAT_STR) {
				subheader = (psx_cdrom_sector_xa_subheader_t *)sector;
			}
			if (subheader) {
				subheader->file = xa_file;
				subheader->channel = xa_channel & PSX_CDROM_SECTOR_XA_CHANNEL_MASK;
				subheader->submode = PSX_CDROM_SECTOR_XA_SUBMODE_DATA | PSX_CDROM_SECTOR_XA_SUBMODE_RT;
				subheader->coding = 0;
				memcpy(subheader + 1, subheader, sizeof(psx_cdrom_sector_xa_subheader_t));
			}
			frames_used += encode_sector_str(enc, (format_t)format, (uint16_t)video_id, frames + frames_used * frame_bytes, sector);
			if (format != FORMAT_STRV)
				psx_cdrom_calculate_checksums((psx_cdrom_sector_t *)sector, PSX_CDROM_SECTOR_TYPE_MODE2_FORM1);   /* filefmt.c:474 */
		} else {
			int samples_length = n_samples - samples_used;                             /* filefmt.c:477-494 */
			if (samples_length > audio_samples_per_sector) samples_length = audio_samples_per_sector;
			psx_audio_xa_encode(xa, &audio_state, pcm + (long)samples_used * channels, samples_length, sector_count, sector);
			samples_used += samples_length;
		}
		memcpy(out + (long)sector_count * sector_size, sector, sector_size);
	}
	if (quant_scale_sum) *quant_scale_sum = enc->state.quant_scale_sum;
	return sector_count;
}

#include <stddef.h>
size_t ref_sizeof_mdec_encoder(void) { return sizeof(mdec_encoder_t); }
