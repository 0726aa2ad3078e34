/*
 * ABI stand-in for FFmpeg's public <libavcodec/avdct.h> (FFmpeg 8.0, libavcodec 62).
 *
 * TEST INFRASTRUCTURE ONLY. The FFmpeg development headers are not installed in this
 * image, but the reference's psxavenc/mdec.c includes this header (mdec.c:31) and uses
 * avcodec_dct_alloc / avcodec_dct_init / av_free and the `fdct` member (mdec.c:524,548,557,640).
 * This file declares just enough of the public AVDCT struct for the unmodified reference
 * source to compile and link against the libavcodec binary bundled with opencv in the image.
 * Field offsets were checked against that binary:
 *   av_class@0 idct@8 idct_permutation@16 fdct@80 dct_algo@88 idct_algo@92
 *   get_pixels@96 bits_per_sample@104 get_pixels_unaligned@112
 */
#ifndef ORACLE_AVDCT_SHIM_H
#define ORACLE_AVDCT_SHIM_H

#include <stddef.h>
#include <stdint.h>

typedef struct AVDCT {
	const void *av_class;
	void (*idct)(int16_t *block);
	uint8_t idct_permutation[64];
	void (*fdct)(int16_t *block);
	int dct_algo;
	int idct_algo;
	void (*get_pixels)(int16_t *block, const uint8_t *pixels, ptrdiff_t line_size);
	int bits_per_sample;
	void (*get_pixels_unaligned)(int16_t *block, const uint8_t *pixels, ptrdiff_t line_size);
} AVDCT;

AVDCT *avcodec_dct_alloc(void);
int avcodec_dct_init(AVDCT *dsp);
void av_free(void *ptr);

#endif
