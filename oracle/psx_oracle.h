/*
 * psx_oracle — CPU restatement of the psxavenc MDEC/BS + SPU/XA-ADPCM encode core.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is linked into, imported by or executed
 * from the product (psxavenc_b200/). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 * Parity pin: the restatement is checked byte-for-byte against the UNMODIFIED reference
 * sources compiled into oracle/_ref/libpsxav_ref.so (see oracle/Makefile) and against the
 * known-answer hashes of SURVEY.md Appendix B (tests/golden/kat.json). The two FDCT models
 * are checked against the libavcodec 62.11.100 binary bundled in the image.
 */
#ifndef PSX_ORACLE_H
#define PSX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* FDCT variants reachable through FFmpeg's AVDCT.fdct (mdec.c:640); SURVEY.md Appendix A. */
enum { ORC_FDCT_ISLOW = 0, ORC_FDCT_SSE2 = 1 };
/* bs_codec_t values (args.h:60-65). */
enum { ORC_BS_V2 = 0, ORC_BS_V3 = 1, ORC_BS_V3DC = 2 };

void orc_fdct_islow(int16_t *block);
void orc_fdct_sse2(int16_t *block);
void orc_fdct_batch(int variant, int16_t *blocks, int count);

typedef struct {
	int bytes_used;          /* rounded up to a multiple of 4 (mdec.c:736) */
	int blocks_used;         /* (uncomp_hwords+1)>>1 (mdec.c:733) */
	int quant_scale;         /* first q in 1..63 whose stream fits, 64 when none does */
	int uncomp_hwords_used;  /* rounded up to a multiple of 64 (mdec.c:726) */
} orc_bs_result_t;

/* encode_frame_bs (mdec.c:580-755). Returns 0, or -1 when no quant scale fits
 * (the reference asserts there, mdec.c:723). out[0..frame_max_size) is fully written. */
int orc_bs_encode_frame(int codec, int fdct_variant, int width, int height,
                        const uint8_t *nv21, int frame_max_size,
                        uint8_t *out, orc_bs_result_t *result);

/* Same for n frames: frame i at frames + i*1.5*W*H, output at out + i*out_stride. */
int orc_bs_encode_batch(int codec, int fdct_variant, int width, int height, int n,
                        const uint8_t *frames, const int *frame_max_sizes,
                        uint8_t *out, long out_stride, orc_bs_result_t *results);

/* psx_audio_encoder_channel_state_t (libpsxav.h:53-57), same layout. */
typedef struct {
	int qerr;
	uint64_t mse;
	int prev1, prev2;
} orc_adpcm_state_t;

/* psx_audio_spu_encode (adpcm.c:356-376). */
int orc_spu_encode(orc_adpcm_state_t *state, const int16_t *samples, int sample_count,
                   int pitch, uint8_t *output);

/* psx_audio_xa_settings_t (libpsxav.h:44-51) flattened into ints. */
typedef struct {
	int format;          /* 0 = XA (2336-byte sectors), 1 = XACD (2352) */
	int stereo;
	int frequency;       /* 18900 / 37800 */
	int bits_per_sample; /* 4 / 8 */
	int file_number;
	int channel_number;
} orc_xa_settings_t;

/* psx_audio_xa_encode (adpcm.c:293-332); state[0]=left, state[1]=right. */
int orc_xa_encode(const orc_xa_settings_t *settings, orc_adpcm_state_t state[2],
                  const int16_t *samples, int sample_count, int lba, uint8_t *output);
/* psx_audio_xa_encode_finalize (adpcm.c:334-340). */
void orc_xa_finalize(const orc_xa_settings_t *settings, uint8_t *output, int output_length);

/* edc_crc32 (cdrom.c:30-41). */
uint32_t orc_edc_crc32(const uint8_t *data, int length);

/* FNV-1a 64 used by the known-answer vectors. */
uint64_t orc_fnv1a64(const uint8_t *data, long length);

#ifdef __cplusplus
}
#endif
#endif
