/*
 * psxav_b200.h — C ABI of libpsxav_b200.so, the B200 (sm_100a) MDEC/BS + SPU/XA-ADPCM encode
 * core. Plain C: pointers, ints and sizes only.
 *
 * Two layers:
 *
 *  1. DROP-IN symbols with the reference's exact names, signatures and struct layouts, so
 *     psxavenc's container/muxing layer (psxavenc/filefmt.c, libpsxav/cdrom.c) links
 *     against this library instead of psxavenc/mdec.c + libpsxav/adpcm.c without edits.
 *     Each declaration cites the reference declaration it replaces.
 *
 *  2. BATCHED entry points (psxb200_*), additive: many frames / many independent ADPCM
 *     streams per call, device-resident or host buffers. The drop-in symbols are thin
 *     wrappers over these with a batch of one.
 *
 * There is no CPU fallback: every entry point needs a CUDA device and aborts with a message
 * on a CUDA error (the reference API has no error channel: encode_frame_bs is void,
 * mdec.h:67; the audio functions return byte counts, libpsxav.h:79-101).
 *
 * All file:line citations are relative to the reference tree (WonderfulToolchain/psxavenc).
 */
#ifndef PSXAV_B200_H
#define PSXAV_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ===================================================================================== */
/* 1. Drop-in layer                                                                       */
/* ===================================================================================== */

#ifndef PSXAV_B200_NO_DROPIN_TYPES

/* psxavenc/args.h:45-57 */
typedef enum {
	FORMAT_INVALID = -1,
	FORMAT_XA, FORMAT_XACD, FORMAT_SPU, FORMAT_VAG, FORMAT_SPUI, FORMAT_VAGI,
	FORMAT_STR, FORMAT_STRCD, FORMAT_STRSPU, FORMAT_STRV, FORMAT_SBS
} format_t;

/* psxavenc/args.h:60-65 */
typedef enum {
	BS_CODEC_INVALID = -1,
	BS_CODEC_V2, BS_CODEC_V3, BS_CODEC_V3DC
} bs_codec_t;

/* psxavenc/mdec.h:32-55. Same field order, types and offsets. The caller owns the struct
 * (on its stack, not zeroed: filefmt.c:424,546,634) and sets frame_output, frame_max_size,
 * frame_index, frame_data_offset, the overflow accumulators and quant_scale_sum itself
 * after init (filefmt.c:428-440, 637-640). The five pointer fields at the end are private
 * to the encoder in the reference as well; this library keeps its GPU context in
 * `dct_context` (an AVDCT* in the reference) and leaves the others NULL. */
typedef struct {
	int frame_index;
	int frame_data_offset;
	int frame_max_size;
	int frame_block_base_overflow;
	int frame_block_overflow_num;
	int frame_block_overflow_den;
	int block_type;
	int16_t last_dc_values[3];
	uint16_t bits_value;
	int bits_left;
	uint8_t *frame_output;
	int bytes_used;
	int blocks_used;
	int uncomp_hwords_used;
	int quant_scale;
	int quant_scale_sum;

	void *dct_context;          /* reference: AVDCT*; here: psxb200_bs_encoder_t* */
	uint32_t *ac_huffman_map;   /* unused (NULL) */
	uint32_t *dc_huffman_map;   /* unused (NULL) */
	int16_t *coeff_clamp_map;   /* unused (NULL) */
	int16_t *dct_block_lists[6];/* unused (NULL) */
} mdec_encoder_state_t;

/* psxavenc/mdec.h:57-63 */
typedef struct {
	bs_codec_t video_codec;
	int video_width;
	int video_height;
	mdec_encoder_state_t state;
} mdec_encoder_t;

/* psxavenc/mdec.h:65 (mdec.c:512). Allocates the GPU context; false on failure. */
bool init_mdec_encoder(mdec_encoder_t *encoder, bs_codec_t video_codec, int video_width, int video_height);
/* psxavenc/mdec.h:66 (mdec.c:553) */
void destroy_mdec_encoder(mdec_encoder_t *encoder);
/* psxavenc/mdec.h:67 (mdec.c:580). video_frame: one NV21 frame in HOST memory. Writes
 * state.frame_output[0..frame_max_size), bytes_used, blocks_used, uncomp_hwords_used,
 * quant_scale and adds to quant_scale_sum. Aborts when no quant scale fits (mdec.c:723). */
void encode_frame_bs(mdec_encoder_t *encoder, const uint8_t *video_frame);
/* psxavenc/mdec.h:68-74 (mdec.c:757). One 2016-byte slice + 32-byte STR header per call;
 * returns the number of frames consumed from video_frames. */
int encode_sector_str(mdec_encoder_t *encoder, format_t format, uint16_t str_video_id,
                      const uint8_t *video_frames, uint8_t *output);

/* libpsxav/libpsxav.h:31-32 */
#define PSX_AUDIO_SPU_BLOCK_SIZE        16
#define PSX_AUDIO_SPU_SAMPLES_PER_BLOCK 28
/* libpsxav/libpsxav.h:34-37 */
enum { PSX_AUDIO_XA_FREQ_SINGLE = 18900, PSX_AUDIO_XA_FREQ_DOUBLE = 37800 };
/* libpsxav/libpsxav.h:39-42 */
typedef enum { PSX_AUDIO_XA_FORMAT_XA, PSX_AUDIO_XA_FORMAT_XACD } psx_audio_xa_format_t;
/* libpsxav/libpsxav.h:44-51 */
typedef struct {
	psx_audio_xa_format_t format;
	bool stereo;
	int frequency;
	int bits_per_sample;
	int file_number;
	int channel_number;
} psx_audio_xa_settings_t;
/* libpsxav/libpsxav.h:53-57 */
typedef struct {
	int qerr;
	uint64_t mse;
	int prev1, prev2;
} psx_audio_encoder_channel_state_t;
/* libpsxav/libpsxav.h:59-62 */
typedef struct {
	psx_audio_encoder_channel_state_t left;
	psx_audio_encoder_channel_state_t right;
} psx_audio_encoder_state_t;
/* libpsxav/libpsxav.h:64-71 */
enum {
	PSX_AUDIO_SPU_LOOP_END    = (1 << 0),
	PSX_AUDIO_SPU_LOOP_REPEAT = (1 << 0) | (1 << 1),
	PSX_AUDIO_SPU_LOOP_START  = (1 << 1) | (1 << 2),
	PSX_AUDIO_SPU_LOOP_TRAP   = (1 << 0) | (1 << 2)
};

/* libpsxav/libpsxav.h:73-77 (adpcm.c:235-260): pure size arithmetic, host only. */
uint32_t psx_audio_xa_get_buffer_size(psx_audio_xa_settings_t settings, int sample_count);
uint32_t psx_audio_spu_get_buffer_size(int sample_count);
uint32_t psx_audio_xa_get_buffer_size_per_sector(psx_audio_xa_settings_t settings);
uint32_t psx_audio_xa_get_samples_per_sector(psx_audio_xa_settings_t settings);
uint32_t psx_audio_xa_get_sector_interleave(psx_audio_xa_settings_t settings);
/* libpsxav/libpsxav.h:78-85 (adpcm.c:293). HOST pointers; the ADPCM search, the sector
 * framing (adpcm.c:262-332, cdrom.c:55-74) and the EDC (cdrom.c:30-41,102-109) all run on
 * the GPU. Bytes the reference leaves untouched keep the caller's content. */
int psx_audio_xa_encode(psx_audio_xa_settings_t settings, psx_audio_encoder_state_t *state,
                        const int16_t *samples, int sample_count, int lba, uint8_t *output);
/* libpsxav/libpsxav.h:86-92 (adpcm.c:342) */
int psx_audio_xa_encode_simple(psx_audio_xa_settings_t settings, const int16_t *samples,
                               int sample_count, int lba, uint8_t *output);
/* libpsxav/libpsxav.h:93-99 (adpcm.c:356) */
int psx_audio_spu_encode(psx_audio_encoder_channel_state_t *state, const int16_t *samples,
                         int sample_count, int pitch, uint8_t *output);
/* libpsxav/libpsxav.h:100 (adpcm.c:378) */
int psx_audio_spu_encode_simple(const int16_t *samples, int sample_count, uint8_t *output, int loop_start);
/* libpsxav/libpsxav.h:101 (adpcm.c:334) */
void psx_audio_xa_encode_finalize(psx_audio_xa_settings_t settings, uint8_t *output, int output_length);

#endif /* PSXAV_B200_NO_DROPIN_TYPES */

/* ===================================================================================== */
/* 2. Batched layer                                                                       */
/* ===================================================================================== */

/* Which FFmpeg AVDCT.fdct the output must match bit-for-bit:
 * ISLOW = ff_jpeg_fdct_islow_8 (the reference's official release binaries, non-x86 builds),
 * SSE2  = ff_fdct_sse2 (a stock SIMD-enabled x86-64 FFmpeg). Default for the drop-in layer:
 * ISLOW, override with the environment variable PSXB200_FDCT=sse2. */
enum { PSXB200_FDCT_ISLOW = 0, PSXB200_FDCT_SSE2 = 1 };

typedef struct {
	int bytes_used;          /* mdec.c:736; 0 when the frame failed */
	int blocks_used;         /* mdec.c:733 */
	int quant_scale;         /* 1..63; 64 = no quant scale fits (reference asserts, mdec.c:723) */
	int uncomp_hwords_used;  /* mdec.c:726 */
} psxb200_bs_result_t;

typedef struct psxb200_bs_encoder psxb200_bs_encoder_t;

/* Returns the CUDA device count (0 = none), without touching any device. */
int psxb200_device_count(void);
/* Last error message of the calling thread ("" if none). */
const char *psxb200_last_error(void);

/* codec: 0 = BS v2, 1 = v3, 2 = v3dc (bs_codec_t). width/height multiples of 16
 * (mdec.c:601-602). max_batch: frames per internal launch (scratch is sized for it).
 * The encoder belongs to the CUDA device that is current at creation; every entry point makes
 * that device current for the duration of the call. NULL on failure. */
psxb200_bs_encoder_t *psxb200_bs_create(int codec, int width, int height, int fdct_variant, int max_batch);
void psxb200_bs_destroy(psxb200_bs_encoder_t *enc);
/* The encoder's CUDA device / the size of one NV21 input frame (1.5 * width * height). */
int psxb200_bs_device(const psxb200_bs_encoder_t *enc);
long long psxb200_bs_frame_bytes(const psxb200_bs_encoder_t *enc);

/* Page-locked host memory usable from every device (cudaHostAlloc, portable): host buffers
 * allocated here are copied from / to directly by the DMA engines. */
void *psxb200_pinned_alloc(size_t bytes);
void psxb200_pinned_free(void *p);

/* Device-resident batch: d_frames = n NV21 frames back to back (1.5*W*H bytes each, base
 * 16-byte aligned), d_max_sizes[n] = per-frame byte budgets (frame_max_size), each
 * <= max_size_bound, or NULL when every frame's budget is max_size_bound; only bytes
 * [0, frame_max_size) of a frame's buffer are defined; d_out = n bitstream buffers out_stride bytes apart (out_stride and
 * base multiples of 4, out_stride >= max_size_bound); d_results[n]. Asynchronous on
 * `stream` (a cudaStream_t, NULL = default stream). Returns 0, or -1 (see
 * psxb200_last_error). Frames for which no quant scale fits get quant_scale 64,
 * bytes_used 0 and an all-zero buffer. */
int psxb200_bs_encode_device(psxb200_bs_encoder_t *enc, int n, const uint8_t *d_frames,
                             const int *d_max_sizes, int max_size_bound, uint8_t *d_out,
                             size_t out_stride, psxb200_bs_result_t *d_results, void *stream);

/* Same with HOST buffers: copies in, encodes and copies out in pipelined chunks on the
 * encoder's own streams; synchronous. Pinned host memory (psxb200_pinned_alloc, cudaHostAlloc)
 * is used in place; pageable memory goes through the driver's staging. h_max_sizes[n] is
 * required (one budget per frame, each <= out_stride). Only the bytes a frame's stream occupies
 * travel back over the bus; the rest of [0, frame_max_size) is zero-filled on the host (the
 * reference clears the whole buffer, mdec.c:676), and bytes at and beyond a frame's own
 * frame_max_size are never written. Returns the number of frames that failed (0 = all good)
 * or -1 on error. */
int psxb200_bs_encode_host(psxb200_bs_encoder_t *enc, int n, const uint8_t *h_frames,
                           const int *h_max_sizes, uint8_t *h_out, size_t out_stride,
                           psxb200_bs_result_t *h_results);

/* STR video sectors straight from the GPU: what calling
 * encode_sector_str (mdec.c:757-836) once per sector produces for a video-only stream driven as
 * encode_file_strspu does (filefmt.c:546-630). Frame k of the call has frame_index
 * first_frame_index + k (1-based, mdec.c:769) and the byte budget
 * 2016 * (floor(K*num/den) - floor((K-1)*num/den)), K its frame_index — the closed form of the
 * overflow accumulator (mdec.c:772-774) with num = frame_block_base_overflow,
 * den = frame_block_overflow_den. Each frame becomes budget/2016 sectors: the 32-byte STR
 * header at the format's offset (FORMAT_STRV: 2048-byte sectors, offset 0; FORMAT_STR: 2336,
 * offset 8; FORMAT_STRCD: 2352, offset 0x18; mdec.c:824-829) followed by a 2016-byte slice. Other
 * bytes of a sector are not written (as in the reference; the caller's mux adds subheaders
 * and EDC, filefmt.c:463-475). psxb200_str_sector_count gives the number of sectors. */
long long psxb200_str_sector_count(int n_frames, int first_frame_index, int sectors_num, int sectors_den);
int psxb200_str_encode_device(psxb200_bs_encoder_t *enc, int n, const uint8_t *d_frames, int format,
                              int first_frame_index, int sectors_num, int sectors_den, int video_id,
                              uint8_t *d_sectors, psxb200_bs_result_t *d_results, void *stream);
int psxb200_str_encode_host(psxb200_bs_encoder_t *enc, int n, const uint8_t *h_frames, int format,
                            int first_frame_index, int sectors_num, int sectors_den, int video_id,
                            uint8_t *h_sectors, psxb200_bs_result_t *h_results);

/* The general form of the two calls above: STR / STRCD / STRV video sectors for one or many
 * independent files, optionally complete (sector framing + EDC) and optionally placed at their
 * LBA slots of the muxed file image. All fields must be set (memset the struct to 0 first).
 *
 *   framing       also write what the reference's mux loop adds around encode_sector_str:
 *                 init_sector_buffer_video (filefmt.c:73-92: FORMAT_STRCD sync + BCD timecode +
 *                 mode 2 + doubled subheader, cdrom.c:55-74; FORMAT_STR the doubled subheader)
 *                 and psx_cdrom_calculate_checksums(.., MODE2_FORM1) (filefmt.c:474,
 *                 cdrom.c:92-100): EDC of bytes [0x10, 0x818) of the buffer at 0x818. For
 *                 FORMAT_STR the reference applies that to the 2336-byte buffer, i.e. shifted
 *                 by 16 bytes against the sector's layout and over 16 bytes it never writes —
 *                 reproduced as is, over whatever the output buffer holds there.
 *   interleave    sectors per mux block: 1 = video only, N = one XA audio sector + N-1 video
 *                 sectors (psx_audio_xa_get_sector_interleave * cd speed, filefmt.c:401-403);
 *                 trailing_audio = FLAG_STR_TRAILING_AUDIO (filefmt.c:456-461). Decides the
 *                 LBA (timecode) of each video sector.
 *   place_at_lba  0: the video sectors of a file follow each other in the output, the first
 *                 one of the call at byte 0; 1: video sector at LBA a sits at byte
 *                 (a - lba_origin) * sector_size, audio slots are left alone.
 *   frames_per_file / file_stride
 *                 0: all n frames are one file. F: the batch is n / F files of F frames, each
 *                 starting over at first_frame_index, file f's output at f * file_stride bytes.
 */
typedef struct {
	int format;              /* FORMAT_STR, FORMAT_STRCD or FORMAT_STRV */
	int first_frame_index;   /* frame_index of a file's first frame in this call, >= 1 (mdec.c:769) */
	int sectors_num;         /* frame_block_base_overflow (filefmt.c:428) */
	int sectors_den;         /* frame_block_overflow_den (filefmt.c:429) */
	int video_id;            /* str_video_id */
	int framing;
	int xa_file, xa_channel; /* subheader file / channel (filefmt.c:84-85) */
	int interleave;
	int trailing_audio;
	int place_at_lba;
	long long lba_origin;
	int frames_per_file;
	long long file_stride;   /* bytes, multiple of 4 */
} psxb200_str_params_t;

/* Slots (sectors) of a file's output region taken by n_frames frames starting at
 * params->first_frame_index: [*first_slot, *end_slot) relative to byte 0 of the region. */
int psxb200_str_slot_range(const psxb200_str_params_t *params, int n_frames, long long *first_slot, long long *end_slot);
int psxb200_str_encode_device_ex(psxb200_bs_encoder_t *enc, int n, const uint8_t *d_frames,
                                 const psxb200_str_params_t *params, uint8_t *d_sectors,
                                 psxb200_bs_result_t *d_results, void *stream);
/* Host buffers; n must be a multiple of frames_per_file when that is set. Bytes of a sector
 * the reference does not write keep the caller's content. */
int psxb200_str_encode_host_ex(psxb200_bs_encoder_t *enc, int n, const uint8_t *h_frames,
                               const psxb200_str_params_t *params, uint8_t *h_sectors,
                               psxb200_bs_result_t *h_results);

/* encode_file_str (filefmt.c:391-520) on the GPU for n_files independent inputs: file f =
 * frames_per_file NV21 frames at h_frames + f * frames_per_file * frame_bytes and
 * samples_per_file XA sample frames (interleaved L,R when stereo) at h_pcm + f * pcm_stride
 * (int16 units), muxed into the image h_images + f * image_stride: slot s is an XA sector when
 * s % interleave == 0 (trailing_audio: == interleave - 1), else the file's next video sector;
 * complete sectors (framing, EDC), frame_index from 1, LBA = slot. params->format, sectors_num,
 * sectors_den, video_id, xa_file, xa_channel, interleave, trailing_audio are used; h_pcm NULL
 * or samples_per_file 0: video only (interleave 1). The video and the XA kernels of a group of
 * files run concurrently on two streams. Bytes neither encoder writes (the ECC area, cdrom.c:98)
 * are zero; psx_audio_xa_encode_finalize is left to the caller. h_xa_states: n_files
 * psx_audio_encoder_state_t in/out, or NULL (zero state in, final state dropped).
 * psxb200_strcd_image_bytes: size of one image. Returns failed frames or -1. */
long long psxb200_strcd_image_bytes(const psxb200_str_params_t *params, int frames_per_file, int xa_bits, int xa_stereo,
                                    int samples_per_file);
int psxb200_strcd_encode_host(psxb200_bs_encoder_t *enc, int n_files, int frames_per_file, const uint8_t *h_frames,
                              const psxb200_str_params_t *params, int xa_frequency, int xa_bits, int xa_stereo,
                              const int16_t *h_pcm, long pcm_stride, int samples_per_file, void *h_xa_states,
                              uint8_t *h_images, long long image_stride, psxb200_bs_result_t *h_results);

/* ---- one process, many devices ----------------------------------------------------------
 * The reference's host is one single-threaded C process (filefmt.c); these entry points are how
 * it reaches every GPU of the box: one encoder and one worker thread per device, frames dealt
 * out in contiguous ranges (they are independent, mdec.c:676-686), every device's results
 * written straight into the caller's host arrays. No collective is involved. device_ids NULL:
 * devices 0 .. n_devices-1; n_devices <= 0: all visible devices. Host buffers should be pinned
 * (psxb200_pinned_alloc). Same contracts and return values as the single-device calls. */
typedef struct psxb200_bs_multi psxb200_bs_multi_t;
psxb200_bs_multi_t *psxb200_bs_multi_create(int codec, int width, int height, int fdct_variant, int max_batch,
                                            int n_devices, const int *device_ids);
void psxb200_bs_multi_destroy(psxb200_bs_multi_t *m);
int psxb200_bs_multi_device_count(const psxb200_bs_multi_t *m);
int psxb200_bs_multi_encode_host(psxb200_bs_multi_t *m, int n, const uint8_t *h_frames, const int *h_max_sizes,
                                 uint8_t *h_out, size_t out_stride, psxb200_bs_result_t *h_results);
int psxb200_bs_multi_str_encode_host(psxb200_bs_multi_t *m, int n, const uint8_t *h_frames,
                                     const psxb200_str_params_t *params, uint8_t *h_sectors,
                                     psxb200_bs_result_t *h_results);
int psxb200_bs_multi_strcd_encode_host(psxb200_bs_multi_t *m, int n_files, int frames_per_file, const uint8_t *h_frames,
                                       const psxb200_str_params_t *params, int xa_frequency, int xa_bits, int xa_stereo,
                                       const int16_t *h_pcm, long pcm_stride, int samples_per_file, void *h_xa_states,
                                       uint8_t *h_images, long long image_stride, psxb200_bs_result_t *h_results);

/* Look-ahead statistics of the drop-in encode_sector_str: frames served
 * from a speculative encode of the frame behind the previous one / frames encoded on demand. */
void psxb200_bs_lookahead_stats(const psxb200_bs_encoder_t *enc, long long *hits, long long *misses);

/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
unsigned long long psxb200_launch_count(void);

/* Optional per-kernel timing for benchmarks: while enabled, every internal launch pair
 * (FDCT kernel, pack kernel) of psxb200_bs_encode_device is bracketed by CUDA events on the
 * launching stream. psxb200_bs_timing_read waits for them, returns the summed durations in
 * milliseconds and the number of launch pairs, and resets the counters. */
void psxb200_bs_timing_enable(psxb200_bs_encoder_t *enc, int on);
int psxb200_bs_timing_read(psxb200_bs_encoder_t *enc, double *dct_ms, double *pack_ms, int *launch_pairs);

/* SPU-ADPCM, n_streams independent mono chains (adpcm.c:356-376 each).
 * Stream s reads sample i at d_samples[(s / pitch) * group_stride + (s % pitch) + i * pitch]
 * i.e. `pitch` interleaved channels per group of streams, groups group_stride samples
 * apart (pitch 1: stream s starts at s * group_stride). sample_count samples per stream
 * (d_counts[s] overrides when non-NULL). d_states[n_streams]: in/out, layout of
 * psx_audio_encoder_channel_state_t. Output: 16 bytes per 28 samples at
 * d_out + s * out_stride. Asynchronous on `stream`. Returns 0 / -1. */
int psxb200_spu_encode_device(int n_streams, const int16_t *d_samples, int pitch, long group_stride,
                              int sample_count, const int *d_counts, void *d_states,
                              uint8_t *d_out, long out_stride, void *stream);
int psxb200_spu_encode_host(int n_streams, const int16_t *h_samples, int pitch, long group_stride,
                            int sample_count, void *h_states, uint8_t *h_out, long out_stride);
/* The same over several devices of this process (see psxb200_bs_multi_*): whole groups of
 * `pitch` interleaved chains per device when there are at least as many groups as devices
 * (vagi x B), else chain c on device c mod G (one vagi file; a chain is strictly sequential,
 * adpcm.c:135-136,186-190, so a single file gains nothing beyond overlap of its channels). */
int psxb200_spu_encode_host_multi(int n_devices, const int *device_ids, int n_streams, const int16_t *h_samples,
                                  int pitch, long group_stride, int sample_count, void *h_states, uint8_t *h_out,
                                  long out_stride);

/* XA-ADPCM, n_streams independent XA streams (adpcm.c:293-332 each, incl. subheaders,
 * sound-group header duplication and EDC). Stream s reads sample_count per-channel frames
 * (interleaved L,R when stereo) at d_samples + s * in_stride (in int16 units; readable up
 * to the end of the last 224/112-sample sound group the reference would touch) and writes
 * sectors of 2336 (format 0) or 2352 (format 1) bytes at d_out + s * out_stride, first
 * sector numbered lba. d_states[n_streams][2] (left, right) in/out. Bytes the reference never
 * writes keep the buffer's content: the coding byte of format 0 is OR-ed into what is there
 * (adpcm.c:278-288) and bytes 8-15 of an 8-bit sound group are copied 8-11 -> 12-15 from it
 * (adpcm.c:322) — start from zeroed buffers for deterministic output, as for the reference.
 * Returns bytes per stream (>= 0) or -1. */
int psxb200_xa_encode_device(int n_streams, int format, int stereo, int frequency, int bits_per_sample,
                             int file_number, int channel_number, const int16_t *d_samples,
                             long in_stride, int sample_count, int lba, void *d_states,
                             uint8_t *d_out, long out_stride, void *stream);
int psxb200_xa_encode_host(int n_streams, int format, int stereo, int frequency, int bits_per_sample,
                           int file_number, int channel_number, const int16_t *h_samples,
                           long in_stride, int sample_count, int lba, void *h_states,
                           uint8_t *h_out, long out_stride);
/* psxb200_xa_encode_device with the sectors of a stream sector_stride bytes apart (0: back to
 * back) and numbered lba + k * lba_step — the XA sectors of a muxed .str image sit in every
 * interleave-th slot (filefmt.c:456-461, 487-494). */
int psxb200_xa_encode_device_ex(int n_streams, int format, int stereo, int frequency, int bits_per_sample,
                                int file_number, int channel_number, const int16_t *d_samples,
                                long in_stride, int sample_count, int lba, int lba_step, void *d_states,
                                uint8_t *d_out, long out_stride, long sector_stride, void *stream);
/* Streams dealt out in contiguous runs over several devices of this process. */
int psxb200_xa_encode_host_multi(int n_devices, const int *device_ids, int n_streams, int format, int stereo,
                                 int frequency, int bits_per_sample, int file_number, int channel_number,
                                 const int16_t *h_samples, long in_stride, int sample_count, int lba, void *h_states,
                                 uint8_t *h_out, long out_stride);

/* ---- front end: decoded pictures -> NV21 --------------------------
 * What the reference's decoder does with libswscale before the encoder sees a frame
 * (psxavenc/decoding.c:286-311, 463-475): bicubic scaling to the encoder's size and conversion
 * to full-range BT.601 NV21 (Y plane + interleaved Cr,Cb plane). DEVICE pointers only — the
 * point is to feed frames that are already in HBM (a decoder's output) without a PCIe round
 * trip; a host RGB source would double the bytes per frame on the link the host entry points
 * are bound by. Same filter structure as libswscale's (cubic B=0 C=0.6 stretched by the scale
 * ratio, chroma of RGB pixel pairs averaged first, edges replicated; an unscaled YUV420P source
 * is re-interleaved without range conversion, as libswscale does), evaluated in float32:
 * results match libswscale within +-1 per sample (tests/test_gpu_color.py), not bit for bit.
 *
 * d_src: n pictures src_frame_stride bytes apart. RGB24/BGR24/RGBA/BGRA: packed rows of src_pitch
 * bytes. YUV420P: Y plane (src_pitch x src_height), then U and V planes (src_pitch/2 x
 * src_height/2); src_full_range 0 = limited ("MPEG") range, expanded to full range. d_frames:
 * n NV21 frames of 1.5 * dst_width * dst_height bytes. d_scratch: psxb200_nv21_scratch_bytes
 * bytes, 8-byte aligned. Asynchronous on `stream`. Returns 0 / -1. */
enum { PSXB200_PIX_RGB24 = 0, PSXB200_PIX_BGR24 = 1, PSXB200_PIX_RGBA = 2, PSXB200_PIX_BGRA = 3, PSXB200_PIX_YUV420P = 4 };
size_t psxb200_nv21_scratch_bytes(int pixfmt, int n, int src_width, int src_height, int dst_width);
int psxb200_nv21_from_device(int pixfmt, int src_full_range, int n, const uint8_t *d_src, size_t src_frame_stride,
                             int src_width, int src_height, int src_pitch, int dst_width, int dst_height,
                             uint8_t *d_frames, void *d_scratch, void *stream);

#ifdef __cplusplus
}
#endif
#endif
