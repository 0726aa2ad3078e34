/*
 * Link-level drop-in check: this file is compiled against the REFERENCE's own headers
 * (psxavenc/mdec.h, libpsxav/libpsxav.h, through the avdct ABI stand-in) and linked against
 * libpsxav_b200.so instead of the reference's mdec.c / adpcm.c. It drives the boundary the way
 * the reference's mux loops do: encode_file_sbs (filefmt.c:633-662), encode_file_strspu's
 * video branch (filefmt.c:546-630) and encode_file_spui (filefmt.c:295-362).
 *
 *   dropin_driver sbs   W H codec frame_max_size n_frames  in.nv21 out.bin
 *   dropin_driver strv  W H codec num den n_sectors         in.nv21 out.bin
 *   dropin_driver spui  channels interleave n_samples        in.pcm  out.bin
 *   dropin_driver multi W H codec frame_max_size n_frames n_devices in.nv21 out.bin
 *                       (psxb200_bs_multi_*: one process driving several devices, pinned buffers)
 *   dropin_driver strvbench W H n_frames caller_work_us in.nv21
 *                       (time spent inside encode_sector_str per frame while the caller keeps the
 *                        decoder's two-frame queue and does caller_work_us of its own work per frame)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "mdec.h"
#include "libpsxav.h"
/* the additive psxb200_* layer; the drop-in types come from the reference's headers above */
#define PSXAV_B200_NO_DROPIN_TYPES
#include "psxav_b200.h"

static double now_us(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e6 + ts.tv_nsec / 1e3;
}

static void *slurp(const char *path, long *size) {
	FILE *f = fopen(path, "rb");
	if (!f) { perror(path); exit(2); }
	fseek(f, 0, SEEK_END);
	*size = ftell(f);
	fseek(f, 0, SEEK_SET);
	void *p = calloc(1, *size + (4 << 20));   /* slack behind the data: the decoder's frame queue has a spare slot (decoding.c:448-451) */
	if (fread(p, 1, *size, f) != (size_t)*size) exit(2);
	fclose(f);
	return p;
}

int main(int argc, char **argv) {
	if (argc < 2) return 2;
	long size;
	if (!strcmp(argv[1], "sbs") && argc == 9) {
		int w = atoi(argv[2]), h = atoi(argv[3]), codec = atoi(argv[4]), max = atoi(argv[5]), n = atoi(argv[6]);
		uint8_t *frames = slurp(argv[7], &size);
		FILE *out = fopen(argv[8], "wb");
		mdec_encoder_t enc;                       /* on the stack, not zeroed, like filefmt.c:634 */
		memset(&enc, 0xA5, sizeof(enc));
		if (!init_mdec_encoder(&enc, (bs_codec_t)codec, w, h)) return 3;
		enc.state.frame_output = malloc(max);
		enc.state.frame_max_size = max;
		enc.state.quant_scale_sum = 0;
		for (int i = 0; i < n; i++) {
			encode_frame_bs(&enc, frames + (long)i * w * h * 3 / 2);
			fwrite(enc.state.frame_output, max, 1, out);
		}
		fprintf(stderr, "quant_scale_sum %d\n", enc.state.quant_scale_sum);
		free(enc.state.frame_output);
		destroy_mdec_encoder(&enc);
		fclose(out);
		return 0;
	}
	if (!strcmp(argv[1], "strv") && argc == 10) {
		int w = atoi(argv[2]), h = atoi(argv[3]), codec = atoi(argv[4]);
		int num = atoi(argv[5]), den = atoi(argv[6]), sectors = atoi(argv[7]);
		uint8_t *frames = slurp(argv[8], &size);
		FILE *out = fopen(argv[9], "wb");
		mdec_encoder_t enc;
		memset(&enc, 0xA5, sizeof(enc));
		if (!init_mdec_encoder(&enc, (bs_codec_t)codec, w, h)) return 3;
		enc.state.frame_output = malloc(2016 * ((num + den - 1) / den));
		enc.state.frame_index = 0;
		enc.state.frame_data_offset = 0;
		enc.state.frame_max_size = 0;
		enc.state.frame_block_base_overflow = num;
		enc.state.frame_block_overflow_num = 0;
		enc.state.frame_block_overflow_den = den;
		enc.state.quant_scale_sum = 0;
		long used = 0;
		for (int s = 0; s < sectors; s++) {
			uint8_t sector[2048];
			memset(sector, 0, sizeof(sector));
			used += encode_sector_str(&enc, FORMAT_STRV, 0x8001, frames + used * w * h * 3 / 2, sector);
			fwrite(sector, sizeof(sector), 1, out);
		}
		free(enc.state.frame_output);
		destroy_mdec_encoder(&enc);
		fclose(out);
		return 0;
	}
	if (!strcmp(argv[1], "spui") && argc == 7) {
		int channels = atoi(argv[2]), interleave = atoi(argv[3]), total = atoi(argv[4]);
		int16_t *pcm = slurp(argv[5], &size);
		FILE *out = fopen(argv[6], "wb");
		int per_chunk = interleave / PSX_AUDIO_SPU_BLOCK_SIZE * PSX_AUDIO_SPU_SAMPLES_PER_BLOCK;
		psx_audio_encoder_channel_state_t *st = calloc(channels, sizeof(*st));
		uint8_t *chunk = malloc((size_t)interleave * channels);
		for (int done = 0, k = 0; done < total; k++) {
			int len = total - done < per_chunk ? total - done : per_chunk;
			uint8_t *ptr = chunk;
			memset(chunk, 0, (size_t)interleave * channels);
			if (k == 0) {                          /* leading silent block, filefmt.c:329-332 */
				ptr += PSX_AUDIO_SPU_BLOCK_SIZE;
				len -= PSX_AUDIO_SPU_SAMPLES_PER_BLOCK;
			}
			for (int ch = 0; ch < channels; ch++, ptr += interleave)
				psx_audio_spu_encode(st + ch, pcm + (long)done * channels + ch, len, channels, ptr);
			fwrite(chunk, (size_t)interleave * channels, 1, out);
			done += len;
		}
		fclose(out);
		return 0;
	}
	if (!strcmp(argv[1], "multi") && argc == 10) {
		int w = atoi(argv[2]), h = atoi(argv[3]), codec = atoi(argv[4]), max = atoi(argv[5]), n = atoi(argv[6]), devs = atoi(argv[7]);
		uint8_t *frames = slurp(argv[8], &size);
		size_t frame_bytes = (size_t)w * h * 3 / 2;
		int have = psxb200_device_count();
		if (have < 1 || devs < 1 || devs > 16) return 3;
		int ids[16];
		for (int i = 0; i < devs; i++) ids[i] = i % have;      /* fewer GPUs than asked for: share them */
		psxb200_bs_multi_t *m = psxb200_bs_multi_create(codec, w, h, PSXB200_FDCT_ISLOW, 64, devs, ids);
		if (!m) { fprintf(stderr, "%s\n", psxb200_last_error()); return 3; }
		uint8_t *pin_in = psxb200_pinned_alloc(frame_bytes * n), *pin_out = psxb200_pinned_alloc((size_t)max * n);
		int *sizes = malloc(sizeof(int) * n);
		psxb200_bs_result_t *res = malloc(sizeof(*res) * n);
		if (!pin_in || !pin_out) return 3;
		memcpy(pin_in, frames, frame_bytes * n);
		memset(pin_out, 0xCC, (size_t)max * n);
		for (int i = 0; i < n; i++) sizes[i] = max;
		int failed = psxb200_bs_multi_encode_host(m, n, pin_in, sizes, pin_out, (size_t)max, res);
		if (failed) { fprintf(stderr, "failed %d: %s\n", failed, psxb200_last_error()); return 4; }
		FILE *out = fopen(argv[9], "wb");
		fwrite(pin_out, (size_t)max, n, out);
		fwrite(res, sizeof(*res), n, out);
		fclose(out);
		fprintf(stderr, "devices %d\n", psxb200_bs_multi_device_count(m));
		psxb200_bs_multi_destroy(m);
		psxb200_pinned_free(pin_in);
		psxb200_pinned_free(pin_out);
		return 0;
	}
	if (!strcmp(argv[1], "strvbench") && argc == 7) {
		int w = atoi(argv[2]), h = atoi(argv[3]), n = atoi(argv[4]);
		double work_us = atof(argv[5]);
		uint8_t *frames = slurp(argv[6], &size);
		size_t frame_bytes = (size_t)w * h * 3 / 2;
		int have = (int)(size / (long)frame_bytes);
		for (int pass = 0; pass < 2; pass++) {          /* pass 0: look-ahead on (default), pass 1: off */
			setenv("PSXB200_STR_LOOKAHEAD", pass ? "0" : "1", 1);
			mdec_encoder_t enc;
			if (!init_mdec_encoder(&enc, BS_CODEC_V2, w, h)) return 3;
			enc.state.frame_output = malloc(2016 * 10);
			enc.state.frame_index = 0;
			enc.state.frame_data_offset = 0;
			enc.state.frame_max_size = 0;
			enc.state.frame_block_base_overflow = 150;      /* strv: 10 sectors per frame */
			enc.state.frame_block_overflow_num = 0;
			enc.state.frame_block_overflow_den = 15;
			enc.state.quant_scale_sum = 0;
			uint8_t *queue = calloc(3, frame_bytes);        /* two frames + the spare slot (decoding.c:448-451) */
			memcpy(queue, frames, frame_bytes);
			memcpy(queue + frame_bytes, frames + frame_bytes * (1 % have), frame_bytes);
			int next = 2;
			double inside = 0, t_start = now_us();
			uint8_t sector[2048];
			for (int s = 0; s < n * 10; s++) {
				double t0 = now_us();
				int used = encode_sector_str(&enc, FORMAT_STRV, 0x8001, queue, sector);
				inside += now_us() - t0;
				if (used) {                                  /* retire_av_data + the decoder refilling the queue */
					memmove(queue, queue + frame_bytes, frame_bytes);
					memcpy(queue + frame_bytes, frames + frame_bytes * (next++ % have), frame_bytes);
					double until = now_us() + work_us;      /* the caller's own work per frame (decode, fwrite) */
					while (now_us() < until) {}
				}
			}
			double wall = now_us() - t_start;
			long long hits = 0, misses = 0;
			psxb200_bs_lookahead_stats((const psxb200_bs_encoder_t *)enc.state.dct_context, &hits, &misses);
			printf("encode_sector_str, look-ahead %s, caller work %.0f us/frame: %.1f us/frame inside the library, %.1f us/frame wall (%lld hits, %lld misses, q sum %d)\n",
			       pass ? "off" : "on ", work_us, inside / n, wall / n, hits, misses, enc.state.quant_scale_sum);
			free(queue);
			free(enc.state.frame_output);
			destroy_mdec_encoder(&enc);
		}
		return 0;
	}
	return 2;
}
