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
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mdec.h"
#include "libpsxav.h"

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
	return 2;
}
