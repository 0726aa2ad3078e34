// Internal launch interface of the ADPCM kernels (adpcm_encode.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace psxb200 {

constexpr int ADPCM_THREADS = 128;

struct ChannelState {   // psx_audio_encoder_channel_state_t (libpsxav.h:53-57)
	int qerr;
	int pad_;
	unsigned long long mse;
	int prev1, prev2;
};

// A short SPU chain handed over in the kernel's parameters (adpcm_launch_spu_small)
constexpr int SPU_SMALL_SAMPLES = 4 * 28;
struct SpuSmallCall {
	int16_t samples[SPU_SMALL_SAMPLES];   // the chain's samples, gathered (pitch 1)
	ChannelState state;                   // incoming state
	int count;                            // samples, <= SPU_SMALL_SAMPLES
	uint32_t seq;                         // value written to *flag when everything else has been written
	// device addresses of mapped host memory:
	uint8_t *out;                         // 16 bytes per 28 samples
	ChannelState *state_out;              // outgoing state
	volatile uint32_t *flag;              // completion flag
};

// The launch's i-th chain (i < n_streams) is stream stream_first + i * stream_step of the arrays.
cudaError_t adpcm_launch_spu(int n_streams, const int16_t *d_samples, int pitch, long group_stride, int sample_count,
                             const int *d_counts, void *d_states, uint8_t *d_out, long out_stride, cudaStream_t stream,
                             int stream_first = 0, int stream_step = 1);

// One chain of at most SPU_SMALL_SAMPLES samples: one launch of one warp, nothing read from memory.
cudaError_t adpcm_launch_spu_small(const SpuSmallCall &call, cudaStream_t stream);

// int16 elements of a stream's input that the reference reads (and so must be readable)
long adpcm_xa_input_extent(int stereo, int bits_per_sample, int sample_count);

// sectors produced per stream (adpcm.c:310,331)
int adpcm_xa_sectors(int stereo, int bits_per_sample, int sample_count);

// Stream s writes sector k at d_out + s * out_stride + k * sector_stride (sector_stride <= 0: back
// to back) and numbers it lba + k * lba_step. d_edc_tables (edc_tables_device()) non-NULL: also
// write sync/header/subheader and the EDC on the device; NULL: sound groups only.
cudaError_t adpcm_launch_xa(int n_streams, int format, int stereo, int frequency, int bits_per_sample, int file_number,
                            int channel_number, const int16_t *d_samples, long in_stride, int sample_count, int lba,
                            int lba_step, void *d_states, uint8_t *d_out, long out_stride, long sector_stride,
                            const uint32_t *d_edc_tables, cudaStream_t stream);

}  // namespace psxb200
