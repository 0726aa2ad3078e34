// Internal launch interface of the ADPCM kernels (adpcm_encode.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace psxb200 {

constexpr int ADPCM_THREADS = 128;

// The launch's i-th chain (i < n_streams) is stream stream_first + i * stream_step of the arrays.
cudaError_t adpcm_launch_spu(int n_streams, const int16_t *d_samples, int pitch, long group_stride, int sample_count,
                             const int *d_counts, void *d_states, uint8_t *d_out, long out_stride, cudaStream_t stream,
                             int stream_first = 0, int stream_step = 1);

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
