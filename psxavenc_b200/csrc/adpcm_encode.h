// Internal launch interface of the ADPCM kernels (adpcm_encode.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace psxb200 {

constexpr int ADPCM_THREADS = 128;

cudaError_t adpcm_launch_spu(int n_streams, const int16_t *d_samples, int pitch, long group_stride, int sample_count,
                             const int *d_counts, void *d_states, uint8_t *d_out, long out_stride, cudaStream_t stream);

// sectors produced per stream (adpcm.c:310,331)
int adpcm_xa_sectors(int stereo, int bits_per_sample, int sample_count);

// frame_sectors: also write sync/header/subheader and EDC on the device
cudaError_t adpcm_launch_xa(int n_streams, int format, int stereo, int frequency, int bits_per_sample, int file_number,
                            int channel_number, const int16_t *d_samples, long in_stride, int sample_count, int lba,
                            void *d_states, uint8_t *d_out, long out_stride, bool frame_sectors, cudaStream_t stream);

}  // namespace psxb200
