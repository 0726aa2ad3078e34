// Internal launch interface of the colour conversion / scaling front end (color_convert.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "psxav_b200.h"

namespace psxb200 {

// float words of scratch per frame
size_t cc_scratch_floats_per_frame(int pixfmt, int src_w, int src_h, int dst_w);

cudaError_t cc_launch(int pixfmt, int full_range, int n, const uint8_t *d_src, size_t src_frame_stride, int src_w, int src_h,
                      int src_pitch, int dst_w, int dst_h, uint8_t *d_frames, float *d_scratch, cudaStream_t stream);

}  // namespace psxb200
