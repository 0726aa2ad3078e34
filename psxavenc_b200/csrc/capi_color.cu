// C ABI of libpsxav_b200.so, front end (declared in include/psxav_b200.h): decoded pictures ->
// the NV21 frames encode_frame_bs takes (reference: libswscale in psxavenc/decoding.c:286-311,
// 463-475). Host code only; the kernels live in color_convert.cu.
#include <cuda_runtime.h>

#include "psxav_b200.h"
#include "capi_util.h"
#include "color_convert.h"

using namespace psxb200;

extern "C" size_t psxb200_nv21_scratch_bytes(int pixfmt, int n, int src_width, int src_height, int dst_width) {
	if (n <= 0) return 0;
	return (size_t)n * cc_scratch_floats_per_frame(pixfmt, src_width, src_height, dst_width) * sizeof(float);
}

extern "C" int psxb200_nv21_from_device(int pixfmt, int src_full_range, int n, const uint8_t *d_src, size_t src_frame_stride,
                                        int src_width, int src_height, int src_pitch, int dst_width, int dst_height,
                                        uint8_t *d_frames, void *d_scratch, void *stream) {
	if (n <= 0) return 0;
	if (pixfmt < PSXB200_PIX_RGB24 || pixfmt > PSXB200_PIX_YUV420P) return fail("psxb200_nv21_from_device: unknown pixel format %d", pixfmt);
	if (src_width < 2 || src_height < 2 || dst_width < 16 || dst_height < 16 || (dst_width % 16) || (dst_height % 16))
		return fail("psxb200_nv21_from_device: bad size (source >= 2x2, destination multiples of 16: mdec.c:601-602)");
	if (pixfmt == PSXB200_PIX_YUV420P && ((src_width | src_height | src_pitch) & 1))
		return fail("psxb200_nv21_from_device: YUV420P needs even width, height and pitch");
	const int bpp = pixfmt == PSXB200_PIX_YUV420P ? 1 : (pixfmt >= PSXB200_PIX_RGBA ? 4 : 3);
	if (src_pitch < src_width * bpp) return fail("psxb200_nv21_from_device: src_pitch %d < row size", src_pitch);
	if (!d_src || !d_frames || !d_scratch || ((uintptr_t)d_scratch & 7)) return fail("psxb200_nv21_from_device: NULL / misaligned buffer");
	CU_TRY(cc_launch(pixfmt, src_full_range, n, d_src, src_frame_stride, src_width, src_height, src_pitch, dst_width, dst_height,
	                 d_frames, static_cast<float *>(d_scratch), static_cast<cudaStream_t>(stream)));
	g_launches += 2;
	return 0;
}
