// Front end of the video path: what the reference's decoder does with
// libswscale before encode_frame_bs sees a frame (psxavenc/decoding.c:286-311, 463-475) —
// scale the decoded picture to the encoder's size with a bicubic filter and convert it to
// full-range BT.601 NV21 (Y plane + interleaved Cr,Cb plane, pitch = width; mdec.c:593-594).
// Device-resident sources only: a host RGB source would double the bytes per frame on the
// PCIe link the host entry points are bound by.
//
// The filter structure follows libswscale's (probed on the libswscale 9.1 binary in this image;
// the tests carry a numpy restatement of it that is pinned against the binary's outputs):
// RGB sources are converted to YCbCr per source pixel, chroma of horizontally adjacent pixel
// pairs is averaged first (chrSrcW = ceil(W/2)); luma is resampled src -> dst and chroma
// (chrSrcW x srcH for RGB, W/2 x H/2 for YUV420P) -> dst/2 x dst/2 with the Mitchell-Netravali
// cubic B = 0, C = 0.6 (SWS_BICUBIC's default), stretched by the scale ratio when shrinking,
// sample centres aligned ((i + 0.5) * ratio - 0.5), edges replicated. Arithmetic is float32 with
// one final rounding — libswscale works in 15-bit fixed point with quantised coefficients, so the
// results agree within +-1 per sample (about 99 % of the samples exactly), not bit for bit
// (tests/test_gpu_color.py).
//
//   cc_hpass_kernel   per source row: horizontal resampling -> float Y[dst_w], Cr[dst_w/2], Cb[dst_w/2]
//   cc_vpass_kernel   vertical resampling of those rows -> NV21 bytes
#include <cuda_runtime.h>
#include <stdint.h>

#include "color_convert.h"

namespace psxb200 {

__device__ __forceinline__ float cubic06(float t) {
	t = fabsf(t);
	if (t < 1.0f) return (1.4f * t - 2.4f) * t * t + 1.0f;
	if (t < 2.0f) return ((-0.6f * t + 3.0f) * t - 4.8f) * t + 2.4f;
	return 0.0f;
}

// Taps of destination sample i when `src` samples are resampled to `dst`.
struct Taps {
	int first, count;
	float centre, inv_stretch;
	__device__ __forceinline__ Taps(int i, int src, int dst) {
		const float ratio = (float)src / (float)dst;
		const float stretch = fmaxf(ratio, 1.0f);
		centre = ((float)i + 0.5f) * ratio - 0.5f;
		inv_stretch = 1.0f / stretch;
		first = (int)ceilf(centre - 2.0f * stretch);
		count = (int)floorf(centre + 2.0f * stretch) - first + 1;
	}
	__device__ __forceinline__ float weight(int k) const { return cubic06(((float)(first + k) - centre) * inv_stretch); }
};

struct CcParams {
	int pixfmt, full_range;
	int src_w, src_h, src_pitch;
	size_t src_frame_stride;
	int dst_w, dst_h;
	int chr_src_w, chr_src_h;    // chroma samples before resampling
	int pair;                    // RGB: chroma of horizontally adjacent pixel pairs is averaged first
	float chroma_bias;           // added to chroma before the final rounding
};

// full-range BT.601 YCbCr of source luma sample (x, y) / chroma sample (cx, y or cy)
__device__ __forceinline__ float src_luma(const CcParams &p, const uint8_t *fr, int x, int y) {
	if (p.pixfmt == PSXB200_PIX_YUV420P) {
		float v = (float)fr[(size_t)y * p.src_pitch + x];
		return p.full_range ? v : (v - 16.0f) * (255.0f / 219.0f);
	}
	const int bpp = p.pixfmt >= PSXB200_PIX_RGBA ? 4 : 3;
	const uint8_t *px = fr + (size_t)y * p.src_pitch + (size_t)x * bpp;
	const bool bgr = p.pixfmt == PSXB200_PIX_BGR24 || p.pixfmt == PSXB200_PIX_BGRA;
	const float r = (float)px[bgr ? 2 : 0], g = (float)px[1], b = (float)px[bgr ? 0 : 2];
	return 0.299f * r + 0.587f * g + 0.114f * b;
}

__device__ __forceinline__ float2 src_chroma(const CcParams &p, const uint8_t *fr, int cx, int cy) {   // .x = Cr, .y = Cb, centred on 0
	if (p.pixfmt == PSXB200_PIX_YUV420P) {
		const uint8_t *u = fr + (size_t)p.src_pitch * p.src_h;
		const uint8_t *v = u + (size_t)(p.src_pitch / 2) * (p.src_h / 2);
		float cb = (float)u[(size_t)cy * (p.src_pitch / 2) + cx] - 128.0f;
		float cr = (float)v[(size_t)cy * (p.src_pitch / 2) + cx] - 128.0f;
		const float s = p.full_range ? 1.0f : 255.0f / 224.0f;
		return make_float2(cr * s, cb * s);
	}
	const int bpp = p.pixfmt >= PSXB200_PIX_RGBA ? 4 : 3;
	const bool bgr = p.pixfmt == PSXB200_PIX_BGR24 || p.pixfmt == PSXB200_PIX_BGRA;
	const int x0 = p.pair ? 2 * cx : cx, x1 = p.pair ? min(2 * cx + 1, p.src_w - 1) : cx;
	const uint8_t *a = fr + (size_t)cy * p.src_pitch + (size_t)x0 * bpp;
	const uint8_t *c = fr + (size_t)cy * p.src_pitch + (size_t)x1 * bpp;
	const float r = 0.5f * ((float)a[bgr ? 2 : 0] + (float)c[bgr ? 2 : 0]);
	const float g = 0.5f * ((float)a[1] + (float)c[1]);
	const float b = 0.5f * ((float)a[bgr ? 0 : 2] + (float)c[bgr ? 0 : 2]);
	const float y = 0.299f * r + 0.587f * g + 0.114f * b;
	return make_float2((r - y) * (0.5f / (1.0f - 0.299f)), (b - y) * (0.5f / (1.0f - 0.114f)));
}

// scratch per frame: luma rows [src_h][dst_w] floats, then chroma rows [chr_src_h][dst_w/2] float2
__global__ void __launch_bounds__(128)
cc_hpass_kernel(CcParams p, const uint8_t *__restrict__ src, float *__restrict__ scratch, size_t scratch_stride) {
	const int f = blockIdx.z, row = blockIdx.y;
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const uint8_t *fr = src + (size_t)f * p.src_frame_stride;
	float *luma = scratch + (size_t)f * scratch_stride;
	float2 *chroma = reinterpret_cast<float2 *>(luma + (size_t)p.src_h * p.dst_w);
	if (row < p.src_h && x < p.dst_w) {
		const Taps t(x, p.src_w, p.dst_w);
		float acc = 0.0f, norm = 0.0f;
		for (int k = 0; k < t.count; k++) {
			const float w = t.weight(k);
			acc += w * src_luma(p, fr, min(max(t.first + k, 0), p.src_w - 1), row);
			norm += w;
		}
		luma[(size_t)row * p.dst_w + x] = acc / norm;
	}
	const int cw = p.dst_w / 2;
	if (row < p.chr_src_h && x < cw) {
		const Taps t(x, p.chr_src_w, cw);
		float2 acc = make_float2(0.0f, 0.0f);
		float norm = 0.0f;
		for (int k = 0; k < t.count; k++) {
			const float w = t.weight(k);
			const float2 c = src_chroma(p, fr, min(max(t.first + k, 0), p.chr_src_w - 1), row);
			acc.x += w * c.x;
			acc.y += w * c.y;
			norm += w;
		}
		chroma[(size_t)row * cw + x] = make_float2(acc.x / norm, acc.y / norm);
	}
}

__device__ __forceinline__ uint8_t to_byte(float v) { return (uint8_t)min(max(__float2int_rn(v), 0), 255); }

__global__ void __launch_bounds__(128)
cc_vpass_kernel(CcParams p, const float *__restrict__ scratch, size_t scratch_stride, uint8_t *__restrict__ dst) {
	const int f = blockIdx.z, row = blockIdx.y;     // rows 0..dst_h-1: luma; dst_h..dst_h+dst_h/2-1: chroma
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const float *luma = scratch + (size_t)f * scratch_stride;
	const float2 *chroma = reinterpret_cast<const float2 *>(luma + (size_t)p.src_h * p.dst_w);
	uint8_t *out = dst + (size_t)f * ((size_t)p.dst_w * p.dst_h * 3 / 2);
	if (row < p.dst_h) {
		if (x >= p.dst_w) return;
		const Taps t(row, p.src_h, p.dst_h);
		float acc = 0.0f, norm = 0.0f;
		for (int k = 0; k < t.count; k++) {
			const float w = t.weight(k);
			acc += w * luma[(size_t)min(max(t.first + k, 0), p.src_h - 1) * p.dst_w + x];
			norm += w;
		}
		out[(size_t)row * p.dst_w + x] = to_byte(acc / norm);
	} else {
		const int cw = p.dst_w / 2, crow = row - p.dst_h;
		if (x >= cw) return;
		const Taps t(crow, p.chr_src_h, p.dst_h / 2);
		float2 acc = make_float2(0.0f, 0.0f);
		float norm = 0.0f;
		for (int k = 0; k < t.count; k++) {
			const float w = t.weight(k);
			const float2 c = chroma[(size_t)min(max(t.first + k, 0), p.chr_src_h - 1) * cw + x];
			acc.x += w * c.x;
			acc.y += w * c.y;
			norm += w;
		}
		uint8_t *o = out + (size_t)p.dst_w * p.dst_h + (size_t)crow * p.dst_w + 2 * x;
		o[0] = to_byte(acc.x / norm + 128.0f + p.chroma_bias);   // Cr first: NV21 (mdec.c:627-628)
		o[1] = to_byte(acc.y / norm + 128.0f + p.chroma_bias);
	}
}

static CcParams make_params(int pixfmt, int full_range, int src_w, int src_h, int src_pitch, size_t src_frame_stride,
                            int dst_w, int dst_h) {
	CcParams p;
	p.pixfmt = pixfmt;
	p.full_range = full_range;
	p.src_w = src_w;
	p.src_h = src_h;
	p.src_pitch = src_pitch;
	p.src_frame_stride = src_frame_stride;
	p.dst_w = dst_w;
	p.dst_h = dst_h;
	const bool yuv = pixfmt == PSXB200_PIX_YUV420P;
	// libswscale quirks, probed on the 9.1 binary and pinned by tests/golden/swscale_nv21.npz:
	// (1) an unscaled YUV420P source is re-interleaved as is — no range conversion even when the
	//     ranges differ (its planar -> semi-planar special case);
	if (yuv && src_w == dst_w && src_h == dst_h) p.full_range = 1;
	// (2) RGB chroma is taken from averaged pixel pairs only while that leaves at least as many
	//     chroma samples as the destination has;
	p.pair = !yuv && dst_w / 2 <= src_w / 2;
	// (3) chroma that went through its limited -> full range expansion (every RGB source, limited
	//     range YUV) comes out truncated rather than rounded.
	p.chroma_bias = (yuv && p.full_range) ? 0.0f : -0.5f;
	p.chr_src_w = yuv ? src_w / 2 : (p.pair ? (src_w + 1) / 2 : src_w);
	p.chr_src_h = yuv ? src_h / 2 : src_h;
	return p;
}

size_t cc_scratch_floats_per_frame(int pixfmt, int src_w, int src_h, int dst_w) {
	const int chr_h = pixfmt == PSXB200_PIX_YUV420P ? src_h / 2 : src_h;
	(void)src_w;
	return (size_t)src_h * dst_w + (size_t)chr_h * (dst_w / 2) * 2;
}

cudaError_t cc_launch(int pixfmt, int full_range, int n, const uint8_t *d_src, size_t src_frame_stride, int src_w, int src_h,
                      int src_pitch, int dst_w, int dst_h, uint8_t *d_frames, float *d_scratch, cudaStream_t stream) {
	const CcParams p = make_params(pixfmt, full_range, src_w, src_h, src_pitch, src_frame_stride, dst_w, dst_h);
	const size_t stride = cc_scratch_floats_per_frame(pixfmt, src_w, src_h, dst_w);
	for (int first = 0; first < n; first += 32768) {    // gridDim.z limit
		const int m = n - first < 32768 ? n - first : 32768;
		const dim3 block(128);
		const dim3 grid_h((dst_w + 127) / 128, src_h, m);
		cc_hpass_kernel<<<grid_h, block, 0, stream>>>(p, d_src + (size_t)first * src_frame_stride, d_scratch + (size_t)first * stride, stride);
		const dim3 grid_v((dst_w + 127) / 128, dst_h + dst_h / 2, m);
		cc_vpass_kernel<<<grid_v, block, 0, stream>>>(p, d_scratch + (size_t)first * stride, stride,
		                                              d_frames + (size_t)first * ((size_t)dst_w * dst_h * 3 / 2));
	}
	return cudaGetLastError();
}

}  // namespace psxb200
