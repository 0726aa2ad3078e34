// C ABI of libpsxav_b200.so (declared in include/psxav_b200.h): the batched psxb200_* entry
// points and the drop-in replacements for the reference's codec-core symbols
// (psxavenc/mdec.h:65-74, libpsxav/libpsxav.h:73-101). Host code only; kernels live in
// bs_encode.cu / adpcm_encode.cu. There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "psxav_b200.h"
#include "adpcm_encode.h"
#include "bs_encode.h"

using namespace psxb200;

namespace {

thread_local char g_error[512] = "";
std::atomic<unsigned long long> g_launches{0};

int fail(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_error, sizeof(g_error), fmt, ap);
	va_end(ap);
	return -1;
}

#define CU_TRY(expr)                                                                                  \
	do {                                                                                              \
		cudaError_t e_ = (expr);                                                                      \
		if (e_ != cudaSuccess) return fail("%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

[[noreturn]] void die(const char *what) {
	fprintf(stderr, "libpsxav_b200: %s: %s\n", what, g_error);
	abort();
}

template <typename T>
struct DeviceBuffer {
	T *ptr = nullptr;
	size_t cap = 0;   // elements
	cudaError_t reserve(size_t n) {
		if (n <= cap) return cudaSuccess;
		if (ptr) cudaFree(ptr);
		ptr = nullptr;
		cap = 0;
		cudaError_t e = cudaMalloc(&ptr, n * sizeof(T));
		if (e == cudaSuccess) cap = n;
		return e;
	}
	void release() {
		if (ptr) cudaFree(ptr);
		ptr = nullptr;
		cap = 0;
	}
};

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// ======================================================================================
// BS video
// ======================================================================================

struct psxb200_bs_encoder {
	int codec, width, height, fdct, max_batch, host_chunk, pack_threads, pack_min_ctas = 0 /* 0: by shared-memory fit */, sm_count = 0;
	bool pack_threads_forced = false;   // PSXB200_PACK_THREADS given: no small-batch override
	size_t frame_bytes;
	BsGeometry geo;
	// coefficient planes: [0] serves the device API and host slot 0, [1] host slot 1
	DeviceBuffer<uint4> coefs[2];
	DeviceBuffer<uint32_t> gstream;   // bitstream images for budgets beyond shared memory
	// host-API pipeline slots
	cudaStream_t streams[2] = {nullptr, nullptr};
	DeviceBuffer<uint8_t> in[2], out[2];
	DeviceBuffer<int> sizes[2];
	DeviceBuffer<psxb200_bs_result_t> res[2];
	// optional per-kernel timing (psxb200_bs_timing_*): three events per internal launch pair
	bool timing = false;
	std::vector<cudaEvent_t> events;
	size_t events_used = 0;
	cudaError_t mark(cudaStream_t st) {
		if (events_used == events.size()) {
			cudaEvent_t e;
			cudaError_t rc = cudaEventCreate(&e);
			if (rc != cudaSuccess) return rc;
			events.push_back(e);
		}
		return cudaEventRecord(events[events_used++], st);
	}

	psxb200_bs_encoder(int c, int w, int h, int f, int mb)
		: codec(c), width(w), height(h), fdct(f), max_batch(mb), host_chunk(mb < 256 ? mb : 256), pack_threads(320),
		  frame_bytes((size_t)w * h * 3 / 2), geo(w, h) {}
};

static int bs_pick_threads(const BsGeometry &geo) {
	// 10 warps per CTA: four such CTAs fit an SM at 48 registers per thread (three at 64 when the
	// shared memory does not allow four, see bs_encode_chunked) and the usual frame sizes' groups
	// of 32 blocks divide with <= 5 % idle warp slots (320x240: 57 groups in 6 rounds, 640x480:
	// 225 in 23); measured best on B200 (profiles/r1_sweeps.md).
	const char *env = getenv("PSXB200_PACK_THREADS");
	if (env && atoi(env) >= 32) return std::min(BS_PACK_MAX_THREADS, atoi(env) / 32 * 32);
	return 32 * std::max(1, std::min(10, geo.ngroups));
}

extern "C" int psxb200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

extern "C" const char *psxb200_last_error(void) { return g_error; }
extern "C" unsigned long long psxb200_launch_count(void) { return g_launches.load(); }

extern "C" psxb200_bs_encoder_t *psxb200_bs_create(int codec, int width, int height, int fdct_variant, int max_batch) {
	if (codec < 0 || codec > 2 || width <= 0 || height <= 0 || (width % 16) || (height % 16)) {
		fail("psxb200_bs_create: bad codec/size (codec %d, %dx%d; multiples of 16 required)", codec, width, height);
		return nullptr;
	}
	if (fdct_variant != PSXB200_FDCT_ISLOW && fdct_variant != PSXB200_FDCT_SSE2) {
		fail("psxb200_bs_create: unknown fdct variant %d", fdct_variant);
		return nullptr;
	}
	if (max_batch < 1) max_batch = 1;
	if (psxb200_device_count() == 0) {
		fail("psxb200_bs_create: no CUDA device (this library has no CPU path)");
		return nullptr;
	}
	auto *enc = new psxb200_bs_encoder(codec, width, height, fdct_variant, max_batch);
	enc->pack_threads = bs_pick_threads(enc->geo);
	enc->pack_threads_forced = getenv("PSXB200_PACK_THREADS") != nullptr;
	{
		int dev = 0;
		if (cudaGetDevice(&dev) != cudaSuccess ||
		    cudaDeviceGetAttribute(&enc->sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
			enc->sm_count = 0;
	}
	if (const char *env = getenv("PSXB200_PACK_MIN_CTAS")) enc->pack_min_ctas = atoi(env);
	if (const char *env = getenv("PSXB200_HOST_CHUNK")) enc->host_chunk = std::max(1, std::min(max_batch, atoi(env)));
	bs_upload_tables();
	cudaError_t e = enc->coefs[0].reserve((size_t)max_batch * enc->geo.frame_stride_u4);
	if (e == cudaSuccess) e = cudaGetLastError();
	if (e != cudaSuccess) {
		fail("psxb200_bs_create: %s", cudaGetErrorString(e));
		delete enc;
		return nullptr;
	}
	return enc;
}

extern "C" void psxb200_bs_destroy(psxb200_bs_encoder_t *enc) {
	if (!enc) return;
	for (int i = 0; i < 2; i++) {
		if (enc->streams[i]) {
			cudaStreamSynchronize(enc->streams[i]);
			cudaStreamDestroy(enc->streams[i]);
		}
		enc->coefs[i].release();
		enc->in[i].release();
		enc->out[i].release();
		enc->sizes[i].release();
		enc->res[i].release();
	}
	enc->gstream.release();
	for (cudaEvent_t e : enc->events) cudaEventDestroy(e);
	delete enc;
}

static int bs_encode_chunked(psxb200_bs_encoder *enc, uint4 *d_coefs, int n, const uint8_t *d_frames,
                             const int *d_max_sizes, int max_size_bound, uint8_t *d_out, size_t out_stride,
                             psxb200_bs_result_t *d_results, cudaStream_t stream, const BsStrLayout *str_batch = nullptr) {
	uint32_t *gstream = nullptr;
	size_t gstride = 0;
	// Few frames (the drop-in calls encode one at a time): every CTA has an SM to itself, so the
	// frame's latency is what counts and the widest CTA wins (88 vs 102 us per drop-in frame).
	int threads = enc->pack_threads, min_ctas = enc->pack_min_ctas;
	if (n <= enc->sm_count && !enc->pack_threads_forced) {
		threads = 32 * std::max(1, std::min(BS_PACK_MAX_THREADS / 32, enc->geo.ngroups));
		min_ctas = 1;
	}
	const size_t smem = bs_pack_smem_bytes(enc->codec != 0, true, enc->geo, max_size_bound, threads);
	if (smem > BS_SMEM_BUDGET) {
		gstride = (size_t)(max_size_bound + 3) / 4 + 2;
		CU_TRY(enc->gstream.reserve(gstride * enc->max_batch));
		gstream = enc->gstream.ptr;
	} else if (min_ctas == 0) {
		// Occupancy beats registers here: four 10-warp CTAs per SM at 48 registers (0.560 ms per
		// 4096 frames) against three at 64 (0.603 ms) — when four fit the SM's shared memory
		// (228 KB, 1 KB reserved per CTA); otherwise the 64-register build at three.
		min_ctas = 4 * (smem + 1024) <= 228 * 1024 ? 4 : 3;
	}
	if (min_ctas == 0) min_ctas = 3;
	for (int first = 0; first < n; first += enc->max_batch) {
		int m = std::min(enc->max_batch, n - first);
		if (enc->timing) CU_TRY(enc->mark(stream));
		CU_TRY(bs_launch_dct(enc->fdct, d_frames + (size_t)first * enc->frame_bytes, enc->frame_bytes, m, enc->width,
		                     enc->height, enc->geo, d_coefs, stream));
		if (enc->timing) CU_TRY(enc->mark(stream));
		BsStrLayout str{};
		if (str_batch) {
			str = *str_batch;
			str.frame_index0 += first;   // sector0 stays the batch's: the kernel positions frames absolutely
		}
		CU_TRY(bs_launch_pack(enc->codec, threads, min_ctas, m, d_coefs, enc->geo,
		                      (str_batch || !d_max_sizes) ? nullptr : d_max_sizes + first, max_size_bound,
		                      str_batch ? d_out : d_out + (size_t)first * out_stride, out_stride, d_results + first, gstream,
		                      gstride, str, stream));
		if (enc->timing) CU_TRY(enc->mark(stream));
		g_launches += 2;
	}
	return 0;
}

extern "C" void psxb200_bs_timing_enable(psxb200_bs_encoder_t *enc, int on) {
	enc->timing = on != 0;
	enc->events_used = 0;
}

extern "C" int psxb200_bs_timing_read(psxb200_bs_encoder_t *enc, double *dct_ms, double *pack_ms, int *launch_pairs) {
	double dct = 0, pack = 0;
	int pairs = 0;
	for (size_t i = 0; i + 3 <= enc->events_used; i += 3) {
		float a = 0, b = 0;
		CU_TRY(cudaEventSynchronize(enc->events[i + 2]));
		CU_TRY(cudaEventElapsedTime(&a, enc->events[i], enc->events[i + 1]));
		CU_TRY(cudaEventElapsedTime(&b, enc->events[i + 1], enc->events[i + 2]));
		dct += a;
		pack += b;
		pairs++;
	}
	enc->events_used = 0;
	*dct_ms = dct;
	*pack_ms = pack;
	*launch_pairs = pairs;
	return 0;
}

extern "C" int psxb200_bs_encode_device(psxb200_bs_encoder_t *enc, int n, const uint8_t *d_frames,
                                        const int *d_max_sizes, int max_size_bound, uint8_t *d_out, size_t out_stride,
                                        psxb200_bs_result_t *d_results, void *stream) {
	if (!enc) return fail("psxb200_bs_encode_device: NULL encoder");
	if (n <= 0) return 0;
	if (max_size_bound < 8) return fail("psxb200_bs_encode_device: max_size_bound %d too small", max_size_bound);
	if (((uintptr_t)d_frames & 15) || ((uintptr_t)d_out & 3) || (out_stride & 3) || out_stride < (size_t)max_size_bound)
		return fail("psxb200_bs_encode_device: alignment/stride contract violated");
	return bs_encode_chunked(enc, enc->coefs[0].ptr, n, d_frames, d_max_sizes, max_size_bound, d_out, out_stride,
	                         d_results, static_cast<cudaStream_t>(stream));
}

extern "C" int psxb200_bs_encode_host(psxb200_bs_encoder_t *enc, int n, const uint8_t *h_frames,
                                      const int *h_max_sizes, uint8_t *h_out, size_t out_stride,
                                      psxb200_bs_result_t *h_results) {
	if (!enc) return fail("psxb200_bs_encode_host: NULL encoder");
	if (n <= 0) return 0;
	for (int i = 0; i < 2; i++) {
		if (!enc->streams[i]) CU_TRY(cudaStreamCreateWithFlags(&enc->streams[i], cudaStreamNonBlocking));
	}
	// chunks of host_chunk frames ping-pong between two streams: copy-in of one chunk overlaps
	// the kernels and copy-out of the other
	const int hc = enc->host_chunk;
	CU_TRY(enc->coefs[1].reserve((size_t)hc * enc->geo.frame_stride_u4));

	for (int first = 0, chunk = 0; first < n; first += hc, chunk++) {
		int slot = chunk & 1;
		int m = std::min(hc, n - first);
		cudaStream_t st = enc->streams[slot];
		int bound = 8;
		bool uniform = true;   // one budget for the whole chunk: no per-frame array to upload
		for (int i = 0; i < m; i++) {
			bound = std::max(bound, h_max_sizes[first + i]);
			uniform = uniform && h_max_sizes[first + i] == h_max_sizes[first];
		}
		uniform = uniform && h_max_sizes[first] >= 8;
		if ((size_t)bound > out_stride && n > 1) return fail("psxb200_bs_encode_host: frame_max_size %d > out_stride", bound);
		size_t dstride = round_up((size_t)bound, 16);

		// the slot's previous chunk has fully drained when its stream is idle
		CU_TRY(cudaStreamSynchronize(st));
		CU_TRY(enc->in[slot].reserve((size_t)hc * enc->frame_bytes));
		CU_TRY(enc->out[slot].reserve((size_t)hc * dstride));
		CU_TRY(enc->sizes[slot].reserve(hc));
		CU_TRY(enc->res[slot].reserve(hc));

		CU_TRY(cudaMemcpyAsync(enc->in[slot].ptr, h_frames + (size_t)first * enc->frame_bytes, (size_t)m * enc->frame_bytes,
		                       cudaMemcpyHostToDevice, st));
		if (!uniform)
			CU_TRY(cudaMemcpyAsync(enc->sizes[slot].ptr, h_max_sizes + first, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, st));
		if (bs_encode_chunked(enc, enc->coefs[slot].ptr, m, enc->in[slot].ptr, uniform ? nullptr : enc->sizes[slot].ptr, bound,
		                      enc->out[slot].ptr, dstride, enc->res[slot].ptr, st))
			return -1;
		CU_TRY(cudaMemcpy2DAsync(h_out + (size_t)first * out_stride, std::max(out_stride, (size_t)bound), enc->out[slot].ptr,
		                         dstride, (size_t)bound, m, cudaMemcpyDeviceToHost, st));
		CU_TRY(cudaMemcpyAsync(h_results + first, enc->res[slot].ptr, (size_t)m * sizeof(psxb200_bs_result_t),
		                       cudaMemcpyDeviceToHost, st));
	}
	CU_TRY(cudaStreamSynchronize(enc->streams[0]));
	CU_TRY(cudaStreamSynchronize(enc->streams[1]));
	int failed = 0;
	for (int i = 0; i < n; i++) failed += h_results[i].quant_scale >= 64;
	return failed;
}

// ---- STR video sectors -------------------------------------------------

static int str_layout(psxb200_bs_encoder *enc, int format, int first_frame_index, int sectors_num, int sectors_den,
                      int video_id, BsStrLayout *out) {
	if (first_frame_index < 1 || sectors_num < 1 || sectors_den < 1)
		return fail("psxb200_str_*: first_frame_index, sectors_num and sectors_den must be >= 1");
	BsStrLayout l{};
	// sector size / header offset per container (mdec.c:824-829; filefmt.c:453,502,572,613)
	if (format == FORMAT_STRV) { l.sector_size = 2048; l.header_offset = 0; }
	else if (format == FORMAT_STR) { l.sector_size = 2336; l.header_offset = 8; }
	else if (format == FORMAT_STRCD) { l.sector_size = 2352; l.header_offset = 0x18; }
	else return fail("psxb200_str_*: format must be FORMAT_STR, FORMAT_STRCD or FORMAT_STRV");
	l.frame_index0 = first_frame_index;
	l.sectors_num = sectors_num;
	l.sectors_den = sectors_den;
	l.sector0 = (long long)(first_frame_index - 1) * sectors_num / sectors_den;
	l.video_id = video_id;
	l.width = enc->width;
	l.height = enc->height;
	*out = l;
	return 0;
}

extern "C" long long psxb200_str_sector_count(int n_frames, int first_frame_index, int sectors_num, int sectors_den) {
	long long a = (long long)(first_frame_index - 1) * sectors_num / sectors_den;
	long long b = (long long)(first_frame_index - 1 + n_frames) * sectors_num / sectors_den;
	return b - a;
}

extern "C" int psxb200_str_encode_device(psxb200_bs_encoder_t *enc, int n, const uint8_t *d_frames, int format,
                                         int first_frame_index, int sectors_num, int sectors_den, int video_id,
                                         uint8_t *d_sectors, psxb200_bs_result_t *d_results, void *stream) {
	if (!enc) return fail("psxb200_str_encode_device: NULL encoder");
	if (n <= 0) return 0;
	BsStrLayout l;
	if (str_layout(enc, format, first_frame_index, sectors_num, sectors_den, video_id, &l)) return -1;
	if (((uintptr_t)d_frames & 15) || ((uintptr_t)d_sectors & 3))
		return fail("psxb200_str_encode_device: alignment contract violated (frames 16 bytes, sectors 4 bytes)");
	int bound = 2016 * ((sectors_num + sectors_den - 1) / sectors_den);
	return bs_encode_chunked(enc, enc->coefs[0].ptr, n, d_frames, nullptr, bound, d_sectors, 0, d_results,
	                         static_cast<cudaStream_t>(stream), &l);
}

extern "C" int psxb200_str_encode_host(psxb200_bs_encoder_t *enc, int n, const uint8_t *h_frames, int format,
                                       int first_frame_index, int sectors_num, int sectors_den, int video_id,
                                       uint8_t *h_sectors, psxb200_bs_result_t *h_results) {
	if (!enc) return fail("psxb200_str_encode_host: NULL encoder");
	if (n <= 0) return 0;
	BsStrLayout batch;
	if (str_layout(enc, format, first_frame_index, sectors_num, sectors_den, video_id, &batch)) return -1;
	for (int i = 0; i < 2; i++) {
		if (!enc->streams[i]) CU_TRY(cudaStreamCreateWithFlags(&enc->streams[i], cudaStreamNonBlocking));
	}
	const int hc = enc->host_chunk;
	const int bound = 2016 * ((sectors_num + sectors_den - 1) / sectors_den);
	CU_TRY(enc->coefs[1].reserve((size_t)hc * enc->geo.frame_stride_u4));
	for (int first = 0, chunk = 0; first < n; first += hc, chunk++) {
		int slot = chunk & 1;
		int m = std::min(hc, n - first);
		cudaStream_t st = enc->streams[slot];
		// this chunk as a batch of its own: sectors land at the start of the slot's buffer
		BsStrLayout l;
		if (str_layout(enc, format, first_frame_index + first, sectors_num, sectors_den, video_id, &l)) return -1;
		long long sectors = psxb200_str_sector_count(m, first_frame_index + first, sectors_num, sectors_den);
		long long before = l.sector0 - batch.sector0;
		size_t bytes = (size_t)sectors * l.sector_size;
		CU_TRY(cudaStreamSynchronize(st));
		CU_TRY(enc->in[slot].reserve((size_t)hc * enc->frame_bytes));
		CU_TRY(enc->out[slot].reserve((size_t)(psxb200_str_sector_count(hc, 1, sectors_num, sectors_den) + 2) * l.sector_size));
		CU_TRY(enc->res[slot].reserve(hc));
		CU_TRY(cudaMemcpyAsync(enc->in[slot].ptr, h_frames + (size_t)first * enc->frame_bytes, (size_t)m * enc->frame_bytes,
		                       cudaMemcpyHostToDevice, st));
		// bytes encode_sector_str never writes (outside header + payload) keep the caller's content
		if (l.sector_size != 2048)
			CU_TRY(cudaMemcpyAsync(enc->out[slot].ptr, h_sectors + (size_t)before * l.sector_size, bytes, cudaMemcpyHostToDevice, st));
		if (bs_encode_chunked(enc, enc->coefs[slot].ptr, m, enc->in[slot].ptr, nullptr, bound, enc->out[slot].ptr, 0,
		                      enc->res[slot].ptr, st, &l))
			return -1;
		CU_TRY(cudaMemcpyAsync(h_sectors + (size_t)before * l.sector_size, enc->out[slot].ptr, bytes, cudaMemcpyDeviceToHost, st));
		CU_TRY(cudaMemcpyAsync(h_results + first, enc->res[slot].ptr, (size_t)m * sizeof(psxb200_bs_result_t),
		                       cudaMemcpyDeviceToHost, st));
	}
	CU_TRY(cudaStreamSynchronize(enc->streams[0]));
	CU_TRY(cudaStreamSynchronize(enc->streams[1]));
	int failed = 0;
	for (int i = 0; i < n; i++) failed += h_results[i].quant_scale >= 64;
	return failed;
}

// ---- drop-in: psxavenc/mdec.h ------------------------------------------------------------

static int dropin_fdct_variant() {
	const char *env = getenv("PSXB200_FDCT");
	if (env && (!strcmp(env, "sse2") || !strcmp(env, "SSE2") || !strcmp(env, "1"))) return PSXB200_FDCT_SSE2;
	return PSXB200_FDCT_ISLOW;
}

extern "C" bool init_mdec_encoder(mdec_encoder_t *encoder, bs_codec_t video_codec, int video_width, int video_height) {
	encoder->video_codec = video_codec;
	encoder->video_width = video_width;
	encoder->video_height = video_height;
	mdec_encoder_state_t *state = &encoder->state;
	state->ac_huffman_map = nullptr;
	state->dc_huffman_map = nullptr;
	state->coeff_clamp_map = nullptr;
	for (int i = 0; i < 6; i++) state->dct_block_lists[i] = nullptr;
	state->dct_context = psxb200_bs_create((int)video_codec, video_width, video_height, dropin_fdct_variant(), 16);
	if (!state->dct_context) {
		fprintf(stderr, "libpsxav_b200: init_mdec_encoder: %s\n", g_error);
		return false;
	}
	return true;
}

extern "C" void destroy_mdec_encoder(mdec_encoder_t *encoder) {
	psxb200_bs_destroy(static_cast<psxb200_bs_encoder_t *>(encoder->state.dct_context));
	encoder->state.dct_context = nullptr;
}

extern "C" void encode_frame_bs(mdec_encoder_t *encoder, const uint8_t *video_frame) {
	mdec_encoder_state_t *state = &encoder->state;
	auto *enc = static_cast<psxb200_bs_encoder_t *>(state->dct_context);
	if (!enc) {
		fail("encoder not initialised (init_mdec_encoder failed or was not called)");
		die("encode_frame_bs");
	}
	psxb200_bs_result_t r;
	int max_size = state->frame_max_size;
	int rc = psxb200_bs_encode_host(enc, 1, video_frame, &max_size, state->frame_output, (size_t)max_size, &r);
	if (rc < 0) die("encode_frame_bs");
	if (rc > 0) {
		// the reference aborts here too: assert(state->quant_scale < 64), mdec.c:723
		fail("frame does not fit %d bytes at any quantization scale", max_size);
		die("encode_frame_bs");
	}
	state->quant_scale = r.quant_scale;
	state->quant_scale_sum += r.quant_scale;
	state->uncomp_hwords_used = r.uncomp_hwords_used;
	state->blocks_used = r.blocks_used;
	state->bytes_used = r.bytes_used;
	// scratch fields the reference leaves behind after a successful frame (mdec.c:678-686, 716)
	state->block_type = 0;
	state->bits_value = 0;
	state->bits_left = 16;
}

// STR video sector packer (mdec.c:757-836): whenever the current frame's payload is used up,
// derive the next frame's byte budget from the sectors-per-frame accumulator and encode it;
// then emit one 32-byte sector header plus the next 2016-byte slice of the frame.
extern "C" int encode_sector_str(mdec_encoder_t *encoder, format_t format, uint16_t str_video_id,
                                 const uint8_t *video_frames, uint8_t *output) {
	mdec_encoder_state_t *st = &encoder->state;
	const size_t frame_advance = (size_t)encoder->video_width * encoder->video_height * 2;   // sic, mdec.c:765
	int consumed = 0;

	while (st->frame_data_offset >= st->frame_max_size) {
		st->frame_index++;
		st->frame_block_overflow_num += st->frame_block_base_overflow;
		st->frame_max_size = st->frame_block_overflow_num / st->frame_block_overflow_den * 2016;
		st->frame_block_overflow_num %= st->frame_block_overflow_den;
		st->frame_data_offset = 0;
		encode_frame_bs(encoder, video_frames + consumed * frame_advance);
		consumed++;
	}

	uint8_t hdr[32];
	auto put16 = [&](int at, uint32_t v) { hdr[at] = (uint8_t)v; hdr[at + 1] = (uint8_t)(v >> 8); };
	auto put32 = [&](int at, uint32_t v) { put16(at, v); put16(at + 2, v >> 16); };
	memset(hdr, 0, sizeof(hdr));
	put16(0x00, 0x0160);                                      // STR magic/version
	put16(0x02, str_video_id);                                // chunk type
	put16(0x04, (uint32_t)(st->frame_data_offset / 2016));    // chunk index within the frame
	put16(0x06, (uint32_t)(st->frame_max_size / 2016));       // chunks in the frame
	put32(0x08, (uint32_t)st->frame_index);
	put32(0x0C, (uint32_t)st->bytes_used);
	put16(0x10, (uint32_t)encoder->video_width);
	put16(0x12, (uint32_t)encoder->video_height);
	memcpy(hdr + 0x14, st->frame_output, 8);                  // copy of the BS header

	int at = format == FORMAT_STR ? 0x008 : (format == FORMAT_STRCD ? 0x018 : 0x000);
	memcpy(output + at, hdr, sizeof(hdr));
	memcpy(output + at + 0x020, st->frame_output + st->frame_data_offset, 2016);
	st->frame_data_offset += 2016;
	return consumed;
}

// ======================================================================================
// ADPCM audio
// ======================================================================================

namespace {

// One process-wide context for the host-pointer audio entry points (the reference API has
// no handle to hang it on: libpsxav.h:78-101).
struct AudioContext {
	std::mutex lock;
	cudaStream_t stream = nullptr;
	DeviceBuffer<int16_t> in;
	DeviceBuffer<uint8_t> out;
	DeviceBuffer<uint8_t> states;
	int ensure() {
		if (psxb200_device_count() == 0) return fail("no CUDA device (this library has no CPU path)");
		if (!stream) CU_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
		return 0;
	}
};
AudioContext g_audio;

constexpr size_t STATE_BYTES = 24;   // sizeof(psx_audio_encoder_channel_state_t)

// Number of int16 elements of `samples` that psx_audio_xa_encode reads (adpcm.c:193-233,
// 310-319): in stereo the per-unit limit shrinks by 28 while the pointer advances by 56, so
// the tail group may be read past sample_count*2 (never past its own 224/112 samples).
long xa_input_extent(int stereo, int bits, int sample_count) {
	const int jump = bits == 8 ? 112 : 224;
	const int units = bits == 8 ? 4 : 8;
	const long total = stereo ? 2L * sample_count : sample_count;
	if (!stereo || total <= 0) return total > 0 ? total : 0;
	long extent = 0;
	long j = (total - 1) / jump;   // only the last group holding samples can over-read
	extent = j * jump;             // all earlier groups are read completely
	long remaining = total - j * jump, furthest = 0;
	for (int step = 0; step < units / 2; step++) {
		long lim = std::min<long>(28, remaining - 28L * step);
		if (lim > 0) furthest = std::max(furthest, 56L * step + 2 * lim);
	}
	return extent + furthest;
}

}  // namespace

extern "C" int psxb200_spu_encode_device(int n_streams, const int16_t *d_samples, int pitch, long group_stride,
                                         int sample_count, const int *d_counts, void *d_states, uint8_t *d_out,
                                         long out_stride, void *stream) {
	if (n_streams <= 0) return 0;
	if (pitch < 1 || ((uintptr_t)d_out & 15) || (out_stride & 15) || ((uintptr_t)d_states & 7))
		return fail("psxb200_spu_encode_device: bad pitch or alignment (out: 16 bytes, states: 8 bytes)");
	CU_TRY(adpcm_launch_spu(n_streams, d_samples, pitch, group_stride, sample_count, d_counts, d_states, d_out, out_stride,
	                        static_cast<cudaStream_t>(stream)));
	g_launches += 1;
	return 0;
}

extern "C" int psxb200_spu_encode_host(int n_streams, const int16_t *h_samples, int pitch, long group_stride,
                                       int sample_count, void *h_states, uint8_t *h_out, long out_stride) {
	if (n_streams <= 0 || sample_count <= 0) return 0;
	if (pitch < 1) return fail("psxb200_spu_encode_host: bad pitch");
	std::lock_guard<std::mutex> guard(g_audio.lock);
	if (g_audio.ensure()) return -1;
	cudaStream_t st = g_audio.stream;

	// highest sample index any stream touches (the last group may be partial)
	const int last = n_streams - 1;
	long top = (long)(last / pitch) * group_stride + last % pitch;
	if (last / pitch > 0) top = std::max(top, (long)(last / pitch - 1) * group_stride + pitch - 1);
	const long extent = top + (long)(sample_count - 1) * pitch + 1;
	const long block_bytes = 16L * ((sample_count + 27) / 28);
	const long dstride = block_bytes;   // multiple of 16
	CU_TRY(g_audio.in.reserve((size_t)extent));
	CU_TRY(g_audio.out.reserve((size_t)n_streams * dstride));
	CU_TRY(g_audio.states.reserve((size_t)n_streams * STATE_BYTES));
	CU_TRY(cudaMemcpyAsync(g_audio.in.ptr, h_samples, (size_t)extent * sizeof(int16_t), cudaMemcpyHostToDevice, st));
	CU_TRY(cudaMemcpyAsync(g_audio.states.ptr, h_states, (size_t)n_streams * STATE_BYTES, cudaMemcpyHostToDevice, st));
	CU_TRY(adpcm_launch_spu(n_streams, g_audio.in.ptr, pitch, group_stride, sample_count, nullptr, g_audio.states.ptr,
	                        g_audio.out.ptr, dstride, st));
	g_launches += 1;
	CU_TRY(cudaMemcpy2DAsync(h_out, n_streams == 1 ? (size_t)block_bytes : (size_t)out_stride, g_audio.out.ptr,
	                         (size_t)dstride, (size_t)block_bytes, n_streams, cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaMemcpyAsync(h_states, g_audio.states.ptr, (size_t)n_streams * STATE_BYTES, cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaStreamSynchronize(st));
	return 0;
}

extern "C" int psxb200_xa_encode_device(int n_streams, int format, int stereo, int frequency, int bits_per_sample,
                                        int file_number, int channel_number, const int16_t *d_samples, long in_stride,
                                        int sample_count, int lba, void *d_states, uint8_t *d_out, long out_stride,
                                        void *stream) {
	if (bits_per_sample != 4 && bits_per_sample != 8) return fail("psxb200_xa_encode_device: bits_per_sample must be 4 or 8");
	if (format != 0 && format != 1) return fail("psxb200_xa_encode_device: format must be 0 (XA) or 1 (XACD)");
	int sectors = adpcm_xa_sectors(stereo, bits_per_sample, sample_count);
	int size = format == 0 ? 2336 : 2352;
	if (n_streams <= 0 || sectors == 0) return 0;
	if (((uintptr_t)d_out & 3) || (out_stride & 3) || ((uintptr_t)d_states & 7))
		return fail("psxb200_xa_encode_device: alignment contract violated (out: 4 bytes, states: 8 bytes)");
	CU_TRY(adpcm_launch_xa(n_streams, format, stereo, frequency, bits_per_sample, file_number, channel_number, d_samples,
	                       in_stride, sample_count, lba, d_states, d_out, out_stride, true,
	                       static_cast<cudaStream_t>(stream)));
	g_launches += 2;
	return sectors * size;
}

extern "C" int psxb200_xa_encode_host(int n_streams, int format, int stereo, int frequency, int bits_per_sample,
                                      int file_number, int channel_number, const int16_t *h_samples, long in_stride,
                                      int sample_count, int lba, void *h_states, uint8_t *h_out, long out_stride) {
	if (bits_per_sample != 4 && bits_per_sample != 8) return fail("psxb200_xa_encode_host: bits_per_sample must be 4 or 8");
	if (format != 0 && format != 1) return fail("psxb200_xa_encode_host: format must be 0 (XA) or 1 (XACD)");
	int sectors = adpcm_xa_sectors(stereo, bits_per_sample, sample_count);
	int size = format == 0 ? 2336 : 2352;
	if (n_streams <= 0 || sectors == 0) return 0;
	std::lock_guard<std::mutex> guard(g_audio.lock);
	if (g_audio.ensure()) return -1;
	cudaStream_t st = g_audio.stream;

	const long extent = xa_input_extent(stereo, bits_per_sample, sample_count);
	const long dstride_in = (long)round_up((size_t)extent, 8);
	const long bytes = (long)sectors * size;
	// 16 bytes of slack in front: the 2336-byte format addresses sectors 16 bytes early
	const long dstride_out = (long)round_up((size_t)bytes, 16);
	CU_TRY(g_audio.in.reserve((size_t)n_streams * dstride_in));
	CU_TRY(g_audio.out.reserve((size_t)n_streams * dstride_out + 16));
	CU_TRY(g_audio.states.reserve((size_t)n_streams * 2 * STATE_BYTES));
	uint8_t *d_out = g_audio.out.ptr + 16;
	const size_t h_in_pitch = n_streams == 1 ? (size_t)extent * 2 : (size_t)in_stride * 2;
	const size_t h_out_pitch = n_streams == 1 ? (size_t)bytes : (size_t)out_stride;
	CU_TRY(cudaMemcpy2DAsync(g_audio.in.ptr, (size_t)dstride_in * 2, h_samples, h_in_pitch, (size_t)extent * 2,
	                         n_streams, cudaMemcpyHostToDevice, st));
	// bytes the reference never writes keep the caller's content: round-trip the output buffer
	CU_TRY(cudaMemcpy2DAsync(d_out, (size_t)dstride_out, h_out, h_out_pitch, (size_t)bytes, n_streams,
	                         cudaMemcpyHostToDevice, st));
	CU_TRY(cudaMemcpyAsync(g_audio.states.ptr, h_states, (size_t)n_streams * 2 * STATE_BYTES, cudaMemcpyHostToDevice, st));
	CU_TRY(adpcm_launch_xa(n_streams, format, stereo, frequency, bits_per_sample, file_number, channel_number,
	                       g_audio.in.ptr, dstride_in, sample_count, lba, g_audio.states.ptr, d_out, dstride_out, true, st));
	g_launches += 2;
	CU_TRY(cudaMemcpy2DAsync(h_out, h_out_pitch, d_out, (size_t)dstride_out, (size_t)bytes, n_streams,
	                         cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaMemcpyAsync(h_states, g_audio.states.ptr, (size_t)n_streams * 2 * STATE_BYTES, cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaStreamSynchronize(st));
	return (int)bytes;
}

// ---- drop-in: libpsxav/libpsxav.h ----------------------------------------------------------

extern "C" uint32_t psx_audio_xa_get_buffer_size_per_sector(psx_audio_xa_settings_t settings) {
	return settings.format == PSX_AUDIO_XA_FORMAT_XA ? 2336 : 2352;
}

extern "C" uint32_t psx_audio_xa_get_samples_per_sector(psx_audio_xa_settings_t settings) {
	// 18 sound groups of 224 (4-bit) or 112 (8-bit) samples, split over the channels
	int per_group = settings.bits_per_sample == 8 ? 112 : 224;
	if (settings.stereo) per_group /= 2;
	return (uint32_t)(per_group * 18);
}

extern "C" uint32_t psx_audio_xa_get_buffer_size(psx_audio_xa_settings_t settings, int sample_count) {
	int per_sector = (int)psx_audio_xa_get_samples_per_sector(settings);
	int sectors = (sample_count + per_sector - 1) / per_sector;
	return (uint32_t)sectors * psx_audio_xa_get_buffer_size_per_sector(settings);
}

extern "C" uint32_t psx_audio_spu_get_buffer_size(int sample_count) {
	return (uint32_t)((sample_count + PSX_AUDIO_SPU_SAMPLES_PER_BLOCK - 1) / PSX_AUDIO_SPU_SAMPLES_PER_BLOCK) *
	       PSX_AUDIO_SPU_BLOCK_SIZE;
}

extern "C" uint32_t psx_audio_xa_get_sector_interleave(psx_audio_xa_settings_t settings) {
	// base 2 (stereo) / 4 (mono) at 37800 Hz 8-bit; halving the data rate doubles the gap
	int interleave = settings.stereo ? 2 : 4;
	if (settings.frequency == PSX_AUDIO_XA_FREQ_SINGLE) interleave *= 2;
	if (settings.bits_per_sample == 4) interleave *= 2;
	return (uint32_t)interleave;
}

extern "C" int psx_audio_xa_encode(psx_audio_xa_settings_t settings, psx_audio_encoder_state_t *state,
                                   const int16_t *samples, int sample_count, int lba, uint8_t *output) {
	int n = psxb200_xa_encode_host(1, settings.format == PSX_AUDIO_XA_FORMAT_XA ? 0 : 1, settings.stereo ? 1 : 0,
	                               settings.frequency, settings.bits_per_sample, settings.file_number,
	                               settings.channel_number, samples, 0, sample_count, lba, state, output, 0);
	if (n < 0) die("psx_audio_xa_encode");
	return n;
}

extern "C" void psx_audio_xa_encode_finalize(psx_audio_xa_settings_t settings, uint8_t *output, int output_length) {
	(void)settings;
	if (output_length >= 2336) {
		// subheader of the last sector, addressed as if it were a full 2352-byte sector
		uint8_t *subheader = output + output_length - 2352 + 16;
		subheader[2] |= 0x80;   // end-of-file submode bit
		memcpy(subheader + 4, subheader, 4);
	}
}

extern "C" int psx_audio_xa_encode_simple(psx_audio_xa_settings_t settings, const int16_t *samples, int sample_count,
                                          int lba, uint8_t *output) {
	psx_audio_encoder_state_t state;
	memset(&state, 0, sizeof(state));
	int length = psx_audio_xa_encode(settings, &state, samples, sample_count, lba, output);
	psx_audio_xa_encode_finalize(settings, output, length);
	return length;
}

extern "C" int psx_audio_spu_encode(psx_audio_encoder_channel_state_t *state, const int16_t *samples,
                                    int sample_count, int pitch, uint8_t *output) {
	if (sample_count <= 0) return 0;
	int bytes = (int)psx_audio_spu_get_buffer_size(sample_count);
	if (psxb200_spu_encode_host(1, samples, pitch, 0, sample_count, state, output, bytes) < 0) die("psx_audio_spu_encode");
	return bytes;
}

extern "C" int psx_audio_spu_encode_simple(const int16_t *samples, int sample_count, uint8_t *output, int loop_start) {
	psx_audio_encoder_channel_state_t state;
	memset(&state, 0, sizeof(state));
	int length = psx_audio_spu_encode(&state, samples, sample_count, 1, output);
	if (length < PSX_AUDIO_SPU_BLOCK_SIZE) return length;

	if (loop_start < 0) {
		// one-shot sample: append a silent block that parks the voice (adpcm.c:385-390)
		memset(output + length, 0, PSX_AUDIO_SPU_BLOCK_SIZE);
		output[length + 1] = PSX_AUDIO_SPU_LOOP_TRAP;
		length += PSX_AUDIO_SPU_BLOCK_SIZE;
	} else {
		// looping sample: flag the last block and the block holding the loop point (adpcm.c:391-396)
		output[length - PSX_AUDIO_SPU_BLOCK_SIZE + 1] |= PSX_AUDIO_SPU_LOOP_REPEAT;
		output[loop_start / PSX_AUDIO_SPU_SAMPLES_PER_BLOCK * PSX_AUDIO_SPU_BLOCK_SIZE + 1] |= PSX_AUDIO_SPU_LOOP_START;
	}
	return length;
}
