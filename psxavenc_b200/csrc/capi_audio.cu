// C ABI of libpsxav_b200.so, audio half (declared in include/psxav_b200.h): the batched
// psxb200_spu_* / psxb200_xa_* entry points and the drop-in replacements for the reference's
// ADPCM symbols (libpsxav/libpsxav.h:73-101). Host code only; the kernels live in
// adpcm_encode.cu. There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>

#include "psxav_b200.h"
#include "adpcm_encode.h"
#include "bs_encode.h"
#include "capi_util.h"

using namespace psxb200;

namespace {

constexpr size_t STATE_BYTES = 24;   // sizeof(psx_audio_encoder_channel_state_t)
constexpr int MAX_AUDIO_DEVICES = 64;
// SPU calls whose input fits this many bytes skip the copy engine: the kernel reads the samples
// and states from, and writes the blocks to, page-locked host memory mapped into the device
// address space (the reference's own callers pass one 28-sample block, filefmt.c:243).
constexpr size_t ZERO_COPY_BYTES = 16 * 1024;

// Context of the host-pointer audio entry points, one per device (the reference API has no
// handle to hang it on: libpsxav.h:78-101).
constexpr int SPU_PIPE_STREAMS = 4;   // chunks of a large SPU call in flight (copy in | kernel | copy out)

struct AudioContext {
	std::mutex lock;
	cudaStream_t stream = nullptr;
	cudaStream_t pipe[SPU_PIPE_STREAMS] = {};   // pipe[0] is `stream`
	DeviceBuffer<int16_t> in;
	DeviceBuffer<uint8_t> out;
	DeviceBuffer<uint8_t> states;
	PinnedBuffer<uint8_t> stage;   // zero-copy staging: samples | states | blocks
	PinnedBuffer<uint8_t> small;   // results of the parameter-borne small SPU calls: blocks | state | flag
	uint32_t seq = 0;              // completion-flag value of the last small call
	const uint32_t *edc = nullptr;
	int ensure() {
		if (!stream) CU_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
		pipe[0] = stream;
		for (int i = 1; i < SPU_PIPE_STREAMS; i++)
			if (!pipe[i]) CU_TRY(cudaStreamCreateWithFlags(&pipe[i], cudaStreamNonBlocking));
		if (!edc) {
			edc = edc_tables_device();
			if (!edc) return fail("EDC tables unavailable: %s", cudaGetErrorString(cudaGetLastError()));
		}
		return 0;
	}
};
AudioContext g_audio[MAX_AUDIO_DEVICES];

// the context of the calling thread's current device
AudioContext *audio_context() {
	int dev = 0;
	if (cudaGetDeviceCount(&dev) != cudaSuccess || dev == 0) {
		cudaGetLastError();
		fail("no CUDA device (this library has no CPU path)");
		return nullptr;
	}
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_AUDIO_DEVICES) {
		fail("cudaGetDevice: %s", cudaGetErrorString(cudaGetLastError()));
		return nullptr;
	}
	return &g_audio[dev];
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

// Streams first, first + step, ... of the caller's arrays on the calling thread's current device
// (step 1, first 0: all of them). The multi-device entry deals the chains of one interleaved
// group out this way (channel c on device c mod G).
int psxb200::spu_encode_host_subset(int n_streams, int first, int step, const int16_t *h_samples, int pitch, long group_stride,
                                    int sample_count, void *h_states, uint8_t *h_out, long out_stride) {
	if (n_streams <= 0 || sample_count <= 0 || first >= n_streams) return 0;
	if (pitch < 1 || step < 1 || first < 0) return fail("psxb200_spu_encode_host: bad pitch / subset");
	AudioContext *ctx = audio_context();
	if (!ctx) return -1;
	std::lock_guard<std::mutex> guard(ctx->lock);
	if (ctx->ensure()) return -1;
	cudaStream_t st = ctx->stream;
	StreamDrain drain(st);
	const long block_bytes = 16L * ((sample_count + 27) / 28);
	const int n_sub = (n_streams - first + step - 1) / step;

	if (n_streams == 1 && sample_count <= SPU_SMALL_SAMPLES) {
		// one block or a few (the reference's own encode_file_spu hands over 28 samples per call,
		// filefmt.c:243): samples and state ride in the kernel's parameters, the kernel writes blocks,
		// state and a completion flag into mapped host memory, and the call returns when the flag
		// shows — one launch, no copies, no stream synchronisation
		constexpr size_t OUT = 0, STATE = 64, FLAG = 96, BYTES = 128;
		if (!ctx->small.ptr) {
			CU_TRY(ctx->small.reserve(BYTES));
			memset(ctx->small.ptr, 0, BYTES);
		}
		uint8_t *d_base = ctx->small.device_ptr();
		if (!d_base) return fail("psxb200_spu_encode_host: mapped staging unavailable");
		SpuSmallCall call;
		for (int i = 0; i < sample_count; i++) call.samples[i] = h_samples[(long)i * pitch];
		memcpy(&call.state, h_states, STATE_BYTES);
		call.count = sample_count;
		if (++ctx->seq == 0) ctx->seq = 1;
		call.seq = ctx->seq;
		call.out = d_base + OUT;
		call.state_out = reinterpret_cast<ChannelState *>(d_base + STATE);
		call.flag = reinterpret_cast<volatile uint32_t *>(d_base + FLAG);
		CU_TRY(adpcm_launch_spu_small(call, st));
		g_launches += 1;
		const volatile uint32_t *flag = reinterpret_cast<const volatile uint32_t *>(ctx->small.ptr + FLAG);
		const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(20);
		bool seen = false;
		for (unsigned spins = 0; !(seen = *flag == call.seq); spins++)
			if ((spins & 0xFFF) == 0xFFF && std::chrono::steady_clock::now() > deadline) break;
		if (!seen) CU_TRY(cudaStreamSynchronize(st));   // a failed launch or a wedged device surfaces here
		std::atomic_thread_fence(std::memory_order_acquire);
		drain.armed = false;
		memcpy(h_out, ctx->small.ptr + OUT, (size_t)block_bytes);
		memcpy(h_states, ctx->small.ptr + STATE, STATE_BYTES);
		return 0;
	}

	if (n_streams == 1 && (size_t)sample_count * 2 <= ZERO_COPY_BYTES) {
		// one short chain (the drop-in's usual call): gather its samples into the mapped staging
		// buffer, one launch, one wait — no copy-engine round trips
		const size_t in_bytes = round_up((size_t)sample_count * 2, 16);
		CU_TRY(ctx->stage.reserve(in_bytes + 32 + (size_t)block_bytes));
		int16_t *s_in = reinterpret_cast<int16_t *>(ctx->stage.ptr);
		uint8_t *s_state = ctx->stage.ptr + in_bytes;
		uint8_t *s_out = ctx->stage.ptr + in_bytes + 32;
		if (pitch == 1) memcpy(s_in, h_samples, (size_t)sample_count * 2);
		else for (int i = 0; i < sample_count; i++) s_in[i] = h_samples[(long)i * pitch];
		memcpy(s_state, h_states, STATE_BYTES);
		uint8_t *d_base = ctx->stage.device_ptr();
		if (!d_base) return fail("psxb200_spu_encode_host: mapped staging unavailable");
		CU_TRY(adpcm_launch_spu(1, reinterpret_cast<const int16_t *>(d_base), 1, 0, sample_count, nullptr, d_base + in_bytes,
		                        d_base + in_bytes + 32, block_bytes, st));
		g_launches += 1;
		CU_TRY(cudaStreamSynchronize(st));
		drain.armed = false;
		memcpy(h_out, s_out, (size_t)block_bytes);
		memcpy(h_states, s_state, STATE_BYTES);
		return 0;
	}

	// highest sample index any stream touches (the last group may be partial)
	const int last = n_streams - 1;
	long top = (long)(last / pitch) * group_stride + last % pitch;
	if (last / pitch > 0) top = std::max(top, (long)(last / pitch - 1) * group_stride + pitch - 1);
	const long extent = top + (long)(sample_count - 1) * pitch + 1;
	const long dstride = block_bytes;   // multiple of 16
	const size_t h_pitch = n_streams == 1 ? (size_t)block_bytes : (size_t)out_stride;
	uint8_t *hs = static_cast<uint8_t *>(h_states);
	CU_TRY(ctx->in.reserve((size_t)extent));
	CU_TRY(ctx->out.reserve((size_t)n_streams * dstride));
	CU_TRY(ctx->states.reserve((size_t)n_streams * STATE_BYTES));

	// Many whole interleave groups that do not overlap in memory (vagi x B): the call is cut into
	// runs of groups that rotate over SPU_PIPE_STREAMS streams, so that the copy-in of a run
	// overlaps the kernel and the copy-out of the runs before it. A chain is sequential in time
	// (adpcm.c:135-136, 186-190), its kernel takes ~45 ns per sample however few chains a launch
	// holds; about a thousand chains per run keep the host->device link busy meanwhile.
	const int n_groups = (n_streams + pitch - 1) / pitch;
	if (first == 0 && step == 1 && n_groups >= 2 && group_stride >= (long)pitch * sample_count) {
		const long chain_bytes = (long)sample_count * 2;
		const long run_chains = std::max(1024L, (4L << 20) / chain_bytes);
		const int run_groups = (int)std::max(1L, (run_chains + pitch - 1) / pitch);
		if (n_groups >= 2 * run_groups) {
			static_assert(SPU_PIPE_STREAMS == 4, "one StreamDrain per extra stream below");
			StreamDrain drains[SPU_PIPE_STREAMS - 1] = {StreamDrain(ctx->pipe[1]), StreamDrain(ctx->pipe[2]), StreamDrain(ctx->pipe[3])};
			int k = 0;
			for (int g0 = 0; g0 < n_groups; g0 += run_groups, k++) {
				const int g1 = std::min(n_groups, g0 + run_groups);
				const int s0 = g0 * pitch, s1 = std::min(n_streams, g1 * pitch);
				const long lo = (long)g0 * group_stride;
				const long hi = g1 == n_groups ? extent : (long)g1 * group_stride;
				cudaStream_t ps = ctx->pipe[k % SPU_PIPE_STREAMS];
				CU_TRY(cudaMemcpyAsync(ctx->in.ptr + lo, h_samples + lo, (size_t)(hi - lo) * sizeof(int16_t), cudaMemcpyHostToDevice, ps));
				CU_TRY(cudaMemcpyAsync(ctx->states.ptr + (size_t)s0 * STATE_BYTES, hs + (size_t)s0 * STATE_BYTES,
				                       (size_t)(s1 - s0) * STATE_BYTES, cudaMemcpyHostToDevice, ps));
				CU_TRY(adpcm_launch_spu(s1 - s0, ctx->in.ptr + lo, pitch, group_stride, sample_count, nullptr,
				                        ctx->states.ptr + (size_t)s0 * STATE_BYTES, ctx->out.ptr + (size_t)s0 * dstride, dstride, ps));
				g_launches += 1;
				CU_TRY(cudaMemcpy2DAsync(h_out + (size_t)s0 * h_pitch, h_pitch, ctx->out.ptr + (size_t)s0 * dstride, (size_t)dstride,
				                         (size_t)block_bytes, (size_t)(s1 - s0), cudaMemcpyDeviceToHost, ps));
				CU_TRY(cudaMemcpyAsync(hs + (size_t)s0 * STATE_BYTES, ctx->states.ptr + (size_t)s0 * STATE_BYTES,
				                       (size_t)(s1 - s0) * STATE_BYTES, cudaMemcpyDeviceToHost, ps));
			}
			for (int i = 0; i < SPU_PIPE_STREAMS; i++) CU_TRY(cudaStreamSynchronize(ctx->pipe[i]));
			drain.armed = false;
			for (auto &d : drains) d.armed = false;
			return 0;
		}
	}

	CU_TRY(cudaMemcpyAsync(ctx->in.ptr, h_samples, (size_t)extent * sizeof(int16_t), cudaMemcpyHostToDevice, st));
	CU_TRY(cudaMemcpy2DAsync(ctx->states.ptr + (size_t)first * STATE_BYTES, (size_t)step * STATE_BYTES, hs + (size_t)first * STATE_BYTES,
	                         (size_t)step * STATE_BYTES, STATE_BYTES, n_sub, cudaMemcpyHostToDevice, st));
	CU_TRY(adpcm_launch_spu(n_sub, ctx->in.ptr, pitch, group_stride, sample_count, nullptr, ctx->states.ptr,
	                        ctx->out.ptr, dstride, st, first, step));
	g_launches += 1;
	CU_TRY(cudaMemcpy2DAsync(h_out + (size_t)first * h_pitch, (size_t)step * h_pitch, ctx->out.ptr + (size_t)first * dstride,
	                         (size_t)step * dstride, (size_t)block_bytes, n_sub, cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaMemcpy2DAsync(hs + (size_t)first * STATE_BYTES, (size_t)step * STATE_BYTES, ctx->states.ptr + (size_t)first * STATE_BYTES,
	                         (size_t)step * STATE_BYTES, STATE_BYTES, n_sub, cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaStreamSynchronize(st));
	drain.armed = false;
	return 0;
}

extern "C" int psxb200_spu_encode_host(int n_streams, const int16_t *h_samples, int pitch, long group_stride,
                                       int sample_count, void *h_states, uint8_t *h_out, long out_stride) {
	return spu_encode_host_subset(n_streams, 0, 1, h_samples, pitch, group_stride, sample_count, h_states, h_out, out_stride);
}

static int xa_check(const char *who, int format, int bits_per_sample) {
	if (bits_per_sample != 4 && bits_per_sample != 8) return fail("%s: bits_per_sample must be 4 or 8", who);
	if (format != 0 && format != 1) return fail("%s: format must be 0 (XA) or 1 (XACD)", who);
	return 0;
}

extern "C" int psxb200_xa_encode_device_ex(int n_streams, int format, int stereo, int frequency, int bits_per_sample,
                                           int file_number, int channel_number, const int16_t *d_samples, long in_stride,
                                           int sample_count, int lba, int lba_step, void *d_states, uint8_t *d_out,
                                           long out_stride, long sector_stride, void *stream) {
	if (xa_check("psxb200_xa_encode_device", format, bits_per_sample)) return -1;
	int sectors = adpcm_xa_sectors(stereo, bits_per_sample, sample_count);
	int size = format == 0 ? 2336 : 2352;
	if (n_streams <= 0 || sectors == 0) return 0;
	if (((uintptr_t)d_out & 3) || (out_stride & 3) || (sector_stride & 3) || ((uintptr_t)d_states & 7))
		return fail("psxb200_xa_encode_device: alignment contract violated (out, strides: 4 bytes, states: 8 bytes)");
	if (sector_stride > 0 && sector_stride < size) return fail("psxb200_xa_encode_device: sector_stride %ld < sector size", sector_stride);
	const uint32_t *edc = edc_tables_device();
	if (!edc) return fail("psxb200_xa_encode_device: EDC tables unavailable");
	CU_TRY(adpcm_launch_xa(n_streams, format, stereo, frequency, bits_per_sample, file_number, channel_number, d_samples,
	                       in_stride, sample_count, lba, lba_step, d_states, d_out, out_stride, sector_stride, edc,
	                       static_cast<cudaStream_t>(stream)));
	g_launches += 2;
	return sectors * size;
}

extern "C" int psxb200_xa_encode_device(int n_streams, int format, int stereo, int frequency, int bits_per_sample,
                                        int file_number, int channel_number, const int16_t *d_samples, long in_stride,
                                        int sample_count, int lba, void *d_states, uint8_t *d_out, long out_stride,
                                        void *stream) {
	return psxb200_xa_encode_device_ex(n_streams, format, stereo, frequency, bits_per_sample, file_number, channel_number,
	                                   d_samples, in_stride, sample_count, lba, 1, d_states, d_out, out_stride, 0, stream);
}

extern "C" int psxb200_xa_encode_host(int n_streams, int format, int stereo, int frequency, int bits_per_sample,
                                      int file_number, int channel_number, const int16_t *h_samples, long in_stride,
                                      int sample_count, int lba, void *h_states, uint8_t *h_out, long out_stride) {
	if (xa_check("psxb200_xa_encode_host", format, bits_per_sample)) return -1;
	int sectors = adpcm_xa_sectors(stereo, bits_per_sample, sample_count);
	int size = format == 0 ? 2336 : 2352;
	if (n_streams <= 0 || sectors == 0) return 0;
	AudioContext *ctx = audio_context();
	if (!ctx) return -1;
	std::lock_guard<std::mutex> guard(ctx->lock);
	if (ctx->ensure()) return -1;
	cudaStream_t st = ctx->stream;
	StreamDrain drain(st);

	const long extent = adpcm_xa_input_extent(stereo, bits_per_sample, sample_count);
	const long dstride_in = (long)round_up((size_t)extent, 8);
	const long bytes = (long)sectors * size;
	// 16 bytes of slack in front: the 2336-byte format addresses sectors 16 bytes early
	const long dstride_out = (long)round_up((size_t)bytes, 16);
	CU_TRY(ctx->in.reserve((size_t)n_streams * dstride_in));
	CU_TRY(ctx->out.reserve((size_t)n_streams * dstride_out + 16));
	CU_TRY(ctx->states.reserve((size_t)n_streams * 2 * STATE_BYTES));
	uint8_t *d_out = ctx->out.ptr + 16;
	const size_t h_in_pitch = n_streams == 1 ? (size_t)extent * 2 : (size_t)in_stride * 2;
	const size_t h_out_pitch = n_streams == 1 ? (size_t)bytes : (size_t)out_stride;
	CU_TRY(cudaMemcpy2DAsync(ctx->in.ptr, (size_t)dstride_in * 2, h_samples, h_in_pitch, (size_t)extent * 2,
	                         n_streams, cudaMemcpyHostToDevice, st));
	// bytes the reference never writes keep the caller's content: round-trip the output buffer
	CU_TRY(cudaMemcpy2DAsync(d_out, (size_t)dstride_out, h_out, h_out_pitch, (size_t)bytes, n_streams,
	                         cudaMemcpyHostToDevice, st));
	CU_TRY(cudaMemcpyAsync(ctx->states.ptr, h_states, (size_t)n_streams * 2 * STATE_BYTES, cudaMemcpyHostToDevice, st));
	CU_TRY(adpcm_launch_xa(n_streams, format, stereo, frequency, bits_per_sample, file_number, channel_number,
	                       ctx->in.ptr, dstride_in, sample_count, lba, 1, ctx->states.ptr, d_out, dstride_out, 0, ctx->edc, st));
	g_launches += 2;
	CU_TRY(cudaMemcpy2DAsync(h_out, h_out_pitch, d_out, (size_t)dstride_out, (size_t)bytes, n_streams,
	                         cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaMemcpyAsync(h_states, ctx->states.ptr, (size_t)n_streams * 2 * STATE_BYTES, cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaStreamSynchronize(st));
	drain.armed = false;
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
