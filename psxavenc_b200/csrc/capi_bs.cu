// C ABI of libpsxav_b200.so, video half (declared in include/psxav_b200.h): the batched
// psxb200_bs_* / psxb200_str_* entry points and the drop-in replacements for the reference's
// MDEC encoder symbols (psxavenc/mdec.h:65-74). Host code only; the kernels live in
// bs_encode.cu. There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <atomic>

#include <algorithm>
#include <cstring>
#include <vector>

#include "psxav_b200.h"
#include "adpcm_encode.h"
#include "bs_encode.h"
#include "capi_bs.h"
#include "capi_util.h"

using namespace psxb200;

namespace psxb200 {

thread_local char g_error[512] = "";
std::atomic<unsigned long long> g_launches{0};

int fail(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_error, sizeof(g_error), fmt, ap);
	va_end(ap);
	return -1;
}

[[noreturn]] void die(const char *what) {
	fprintf(stderr, "libpsxav_b200: %s: %s\n", what, g_error);
	abort();
}

}  // namespace psxb200

static int bs_pick_threads(const BsGeometry &geo) {
	// 10 warps per CTA: four such CTAs fit an SM at 48 registers per thread (three at 64 when the
	// shared memory does not allow four, see bs_encode_chunked) and the usual frame sizes' groups
	// of 32 blocks divide with <= 5 % idle warp slots (320x240: 57 groups in 6 rounds, 640x480:
	// 225 in 23); measured best on B200 (profiles/r1_sweeps.md).
	const char *env = getenv("PSXB200_PACK_THREADS");
	if (env && atoi(env) >= 32) return std::min(BS_PACK_MAX_THREADS, atoi(env) / 32 * 32);
	return 32 * std::max(1, std::min(10, geo.ngroups));
}

cudaError_t psxb200_bs_encoder::mark(cudaStream_t st) {
	if (events_used == events.size()) {
		cudaEvent_t e;
		cudaError_t rc = cudaEventCreate(&e);
		if (rc != cudaSuccess) return rc;
		events.push_back(e);
	}
	return cudaEventRecord(events[events_used++], st);
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

extern "C" void *psxb200_pinned_alloc(size_t bytes) {
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
		fail("psxb200_pinned_alloc(%zu): %s", bytes, cudaGetErrorString(cudaGetLastError()));
		return nullptr;
	}
	return p;
}

extern "C" void psxb200_pinned_free(void *p) {
	if (p) cudaFreeHost(p);
}

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
	cudaError_t e = cudaGetDevice(&enc->device);
	if (e == cudaSuccess) e = cudaDeviceGetAttribute(&enc->sm_count, cudaDevAttrMultiProcessorCount, enc->device);
	if (const char *env = getenv("PSXB200_PACK_MIN_CTAS")) enc->pack_min_ctas = atoi(env);
	if (const char *env = getenv("PSXB200_DEVICE_PIPELINE")) enc->device_pipeline = atoi(env);
	if (const char *env = getenv("PSXB200_HOST_CHUNK")) enc->host_chunk = std::max(1, std::min(max_batch, atoi(env)));
	if (e == cudaSuccess) e = bs_upload_tables();
	if (e == cudaSuccess) e = enc->coefs.reserve((size_t)max_batch * enc->geo.frame_stride_u4);
	if (e != cudaSuccess) {
		fail("psxb200_bs_create: %s", cudaGetErrorString(e));
		cudaGetLastError();
		delete enc;
		return nullptr;
	}
	return enc;
}

extern "C" void psxb200_bs_destroy(psxb200_bs_encoder_t *enc) {
	if (!enc) return;
	DeviceGuard guard(enc->device);
	for (BsSlot &s : enc->slots) {
		if (s.stream) {
			cudaStreamSynchronize(s.stream);
			cudaStreamDestroy(s.stream);
		}
		if (s.res_ready) cudaEventDestroy(s.res_ready);
		if (s.audio_stream) {
			cudaStreamSynchronize(s.audio_stream);
			cudaStreamDestroy(s.audio_stream);
		}
		if (s.audio_done) cudaEventDestroy(s.audio_done);
		if (s.image_ready) cudaEventDestroy(s.image_ready);
		s.coefs.release();
		s.gstream.release();
		s.in.release();
		s.out.release();
		s.sizes.release();
		s.res.release();
		s.h_res.release();
		s.pcm.release();
		s.states.release();
		s.h_states.release();
	}
	enc->coefs.release();
	enc->gstream.release();
	enc->coefs2.release();
	enc->gstream2.release();
	for (int i = 0; i < 2; i++) {
		if (enc->pipe[i]) {
			cudaStreamSynchronize(enc->pipe[i]);
			cudaStreamDestroy(enc->pipe[i]);
		}
		if (enc->pipe_join[i]) cudaEventDestroy(enc->pipe_join[i]);
	}
	if (enc->pipe_fork) cudaEventDestroy(enc->pipe_fork);
	BsLookahead &a = enc->ahead;
	if (a.stream) {
		cudaStreamSynchronize(a.stream);
		cudaStreamDestroy(a.stream);
	}
	for (auto &g : a.graphs)
		if (g.exec) cudaGraphExecDestroy(g.exec);
	a.staged.release();
	a.h_out.release();
	a.h_res.release();
	a.in.release();
	a.out.release();
	a.res.release();
	a.coefs.release();
	a.gstream.release();
	for (cudaEvent_t e : enc->events) cudaEventDestroy(e);
	delete enc;
}

extern "C" int psxb200_bs_device(const psxb200_bs_encoder_t *enc) { return enc ? enc->device : -1; }
extern "C" long long psxb200_bs_frame_bytes(const psxb200_bs_encoder_t *enc) { return enc ? (long long)enc->frame_bytes : 0; }

// FDCT + pack (+ STR framing) kernels for n device-resident frames, in launches of at most
// max_batch frames; `coefs` and `gstream` are the scratch of the calling pipeline slot (kernels
// of different slots may overlap, so they never share scratch).
static int bs_encode_chunked(psxb200_bs_encoder *enc, DeviceBuffer<uint4> &coefs, DeviceBuffer<uint32_t> &gstream_buf, int n,
                             const uint8_t *d_frames, const int *d_max_sizes, int max_size_bound, uint8_t *d_out,
                             size_t out_stride, psxb200_bs_result_t *d_results, cudaStream_t stream,
                             const BsStrLayout *str_batch = nullptr) {
	uint32_t *gstream = nullptr;
	size_t gstride = 0;
	const int launch = std::min(enc->max_batch, n);
	CU_TRY(coefs.reserve((size_t)launch * enc->geo.frame_stride_u4));
	// Few frames (the drop-in calls encode one at a time): every CTA has an SM to itself, so the
	// frame's latency is what counts and the widest CTA wins (88 vs 102 us per drop-in frame).
	int threads = enc->pack_threads, min_ctas = enc->pack_min_ctas;
	if (n <= enc->sm_count && !enc->pack_threads_forced) {
		threads = 32 * std::max(1, std::min(BS_PACK_MAX_THREADS / 32, enc->geo.ngroups));
		min_ctas = 1;
	}
	// Very few frames (the drop-in symbols): a cluster of CTAs per frame, see bs_pack_kernel
	static const bool cluster_off = getenv("PSXB200_NO_CLUSTER") != nullptr;
	static std::atomic<bool> cluster_refused{false};   // a cluster launch was turned down once (a partitioned device, say): never again
	int cluster = 1;
	// (clusters are placed within a GPC, so they do not tile all SMs: with half of the SMs asked for, every cluster of
	// the launch is resident at once — 37 clusters of 4 on 148 SMs ran in two waves, 36.9 us against 32.8 us without)
	if (2 * n * BS_PACK_CLUSTER <= enc->sm_count && !enc->pack_threads_forced && !cluster_off && !cluster_refused.load()) {
		const int cl_threads = 32 * std::max(1, std::min(BS_PACK_MAX_THREADS / 32, (enc->geo.ngroups + BS_PACK_CLUSTER - 1) / BS_PACK_CLUSTER));
		if (bs_pack_smem_bytes(enc->codec != 0, true, enc->geo, max_size_bound, cl_threads) <= BS_SMEM_BUDGET) {
			cluster = BS_PACK_CLUSTER;
			threads = cl_threads;
		}
	}
	const size_t smem = bs_pack_smem_bytes(enc->codec != 0, true, enc->geo, max_size_bound, threads);
	if (smem > BS_SMEM_BUDGET) {
		gstride = (size_t)(max_size_bound + 3) / 4 + 2;
		CU_TRY(gstream_buf.reserve(gstride * launch));
		gstream = gstream_buf.ptr;
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
		                     enc->height, enc->geo, coefs.ptr, stream));
		if (enc->timing) CU_TRY(enc->mark(stream));
		BsStrLayout str{};
		if (str_batch) {
			str = *str_batch;
			str.frame_base += first;   // the kernel positions every frame absolutely within the batch
		}
		if (cluster > 1) {
			// launch-configuration errors come back synchronously: fall back to one CTA per frame
			if (bs_launch_pack_cluster(enc->codec, threads, m, coefs.ptr, enc->geo,
			                           (str_batch || !d_max_sizes) ? nullptr : d_max_sizes + first, max_size_bound,
			                           str_batch ? d_out : d_out + (size_t)first * out_stride, out_stride, d_results + first,
			                           str, stream) != cudaSuccess) {
				cudaGetLastError();
				cluster_refused.store(true);
				cluster = 1;
			}
		}
		if (cluster == 1)
			CU_TRY(bs_launch_pack(enc->codec, threads, min_ctas, m, coefs.ptr, enc->geo,
			                      (str_batch || !d_max_sizes) ? nullptr : d_max_sizes + first, max_size_bound,
			                      str_batch ? d_out : d_out + (size_t)first * out_stride, out_stride, d_results + first, gstream,
			                      gstride, str, stream));
		if (enc->timing) CU_TRY(enc->mark(stream));
		g_launches += (gstream || cluster > 1) ? 2 : 3;   // FDCT, pack, and (shared-memory image) the pack kernel for deferred frames
		if (str_batch && str.framing && str.format != FORMAT_STRV) {
			CU_TRY(bs_launch_str_framing(m, max_size_bound / 2016, d_out, str, stream));
			g_launches += 1;
		}
	}
	return 0;
}

extern "C" void psxb200_bs_timing_enable(psxb200_bs_encoder_t *enc, int on) {
	enc->timing = on != 0;
	enc->events_used = 0;
}

extern "C" int psxb200_bs_timing_read(psxb200_bs_encoder_t *enc, double *dct_ms, double *pack_ms, int *launch_pairs) {
	DeviceGuard guard(enc->device);
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
	DeviceGuard guard(enc->device);
	CU_TRY(guard.status);
	cudaStream_t user = static_cast<cudaStream_t>(stream);
	if (!enc->device_pipeline || n <= enc->max_batch || enc->timing)
		return bs_encode_chunked(enc, enc->coefs, enc->gstream, n, d_frames, d_max_sizes, max_size_bound, d_out, out_stride,
		                         d_results, user);
	// several launches: alternate them between two forked streams (see psxb200_bs_encoder::pipe)
	for (int i = 0; i < 2; i++) {
		if (!enc->pipe[i]) CU_TRY(cudaStreamCreateWithFlags(&enc->pipe[i], cudaStreamNonBlocking));
		if (!enc->pipe_join[i]) CU_TRY(cudaEventCreateWithFlags(&enc->pipe_join[i], cudaEventDisableTiming));
	}
	if (!enc->pipe_fork) CU_TRY(cudaEventCreateWithFlags(&enc->pipe_fork, cudaEventDisableTiming));
	CU_TRY(cudaEventRecord(enc->pipe_fork, user));
	for (int i = 0; i < 2; i++) CU_TRY(cudaStreamWaitEvent(enc->pipe[i], enc->pipe_fork, 0));
	int k = 0;
	for (int first = 0; first < n; first += enc->max_batch, k ^= 1) {
		const int m = std::min(enc->max_batch, n - first);
		if (bs_encode_chunked(enc, k ? enc->coefs2 : enc->coefs, k ? enc->gstream2 : enc->gstream, m,
		                      d_frames + (size_t)first * enc->frame_bytes, d_max_sizes ? d_max_sizes + first : nullptr, max_size_bound,
		                      d_out + (size_t)first * out_stride, out_stride, d_results + first, enc->pipe[k]))
			return -1;
	}
	for (int i = 0; i < 2; i++) {
		CU_TRY(cudaEventRecord(enc->pipe_join[i], enc->pipe[i]));
		CU_TRY(cudaStreamWaitEvent(user, enc->pipe_join[i], 0));
	}
	return 0;
}

// ---- host pipeline ----------------------------------------------------------------------

static int slot_prepare(psxb200_bs_encoder *enc, BsSlot &s) {
	if (!s.stream) CU_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
	if (!s.res_ready) CU_TRY(cudaEventCreateWithFlags(&s.res_ready, cudaEventDisableTiming));
	(void)enc;
	return 0;
}

namespace {

// Early (error) return of a host entry point: waits for whatever the slots still have in flight.
// Those copies read and write the caller's buffers, which it may release once the call is back.
struct SlotDrain {
	psxb200_bs_encoder *enc;
	bool armed = true;
	explicit SlotDrain(psxb200_bs_encoder *e) : enc(e) {}
	~SlotDrain() {
		if (!armed) return;
		for (BsSlot &s : enc->slots) {
			if (s.stream) cudaStreamSynchronize(s.stream);
			if (s.audio_stream) cudaStreamSynchronize(s.audio_stream);
		}
		cudaGetLastError();
	}
	SlotDrain(const SlotDrain &) = delete;
	SlotDrain &operator=(const SlotDrain &) = delete;
};

// One chunk of psxb200_bs_encode_host in flight on a slot.
struct BsChunk {
	int first = 0, m = 0, bound = 0;
	size_t dstride = 0;
};

}  // namespace

extern "C" int psxb200_bs_encode_host(psxb200_bs_encoder_t *enc, int n, const uint8_t *h_frames,
                                      const int *h_max_sizes, uint8_t *h_out, size_t out_stride,
                                      psxb200_bs_result_t *h_results) {
	if (!enc) return fail("psxb200_bs_encode_host: NULL encoder");
	if (n <= 0) return 0;
	if (!h_frames || !h_max_sizes || !h_out || !h_results) return fail("psxb200_bs_encode_host: NULL argument");
	DeviceGuard guard(enc->device);
	CU_TRY(guard.status);
	for (BsSlot &s : enc->slots)
		if (slot_prepare(enc, s)) return -1;
	for (int i = 0; i < n; i++)
		if (n > 1 && (size_t)std::max(h_max_sizes[i], 0) > out_stride)
			return fail("psxb200_bs_encode_host: frame_max_size %d of frame %d > out_stride", h_max_sizes[i], i);
	SlotDrain drain(enc);

	// Chunks of host_chunk frames rotate through BS_SLOTS streams. A chunk goes through two
	// phases: (A) frames in, kernels, result rows out; (B) once the rows are on the host, only
	// the bytes the frames actually produced are copied back — a strided copy as wide as the
	// chunk's longest stream, never past a frame's own budget — while this thread zero-fills the
	// rest of every budget (the reference clears the whole buffer, mdec.c:676). Phase B of a
	// chunk is issued after phase A of the next one, so the copy engines stay busy.
	const int hc = enc->host_chunk;
	// experiments only (tools/e2e_probe.py): leave the zero fill of the tails out to see what it costs
	static const bool skip_tail_zero = getenv("PSXB200_EXPERIMENT_SKIP_TAIL_ZERO") != nullptr;
	BsChunk inflight[BS_SLOTS];
	bool busy[BS_SLOTS] = {};

	auto phase_a = [&](int slot, int first, int m) -> int {
		BsSlot &s = enc->slots[slot];
		int bound = 8;
		bool uniform = true;   // one budget for the whole chunk: no per-frame array to upload
		for (int i = 0; i < m; i++) {
			bound = std::max(bound, h_max_sizes[first + i]);
			uniform = uniform && h_max_sizes[first + i] == h_max_sizes[first];
		}
		uniform = uniform && h_max_sizes[first] >= 8;
		const size_t dstride = round_up((size_t)bound, 16);
		CU_TRY(s.in.reserve((size_t)hc * enc->frame_bytes));
		CU_TRY(s.out.reserve((size_t)hc * dstride));
		CU_TRY(s.sizes.reserve(hc));
		CU_TRY(s.res.reserve(hc));
		CU_TRY(s.h_res.reserve(hc));
		CU_TRY(cudaMemcpyAsync(s.in.ptr, h_frames + (size_t)first * enc->frame_bytes, (size_t)m * enc->frame_bytes,
		                       cudaMemcpyHostToDevice, s.stream));
		if (!uniform)
			CU_TRY(cudaMemcpyAsync(s.sizes.ptr, h_max_sizes + first, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, s.stream));
		if (bs_encode_chunked(enc, s.coefs, s.gstream, m, s.in.ptr, uniform ? nullptr : s.sizes.ptr, bound, s.out.ptr, dstride,
		                      s.res.ptr, s.stream))
			return -1;
		CU_TRY(cudaMemcpyAsync(s.h_res.ptr, s.res.ptr, (size_t)m * sizeof(psxb200_bs_result_t), cudaMemcpyDeviceToHost, s.stream));
		CU_TRY(cudaEventRecord(s.res_ready, s.stream));
		inflight[slot] = BsChunk{first, m, bound, dstride};
		busy[slot] = true;
		return 0;
	};

	auto phase_b = [&](int slot) -> int {
		BsSlot &s = enc->slots[slot];
		const BsChunk c = inflight[slot];
		CU_TRY(cudaEventSynchronize(s.res_ready));
		// bytes of frame i that hold its stream: min(bytes_used, budget) (bytes_used is rounded up
		// to a multiple of 4 and may pass an odd budget by up to 3, mdec.c:736)
		int longest = 0, smallest_budget = INT32_MAX;
		for (int i = 0; i < c.m; i++) {
			const int budget = std::max(h_max_sizes[c.first + i], 0);
			longest = std::max(longest, std::min(s.h_res.ptr[i].bytes_used, budget));
			smallest_budget = std::min(smallest_budget, budget);
		}
		const int width = std::min((int)round_up((size_t)longest, 64), smallest_budget);
		uint8_t *dst = h_out + (size_t)c.first * out_stride;
		if (width > 0) {
			if (c.m == 1)
				CU_TRY(cudaMemcpyAsync(dst, s.out.ptr, (size_t)width, cudaMemcpyDeviceToHost, s.stream));
			else
				CU_TRY(cudaMemcpy2DAsync(dst, out_stride, s.out.ptr, c.dstride, (size_t)width, c.m, cudaMemcpyDeviceToHost, s.stream));
		}
		for (int i = 0; i < c.m; i++) {
			const int budget = std::max(h_max_sizes[c.first + i], 0);
			const int used = std::min(s.h_res.ptr[i].bytes_used, budget);
			uint8_t *row = dst + (size_t)i * out_stride;
			if (used > width)   // budgets differ within the chunk and this frame is longer than the smallest
				CU_TRY(cudaMemcpyAsync(row + width, s.out.ptr + (size_t)i * c.dstride + width, (size_t)(used - width),
				                       cudaMemcpyDeviceToHost, s.stream));
			const int clean_from = std::max(used, width);
			if (budget > clean_from && !skip_tail_zero) memset(row + clean_from, 0, (size_t)(budget - clean_from));
			h_results[c.first + i] = s.h_res.ptr[i];
		}
		return 0;
	};

	auto finish = [&](int slot) -> int {
		if (busy[slot]) CU_TRY(cudaStreamSynchronize(enc->slots[slot].stream));
		busy[slot] = false;
		return 0;
	};

	int chunk = 0;
	for (int first = 0; first < n; first += hc, chunk++) {
		const int slot = chunk % BS_SLOTS;
		if (finish(slot)) return -1;
		if (phase_a(slot, first, std::min(hc, n - first))) return -1;
		if (chunk > 0 && phase_b((chunk - 1) % BS_SLOTS)) return -1;
	}
	if (phase_b((chunk - 1) % BS_SLOTS)) return -1;
	for (int slot = 0; slot < BS_SLOTS; slot++)
		if (finish(slot)) return -1;
	drain.armed = false;
	int failed = 0;
	for (int i = 0; i < n; i++) failed += h_results[i].quant_scale >= 64;
	return failed;
}

// ---- STR video sectors ---------------------------------------------

static int str_layout(const psxb200_bs_encoder *enc, const psxb200_str_params_t *p, BsStrLayout *out) {
	if (!p) return fail("psxb200_str_*: NULL params");
	if (p->first_frame_index < 1 || p->sectors_num < 1 || p->sectors_den < 1)
		return fail("psxb200_str_*: first_frame_index, sectors_num and sectors_den must be >= 1");
	if (p->interleave < 0 || p->frames_per_file < 0 || (p->file_stride & 3))
		return fail("psxb200_str_*: bad interleave / frames_per_file / file_stride (multiple of 4)");
	BsStrLayout l{};
	// sector size / header offset per container (mdec.c:824-829; filefmt.c:453,502,572,613)
	if (p->format == FORMAT_STRV) { l.sector_size = 2048; l.header_offset = 0; }
	else if (p->format == FORMAT_STR) { l.sector_size = 2336; l.header_offset = 8; }
	else if (p->format == FORMAT_STRCD) { l.sector_size = 2352; l.header_offset = 0x18; }
	else return fail("psxb200_str_*: format must be FORMAT_STR, FORMAT_STRCD or FORMAT_STRV");
	l.format = p->format;
	l.frame_index0 = p->first_frame_index;
	l.sectors_num = p->sectors_num;
	l.sectors_den = p->sectors_den;
	l.video_id = p->video_id;
	l.width = enc->width;
	l.height = enc->height;
	l.frame_base = 0;
	l.frames_per_file = p->frames_per_file;
	l.file_stride = p->file_stride;
	l.interleave = p->interleave > 1 ? p->interleave : 1;
	l.audio_first = p->trailing_audio ? 0 : 1;
	l.place_at_lba = p->place_at_lba ? 1 : 0;
	l.framing = p->framing ? 1 : 0;
	l.xa_file = p->xa_file;
	l.xa_channel = p->xa_channel;
	const long long v_first = (long long)(p->first_frame_index - 1) * p->sectors_num / p->sectors_den;
	l.slot0 = l.place_at_lba ? p->lba_origin : v_first;
	if (l.frames_per_file > 0 && l.file_stride <= 0) return fail("psxb200_str_*: frames_per_file needs file_stride");
	*out = l;
	return 0;
}

// slots [lo, hi) taken by the video sectors of m frames starting at frame_index K0 of a file
static void str_slot_range(const BsStrLayout &l, long long K0, int m, long long *lo, long long *hi) {
	const long long v_lo = (K0 - 1) * l.sectors_num / l.sectors_den;
	const long long v_hi = (K0 - 1 + m) * l.sectors_num / l.sectors_den;
	*lo = bs_str_slot(l, v_lo);
	*hi = v_hi > v_lo ? bs_str_slot(l, v_hi - 1) + 1 : *lo;
}

static int str_max_budget(const BsStrLayout &l) { return 2016 * ((l.sectors_num + l.sectors_den - 1) / l.sectors_den); }

extern "C" long long psxb200_str_sector_count(int n_frames, int first_frame_index, int sectors_num, int sectors_den) {
	long long a = (long long)(first_frame_index - 1) * sectors_num / sectors_den;
	long long b = (long long)(first_frame_index - 1 + n_frames) * sectors_num / sectors_den;
	return b - a;
}

extern "C" int psxb200_str_slot_range(const psxb200_str_params_t *p, int n_frames, long long *first_slot, long long *end_slot) {
	if (!p || p->sectors_num < 1 || p->sectors_den < 1 || p->first_frame_index < 1) return fail("psxb200_str_slot_range: bad params");
	BsStrLayout l{};
	l.sectors_num = p->sectors_num;
	l.sectors_den = p->sectors_den;
	l.interleave = p->interleave > 1 ? p->interleave : 1;
	l.audio_first = p->trailing_audio ? 0 : 1;
	l.place_at_lba = p->place_at_lba ? 1 : 0;
	long long lo, hi;
	str_slot_range(l, p->first_frame_index, n_frames, &lo, &hi);
	const long long v_first = (long long)(p->first_frame_index - 1) * p->sectors_num / p->sectors_den;
	const long long origin = l.place_at_lba ? p->lba_origin : v_first;
	if (first_slot) *first_slot = lo - origin;
	if (end_slot) *end_slot = hi - origin;
	return 0;
}

extern "C" int psxb200_str_encode_device_ex(psxb200_bs_encoder_t *enc, int n, const uint8_t *d_frames,
                                            const psxb200_str_params_t *params, uint8_t *d_sectors,
                                            psxb200_bs_result_t *d_results, void *stream) {
	if (!enc) return fail("psxb200_str_encode_device: NULL encoder");
	if (n <= 0) return 0;
	BsStrLayout l;
	if (str_layout(enc, params, &l)) return -1;
	if (((uintptr_t)d_frames & 15) || ((uintptr_t)d_sectors & 3))
		return fail("psxb200_str_encode_device: alignment contract violated (frames 16 bytes, sectors 4 bytes)");
	DeviceGuard guard(enc->device);
	CU_TRY(guard.status);
	return bs_encode_chunked(enc, enc->coefs, enc->gstream, n, d_frames, nullptr, str_max_budget(l), d_sectors, 0, d_results,
	                         static_cast<cudaStream_t>(stream), &l);
}

static psxb200_str_params_t str_simple_params(int format, int first_frame_index, int sectors_num, int sectors_den, int video_id) {
	psxb200_str_params_t p;
	memset(&p, 0, sizeof(p));
	p.format = format;
	p.first_frame_index = first_frame_index;
	p.sectors_num = sectors_num;
	p.sectors_den = sectors_den;
	p.video_id = video_id;
	p.interleave = 1;
	return p;
}

extern "C" int psxb200_str_encode_device(psxb200_bs_encoder_t *enc, int n, const uint8_t *d_frames, int format,
                                         int first_frame_index, int sectors_num, int sectors_den, int video_id,
                                         uint8_t *d_sectors, psxb200_bs_result_t *d_results, void *stream) {
	psxb200_str_params_t p = str_simple_params(format, first_frame_index, sectors_num, sectors_den, video_id);
	return psxb200_str_encode_device_ex(enc, n, d_frames, &p, d_sectors, d_results, stream);
}

// Host pipeline for STR sectors. A chunk is either a run of whole files (frames_per_file <=
// host_chunk) or a run of frames inside one file. Its sectors are produced in a compact device
// region and come back with one (strided) copy per chunk covering only the bytes of a sector
// the encoder writes; layouts in which that range would also cover bytes the encoder leaves
// alone (audio slots of a muxed image, the 16 unwritten bytes inside FORMAT_STR's EDC range) first
// upload what the caller's buffer holds, so those bytes keep their content and the EDC is
// computed over them exactly as in the reference.
extern "C" int psxb200_str_encode_host_ex(psxb200_bs_encoder_t *enc, int n, const uint8_t *h_frames,
                                          const psxb200_str_params_t *params, uint8_t *h_sectors,
                                          psxb200_bs_result_t *h_results) {
	if (!enc) return fail("psxb200_str_encode_host: NULL encoder");
	if (n <= 0) return 0;
	if (!h_frames || !h_sectors || !h_results) return fail("psxb200_str_encode_host: NULL argument");
	BsStrLayout batch;
	if (str_layout(enc, params, &batch)) return -1;
	DeviceGuard guard(enc->device);
	CU_TRY(guard.status);
	for (BsSlot &s : enc->slots)
		if (slot_prepare(enc, s)) return -1;
	SlotDrain drain(enc);

	const int hc = enc->host_chunk;
	const int ss = batch.sector_size;
	const int bound = str_max_budget(batch);
	const int fpf = batch.frames_per_file;
	if (fpf > 0 && n % fpf) return fail("psxb200_str_encode_host: n (%d) is not a multiple of frames_per_file (%d)", n, fpf);
	const bool whole_files = fpf > 0 && fpf <= hc;
	const bool roundtrip = batch.format == FORMAT_STR ? batch.framing : (batch.place_at_lba && batch.interleave > 1);
	// column range of a sector that is written (and copied back when not round-tripping)
	int col_lo = 0, col_hi = ss;
	if (batch.format == FORMAT_STRCD) {
		col_lo = batch.framing ? 0 : 0x18;
		col_hi = batch.framing ? 0x81C : 0x818;
	} else if (batch.format == FORMAT_STR) {
		col_lo = 8;
		col_hi = 0x808;   // framing: round trip (the EDC lands at 0x818, past 16 untouched bytes)
	}
	const bool strided = !roundtrip && (col_lo != 0 || col_hi != ss);

	struct Pending { int first, m; };
	Pending pending[BS_SLOTS] = {};
	bool busy[BS_SLOTS] = {};
	auto finish = [&](int slot) -> int {
		if (!busy[slot]) return 0;
		BsSlot &s = enc->slots[slot];
		CU_TRY(cudaStreamSynchronize(s.stream));
		memcpy(h_results + pending[slot].first, s.h_res.ptr, (size_t)pending[slot].m * sizeof(psxb200_bs_result_t));
		busy[slot] = false;
		return 0;
	};

	int chunk = 0;
	for (int first = 0; first < n; chunk++) {
		const int slot = chunk % BS_SLOTS;
		BsSlot &s = enc->slots[slot];
		if (finish(slot)) return -1;

		// frames [first, first + m) of the batch; files [file0, file0 + files) when whole_files
		int m, files = 1, file0 = 0;
		long long k0 = batch.frame_index0;    // frame_index of the chunk's first frame within its file
		if (whole_files) {
			file0 = first / fpf;
			files = std::min(hc / fpf, (n - first) / fpf);
			m = files * fpf;
		} else if (fpf > 0) {
			file0 = first / fpf;
			const int in_file = first - file0 * fpf;
			k0 += in_file;
			m = std::min(std::min(hc, fpf - in_file), n - first);
		} else {
			k0 += first;
			m = std::min(hc, n - first);
		}
		// slots of one file's share of the chunk (whole_files: every file has the same range)
		long long lo, hi;
		str_slot_range(batch, k0, whole_files ? fpf : m, &lo, &hi);
		const size_t region = (size_t)(hi - lo) * ss;
		const size_t dev_stride = round_up(region, 16);

		BsStrLayout l = batch;
		l.frame_index0 = (int)k0;
		l.frame_base = 0;
		l.frames_per_file = whole_files ? fpf : 0;
		l.file_stride = (long long)dev_stride;
		l.slot0 = lo;

		CU_TRY(s.in.reserve((size_t)hc * enc->frame_bytes));
		CU_TRY(s.out.reserve(dev_stride * files + 16));
		CU_TRY(s.res.reserve(hc));
		CU_TRY(s.h_res.reserve(hc));
		CU_TRY(cudaMemcpyAsync(s.in.ptr, h_frames + (size_t)first * enc->frame_bytes, (size_t)m * enc->frame_bytes,
		                       cudaMemcpyHostToDevice, s.stream));
		uint8_t *h_region = h_sectors + (size_t)file0 * (size_t)batch.file_stride + (size_t)(lo - batch.slot0) * ss;
		const size_t h_pitch = whole_files ? (size_t)batch.file_stride : region;
		if (roundtrip && region)
			CU_TRY(cudaMemcpy2DAsync(s.out.ptr, dev_stride, h_region, std::max(h_pitch, region), region, files,
			                         cudaMemcpyHostToDevice, s.stream));
		if (bs_encode_chunked(enc, s.coefs, s.gstream, m, s.in.ptr, nullptr, bound, s.out.ptr, 0, s.res.ptr, s.stream, &l))
			return -1;
		if (region) {
			if (!strided) {
				CU_TRY(cudaMemcpy2DAsync(h_region, std::max(h_pitch, region), s.out.ptr, dev_stride, region, files,
				                         cudaMemcpyDeviceToHost, s.stream));
			} else {
				// rows = sectors: one strided copy per file (rows of different files are file_stride apart)
				for (int f = 0; f < files; f++)
					CU_TRY(cudaMemcpy2DAsync(h_region + (size_t)f * h_pitch + col_lo, ss, s.out.ptr + (size_t)f * dev_stride + col_lo,
					                         ss, (size_t)(col_hi - col_lo), (size_t)(hi - lo), cudaMemcpyDeviceToHost, s.stream));
			}
		}
		CU_TRY(cudaMemcpyAsync(s.h_res.ptr, s.res.ptr, (size_t)m * sizeof(psxb200_bs_result_t), cudaMemcpyDeviceToHost, s.stream));
		pending[slot] = Pending{first, m};
		busy[slot] = true;
		first += m;
	}
	for (int slot = 0; slot < BS_SLOTS; slot++)
		if (finish(slot)) return -1;
	drain.armed = false;
	int failed = 0;
	for (int i = 0; i < n; i++) failed += h_results[i].quant_scale >= 64;
	return failed;
}

extern "C" int psxb200_str_encode_host(psxb200_bs_encoder_t *enc, int n, const uint8_t *h_frames, int format,
                                       int first_frame_index, int sectors_num, int sectors_den, int video_id,
                                       uint8_t *h_sectors, psxb200_bs_result_t *h_results) {
	psxb200_str_params_t p = str_simple_params(format, first_frame_index, sectors_num, sectors_den, video_id);
	return psxb200_str_encode_host_ex(enc, n, h_frames, &p, h_sectors, h_results);
}

// ---- .str / .strcd file images: video + XA audio muxed on the GPU (SURVEY.md 8e "strcd") ---
//
// encode_file_str (filefmt.c:391-520) for n_files independent inputs: file f holds
// frames_per_file video frames and samples_per_file XA sample frames and becomes the image
// h_images + f * image_stride: slot s of the image is an XA audio sector when s % interleave == 0
// (trailing_audio: == interleave - 1), else the next video sector. Files are processed in groups;
// a group's video sectors (FDCT/pack/framing kernels) and XA sectors (ADPCM/framing kernels) are
// produced concurrently on two streams into the same zeroed device image, which then returns
// with one strided copy. Bytes neither encoder writes (the ECC area of video sectors,
// cdrom.c:98 "TODO: ECC") are zero.
extern "C" int psxb200_strcd_encode_host(psxb200_bs_encoder_t *enc, int n_files, int frames_per_file, const uint8_t *h_frames,
                                         const psxb200_str_params_t *params, int xa_frequency, int xa_bits, int xa_stereo,
                                         const int16_t *h_pcm, long pcm_stride, int samples_per_file, void *h_xa_states,
                                         uint8_t *h_images, long long image_stride, psxb200_bs_result_t *h_results) {
	if (!enc) return fail("psxb200_strcd_encode_host: NULL encoder");
	if (n_files <= 0 || frames_per_file <= 0) return 0;
	if (!params || !h_frames || !h_images || !h_results) return fail("psxb200_strcd_encode_host: NULL argument");
	if (params->format != FORMAT_STR && params->format != FORMAT_STRCD)
		return fail("psxb200_strcd_encode_host: format must be FORMAT_STR or FORMAT_STRCD");
	if (xa_bits != 4 && xa_bits != 8) return fail("psxb200_strcd_encode_host: xa_bits must be 4 or 8");
	psxb200_str_params_t p = *params;
	p.place_at_lba = 1;
	p.framing = 1;
	p.lba_origin = 0;
	p.first_frame_index = 1;
	p.frames_per_file = frames_per_file;
	const bool audio = h_pcm && samples_per_file > 0 && p.interleave > 1;
	if (!audio) p.interleave = 1;
	BsStrLayout batch;
	p.file_stride = 4;   // placeholder, replaced per group below
	if (str_layout(enc, &p, &batch)) return -1;
	DeviceGuard guard(enc->device);
	CU_TRY(guard.status);
	for (BsSlot &s : enc->slots)
		if (slot_prepare(enc, s)) return -1;
	for (BsSlot &s : enc->slots) {
		if (!s.audio_stream) CU_TRY(cudaStreamCreateWithFlags(&s.audio_stream, cudaStreamNonBlocking));
		if (!s.audio_done) CU_TRY(cudaEventCreateWithFlags(&s.audio_done, cudaEventDisableTiming));
		if (!s.image_ready) CU_TRY(cudaEventCreateWithFlags(&s.image_ready, cudaEventDisableTiming));
	}
	const uint32_t *edc = edc_tables_device();
	if (!edc) return fail("psxb200_strcd_encode_host: EDC tables unavailable");
	SlotDrain drain(enc);

	const int ss = batch.sector_size;
	const int xa_format = p.format == FORMAT_STRCD ? 1 : 0;
	long long v_lo, v_hi;
	str_slot_range(batch, 1, frames_per_file, &v_lo, &v_hi);
	const int audio_sectors = audio ? adpcm_xa_sectors(xa_stereo, xa_bits, samples_per_file) : 0;
	const int audio_slot0 = batch.audio_first ? 0 : batch.interleave - 1;
	long long slots = v_hi;
	if (audio_sectors) slots = std::max(slots, (long long)audio_slot0 + (long long)(audio_sectors - 1) * batch.interleave + 1);
	const size_t image_bytes = (size_t)slots * ss;
	if ((long long)image_bytes > image_stride && n_files > 1) return fail("psxb200_strcd_encode_host: image_stride %lld < %zu", image_stride, image_bytes);
	const size_t dev_stride = round_up(image_bytes, 16);
	// int16 elements of a file's PCM the reference reads (adpcm.c:193-233 may run a little past
	// 2 * samples_per_file in the tail group of a stereo stream; the same bytes are read here)
	const long pcm_extent = audio ? adpcm_xa_input_extent(xa_stereo, xa_bits, samples_per_file) : 0;
	const long pcm_dev_stride = (long)round_up((size_t)pcm_extent + 224, 8);   // zero tail: the last sound group may read past the end

	// Files per chunk: a chunk's XA chains take their full serial latency (72 units per sector and
	// channel) however few files it holds, so chunks are made large enough — four host chunks,
	// 1024 frames by default — for the copies of the next chunk to cover it.
	int group_frames = 4 * enc->host_chunk;
	if (const char *env = getenv("PSXB200_STRCD_GROUP_FRAMES")) group_frames = std::max(1, atoi(env));
	const int group = std::max(1, std::min(n_files, group_frames / frames_per_file));
	const int bound = str_max_budget(batch);
	struct Pending { int file0, files; };
	Pending pending[BS_SLOTS] = {};
	bool busy[BS_SLOTS] = {};
	auto finish = [&](int slot) -> int {
		if (!busy[slot]) return 0;
		BsSlot &s = enc->slots[slot];
		CU_TRY(cudaStreamSynchronize(s.stream));
		memcpy(h_results + (size_t)pending[slot].file0 * frames_per_file, s.h_res.ptr,
		       (size_t)pending[slot].files * frames_per_file * sizeof(psxb200_bs_result_t));
		if (audio && h_xa_states)
			memcpy((uint8_t *)h_xa_states + (size_t)pending[slot].file0 * 48, s.h_states.ptr, (size_t)pending[slot].files * 48);
		busy[slot] = false;
		return 0;
	};

	int chunk = 0;
	for (int file0 = 0; file0 < n_files; file0 += group, chunk++) {
		const int slot = chunk % BS_SLOTS;
		BsSlot &s = enc->slots[slot];
		if (finish(slot)) return -1;
		const int files = std::min(group, n_files - file0);
		const int m = files * frames_per_file;
		CU_TRY(s.in.reserve((size_t)group * frames_per_file * enc->frame_bytes));
		CU_TRY(s.out.reserve(dev_stride * group + 16));
		CU_TRY(s.res.reserve((size_t)group * frames_per_file));
		CU_TRY(s.h_res.reserve((size_t)group * frames_per_file));
		// the image starts out zeroed: both encoders leave some bytes alone
		CU_TRY(cudaMemsetAsync(s.out.ptr, 0, dev_stride * files, s.stream));
		CU_TRY(cudaEventRecord(s.image_ready, s.stream));
		CU_TRY(cudaMemcpyAsync(s.in.ptr, h_frames + (size_t)file0 * frames_per_file * enc->frame_bytes, (size_t)m * enc->frame_bytes,
		                       cudaMemcpyHostToDevice, s.stream));
		if (audio) {
			// XA chains of this group on the audio stream, concurrent with the video kernels
			cudaStream_t as = s.audio_stream;
			CU_TRY(s.pcm.reserve((size_t)group * pcm_dev_stride));
			CU_TRY(s.states.reserve((size_t)group * 48));
			CU_TRY(s.h_states.reserve((size_t)group * 48));
			CU_TRY(cudaStreamWaitEvent(as, s.image_ready, 0));
			CU_TRY(cudaMemsetAsync(s.pcm.ptr, 0, (size_t)files * pcm_dev_stride * sizeof(int16_t), as));
			CU_TRY(cudaMemcpy2DAsync(s.pcm.ptr, (size_t)pcm_dev_stride * 2, h_pcm + (size_t)file0 * pcm_stride,
			                         (size_t)(files > 1 ? pcm_stride : pcm_extent) * 2, (size_t)pcm_extent * 2, files,
			                         cudaMemcpyHostToDevice, as));
			if (h_xa_states) {
				memcpy(s.h_states.ptr, (const uint8_t *)h_xa_states + (size_t)file0 * 48, (size_t)files * 48);
				CU_TRY(cudaMemcpyAsync(s.states.ptr, s.h_states.ptr, (size_t)files * 48, cudaMemcpyHostToDevice, as));
			} else {
				CU_TRY(cudaMemsetAsync(s.states.ptr, 0, (size_t)files * 48, as));
			}
			// sector k of a file's audio goes to slot audio_slot0 + k * interleave and carries that LBA
			uint8_t *a_out = s.out.ptr + (size_t)audio_slot0 * ss;
			CU_TRY(adpcm_launch_xa(files, xa_format, xa_stereo, xa_frequency, xa_bits, p.xa_file, p.xa_channel, s.pcm.ptr,
			                       pcm_dev_stride, samples_per_file, audio_slot0, batch.interleave, s.states.ptr, a_out,
			                       (long)dev_stride, (long)batch.interleave * ss, edc, as));
			g_launches += 2;
			CU_TRY(cudaMemcpyAsync(s.h_states.ptr, s.states.ptr, (size_t)files * 48, cudaMemcpyDeviceToHost, as));
			CU_TRY(cudaEventRecord(s.audio_done, as));
		}
		BsStrLayout l = batch;
		l.file_stride = (long long)dev_stride;
		l.slot0 = 0;
		if (bs_encode_chunked(enc, s.coefs, s.gstream, m, s.in.ptr, nullptr, bound, s.out.ptr, 0, s.res.ptr, s.stream, &l)) return -1;
		if (audio) CU_TRY(cudaStreamWaitEvent(s.stream, s.audio_done, 0));
		CU_TRY(cudaMemcpy2DAsync(h_images + (size_t)file0 * (size_t)image_stride, files > 1 ? (size_t)image_stride : image_bytes,
		                         s.out.ptr, dev_stride, image_bytes, files, cudaMemcpyDeviceToHost, s.stream));
		CU_TRY(cudaMemcpyAsync(s.h_res.ptr, s.res.ptr, (size_t)m * sizeof(psxb200_bs_result_t), cudaMemcpyDeviceToHost, s.stream));
		pending[slot] = Pending{file0, files};
		busy[slot] = true;
	}
	for (int slot = 0; slot < BS_SLOTS; slot++)
		if (finish(slot)) return -1;
	drain.armed = false;
	int failed = 0;
	for (long long i = 0; i < (long long)n_files * frames_per_file; i++) failed += h_results[i].quant_scale >= 64;
	return failed;
}

extern "C" long long psxb200_strcd_image_bytes(const psxb200_str_params_t *params, int frames_per_file, int xa_bits, int xa_stereo,
                                               int samples_per_file) {
	if (!params || params->sectors_num < 1 || params->sectors_den < 1) return -1;
	BsStrLayout l{};
	l.sectors_num = params->sectors_num;
	l.sectors_den = params->sectors_den;
	const bool audio = samples_per_file > 0 && params->interleave > 1;
	l.interleave = audio ? params->interleave : 1;
	l.audio_first = params->trailing_audio ? 0 : 1;
	l.place_at_lba = 1;
	long long lo, hi;
	str_slot_range(l, 1, frames_per_file, &lo, &hi);
	long long slots = hi;
	if (audio) {
		const int sectors = adpcm_xa_sectors(xa_stereo, xa_bits, samples_per_file);
		const int slot0 = l.audio_first ? 0 : l.interleave - 1;
		if (sectors) slots = std::max(slots, (long long)slot0 + (long long)(sectors - 1) * l.interleave + 1);
	}
	const int ss = params->format == FORMAT_STRCD ? 2352 : params->format == FORMAT_STR ? 2336 : 2048;
	return slots * ss;
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
	psxb200_bs_encoder_t *enc = psxb200_bs_create((int)video_codec, video_width, video_height, dropin_fdct_variant(), 16);
	state->dct_context = enc;
	if (!enc) {
		fprintf(stderr, "libpsxav_b200: init_mdec_encoder: %s\n", g_error);
		return false;
	}
	// Speculative look-ahead of encode_sector_str (see there); PSXB200_STR_LOOKAHEAD=0 turns it off.
	const char *env = getenv("PSXB200_STR_LOOKAHEAD");
	enc->ahead.enabled = !(env && atoi(env) == 0);
	return true;
}

extern "C" void destroy_mdec_encoder(mdec_encoder_t *encoder) {
	psxb200_bs_destroy(static_cast<psxb200_bs_encoder_t *>(encoder->state.dct_context));
	encoder->state.dct_context = nullptr;
}

static psxb200_bs_encoder_t *dropin_encoder(mdec_encoder_t *encoder, const char *who) {
	auto *enc = static_cast<psxb200_bs_encoder_t *>(encoder->state.dct_context);
	if (!enc) {
		fail("encoder not initialised (init_mdec_encoder failed or was not called)");
		die(who);
	}
	return enc;
}

static void dropin_store_result(mdec_encoder_state_t *state, const psxb200_bs_result_t &r, int max_size) {
	if (r.quant_scale >= 64) {
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

extern "C" void encode_frame_bs(mdec_encoder_t *encoder, const uint8_t *video_frame) {
	mdec_encoder_state_t *state = &encoder->state;
	psxb200_bs_encoder_t *enc = dropin_encoder(encoder, "encode_frame_bs");
	psxb200_bs_result_t r;
	int max_size = state->frame_max_size;
	int rc = psxb200_bs_encode_host(enc, 1, video_frame, &max_size, state->frame_output, (size_t)max_size, &r);
	if (rc < 0) die("encode_frame_bs");
	dropin_store_result(state, r, max_size);
}

// Look-ahead of encode_sector_str. The function is handed the head of the caller's frame queue
// and the reference's mux loops keep at least two frames in it (frames_needed >= 2,
// filefmt.c:443-446, 565-568; decoding.c:448-451 always allocates one slot more than it holds),
// stored 1.5*W*H bytes apart (decoding.c:317, 463). When a frame is encoded, the frame behind it
// is therefore copied to a pinned staging buffer and its encode is started on a side stream with
// the budget the accumulator will give it. The next call that needs a new frame compares the
// frame it is handed with the staged copy (all bytes) and its budget with the predicted one; on
// a match the finished result is taken, otherwise the speculation is dropped and the frame is
// encoded synchronously. Output bytes are the same either way.
static int lookahead_enqueue(psxb200_bs_encoder *enc, int max_size) {
	BsLookahead &a = enc->ahead;
	CU_TRY(cudaMemcpyAsync(a.in.ptr, a.staged.ptr, enc->frame_bytes, cudaMemcpyHostToDevice, a.stream));
	if (bs_encode_chunked(enc, a.coefs, a.gstream, 1, a.in.ptr, nullptr, max_size, a.out.ptr, round_up((size_t)max_size, 16),
	                      a.res.ptr, a.stream))
		return -1;
	CU_TRY(cudaMemcpyAsync(a.h_out.ptr, a.out.ptr, (size_t)max_size, cudaMemcpyDeviceToHost, a.stream));
	CU_TRY(cudaMemcpyAsync(a.h_res.ptr, a.res.ptr, sizeof(psxb200_bs_result_t), cudaMemcpyDeviceToHost, a.stream));
	return 0;
}

static int lookahead_start(psxb200_bs_encoder *enc, const uint8_t *frame, int max_size) {
	BsLookahead &a = enc->ahead;
	a.valid = false;
	if (max_size < 8) return 0;
	if (!a.stream) CU_TRY(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
	// buffers are sized for the largest budget seen so far; growing one invalidates the graphs
	// (they hold the old addresses)
	const size_t want_out = round_up((size_t)max_size, 16);
	if (a.h_out.cap < (size_t)max_size || a.out.cap < want_out || !a.staged.ptr) {
		for (auto &g : a.graphs) {
			if (g.exec) cudaGraphExecDestroy(g.exec);
			g = BsLookahead::Graph();
		}
		CU_TRY(cudaStreamSynchronize(a.stream));
		CU_TRY(a.staged.reserve(enc->frame_bytes));
		CU_TRY(a.h_out.reserve((size_t)max_size));
		CU_TRY(a.h_res.reserve(1));
		CU_TRY(a.in.reserve(enc->frame_bytes));
		CU_TRY(a.out.reserve(want_out));
		CU_TRY(a.res.reserve(1));
	}
	memcpy(a.staged.ptr, frame, enc->frame_bytes);
	a.max_size = max_size;

	BsLookahead::Graph *slot = nullptr;
	for (auto &g : a.graphs)
		if (g.max_size == max_size) slot = &g;
	if (!slot)
		for (auto &g : a.graphs)
			if (!slot && g.max_size == 0) slot = &g;
	if (slot && !enc->timing) {
		slot->max_size = max_size;
		if (!slot->exec && slot->uses++ > 0) {
			// second speculation with this budget (the first one ran eagerly and sized every
			// scratch buffer): record the sequence once
			cudaGraph_t graph = nullptr;
			CU_TRY(cudaStreamBeginCapture(a.stream, cudaStreamCaptureModeThreadLocal));
			const int rc = lookahead_enqueue(enc, max_size);
			cudaError_t e = cudaStreamEndCapture(a.stream, &graph);
			if (rc || e != cudaSuccess) {
				if (graph) cudaGraphDestroy(graph);
				cudaGetLastError();
				slot->uses = -1000000;   // do not try again; run eagerly
			} else {
				e = cudaGraphInstantiate(&slot->exec, graph, 0);
				cudaGraphDestroy(graph);
				if (e != cudaSuccess) {
					slot->exec = nullptr;
					slot->uses = -1000000;
					cudaGetLastError();
				}
			}
		}
		if (slot->exec) {
			CU_TRY(cudaGraphLaunch(slot->exec, a.stream));
			a.valid = true;
			return 0;
		}
	}
	if (lookahead_enqueue(enc, max_size)) return -1;
	a.valid = true;
	return 0;
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
		const uint8_t *frame = video_frames + consumed * frame_advance;
		consumed++;

		psxb200_bs_encoder_t *enc = dropin_encoder(encoder, "encode_sector_str");
		BsLookahead &a = enc->ahead;
		if (!a.enabled || st->frame_max_size < 8) {
			encode_frame_bs(encoder, frame);
			continue;
		}
		DeviceGuard guard(enc->device);
		if (a.valid && a.max_size == st->frame_max_size && memcmp(a.staged.ptr, frame, enc->frame_bytes) == 0) {
			if (cudaStreamSynchronize(a.stream) != cudaSuccess) {
				fail("look-ahead stream: %s", cudaGetErrorString(cudaGetLastError()));
				die("encode_sector_str");
			}
			memcpy(st->frame_output, a.h_out.ptr, (size_t)st->frame_max_size);
			dropin_store_result(st, a.h_res.ptr[0], st->frame_max_size);
			a.hits++;
		} else {
			if (a.valid) {
				cudaStreamSynchronize(a.stream);   // drop a stale speculation before its buffers are reused
				a.misses++;
			}
			a.valid = false;
			encode_frame_bs(encoder, frame);
		}
		// speculate on the frame behind this one, with the budget the accumulator will hand out
		// next — only while the loop is not going to consume another frame right away
		int num = st->frame_block_overflow_num + st->frame_block_base_overflow;
		const int next_size = st->frame_block_overflow_den > 0 ? num / st->frame_block_overflow_den * 2016 : 0;
		if (st->frame_max_size > 0 && next_size >= 8) {
			if (lookahead_start(enc, frame + enc->frame_bytes, next_size)) die("encode_sector_str (look-ahead)");
		} else {
			a.valid = false;
		}
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

extern "C" void psxb200_bs_lookahead_stats(const psxb200_bs_encoder_t *enc, long long *hits, long long *misses) {
	if (hits) *hits = enc ? enc->ahead.hits : 0;
	if (misses) *misses = enc ? enc->ahead.misses : 0;
}
