// The BS encoder handle behind psxb200_bs_encoder_t (shared by capi_bs.cu and capi_multi.cu).
#pragma once

#include <cuda_runtime.h>

#include <vector>

#include "psxav_b200.h"
#include "bs_encode.h"
#include "capi_util.h"

// host-API pipeline depth: chunks in flight per encoder
constexpr int BS_SLOTS = 3;

// One pipeline slot of the host entry points: a stream with its own scratch, so that the
// kernels and copies of consecutive chunks overlap without sharing anything.
struct BsSlot {
	cudaStream_t stream = nullptr;
	cudaEvent_t res_ready = nullptr;
	psxb200::DeviceBuffer<uint4> coefs;          // coefficient plane of the chunk
	psxb200::DeviceBuffer<uint32_t> gstream;     // bitstream images for budgets beyond shared memory
	psxb200::DeviceBuffer<uint8_t> in, out;
	psxb200::DeviceBuffer<int> sizes;
	psxb200::DeviceBuffer<psxb200_bs_result_t> res;
	psxb200::PinnedBuffer<psxb200_bs_result_t> h_res;
	// psxb200_strcd_encode_host: the chunk's XA chains run beside its video kernels on a stream of
	// their own (XA chains of different chunks overlap as well: a chain is latency-bound)
	cudaStream_t audio_stream = nullptr;
	cudaEvent_t audio_done = nullptr, image_ready = nullptr;
	psxb200::DeviceBuffer<int16_t> pcm;
	psxb200::DeviceBuffer<uint8_t> states;
	psxb200::PinnedBuffer<uint8_t> h_states;
};

// Speculative next-frame encode of the drop-in encode_sector_str (capi_bs.cu).
struct BsLookahead {
	bool enabled = false, valid = false;
	int max_size = 0;
	long long hits = 0, misses = 0;
	cudaStream_t stream = nullptr;
	psxb200::PinnedBuffer<uint8_t> staged, h_out;
	psxb200::PinnedBuffer<psxb200_bs_result_t> h_res;
	psxb200::DeviceBuffer<uint8_t> in, out;
	psxb200::DeviceBuffer<psxb200_bs_result_t> res;
	psxb200::DeviceBuffer<uint4> coefs;
	psxb200::DeviceBuffer<uint32_t> gstream;
	// the copy-in / kernels / copy-out sequence of one speculation as an instantiated CUDA graph
	// per byte budget (all its pointers are the fixed staging buffers above): one launch call
	// instead of six, and no gaps between the dependent operations on the device
	struct Graph {
		int max_size = 0, uses = 0;
		cudaGraphExec_t exec = nullptr;
	} graphs[4];
};

struct psxb200_bs_encoder {
	int codec, width, height, fdct, max_batch, host_chunk, pack_threads;
	int pack_min_ctas = 0;              // 0: by shared-memory fit
	int sm_count = 0, device = 0;
	bool pack_threads_forced = false;   // PSXB200_PACK_THREADS given: no small-batch override
	size_t frame_bytes;
	psxb200::BsGeometry geo;
	// scratch of the device-pointer entry points (they run on the caller's stream)
	psxb200::DeviceBuffer<uint4> coefs;
	psxb200::DeviceBuffer<uint32_t> gstream;
	// device-pointer entry points, batches of more than one launch: the launches alternate
	// between two internal streams forked from (and joined back into) the caller's stream, so
	// that the FDCT kernel of one launch fills the SMs the pack kernel of the previous one leaves
	// idle in its last wave, and the coefficient plane of a launch (reused every other launch)
	// stays in L2 between the two kernels
	cudaStream_t pipe[2] = {nullptr, nullptr};
	cudaEvent_t pipe_fork = nullptr, pipe_join[2] = {nullptr, nullptr};
	psxb200::DeviceBuffer<uint4> coefs2;
	psxb200::DeviceBuffer<uint32_t> gstream2;
	int device_pipeline = 1;            // PSXB200_DEVICE_PIPELINE=0 turns it off
	BsSlot slots[BS_SLOTS];
	BsLookahead ahead;
	// optional per-kernel timing (psxb200_bs_timing_*): three events per internal launch pair
	bool timing = false;
	std::vector<cudaEvent_t> events;
	size_t events_used = 0;
	cudaError_t mark(cudaStream_t st);

	psxb200_bs_encoder(int c, int w, int h, int f, int mb)
		: codec(c), width(w), height(h), fdct(f), max_batch(mb), host_chunk(mb < 256 ? mb : 256), pack_threads(320),
		  frame_bytes((size_t)w * h * 3 / 2), geo(w, h) {}
};
