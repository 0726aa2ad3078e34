// C ABI of libpsxav_b200.so, multi-device half (declared in include/psxav_b200.h): ONE process
// driving every visible GPU, which is how the reference's single-process C host (filefmt.c)
// reaches more than one device. No collective: frames are independent
// (mdec.c:676-686 resets all per-frame state) and ADPCM chains are independent of each other,
// so the work is dealt out in contiguous frame ranges / whole chains and every device's results
// land directly in the caller's host arrays.
//
// psxb200_bs_multi_t owns one psxb200_bs_encoder_t and one worker thread per device (the worker
// keeps its device current and runs the single-device host pipeline of capi_bs.cu on its share).
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "psxav_b200.h"
#include "adpcm_encode.h"
#include "capi_util.h"

using namespace psxb200;

namespace {

// Runs fn(i) for i in [0, n) on n threads, thread i with device ids[i] current; returns the
// first failure (its message becomes the caller's last error).
int run_on_devices(const std::vector<int> &ids, const std::function<int(int)> &fn) {
	const int n = (int)ids.size();
	std::vector<int> rc(n, 0);
	std::vector<std::string> err(n);
	std::vector<std::thread> threads;
	threads.reserve(n);
	for (int i = 0; i < n; i++) {
		threads.emplace_back([&, i] {
			if (cudaSetDevice(ids[i]) != cudaSuccess) {
				err[i] = std::string("cudaSetDevice: ") + cudaGetErrorString(cudaGetLastError());
				rc[i] = -1;
				return;
			}
			rc[i] = fn(i);
			if (rc[i] < 0) err[i] = g_error;
		});
	}
	for (auto &t : threads) t.join();
	int total = 0;
	for (int i = 0; i < n; i++) {
		if (rc[i] < 0) return fail("device %d: %s", ids[i], err[i].c_str());
		total += rc[i];
	}
	return total;
}

int device_list(int n_devices, const int *device_ids, std::vector<int> *out) {
	const int have = psxb200_device_count();
	if (have == 0) return fail("no CUDA device (this library has no CPU path)");
	if (n_devices <= 0) n_devices = have;
	out->clear();
	for (int i = 0; i < n_devices; i++) {
		const int d = device_ids ? device_ids[i] : i;
		if (d < 0 || d >= have) return fail("device %d out of range (%d visible)", d, have);
		out->push_back(d);
	}
	return 0;
}

}  // namespace

struct psxb200_bs_multi {
	std::vector<int> devices;
	std::vector<psxb200_bs_encoder_t *> encoders;
	// persistent workers: one per device
	std::vector<std::thread> workers;
	std::mutex lock;
	std::condition_variable wake, done;
	std::function<int(int)> job;
	unsigned long long generation = 0;
	int outstanding = 0;
	bool quit = false;
	std::vector<int> rc;
	std::vector<std::string> err;

	void worker(int i) {
		cudaSetDevice(devices[i]);
		unsigned long long seen = 0;
		for (;;) {
			std::unique_lock<std::mutex> hold(lock);
			wake.wait(hold, [&] { return quit || generation != seen; });
			if (quit) return;
			seen = generation;
			hold.unlock();
			int r = job(i);
			hold.lock();
			rc[i] = r;
			err[i] = r < 0 ? g_error : "";
			if (--outstanding == 0) done.notify_all();
		}
	}

	// runs fn(i) on every worker and waits; returns the sum of the results or -1
	int run(const std::function<int(int)> &fn) {
		std::unique_lock<std::mutex> hold(lock);
		job = fn;
		outstanding = (int)devices.size();
		generation++;
		wake.notify_all();
		done.wait(hold, [&] { return outstanding == 0; });
		int total = 0;
		for (size_t i = 0; i < devices.size(); i++) {
			if (rc[i] < 0) return fail("device %d: %s", devices[i], err[i].c_str());
			total += rc[i];
		}
		return total;
	}
};

extern "C" psxb200_bs_multi_t *psxb200_bs_multi_create(int codec, int width, int height, int fdct_variant, int max_batch,
                                                       int n_devices, const int *device_ids) {
	auto *m = new psxb200_bs_multi;
	if (device_list(n_devices, device_ids, &m->devices)) {
		delete m;
		return nullptr;
	}
	const int n = (int)m->devices.size();
	m->encoders.assign(n, nullptr);
	m->rc.assign(n, 0);
	m->err.assign(n, "");
	int prev = 0;
	cudaGetDevice(&prev);
	bool ok = true;
	for (int i = 0; i < n && ok; i++) {
		ok = cudaSetDevice(m->devices[i]) == cudaSuccess;
		if (!ok) fail("cudaSetDevice(%d): %s", m->devices[i], cudaGetErrorString(cudaGetLastError()));
		if (ok) m->encoders[i] = psxb200_bs_create(codec, width, height, fdct_variant, max_batch);
		ok = ok && m->encoders[i];
	}
	cudaSetDevice(prev);
	if (!ok) {
		for (auto *e : m->encoders) psxb200_bs_destroy(e);
		delete m;
		return nullptr;
	}
	for (int i = 0; i < n; i++) m->workers.emplace_back(&psxb200_bs_multi::worker, m, i);
	return m;
}

extern "C" void psxb200_bs_multi_destroy(psxb200_bs_multi_t *m) {
	if (!m) return;
	{
		std::lock_guard<std::mutex> hold(m->lock);
		m->quit = true;
	}
	m->wake.notify_all();
	for (auto &t : m->workers) t.join();
	for (auto *e : m->encoders) psxb200_bs_destroy(e);
	delete m;
}

extern "C" int psxb200_bs_multi_device_count(const psxb200_bs_multi_t *m) { return m ? (int)m->devices.size() : 0; }

// contiguous share of `n` units for worker i of g (shares differ by at most one unit)
static void share(long long n, int i, int g, long long *first, long long *count) {
	const long long base = n / g, extra = n % g;
	*first = i * base + std::min<long long>(i, extra);
	*count = base + (i < extra ? 1 : 0);
}

extern "C" int psxb200_bs_multi_encode_host(psxb200_bs_multi_t *m, int n, const uint8_t *h_frames, const int *h_max_sizes,
                                            uint8_t *h_out, size_t out_stride, psxb200_bs_result_t *h_results) {
	if (!m) return fail("psxb200_bs_multi_encode_host: NULL handle");
	if (n <= 0) return 0;
	if (!h_frames || !h_max_sizes || !h_out || !h_results) return fail("psxb200_bs_multi_encode_host: NULL argument");
	const int g = (int)m->devices.size();
	const size_t frame_bytes = (size_t)psxb200_bs_frame_bytes(m->encoders[0]);
	return m->run([&](int i) -> int {
		long long first, count;
		share(n, i, g, &first, &count);
		if (count == 0) return 0;
		return psxb200_bs_encode_host(m->encoders[i], (int)count, h_frames + (size_t)first * frame_bytes, h_max_sizes + first,
		                              h_out + (size_t)first * out_stride, out_stride, h_results + first);
	});
}

extern "C" int psxb200_bs_multi_str_encode_host(psxb200_bs_multi_t *m, int n, const uint8_t *h_frames,
                                                const psxb200_str_params_t *params, uint8_t *h_sectors,
                                                psxb200_bs_result_t *h_results) {
	if (!m) return fail("psxb200_bs_multi_str_encode_host: NULL handle");
	if (n <= 0) return 0;
	if (!params || !h_frames || !h_sectors || !h_results) return fail("psxb200_bs_multi_str_encode_host: NULL argument");
	const int g = (int)m->devices.size();
	const size_t frame_bytes = (size_t)psxb200_bs_frame_bytes(m->encoders[0]);
	const int fpf = params->frames_per_file;
	if (fpf > 0 && n % fpf) return fail("psxb200_bs_multi_str_encode_host: n is not a multiple of frames_per_file");
	const int ss = params->format == FORMAT_STRCD ? 2352 : params->format == FORMAT_STR ? 2336 : 2048;
	return m->run([&](int i) -> int {
		psxb200_str_params_t p = *params;
		long long first, count;
		uint8_t *dst = h_sectors;
		if (fpf > 0) {
			// whole files per device
			share(n / fpf, i, g, &first, &count);
			dst += (size_t)first * (size_t)p.file_stride;
			first *= fpf;
			count *= fpf;
		} else {
			share(n, i, g, &first, &count);
			if (count == 0) return 0;
			// this device's frames start `first` frames into the stream: their first sector's position
			long long lo = 0;
			if (psxb200_str_slot_range(params, (int)first, nullptr, &lo)) return -1;
			if (first == 0) lo = 0;
			p.first_frame_index += (int)first;
			if (!p.place_at_lba) dst += (size_t)lo * ss;   // place_at_lba: positions are absolute (lba_origin)
		}
		if (count == 0) return 0;
		return psxb200_str_encode_host_ex(m->encoders[i], (int)count, h_frames + (size_t)first * frame_bytes, &p, dst,
		                                  h_results + first);
	});
}

extern "C" int psxb200_bs_multi_strcd_encode_host(psxb200_bs_multi_t *m, int n_files, int frames_per_file, const uint8_t *h_frames,
                                                  const psxb200_str_params_t *params, int xa_frequency, int xa_bits, int xa_stereo,
                                                  const int16_t *h_pcm, long pcm_stride, int samples_per_file, void *h_xa_states,
                                                  uint8_t *h_images, long long image_stride, psxb200_bs_result_t *h_results) {
	if (!m) return fail("psxb200_bs_multi_strcd_encode_host: NULL handle");
	if (n_files <= 0) return 0;
	const int g = (int)m->devices.size();
	const size_t frame_bytes = (size_t)psxb200_bs_frame_bytes(m->encoders[0]);
	return m->run([&](int i) -> int {
		long long first, count;
		share(n_files, i, g, &first, &count);
		if (count == 0) return 0;
		return psxb200_strcd_encode_host(m->encoders[i], (int)count, frames_per_file,
		                                 h_frames + (size_t)first * frames_per_file * frame_bytes, params, xa_frequency, xa_bits,
		                                 xa_stereo, h_pcm ? h_pcm + (size_t)first * pcm_stride : nullptr, pcm_stride, samples_per_file,
		                                 h_xa_states ? (uint8_t *)h_xa_states + (size_t)first * 48 : nullptr,
		                                 h_images + (size_t)first * (size_t)image_stride, image_stride,
		                                 h_results + (size_t)first * frames_per_file);
	});
}

// ---- ADPCM ---------------------------------------------------------------------------------

extern "C" int psxb200_spu_encode_host_multi(int n_devices, const int *device_ids, int n_streams, const int16_t *h_samples,
                                             int pitch, long group_stride, int sample_count, void *h_states, uint8_t *h_out,
                                             long out_stride) {
	if (n_streams <= 0 || sample_count <= 0) return 0;
	if (pitch < 1) return fail("psxb200_spu_encode_host_multi: bad pitch");
	std::vector<int> ids;
	if (device_list(n_devices, device_ids, &ids)) return -1;
	const int g = (int)ids.size();
	const int groups = (n_streams + pitch - 1) / pitch;
	if (groups >= g) {
		// many interleaved groups (vagi x B): contiguous runs of whole groups per device
		return run_on_devices(ids, [&](int i) -> int {
			long long first, count;
			share(groups, i, g, &first, &count);
			if (count == 0) return 0;
			const int s0 = (int)first * pitch;
			const int ns = std::min(n_streams, (int)(first + count) * pitch) - s0;
			return spu_encode_host_subset(ns, 0, 1, h_samples + first * group_stride, pitch, group_stride, sample_count,
			                              (uint8_t *)h_states + (size_t)s0 * 24, h_out + (size_t)s0 * out_stride, out_stride);
		});
	}
	// fewer groups than devices (one vagi file): chain c goes to device c mod G;
	// a chain cannot be cut in time (adpcm.c:135-136, 186-190)
	return run_on_devices(ids, [&](int i) -> int {
		return spu_encode_host_subset(n_streams, i, g, h_samples, pitch, group_stride, sample_count, h_states, h_out, out_stride);
	});
}

extern "C" int psxb200_xa_encode_host_multi(int n_devices, const int *device_ids, int n_streams, int format, int stereo,
                                            int frequency, int bits_per_sample, int file_number, int channel_number,
                                            const int16_t *h_samples, long in_stride, int sample_count, int lba, void *h_states,
                                            uint8_t *h_out, long out_stride) {
	if (n_streams <= 0) return 0;
	std::vector<int> ids;
	if (device_list(n_devices, device_ids, &ids)) return -1;
	const int g = (int)ids.size();
	int bytes = 0;
	int rc = run_on_devices(ids, [&](int i) -> int {
		long long first, count;
		share(n_streams, i, g, &first, &count);
		if (count == 0) return 0;
		int r = psxb200_xa_encode_host((int)count, format, stereo, frequency, bits_per_sample, file_number, channel_number,
		                               h_samples + first * in_stride, in_stride, sample_count, lba,
		                               (uint8_t *)h_states + (size_t)first * 48, h_out + first * out_stride, out_stride);
		if (r >= 0 && i == 0) bytes = r;
		return r < 0 ? -1 : 0;
	});
	return rc < 0 ? -1 : bytes;
}
