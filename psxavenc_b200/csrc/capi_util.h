// Shared host-side helpers of the C ABI (capi_*.cu): error reporting, device selection, buffers.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace psxb200 {

// Last error message of the calling thread (psxb200_last_error).
extern thread_local char g_error[512];
// Kernels launched by this library since load (psxb200_launch_count).
extern std::atomic<unsigned long long> g_launches;

int fail(const char *fmt, ...);
[[noreturn]] void die(const char *what);

#define CU_TRY(expr)                                                                                  \
	do {                                                                                              \
		cudaError_t e_ = (expr);                                                                      \
		if (e_ != cudaSuccess)                                                                        \
			return ::psxb200::fail("%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

// capi_audio.cu: SPU chains first, first + step, ... of n_streams on the current device
int spu_encode_host_subset(int n_streams, int first, int step, const int16_t *h_samples, int pitch, long group_stride,
                           int sample_count, void *h_states, uint8_t *h_out, long out_stride);

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Makes `device` current for the calling thread for the lifetime of the guard.
struct DeviceGuard {
	int prev = -1;
	bool switched = false;
	cudaError_t status = cudaSuccess;
	explicit DeviceGuard(int device) {
		status = cudaGetDevice(&prev);
		if (status == cudaSuccess && prev != device) {
			status = cudaSetDevice(device);
			switched = status == cudaSuccess;
		}
	}
	~DeviceGuard() {
		if (switched) cudaSetDevice(prev);
	}
	DeviceGuard(const DeviceGuard &) = delete;
	DeviceGuard &operator=(const DeviceGuard &) = delete;
};

// Error return of a host entry point with copies still queued on `stream`: they touch the
// caller's buffers, so the stream is drained before the call goes back.
struct StreamDrain {
	cudaStream_t stream;
	bool armed = true;
	explicit StreamDrain(cudaStream_t st) : stream(st) {}
	~StreamDrain() {
		if (!armed) return;
		cudaStreamSynchronize(stream);
		cudaGetLastError();
	}
	StreamDrain(const StreamDrain &) = delete;
	StreamDrain &operator=(const StreamDrain &) = delete;
};

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

// Page-locked host memory, mapped into the device address space (zero-copy for tiny transfers).
template <typename T>
struct PinnedBuffer {
	T *ptr = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t n) {
		if (n <= cap) return cudaSuccess;
		if (ptr) cudaFreeHost(ptr);
		ptr = nullptr;
		cap = 0;
		cudaError_t e = cudaHostAlloc(&ptr, n * sizeof(T), cudaHostAllocPortable | cudaHostAllocMapped);
		if (e == cudaSuccess) cap = n;
		return e;
	}
	T *device_ptr() const {
		void *d = nullptr;
		if (!ptr || cudaHostGetDevicePointer(&d, ptr, 0) != cudaSuccess) return nullptr;
		return static_cast<T *>(d);
	}
	void release() {
		if (ptr) cudaFreeHost(ptr);
		ptr = nullptr;
		cap = 0;
	}
};

}  // namespace psxb200
