// Internal launch interface of the BS encoder kernels (bs_encode.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "psxav_b200.h"

namespace psxb200 {

constexpr int BS_DCT_THREADS = 128;
constexpr int BS_PACK_MAX_THREADS = 640;
// per block in the coefficient plane: 8 uint4 of |coef| (u16 pairs, zig-zag order) + 1 uint4
// holding the 64-bit sign mask
constexpr int BS_U4_PER_BLOCK = 9;
// bitstream images above this size are built in global memory instead of shared memory
constexpr int BS_SMEM_STREAM_LIMIT = 96 * 1024;

struct BsGeometry {
	int mbw, mbh;            // macroblocks per row / column
	int nblk;                // 8x8 blocks per frame = 6 * mbw * mbh
	int ngroups;             // ceil(nblk / 32)
	size_t frame_stride_u4;  // coefficient plane stride between frames, in uint4

	BsGeometry(int width, int height)
		: mbw(width / 16), mbh(height / 16), nblk(6 * (width / 16) * (height / 16)),
		  ngroups((6 * (width / 16) * (height / 16) + 31) / 32),
		  frame_stride_u4((size_t)((6 * (width / 16) * (height / 16) + 31) / 32) * BS_U4_PER_BLOCK * 32) {}
};

void bs_upload_tables();
size_t bs_pack_smem_bytes(bool v3, bool smem_stream, int ngroups, int max_size_bound);

cudaError_t bs_launch_dct(int fdct_variant, const uint8_t *d_frames, size_t frame_bytes, int n, int width, int height,
                          const BsGeometry &geo, uint4 *d_coefs, cudaStream_t stream);

// d_gstream == nullptr: bitstream image in shared memory (max_size_bound <= BS_SMEM_STREAM_LIMIT)
cudaError_t bs_launch_pack(int codec, int threads, int n, const uint4 *d_coefs, const BsGeometry &geo,
                           const int *d_max_sizes, int max_size_bound, uint8_t *d_out, size_t out_stride,
                           psxb200_bs_result_t *d_results, uint32_t *d_gstream, size_t gstream_stride,
                           cudaStream_t stream);

}  // namespace psxb200
