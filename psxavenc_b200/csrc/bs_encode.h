// Internal launch interface of the BS encoder kernels (bs_encode.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "psxav_b200.h"

namespace psxb200 {

// FDCT kernel CTA: 3 warps divide the usual frames' plane groups exactly (320x240: 57 groups =
// 19 CTAs, 640x480: 225 = 75) and 9 such CTAs per SM keep the 72-register budget (27 warps);
// measured: 96 x 9 0.416 ms, 64 x 14 0.416 ms, 128 x 7 0.423 ms per 4096 frames; 128 x 6
// (80 registers) and 128 x 8 (64 registers, spills) are slower (profiles/r1_sweeps.md).
#ifndef BS_DCT_THREADS_PER_CTA
#define BS_DCT_THREADS_PER_CTA 96
#endif
constexpr int BS_DCT_THREADS = BS_DCT_THREADS_PER_CTA;
#ifndef BS_DCT_MIN_CTAS
#define BS_DCT_MIN_CTAS 9
#endif
#ifndef BS_PACK_MAX_THREADS_PER_CTA
#define BS_PACK_MAX_THREADS_PER_CTA 640
#endif
constexpr int BS_PACK_MAX_THREADS = BS_PACK_MAX_THREADS_PER_CTA;
// per block in the coefficient plane: up to 8 uint4 rows of list entries ((y << 6) | zig-zag
// position, 16 bits each, highest position first, zero padded; only the rows the group's longest
// list needs are written) + 1 meta uint4: sign mask lo/hi, |DC| | list length << 16, longest list
constexpr int BS_U4_PER_BLOCK = 9;
constexpr int BS_META_ROW = 8;
// groups whose longest list reaches this many entries are stored dense (all 64 y values,
// position implicit) instead of as lists; flagged in meta.w
#ifndef BS_DENSE_MIN
#define BS_DENSE_MIN 49
#endif
constexpr unsigned BS_DENSE_FLAG = 0x100;
// emit table: rows = levels 0..40 plus one row for everything above, columns = runs 0..31 plus one
// for everything above; the extra row and column hold the escape marker
constexpr int BS_VLC_ROWS = 42, BS_VLC_COLS = 33;
// the bitstream image lives in shared memory while the CTA's total stays below this, else in
// global memory
constexpr size_t BS_SMEM_BUDGET = 200 * 1024;

// The coefficient plane of a frame holds its blocks type-major: first the chroma blocks
// (plane index 2*mb + k, k = 0 Cr / 1 Cb), padded to a whole group of 32, then the luma blocks
// (cpad + 4*mb + y, y = 0..3 for Y1..Y4); mb counts macroblocks in bitstream order (columns
// outermost, mdec.c:689-704). Every group of 32 lanes is thus of one kind: the FDCT kernel's
// gather does not diverge and the pack kernel's dead-row skipping follows the (usually much
// smoother) chroma separately from the luma.
struct BsGeometry {
	int mbw, mbh;            // macroblocks per row / column
	int nmb;                 // macroblocks per frame
	int nblk;                // 8x8 blocks per frame = 6 * nmb
	int cgroups;             // plane groups holding chroma = ceil(2 * nmb / 32)
	int ngroups;             // plane groups per frame = cgroups + ceil(4 * nmb / 32)
	int nsgroups;            // groups of 32 blocks in bitstream order = ceil(nblk / 32)
	size_t frame_stride_u4;  // coefficient plane stride between frames, in uint4

	BsGeometry(int width, int height)
		: mbw(width / 16), mbh(height / 16), nmb(mbw * mbh), nblk(6 * nmb), cgroups((2 * nmb + 31) / 32),
		  ngroups(cgroups + (4 * nmb + 31) / 32), nsgroups((nblk + 31) / 32),
		  frame_stride_u4((size_t)ngroups * BS_U4_PER_BLOCK * 32) {}
};

// Bitstream-order block index (6*mb + k, k = Cr Cb Y1 Y2 Y3 Y4) of plane index pi, -1 for padding.
__host__ __device__ inline int bs_plane_to_block(int pi, int cpad, int nmb) {
	if (pi < cpad) return pi < 2 * nmb ? 6 * (pi >> 1) + (pi & 1) : -1;
	int li = pi - cpad;
	return li < 4 * nmb ? 6 * (li >> 2) + 2 + (li & 3) : -1;
}

// STR sector output mode of the pack kernel (encode_sector_str, mdec.c:757-836, driven as
// encode_file_str / encode_file_strspu do, filefmt.c:391-520, 546-630). sector_size == 0: off.
// A launch covers frames [frame_base, frame_base + n) of a batch that holds one or more
// independent files of frames_per_file frames each (0: the whole batch is one file). Frame k of
// a file has frame_index K = frame_index0 + k (1-based, mdec.c:769), the byte budget
// 2016 * (floor(K*num/den) - floor((K-1)*num/den)) (mdec.c:772-774 in closed form) and its
// sectors are the file's video sectors v = floor((K-1)*num/den) + j. Video sector v sits in slot
// v of the file's output region, or — place_at_lba — in the slot of its LBA in the muxed file
// (one XA audio sector per `interleave` sectors, filefmt.c:456-461); slot0 is the slot that maps
// to byte 0 of the region.
struct BsStrLayout {
	int sector_size;        // bytes between consecutive slots
	int header_offset;      // offset of the 32-byte STR header inside a sector (mdec.c:824-829)
	int frame_index0;       // frame_index of a file's first frame in this batch
	int sectors_num;        // frame_block_base_overflow
	int sectors_den;        // frame_block_overflow_den
	int video_id, width, height;
	int frame_base;         // batch index of the launch's first frame
	int frames_per_file;    // 0: one file
	long long file_stride;  // bytes between the output regions of consecutive files
	int interleave;         // 1: video only; N: 1 audio + N-1 video sectors per block
	int audio_first;        // the audio sector opens the block (default) / closes it (FLAG_STR_TRAILING_AUDIO)
	int place_at_lba;
	long long slot0;
	int framing;            // also write what init_sector_buffer_video + FORM1 checksums write (filefmt.c:73-92, 474)
	int xa_file, xa_channel;
	int format;             // FORMAT_STR / FORMAT_STRCD / FORMAT_STRV
};

// LBA of video sector v of a file (filefmt.c:456-461 with a constant video_sectors_per_block)
__host__ __device__ inline long long bs_str_lba(const BsStrLayout &l, long long v) {
	if (l.interleave <= 1) return v;
	return v + v / (l.interleave - 1) + (l.audio_first ? 1 : 0);
}
__host__ __device__ inline long long bs_str_slot(const BsStrLayout &l, long long v) {
	return l.place_at_lba ? bs_str_lba(l, v) : v;
}

// EDC tables (edc.cuh) of the current device:
const uint32_t *edc_tables_device();   // device pointer to the tables of the current device (after bs_upload_tables)

// uploads the constant tables to the current device (once per device; thread-safe)
cudaError_t bs_upload_tables();
size_t bs_pack_smem_bytes(bool v3, bool smem_stream, const BsGeometry &geo, int max_size_bound, int threads);

cudaError_t bs_launch_dct(int fdct_variant, const uint8_t *d_frames, size_t frame_bytes, int n, int width, int height,
                          const BsGeometry &geo, uint4 *d_coefs, cudaStream_t stream);

// d_gstream == nullptr: bitstream image in shared memory (bs_pack_smem_bytes(.., true, ..) <= BS_SMEM_BUDGET)
cudaError_t bs_launch_pack(int codec, int threads, int min_ctas, int n, const uint4 *d_coefs, const BsGeometry &geo,
                           const int *d_max_sizes, int max_size_bound, uint8_t *d_out, size_t out_stride,
                           psxb200_bs_result_t *d_results, uint32_t *d_gstream, size_t gstream_stride,
                           const BsStrLayout &str, cudaStream_t stream);

// The same for calls with a few frames only: a cluster of BS_PACK_CLUSTER CTAs per frame (shared-memory image,
// no second kernel: a cluster finishes its frame itself). `threads` per CTA, at most BS_PACK_MAX_THREADS.
#ifndef BS_PACK_CLUSTER_CTAS
#define BS_PACK_CLUSTER_CTAS 4
#endif
constexpr int BS_PACK_CLUSTER = BS_PACK_CLUSTER_CTAS;
cudaError_t bs_launch_pack_cluster(int codec, int threads, int n, const uint4 *d_coefs, const BsGeometry &geo,
                                   const int *d_max_sizes, int max_size_bound, uint8_t *d_out, size_t out_stride,
                                   psxb200_bs_result_t *d_results, const BsStrLayout &str, cudaStream_t stream);

// STR mode, framing: sync/header/subheader of the video sectors and their FORM1 EDC
// (filefmt.c:73-92, cdrom.c:55-74, 92-100), after bs_launch_pack wrote headers and payloads.
cudaError_t bs_launch_str_framing(int n, int max_chunks, uint8_t *d_out, const BsStrLayout &str, cudaStream_t stream);

}  // namespace psxb200
