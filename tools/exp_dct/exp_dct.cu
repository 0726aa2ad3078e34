// Measured A/B of the north-star's kernel techniques for the FDCT stage:
//
//   A  scalar, __ldg gather, lanes follow the bitstream's macroblock order (the product's mapping)
//   B  scalar, __ldg gather, lanes run along a row of blocks (coalesced 256-byte rows)
//   C  scalar, rows of macroblocks staged in shared memory by TMA (cp.async.bulk.tensor.2d + mbarrier)
//   D  tensor cores: both passes of ff_jpeg_fdct_islow_8 as exact integer MMAs
//      (mma.sync.m16n8k32.s8: the 13-bit constants and the 16-bit intermediates are split into
//      signed byte digits, SURVEY.md Appendix A.1 proves the passes are integer matrix products)
//
// All four compute the same thing — NV21 frame -> level shift -> 8x8 blocks in bitstream order ->
// ff_jpeg_fdct_islow_8 (reference psxavenc/mdec.c:605-643) -> int16 coefficients, 128 bytes per
// block — are checked against each other bit for bit, and timed with CUDA events. This is an
// experiment harness, not part of the product library.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I psxavenc_b200/csrc -o exp_dct tools/exp_dct/exp_dct.cu
//   ./exp_dct [frames] [reps] [variant letters]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fdct.cuh"

#define CHECK(x)                                                                                   \
	do {                                                                                           \
		cudaError_t e_ = (x);                                                                      \
		if (e_ != cudaSuccess) {                                                                   \
			fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
			exit(1);                                                                               \
		}                                                                                          \
	} while (0)

constexpr int W = 320, H = 240, MBW = W / 16, MBH = H / 16, NMB = MBW * MBH, NBLK = 6 * NMB;
constexpr int FRAME_BYTES = W * H * 3 / 2;

// bitstream index of a block: macroblocks column-major (mdec.c:689-704), Cr Cb Y1 Y2 Y3 Y4
__host__ __device__ inline int block_index(int mx, int my, int k) { return 6 * (mx * MBH + my) + k; }

__device__ __forceinline__ void store_block(int16_t *dst, const int (&v)[64]) {
#pragma unroll
	for (int j = 0; j < 8; j++) {
		uint4 r;
		r.x = (uint32_t)(uint16_t)v[8 * j + 0] | ((uint32_t)(uint16_t)v[8 * j + 1] << 16);
		r.y = (uint32_t)(uint16_t)v[8 * j + 2] | ((uint32_t)(uint16_t)v[8 * j + 3] << 16);
		r.z = (uint32_t)(uint16_t)v[8 * j + 4] | ((uint32_t)(uint16_t)v[8 * j + 5] << 16);
		r.w = (uint32_t)(uint16_t)v[8 * j + 6] | ((uint32_t)(uint16_t)v[8 * j + 7] << 16);
		reinterpret_cast<uint4 *>(dst)[j] = r;
	}
}

__device__ __forceinline__ void unpack_luma_row(uint2 r, int *v) {
#pragma unroll
	for (int x = 0; x < 4; x++) {
		v[x] = (int)((r.x >> (8 * x)) & 0xFF) - 128;
		v[4 + x] = (int)((r.y >> (8 * x)) & 0xFF) - 128;
	}
}

__device__ __forceinline__ void unpack_chroma_row(uint4 r, int comp, int *v) {
	const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
	for (int x = 0; x < 8; x++) v[x] = (int)((w[x >> 1] >> (16 * (x & 1) + 8 * comp)) & 0xFF) - 128;
}

// ---- A: the product's mapping: consecutive lanes = consecutive blocks of the type-major plane ----
__global__ void __launch_bounds__(96, 9) dct_scalar_mbmajor(const uint8_t *__restrict__ frames, int16_t *__restrict__ out) {
	const int f = blockIdx.y, b = blockIdx.x * 96 + threadIdx.x;   // 0..599 chroma, 600..1799 luma
	if (b >= NBLK) return;
	const bool chroma = b < 2 * NMB;
	const int mb = chroma ? b >> 1 : (b - 2 * NMB) >> 2, k = chroma ? b & 1 : 2 + ((b - 2 * NMB) & 3);
	const int mx = mb / MBH, my = mb - mx * MBH;
	const uint8_t *fr = frames + (size_t)f * FRAME_BYTES;
	int v[64];
	if (chroma) {
		const uint8_t *p = fr + W * H + W * (my * 8) + mx * 16;
#pragma unroll
		for (int y = 0; y < 8; y++) unpack_chroma_row(__ldg(reinterpret_cast<const uint4 *>(p + y * W)), k, v + 8 * y);
	} else {
		const uint8_t *p = fr + W * (my * 16 + ((k - 2) >> 1) * 8) + mx * 16 + ((k - 2) & 1) * 8;
#pragma unroll
		for (int y = 0; y < 8; y++) unpack_luma_row(__ldg(reinterpret_cast<const uint2 *>(p + y * W)), v + 8 * y);
	}
	fdct8x8<FDCT_ISLOW>(v);
	store_block(out + ((size_t)f * NBLK + block_index(mx, my, k)) * 64, v);
}

// ---- B / C: one CTA per row of macroblocks; lanes run along the row ----------------------------
// 120 blocks per macroblock row: 80 luma (2 block rows x 40) then 40 chroma (20 MB x Cr, Cb)
template <bool TMA>
__global__ void __launch_bounds__(128, 7) dct_scalar_rowmajor(const uint8_t *__restrict__ frames, int16_t *__restrict__ out,
                                                              const __grid_constant__ CUtensorMap tmap) {
	__shared__ __align__(128) uint8_t tile[24 * W];    // 16 luma rows + 8 chroma rows of this macroblock row
	__shared__ __align__(8) uint64_t bar;
	const int f = blockIdx.y, my = blockIdx.x, t = threadIdx.x;
	const uint8_t *fr = frames + (size_t)f * FRAME_BYTES;
	if (TMA) {
		const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), tile_a = (uint32_t)__cvta_generic_to_shared(tile);
		if (t == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(24 * W) : "memory");
			// the frames form one 2-D tensor of 32-bit words: 80 wide, 360 rows per frame (240 luma + 120 chroma)
			// the map's box is 80 words x 8 rows: two boxes of luma rows, one of chroma rows
			const int rows[3] = {f * 360 + my * 16, f * 360 + my * 16 + 8, f * 360 + 240 + my * 8};
#pragma unroll
			for (int i = 0; i < 3; i++)
				asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
				                 tile_a + i * 8 * W),
				             "l"(&tmap), "r"(0), "r"(rows[i]), "r"(bar_a)
				             : "memory");
		}
		__syncthreads();
		uint32_t done = 0;
		while (!done)
			asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_a) : "memory");
	}
	if (t >= 120) return;
	int v[64], mx, k;
	if (t < 80) {
		const int by = t / 40, bx = t - 40 * by;
		mx = bx >> 1;
		k = 2 + 2 * by + (bx & 1);
#pragma unroll
		for (int y = 0; y < 8; y++) {
			const uint2 r = TMA ? *reinterpret_cast<const uint2 *>(tile + (8 * by + y) * W + 8 * bx)
			                    : __ldg(reinterpret_cast<const uint2 *>(fr + (size_t)(16 * my + 8 * by + y) * W + 8 * bx));
			unpack_luma_row(r, v + 8 * y);
		}
	} else {
		mx = (t - 80) >> 1;
		k = (t - 80) & 1;
#pragma unroll
		for (int y = 0; y < 8; y++) {
			const uint4 r = TMA ? *reinterpret_cast<const uint4 *>(tile + (16 + y) * W + 16 * mx)
			                    : __ldg(reinterpret_cast<const uint4 *>(fr + (size_t)W * H + (size_t)(8 * my + y) * W + 16 * mx));
			unpack_chroma_row(r, k, v + 8 * y);
		}
	}
	fdct8x8<FDCT_ISLOW>(v);
	store_block(out + ((size_t)f * NBLK + block_index(mx, my, k)) * 64, v);
}

// ---- D: tensor cores -----------------------------------------------------------------------------
// C[u][k]: the islow pass as an integer matrix (out_u = (sum_k C[u][k] in_k + round) >> shift, shift 9
// for the row pass and 17 for the column pass; rows 0 and 4 hold +-8192, which reproduces their
// "<< 4" and "(x + 8) >> 4"). Digits: C = 256 * hi + lo with lo in [-128, 127].
__constant__ int8_t c_dig[2][8][8];

__device__ __forceinline__ void mma_s8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
	asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
	             : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
	             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A strip = 8 rows x 32 bytes = 4 blocks. Luma: 4 horizontally adjacent blocks; chroma: Cr, Cb of
// two adjacent macroblocks (their samples interleaved in the 32 bytes). One warp per strip,
// several strips per warp.
template <bool CHROMA>
__global__ void __launch_bounds__(128) dct_mma(const uint8_t *__restrict__ frames, int16_t *__restrict__ out, int n_frames) {
	const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
	// ---- constant A fragments ----
	// element (reg r, byte i) of an m16n8k32 A fragment: row g + 8 (r & 1), k = 4 q + i + 16 (r >> 1)
	uint32_t a1[2][2][4];   // pass 1: [block pair j][digit]
	uint32_t a2[3][4];      // pass 2: classes hh, mid, ll
#pragma unroll
	for (int r = 0; r < 4; r++) {
#pragma unroll
		for (int j = 0; j < 2; j++)
#pragma unroll
			for (int d = 0; d < 2; d++) a1[j][d][r] = 0;
#pragma unroll
		for (int c = 0; c < 3; c++) a2[c][r] = 0;
#pragma unroll
		for (int i = 0; i < 4; i++) {
			const int row = g + 8 * (r & 1), ks = 4 * q + i + 16 * (r >> 1);
			const int bp = row >> 3, u = row & 7;
			// pass 1: which input sample does k-slot ks hold for output block (2 j + bp)?
#pragma unroll
			for (int j = 0; j < 2; j++) {
				int k = -1;
				if (!CHROMA) {
					if ((ks >> 3) == 2 * j + bp) k = ks & 7;
				} else {
					if ((ks >> 4) == j && (ks & 1) == bp) k = (ks & 15) >> 1;
				}
#pragma unroll
				for (int d = 0; d < 2; d++)
					if (k >= 0) a1[j][d][r] |= (uint32_t)(uint8_t)c_dig[d][u][k] << (8 * i);
			}
			// pass 2: k-slot ks holds digit (i >> 1) of T[n = 2 (ks & 15) / 4 + (i & 1)] of block (ks >> 4) of the pair
			const int blk = ks >> 4, n = 2 * ((ks & 15) >> 2) + (i & 1), dig = i >> 1;
			if (blk == bp) {
				if (dig == 0) {
					a2[0][r] |= (uint32_t)(uint8_t)c_dig[0][u][n] << (8 * i);   // hh
					a2[1][r] |= (uint32_t)(uint8_t)c_dig[1][u][n] << (8 * i);   // mid: lo(C) x hi(T)
				} else {
					a2[1][r] |= (uint32_t)(uint8_t)c_dig[0][u][n] << (8 * i);   // mid: hi(C) x lo(T)
					a2[2][r] |= (uint32_t)(uint8_t)c_dig[1][u][n] << (8 * i);   // ll
				}
			}
		}
	}

	constexpr int BANDS = CHROMA ? H / 16 : H / 8, STRIPS = W / 32;   // 8-row bands, 32-byte strips per band
	const long warps = (long)gridDim.x * (blockDim.x >> 5), warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const long total = (long)n_frames * BANDS * STRIPS;
	for (long s = warp; s < total; s += warps) {
		const int f = (int)(s / (BANDS * STRIPS)), rem = (int)(s - (long)f * BANDS * STRIPS);
		const int band = rem / STRIPS, strip = rem - band * STRIPS;
		const uint8_t *p = frames + (size_t)f * FRAME_BYTES + (CHROMA ? W * H : 0) + (size_t)(8 * band + g) * W + 32 * strip + 4 * q;
		// B fragment of pass 1: this lane's row g of the strip, bytes 4q..4q+3 and 16+4q..; "- 128" = flip the top bit
		const uint32_t b0 = __ldg(reinterpret_cast<const uint32_t *>(p)) ^ 0x80808080u;
		const uint32_t b1 = __ldg(reinterpret_cast<const uint32_t *>(p + 16)) ^ 0x80808080u;
		uint32_t tb[2][2];   // B fragments of pass 2: [block pair][register]
#pragma unroll
		for (int j = 0; j < 2; j++) {
			int dh[4] = {0, 0, 0, 0}, dl[4] = {0, 0, 0, 0};
			mma_s8(dh, a1[j][0], b0, b1);
			mma_s8(dl, a1[j][1], b0, b1);
			// T[n][u = g] for n = 2q, 2q+1 of blocks 2j (c0, c1) and 2j+1 (c2, c3)
			uint32_t hi[4], lo[4];
#pragma unroll
			for (int c = 0; c < 4; c++) {
				const int t = (dh[c] * 256 + dl[c] + 256) >> 9;
				const int th = (t + 128) >> 8;
				hi[c] = (uint32_t)th & 0xFFu;
				lo[c] = (uint32_t)(t - 256 * th) & 0xFFu;
			}
			tb[j][0] = hi[0] | (hi[1] << 8) | (lo[0] << 16) | (lo[1] << 24);
			tb[j][1] = hi[2] | (hi[3] << 8) | (lo[2] << 16) | (lo[3] << 24);
		}
#pragma unroll
		for (int j = 0; j < 2; j++) {
			int hh[4] = {0, 0, 0, 0}, mid[4] = {0, 0, 0, 0}, ll[4] = {0, 0, 0, 0};
			mma_s8(hh, a2[0], tb[j][0], tb[j][1]);
			mma_s8(mid, a2[1], tb[j][0], tb[j][1]);
			mma_s8(ll, a2[2], tb[j][0], tb[j][1]);
			// Y[v = g][u = 2q, 2q+1] of blocks 2j (c0, c1) and 2j+1 (c2, c3)
#pragma unroll
			for (int bp = 0; bp < 2; bp++) {
				const int y0 = (int)(((uint32_t)hh[2 * bp] << 16) + ((uint32_t)mid[2 * bp] << 8) + (uint32_t)ll[2 * bp] + 65536u) >> 17;
				const int y1 = (int)(((uint32_t)hh[2 * bp + 1] << 16) + ((uint32_t)mid[2 * bp + 1] << 8) + (uint32_t)ll[2 * bp + 1] + 65536u) >> 17;
				const int b = 2 * j + bp;
				int mx, my, k;
				if (!CHROMA) {
					const int bx = 4 * strip + b;
					mx = bx >> 1;
					my = band >> 1;
					k = 2 + 2 * (band & 1) + (bx & 1);
				} else {
					mx = 2 * strip + (b >> 1);
					my = band;
					k = b & 1;
				}
				int16_t *dst = out + ((size_t)f * NBLK + block_index(mx, my, k)) * 64;
				reinterpret_cast<uint32_t *>(dst)[4 * g + q] = (uint32_t)(uint16_t)y0 | ((uint32_t)(uint16_t)y1 << 16);
			}
		}
	}
}

// ---- host ----------------------------------------------------------------------------------------

static void islow_matrix(int C[8][8]) {
	// the pass is linear before its rounding shift: feed unit vectors through the same arithmetic
	for (int k = 0; k < 8; k++) {
		long d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		d[k] = 1;
		long e0 = d[0] + d[7], o0 = d[0] - d[7], e1 = d[1] + d[6], o1 = d[1] - d[6], e2 = d[2] + d[5], o2 = d[2] - d[5], e3 = d[3] + d[4], o3 = d[3] - d[4];
		long ee0 = e0 + e3, eo0 = e0 - e3, ee1 = e1 + e2, eo1 = e1 - e2;
		C[0][k] = (int)((ee0 + ee1) * 8192);
		C[4][k] = (int)((ee0 - ee1) * 8192);
		long z = (eo1 + eo0) * 4433;
		C[2][k] = (int)(z + eo0 * 6270);
		C[6][k] = (int)(z - eo1 * 15137);
		long z1 = o3 + o0, z2 = o2 + o1, z3 = o3 + o1, z4 = o2 + o0, z5 = (z3 + z4) * 9633;
		z3 = z3 * -16069 + z5;
		z4 = z4 * -3196 + z5;
		z1 *= -7373;
		z2 *= -20995;
		C[7][k] = (int)(o3 * 2446 + z1 + z3);
		C[5][k] = (int)(o2 * 16819 + z2 + z4);
		C[3][k] = (int)(o1 * 25172 + z2 + z3);
		C[1][k] = (int)(o0 * 12299 + z1 + z4);
	}
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
	const int n = argc > 1 ? atoi(argv[1]) : 2048, reps = argc > 2 ? atoi(argv[2]) : 20;
	const char *which = argc > 3 ? argv[3] : "ABCD";
	std::vector<uint8_t> h_frames((size_t)n * FRAME_BYTES);
	uint32_t lcg = 12345;
	for (int f = 0; f < n; f++) {
		uint8_t *fr = h_frames.data() + (size_t)f * FRAME_BYTES;
		for (int y = 0; y < H * 3 / 2; y++)
			for (int x = 0; x < W; x++) {
				lcg = lcg * 1664525u + 1013904223u;
				int v = 64 + ((x + 2 * f) & 127) + ((y * 3 + f) & 63) + (int)(lcg >> 28);
				if ((f & 7) == 7) v = (int)(lcg >> 24);   // every eighth frame: white noise (full range)
				fr[y * W + x] = (uint8_t)(v > 255 ? 255 : v);
			}
	}
	uint8_t *d_frames;
	int16_t *d_out[4];
	const size_t out_elems = (size_t)n * NBLK * 64;
	CHECK(cudaMalloc(&d_frames, h_frames.size()));
	CHECK(cudaMemcpy(d_frames, h_frames.data(), h_frames.size(), cudaMemcpyHostToDevice));
	for (auto &p : d_out) {
		CHECK(cudaMalloc(&p, out_elems * 2));
		CHECK(cudaMemset(p, 0xEE, out_elems * 2));
	}

	int C[8][8];
	islow_matrix(C);
	int8_t dig[2][8][8];
	for (int u = 0; u < 8; u++)
		for (int k = 0; k < 8; k++) {
			const int hi = (C[u][k] + 128) >> 8, lo = C[u][k] - 256 * hi;
			if (hi < -128 || hi > 127 || lo < -128 || lo > 127) {
				fprintf(stderr, "digit overflow C[%d][%d] = %d\n", u, k, C[u][k]);
				return 1;
			}
			dig[0][u][k] = (int8_t)hi;
			dig[1][u][k] = (int8_t)lo;
		}
	CHECK(cudaMemcpyToSymbol(c_dig, dig, sizeof(dig)));

	// TMA descriptor: all frames as one 2-D tensor of 32-bit words, 80 x (360 n), box 80 x 8 rows
	CUtensorMap tmap_luma;
	{
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
		EncodeTiled encode = reinterpret_cast<EncodeTiled>(fn);
		const cuuint64_t dims[2] = {W / 4, (cuuint64_t)360 * n}, strides[1] = {W};
		const cuuint32_t elem[2] = {1, 1};
		const cuuint32_t box[2] = {W / 4, 8};
		if (!encode || encode(&tmap_luma, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d_frames, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
		                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
			fprintf(stderr, "cuTensorMapEncodeTiled failed\n");
			return 1;
		}
	}

	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	auto launch = [&](char v, int16_t *out) {
		switch (v) {
		case 'A': dct_scalar_mbmajor<<<dim3((NBLK + 95) / 96, n), 96>>>(d_frames, out); break;
		case 'B': dct_scalar_rowmajor<false><<<dim3(MBH, n), 128>>>(d_frames, out, tmap_luma); break;
		case 'C': dct_scalar_rowmajor<true><<<dim3(MBH, n), 128>>>(d_frames, out, tmap_luma); break;
		case 'D':
			dct_mma<false><<<sms * 16, 128>>>(d_frames, out, n);
			dct_mma<true><<<sms * 8, 128>>>(d_frames, out, n);
			break;
		}
	};
	const char *names[4] = {"A scalar, ldg, macroblock-order lanes (product mapping)", "B scalar, ldg, lanes along a block row",
	                        "C scalar, TMA-staged macroblock rows", "D tensor cores (mma.sync s8 digits), both passes"};
	cudaEvent_t e0, e1;
	CHECK(cudaEventCreate(&e0));
	CHECK(cudaEventCreate(&e1));
	std::vector<int16_t> ref, got(out_elems);
	printf("%d frames of %dx%d (%d blocks), %d timed launches per variant\n", n, W, H, n * NBLK, reps);
	for (int i = 0; i < 4; i++) {
		const char v = "ABCD"[i];
		if (!strchr(which, v)) continue;
		for (int k = 0; k < 3; k++) launch(v, d_out[i]);
		CHECK(cudaDeviceSynchronize());
		CHECK(cudaEventRecord(e0));
		for (int k = 0; k < reps; k++) launch(v, d_out[i]);
		CHECK(cudaEventRecord(e1));
		CHECK(cudaDeviceSynchronize());
		float ms = 0;
		CHECK(cudaEventElapsedTime(&ms, e0, e1));
		CHECK(cudaMemcpy(got.data(), d_out[i], out_elems * 2, cudaMemcpyDeviceToHost));
		const char *verdict = "reference";
		if (ref.empty()) ref = got;
		else verdict = memcmp(ref.data(), got.data(), out_elems * 2) == 0 ? "bit-identical" : "DIFFERENT";
		printf("%-62s %8.4f ms per launch  %7.2f ms per 4096 frames  %6.1f Gblocks/s  [%s]\n", names[i], ms / reps, ms / reps * 4096.0 / n,
		       (double)n * NBLK / (ms / reps) / 1e6, verdict);
	}
	return 0;
}
