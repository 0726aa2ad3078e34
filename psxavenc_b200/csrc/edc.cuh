// CD-ROM XA error detection code (reference libpsxav/cdrom.c:30-41: reflected CRC-32, polynomial
// 0xD8018001, zero initial value, no final xor) evaluated by one warp per sector.
//
// The CRC is linear in the message, so the byte range is cut into 31 pieces of PIECE 32-bit
// words (the last one shorter), lane l runs the table-driven CRC over piece l, and the pieces are
// chained front to back: crc = advance(crc, bytes of the next piece) ^ crc(next piece).
// `advance` by a full piece is four lookups in byte-indexed tables built on the host
// (edc_build_tables); the short last piece is advanced byte by byte.
#pragma once

#include <stdint.h>

namespace psxb200 {

constexpr int EDC_PIECE_FORM1 = 17;   // 0x808 bytes = 514 words -> 30 pieces of 17 + 4 words
constexpr int EDC_PIECE_FORM2 = 19;   // 0x91C bytes = 583 words -> 30 pieces of 19 + 13 words
// device table layout: [0..255] the byte table; [256..1279] advance by 4*EDC_PIECE_FORM1 bytes;
// [1280..2303] advance by 4*EDC_PIECE_FORM2 bytes
constexpr int EDC_TABLE_WORDS = 256 + 1024 + 1024;

__host__ __device__ inline uint32_t edc_byte_table_entry(uint32_t i) {
	uint32_t c = i;
	for (int b = 0; b < 8; b++) c = (c >> 1) ^ ((c & 1) ? 0xD8018001u : 0u);
	return c;
}

// tab: [0..255] byte table, [256 + 256*k + b] = the CRC state (b << 8k) advanced over 4*PIECE zero bytes
inline void edc_build_advance_table(const uint32_t *byte_tab, int piece_words, uint32_t *out /* 1024 */) {
	for (int k = 0; k < 4; k++) {
		for (uint32_t b = 0; b < 256; b++) {
			uint32_t c = b << (8 * k);
			for (int i = 0; i < 4 * piece_words; i++) c = (c >> 8) ^ byte_tab[c & 0xFF];
			out[256 * k + b] = c;
		}
	}
}

__device__ __forceinline__ uint32_t edc_word(uint32_t crc, uint32_t v, const uint32_t *tab) {
	crc ^= v;
	crc = (crc >> 8) ^ tab[crc & 0xFF];
	crc = (crc >> 8) ^ tab[crc & 0xFF];
	crc = (crc >> 8) ^ tab[crc & 0xFF];
	crc = (crc >> 8) ^ tab[crc & 0xFF];
	return crc;
}

// All 32 lanes call this. words: the 4-byte aligned start of the range (global or shared memory),
// nwords its length in 32-bit words (<= 31 * PIECE); tab = byte table, adv = the advance table of
// this PIECE (both in shared memory). Returns the EDC in every lane.
template <int PIECE>
__device__ __forceinline__ uint32_t warp_edc(const uint32_t *words, int nwords, const uint32_t *tab, const uint32_t *adv) {
	const int lane = threadIdx.x & 31;
	const int first = lane * PIECE;
	uint32_t c = 0;
#pragma unroll 1
	for (int i = 0; i < PIECE; i++)
		if (first + i < nwords) c = edc_word(c, words[first + i], tab);
	const int full = nwords / PIECE;            // pieces of full length
	const int tail = nwords - full * PIECE;     // words of the short last piece (may be 0)
	uint32_t acc = __shfl_sync(0xFFFFFFFFu, c, 0);
	if (full == 0) return acc;                  // the whole range is lane 0's (short) piece
	for (int l = 1; l < full; l++) {
		const uint32_t next = __shfl_sync(0xFFFFFFFFu, c, l);
		acc = adv[acc & 0xFF] ^ adv[256 + ((acc >> 8) & 0xFF)] ^ adv[512 + ((acc >> 16) & 0xFF)] ^ adv[768 + (acc >> 24)] ^ next;
	}
	if (tail) {
		const uint32_t next = __shfl_sync(0xFFFFFFFFu, c, full);
		for (int i = 0; i < 4 * tail; i++) acc = (acc >> 8) ^ tab[acc & 0xFF];
		acc ^= next;
	}
	return acc;
}

}  // namespace psxb200
