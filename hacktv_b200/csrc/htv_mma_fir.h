/* The 51-tap video filter (ref _vid_filter_process video.c:3235-3248, fir_int16_scomplex_process
 * fir.c:564-615) as an exact int8 tensor-core contraction. Index arithmetic only, shared by the
 * CUDA kernel (k_mod_mma, htv_kernels.cu) and the host-side emulation in tests/ so the fragment
 * mapping can be checked without a GPU.
 *
 * out[x] = sum_{y=0..50} comp[x - 25 + y] * tap[y]                       (centred FIR)
 *
 * A line is cut into rows of MF_T = 32 outputs: x = 32 r + c. With k' = c + y + MF_SHIFT
 *   A[r][k'] = comp[32 r + k' - MF_LEAD]      (overlapping windows of one contiguous stream)
 *   B[k'][c] = tap[k' - MF_SHIFT - c]         (banded Toeplitz, zero outside 0..50)
 * so out = A x B with K = 96 (3 k-steps of 32). MF_LEAD = 25 + MF_SHIFT = 32 puts byte 0 of the
 * window 32 samples before the line: every row starts on a 32-byte boundary.
 *
 * int16 x int16 products are made exact by the byte split v = 256 hi + lo (hi signed, lo
 * unsigned): x h = 65536 xh hh + 256 (xh hl + xl hh) + xl hl - three int32 accumulators fed by
 * four mma.sync.m16n8k32 (s8.s8, s8.u8, u8.s8, u8.u8); each partial sum stays below 2^23.
 *
 * Fragment use (PTX ISA, m16n8k32 integer): lane = 4 g + t. A register a0/a2 carries k-slots
 * 4t..4t+3 / 16+4t..16+4t+3 of row g, a1/a3 the same of row g + 8; B register b0/b1 carries
 * those slots of column g. The k index is a summation index, so slots may name any k' as long
 * as A and B agree: here lane t's (a0, a2) = the 8 consecutive bytes k' = 32 s + 8 t .. + 7
 * (one 64-bit shared-memory load; a warp reads 256 contiguous bytes, conflict-free) and
 * (b0, b1) the matching taps. */
#ifndef HTV_MMA_FIR_H
#define HTV_MMA_FIR_H

#include <stdint.h>

#if defined(__CUDACC__)
#define MF_HD __host__ __device__ __forceinline__
#else
#define MF_HD static inline
#endif

#define MF_T       32      /* outputs per A row */
#define MF_NTAPS   51
#define MF_SHIFT   7
#define MF_LEAD    32      /* window byte 0 = composite sample -32 of the line */
#define MF_KSTEPS  3       /* K = 96 */
#define MF_ROWW    40      /* exchange buffer: words per row of 32 outputs (bank-conflict-free both ways) */

/* rows of 32 outputs in a line, m16 tiles, bytes of one plane buffer in shared memory */
MF_HD int mf_rows(int W) { return(W / MF_T); }
MF_HD int mf_mtiles(int W) { return((W / MF_T + 15) / 16); }
MF_HD int mf_plane_bytes(int W) { return(mf_mtiles(W) * 16 * MF_T + 64); }    /* >= W + 64, multiple of 16 */
MF_HD int mf_window_bytes(int W) { return(W + 2 * MF_LEAD); }                  /* what the TMA brings per plane */

/* byte offset, inside a plane window, of the 8 bytes lane (g, t) loads for m-tile mt, k-step s;
 * half = 0: row g, half = 1: row g + 8 */
MF_HD int mf_a_offset(int mt, int s, int lane, int half)
{
	const int g = lane >> 2, t = lane & 3;
	return(MF_T * (mt * 16 + g + 8 * half) + 32 * s + 8 * t);
}

/* one 32-bit B register. taps[51] in application order (tap[y] multiplies comp[x - 25 + y]).
 * j = n-tile (columns 8 j .. 8 j + 7), s = k-step, w: bit 0 = b1 (second four k'), bit 1 = low byte plane */
MF_HD uint32_t mf_b_word(const int32_t *taps, int j, int s, int lane, int w)
{
	const int g = lane >> 2, t = lane & 3;
	const int c = 8 * j + g;
	const int kp0 = 32 * s + 8 * t + (w & 1) * 4;
	uint32_t r = 0;
	for(int e = 0; e < 4; e++)
	{
		const int y = kp0 + e - MF_SHIFT - c;
		const int h = (y >= 0 && y < MF_NTAPS) ? taps[y] : 0;
		const uint32_t b = (w & 2) ? ((uint32_t) h & 0xFFu) : (((uint32_t) h >> 8) & 0xFFu);
		r |= b << (8 * e);
	}
	return(r);
}

/* sample index of accumulator register ci (0..3) of lane (g, t) for m-tile mt, n-tile j */
MF_HD int mf_out_x(int mt, int j, int lane, int ci)
{
	const int g = lane >> 2, t = lane & 3;
	return(MF_T * (mt * 16 + g + ((ci & 2) ? 8 : 0)) + 8 * j + 2 * t + (ci & 1));
}

/* word index of sample x in the exchange buffer */
MF_HD int mf_fir_index(int x) { return((x >> 5) * MF_ROWW + (x & 31)); }

/* the three partial sums back to the int32 the reference accumulates (wraps like it) */
MF_HD int32_t mf_combine(int32_t hh, int32_t mid, int32_t ll)
{
	return((int32_t) (((uint32_t) hh << 16) + ((uint32_t) mid << 8) + (uint32_t) ll));
}

#endif
