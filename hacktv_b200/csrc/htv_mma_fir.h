/* The 51-tap video filter (ref _vid_filter_process video.c:3235-3248, fir_int16_scomplex_process
 * fir.c:564-615) as an exact int8 tensor-core contraction. Index arithmetic only, shared by the
 * CUDA kernel (k_mod_mma, htv_kernels.cu), the host code that builds the tap operand, and the
 * host-side emulation in tests/ so the fragment mapping can be checked without a GPU.
 *
 * out[x] = sum_{y=0..50} comp[x - 25 + y] * tap[y]                       (centred FIR)
 *
 * A line is cut into stream rows of MF_M = 16 outputs: x = 16 r + m. With k' = m + y + MF_SHIFT
 *   A[m][k'] = tap[k' - MF_SHIFT - m]         (banded Toeplitz, zero outside 0..50, constant)
 *   B[k'][r] = comp[16 r + k' - MF_LEAD]      (overlapping windows of one contiguous stream)
 * so out = A x B with K = 96 (3 k-steps of 32; taps reach k' = 72). MF_LEAD = 25 + MF_SHIFT = 32
 * puts byte 0 of the window 32 samples before the line: every row starts on a 16-byte boundary.
 * One mma tile is 16 outputs x 8 rows = 128 consecutive samples; a warp owns one tile of a line.
 *
 * int16 x int16 products are made exact by the byte split v = 256 hi + lo (hi signed, lo
 * unsigned): h x = 65536 hh xh + 256 (hh xl + hl xh) + hl xl - three int32 accumulators fed by
 * four mma.sync.m16n8k32 (s8.s8, s8.u8, u8.s8, u8.u8); each partial sum stays below 2^23.
 *
 * Fragment use (PTX ISA, m16n8k32 integer): lane = 4 g + t. B register b0 / b1 carries k-slots
 * 4t..4t+3 / 16+4t..16+4t+3 of column g; A register a0 / a2 those slots of row g, a1 / a3 of row
 * g + 8. The k index is a summation index, so slots may name any k' as long as A and B agree:
 * here lane t's (b0, b1) = the 8 consecutive stream bytes k' = 32 s + 8 t .. + 7 of row g - one
 * 64-bit shared-memory load straight into the register pair, no shuffling - and (a0 .. a3) the
 * matching taps, prepared once on the host in fragment order (one 128-bit load per fragment). */
#ifndef HTV_MMA_FIR_H
#define HTV_MMA_FIR_H

#include <stdint.h>

#if defined(__CUDACC__)
#define MF_HD __host__ __device__ __forceinline__
#else
#define MF_HD static inline
#endif

#define MF_M       16      /* outputs per stream row */
#define MF_TILE    128     /* samples per mma tile (16 outputs x 8 rows) */
#define MF_NTAPS   51
#define MF_SHIFT   7
#define MF_LEAD    32      /* window byte 0 = composite sample -32 of the line */
#define MF_KSTEPS  3       /* K = 96 */
#define MF_ROWW    40      /* exchange buffer: words per 32 outputs (bank-conflict-free both ways) */
#define MF_ATAB_WORDS (MF_KSTEPS * 4 * 32 * 4)   /* tap operand: [k-step][I hi, I lo, Q hi, Q lo][lane] x 4 registers */

/* Two layouts of the byte planes in device memory:
 *  contiguous (128 | W): the planes are the composite stream itself, line after line; the window of a
 *    line starts MF_LEAD bytes before it (16-byte aligned because 16 | W). mf_window_bytes() arrive
 *    by TMA, the buffer is mf_plane_bytes() long and its tail is read by zero taps only.
 *  pitched (any W): one row of `pitch` bytes per line, byte i = sample i - MF_LEAD of that line;
 *    k_raster also writes a line's first / last MF_LEAD samples into the halo of the previous / next
 *    row, so a row is self-contained and 16-byte aligned whatever W is (NTSC: 858). mf_row_bytes()
 *    arrive by TMA; bytes past sample W + MF_LEAD meet zero taps only (any finite value will do). */
MF_HD int mf_tiles(int W) { return((W + MF_TILE - 1) / MF_TILE); }
MF_HD int mf_plane_bytes(int W) { return(W + 96); }
MF_HD int mf_window_bytes(int W) { return(W + 2 * MF_LEAD); }
MF_HD int mf_row_bytes(int W) { return(MF_TILE * mf_tiles(W) + 96); }
MF_HD int mf_pitch(int W) { return(MF_TILE * mf_tiles(W) + 128); }

/* byte offset, inside a plane window, of the 8 stream bytes lane (g, t) loads for tile nt, k-step s */
MF_HD int mf_b_offset(int nt, int s, int lane)
{
	const int g = lane >> 2, t = lane & 3;
	return(MF_M * (8 * nt + g) + 32 * s + 8 * t);
}

/* one 32-bit register of the tap operand. taps[51] in application order (tap[y] multiplies
 * comp[x - 25 + y]); s = k-step, reg = 0..3 (a0..a3), lo = low byte plane */
MF_HD uint32_t mf_a_word(const int32_t *taps, int s, int lane, int reg, int lo)
{
	const int g = lane >> 2, t = lane & 3;
	const int m = g + ((reg & 1) ? 8 : 0);
	const int kp0 = 32 * s + 8 * t + ((reg & 2) ? 4 : 0);
	uint32_t r = 0;
	for(int e = 0; e < 4; e++)
	{
		const int y = kp0 + e - MF_SHIFT - m;
		const int h = (y >= 0 && y < MF_NTAPS) ? taps[y] : 0;
		const uint32_t b = lo ? ((uint32_t) h & 0xFFu) : (((uint32_t) h >> 8) & 0xFFu);
		r |= b << (8 * e);
	}
	return(r);
}

/* the whole tap operand in the order k_mod_mma reads it; out[MF_ATAB_WORDS] */
MF_HD void mf_build_atab(const int32_t *itaps, const int32_t *qtaps, uint32_t *out)
{
	for(int s = 0; s < MF_KSTEPS; s++) for(int kind = 0; kind < 4; kind++) for(int lane = 0; lane < 32; lane++)
		for(int reg = 0; reg < 4; reg++)
			out[((s * 4 + kind) * 32 + lane) * 4 + reg] = mf_a_word((kind & 2) ? qtaps : itaps, s, lane, reg, kind & 1);
}

/* sample index of accumulator register ci (0..3) of lane (g, t) for tile nt */
MF_HD int mf_out_x(int nt, int lane, int ci)
{
	const int g = lane >> 2, t = lane & 3;
	return(MF_TILE * nt + MF_M * (2 * t + (ci & 1)) + g + ((ci & 2) ? 8 : 0));
}

/* word index of sample x in the exchange buffer */
MF_HD int mf_fir_index(int x) { return((x >> 5) * MF_ROWW + (x & 31)); }

/* the three partial sums back to the int32 the reference accumulates (wraps like it) */
MF_HD int32_t mf_combine(int32_t hh, int32_t mid, int32_t ll)
{
	return((int32_t) (((uint32_t) hh << 16) + ((uint32_t) mid << 8) + (uint32_t) ll));
}

/* ---- the fused line kernels' lane layout and their short low-pass (k_line's chroma low-pass, k_sec_raster's baseband
 * low-pass: up to 17 taps, one k-step of 32) ------------------------------------------------------------------------
 * A lane keeps four samples of its warp's tile in the order the accumulators of an m16n8k32 contraction arrive:
 * sample j of lane (g, t) is x = 128 nt + 32 t + g + 8 j, held by accumulator register ci = ((j & 1) << 1) | (j >> 1).
 * The low-pass reads byte planes whose byte 0 is sample -MF_LP_LEAD of the line:
 *   out[x] = sum_y u[x - h + y] tap[y], h = ntaps / 2;   A[m][k'] = tap[k' - m - (MF_LP_LEAD - h)]. */
#define MF_LP_LEAD 8

MF_HD int mf_lane_x(int nt, int lane, int j) { return(MF_TILE * nt + 32 * (lane & 3) + (lane >> 2) + 8 * j); }
MF_HD int mf_lane_ci(int j) { return(((j & 1) << 1) | (j >> 1)); }

/* byte offset, inside a low-pass plane, of the 8 stream bytes lane (g, t) loads for tile nt */
MF_HD int mf_lp_b_offset(int nt, int lane) { return(MF_M * (8 * nt + (lane >> 2)) + 8 * (lane & 3)); }

/* one 32-bit register of the low-pass tap operand (same fragment convention as mf_a_word) */
MF_HD uint32_t mf_lp_a_word(const int32_t *taps, int ntaps, int lane, int reg, int lo)
{
	const int g = lane >> 2, t = lane & 3, h = ntaps / 2;
	const int m = g + ((reg & 1) ? 8 : 0);
	const int kp0 = 8 * t + ((reg & 2) ? 4 : 0);
	uint32_t r = 0;
	for(int e = 0; e < 4; e++)
	{
		const int y = kp0 + e - (MF_LP_LEAD - h) - m;
		const int v = (y >= 0 && y < ntaps) ? taps[y] : 0;
		const uint32_t b = lo ? ((uint32_t) v & 0xFFu) : (((uint32_t) v >> 8) & 0xFFu);
		r |= b << (8 * e);
	}
	return(r);
}

#endif
