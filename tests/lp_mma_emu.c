/* Host emulation of the short low-pass of the fused line kernels (k_line's chroma low-pass, k_sec_raster's SECAM
 * baseband low-pass; hacktv_b200/csrc/htv_mma_fir.h: mf_lp_a_word, mf_lp_b_offset, mf_lane_x, mf_lane_ci): the same
 * index helpers the kernels use drive a lane-by-lane model of mma.sync.m16n8k32 written from the PTX ISA fragment
 * tables; the result - and which lane ends up with which sample - is compared with the direct sum
 * out[x] = sum_y u[x - ntaps/2 + y] tap[y] over a zero-padded line (ref fir.c:357-375). No GPU involved.
 *
 * usage: lp_mma_emu W ntaps seed extreme(0|1)   -> prints "OK <checked>" or the first mismatch */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "htv_mma_fir.h"

static uint32_t rng_state;
static uint32_t rng(void) { rng_state = rng_state * 1664525u + 1013904223u; return(rng_state >> 8); }

/* D += A x B for one warp, registers laid out as the PTX ISA specifies for m16n8k32 (.s8/.u8) */
static void mma_emu(int32_t d[32][4], uint32_t a[32][4], uint32_t b[32][2], int a_signed, int b_signed)
{
	int A[16][32], B[32][8];
	for(int lane = 0; lane < 32; lane++)
	{
		const int g = lane >> 2, t = lane & 3;
		for(int reg = 0; reg < 4; reg++) for(int e = 0; e < 4; e++)
		{
			const int row = (reg & 1) ? g + 8 : g;
			const int col = t * 4 + e + (reg >= 2 ? 16 : 0);
			const uint32_t by = (a[lane][reg] >> (8 * e)) & 0xFF;
			A[row][col] = a_signed ? (int) (int8_t) by : (int) by;
		}
		for(int reg = 0; reg < 2; reg++) for(int e = 0; e < 4; e++)
		{
			const int k = t * 4 + e + reg * 16;
			const uint32_t by = (b[lane][reg] >> (8 * e)) & 0xFF;
			B[k][g] = b_signed ? (int) (int8_t) by : (int) by;
		}
	}
	for(int lane = 0; lane < 32; lane++)
	{
		const int g = lane >> 2, t = lane & 3;
		for(int ci = 0; ci < 4; ci++)
		{
			const int row = g + ((ci & 2) ? 8 : 0), col = 2 * t + (ci & 1);
			int64_t acc = 0;
			for(int k = 0; k < 32; k++) acc += (int64_t) A[row][k] * B[k][col];
			d[lane][ci] = (int32_t) ((uint32_t) d[lane][ci] + (uint32_t) acc);
		}
	}
}

int main(int argc, char **argv)
{
	const int W = argc > 1 ? atoi(argv[1]) : 1024;
	const int ntaps = argc > 2 ? atoi(argv[2]) : 15;
	rng_state = argc > 3 ? (uint32_t) atoi(argv[3]) : 1;
	const int extreme = argc > 4 ? atoi(argv[4]) : 0;
	if(ntaps < 3 || ntaps > 17 || !(ntaps & 1)) { printf("3 <= ntaps <= 17, odd\n"); return(2); }
	const int T = mf_tiles(W), PB = MF_TILE * T + 32, h = ntaps / 2;

	int16_t *u = calloc((size_t) W, sizeof(int16_t));
	for(int x = 0; x < W; x++)
	{
		const uint32_t r = rng();
		u[x] = (int16_t) (r & 0xFFFF);
		if((r >> 16 & 15) == 0) u[x] = (r & 1) ? 32767 : -32768;
	}
	int32_t taps[17];
	for(int y = 0; y < ntaps; y++)
	{
		const int span = extreme ? 65536 : 6000;
		taps[y] = (int) (rng() % span) - span / 2;
		if(extreme && y % 5 == 0) taps[y] = (y & 1) ? 32767 : -32768;
	}
	/* byte planes as the kernels write them: byte i = sample i - MF_LP_LEAD, zero outside 0 .. W-1 */
	uint8_t *ph = calloc(PB, 1), *pl = calloc(PB, 1);
	for(int x = 0; x < W; x++) { ph[MF_LP_LEAD + x] = (uint8_t) ((u[x] >> 8) & 0xFF); pl[MF_LP_LEAD + x] = (uint8_t) (u[x] & 0xFF); }

	int32_t *out = calloc((size_t) MF_TILE * T, sizeof(int32_t));
	uint8_t *seen = calloc((size_t) MF_TILE * T, 1);
	for(int nt = 0; nt < T; nt++)
	{
		int32_t acc[3][32][4];
		uint32_t ah[32][4], al[32][4], bh[32][2], bl[32][2];
		memset(acc, 0, sizeof(acc));
		for(int lane = 0; lane < 32; lane++)
		{
			const int off = mf_lp_b_offset(nt, lane);
			if(off < 0 || off + 8 > PB) { printf("stream load out of the plane\n"); return(1); }
			memcpy(bh[lane], ph + off, 8);
			memcpy(bl[lane], pl + off, 8);
			for(int reg = 0; reg < 4; reg++) { ah[lane][reg] = mf_lp_a_word(taps, ntaps, lane, reg, 0); al[lane][reg] = mf_lp_a_word(taps, ntaps, lane, reg, 1); }
		}
		mma_emu(acc[0], ah, bh, 1, 1);
		mma_emu(acc[1], ah, bl, 1, 0);
		mma_emu(acc[1], al, bh, 0, 1);
		mma_emu(acc[2], al, bl, 0, 0);
		for(int lane = 0; lane < 32; lane++) for(int j = 0; j < 4; j++)
		{
			const int ci = mf_lane_ci(j), x = mf_lane_x(nt, lane, j);
			if(x < 0 || x >= MF_TILE * T) { printf("sample index %d out of the tiles\n", x); return(1); }
			if(seen[x]++) { printf("sample %d owned twice\n", x); return(1); }
			out[x] = mf_combine(acc[0][lane][ci], acc[1][lane][ci], acc[2][lane][ci]);
		}
	}
	long checked = 0;
	for(int x = 0; x < W; x++)
	{
		uint32_t s = 0;
		for(int y = 0; y < ntaps; y++)
		{
			const int i = x - h + y;
			if(i >= 0 && i < W) s += (uint32_t) ((int32_t) u[i] * taps[y]);
		}
		if(!seen[x]) { printf("sample %d never written\n", x); return(1); }
		if(out[x] != (int32_t) s) { printf("MISMATCH x=%d got %d want %d\n", x, out[x], (int32_t) s); return(1); }
		checked++;
	}
	printf("OK %ld\n", checked);
	return(0);
}
