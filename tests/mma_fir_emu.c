/* Host emulation of k_mod_mma's tensor-core FIR (hacktv_b200/csrc/htv_mma_fir.h): the same index
 * helpers the kernel uses drive a lane-by-lane model of mma.sync.m16n8k32 written from the PTX
 * ISA fragment tables, and the result is compared with the direct 51-tap sum the reference
 * computes (ref fir.c:564-615). Built and run by tests/test_mma_fir_host.py; no GPU involved.
 *
 * usage: mma_fir_emu W seed extreme(0|1)   -> prints "OK <checked>" or the first mismatch */
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

static int sat16(int v) { return(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

int main(int argc, char **argv)
{
	const int W = argc > 1 ? atoi(argv[1]) : 1024;
	rng_state = argc > 2 ? (uint32_t) atoi(argv[2]) : 1;
	const int extreme = argc > 3 ? atoi(argv[3]) : 0;
	if(W % MF_TILE) { printf("W must be a multiple of %d\n", MF_TILE); return(2); }

	/* a composite stream of three lines; the middle one is filtered */
	int16_t *comp = malloc(sizeof(int16_t) * 3 * W);
	for(int i = 0; i < 3 * W; i++)
	{
		const uint32_t r = rng();
		comp[i] = (int16_t) (r & 0xFFFF);
		if((r >> 16 & 15) == 0) comp[i] = (r & 1) ? 32767 : -32768;
	}
	int32_t taps[2][MF_NTAPS];
	for(int q = 0; q < 2; q++) for(int y = 0; y < MF_NTAPS; y++)
	{
		const int span = extreme ? 65536 : 2600;       /* sum |taps| < 2^16 keeps the int32 sum in range */
		taps[q][y] = (int) (rng() % span) - span / 2;
		if(extreme && y % 7 == 0) taps[q][y] = (y & 1) ? 32767 : -32768;
	}

	/* byte planes of the window, as k_raster writes them and the TMA delivers them */
	const int PB = mf_plane_bytes(W);
	uint8_t *ph = malloc(PB), *pl = malloc(PB);
	for(int i = 0; i < PB; i++) { ph[i] = (uint8_t) rng(); pl[i] = (uint8_t) rng(); }   /* garbage past the window */
	const int16_t *line = comp + W;
	for(int i = 0; i < mf_window_bytes(W); i++)
	{
		const int v = line[i - MF_LEAD];
		ph[i] = (uint8_t) ((v >> 8) & 0xFF);
		pl[i] = (uint8_t) (v & 0xFF);
	}

	uint32_t atab[MF_ATAB_WORDS];
	mf_build_atab(taps[0], taps[1], atab);
	uint32_t *fir = calloc((size_t) (W / 32) * MF_ROWW, sizeof(uint32_t));
	uint8_t *seen = calloc((size_t) W, 1);
	for(int nt = 0; nt < mf_tiles(W); nt++)
	{
		int32_t acc[2][3][32][4];
		memset(acc, 0, sizeof(acc));
		for(int s = 0; s < MF_KSTEPS; s++)
		{
			uint32_t a[4][32][4], bh[32][2], bl[32][2];
			for(int lane = 0; lane < 32; lane++)
			{
				const int off = mf_b_offset(nt, s, lane);
				if(off + 8 > PB) { printf("stream load out of the plane buffer\n"); return(1); }
				memcpy(bh[lane], ph + off, 8);
				memcpy(bl[lane], pl + off, 8);
				for(int kind = 0; kind < 4; kind++)
					memcpy(a[kind][lane], atab + ((s * 4 + kind) * 32 + lane) * 4, 16);
			}
			for(int q = 0; q < 2; q++)
			{
				mma_emu(acc[q][0], a[2 * q], bh, 1, 1);        /* taps hi x stream hi */
				mma_emu(acc[q][1], a[2 * q], bl, 1, 0);        /* taps hi x stream lo */
				mma_emu(acc[q][1], a[2 * q + 1], bh, 0, 1);    /* taps lo x stream hi */
				mma_emu(acc[q][2], a[2 * q + 1], bl, 0, 0);    /* taps lo x stream lo */
			}
		}
		for(int lane = 0; lane < 32; lane++) for(int ci = 0; ci < 4; ci++)
		{
			const int x = mf_out_x(nt, lane, ci);
			if(x < 0 || x >= W) { printf("output index %d out of the line\n", x); return(1); }
			const int vi = mf_combine(acc[0][0][lane][ci], acc[0][1][lane][ci], acc[0][2][lane][ci]);
			const int vq = mf_combine(acc[1][0][lane][ci], acc[1][1][lane][ci], acc[1][2][lane][ci]);
			if(seen[x]++) { printf("sample %d written twice\n", x); return(1); }
			fir[mf_fir_index(x)] = ((uint32_t) sat16(vi >> 15) & 0xFFFF) | ((uint32_t) sat16(vq >> 15) << 16);
		}
	}

	long checked = 0;
	for(int x = 0; x < W; x++)
	{
		uint32_t si = 0, sq = 0;
		for(int y = 0; y < MF_NTAPS; y++)
		{
			si += (uint32_t) ((int32_t) line[x - 25 + y] * taps[0][y]);
			sq += (uint32_t) ((int32_t) line[x - 25 + y] * taps[1][y]);
		}
		const int ei = sat16((int32_t) si >> 15), eq = sat16((int32_t) sq >> 15);
		if(!seen[x]) { printf("sample %d never written\n", x); return(1); }
		const uint32_t got = fir[mf_fir_index(x)];
		if((int16_t) (got & 0xFFFF) != ei || (int16_t) (got >> 16) != eq)
		{
			printf("MISMATCH x=%d got (%d, %d) want (%d, %d)\n", x, (int16_t) (got & 0xFFFF), (int16_t) (got >> 16), ei, eq);
			return(1);
		}
		checked++;
	}
	printf("OK %ld\n", checked);
	return(0);
}
