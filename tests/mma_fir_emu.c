/* Host emulation of k_mod_mma's tensor-core FIR (hacktv_b200/csrc/htv_mma_fir.h): the same index
 * helpers the kernel uses drive a lane-by-lane model of mma.sync.m16n8k32 written from the PTX
 * ISA fragment tables, and the result is compared with the direct 51-tap sum the reference
 * computes (ref fir.c:564-615). Built and run by tests/test_mma_fir_host.py; no GPU involved.
 *
 * usage: mma_fir_emu W seed extreme(0|1) [pitched(0|1)]   -> prints "OK <checked>" or the first mismatch
 * pitched = 1: the planes are built row by row the way k_raster's pitched branch writes them (own
 * samples + halos into the neighbouring rows) and any W is allowed; 0: contiguous stream, 128 | W. */
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
	const int pitched = argc > 4 ? atoi(argv[4]) : 0;
	if(!pitched && W % MF_TILE) { printf("W must be a multiple of %d\n", MF_TILE); return(2); }

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
	const int PB = pitched ? mf_row_bytes(W) : mf_plane_bytes(W);
	uint8_t *ph = malloc(PB), *pl = malloc(PB);
	const int16_t *line = comp + W;
	if(!pitched)
	{
		for(int i = 0; i < PB; i++) { ph[i] = (uint8_t) rng(); pl[i] = (uint8_t) rng(); }   /* garbage past the window */
		for(int i = 0; i < mf_window_bytes(W); i++)
		{
			const int v = line[i - MF_LEAD];
			ph[i] = (uint8_t) ((v >> 8) & 0xFF);
			pl[i] = (uint8_t) (v & 0xFF);
		}
	}
	else
	{
		/* three rows of `pitch` bytes; every "thread" (x0 = 0, 4, ...) of every line writes its own
		 * samples and, at the line's edges, the neighbours' halos - k_raster's pitched branch */
		const int pitch = mf_pitch(W), nrows = 3;
		if(pitch < PB || pitch % 16) { printf("bad pitch\n"); return(1); }
		uint8_t *rows_h = malloc((size_t) nrows * pitch), *rows_l = malloc((size_t) nrows * pitch);
		uint8_t *wr = calloc((size_t) nrows * pitch, 1);
		for(int i = 0; i < nrows * pitch; i++) { rows_h[i] = (uint8_t) rng(); rows_l[i] = (uint8_t) rng(); }   /* stale bytes */
		for(int b = 0; b < nrows; b++) for(int x0 = 0; x0 < W; x0 += 4)
		{
			const int16_t *val = comp + b * W + x0;
			uint8_t *rh = rows_h + (size_t) b * pitch, *rl = rows_l + (size_t) b * pitch, *rw = wr + (size_t) b * pitch;
			if(x0 + 4 <= W)
				for(int k = 0; k < 4; k++) { rh[MF_LEAD + x0 + k] = (val[k] >> 8) & 0xFF; rl[MF_LEAD + x0 + k] = val[k] & 0xFF; rw[MF_LEAD + x0 + k]++; }
			const int head = x0 < MF_LEAD && b > 0, tail = x0 + 4 > W - MF_LEAD && b + 1 < nrows;
			if(head || tail || x0 + 4 > W)
				for(int k = 0; k < 4; k++)
				{
					const int x = x0 + k;
					if(x >= W) break;
					const uint8_t hi = (val[k] >> 8) & 0xFF, lo = val[k] & 0xFF;
					if(x0 + 4 > W) { rh[MF_LEAD + x] = hi; rl[MF_LEAD + x] = lo; rw[MF_LEAD + x]++; }
					if(head && x < MF_LEAD) { rh[-pitch + MF_LEAD + W + x] = hi; rl[-pitch + MF_LEAD + W + x] = lo; rw[-pitch + MF_LEAD + W + x]++; }
					if(tail && x >= W - MF_LEAD) { rh[pitch + x - (W - MF_LEAD)] = hi; rl[pitch + x - (W - MF_LEAD)] = lo; rw[pitch + x - (W - MF_LEAD)]++; }
				}
		}
		/* the middle row: every byte the non-zero taps can reach is written exactly once */
		for(int i = 0; i < W + 2 * MF_LEAD; i++)
			if(wr[pitch + i] != 1) { printf("row byte %d written %d times\n", i, wr[pitch + i]); return(1); }
		memcpy(ph, rows_h + pitch, PB);
		memcpy(pl, rows_l + pitch, PB);
	}

	uint32_t atab[MF_ATAB_WORDS];
	mf_build_atab(taps[0], taps[1], atab);
	uint32_t *fir = calloc((size_t) mf_tiles(W) * 4 * MF_ROWW, sizeof(uint32_t));
	uint8_t *seen = calloc((size_t) mf_tiles(W) * MF_TILE, 1);
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
			if(x < 0 || x >= mf_tiles(W) * MF_TILE) { printf("output index %d out of the tiles\n", x); return(1); }
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
