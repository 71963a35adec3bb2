/* CPU check of k_resample's arithmetic (hacktv_b200/csrc/htv_resample.h): rows of a random pixel-rate
 * stream laid out as the raster context leaves them (row b = raster line first - 1 + b), resampled row by
 * row with the kernel's closed form, against a streaming restatement of the reference's loop
 * (fir_int16_process fir.c:304-355: consume an input when d >= I, emit while d < I) run over the same
 * stream from its start. Taps are random int16 in the reference's layout. No GPU involved.
 *
 * usage: resample_emu Wp I D seed  -> "OK <outputs checked>" */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "htv_resample.h"

static uint32_t st;
static uint32_t rng(void) { st = st * 1664525u + 1013904223u; return(st >> 8); }

int main(int argc, char **argv)
{
	const int Wp = argc > 1 ? atoi(argv[1]) : 864, I = argc > 2 ? atoi(argv[2]) : 32, D = argc > 3 ? atoi(argv[3]) : 27;
	const int rows = 6;
	st = argc > 4 ? (uint32_t) atoi(argv[4]) : 1;
	if(((long long) Wp * I) % D) { printf("line width would vary\n"); return(2); }
	const int Ws = (int) ((long long) Wp * I / D);
	const int ntaps = (21 * I) | 1, A = (ntaps + I - 1) / I;
	int16_t *taps = calloc((size_t) I * A, sizeof(int16_t));
	for(int i = 0; i < I * A; i++) taps[i] = (int16_t) ((rng() % 4001) - 2000);
	/* 64 zeros in front: the stream starts at row 0 with zero history (the reference's calloc'd window) */
	int16_t *buf = calloc((size_t) rows * Wp + 64, sizeof(int16_t)), *comp = buf + 64;
	for(int i = 0; i < rows * Wp; i++) comp[i] = (int16_t) (rng() & 0xFFFF);

	/* streaming form */
	int16_t *ref = malloc(sizeof(int16_t) * (size_t) rows * Ws);
	{
		int16_t *win = calloc(A, sizeof(int16_t));
		int d = I, n = 0;
		long long in = 0;
		while(in < (long long) rows * Wp || d < I)
		{
			if(d >= I)
			{
				if(in >= (long long) rows * Wp) break;
				d -= I;
				memmove(win, win + 1, sizeof(int16_t) * (A - 1));
				win[A - 1] = comp[in++];
			}
			for(; d < I; d += D)
			{
				int a = 0;
				for(int c = 0; c < A; c++) a += win[c] * taps[d * A + c];
				a >>= 15;
				ref[n++] = a < -32768 ? -32768 : (a > 32767 ? 32767 : a);
			}
		}
		if(n != rows * Ws) { printf("streaming form produced %d outputs, expected %d\n", n, rows * Ws); return(1); }
		free(win);
	}

	/* closed form, as k_resample: output row b <- comp rows b, b + 1 */
	long checked = 0;
	for(int b = 0; b + 1 < rows; b++)
	{
		const int16_t *in = comp + (size_t) (b + 1) * Wp;
		for(int x = 0; x < Ws; x++)
		{
			const int got = rs_output(in, x, I, D, A, taps), want = ref[(size_t) (b + 1) * Ws + x];
			if(got != want) { printf("MISMATCH row %d x %d: %d != %d\n", b, x, got, want); return(1); }
			checked++;
		}
	}
	/* and the stream's very first line: zero history */
	for(int x = 0; x < Ws; x++)
		if(rs_output(comp, x, I, D, A, taps) != ref[x]) { printf("MISMATCH first line x %d\n", x); return(1); }
	printf("OK %ld\n", checked);
	return(0);
}
