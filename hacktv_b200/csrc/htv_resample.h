/* The --pixelrate resampler (ref _init_vresampler video.c:3627-3651, fir_int16_process fir.c:304-355
 * with the taps of fir_int16_resampler_init fir.c:393-428) in closed form of the output index, shared by
 * k_resample (htv_kernels.cu) and its CPU check (tests/resample_emu.c).
 *
 * The reference consumes an input whenever its counter d >= I (d -= I) and emits outputs while d < I
 * (d += D), starting from d = I: output j follows input floor(j D / I) and uses phase (j D) mod I.
 * When Ws D = Wp I every line is a whole number of periods, so within a line output x uses the
 * rs_ataps inputs ending at floor(x D / I) of that line (reaching back into the previous line) and
 * phase (x D) mod I. taps[phase * A + c] multiplies input (newest - A + 1 + c). */
#ifndef HTV_RESAMPLE_H
#define HTV_RESAMPLE_H

#include <stdint.h>

#if defined(__CUDACC__)
#define RS_HD __host__ __device__ __forceinline__
#else
#define RS_HD static inline
#endif

/* in: input 0 of the resampled line (the previous line's samples precede it in memory) */
RS_HD int rs_output(const int16_t *in, int x, int I, int D, int A, const int16_t *taps)
{
	const int xd = x * D, pos = xd / I, ph = xd - pos * I;              /* x D < 2^31 for any line width here */
	const int16_t *w = in + pos - A + 1;
	const int16_t *t = taps + ph * A;
	int a = 0, c;
	for(c = 0; c < A; c++) a += (int) w[c] * (int) t[c];
	a >>= 15;
	return(a < -32768 ? -32768 : (a > 32767 ? 32767 : a));
}

#endif
