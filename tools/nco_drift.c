#define _GNU_SOURCE
// Experiment (round 2): the phase of the reference's Q31 FM NCO (a floor in every step + a renormalisation every
// 32 767 samples, ref video.c:2259-2276) against the closed form the CUDA path uses - the running sum of
// atan2(rounded LUT entry). With a silent / periodic modulating signal and the carrier at a small rational fraction
// of the sample rate (NTSC-M at 13.5 Msps: 4.5 MHz = fs / 3) the phasor revisits the same few states, its floor
// errors stop averaging out and the reference drifts away from its own ideal carrier; per-segment simulation from
// approximate start states recovers only part of it (the drift lives in the low bits of every state).
//     gcc -O2 -o nco_drift tools/nco_drift.c -lm && ./nco_drift 13.5e6 4.5e6 25e3 10 && ./nco_drift 16e6 6e6 50e3 10
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>
typedef struct { int32_t i, q; } c32;
static inline void mul(c32 *r, const c32 *a, const c32 *b)
{
	int64_t i = (int64_t) a->i * b->i - (int64_t) a->q * b->q, q = (int64_t) a->i * b->q + (int64_t) a->q * b->i;
	r->i = i >> 31; r->q = q >> 31;
}
int main(int argc, char **argv)
{
	double fs = argc > 1 ? atof(argv[1]) : 13.5e6, fc = argc > 2 ? atof(argv[2]) : 4.5e6, dev = argc > 3 ? atof(argv[3]) : 25e3;
	double seconds = argc > 4 ? atof(argv[4]) : 10;
	static c32 lut[65536]; static long double ang[65536];
	for(int r = -32768; r <= 32767; r++)
	{
		double d = 2.0 * M_PI / fs * (fc + (double) r / 32767 * dev);
		lut[r + 32768].i = lround(cos(d) * INT32_MAX); lut[r + 32768].q = lround(sin(d) * INT32_MAX);
		ang[r + 32768] = atan2l((long double) lut[r + 32768].q, (long double) lut[r + 32768].i);
	}
	c32 ph = { INT32_MAX, 0 };
	int counter = 32767;
	long long N = (long long) (fs * seconds);
	long double model = 0;            // closed-form accumulated phase
	long double corr = 0;             // sum of per-segment corrections measured from approximate starts
	long long interp = 0; int sample = 0; long long j = 0;
	// per-segment independent simulation state
	long long seg_n = 0;
	c32 sim = ph; int sim_counter = counter;
	long double seg_model = 0;
	long double sim_start_ang = 0;
	for(long long n = 0; n < N; n++)
	{
		interp += 32000;
		if(interp >= (long long) fs)
		{
			interp -= (long long) fs;
			// close the previous segment: measured rotation of the independent simulation vs the model
			if(seg_n > 0)
			{
				long double a1 = atan2l((long double) sim.q, (long double) sim.i);
				long double d = a1 - sim_start_ang - seg_model;
				d -= 2 * M_PIl * roundl(d / (2 * M_PIl));
				corr += d;
			}
			j++;
			sample = ((j / 16000) & 1) ? (int) lround(8000 * sin(2 * M_PI * 1000.0 * j / 32000.0)) : 0;
			// restart the independent simulation from the APPROXIMATE state: model phase + corrections so far,
			// amplitude from the renormalisation counter
			{
				long double a = model + corr;
				double amp = 2147483647.0 * (1.0 - (double) (32767 - counter) * 4.656612873077393e-10);
				sim.i = (int32_t) floor(amp * cosl(a)); sim.q = (int32_t) floor(amp * sinl(a));
				sim_counter = counter; seg_model = 0; seg_n = 0;
				sim_start_ang = atan2l((long double) sim.q, (long double) sim.i);
			}
		}
		mul(&ph, &ph, &lut[sample + 32768]);
		mul(&sim, &sim, &lut[sample + 32768]);
		model += ang[sample + 32768]; seg_model += ang[sample + 32768]; seg_n++;
		if(--counter == 0)
		{
			double ra = atan2(ph.q, ph.i);
			ph.i = lround(cos(ra) * INT32_MAX); ph.q = lround(sin(ra) * INT32_MAX);
			counter = 32767;
		}
		if(--sim_counter == 0)
		{
			double ra = atan2(sim.q, sim.i);
			sim.i = lround(cos(ra) * INT32_MAX); sim.q = lround(sin(ra) * INT32_MAX);
			sim_counter = 32767;
		}
		if((n + 1) % (long long) (fs * (seconds / 10)) == 0)
		{
			long double a = atan2l((long double) ph.q, (long double) ph.i);
			long double e0 = a - model; e0 -= 2 * M_PIl * roundl(e0 / (2 * M_PIl));
			long double e1 = a - (model + corr); e1 -= 2 * M_PIl * roundl(e1 / (2 * M_PIl));
			printf("t=%6.2f s  true-model = %+.3Le rad   true-(model+segment corrections) = %+.3Le rad\n", (double) (n + 1) / fs, e0, e1);
		}
	}
	return(0);
}
