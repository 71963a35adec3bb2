/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Harness that drives the UNMODIFIED reference encoder through its own public
 * API: vid_init() / av_test_open() / vid_next_line() / vid_free()
 * (reference video.h:510-516, av_test.h:21, hacktv.c:1440-1601). It is linked
 * against objects compiled in place from /root/reference/src by
 * oracle/Makefile; nothing from the reference is copied into this repo.
 *
 * Uses:
 *   - dump mode: write the emitted int16 stream (I only for real modes, IQ
 *     for complex modes - what rf_file.c:97-116/226-233 would put in a file)
 *     for a window of lines -> golden vectors / parity checks;
 *   - bench mode: time vid_next_line() with no sink I/O -> cpu_baseline
 *     ("kind": "reference") in bench.py;
 *   - custom source: feed arbitrary RGB32 frames and 32 kHz stereo PCM through
 *     the reference's av_t callbacks (reference av.h:64-116) so parity can be
 *     checked on inputs other than the built-in test pattern.
 *
 * Config overrides mirror hacktv.c:1107-1437 for the options in scope.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include <unistd.h>
#include "hacktv.h"
#include "av_test.h"

typedef struct {
	int width, height;
	uint32_t *frames;
	size_t nframes, cur;
	int16_t *audio;
	size_t audio_samples;   /* stereo pairs in the whole file */
	size_t audio_block;     /* pairs handed out per read */
	size_t audio_pos;
} src_t;

static int _src_read_video(void *ctx, av_frame_t *frame)
{
	src_t *s = ctx;
	uint32_t *fb = s->frames + (size_t) s->width * s->height * s->cur;
	s->cur = (s->cur + 1) % s->nframes;
	av_frame_init(frame, s->width, s->height, fb, 1, s->width);
	return(AV_OK);
}

static int _src_read_audio(void *ctx, int16_t **samples, size_t *nsamples)
{
	src_t *s = ctx;
	size_t n = s->audio_block;
	if(s->audio_pos + n > s->audio_samples) n = s->audio_samples - s->audio_pos;
	*samples = s->audio + s->audio_pos * 2;
	*nsamples = n;
	s->audio_pos += n;
	if(s->audio_pos >= s->audio_samples) s->audio_pos = 0;
	return(AV_OK);
}

static int _src_close(void *ctx)
{
	src_t *s = ctx;
	free(s->frames);
	free(s->audio);
	free(s);
	return(AV_OK);
}

static void *_slurp(const char *fn, size_t *len)
{
	FILE *f = fopen(fn, "rb");
	void *d;
	if(!f) { perror(fn); exit(2); }
	fseek(f, 0, SEEK_END);
	*len = ftell(f);
	fseek(f, 0, SEEK_SET);
	d = malloc(*len ? *len : 1);
	if(fread(d, 1, *len, f) != *len) { perror("fread"); exit(2); }
	fclose(f);
	return(d);
}

static double _now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return(ts.tv_sec + ts.tv_nsec * 1e-9);
}

static void _usage(void)
{
	fprintf(stderr,
		"ref_harness -m MODE -s RATE [--pixelrate N] [--filter] [--noaudio] [--nonicam]\n"
		"            [--nocolour] [--offset HZ] [--swap-iq] [--level F] [--volume F] [--passthru FILE] [--wss MODE]\n"
		"            [--skip LINES] [--lines N] [-o FILE] [--bench] [--geometry]\n"
		"            [--frames FILE.rgb32] [--audio FILE.s16le] [--audio-block N]\n");
	exit(2);
}

int main(int argc, char **argv)
{
	static hacktv_t s;
	const vid_configs_t *vc;
	vid_config_t conf;
	const char *mode = "i", *out = NULL, *frames_fn = NULL, *audio_fn = NULL, *passthru_fn = NULL, *wss_mode = NULL;
	unsigned int rate = 16000000, pixelrate = 0;
	int filter = 0, noaudio = 0, nonicam = 0, nocolour = 0, swap_iq = 0, bench = 0, geometry = 0;
	long long offset = 0, skip = 0, lines = 625;
	double level = 1.0, volume = 1.0;
	size_t audio_block = 0;
	int i, r, complex;
	FILE *fo = NULL;
	long long n;
	double t0, t1;
	uint64_t samples = 0;
	int16_t *real = NULL;

	for(i = 1; i < argc; i++)
	{
		if(!strcmp(argv[i], "-m") && i + 1 < argc) mode = argv[++i];
		else if(!strcmp(argv[i], "-s") && i + 1 < argc) rate = strtoul(argv[++i], NULL, 10);
		else if(!strcmp(argv[i], "--pixelrate") && i + 1 < argc) pixelrate = strtoul(argv[++i], NULL, 10);
		else if(!strcmp(argv[i], "--filter")) filter = 1;
		else if(!strcmp(argv[i], "--noaudio")) noaudio = 1;
		else if(!strcmp(argv[i], "--nonicam")) nonicam = 1;
		else if(!strcmp(argv[i], "--nocolour")) nocolour = 1;
		else if(!strcmp(argv[i], "--swap-iq")) swap_iq = 1;
		else if(!strcmp(argv[i], "--offset") && i + 1 < argc) offset = strtoll(argv[++i], NULL, 10);
		else if(!strcmp(argv[i], "--level") && i + 1 < argc) level = atof(argv[++i]);
		else if(!strcmp(argv[i], "--volume") && i + 1 < argc) volume = atof(argv[++i]);
		else if(!strcmp(argv[i], "--skip") && i + 1 < argc) skip = strtoll(argv[++i], NULL, 10);
		else if(!strcmp(argv[i], "--lines") && i + 1 < argc) lines = strtoll(argv[++i], NULL, 10);
		else if(!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
		else if(!strcmp(argv[i], "--bench")) bench = 1;
		else if(!strcmp(argv[i], "--geometry")) geometry = 1;
		else if(!strcmp(argv[i], "--frames") && i + 1 < argc) frames_fn = argv[++i];
		else if(!strcmp(argv[i], "--audio") && i + 1 < argc) audio_fn = argv[++i];
		else if(!strcmp(argv[i], "--audio-block") && i + 1 < argc) audio_block = strtoul(argv[++i], NULL, 10);
		else if(!strcmp(argv[i], "--passthru") && i + 1 < argc) passthru_fn = argv[++i];
		else if(!strcmp(argv[i], "--wss") && i + 1 < argc) wss_mode = argv[++i];
		else _usage();
	}

	for(vc = vid_configs; vc->id != NULL; vc++)
	{
		if(strcmp(mode, vc->id) == 0) break;
	}
	if(vc->id == NULL) { fprintf(stderr, "Unrecognised TV mode.\n"); return(2); }

	/* hacktv.c:1107-1437 for the options in scope */
	memcpy(&conf, vc->conf, sizeof(vid_config_t));
	if(nocolour && (conf.colour_mode == VID_PAL || conf.colour_mode == VID_SECAM || conf.colour_mode == VID_NTSC))
	{
		conf.colour_mode = VID_NONE;
	}
	if(noaudio)
	{
		conf.fm_mono_level = conf.fm_left_level = conf.fm_right_level = 0;
		conf.am_audio_level = conf.nicam_level = conf.dance_level = 0;
		conf.fm_mono_carrier = conf.fm_left_carrier = conf.fm_right_carrier = 0;
		conf.nicam_carrier = conf.dance_carrier = conf.am_mono_carrier = 0;
	}
	if(nonicam) { conf.nicam_level = 0; conf.nicam_carrier = 0; }
	conf.level *= (float) level;
	if(filter) conf.vfilter = 1;
	conf.swap_iq = swap_iq;
	conf.offset = offset;
	conf.wss = (char *) wss_mode;                          /* hacktv.c --wss; video.c:4234-4242 */
	conf.passthru = (char *) passthru_fn;                 /* hacktv.c --passthru; video.c:4607-4634 */
	conf.volume = (float) volume * 256 + 0.5;
	conf.raw_bb_white_level = INT16_MAX;

	memset(&s, 0, sizeof(s));
	t0 = _now();
	r = vid_init(&s.vid, rate, pixelrate, &conf);
	t1 = _now();
	if(r != VID_OK) { fprintf(stderr, "vid_init failed (%d)\n", r); return(1); }

	complex = s.vid.conf.output_type == RF_INT16_COMPLEX;

	if(geometry)
	{
		printf("{\"width\": %d, \"half_width\": %d, \"lines\": %d, \"active_left\": %d, \"active_width\": %d, "
		       "\"active_lines\": %d, \"burst_left\": %d, \"burst_width\": %d, \"complex\": %d, "
		       "\"white\": %d, \"black\": %d, \"blank\": %d, \"sync\": %d, \"olines\": %d, \"nthreads\": %d, "
		       "\"init_s\": %.4f}\n",
			s.vid.width, s.vid.half_width, s.vid.conf.lines, s.vid.active_left, s.vid.active_width,
			s.vid.conf.active_lines, s.vid.burst_left, s.vid.burst_width, complex,
			s.vid.white_level, s.vid.black_level, s.vid.blanking_level, s.vid.sync_level,
			s.vid.olines, s.vid.nthreads, t1 - t0);
		fflush(stdout);
	}

	/* hacktv.c:1503-1518 */
	s.vid.av = (av_t) {
		.frame_rate = (r64_t) { s.vid.conf.frame_rate.num * (s.vid.conf.interlace ? 2 : 1), s.vid.conf.frame_rate.den },
		.display_aspect_ratios = { s.vid.conf.frame_aspects[0], s.vid.conf.frame_aspects[1] },
		.fit_mode = AV_FIT_STRETCH,
		.width = s.vid.active_width,
		.height = s.vid.conf.active_lines,
		.sample_rate = (r64_t) { HACKTV_AUDIO_SAMPLE_RATE, 1 },
	};

	if(frames_fn || audio_fn)
	{
		src_t *src = calloc(1, sizeof(src_t));
		size_t len;

		src->width = s.vid.active_width;
		src->height = s.vid.conf.active_lines;
		if(frames_fn)
		{
			src->frames = _slurp(frames_fn, &len);
			src->nframes = len / ((size_t) src->width * src->height * 4);
			if(src->nframes < 1) { fprintf(stderr, "frame file too short\n"); return(2); }
			s.vid.av.read_video = _src_read_video;
		}
		if(audio_fn)
		{
			src->audio = _slurp(audio_fn, &len);
			src->audio_samples = len / 4;
			src->audio_block = audio_block ? audio_block : src->audio_samples;
			if(src->audio_samples < 1) { fprintf(stderr, "audio file too short\n"); return(2); }
			s.vid.av.read_audio = _src_read_audio;
		}
		s.vid.av.av_source_ctx = src;
		s.vid.av.close = _src_close;
	}
	else
	{
		if(av_test_open(&s.vid.av) != AV_OK) { fprintf(stderr, "av_test_open failed\n"); return(1); }
	}

	if(out)
	{
		fo = strcmp(out, "-") == 0 ? stdout : fopen(out, "wb");
		if(!fo) { perror(out); return(1); }
		real = malloc(sizeof(int16_t) * s.vid.max_width);
	}

	/* Skip (still generated - the encoder is a stream) */
	for(n = 0; n < skip; n++)
	{
		if(vid_next_line(&s.vid) == NULL) break;
	}

	t0 = _now();
	for(n = 0; n < lines; n++)
	{
		vid_line_t *l = vid_next_line(&s.vid);
		if(l == NULL) break;
		samples += l->width;

		if(fo)
		{
			if(complex)
			{
				/* rf_file.c:226-233 */
				fwrite(l->output, sizeof(int16_t) * 2, l->width, fo);
			}
			else
			{
				/* rf_file.c:97-116 */
				int x;
				for(x = 0; x < l->width; x++) real[x] = l->output[x * 2];
				fwrite(real, sizeof(int16_t), l->width, fo);
			}
		}
	}
	t1 = _now();

	if(bench)
	{
		printf("{\"mode\": \"%s\", \"rate\": %u, \"filter\": %d, \"lines\": %lld, \"samples\": %llu, "
		       "\"seconds\": %.6f, \"msamples_per_s\": %.4f, \"threads\": %d}\n",
			mode, rate, filter, n, (unsigned long long) samples, t1 - t0,
			samples / (t1 - t0) / 1e6, 1 + s.vid.nthreads);
	}

	if(fo && fo != stdout) fclose(fo);
	free(real);

	/* Deliberately no vid_free(): the reference's thread shutdown handshake
	 * (video.c:3604-3613 vs 4713-4721) can deadlock when the main thread sees
	 * nthreads == 0 before the worker reaches its final barrier. The process
	 * is about to exit anyway. */
	fflush(NULL);
	_exit(0);
}
