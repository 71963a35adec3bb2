/* hacktv_b200 - the built-in test source (colour bars + 1 kHz tone), the synthetic
 * input every BASELINE config is quoted on. Restates ref av_test.c:71-205 behind
 * the htv_av_t pull interface; the picture never changes, so its serial is constant
 * and the encoder uploads it once. */
#include <stdlib.h>
#include <math.h>
#include "hacktv_b200.h"

typedef struct {
	int width, height;
	uint32_t *video;
	int16_t *audio;
	size_t audio_pairs;
} test_src_t;

static const char *logo_rows[9] = {
	"                                                ",
	" ##  ##    ##     ####   ##  ##  ######  ##  ## ",
	" ##  ##   ####   ##  ##  ## ##     ##    ##  ## ",
	" ##  ##  ##  ##  ##      ####      ##    ##  ## ",
	" ######  ######  ##      ###       ##    ##  ## ",
	" ##  ##  ##  ##  ##      ####      ##    ##  ## ",
	" ##  ##  ##  ##  ##  ##  ## ##     ##     ####  ",
	" ##  ##  ##  ##   ####   ##  ##    ##      ##   ",
	"                                                ",
};

void htv_test_pattern(int width, int height, uint32_t *rgb)
{
	/* 75% bars, a red strip, a ramp, an 8-step grey scale, and the 48x9 logo at x4 */
	static const uint32_t bars[8] = { 0x000000, 0x0000BF, 0xBF0000, 0xBF00BF, 0x00BF00, 0x00BFBF, 0xBFBF00, 0xFFFFFF };
	int x, y;
	for(y = 0; y < height; y++)
	{
		uint32_t *row = rgb + (size_t) y * width;
		for(x = 0; x < width; x++)
		{
			uint32_t c;
			if(y < height - 140) c = bars[7 - x * 8 / width];
			else if(y < height - 120) c = 0xBF0000;
			else
			{
				uint32_t g = x * 0xFF / (width - 1);
				if(y >= height - 100) { g &= 0xE0; g |= (g >> 3) | (g >> 6); }
				c = g << 16 | g << 8 | g;
			}
			row[x] = c;
		}
	}
	if(width >= 48 * 4 && height >= 9 * 4)
	{
		const int left = (width - 48 * 4) / 2, top = height / 10;
		for(y = 0; y < 9 * 4; y++)
		{
			for(x = 0; x < 48 * 4; x++)
			{
				rgb[(size_t) (top + y) * width + left + x] = logo_rows[y / 4][x / 4] == ' ' ? 0x000000 : 0xFFFFFF;
			}
		}
	}
}

size_t htv_test_tone_pairs(void) { return(32000 * 64 / 100 * 10); }   /* ten 640 ms segments */

void htv_test_tone(int16_t *pcm)
{
	/* 1 kHz at 0.1 FS; left muted in segment 0, right muted in segments 2 and 4 */
	const double w = 1000.0 * 2 * M_PI * 1 / 32000;
	const int seg = 32000 * 64 / 100;
	int x;
	for(x = 0; x < seg * 10; x++)
	{
		int16_t v = sin(x * w) * INT16_MAX * 0.1;
		int k = x / seg;
		pcm[x * 2 + 0] = k == 0 ? 0 : v;
		pcm[x * 2 + 1] = (k == 2 || k == 4) ? 0 : v;
	}
}

static int test_read_video(void *ctx, htv_frame_t *frame)
{
	test_src_t *s = ctx;
	frame->width = s->width;
	frame->height = s->height;
	frame->framebuffer = s->video;
	frame->serial = 1;
	return(HTV_OK);
}

static int test_read_audio(void *ctx, const int16_t **samples, size_t *npairs)
{
	test_src_t *s = ctx;
	*samples = s->audio;
	*npairs = s->audio_pairs;
	return(HTV_OK);
}

static int test_close(void *ctx)
{
	test_src_t *s = ctx;
	free(s->video);
	free(s->audio);
	free(s);
	return(HTV_OK);
}

int htv_av_test_open(htv_av_t *av)
{
	test_src_t *s;
	if(!av || av->width < 1 || av->height < 1) return(HTV_ERROR);
	s = calloc(1, sizeof(*s));
	if(!s) return(HTV_OUT_OF_MEMORY);
	s->width = av->width;
	s->height = av->height;
	s->video = malloc((size_t) s->width * s->height * sizeof(uint32_t));
	s->audio_pairs = htv_test_tone_pairs();
	s->audio = malloc(s->audio_pairs * 2 * sizeof(int16_t));
	if(!s->video || !s->audio) { test_close(s); return(HTV_OUT_OF_MEMORY); }
	htv_test_pattern(s->width, s->height, s->video);
	htv_test_tone(s->audio);
	av->ctx = s;
	av->read_video = test_read_video;
	av->read_audio = test_read_audio;
	av->close = test_close;
	return(HTV_OK);
}

/* ---- in-memory source ---------------------------------------------------
 * Caller-owned pictures (RGBx, width x height each, used cyclically one per video frame)
 * and 32 kHz stereo PCM (used cyclically, handed out `audio_block` pairs at a time).
 * The C counterpart of what the tests install through Python callbacks; lets a host
 * program (and bench.py's end-to-end leg) drive the encoder without an interpreter in
 * the per-frame path. `static_video` keeps the picture serial constant (one upload). */
typedef struct {
	int width, height;
	const uint32_t *frames;
	size_t nframes, cur;
	const int16_t *audio;
	size_t audio_pairs, audio_block, audio_pos;
	int static_video;
	uint64_t serial;
} mem_src_t;

static int mem_read_video(void *ctx, htv_frame_t *frame)
{
	mem_src_t *s = ctx;
	frame->width = s->width;
	frame->height = s->height;
	frame->framebuffer = s->frames + (size_t) s->width * s->height * s->cur;
	s->cur = (s->cur + 1) % s->nframes;
	frame->serial = s->static_video ? 1 : ++s->serial;
	return(HTV_OK);
}

static int mem_read_audio(void *ctx, const int16_t **samples, size_t *npairs)
{
	mem_src_t *s = ctx;
	size_t n = s->audio_block;
	if(s->audio_pos + n > s->audio_pairs) n = s->audio_pairs - s->audio_pos;
	*samples = s->audio + s->audio_pos * 2;
	*npairs = n;
	s->audio_pos += n;
	if(s->audio_pos >= s->audio_pairs) s->audio_pos = 0;
	return(HTV_OK);
}

static int mem_close(void *ctx) { free(ctx); return(HTV_OK); }

int htv_av_memory_open(htv_av_t *av, const uint32_t *frames, size_t nframes,
	const int16_t *audio, size_t audio_pairs, size_t audio_block, int static_video)
{
	mem_src_t *s;
	if(!av || av->width < 1 || av->height < 1) return(HTV_ERROR);
	s = calloc(1, sizeof(*s));
	if(!s) return(HTV_OUT_OF_MEMORY);
	s->width = av->width; s->height = av->height;
	s->frames = frames; s->nframes = nframes;
	s->audio = audio; s->audio_pairs = audio_pairs;
	s->audio_block = audio_block && audio_block < audio_pairs ? audio_block : audio_pairs;
	s->static_video = static_video;
	av->ctx = s;
	av->read_video = frames && nframes ? mem_read_video : NULL;
	av->read_audio = audio && audio_pairs ? mem_read_audio : NULL;
	av->close = mem_close;
	return(HTV_OK);
}
