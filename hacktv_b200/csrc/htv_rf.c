/* hacktv_b200 - RF sink dispatch and the int16 file sink (ref rf.c:23-51,
 * rf_file.c:97-116, 226-233, 290-373). Only what the -o file path needs. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "hacktv_b200.h"

typedef struct {
	FILE *f;
	int complex;
	int16_t *buf;
} file_sink_t;

int htv_rf_write(htv_rf_t *s, const int16_t *iq_data, size_t samples)
{
	if(s && s->write) return(s->write(s->ctx, iq_data, samples));
	return(HTV_ERROR);
}

int htv_rf_close(htv_rf_t *s)
{
	if(s && s->close) return(s->close(s->ctx));
	return(HTV_OK);
}

static int file_write(void *ctx, const int16_t *iq, size_t samples)
{
	file_sink_t *fs = ctx;
	if(fs->complex)
	{
		if(fwrite(iq, sizeof(int16_t) * 2, samples, fs->f) != samples) return(HTV_ERROR);
		return(HTV_OK);
	}
	while(samples)
	{
		/* real signals: the Q half of each pair is dropped */
		size_t i, n = samples < 4096 ? samples : 4096;
		for(i = 0; i < n; i++, iq += 2) fs->buf[i] = iq[0];
		if(fwrite(fs->buf, sizeof(int16_t), n, fs->f) != n) return(HTV_ERROR);
		samples -= n;
	}
	return(HTV_OK);
}

static int file_close(void *ctx)
{
	file_sink_t *fs = ctx;
	if(fs->f && fs->f != stdout) fclose(fs->f);
	else if(fs->f) fflush(fs->f);
	free(fs->buf);
	free(fs);
	return(HTV_OK);
}

int htv_rf_file_open(htv_rf_t *s, const char *filename, int complex)
{
	file_sink_t *fs;
	if(!s) return(HTV_ERROR);
	if(filename == NULL)
	{
		fprintf(stderr, "No output filename provided.\n");
		return(HTV_ERROR);
	}
	fs = calloc(1, sizeof(*fs));
	if(!fs) return(HTV_OUT_OF_MEMORY);
	fs->complex = complex != 0;
	fs->f = strcmp(filename, "-") == 0 ? stdout : fopen(filename, "wb");
	if(!fs->f) { perror("fopen"); free(fs); return(HTV_ERROR); }
	fs->buf = malloc(sizeof(int16_t) * 4096);
	s->ctx = fs;
	s->write = file_write;
	s->close = file_close;
	return(HTV_OK);
}
