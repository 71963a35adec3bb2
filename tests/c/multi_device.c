/* In-process multi-channel through the C-ABI (channel c -> device c mod N, one host
 * thread per device): N pthreads, each driving its own encoder placed with htv_init_on() on device
 * (thread mod htv_device_count()), none of them ever calling cudaSetDevice. All channels carry the same
 * programme, so every thread must produce the bytes of the single-threaded run on device 0 (+-0: the same
 * kernels on the same inputs). Usage: multi_device [threads] [lines]; prints one JSON line, exit 0 = pass. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "hacktv_b200.h"

typedef struct { int id, device, lines, rc; int16_t *out; size_t nvalues; } job_t;

static int render(job_t *j)
{
	const htv_config_t *mc = htv_find_mode("i");
	htv_config_t conf;
	htv_t *v = NULL;
	size_t n = 0;
	int done = 0;
	if(!mc) return(-1);
	memcpy(&conf, mc, sizeof(conf));
	conf.vfilter = 1;
	if(htv_init_on(&v, j->device, 16000000, 0, &conf) != HTV_OK) return(-2);
	if(htv_device(v) != j->device) { htv_free(v); return(-3); }
	if(htv_av_test_open(htv_av(v)) != HTV_OK) { htv_free(v); return(-4); }
	j->nvalues = (size_t) j->lines * htv_samples_per_line(v) * 2;
	j->out = malloc(j->nvalues * sizeof(int16_t));
	/* uneven pieces, so the threads interleave their calls */
	while(done < j->lines)
	{
		int piece = 97 + 31 * j->id;
		if(piece > j->lines - done) piece = j->lines - done;
		if(htv_render_host(v, piece, j->out + (size_t) done * htv_samples_per_line(v) * 2, &n) != HTV_OK) { htv_free(v); return(-5); }
		done += piece;
	}
	htv_free(v);
	return(0);
}

static void *thread_main(void *arg) { job_t *j = arg; j->rc = render(j); return(NULL); }

int main(int argc, char **argv)
{
	int nthreads = argc > 1 ? atoi(argv[1]) : 2, lines = argc > 2 ? atoi(argv[2]) : 700, i, bad = 0;
	const int ndev = htv_device_count();
	job_t ref, *jobs;
	pthread_t *th;
	if(ndev < 1) { fprintf(stderr, "no CUDA device\n"); return(2); }
	if(nthreads < 1 || nthreads > 64) return(2);
	memset(&ref, 0, sizeof(ref));
	ref.id = 0; ref.device = 0; ref.lines = lines;
	if(render(&ref) != 0) { fprintf(stderr, "reference run failed\n"); return(1); }
	jobs = calloc(nthreads, sizeof(job_t));
	th = calloc(nthreads, sizeof(pthread_t));
	for(i = 0; i < nthreads; i++)
	{
		jobs[i].id = i + 1; jobs[i].device = i % ndev; jobs[i].lines = lines;
		pthread_create(&th[i], NULL, thread_main, &jobs[i]);
	}
	for(i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
	for(i = 0; i < nthreads; i++)
	{
		if(jobs[i].rc != 0 || jobs[i].nvalues != ref.nvalues || memcmp(jobs[i].out, ref.out, ref.nvalues * sizeof(int16_t)) != 0) bad++;
	}
	printf("{\"threads\": %d, \"devices\": %d, \"lines\": %d, \"mismatching_threads\": %d}\n", nthreads, ndev, lines, bad);
	return(bad ? 1 : 0);
}
