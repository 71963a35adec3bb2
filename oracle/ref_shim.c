/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Link-time shims used when the UNMODIFIED reference (fsphil/hacktv, compiled
 * in place from /root/reference/src by oracle/Makefile) is built without
 * libav*: the three av_ffmpeg_* entry points hacktv.c refers to
 * (reference av_ffmpeg.h:21-23) are stubbed, exactly as SURVEY.md §8(c)
 * describes.
 *
 * The optional "zero heap" allocator (-DREF_ZERO_HEAP, linked with
 * -Wl,--wrap=malloc) makes the reference's two documented heap over-reads
 * deterministic WITHOUT touching its sources: every malloc() is zero-filled
 * and over-allocated by 256 bytes, so
 *   - fir_int16_process_block()'s read of ataps/2 samples past
 *     chrominance_buffer (reference fir.c:365-372, video.c:3019-3020), and
 *   - the SECAM bell LUT's entry 65535 of a 65535-entry allocation
 *     (reference video.c:4115,4122-4127)
 * land in zeroed, owned memory. This is the "ASan-clean" parity oracle of
 * out-of-line chroma samples are 0.
 */
#include <stddef.h>
#include <stdlib.h>

struct av_t_fwd;

int av_ffmpeg_open(void *av, char *input_url, char *format, char *options)
{
	(void) av; (void) input_url; (void) format; (void) options;
	return(-1); /* AV_ERROR */
}

void av_ffmpeg_init(void) { }
void av_ffmpeg_deinit(void) { }

#ifdef REF_ZERO_HEAP
void *__real_malloc(size_t n);

void *__wrap_malloc(size_t n)
{
	return(calloc(1, n + 256));
}
#endif
