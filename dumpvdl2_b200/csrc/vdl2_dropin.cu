/*
 * vdl2_dropin.cu — drop-in shim: the reference's own entry points for this path (include/vdl2_dropin.h) on top
 * of the batch API.  Host code only; written in the C subset of C++ so that it reads like the reference's C.
 *
 * Thread protocol kept from the reference (src/demod.c:299-301,339-347; src/dumpvdl2.c:117-135,1170):
 *   producer  process_buf_*():  wait(demods_ready); hand the buffer over; wait(samples_ready)
 *   consumers process_samples(): wait(demods_ready); wait(samples_ready); work
 * Here "hand the buffer over" = vdl2gpu_submit() (copy into the pinned ring, async H2D + kernels) and the
 * "work" of channel 0's thread = vdl2gpu_flush() + avlc_decoder_queue_push() of the frames, so that when main()
 * passes its final demods_ready barrier every frame has been pushed, exactly as with the CPU demodulators.
 * The other channel threads only keep the barrier counts right.
 */
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/vdl2gpu.h"
#include "../../include/vdl2_dropin.h"

extern "C" {
/* provided by the host program; weak so that the library loads without it (python, tests) */
extern pthread_barrier_t demods_ready __attribute__((weak));
extern pthread_barrier_t samples_ready __attribute__((weak));
void avlc_decoder_queue_push(vdl2_msg_metadata *metadata, octet_string_t *frame, int flags) __attribute__((weak));
#ifdef VDL2_DROPIN_USE_REFERENCE_HEADERS
extern dumpvdl2_config_t Config __attribute__((weak));
#endif

float *sbuf = NULL;
}

#define DROPIN_MAX_CHANNELS 65536

static struct {
	vdl2_channel_t *chan[DROPIN_MAX_CHANNELS];
	uint32_t freq[DROPIN_MAX_CHANNELS];
	uint32_t n_channels;
	uint32_t centerfreq, rate, oversample;
	vdl2gpu_ctx *ctx;
	int fmt;                      /* -1 until the first buffer shows which converter the front-end uses */
	int status;
	char *station_id;
	float max_ppm;
	pthread_mutex_t lock;
} D = { {0}, {0}, 0, 0, 0, 0, NULL, -1, 0, NULL, 0.f, PTHREAD_MUTEX_INITIALIZER };

extern "C" vdl2_channel_t *vdl2_channel_init(uint32_t centerfreq, uint32_t freq, uint32_t source_rate, uint32_t oversample) {
	/* src/demod.c:379-392: the struct is what main() keeps in its channel list and passes to pthread_create */
	vdl2_channel_t *v = (vdl2_channel_t *)calloc(1, sizeof(vdl2_channel_t));
	if(v == NULL || D.n_channels >= DROPIN_MAX_CHANNELS) { free(v); return NULL; }
	v->mag_nf = 2.0f;
	v->downmix_dphi = (uint32_t)(int)(((float)centerfreq - (float)freq) / (float)source_rate * 256.0f * 65536.0f);
	v->offset_tuning = (centerfreq != freq);
	v->oversample = (uint16_t)oversample;
	v->freq = freq;
	pthread_mutex_lock(&D.lock);
	D.centerfreq = centerfreq; D.rate = source_rate; D.oversample = oversample;
	D.chan[D.n_channels] = v;
	D.freq[D.n_channels] = freq;
	D.n_channels++;
	pthread_mutex_unlock(&D.lock);
	return v;
}

/* The tables these four build in the reference (src/demod.c:349-377,84-96) are computed inside
 * vdl2gpu_create(); the calls are kept so that main()'s init sequence (src/dumpvdl2.c:1149-1151) links. */
extern "C" void sincosf_lut_init(void) {}
extern "C" void input_lpf_init(uint32_t sample_rate) { (void)sample_rate; }
extern "C" void demod_sync_init(void) {}
extern "C" void process_buf_uchar_init(void) {}

extern "C" int rs_init(void) { return 0; }       /* GF tables live in the library (src/rs.c:27-30) */

extern "C" int rs_verify(uint8_t *data, int fec_octets) {
	/* src/rs.c:32-49 for callers outside the burst kernel: one block through the device RS routine */
	uint8_t *d_blk = NULL; int32_t *d_aux = NULL; int32_t h[2] = { fec_octets, 0 };
	int ret = -1;
	if(cudaMalloc((void **)&d_blk, 256) != cudaSuccess) return -1;
	if(cudaMalloc((void **)&d_aux, 8) == cudaSuccess
			&& cudaMemcpy(d_blk, data, 255, cudaMemcpyHostToDevice) == cudaSuccess
			&& cudaMemcpy(d_aux, h, 8, cudaMemcpyHostToDevice) == cudaSuccess
			&& vdl2gpu_launch_rs_verify(d_blk, d_aux, 1, d_aux + 1, NULL) == VDL2GPU_OK
			&& cudaMemcpy(h, d_aux, 8, cudaMemcpyDeviceToHost) == cudaSuccess
			&& cudaMemcpy(data, d_blk, 255, cudaMemcpyDeviceToHost) == cudaSuccess)
		ret = h[1];
	cudaFree(d_blk); cudaFree(d_aux);
	return ret;
}

static int ensure_ctx(int fmt, uint32_t len) {
	if(D.ctx != NULL) return D.fmt == fmt ? 0 : VDL2GPU_EINVAL;
	vdl2gpu_config cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.sample_rate = D.rate; cfg.oversample = D.oversample; cfg.sample_fmt = (uint32_t)fmt;
	cfg.centerfreq = D.centerfreq; cfg.n_channels = D.n_channels; cfg.freqs = D.freq;
	cfg.max_ppm = D.max_ppm;
#ifdef VDL2_DROPIN_USE_REFERENCE_HEADERS
	if(&Config != NULL) cfg.max_ppm = Config.max_ppm;
#endif
	cfg.max_chunk_bytes = len > (1u << 20) ? len : (1u << 20);
	cfg.device = -1;
	cfg.n_inflight = 2;
	D.fmt = fmt;
	return vdl2gpu_create(&cfg, &D.ctx);
}

static void feed(int fmt, unsigned char *buf, uint32_t len) {
	if(len == 0) return;                                       /* src/demod.c:341,358 */
	if(&demods_ready != NULL) pthread_barrier_wait(&demods_ready);
	int rc = ensure_ctx(fmt, len);
	if(rc == 0) rc = vdl2gpu_submit(D.ctx, buf, len);
	if(rc != 0 && D.status == 0) {
		D.status = rc;
		fprintf(stderr, "libvdl2gpu drop-in: %s (%s)\n", vdl2gpu_strerror(rc), vdl2gpu_last_error());
	}
	if(&samples_ready != NULL) pthread_barrier_wait(&samples_ready);
}

extern "C" void process_buf_uchar(unsigned char *buf, uint32_t len, void *ctx) { (void)ctx; feed(VDL2GPU_FMT_U8, buf, len); }
extern "C" void process_buf_short(unsigned char *buf, uint32_t len, void *ctx) { (void)ctx; feed(VDL2GPU_FMT_S16_LE, buf, len); }

/* frame -> avlc_decoder_queue_push, ownership rules of src/decode.c:173-194 (consumer frees all three) */
static void push_frame(const vdl2gpu_frame *f, void *user) {
	(void)user;
	if(avlc_decoder_queue_push == NULL) return;
	vdl2_msg_metadata *m = (vdl2_msg_metadata *)calloc(1, sizeof(*m));
	octet_string_t *o = (octet_string_t *)calloc(1, sizeof(*o));
	uint8_t *copy = (uint8_t *)calloc(f->len ? f->len : 1, 1);
	if(!m || !o || !copy) { free(m); free(o); free(copy); return; }
	m->version = 1;
	m->station_id = D.station_id;
#ifdef VDL2_DROPIN_USE_REFERENCE_HEADERS
	if(&Config != NULL) m->station_id = Config.station_id;
#endif
	m->freq = f->freq;
	m->frame_pwr_dbfs = f->frame_pwr_dbfs;
	m->nf_pwr_dbfs = f->nf_pwr_dbfs;
	m->ppm_error = f->ppm_error;
	m->burst_timestamp = f->burst_timestamp;
	m->datalen_octets = f->datalen_octets;
	m->synd_weight = f->synd_weight;
	m->num_fec_corrections = f->num_fec_corrections;
	m->idx = f->idx;
	memcpy(copy, f->data, f->len);
	o->buf = copy; o->len = f->len;
	avlc_decoder_queue_push(m, o, 0);
}

extern "C" void *process_samples(void *arg) {
	/* src/demod.c:288-337: one thread per channel.  Only channel 0's thread has work to do. */
	int is_first = (D.n_channels > 0 && arg == (void *)D.chan[0]);
	if(&demods_ready == NULL || &samples_ready == NULL) return NULL;
	for(;;) {
		pthread_barrier_wait(&demods_ready);
		pthread_barrier_wait(&samples_ready);
		if(is_first && D.ctx != NULL) {
			int rc = vdl2gpu_flush(D.ctx, push_frame, NULL);
			if(rc < 0 && D.status == 0) D.status = rc;
		}
	}
	return NULL;
}

extern "C" void vdl2gpu_dropin_set_station_id(char *station_id) { D.station_id = station_id; }
extern "C" void vdl2gpu_dropin_set_max_ppm(float max_ppm) { D.max_ppm = max_ppm; }
extern "C" int vdl2gpu_dropin_last_status(void) { return D.status; }
extern "C" void vdl2gpu_dropin_reset(void) {
	if(D.ctx) vdl2gpu_destroy(D.ctx);
	D.ctx = NULL; D.fmt = -1; D.status = 0;
	for(uint32_t i = 0; i < D.n_channels; i++) { free(D.chan[i]); D.chan[i] = NULL; }
	D.n_channels = 0;
}
