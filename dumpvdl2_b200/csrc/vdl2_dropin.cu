/*
 * vdl2_dropin.cu — drop-in shim: the reference's own entry points for this path (include/vdl2_dropin.h) on top
 * of the batch API.  Host code only; written in the C subset of C++ so that it reads like the reference's C.
 *
 * Thread protocol kept from the reference (src/demod.c:299-301,339-347; src/dumpvdl2.c:117-135,1170):
 *   producer  process_buf_*():  wait(demods_ready); hand the buffer over; wait(samples_ready)
 *   consumers process_samples(): wait(demods_ready); wait(samples_ready); work
 * Here "hand the buffer over" = vdl2gpu_submit() (copy into the pinned ring, async H2D + kernels) and the
 * "work" of channel 0's thread = pushing finished frames through avlc_decoder_queue_push().  The reference's
 * end-of-stream drain is main() waiting once more on demods_ready (src/dumpvdl2.c:1170): a consumer that arrives
 * at that barrier must have finished ALL its work.  Channel 0's thread therefore arrives at demods_ready only
 * when either (a) every submitted buffer has been processed and its frames pushed, or (b) the producer is already
 * waiting there with the next buffer (feed() raises `producer_waiting` before its barrier wait; main()'s final
 * wait never does).  (b) lets buffer n+1 enter the GPU pipeline while buffer n is still in flight (K1 of n+1 beside
 * K2/K3 of n); (a) makes the final barrier the same complete drain it is with the CPU demodulators.
 * The other channel threads only keep the barrier counts right.
 */
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <atomic>
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
	uint32_t max_chunk_bytes;
	int nonfatal;                 /* VDL2GPU_DROPIN_NONFATAL=1: record failures in `status` instead of terminating (tests) */
} D = { {0}, {0}, 0, 0, 0, 0, NULL, -1, 0, NULL, 0.f, PTHREAD_MUTEX_INITIALIZER, 0, 0 };

static std::atomic<int> g_producer_waiting(0);

/* Like the reference (xcalloc -> _exit(1), src/util.c:32-40; init failures -> _exit(2/3), src/dumpvdl2.c:1090-1099)
 * the drop-in has no error channel towards its callers: a demodulator that cannot run must not keep the program
 * alive without output. */
static void shim_fail(int rc, const char *what) {
	if(D.status == 0) D.status = rc;
	fprintf(stderr, "libvdl2gpu drop-in: %s: %s (%s)\n", what, vdl2gpu_strerror(rc), vdl2gpu_last_error());
	if(!D.nonfatal) _exit(3);
}

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

/* src/rs.c:32-49 for callers outside the burst kernel: one block through the device RS routine.  Scratch
 * (pinned host + device, 264 bytes each) is allocated once; a call is two async copies and one kernel on the
 * legacy stream. */
static struct { uint8_t *h; uint8_t *d; int dev; pthread_mutex_t lock; } RSV = { NULL, NULL, -1, PTHREAD_MUTEX_INITIALIZER };

extern "C" int rs_verify(uint8_t *data, int fec_octets) {
	int ret = -1, dev = 0;
	pthread_mutex_lock(&RSV.lock);
	if(cudaGetDevice(&dev) != cudaSuccess) { pthread_mutex_unlock(&RSV.lock); return -1; }
	if(RSV.h == NULL || RSV.dev != dev) {
		if(RSV.h) { cudaFreeHost(RSV.h); cudaFree(RSV.d); RSV.h = NULL; RSV.d = NULL; }
		if(cudaHostAlloc((void **)&RSV.h, 264, cudaHostAllocDefault) != cudaSuccess || cudaMalloc((void **)&RSV.d, 264) != cudaSuccess) {
			if(RSV.h) cudaFreeHost(RSV.h);
			RSV.h = NULL; RSV.d = NULL;
			pthread_mutex_unlock(&RSV.lock);
			return -1;
		}
		RSV.dev = dev;
	}
	/* layout: [0,255) block, [256,260) fec_octets, [260,264) return value */
	memcpy(RSV.h, data, 255);
	int32_t fo = fec_octets, rv = -1;
	memcpy(RSV.h + 256, &fo, 4); memcpy(RSV.h + 260, &rv, 4);
	if(cudaMemcpyAsync(RSV.d, RSV.h, 264, cudaMemcpyHostToDevice, 0) == cudaSuccess
			&& vdl2gpu_launch_rs_verify(RSV.d, (const int32_t *)(RSV.d + 256), 1, (int32_t *)(RSV.d + 260), NULL) == VDL2GPU_OK
			&& cudaMemcpyAsync(RSV.h, RSV.d, 264, cudaMemcpyDeviceToHost, 0) == cudaSuccess
			&& cudaStreamSynchronize(0) == cudaSuccess) {
		memcpy(&rv, RSV.h + 260, 4);
		memcpy(data, RSV.h, 255);
		ret = rv;
	}
	pthread_mutex_unlock(&RSV.lock);
	return ret;
}

/* The context is sized for the largest buffer any front-end of the reference hands over (SoapySDR / SDRplay:
 * 524288 int16 = 1 MiB, src/soapysdr.h:21; RTL / file: 320000 bytes, src/rtl.h:21, src/dumpvdl2.h:48) or the first
 * buffer seen, whichever is larger; a later, even larger buffer is fed in pieces (feed()). */
static int ensure_ctx(int fmt, uint32_t len) {
	if(D.ctx != NULL) return D.fmt == fmt ? 0 : VDL2GPU_EINVAL;
	const char *nf = getenv("VDL2GPU_DROPIN_NONFATAL");
	D.nonfatal = (nf && atoi(nf)) ? 1 : 0;
	vdl2gpu_config cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.sample_rate = D.rate; cfg.oversample = D.oversample; cfg.sample_fmt = (uint32_t)fmt;
	cfg.centerfreq = D.centerfreq; cfg.n_channels = D.n_channels; cfg.freqs = D.freq;
	cfg.max_ppm = D.max_ppm;
#ifdef VDL2_DROPIN_USE_REFERENCE_HEADERS
	if(&Config != NULL) cfg.max_ppm = Config.max_ppm;
#endif
	cfg.max_chunk_bytes = len > (1u << 20) ? len : (1u << 20);
	cfg.max_chunk_bytes &= ~3u;                                 /* whole samples in either format */
	cfg.device = -1;
	cfg.n_inflight = 3;
	D.fmt = fmt;
	D.max_chunk_bytes = cfg.max_chunk_bytes;
	return vdl2gpu_create(&cfg, &D.ctx);
}

static void feed(int fmt, unsigned char *buf, uint32_t len) {
	if(len == 0) return;                                       /* src/demod.c:341,358 */
	g_producer_waiting.store(1, std::memory_order_release);    /* see process_samples */
	if(&demods_ready != NULL) pthread_barrier_wait(&demods_ready);
	int rc = ensure_ctx(fmt, len);
	if(rc != 0) shim_fail(rc, "vdl2gpu_create");
	/* a buffer larger than the context was sized for goes in as several chunks: chunking does not change the result */
	for(uint32_t off = 0; rc == 0 && off < len; off += D.max_chunk_bytes) {
		const uint32_t n = len - off < D.max_chunk_bytes ? len - off : D.max_chunk_bytes;
		rc = vdl2gpu_submit(D.ctx, buf + off, n);
		if(rc != 0) shim_fail(rc, "vdl2gpu_submit");
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
		g_producer_waiting.store(0, std::memory_order_release);    /* whoever was waiting has been let through */
		pthread_barrier_wait(&samples_ready);
		if(is_first && D.ctx != NULL) {
			/* push what has finished; go back to the barrier once everything has (final drain) or as soon as the
			 * producer stands there with the next buffer (the GPU then works on two buffers at once) */
			for(;;) {
				int rc = vdl2gpu_poll(D.ctx, push_frame, NULL);
				if(rc == VDL2GPU_EOVERFLOW) {                  /* bursts were dropped on the device: report, keep running */
					if(D.status == 0) D.status = rc;
					fprintf(stderr, "libvdl2gpu drop-in: %s\n", vdl2gpu_last_error());
				} else if(rc < 0) { shim_fail(rc, "vdl2gpu_poll"); break; }
				if(vdl2gpu_chunks_in_flight(D.ctx) == 0) break;
				if(g_producer_waiting.load(std::memory_order_acquire)) break;
				usleep(200);
			}
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
