/*
 * include/vdl2_dropin.h — the drop-in face of libvdl2gpu.so: the exact symbols the rest of dumpvdl2 expects
 * from src/demod.c, src/rs.c (declared in reference src/dumpvdl2.h:371-389) so that the unmodified SDR
 * front-ends (src/rtl.c:194-196, src/mirics.c:203-204, src/sdrplay.c:132, src/sdrplay3.c:110,
 * src/soapysdr.c:220-223, process_iq_file src/dumpvdl2.c:323-358) and main() (src/dumpvdl2.c:1086-1170) link and
 * run unchanged, while every sample is demodulated on the GPU.
 *
 * A maintainer builds the reference with src/demod.c, src/chebyshev.c, src/rs.c and src/libfec removed from
 * the source list and libvdl2gpu.so added to the link line.  src/decode.c stays: it owns
 * avlc_decoder_queue_push() and the AVLC decoder thread, which this library calls / feeds.
 *
 * Symbols resolved from the host program (weak here, so the library also loads stand-alone):
 *   pthread_barrier_t demods_ready, samples_ready;           src/dumpvdl2.c:67
 *   void avlc_decoder_queue_push(vdl2_msg_metadata *, octet_string_t *, int);   src/decode.c:165-171
 *   dumpvdl2_config_t Config;   (only when built with -DVDL2_DROPIN_USE_REFERENCE_HEADERS, for station_id / max_ppm)
 *
 * Build modes:
 *   default                               : the layout mirrors below are used (checked against the reference's
 *                                           headers by tests/test_dropin.py when /root/reference is mounted);
 *   -DVDL2_DROPIN_USE_REFERENCE_HEADERS   : "dumpvdl2.h", "output-common.h" of the reference are included instead.
 */
#ifndef VDL2_DROPIN_H
#define VDL2_DROPIN_H
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <sys/time.h>

#ifdef VDL2_DROPIN_USE_REFERENCE_HEADERS
#ifdef __cplusplus
extern "C" {                           /* the reference's headers are plain C */
#endif
#include "dumpvdl2.h"
#include "output-common.h"
#ifdef __cplusplus
}
#endif
#else
/* ---- layout mirrors (interface declarations; field order and types as in the reference) ---- */
typedef struct {                       /* src/dumpvdl2.h:290-293 */
	uint8_t *buf;
	uint32_t start, end, len, descrambler_pos;
} bitstream_t;
typedef struct {                       /* src/dumpvdl2.h:321-352; main() touches only demod_thread */
	long long unsigned samplenum;
	bitstream_t *bs, *frame_bs;
	float syncbuf[160];
	float prev_phi;
	float prev_dphi, dphi;
	float pherr[3];
	float ppm_error;
	float mag_lp;
	float mag_nf;
	float frame_pwr;
	int bufnum;
	int nfcnt;
	int syncbufidx;
	int frame_pwr_cnt;
	int sclk;
	int offset_tuning;
	int num_fec_corrections;
	int demod_state;               /* enum demod_states */
	int decoder_state;             /* enum decoder_states */
	uint32_t freq;
	uint32_t downmix_phi, downmix_dphi;
	uint32_t requested_bits;
	uint32_t datalen, datalen_octets, last_block_len_octets, fec_octets;
	uint32_t num_blocks;
	uint32_t syndrome;
	uint16_t lfsr;
	uint16_t oversample;
	struct timeval tstart;
	struct timeval burst_timestamp;
	pthread_t demod_thread;
} vdl2_channel_t;
typedef struct {                       /* src/dumpvdl2.h:422-425 */
	uint8_t *buf;
	size_t len;
} octet_string_t;
typedef struct {                       /* src/output-common.h:31-43 */
	char *station_id;
	uint32_t freq;
	uint32_t synd_weight;
	uint32_t datalen_octets;
	float frame_pwr_dbfs;
	float nf_pwr_dbfs;
	float ppm_error;
	int version;
	int num_fec_corrections;
	int idx;
	struct timeval burst_timestamp;
} vdl2_msg_metadata;
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- src/demod.c replacements (src/dumpvdl2.h:371-381) ---- */
extern float *sbuf;                    /* front-ends assign it; the GPU path ignores it */
vdl2_channel_t *vdl2_channel_init(uint32_t centerfreq, uint32_t freq, uint32_t source_rate, uint32_t oversample);
void sincosf_lut_init(void);
void input_lpf_init(uint32_t sample_rate);
void demod_sync_init(void);
void process_buf_uchar_init(void);
void process_buf_uchar(unsigned char *buf, uint32_t len, void *ctx);
void process_buf_short(unsigned char *buf, uint32_t len, void *ctx);
void *process_samples(void *arg);      /* pthread entry, one per channel: keeps the two-barrier protocol alive */

/* ---- src/rs.c replacements (src/dumpvdl2.h:387-389); rs_verify runs the K3 RS routine on the device ---- */
int rs_init(void);
int rs_verify(uint8_t *data, int fec_octets);

/* ---- extras (not in the reference) ---- */
void vdl2gpu_dropin_set_station_id(char *station_id);   /* only needed in the mirror build mode */
void vdl2gpu_dropin_set_max_ppm(float max_ppm);         /* Config.max_ppm, mirror build mode */
int vdl2gpu_dropin_last_status(void);                   /* last VDL2GPU_E* code seen by the shim (0 = fine) */
void vdl2gpu_dropin_reset(void);                        /* tear down the GPU context and forget the channels (tests) */

#ifdef __cplusplus
}
#endif
#endif
