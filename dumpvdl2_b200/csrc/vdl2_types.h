/*
 * vdl2_types.h — data layout shared by the kernels and the host runtime of libvdl2gpu.so.
 * Reference citations are file:line under /root/reference.
 *
 * HBM layout per context (one IQ stream, n_ch channels, n_chp = n_ch rounded up to 32):
 *   samples   float2[max_pairs]            K0 -> K1   converted stream {re, im} (src/demod.c:339-365); one stream per
 *                                                     channel: float2[max_pairs][n_chp], time-major
 *   dec       float2[max_dec][n_chp]       K1 -> K2   decimated samples, TIME-MAJOR: a warp of 32 channels
 *                                                     reads/writes 256 contiguous bytes per time step
 *   phase     float[160+max_dec][n_chp]    K2a -> K2  atan2 of every decimated sample; the first 160 rows carry the
 *                                                     tail of the previous chunk (the preamble metric looks 150 back)
 *   mag       float[max_dec][n_chp]        K2a -> K2  hypot of every decimated sample
 *   k1 state  u32[K1_NFIELDS][n_chp]       SoA        filter delay lines + NCO (src/demod.c:289-298)
 *   k2 state  u32[K2_NFIELDS][n_chp]       SoA        demodulator/decoder scalars (src/dumpvdl2.h:321-352)
 *   ring      float[160][n_chp]                       syncbuf phase ring (src/dumpvdl2.h:324)
 *   pool      BurstSlot[n_slots]           K2 -> K3   packed burst bits + metadata snapshot
 *   out       mapped pinned host memory    K3 -> host burst records (frames + metadata)
 */
#ifndef VDL2_TYPES_H
#define VDL2_TYPES_H
#include <stdint.h>

#define VDL2_RS_K 249
#define VDL2_RS_N 255
#define VDL2_TRLEN 17
#define VDL2_HDRFECLEN 5
#define VDL2_HEADER_LEN 25
#define VDL2_PREAMBLE_SYMS 16
#define VDL2_SPS 10
#define VDL2_BPS 3
#define VDL2_SYNC_BUFLEN 160
#define VDL2_SYNC_SKIP 3
#define VDL2_SYMBOL_RATE 10500
#define VDL2_MAX_FRAME_LENGTH 0x3FFFu            /* src/decode.c:45 */
#define VDL2_MAX_FRAME_LENGTH_CORRECTED 0x1FFFu  /* src/decode.c:48 */
#define VDL2_LFSR_IV 0x6959u                     /* src/decode.c:50 */
#define VDL2_MAX_BLOCKS 9                        /* ceil(2048 / 249) */
/* 25 header bits + 8 * (2048 data + 54 fec octets) + up to 2 spare bits of the last symbol */
#define VDL2_MAX_BURST_BITS (25 + 8 * (2048 + 54) + 2)
#define VDL2_MAX_BURST_WORDS ((VDL2_MAX_BURST_BITS + 31) / 32)     /* 527 */
#define VDL2_UNWRAP_STATES 80
#define VDL2_MAX_FRAMES 1032                     /* >= 16384 bits / 16 */

/* burst decoder status; names follow the statsd counters of src/decode.c:204-369 */
enum {
	VDL2_BURST_OK = 0,
	VDL2_ERR_NO_HEADER = 1, VDL2_ERR_CRC_BAD = 2, VDL2_ERR_TOO_LONG = 3, VDL2_ERR_NO_FEC = 4,
	VDL2_ERR_DATA_TRUNCATED = 5, VDL2_ERR_FEC_TRUNCATED = 6, VDL2_ERR_DEINTERLEAVE_DATA = 7,
	VDL2_ERR_DEINTERLEAVE_FEC = 8, VDL2_ERR_FEC_BAD = 9, VDL2_ERR_BITSTREAM = 10,
	VDL2_ERR_TRUNCATED_OCTETS = 11, VDL2_ERR_UNSTUFF = 12
};

/* K1 per-channel state fields (SoA, one u32 plane per field) */
enum {
	K1_XR1 = 0, K1_XR2, K1_XI1, K1_XI2, K1_YR1, K1_YR2, K1_YI1, K1_YI2, K1_PHI, K1_DPHI, K1_NFIELDS
};

/* K2 per-channel state fields */
enum {
	K2_PREV_PHI = 0, K2_PREV_DPHI, K2_DPHI, K2_PHERR0, K2_PHERR1, K2_PHERR2, K2_PPM, K2_MAG_LP, K2_MAG_NF,
	K2_FRAME_PWR, K2_RING_POS, K2_SCLK, K2_NFCNT, K2_FRAME_PWR_CNT, K2_STATE, K2_ACC_LO, K2_ACC_HI, K2_NBITS,
	K2_NEED_BITS, K2_DATALEN, K2_SYNDROME, K2_SLOT, K2_BURST_SEQ, K2_SYNC_LO, K2_SYNC_HI, K2_FREQ,
	K2_CNT_SYNC, K2_CNT_HDR_GOOD, K2_PURE_RUN, K2_NFIELDS
};
/* K2_STATE bit layout */
#define VDL2_ST_LOCKED 1u            /* demod_state == DM_SYNC (src/dumpvdl2.h:294) */
#define VDL2_DEC_SHIFT 1
#define VDL2_DEC_HEADER 0u           /* src/dumpvdl2.h:295 */
#define VDL2_DEC_DATA 1u
#define VDL2_DEC_IDLE 2u

/* per-channel counters kept on the device (u32 planes, SoA) */
enum {
	VDL2_CNT_SYNC_GOOD = 0, VDL2_CNT_HDR_CRC_GOOD, VDL2_CNT_BURSTS, VDL2_CNT_BURST_ERR, VDL2_CNT_BLOCKS_PROCESSED,
	VDL2_CNT_BLOCKS_FEC_OK, VDL2_CNT_MSG_GOOD, VDL2_CNT_FCS_GOOD, VDL2_CNT_FCS_BAD, VDL2_NUM_COUNTERS
};

/* burst hand-off K2 -> K3 */
typedef struct {
	uint32_t channel;
	uint32_t burst_seq;
	uint32_t datalen_bits;       /* transmission length from the header (src/decode.c:222) */
	uint32_t syndrome;           /* header syndrome (src/decode.c:210) */
	uint32_t nbits;              /* bits stored in words[], header included */
	float frame_pwr;             /* snapshots taken when the last symbol arrived (src/decode.c:180-182) */
	float mag_nf;
	float ppm_error;
	uint32_t sync_lo, sync_hi;   /* decimated-sample index of the sync decision */
	uint32_t freq;
	uint32_t pad;
	uint32_t words[VDL2_MAX_BURST_WORDS + 1];   /* bit i of the burst = words[i/32] >> (31 - i%32) & 1, still scrambled */
} vdl2_burst_slot;

/* device-side queue bookkeeping */
typedef struct {
	int32_t free_top;            /* number of entries in free_list */
	uint32_t n_ready;            /* entries in ready[] for the chunk being processed */
	uint32_t pool_overflows;
	uint32_t out_overflows;
	uint32_t out_used;           /* bytes used in the current output region */
	uint32_t out_records;
	uint32_t n_events;
	uint32_t pad;
} vdl2_queue_ctl;

/* burst record, K3 -> host (in mapped pinned memory), followed by
 *   n_frames x { uint16 len; uint16 fcs_residue; }   then the frame octets back to back, padded to 16 bytes */
typedef struct {
	uint32_t rec_bytes;          /* whole record, multiple of 16 */
	uint32_t channel;
	uint32_t burst_seq;
	int32_t status;              /* VDL2_BURST_OK / VDL2_ERR_* */
	uint32_t n_frames;
	uint32_t datalen_bits;
	uint32_t syndrome;
	int32_t num_fec_corrections;
	float frame_pwr, mag_nf, ppm_error;
	uint32_t num_blocks;
	uint32_t sync_lo, sync_hi;
	uint32_t freq;
	uint32_t frame_bytes;        /* total octets of all frames */
	int8_t rs_ret[12];           /* per block: decode_rs_char return value, -128 = not run */
	uint32_t pad;
} vdl2_burst_record;             /* 80 bytes */

/* header of each output region */
typedef struct {
	uint32_t bytes_used;
	uint32_t n_records;
	uint32_t pool_overflows;
	uint32_t out_overflows;
	uint32_t n_events_total;
	uint32_t pad[3];
} vdl2_out_header;               /* 32 bytes */

/* read-only tables resident in HBM */
typedef struct {
	float levels[256];           /* src/demod.c:349-354 */
	float lut[257][4];           /* {cos, sin, dcos*2^-16, dsin*2^-16} per 1/256 turn (src/demod.c:58-72,372-377) */
	float A[3], B[3];            /* src/demod.c:367-370 */
	float lr_X[16];              /* src/demod.c:84-96 */
	float lr_denom;
	float pr_phase[16];          /* src/demod.c:107-124 */
	uint32_t lfsr_words[1056];   /* scrambler output, bit i at words[i/32] >> (31 - i%32), 33792 bits (period 32767 wraps) */
	uint8_t gf_exp[512];         /* src/libfec/init_rs.h:48-58, doubled to skip the modulo */
	uint8_t gf_log[256];
	/* got_sync's `unwrap` accumulator (src/demod.c:137-141) as a finite automaton: it only takes the values reachable
	 * from 0 by at most 15 steps of fl32((double)u -/+ 2pi) (77 of them).  Row r (24 bytes) = three {next row byte
	 * offset, value bits of the next state} pairs: +0 no wrap, +8 step > pi (u -= 2pi), +16 step < -pi (u += 2pi). */
	uint32_t unwrap_lut[VDL2_UNWRAP_STATES * 6];
} vdl2_tables;

#endif
