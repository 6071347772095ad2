/* oracle/layout_check.c — TEST INFRASTRUCTURE: compares the layout mirrors of include/vdl2_dropin.h with the
 * reference's own declarations (src/dumpvdl2.h:321-352,422-425, src/output-common.h:31-43).  Prints one line per
 * check and exits non-zero on any mismatch. */
#include <stdio.h>
#include <stddef.h>
#include "dumpvdl2.h"
#include "output-common.h"
#define vdl2_channel_t mirror_channel_t
#define bitstream_t mirror_bitstream_t
#define octet_string_t mirror_octet_string_t
#define vdl2_msg_metadata mirror_metadata
#define sbuf mirror_sbuf
#define vdl2_channel_init mirror_channel_init
#define sincosf_lut_init mirror_f1
#define input_lpf_init mirror_f2
#define demod_sync_init mirror_f3
#define process_buf_uchar_init mirror_f4
#define process_buf_uchar mirror_f5
#define process_buf_short mirror_f6
#define process_samples mirror_f7
#define rs_init mirror_f8
#define rs_verify mirror_f9
#include "vdl2_dropin.h"
#undef vdl2_channel_t
#undef bitstream_t
#undef octet_string_t
#undef vdl2_msg_metadata

static int bad;
#define SAME(expr_a, expr_b, what) do { size_t a_ = (expr_a), b_ = (expr_b); \
	printf("%-44s reference %4zu  mirror %4zu  %s\n", what, a_, b_, a_ == b_ ? "ok" : "MISMATCH"); if(a_ != b_) bad = 1; } while(0)

int main(void) {
	SAME(sizeof(vdl2_channel_t), sizeof(mirror_channel_t), "sizeof(vdl2_channel_t)");
	SAME(offsetof(vdl2_channel_t, demod_thread), offsetof(mirror_channel_t, demod_thread), "offsetof(vdl2_channel_t, demod_thread)");
	SAME(offsetof(vdl2_channel_t, freq), offsetof(mirror_channel_t, freq), "offsetof(vdl2_channel_t, freq)");
	SAME(offsetof(vdl2_channel_t, oversample), offsetof(mirror_channel_t, oversample), "offsetof(vdl2_channel_t, oversample)");
	SAME(offsetof(vdl2_channel_t, burst_timestamp), offsetof(mirror_channel_t, burst_timestamp), "offsetof(vdl2_channel_t, burst_timestamp)");
	SAME(sizeof(octet_string_t), sizeof(mirror_octet_string_t), "sizeof(octet_string_t)");
	SAME(offsetof(octet_string_t, len), offsetof(mirror_octet_string_t, len), "offsetof(octet_string_t, len)");
	SAME(sizeof(vdl2_msg_metadata), sizeof(mirror_metadata), "sizeof(vdl2_msg_metadata)");
	SAME(offsetof(vdl2_msg_metadata, idx), offsetof(mirror_metadata, idx), "offsetof(vdl2_msg_metadata, idx)");
	SAME(offsetof(vdl2_msg_metadata, burst_timestamp), offsetof(mirror_metadata, burst_timestamp), "offsetof(vdl2_msg_metadata, burst_timestamp)");
	SAME(offsetof(vdl2_msg_metadata, ppm_error), offsetof(mirror_metadata, ppm_error), "offsetof(vdl2_msg_metadata, ppm_error)");
	return bad;
}
