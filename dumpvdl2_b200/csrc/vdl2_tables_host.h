/*
 * vdl2_tables_host.h — start-up tables of libvdl2gpu.so, computed on the host by restating the reference's
 * init code read strictly (no fast-math, no contraction).  Pure C++ (no CUDA) so that tests/hostsim can
 * include it too.  Citations are file:line under /root/reference.
 */
#ifndef VDL2_TABLES_HOST_H
#define VDL2_TABLES_HOST_H
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "vdl2_types.h"

/* ------------------------------------------------------------------------------------------------
 * start-up tables
 * ---------------------------------------------------------------------------------------------- */
struct host_tables {
	float sin_lut[257], cos_lut[257];
	vdl2_tables t;
	uint32_t s27;
};

/* src/demod.c:349-354 */
static inline void make_levels(float *levels) {
	for(int code = 0; code < 256; code++) levels[code] = ((float)code - 127.5f) / 127.5f;
}

/* src/demod.c:372-377 and the interpolation of src/demod.c:58-72 re-expressed per table slot:
 * value = v1 + (v2 - v1) * (frac / 65536) == v1 + ((v2 - v1) * 2^-16) * frac, the scaling by a power of
 * two being exact. */
static inline void make_nco_lut(host_tables &h) {
	for(uint32_t i = 0; i < 256; i++) {
		float arg = (float)(2.0f * M_PI * (float)i / 256.0f);
		sincosf(arg, &h.sin_lut[i], &h.cos_lut[i]);
	}
	h.sin_lut[256] = h.sin_lut[0];
	h.cos_lut[256] = h.cos_lut[0];
	for(int i = 0; i < 256; i++) {
		h.t.lut[i][0] = h.cos_lut[i];
		h.t.lut[i][1] = h.sin_lut[i];
		h.t.lut[i][2] = ldexpf(h.cos_lut[i + 1] - h.cos_lut[i], -16);
		h.t.lut[i][3] = ldexpf(h.sin_lut[i + 1] - h.sin_lut[i], -16);
	}
	h.t.lut[256][0] = h.cos_lut[256]; h.t.lut[256][1] = h.sin_lut[256]; h.t.lut[256][2] = 0.f; h.t.lut[256][3] = 0.f;
}

/* src/demod.c:367-370 -> src/chebyshev.c:32-119 for the hard-wired two poles, 8 kHz cut-off, 0.5 % ripple
 * (Smith, DSP Guide ch. 20).  With one biquad section the cascade of chebyshev.c:94-105 is the identity,
 * and the gain step of :107-118 divides the feed-forward taps by sum(A)/(1-sum(B)). */
static inline void make_lpf(uint32_t rate, float *A, float *B) {
	const float fc = (float)8000 / (float)rate, ripple_pct = 0.5f;
	const int np = 2;
	float im_p, re_p;
	sincosf((float)(M_PI / (2 * np) + (1 - 1) * M_PI / np), &im_p, &re_p);
	re_p = -re_p;
	float es = sqrtf(powf(100.f / (100.f - ripple_pct), 2.f) - 1.f);
	float vx = (1.f / np) * logf((1.f / es) + sqrtf(1.f / (es * es) + 1.f));
	float kx = (1.f / np) * logf((1.f / es) + sqrtf(1.f / (es * es) - 1.f));
	kx = (expf(kx) + expf(-kx)) / 2.f;
	re_p *= ((expf(vx) - expf(-vx)) / 2.f) / kx;
	im_p *= ((expf(vx) + expf(-vx)) / 2.f) / kx;
	float t = 2.f * tanf(0.5f);
	float w = (float)(2.f * M_PI * fc);
	float m = re_p * re_p + im_p * im_p;
	float d = 4.f - 4.f * re_p * t + m * t * t;
	float x0 = t * t / d, x1 = 2.f * x0, x2 = x0;
	float y1 = (8.f - 2.f * m * t * t) / d;
	float y2 = (-4.f - 4.f * re_p * t - m * t * t) / d;
	float k = sinf(0.5f - w / 2.f) / sinf(0.5f + w / 2.f);
	d = 1 + y1 * k - y2 * k * k;
	float f0 = (x0 - x1 * k + x2 * k * k) / d;
	float f1 = (-2.f * x0 * k + x1 + x1 * k * k - 2.f * x2 * k) / d;
	float f2 = (x0 * k * k - x1 * k + x2) / d;
	float g1 = (2.f * k + y1 + y1 * k * k - 2.f * y2 * k) / d;
	float g2 = (-(k * k) - y1 * k + y2) / d;
	float sa = 0.f, sb = 0.f;
	sa += f0; sa += f1; sa += f2;
	sb += -0.f; sb += g1; sb += g2;
	float gain = sa / (1.f - sb);
	A[0] = f0 / gain; A[1] = f1 / gain; A[2] = f2 / gain;
	B[0] = -0.f; B[1] = g1; B[2] = g2;
}

/* src/demod.c:84-96, 107-124 */
static inline void make_sync_consts(vdl2_tables &t) {
	static const int steps[16] = { 0, 3, -3, 1, 1, 2, 0, 4, -3, 4, -2, 3, 1, -2, -3, 0 };
	float mean_x = 0.f;
	for(int i = 0; i < 16; i++) mean_x += i;
	mean_x /= 16;
	t.lr_denom = 0.f;
	for(int i = 0; i < 16; i++) {
		t.lr_X[i] = i - mean_x;
		t.lr_denom += (i - mean_x) * (i - mean_x);
		t.pr_phase[i] = (float)(steps[i] * M_PI / 4);
	}
}

/* src/bitstream.c:94-107: scrambler output sequence from the fixed IV (src/decode.c:50), MSB-first words */
static inline void make_lfsr(host_tables &h) {
	memset(h.t.lfsr_words, 0, sizeof(h.t.lfsr_words));
	uint16_t s = VDL2_LFSR_IV;
	for(uint32_t i = 0; i < 1056u * 32u; i++) {
		uint32_t bit = (s ^ (s >> 14)) & 1u;
		s = (uint16_t)((s >> 1) | (bit << 14));
		h.t.lfsr_words[i >> 5] |= bit << (31u - (i & 31u));
	}
	h.s27 = h.t.lfsr_words[0] >> 5;
}

/* src/libfec/init_rs.h:48-58 with gfpoly 0x187 (src/rs.c:28) */
static inline void make_gf(vdl2_tables &t) {
	int v = 1;
	memset(t.gf_log, 0, sizeof(t.gf_log));
	for(int e = 0; e < 255; e++) {
		t.gf_exp[e] = (uint8_t)v;
		t.gf_log[v] = (uint8_t)e;
		v <<= 1;
		if(v & 0x100) v ^= 0x187;
	}
	for(int e = 255; e < 512; e++) t.gf_exp[e] = t.gf_exp[e - 255];
}

/* src/demod.c:137-141: `unwrap -= 2.0f * M_PI` / `unwrap += 2.0f * M_PI` is double arithmetic narrowed on store.
 * Breadth-first enumeration of the values reachable within the 15 steps one got_sync evaluation can take; the
 * transition table replaces the arithmetic in the K2 walk.  Returns the number of states (77), 0 on overflow. */
static inline int make_unwrap_lut(vdl2_tables &t) {
	float vals[VDL2_UNWRAP_STATES];
	int depth[VDL2_UNWRAP_STATES], next[VDL2_UNWRAP_STATES][3];
	int n = 1;
	vals[0] = 0.f; depth[0] = 0;
	for(int head = 0; head < n; head++) {
		next[head][0] = head;
		for(int j = 1; j <= 2; j++) {
			next[head][j] = 0;                              /* transitions out of a depth-15 state are never taken */
			if(depth[head] >= 15) continue;
			const float nv = (float)((double)vals[head] + (j == 1 ? -1.0 : 1.0) * (2.0f * M_PI));
			int k = 0;
			while(k < n && memcmp(&vals[k], &nv, 4) != 0) k++;
			if(k == n) {
				if(n == VDL2_UNWRAP_STATES) return 0;
				vals[n] = nv; depth[n] = depth[head] + 1; n++;
			}
			next[head][j] = k;
		}
	}
	memset(t.unwrap_lut, 0, sizeof(t.unwrap_lut));
	for(int s = 0; s < n; s++)
		for(int j = 0; j < 3; j++) {
			t.unwrap_lut[s * 6 + j * 2] = (uint32_t)next[s][j] * 24u;
			memcpy(&t.unwrap_lut[s * 6 + j * 2 + 1], &vals[next[s][j]], 4);
		}
	return n;
}

static inline void make_tables(host_tables &h, uint32_t rate) {
	memset(&h, 0, sizeof(h));
	make_levels(h.t.levels);
	make_nco_lut(h);
	make_lpf(rate, h.t.A, h.t.B);
	make_sync_consts(h.t);
	make_lfsr(h);
	make_gf(h.t);
	make_unwrap_lut(h.t);
}


#endif
