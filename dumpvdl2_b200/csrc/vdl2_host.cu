/*
 * vdl2_host.cu — host runtime of libvdl2gpu.so: start-up tables, device-resident channel state, the
 * pinned staging ring, stream-ordered kernel chain per IQ chunk, and harvesting of burst records into
 * avlc_decoder_queue_push-shaped frames.  C-ABI in include/vdl2gpu.h.
 *
 * Host arithmetic for the start-up tables restates the reference's init code read strictly
 * (compiled with -ffp-contract=off, no fast-math); citations are file:line under /root/reference.
 * There is no CPU implementation of the sample path in this library: every data-path entry point
 * needs a CUDA device and fails with VDL2GPU_ENODEV / VDL2GPU_ECUDA otherwise.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <algorithm>
#include <deque>
#include <string>
#include <vector>
#include "../../include/vdl2gpu.h"
#include "vdl2_kernels.h"
#include "vdl2_types.h"
#include "vdl2_tables_host.h"

static thread_local char g_last_error[512] = "";

static int fail_cuda(cudaError_t e, const char *what, int line) {
	snprintf(g_last_error, sizeof(g_last_error), "%s failed at vdl2_host.cu:%d: %s", what, line, cudaGetErrorString(e));
	return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? VDL2GPU_ENODEV : VDL2GPU_ECUDA;
}
#define CU(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) return fail_cuda(e_, #call, __LINE__); } while(0)
#define KL(call) do { int e_ = (call); if(e_ != 0) return fail_cuda((cudaError_t)e_, #call, __LINE__); } while(0)

/* ------------------------------------------------------------------------------------------------
 * context
 * ---------------------------------------------------------------------------------------------- */
/* frames harvested but not yet delivered; their octets live in per-chunk blobs (one copy per chunk) */
struct pending_frame {
	vdl2gpu_frame f;
	uint32_t blob, offset;
};

struct chunk_slot {
	uint8_t *h_raw = nullptr, *d_raw = nullptr;
	uint8_t *h_out = nullptr, *d_out = nullptr;
	cudaEvent_t done = nullptr;
	cudaEvent_t tk[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };   /* front: 0,1,2  back: 3,4,5 */
	bool busy = false, timed = false;
	uint64_t first_pair = 0, dec_base = 0;
	uint32_t n_pairs = 0, n_dec = 0;
	struct timeval arrival = { 0, 0 };
};

struct vdl2gpu_ctx {
	vdl2gpu_config cfg;
	std::vector<uint32_t> freqs;
	int device = 0;
	/* `stream` (front) carries H2D + K0 + K1 of chunk c+1 while `s_back` carries K2a/K2/K3 of chunk c.  Both stages
	 * are latency-bound with one warp per SM sub-partition at 16 k channels, so co-residency raises issue-slot use:
	 * 12.1 -> 8.7 ms per chunk on B200.  This only works because every kernel of the chain requests the SAME
	 * shared-memory carve-out (vdl2_kernels.cu): with per-kernel defaults the SMs drain to re-partition L1/shared
	 * memory and the overlap is a 1.7x slow-down.  VDL2GPU_FLAG_NO_OVERLAP puts everything on one stream. */
	cudaStream_t stream = nullptr, s_back = nullptr;
	cudaEvent_t ev_k1_done[2] = { nullptr, nullptr }, ev_back_done[2] = { nullptr, nullptr }, ev_k2a_done = nullptr;
	uint64_t chunk_seq = 0;
	cudaEvent_t ev_input_ready = nullptr, ev_input_consumed = nullptr;
	uint32_t n_ch = 0, n_chp = 0, max_pairs = 0, max_dec = 0, n_slots = 0, out_cap = 0, event_cap = 0;
	host_tables tab;
	vdl2_tables *d_tab = nullptr;
	float4 *d_samples = nullptr;
	float2 *d_dec2[2] = { nullptr, nullptr };
	float *d_phase = nullptr, *d_mag = nullptr, *d_hist_tmp = nullptr;
	uint32_t *d_k1 = nullptr, *d_k2 = nullptr, *d_counters = nullptr, *d_ready = nullptr;
	float *d_ring = nullptr;
	vdl2_burst_slot *d_pool = nullptr;
	int32_t *d_free = nullptr;
	vdl2_queue_ctl *d_ctl = nullptr;
	void *d_events = nullptr;
	std::vector<chunk_slot> chunks;
	std::deque<uint32_t> inflight;
	uint32_t next_slot = 0;
	uint32_t decim_cnt = 0, last_n_dec = 0;
	uint64_t total_pairs = 0, total_dec = 0;
	uint32_t events_read = 0;
	bool timing = false;
	double k_ms[4] = { 0, 0, 0, 0 };
	uint64_t k_launches[4] = { 0, 0, 0, 0 };
	vdl2gpu_stats stats;
	std::vector<pending_frame> pending;
	std::vector<std::vector<uint8_t>> blobs;
	cudaEvent_t ev_drain = nullptr;
};

static uint32_t dphi_for(uint32_t centerfreq, uint32_t freq, uint32_t rate) {      /* src/demod.c:385 */
	return (uint32_t)(int)(((float)centerfreq - (float)freq) / (float)rate * 256.0f * 65536.0f);
}

extern "C" int vdl2gpu_abi_version(void) { return VDL2GPU_ABI_VERSION; }

extern "C" int vdl2gpu_device_count(void) {
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

extern "C" const char *vdl2gpu_last_error(void) { return g_last_error; }

extern "C" const char *vdl2gpu_strerror(int code) {
	switch(code) {
		case VDL2GPU_OK: return "ok";
		case VDL2GPU_EINVAL: return "invalid argument";
		case VDL2GPU_ENODEV: return "no usable CUDA device";
		case VDL2GPU_ECUDA: return "CUDA runtime error";
		case VDL2GPU_ENOMEM: return "out of memory";
		case VDL2GPU_ETOOBIG: return "chunk larger than max_chunk_bytes";
		case VDL2GPU_EOVERFLOW: return "device queue overflow, bursts dropped";
		default: return "unknown error";
	}
}

static int free_ctx(vdl2gpu_ctx *c) {
	if(!c) return VDL2GPU_OK;
	cudaSetDevice(c->device);
	if(c->stream) cudaStreamSynchronize(c->stream);
	if(c->s_back) cudaStreamSynchronize(c->s_back);
	for(auto &s : c->chunks) {
		if(s.h_raw) cudaFreeHost(s.h_raw);
		if(s.d_raw) cudaFree(s.d_raw);
		if(s.h_out) cudaFreeHost(s.h_out);
		if(s.done) cudaEventDestroy(s.done);
		for(auto &e : s.tk) if(e) cudaEventDestroy(e);
	}
	cudaFree(c->d_tab); cudaFree(c->d_samples); cudaFree(c->d_dec2[0]); cudaFree(c->d_dec2[1]); cudaFree(c->d_phase); cudaFree(c->d_mag); cudaFree(c->d_hist_tmp); cudaFree(c->d_k1); cudaFree(c->d_k2);
	cudaFree(c->d_counters); cudaFree(c->d_ready); cudaFree(c->d_ring); cudaFree(c->d_pool); cudaFree(c->d_free);
	cudaFree(c->d_ctl); cudaFree(c->d_events);
	if(c->ev_input_ready) cudaEventDestroy(c->ev_input_ready);
	if(c->ev_input_consumed) cudaEventDestroy(c->ev_input_consumed);
	if(c->ev_drain) cudaEventDestroy(c->ev_drain);
	if(c->ev_k2a_done) cudaEventDestroy(c->ev_k2a_done);
	for(int i = 0; i < 2; i++) { if(c->ev_k1_done[i]) cudaEventDestroy(c->ev_k1_done[i]); if(c->ev_back_done[i]) cudaEventDestroy(c->ev_back_done[i]); }
	if(c->s_back && c->s_back != c->stream) cudaStreamDestroy(c->s_back);
	if(c->stream) cudaStreamDestroy(c->stream);
	delete c;
	return VDL2GPU_OK;
}

static int create_impl(const vdl2gpu_config *cfg, vdl2gpu_ctx *c) {
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if(e != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		snprintf(g_last_error, sizeof(g_last_error), "no CUDA device (%s); libvdl2gpu has no CPU fallback",
				e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
		return VDL2GPU_ENODEV;
	}
	if(cfg->device >= 0) { c->device = cfg->device; CU(cudaSetDevice(c->device)); }
	else CU(cudaGetDevice(&c->device));
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, c->device));
	if(prop.major < 10) {
		snprintf(g_last_error, sizeof(g_last_error), "device %d is sm_%d%d; this library carries sm_100a code only", c->device, prop.major, prop.minor);
		return VDL2GPU_ENODEV;
	}
	c->cfg = *cfg;
	c->freqs.assign(cfg->freqs, cfg->freqs + cfg->n_channels);
	c->cfg.freqs = c->freqs.data();
	c->n_ch = cfg->n_channels;
	c->n_chp = (c->n_ch + 31u) & ~31u;
	const uint32_t max_bytes = cfg->max_chunk_bytes ? cfg->max_chunk_bytes : (1u << 20);
	c->cfg.max_chunk_bytes = max_bytes;
	c->max_pairs = max_bytes / (cfg->sample_fmt == VDL2GPU_FMT_S16_LE ? 4u : 2u);
	c->max_dec = c->max_pairs / cfg->oversample + 2;
	if((uint64_t)(c->max_dec + VDL2_SYNC_BUFLEN) * c->n_chp >= (1ull << 32)) {
		snprintf(g_last_error, sizeof(g_last_error), "%u channels x %u decimated samples per chunk exceed the 32-bit element index of the kernels; "
				"use a smaller max_chunk_bytes or shard the channels", c->n_ch, c->max_dec);
		return VDL2GPU_ETOOBIG;
	}
	c->n_slots = std::max(256u, 3u * c->n_ch);
	c->out_cap = std::max(4u << 20, c->n_ch * 512u);
	const uint32_t n_inflight = cfg->n_inflight ? cfg->n_inflight : 4u;
	c->cfg.n_inflight = n_inflight;
	c->event_cap = (cfg->flags & VDL2GPU_FLAG_TRACE) ? (1u << 20) : 1u;
	make_tables(c->tab, cfg->sample_rate);
	memset(&c->stats, 0, sizeof(c->stats));

	/* equal (default) priorities: raising either stage's priority slowed the pair down (tools/probe_overlap.py) */
	CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	if(cfg->flags & VDL2GPU_FLAG_NO_OVERLAP) c->s_back = c->stream;
	else CU(cudaStreamCreateWithFlags(&c->s_back, cudaStreamNonBlocking));
	CU(cudaEventCreateWithFlags(&c->ev_k2a_done, cudaEventDisableTiming));
	for(int i = 0; i < 2; i++) {
		CU(cudaEventCreateWithFlags(&c->ev_k1_done[i], cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&c->ev_back_done[i], cudaEventDisableTiming));
	}
	CU(cudaEventCreateWithFlags(&c->ev_input_ready, cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&c->ev_input_consumed, cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&c->ev_drain, cudaEventDisableTiming));
	CU(cudaMalloc(&c->d_tab, sizeof(vdl2_tables)));
	CU(cudaMemcpy(c->d_tab, &c->tab.t, sizeof(vdl2_tables), cudaMemcpyHostToDevice));
	CU(cudaMalloc(&c->d_samples, (size_t)c->max_pairs * sizeof(float4)));
	for(int i = 0; i < 2; i++) {
		CU(cudaMalloc(&c->d_dec2[i], (size_t)c->max_dec * c->n_chp * sizeof(float2)));
		CU(cudaMemset(c->d_dec2[i], 0, (size_t)c->max_dec * c->n_chp * sizeof(float2)));
	}
	CU(cudaMalloc(&c->d_phase, (size_t)(c->max_dec + VDL2_SYNC_BUFLEN) * c->n_chp * sizeof(float)));
	CU(cudaMemset(c->d_phase, 0, (size_t)(c->max_dec + VDL2_SYNC_BUFLEN) * c->n_chp * sizeof(float)));
	CU(cudaMalloc(&c->d_mag, (size_t)c->max_dec * c->n_chp * sizeof(float)));
	CU(cudaMalloc(&c->d_hist_tmp, (size_t)VDL2_SYNC_BUFLEN * c->n_chp * sizeof(float)));
	CU(cudaMalloc(&c->d_k1, (size_t)K1_NFIELDS * c->n_chp * 4));
	CU(cudaMalloc(&c->d_k2, (size_t)K2_NFIELDS * c->n_chp * 4));
	CU(cudaMalloc(&c->d_counters, (size_t)VDL2_NUM_COUNTERS * c->n_chp * 4));
	CU(cudaMalloc(&c->d_ring, (size_t)VDL2_SYNC_BUFLEN * c->n_chp * 4));
	CU(cudaMalloc(&c->d_pool, (size_t)c->n_slots * sizeof(vdl2_burst_slot)));
	CU(cudaMalloc(&c->d_free, (size_t)c->n_slots * 4));
	CU(cudaMalloc(&c->d_ready, (size_t)c->n_slots * 4));
	CU(cudaMalloc(&c->d_ctl, sizeof(vdl2_queue_ctl)));
	CU(cudaMalloc(&c->d_events, (size_t)c->event_cap * sizeof(vdl2gpu_event)));
	CU(cudaMemset(c->d_counters, 0, (size_t)VDL2_NUM_COUNTERS * c->n_chp * 4));
	CU(cudaMemset(c->d_ring, 0, (size_t)VDL2_SYNC_BUFLEN * c->n_chp * 4));
	CU(cudaMemset(c->d_pool, 0, (size_t)c->n_slots * sizeof(vdl2_burst_slot)));

	/* per-channel state: vdl2_channel_init + demod_reset (src/demod.c:205-220,379-392), process_samples
	 * locals (src/demod.c:289-298) */
	std::vector<uint32_t> k1((size_t)K1_NFIELDS * c->n_chp, 0), k2((size_t)K2_NFIELDS * c->n_chp, 0);
	auto fbits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
	for(uint32_t ch = 0; ch < c->n_ch; ch++) {
		/* a channel on the centre frequency skips the mixer in the reference (src/demod.c:312); with a zero
		 * phase step the table gives cos = 1, sin = 0 and the products are exact, so no branch is needed */
		k1[(size_t)K1_DPHI * c->n_chp + ch] = (cfg->centerfreq != c->freqs[ch]) ? dphi_for(cfg->centerfreq, c->freqs[ch], cfg->sample_rate) : 0u;
		k2[(size_t)K2_MAG_NF * c->n_chp + ch] = fbits(2.0f);
		k2[(size_t)K2_PHERR1 * c->n_chp + ch] = fbits(1000.f);
		k2[(size_t)K2_PHERR2 * c->n_chp + ch] = fbits(1000.f);
		k2[(size_t)K2_STATE * c->n_chp + ch] = VDL2_DEC_HEADER << VDL2_DEC_SHIFT;
		k2[(size_t)K2_NEED_BITS * c->n_chp + ch] = VDL2_HEADER_LEN;
		k2[(size_t)K2_SLOT * c->n_chp + ch] = (uint32_t)-1;
		k2[(size_t)K2_FREQ * c->n_chp + ch] = c->freqs[ch];
		k2[(size_t)K2_PURE_RUN * c->n_chp + ch] = 0x40000000u;      /* VDL2_PURE_SATURATED: ring of zeros == history of zeros */
	}
	CU(cudaMemcpy(c->d_k1, k1.data(), k1.size() * 4, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(c->d_k2, k2.data(), k2.size() * 4, cudaMemcpyHostToDevice));
	std::vector<int32_t> fl(c->n_slots);
	for(uint32_t i = 0; i < c->n_slots; i++) fl[i] = (int32_t)i;
	CU(cudaMemcpy(c->d_free, fl.data(), fl.size() * 4, cudaMemcpyHostToDevice));
	vdl2_queue_ctl ctl;
	memset(&ctl, 0, sizeof(ctl));
	ctl.free_top = (int32_t)c->n_slots;
	CU(cudaMemcpy(c->d_ctl, &ctl, sizeof(ctl), cudaMemcpyHostToDevice));

	c->chunks.resize(n_inflight);
	for(auto &s : c->chunks) {
		CU(cudaHostAlloc((void **)&s.h_raw, max_bytes, cudaHostAllocDefault));
		CU(cudaMalloc(&s.d_raw, max_bytes));
		CU(cudaHostAlloc((void **)&s.h_out, sizeof(vdl2_out_header) + c->out_cap, cudaHostAllocMapped));
		CU(cudaHostGetDevicePointer((void **)&s.d_out, s.h_out, 0));
		memset(s.h_out, 0, sizeof(vdl2_out_header));
		CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
		for(auto &ev : s.tk) CU(cudaEventCreate(&ev));
	}
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_create(const vdl2gpu_config *cfg, vdl2gpu_ctx **out) {
	if(!cfg || !out || !cfg->freqs || cfg->n_channels == 0 || cfg->oversample == 0 || cfg->sample_fmt > 1
			|| cfg->sample_rate != (uint32_t)VDL2_SYMBOL_RATE * VDL2_SPS * cfg->oversample) {
		snprintf(g_last_error, sizeof(g_last_error), "bad vdl2gpu_config (sample_rate must be 105000*oversample)");
		return VDL2GPU_EINVAL;
	}
	vdl2gpu_ctx *c = new vdl2gpu_ctx();
	int rc = create_impl(cfg, c);
	if(rc != VDL2GPU_OK) { free_ctx(c); *out = nullptr; return rc; }
	*out = c;
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_destroy(vdl2gpu_ctx *ctx) { return free_ctx(ctx); }

/* ------------------------------------------------------------------------------------------------
 * harvesting: burst records (mapped pinned memory) -> frames
 * ---------------------------------------------------------------------------------------------- */
static uint32_t synd_weight_of(uint32_t syn) {       /* src/decode.c:98-100 */
	static const uint8_t w[32] = { 0, 1, 1, 2, 1, 2, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1, 1, 2, 1, 2, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1, 1 };
	return w[syn & 31u];
}

static void harvest(vdl2gpu_ctx *c, chunk_slot &s) {
	if(s.timed) {
		static const int from[4] = { 0, 1, 3, 4 };              /* K0, K1 on the front stream; K2(+K2a), K3 on the back stream */
		for(int k = 0; k < 4; k++) {
			float ms = 0.f;
			if(cudaEventElapsedTime(&ms, s.tk[from[k]], s.tk[from[k] + 1]) == cudaSuccess) { c->k_ms[k] += ms; c->k_launches[k]++; }
		}
		s.timed = false;
	}
	const vdl2_out_header *h = reinterpret_cast<const vdl2_out_header *>(s.h_out);
	const uint8_t *base = s.h_out + sizeof(vdl2_out_header);
	const vdl2_out_header hcopy = *h; h = &hcopy;
	/* one copy out of the mapped region, so the region can be handed back to the device at once */
	c->blobs.emplace_back(base, base + std::min(h->bytes_used, c->out_cap));
	const uint32_t blob_id = (uint32_t)c->blobs.size() - 1;
	base = c->blobs.back().data();
	c->stats.out_bytes += h->bytes_used;            /* out_bytes: record bytes written by K3 (D2H traffic) */
	std::vector<const vdl2_burst_record *> recs;
	uint32_t off = 0;
	for(uint32_t k = 0; k < h->n_records && off + sizeof(vdl2_burst_record) <= h->bytes_used; k++) {
		const vdl2_burst_record *r = reinterpret_cast<const vdl2_burst_record *>(base + off);
		if(r->rec_bytes < sizeof(vdl2_burst_record) || off + r->rec_bytes > h->bytes_used) break;
		recs.push_back(r);
		off += r->rec_bytes;
	}
	std::sort(recs.begin(), recs.end(), [](const vdl2_burst_record *a, const vdl2_burst_record *b) {
		return a->channel != b->channel ? a->channel < b->channel : a->burst_seq < b->burst_seq;
	});
	c->stats.pool_overflows = h->pool_overflows;
	c->stats.out_overflows = h->out_overflows;
	const double rate = (double)c->cfg.sample_rate / (double)c->cfg.oversample;     /* decimated samples per second */
	for(const vdl2_burst_record *r : recs) {
		c->stats.bursts++;
		if(r->status != VDL2_BURST_OK) c->stats.burst_errors++;
		for(uint32_t q = 0; q < r->num_blocks && q < VDL2_MAX_BLOCKS; q++)
			if(r->rs_ret[q] != -128) { c->stats.blocks_processed++; if(r->rs_ret[q] >= 0) c->stats.blocks_fec_ok++; }
		const uint32_t *tab = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(r) + sizeof(vdl2_burst_record));
		const uint8_t *bytes = reinterpret_cast<const uint8_t *>(tab + r->n_frames);
		uint32_t foff = 0;
		const uint64_t sync_idx = (uint64_t)r->sync_lo | ((uint64_t)r->sync_hi << 32);
		for(uint32_t k = 0; k < r->n_frames; k++) {
			const uint32_t len = tab[k] & 0xFFFFu, crc = tab[k] >> 16;
			pending_frame pf;
			memset(&pf.f, 0, sizeof(pf.f));
			pf.blob = blob_id;
			pf.offset = (uint32_t)((bytes + foff) - base);
			foff += len;
			pf.f.channel = r->channel; pf.f.freq = r->freq; pf.f.burst_seq = r->burst_seq; pf.f.idx = (int32_t)k;
			pf.f.len = len;
			pf.f.synd_weight = synd_weight_of(r->syndrome);
			pf.f.datalen_octets = r->datalen_bits / 8 + ((r->datalen_bits % 8) != 0);
			pf.f.num_fec_corrections = r->num_fec_corrections;
			pf.f.frame_pwr = r->frame_pwr; pf.f.mag_nf = r->mag_nf;
			pf.f.frame_pwr_dbfs = 10.0f * log10f(r->frame_pwr);                  /* src/decode.c:180 */
			pf.f.nf_pwr_dbfs = 20.0f * log10f(r->mag_nf + 0.001f);               /* src/decode.c:181 */
			pf.f.ppm_error = r->ppm_error;
			pf.f.sync_dec_index = sync_idx;
			/* the reference stamps gettimeofday() at sync (src/demod.c:246); here: arrival time of the chunk
			 * being processed when the burst completed, moved back by the distance to the sync sample */
			double back = ((double)(s.dec_base + s.n_dec) - (double)sync_idx) / rate;
			double ts = (double)s.arrival.tv_sec + 1e-6 * (double)s.arrival.tv_usec - back;
			pf.f.burst_timestamp.tv_sec = (time_t)floor(ts);
			pf.f.burst_timestamp.tv_usec = (suseconds_t)((ts - floor(ts)) * 1e6);
			pf.f.fcs_residue = (uint16_t)crc;
			pf.f.fcs_ok = (len >= 11 && crc == 0xF0B8u) ? 1 : 0;
			c->stats.msg_good++;
			if(len >= 11) { if(crc == 0xF0B8u) c->stats.fcs_good++; else c->stats.fcs_bad++; }
			c->pending.push_back(pf);
		}
	}
	c->stats.chunks_completed++;
	s.busy = false;
}

static int deliver(vdl2gpu_ctx *c, vdl2gpu_frame_cb cb, void *user) {
	int n = (int)c->pending.size();
	if(cb) {
		for(auto &pf : c->pending) {
			pf.f.data = c->blobs[pf.blob].data() + pf.offset;
			cb(&pf.f, user);
		}
	}
	c->pending.clear();
	c->blobs.clear();
	return n;
}

/* ------------------------------------------------------------------------------------------------
 * data path
 * ---------------------------------------------------------------------------------------------- */
static int acquire_slot(vdl2gpu_ctx *c, chunk_slot **out) {
	chunk_slot &s = c->chunks[c->next_slot];
	if(s.busy) {
		/* back-pressure: the producer blocks until the oldest chunk has drained (cf. the demods_ready
		 * barrier in src/demod.c:342) */
		CU(cudaEventSynchronize(s.done));
		harvest(c, s);
		if(!c->inflight.empty() && c->inflight.front() == c->next_slot) c->inflight.pop_front();
	}
	*out = &s;
	return VDL2GPU_OK;
}

static int run_chain(vdl2gpu_ctx *c, chunk_slot &s, const void *d_raw, uint32_t n_pairs, uint32_t k0_fmt = 0xFFFFFFFFu) {
	const uint32_t os = c->cfg.oversample;
	s.first_pair = c->total_pairs;
	s.n_pairs = n_pairs;
	s.dec_base = c->total_dec;
	s.n_dec = (c->decim_cnt + n_pairs) / os;
	gettimeofday(&s.arrival, NULL);
	s.timed = c->timing;
	const int db = (int)(c->chunk_seq & 1u);                 /* decimated-sample buffer of this chunk */
	float2 *d_dec = c->d_dec2[db];
	/* ---- front stage: K0, K1 ---- */
	if(s.timed) CU(cudaEventRecord(s.tk[0], c->stream));
	KL(vdl2_launch_k0(d_raw, n_pairs, k0_fmt == 0xFFFFFFFFu ? c->cfg.sample_fmt : k0_fmt, c->d_tab->levels, reinterpret_cast<float *>(c->d_samples), c->stream));
	CU(cudaEventRecord(c->ev_input_consumed, c->stream));
	if(s.timed) CU(cudaEventRecord(s.tk[1], c->stream));
	if(c->chunk_seq >= 2) CU(cudaStreamWaitEvent(c->stream, c->ev_back_done[db], 0));   /* K2 of chunk c-2 has read this buffer */
	/* K1 (one warp per SM sub-partition, latency-bound) pairs well with the equally latency-bound walker K2 and K3 of
	 * the previous chunk, but not with K2a, a full-occupancy issue-bound pass: let K2a of chunk c-1 finish first */
	if(c->chunk_seq >= 1 && c->s_back != c->stream) CU(cudaStreamWaitEvent(c->stream, c->ev_k2a_done, 0));
	vdl2_k1_params p1;
	p1.samples = c->d_samples; p1.n_pairs = n_pairs; p1.oversample = os; p1.cnt0 = c->decim_cnt;
	p1.n_ch = c->n_ch; p1.n_chp = c->n_chp; p1.dec = d_dec; p1.state = c->d_k1;
	p1.lut = reinterpret_cast<const float4 *>(c->d_tab->lut);
	p1.a0 = c->tab.t.A[0]; p1.a1 = c->tab.t.A[1]; p1.a2 = c->tab.t.A[2]; p1.b1 = c->tab.t.B[1]; p1.b2 = c->tab.t.B[2];
	p1.one = 1.0f; p1.neg_one = -1.0f; p1.two = 2.0f;
	KL(vdl2_launch_k1(&p1, (c->cfg.flags & VDL2GPU_FLAG_K1_SCALAR) ? 1 : 0, c->stream));
	if(s.timed) CU(cudaEventRecord(s.tk[2], c->stream));
	CU(cudaEventRecord(c->ev_k1_done[db], c->stream));
	/* ---- back stage: K2a, K2, K3 ---- */
	CU(cudaStreamWaitEvent(c->s_back, c->ev_k1_done[db], 0));
	if(s.timed) CU(cudaEventRecord(s.tk[3], c->s_back));
	vdl2_k2_params p2;
	p2.dec = d_dec; p2.phase = c->d_phase; p2.mag = c->d_mag; p2.hist_tmp = c->d_hist_tmp; p2.n_dec = s.n_dec; p2.n_ch = c->n_ch; p2.n_chp = c->n_chp; p2.dec_base = s.dec_base;
	p2.state = c->d_k2; p2.ring = c->d_ring; p2.tables = c->d_tab; p2.max_ppm = c->cfg.max_ppm; p2.s27 = c->tab.s27;
	p2.pool = c->d_pool; p2.free_list = c->d_free; p2.ready = c->d_ready; p2.ctl = c->d_ctl;
	p2.events = c->d_events; p2.event_cap = c->event_cap; p2.trace = (c->cfg.flags & VDL2GPU_FLAG_TRACE) ? 1u : 0u;
	KL(vdl2_launch_k2a(&p2, c->s_back));
	CU(cudaEventRecord(c->ev_k2a_done, c->s_back));
	KL(vdl2_launch_k2(&p2, c->s_back));
	CU(cudaEventRecord(c->ev_back_done[db], c->s_back));
	if(s.timed) CU(cudaEventRecord(s.tk[4], c->s_back));
	vdl2_k3_params p3;
	p3.pool = c->d_pool; p3.free_list = c->d_free; p3.ready = c->d_ready; p3.ctl = c->d_ctl; p3.tables = c->d_tab;
	p3.out = s.d_out; p3.out_cap = c->out_cap; p3.n_chp = c->n_chp; p3.counters = c->d_counters;
	KL(vdl2_launch_k3(&p3, 148u * 16u, c->s_back));      /* 16 resident blocks per SM (11.4 KB shared memory each): ~1.5 bursts per block per chunk at the bench traffic */
	if(s.timed) CU(cudaEventRecord(s.tk[5], c->s_back));
	CU(cudaEventRecord(s.done, c->s_back));
	c->chunk_seq++;
	s.busy = true;
	c->inflight.push_back(c->next_slot);
	c->next_slot = (c->next_slot + 1) % (uint32_t)c->chunks.size();
	c->decim_cnt = (c->decim_cnt + n_pairs) % os;
	c->total_pairs += n_pairs;
	c->total_dec += s.n_dec;
	c->last_n_dec = s.n_dec;
	c->stats.chunks_submitted++;
	c->stats.iq_samples += n_pairs;
	c->stats.dec_samples += s.n_dec;
	c->stats.kernel_launches += (n_pairs ? 2 : 0) + (s.n_dec ? (s.n_dec >= VDL2_SYNC_BUFLEN ? 3 : 4) : 0) + 2;
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_submit(vdl2gpu_ctx *c, const void *iq, uint32_t len) {
	if(!c || (!iq && len)) return VDL2GPU_EINVAL;
	if(len == 0) return VDL2GPU_OK;                                   /* src/demod.c:341,358 */
	if(len > c->cfg.max_chunk_bytes) return VDL2GPU_ETOOBIG;
	const uint32_t n_pairs = len / (c->cfg.sample_fmt == VDL2GPU_FMT_S16_LE ? 4u : 2u);
	if(n_pairs == 0) return VDL2GPU_OK;
	CU(cudaSetDevice(c->device));
	chunk_slot *s;
	int rc = acquire_slot(c, &s);
	if(rc) return rc;
	memcpy(s->h_raw, iq, len);
	CU(cudaMemcpyAsync(s->d_raw, s->h_raw, len, cudaMemcpyHostToDevice, c->stream));
	return run_chain(c, *s, s->d_raw, n_pairs);
}

extern "C" int vdl2gpu_submit_planar_s16(vdl2gpu_ctx *c, const int16_t *xi, const int16_t *xq, uint32_t n_pairs) {
	if(!c || ((!xi || !xq) && n_pairs)) return VDL2GPU_EINVAL;
	if(c->cfg.sample_fmt != VDL2GPU_FMT_S16_LE) return VDL2GPU_EINVAL;
	if(n_pairs == 0) return VDL2GPU_OK;
	if((uint64_t)n_pairs * 4u > c->cfg.max_chunk_bytes) return VDL2GPU_ETOOBIG;
	CU(cudaSetDevice(c->device));
	chunk_slot *s;
	int rc = acquire_slot(c, &s);
	if(rc) return rc;
	memcpy(s->h_raw, xi, (size_t)n_pairs * 2);
	memcpy(s->h_raw + (size_t)n_pairs * 2, xq, (size_t)n_pairs * 2);
	CU(cudaMemcpyAsync(s->d_raw, s->h_raw, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, c->stream));
	return run_chain(c, *s, s->d_raw, n_pairs, 2u);
}

/* src/fmtr-binary.c + src/output-file.c:181-189: one raw frame in the reference's archive format, i.e. a 2-octet
 * big-endian length (its own two octets included) followed by the proto3 message dumpvdl2.raw_avlc_frame
 * (proto/dumpvdl2.proto:25-48).  Files made of these records replay through an unmodified
 * `dumpvdl2 --raw-frames-file` (src/input-raw_frames_file.c:33-75).  Returns the record size or a negative error. */
static size_t pb_varint(uint8_t *o, uint64_t v) { size_t n = 0; do { uint8_t b = v & 0x7Fu; v >>= 7; o[n++] = (uint8_t)(b | (v ? 0x80u : 0u)); } while(v); return n; }
static size_t pb_u32_field(uint8_t *o, int field, uint32_t v) { if(!v) return 0; size_t n = pb_varint(o, (uint64_t)field << 3); return n + pb_varint(o + n, v); }
static size_t pb_i64_field(uint8_t *o, int field, int64_t v) { if(!v) return 0; size_t n = pb_varint(o, (uint64_t)field << 3); return n + pb_varint(o + n, (uint64_t)v); }
static size_t pb_f32_field(uint8_t *o, int field, float v) { uint32_t u; memcpy(&u, &v, 4); if(!u) return 0; size_t n = pb_varint(o, ((uint64_t)field << 3) | 5u); memcpy(o + n, &u, 4); return n + 4; }

extern "C" int vdl2gpu_serialize_raw_frame(const vdl2gpu_frame *f, const char *station_id, uint8_t *out, size_t cap) {
	if(!f || !out || (f->len && !f->data)) return VDL2GPU_EINVAL;
	uint8_t ts[24], md[96 + 256];
	size_t nts = 0, nmd = 0;
	nts += pb_i64_field(ts + nts, 1, (int64_t)f->burst_timestamp.tv_sec);
	nts += pb_i64_field(ts + nts, 2, (int64_t)f->burst_timestamp.tv_usec);
	size_t sl = station_id ? strlen(station_id) : 0;
	if(sl > 255) sl = 255;                                          /* STATION_ID_LEN_MAX, src/dumpvdl2.h:193 */
	if(sl) { nmd += pb_varint(md + nmd, (1u << 3) | 2u); nmd += pb_varint(md + nmd, sl); memcpy(md + nmd, station_id, sl); nmd += sl; }
	nmd += pb_u32_field(md + nmd, 2, f->freq);
	nmd += pb_u32_field(md + nmd, 3, f->synd_weight);
	nmd += pb_u32_field(md + nmd, 4, f->datalen_octets);
	nmd += pb_f32_field(md + nmd, 5, f->frame_pwr_dbfs);
	nmd += pb_f32_field(md + nmd, 6, f->nf_pwr_dbfs);
	nmd += pb_f32_field(md + nmd, 7, f->ppm_error);
	nmd += pb_i64_field(md + nmd, 8, 1);                            /* metadata version, src/decode.c:177 */
	nmd += pb_i64_field(md + nmd, 9, (int64_t)f->num_fec_corrections);
	nmd += pb_i64_field(md + nmd, 10, (int64_t)f->idx);
	nmd += pb_varint(md + nmd, (11u << 3) | 2u); nmd += pb_varint(md + nmd, nts); memcpy(md + nmd, ts, nts); nmd += nts;
	uint8_t hdr[16];
	size_t nh = pb_varint(hdr, (1u << 3) | 2u); nh += pb_varint(hdr + nh, nmd);
	uint8_t dh[16];
	size_t nd = f->len ? pb_varint(dh, (2u << 3) | 2u) : 0;
	if(f->len) nd += pb_varint(dh + nd, f->len);
	const size_t total = 2 + nh + nmd + nd + f->len;
	if(total > 65536 || total > cap) return VDL2GPU_ETOOBIG;         /* OUT_BINARY_FRAME_LEN_MAX, src/output-file.h:26 */
	uint8_t *o = out;
	*o++ = (uint8_t)(total >> 8); *o++ = (uint8_t)total;
	memcpy(o, hdr, nh); o += nh; memcpy(o, md, nmd); o += nmd;
	memcpy(o, dh, nd); o += nd; if(f->len) memcpy(o, f->data, f->len);
	return (int)total;
}

extern "C" int vdl2gpu_submit_device(vdl2gpu_ctx *c, const void *dev_iq, uint32_t len, void *producer_stream) {
	if(!c || (!dev_iq && len)) return VDL2GPU_EINVAL;
	if(len == 0) return VDL2GPU_OK;
	if(len > c->cfg.max_chunk_bytes) return VDL2GPU_ETOOBIG;
	const uint32_t n_pairs = len / (c->cfg.sample_fmt == VDL2GPU_FMT_S16_LE ? 4u : 2u);
	if(n_pairs == 0) return VDL2GPU_OK;
	CU(cudaSetDevice(c->device));
	chunk_slot *s;
	int rc = acquire_slot(c, &s);
	if(rc) return rc;
	CU(cudaEventRecord(c->ev_input_ready, (cudaStream_t)producer_stream));
	CU(cudaStreamWaitEvent(c->stream, c->ev_input_ready, 0));
	return run_chain(c, *s, dev_iq, n_pairs);
}

extern "C" int vdl2gpu_wait_input_consumed(vdl2gpu_ctx *c, void *stream) {
	if(!c) return VDL2GPU_EINVAL;
	CU(cudaStreamWaitEvent((cudaStream_t)stream, c->ev_input_consumed, 0));
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_stream_wait(vdl2gpu_ctx *c, void *stream) {
	if(!c) return VDL2GPU_EINVAL;
	CU(cudaSetDevice(c->device));
	CU(cudaEventRecord(c->ev_drain, c->s_back));        /* every chunk's chain ends on the back stream */
	CU(cudaStreamWaitEvent((cudaStream_t)stream, c->ev_drain, 0));
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_poll(vdl2gpu_ctx *c, vdl2gpu_frame_cb cb, void *user) {
	if(!c) return VDL2GPU_EINVAL;
	CU(cudaSetDevice(c->device));
	while(!c->inflight.empty()) {
		chunk_slot &s = c->chunks[c->inflight.front()];
		if(!s.busy) { c->inflight.pop_front(); continue; }
		cudaError_t q = cudaEventQuery(s.done);
		if(q == cudaErrorNotReady) break;
		if(q != cudaSuccess) return fail_cuda(q, "cudaEventQuery", __LINE__);
		harvest(c, s);
		c->inflight.pop_front();
	}
	return deliver(c, cb, user);
}

extern "C" int vdl2gpu_flush(vdl2gpu_ctx *c, vdl2gpu_frame_cb cb, void *user) {
	if(!c) return VDL2GPU_EINVAL;
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaStreamSynchronize(c->s_back));
	while(!c->inflight.empty()) {
		chunk_slot &s = c->chunks[c->inflight.front()];
		if(s.busy) harvest(c, s);
		c->inflight.pop_front();
	}
	return deliver(c, cb, user);
}

static int read_counters(vdl2gpu_ctx *c, std::vector<uint32_t> &k3, std::vector<uint32_t> &sync, std::vector<uint32_t> &hdr) {
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaStreamSynchronize(c->s_back));
	k3.resize((size_t)VDL2_NUM_COUNTERS * c->n_chp);
	sync.resize(c->n_chp); hdr.resize(c->n_chp);
	CU(cudaMemcpy(k3.data(), c->d_counters, k3.size() * 4, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(sync.data(), c->d_k2 + (size_t)K2_CNT_SYNC * c->n_chp, c->n_chp * 4, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(hdr.data(), c->d_k2 + (size_t)K2_CNT_HDR_GOOD * c->n_chp, c->n_chp * 4, cudaMemcpyDeviceToHost));
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_get_stats(vdl2gpu_ctx *c, vdl2gpu_stats *out) {
	if(!c || !out) return VDL2GPU_EINVAL;
	std::vector<uint32_t> k3, sync, hdr;
	int rc = read_counters(c, k3, sync, hdr);
	if(rc) return rc;
	uint64_t a = 0, b = 0;
	for(uint32_t ch = 0; ch < c->n_ch; ch++) { a += sync[ch]; b += hdr[ch]; }
	c->stats.demod_sync_good = a;
	c->stats.decoder_crc_good = b;
	*out = c->stats;
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_get_channel_counters(vdl2gpu_ctx *c, uint64_t *out, uint32_t n_channels) {
	if(!c || !out || n_channels > c->n_ch) return VDL2GPU_EINVAL;
	std::vector<uint32_t> k3, sync, hdr;
	int rc = read_counters(c, k3, sync, hdr);
	if(rc) return rc;
	for(uint32_t ch = 0; ch < n_channels; ch++) {
		uint64_t *o = out + (size_t)ch * VDL2_NUM_COUNTERS;
		for(int k = 0; k < VDL2_NUM_COUNTERS; k++) o[k] = k3[(size_t)k * c->n_chp + ch];
		o[VDL2_CNT_SYNC_GOOD] = sync[ch];
		o[VDL2_CNT_HDR_CRC_GOOD] = hdr[ch];
	}
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_get_tables(vdl2gpu_ctx *c, float levels[256], float sin_lut[257], float cos_lut[257],
		float A[3], float B[3], float lr_X[16], float *lr_denom, float pr_phase[16]) {
	if(!c) return VDL2GPU_EINVAL;
	if(levels) memcpy(levels, c->tab.t.levels, sizeof(c->tab.t.levels));
	if(sin_lut) memcpy(sin_lut, c->tab.sin_lut, sizeof(c->tab.sin_lut));
	if(cos_lut) memcpy(cos_lut, c->tab.cos_lut, sizeof(c->tab.cos_lut));
	if(A) memcpy(A, c->tab.t.A, 12);
	if(B) memcpy(B, c->tab.t.B, 12);
	if(lr_X) memcpy(lr_X, c->tab.t.lr_X, 64);
	if(lr_denom) *lr_denom = c->tab.t.lr_denom;
	if(pr_phase) memcpy(pr_phase, c->tab.t.pr_phase, 64);
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_read_dec(vdl2gpu_ctx *c, float *out, size_t cap_floats, uint32_t *n_dec) {
	if(!c || !out || !n_dec) return VDL2GPU_EINVAL;
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaStreamSynchronize(c->s_back));
	*n_dec = c->last_n_dec;
	if((size_t)c->last_n_dec * c->n_ch * 2 > cap_floats) return VDL2GPU_ETOOBIG;
	if(c->last_n_dec == 0) return VDL2GPU_OK;
	CU(cudaMemcpy2D(out, (size_t)c->n_ch * sizeof(float2), c->d_dec2[(c->chunk_seq + 1) & 1u], (size_t)c->n_chp * sizeof(float2),
			(size_t)c->n_ch * sizeof(float2), c->last_n_dec, cudaMemcpyDeviceToHost));
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_read_events(vdl2gpu_ctx *c, vdl2gpu_event *out, uint32_t cap) {
	if(!c || !out) return VDL2GPU_EINVAL;
	if(!(c->cfg.flags & VDL2GPU_FLAG_TRACE)) return 0;
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaStreamSynchronize(c->s_back));
	vdl2_queue_ctl ctl;
	CU(cudaMemcpy(&ctl, c->d_ctl, sizeof(ctl), cudaMemcpyDeviceToHost));
	uint32_t total = std::min(ctl.n_events, c->event_cap);
	uint32_t n = total > c->events_read ? total - c->events_read : 0;
	if(n > cap) n = cap;
	if(n) CU(cudaMemcpy(out, (const vdl2gpu_event *)c->d_events + c->events_read, (size_t)n * sizeof(vdl2gpu_event), cudaMemcpyDeviceToHost));
	c->events_read += n;
	return (int)n;
}

extern "C" int vdl2gpu_enable_timing(vdl2gpu_ctx *c, int on) {
	if(!c) return VDL2GPU_EINVAL;
	c->timing = on != 0;
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_get_kernel_ms(vdl2gpu_ctx *c, double ms[4], uint64_t launches[4]) {
	if(!c || !ms) return VDL2GPU_EINVAL;
	for(int k = 0; k < 4; k++) { ms[k] = c->k_ms[k]; if(launches) launches[k] = c->k_launches[k]; }
	return VDL2GPU_OK;
}

/* ------------------------------------------------------------------------------------------------
 * raw launch stubs (device pointers in, device pointers out)
 * ---------------------------------------------------------------------------------------------- */
static vdl2_tables *g_stub_tables = nullptr;       /* GF tables for the stand-alone RS stub, per process */

extern "C" int vdl2gpu_launch_convert(const void *raw, uint32_t n_pairs, uint32_t sample_fmt, const float *levels256,
		float *samples_out, void *stream) {
	/* stand-alone form writes float2 {re, im}: convert into a temporary float4 layout is not needed by callers,
	 * so this stub runs the K0 kernel into a scratch buffer and compacts */
	if(!raw || !samples_out || sample_fmt > 1 || (sample_fmt == 0 && !levels256)) return VDL2GPU_EINVAL;
	float4 *tmp = nullptr;
	CU(cudaMalloc(&tmp, (size_t)std::max(n_pairs, 1u) * sizeof(float4)));
	int rc = vdl2_launch_k0(raw, n_pairs, sample_fmt, levels256, reinterpret_cast<float *>(tmp), (cudaStream_t)stream);
	if(rc == 0 && n_pairs)
		rc = (int)cudaMemcpy2DAsync(samples_out, sizeof(float2), tmp, sizeof(float4), sizeof(float2), n_pairs, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
	cudaStreamSynchronize((cudaStream_t)stream);
	cudaFree(tmp);
	if(rc) return fail_cuda((cudaError_t)rc, "vdl2gpu_launch_convert", __LINE__);
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_launch_fcs_crc16(const uint8_t *frames, const uint32_t *offsets, const uint32_t *lens,
		uint32_t n_frames, uint16_t *residues_out, void *stream) {
	if(n_frames && (!frames || !offsets || !lens || !residues_out)) return VDL2GPU_EINVAL;
	KL(vdl2_launch_k4(frames, offsets, lens, n_frames, residues_out, (cudaStream_t)stream));
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_launch_rs_verify(uint8_t *blocks, const int32_t *fec_octets, uint32_t n_blocks, int32_t *ret_out, void *stream) {
	if(n_blocks && (!blocks || !fec_octets || !ret_out)) return VDL2GPU_EINVAL;
	if(!g_stub_tables) {
		host_tables *h = new host_tables();
		memset(h, 0, sizeof(*h));
		make_gf(h->t);
		CU(cudaMalloc(&g_stub_tables, sizeof(vdl2_tables)));
		CU(cudaMemcpy(g_stub_tables, &h->t, sizeof(vdl2_tables), cudaMemcpyHostToDevice));
		delete h;
	}
	KL(vdl2_launch_rs(blocks, fec_octets, n_blocks, ret_out, g_stub_tables, (cudaStream_t)stream));
	return VDL2GPU_OK;
}
