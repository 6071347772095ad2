/*
 * vdl2_mg.cu — multi-GPU ingest helper of libvdl2gpu.so (C-ABI: vdl2gpu_mg_* in include/vdl2gpu.h).
 *
 * The path shards by channel (global channel k lives on GPU k mod N, one process per GPU): the reference's channel
 * threads share nothing but the read-only sample buffer (src/dumpvdl2.c:117-135, src/demod.c:50,300-301).  The only
 * data that crosses GPUs is therefore the raw IQ itself: the rank that owns the SDR / file (rank 0) hands every
 * step's chunks to all ranks, each rank demodulates its shard straight out of the receive buffer
 * (vdl2gpu_submit_device).  This file owns that fan-out: a double-buffered receive area per rank, filled one step
 * ahead of the kernels that read it, in one of two ways
 *
 *   VDL2GPU_MG_NCCL         ncclBroadcast on a communication stream of its own (libnccl is dlopen()ed: the library
 *                           loads without it);
 *   VDL2GPU_MG_COPY_ENGINE  rank 0 writes into the peers' receive buffers (CUDA IPC mappings) with cudaMemcpyAsync
 *                           peer copies, i.e. on the copy engines over NVLink, and raises a sequence number in the
 *                           peer's memory with a stream memory operation (cuStreamWriteValue32); the peer's stream
 *                           waits for it (cuStreamWaitValue32) and acknowledges the same way once its K0 kernels have
 *                           read the buffer.  No kernel runs for the transfer and no host thread waits: nothing of the
 *                           fan-out occupies an SM beside the latency-bound K1/K2 kernels.
 *
 * Host code only.  The caller moves the small opaque blobs between the processes (ncclUniqueId / IPC handles) with
 * whatever it has: MPI, a socket, torch.distributed.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/vdl2gpu.h"

extern "C" void vdl2gpu_set_last_error(const char *msg);

#define MG_FAIL(code, ...) do { char b_[400]; snprintf(b_, sizeof(b_), __VA_ARGS__); vdl2gpu_set_last_error(b_); return (code); } while(0)
#define MG_CU(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) MG_FAIL(VDL2GPU_ECUDA, "%s failed at vdl2_mg.cu:%d: %s", #call, __LINE__, cudaGetErrorString(e_)); } while(0)

/* ---- NCCL through dlopen ---- */
typedef struct { char internal[128]; } mg_nccl_id;
typedef void *mg_nccl_comm;
static struct {
	void *h;
	int (*GetUniqueId)(mg_nccl_id *);
	int (*CommInitRank)(mg_nccl_comm *, int, mg_nccl_id, int);
	int (*Broadcast)(const void *, void *, size_t, int, int, mg_nccl_comm, cudaStream_t);
	int (*CommDestroy)(mg_nccl_comm);
	const char *(*GetErrorString)(int);
} NC = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };

static int nccl_load(void) {
	if(NC.h) return VDL2GPU_OK;
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if(!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if(!h) MG_FAIL(VDL2GPU_ENODEV, "libnccl.so.2 not found: %s", dlerror());
	*(void **)&NC.GetUniqueId = dlsym(h, "ncclGetUniqueId");
	*(void **)&NC.CommInitRank = dlsym(h, "ncclCommInitRank");
	*(void **)&NC.Broadcast = dlsym(h, "ncclBroadcast");
	*(void **)&NC.CommDestroy = dlsym(h, "ncclCommDestroy");
	*(void **)&NC.GetErrorString = dlsym(h, "ncclGetErrorString");
	if(!NC.GetUniqueId || !NC.CommInitRank || !NC.Broadcast || !NC.CommDestroy) MG_FAIL(VDL2GPU_ENODEV, "libnccl.so.2 lacks the expected symbols");
	NC.h = h;
	return VDL2GPU_OK;
}
#define MG_NC(call) do { int e_ = (call); if(e_ != 0) MG_FAIL(VDL2GPU_ECUDA, "%s failed at vdl2_mg.cu:%d: %s", #call, __LINE__, NC.GetErrorString ? NC.GetErrorString(e_) : "nccl error"); } while(0)

/* ---- stream memory operations (driver API entry points, resolved at run time: the library does not link libcuda) ---- */
typedef CUresult (*mg_memop_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static mg_memop_fn g_write32 = nullptr, g_wait32 = nullptr;

static int memops_load(void) {
	if(g_write32 && g_wait32) return VDL2GPU_OK;
	cudaDriverEntryPointQueryResult q;
	void *f = nullptr;
	MG_CU(cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &q));
	g_write32 = (mg_memop_fn)f;
	MG_CU(cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q));
	g_wait32 = (mg_memop_fn)f;
	if(!g_write32 || !g_wait32) MG_FAIL(VDL2GPU_ENODEV, "stream memory operations are not available in this driver");
	return VDL2GPU_OK;
}
#define MG_DRV(call) do { CUresult e_ = (call); if(e_ != CUDA_SUCCESS) MG_FAIL(VDL2GPU_ECUDA, "%s failed at vdl2_mg.cu:%d: CUresult %d", #call, __LINE__, (int)e_); } while(0)

struct mg_peer {
	uint8_t *recv[2] = { nullptr, nullptr };     /* the peer's receive halves, mapped into this (rank 0) process */
	uint32_t *flags = nullptr;                   /* the peer's flag words, mapped here: [0..1] filled[half] */
	cudaStream_t s_copy = nullptr;
};

struct vdl2gpu_mg {
	vdl2gpu_ctx *ctx = nullptr;
	int rank = 0, world = 1, mode = VDL2GPU_MG_NCCL, device = 0;
	uint32_t stage_bytes = 0;
	uint8_t *recv[2] = { nullptr, nullptr };     /* this rank's double-buffered receive / staging area */
	uint32_t *flags = nullptr;                   /* device words: [0..1] filled[half] (written by rank 0), then on rank 0 only
	                                              * [2 + 2 r + half] consumed by rank r (written by rank r) */
	uint32_t *root_flags = nullptr;              /* rank != 0: rank 0's flag words mapped here */
	cudaStream_t s_comm = nullptr, s_half[2] = { nullptr, nullptr };
	cudaEvent_t ev_staged[2] = { nullptr, nullptr }, ev_gathered = nullptr;
	uint32_t seq[2] = { 0, 0 };                  /* how many times each half has been filled */
	uint32_t submitted[2] = { 0, 0 };            /* the filling of each half that vdl2gpu_mg_submit last consumed */
	uint32_t next_half = 0;
	bool imported = false;
	mg_nccl_comm comm = nullptr;
	std::vector<mg_peer> peers;                  /* rank 0, copy-engine mode */
};

struct mg_blob {                                  /* what every rank publishes in copy-engine mode */
	cudaIpcMemHandle_t recv[2];
	cudaIpcMemHandle_t flags;
};

extern "C" size_t vdl2gpu_mg_blob_bytes(void) { return sizeof(mg_blob); }

extern "C" int vdl2gpu_mg_unique_id(uint8_t *id, size_t cap) {
	if(!id || cap < sizeof(mg_nccl_id)) return VDL2GPU_EINVAL;
	int rc = nccl_load();
	if(rc) return rc;
	mg_nccl_id u;
	MG_NC(NC.GetUniqueId(&u));
	memcpy(id, &u, sizeof(u));
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_mg_destroy(vdl2gpu_mg *m) {
	if(!m) return VDL2GPU_OK;
	cudaSetDevice(m->device);
	cudaDeviceSynchronize();
	for(auto &p : m->peers) {
		for(int h = 0; h < 2; h++) if(p.recv[h]) cudaIpcCloseMemHandle(p.recv[h]);
		if(p.flags) cudaIpcCloseMemHandle(p.flags);
		if(p.s_copy) cudaStreamDestroy(p.s_copy);
	}
	if(m->root_flags) cudaIpcCloseMemHandle(m->root_flags);
	if(m->comm && NC.CommDestroy) NC.CommDestroy(m->comm);
	for(int h = 0; h < 2; h++) {
		if(m->recv[h]) cudaFree(m->recv[h]);
		if(m->s_half[h]) cudaStreamDestroy(m->s_half[h]);
		if(m->ev_staged[h]) cudaEventDestroy(m->ev_staged[h]);
	}
	if(m->ev_gathered) cudaEventDestroy(m->ev_gathered);
	if(m->flags) cudaFree(m->flags);
	if(m->s_comm) cudaStreamDestroy(m->s_comm);
	delete m;
	return VDL2GPU_OK;
}

static int mg_create_impl(vdl2gpu_mg *m, const uint8_t *nccl_id) {
	MG_CU(cudaGetDevice(&m->device));
	MG_CU(cudaStreamCreateWithFlags(&m->s_comm, cudaStreamNonBlocking));
	for(int h = 0; h < 2; h++) {
		MG_CU(cudaMalloc(&m->recv[h], m->stage_bytes));
		MG_CU(cudaStreamCreateWithFlags(&m->s_half[h], cudaStreamNonBlocking));
		MG_CU(cudaEventCreateWithFlags(&m->ev_staged[h], cudaEventDisableTiming));
	}
	MG_CU(cudaEventCreateWithFlags(&m->ev_gathered, cudaEventDisableTiming));
	const size_t n_flags = 2 + 2 * (size_t)m->world;
	MG_CU(cudaMalloc(&m->flags, n_flags * sizeof(uint32_t)));
	MG_CU(cudaMemset(m->flags, 0, n_flags * sizeof(uint32_t)));
	if(m->mode == VDL2GPU_MG_NCCL) {
		if(!nccl_id) MG_FAIL(VDL2GPU_EINVAL, "vdl2gpu_mg_create: NCCL mode needs the unique id of vdl2gpu_mg_unique_id()");
		int rc = nccl_load();
		if(rc) return rc;
		mg_nccl_id u;
		memcpy(&u, nccl_id, sizeof(u));
		MG_NC(NC.CommInitRank(&m->comm, m->world, u, m->rank));
		m->imported = true;
	} else {
		int rc = memops_load();
		if(rc) return rc;
		if(m->rank == 0) {
			m->peers.resize(m->world);
			for(int r = 1; r < m->world; r++) MG_CU(cudaStreamCreateWithFlags(&m->peers[r].s_copy, cudaStreamNonBlocking));
		}
	}
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_mg_create(vdl2gpu_ctx *ctx, int rank, int world, int mode, const uint8_t *nccl_id, uint32_t stage_bytes, vdl2gpu_mg **out) {
	if(!ctx || !out || world < 1 || rank < 0 || rank >= world || stage_bytes == 0 || (mode != VDL2GPU_MG_NCCL && mode != VDL2GPU_MG_COPY_ENGINE))
		return VDL2GPU_EINVAL;
	vdl2gpu_mg *m = new vdl2gpu_mg();
	m->ctx = ctx; m->rank = rank; m->world = world; m->mode = mode; m->stage_bytes = stage_bytes;
	int rc = mg_create_impl(m, nccl_id);
	if(rc) { vdl2gpu_mg_destroy(m); *out = nullptr; return rc; }
	*out = m;
	return VDL2GPU_OK;
}

/* copy-engine mode, step 1: what this rank lets the others map */
extern "C" int vdl2gpu_mg_export(vdl2gpu_mg *m, uint8_t *blob, size_t cap) {
	if(!m || !blob || cap < sizeof(mg_blob) || m->mode != VDL2GPU_MG_COPY_ENGINE) return VDL2GPU_EINVAL;
	mg_blob b;
	memset(&b, 0, sizeof(b));
	MG_CU(cudaIpcGetMemHandle(&b.recv[0], m->recv[0]));
	MG_CU(cudaIpcGetMemHandle(&b.recv[1], m->recv[1]));
	MG_CU(cudaIpcGetMemHandle(&b.flags, m->flags));
	memcpy(blob, &b, sizeof(b));
	return VDL2GPU_OK;
}

/* copy-engine mode, step 2: `all` = the blobs of ranks 0..world-1 back to back */
extern "C" int vdl2gpu_mg_import(vdl2gpu_mg *m, const uint8_t *all, size_t bytes) {
	if(!m || !all || bytes < sizeof(mg_blob) * (size_t)m->world || m->mode != VDL2GPU_MG_COPY_ENGINE) return VDL2GPU_EINVAL;
	const mg_blob *b = reinterpret_cast<const mg_blob *>(all);
	if(m->rank == 0) {
		for(int r = 1; r < m->world; r++) {
			MG_CU(cudaIpcOpenMemHandle((void **)&m->peers[r].recv[0], b[r].recv[0], cudaIpcMemLazyEnablePeerAccess));
			MG_CU(cudaIpcOpenMemHandle((void **)&m->peers[r].recv[1], b[r].recv[1], cudaIpcMemLazyEnablePeerAccess));
			MG_CU(cudaIpcOpenMemHandle((void **)&m->peers[r].flags, b[r].flags, cudaIpcMemLazyEnablePeerAccess));
		}
	} else {
		MG_CU(cudaIpcOpenMemHandle((void **)&m->root_flags, b[0].flags, cudaIpcMemLazyEnablePeerAccess));
	}
	m->imported = true;
	return VDL2GPU_OK;
}

/* Fill the free half of the receive area with the next step's chunks.  Rank 0 passes the pieces that make up the
 * step (device pointers, or pinned host pointers with src_is_host), `total` bytes in all; the other ranks pass
 * n_src = 0 and the same total.  Returns the half (0/1) to hand to vdl2gpu_mg_submit, or a negative error.
 * Everything is asynchronous; the call may be issued a whole step before the kernels that read the buffer. */
extern "C" int vdl2gpu_mg_stage(vdl2gpu_mg *m, const void *const *src, const uint32_t *src_bytes, uint32_t n_src, int src_is_host, uint32_t total) {
	if(!m || !m->imported || total == 0 || total > m->stage_bytes) return VDL2GPU_EINVAL;
	MG_CU(cudaSetDevice(m->device));
	const int h = (int)(m->next_half & 1u);
	const uint32_t prev = m->seq[h];                 /* this half's previous filling must have been read everywhere */
	const uint32_t seq = prev + 1;
	/* this rank's own K0 kernels of the previous use: the context's "input consumed" event (recorded per chunk) */
	int rc = vdl2gpu_wait_input_consumed(m->ctx, m->s_comm);
	if(rc) return rc;
	if(m->mode == VDL2GPU_MG_COPY_ENGINE && m->rank != 0 && prev != 0 && m->submitted[h] != prev) {
		/* the previous filling of this half was staged but never submitted here: acknowledge it anyway, rank 0 waits for
		 * every rank's acknowledgement before it overwrites the half */
		MG_DRV(g_write32((CUstream)m->s_half[h], (CUdeviceptr)(m->root_flags + 2 + 2 * m->rank + h), prev, CU_STREAM_WRITE_VALUE_DEFAULT));
	}
	if(m->rank == 0) {
		if(n_src == 0 || !src || !src_bytes) return VDL2GPU_EINVAL;
		uint32_t off = 0;
		for(uint32_t k = 0; k < n_src; k++) {
			if(off + src_bytes[k] > total) return VDL2GPU_EINVAL;
			MG_CU(cudaMemcpyAsync(m->recv[h] + off, src[k], src_bytes[k], src_is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, m->s_comm));
			off += src_bytes[k];
		}
		if(off != total) return VDL2GPU_EINVAL;
	}
	if(m->mode == VDL2GPU_MG_NCCL) {
		if(m->world > 1) MG_NC(NC.Broadcast(m->recv[h], m->recv[h], total, /* ncclUint8 */ 1, 0, m->comm, m->s_comm));
		MG_CU(cudaEventRecord(m->ev_staged[h], m->s_comm));
	} else if(m->rank == 0) {
		MG_CU(cudaEventRecord(m->ev_gathered, m->s_comm));
		MG_CU(cudaEventRecord(m->ev_staged[h], m->s_comm));
		for(int r = 1; r < m->world; r++) {
			mg_peer &p = m->peers[r];
			MG_CU(cudaStreamWaitEvent(p.s_copy, m->ev_gathered, 0));
			if(prev) MG_DRV(g_wait32((CUstream)p.s_copy, (CUdeviceptr)(m->flags + 2 + 2 * r + h), prev, CU_STREAM_WAIT_VALUE_GEQ));
			MG_CU(cudaMemcpyAsync(p.recv[h], m->recv[h], total, cudaMemcpyDeviceToDevice, p.s_copy));      /* copy engine, NVLink */
			MG_DRV(g_write32((CUstream)p.s_copy, (CUdeviceptr)(p.flags + h), seq, CU_STREAM_WRITE_VALUE_DEFAULT));
		}
		/* the staging half on rank 0 may only be refilled once every peer copy out of it has run: make the next
		 * gather (on s_comm) wait for them */
		for(int r = 1; r < m->world; r++) {
			cudaEvent_t e = nullptr;
			MG_CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
			MG_CU(cudaEventRecord(e, m->peers[r].s_copy));
			MG_CU(cudaStreamWaitEvent(m->s_comm, e, 0));
			MG_CU(cudaEventDestroy(e));                /* released once the wait has been satisfied */
		}
	}
	m->seq[h] = seq;
	m->next_half++;
	return h;
}

/* Submit the n_chunks chunks of half `h` (staged earlier) to the context, in order. */
extern "C" int vdl2gpu_mg_submit(vdl2gpu_mg *m, int h, uint32_t n_chunks, uint32_t chunk_bytes) {
	if(!m || h < 0 || h > 1 || (uint64_t)n_chunks * chunk_bytes > m->stage_bytes) return VDL2GPU_EINVAL;
	MG_CU(cudaSetDevice(m->device));
	cudaStream_t s = m->s_half[h];
	if(m->mode == VDL2GPU_MG_NCCL || m->rank == 0) {
		MG_CU(cudaStreamWaitEvent(s, m->ev_staged[h], 0));
	} else {
		MG_DRV(g_wait32((CUstream)s, (CUdeviceptr)(m->flags + h), m->seq[h], CU_STREAM_WAIT_VALUE_GEQ));
	}
	for(uint32_t k = 0; k < n_chunks; k++) {
		int rc = vdl2gpu_submit_device(m->ctx, m->recv[h] + (size_t)k * chunk_bytes, chunk_bytes, s);
		if(rc) return rc;
	}
	m->submitted[h] = m->seq[h];
	if(m->mode == VDL2GPU_MG_COPY_ENGINE && m->rank != 0) {
		/* tell rank 0 that this half has been read (after the last chunk's K0) */
		int rc = vdl2gpu_wait_input_consumed(m->ctx, s);
		if(rc) return rc;
		MG_DRV(g_write32((CUstream)s, (CUdeviceptr)(m->root_flags + 2 + 2 * m->rank + h), m->seq[h], CU_STREAM_WRITE_VALUE_DEFAULT));
	}
	return VDL2GPU_OK;
}

extern "C" int vdl2gpu_mg_mode(vdl2gpu_mg *m) { return m ? m->mode : VDL2GPU_EINVAL; }
