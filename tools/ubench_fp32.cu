// tools/ubench_fp32.cu — issue-rate microbenchmark for the FP32 forms K1 can be built from
// (scalar FMUL/FADD/FFMA vs packed FMUL2/FADD2/FFMA2 on sm_100a), in SM cycles per warp-instruction
// per SM sub-partition, at 1/2/4 warps per sub-partition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define DI __device__ __forceinline__
DI u64 mul2(u64 a,u64 b){u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;}
DI u64 add2(u64 a,u64 b){u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;}
DI u64 fma2(u64 a,u64 b,u64 c){u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;}
DI u64 pack(float x,float y){u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y)); return r;}
DI float lo(u64 v){float x,y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); return x;}
DI float hi(u64 v){float x,y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); return y;}

#define NCH 8
template<int OP> __global__ void bench(float *out, long long *cyc, int iters, float a, float b) {
    float x[NCH]; u64 p[NCH];
    for (int i = 0; i < NCH; i++) { x[i] = a + i + threadIdx.x; p[i] = pack(x[i], x[i] + 1.f); }
    u64 pa = pack(a, a), pb = pack(b, b);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            if (OP == 0) x[i] = __fmul_rn(x[i], a);
            if (OP == 1) x[i] = __fadd_rn(x[i], b);
            if (OP == 2) x[i] = __fmaf_rn(x[i], a, b);
            if (OP == 3) p[i] = mul2(p[i], pa);
            if (OP == 4) p[i] = add2(p[i], pb);
            if (OP == 5) p[i] = fma2(p[i], pa, pb);
            if (OP == 6) x[i] = __fmaf_rn(x[i], x[(i + 1) % NCH], x[(i + 3) % NCH]);   // 3 distinct regs
            if (OP == 7) p[i] = fma2(p[i], p[(i + 1) % NCH], p[(i + 3) % NCH]);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < NCH; i++) s += x[i] + lo(p[i]) + hi(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x % 32 == 0) cyc[(blockIdx.x * blockDim.x + threadIdx.x) / 32] = t1 - t0;
}

// candidate K1 inner loops: one thread = one channel, sample stream broadcast from shared memory
struct K1P { float a0, a1, a2, b1, b2, one, neg1; };
__global__ void k1_scalar(const float2 *samp, const float4 *lut, float2 *out, long long *cyc, int n, K1P P, unsigned dphi0) {
    extern __shared__ float4 sm[];
    float4 *slut = sm; float2 *ss = (float2 *)(sm + 257);
    for (int i = threadIdx.x; i < 257; i += blockDim.x) slut[i] = lut[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) ss[i] = samp[i];
    __syncthreads();
    unsigned phi = 0, dphi = dphi0 * (threadIdx.x + 1 + blockIdx.x);
    float xr1 = 0, xr2 = 0, xi1 = 0, xi2 = 0, yr1 = 0, yr2 = 0, yi1 = 0, yi2 = 0, accr = 0, acci = 0;
    long long t0 = clock64();
#pragma unroll 4
    for (int k = 0; k < n; k++) {
        float2 s = ss[k];
        unsigned idx = (phi >> 16) & 0xff;
        float fr = (float)(phi & 0xffff);
        float4 e = slut[idx];                    // {s0, c0, ds', dc'}
        float sn = __fadd_rn(e.x, __fmul_rn(e.z, fr));
        float cs = __fadd_rn(e.y, __fmul_rn(e.w, fr));
        phi += dphi;
        float re = __fadd_rn(__fmul_rn(s.x, cs), -__fmul_rn(s.y, sn));
        float im = __fadd_rn(__fmul_rn(s.y, cs), __fmul_rn(s.x, sn));
        float r = __fmul_rn(P.a0, re);
        r = __fadd_rn(r, __fadd_rn(__fmul_rn(P.a1, xr1), __fmul_rn(P.a2, xr2)));
        r = __fadd_rn(r, __fadd_rn(__fmul_rn(P.b1, yr1), __fmul_rn(P.b2, yr2)));
        xr2 = xr1; xr1 = re; yr2 = yr1; yr1 = r;
        float q = __fmul_rn(P.a0, im);
        q = __fadd_rn(q, __fadd_rn(__fmul_rn(P.a1, xi1), __fmul_rn(P.a2, xi2)));
        q = __fadd_rn(q, __fadd_rn(__fmul_rn(P.b1, yi1), __fmul_rn(P.b2, yi2)));
        xi2 = xi1; xi1 = im; yi2 = yi1; yi1 = q;
        if ((k % 20) == 19) { accr += r; acci += q; }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(accr, acci);
    if (threadIdx.x % 32 == 0) cyc[(blockIdx.x * blockDim.x + threadIdx.x) / 32] = t1 - t0;
}

__global__ void k1_packed(const float2 *samp, const float4 *lut, float2 *out, long long *cyc, int n, K1P P, unsigned dphi0) {
    extern __shared__ float4 sm[];
    float4 *slut = sm; float4 *ss = sm + 257;    // {re, im, im, re}
    for (int i = threadIdx.x; i < 257; i += blockDim.x) slut[i] = lut[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) { float2 v = samp[i]; ss[i] = make_float4(v.x, v.y, v.y, v.x); }
    __syncthreads();
    unsigned phi = 0, dphi = dphi0 * (threadIdx.x + 1 + blockIdx.x);
    u64 x1 = 0, x2 = 0, y1 = 0, y2 = 0, acc = 0;
    const u64 ONE = pack(P.one, P.one), SGN = pack(P.neg1, P.one);
    const u64 A0 = pack(P.a0, P.a0), A1 = pack(P.a1, P.a1), A2 = pack(P.a2, P.a2), B1 = pack(P.b1, P.b1), B2 = pack(P.b2, P.b2);
    long long t0 = clock64();
#pragma unroll 4
    for (int k = 0; k < n; k++) {
        float4 s = ss[k];
        u64 X = pack(s.x, s.y), XS = pack(s.z, s.w);
        unsigned idx = (phi >> 16) & 0xff;
        float fr = (float)(phi & 0xffff);
        float4 e = slut[idx];                    // {c0, s0, dc', ds'}
        u64 CS = fma2(mul2(pack(e.z, e.w), pack(fr, fr)), ONE, pack(e.x, e.y));
        phi += dphi;
        float c = lo(CS), sn = hi(CS);
        u64 Pm = mul2(X, pack(c, c));
        u64 Qm = mul2(XS, pack(sn, sn));
        u64 x0 = fma2(Qm, SGN, Pm);
        u64 t = fma2(mul2(A1, x1), ONE, mul2(A2, x2));
        u64 r = fma2(mul2(A0, x0), ONE, t);
        u64 u = fma2(mul2(B1, y1), ONE, mul2(B2, y2));
        u64 y0 = fma2(r, ONE, u);
        x2 = x1; x1 = x0; y2 = y1; y1 = y0;
        if ((k % 20) == 19) acc = fma2(acc, ONE, y0);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(lo(acc), hi(acc));
    if (threadIdx.x % 32 == 0) cyc[(blockIdx.x * blockDim.x + threadIdx.x) / 32] = t1 - t0;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template<int OP> void run(const char *name, int wps, float *out, long long *cyc, long long *hc) {
    int threads = 128 * wps, blocks = 148, iters = 4096;
    bench<OP><<<blocks, threads>>>(out, cyc, iters, 1.0000001f, 1e-9f);
    CK(cudaDeviceSynchronize());
    int nw = blocks * threads / 32;
    CK(cudaMemcpy(hc, cyc, nw * sizeof(long long), cudaMemcpyDeviceToHost));
    double mx = 0; for (int i = 0; i < nw; i++) if (hc[i] > mx) mx = hc[i];
    // per SMSP: wps warps each issue iters*NCH instructions in mx cycles
    printf("%-22s warps/SMSP=%d  cycles/warp-instr/SMSP = %.3f   (single-warp view: %.3f cyc/instr)\n", name, wps,
           mx / ((double)iters * NCH * wps), mx / ((double)iters * NCH));
}

int main() {
    float *out; long long *cyc; CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&cyc, 1 << 20));
    long long *hc = (long long *)malloc(1 << 20);
    for (int wps = 1; wps <= 4; wps *= 2) {
        run<0>("FMUL r,r,r", wps, out, cyc, hc); run<1>("FADD", wps, out, cyc, hc); run<2>("FFMA r*c+c", wps, out, cyc, hc);
        run<6>("FFMA 3 distinct regs", wps, out, cyc, hc);
        run<3>("FMUL2", wps, out, cyc, hc); run<4>("FADD2", wps, out, cyc, hc); run<5>("FFMA2 r*c+c", wps, out, cyc, hc);
        run<7>("FFMA2 3 distinct regs", wps, out, cyc, hc);
    }
    // K1 candidates
    int n = 8000; float2 *samp; float4 *lut; float2 *o2;
    CK(cudaMalloc(&samp, n * sizeof(float2))); CK(cudaMalloc(&lut, 257 * sizeof(float4))); CK(cudaMalloc(&o2, 1 << 22));
    float2 *hs = (float2 *)malloc(n * sizeof(float2)); for (int i = 0; i < n; i++) hs[i] = make_float2((rand() % 256 - 127.5f) / 127.5f, (rand() % 256 - 127.5f) / 127.5f);
    float4 *hl = (float4 *)malloc(257 * sizeof(float4)); for (int i = 0; i < 257; i++) hl[i] = make_float4(0.5f, 0.5f, 1e-6f, -1e-6f);
    CK(cudaMemcpy(samp, hs, n * sizeof(float2), cudaMemcpyHostToDevice)); CK(cudaMemcpy(lut, hl, 257 * sizeof(float4), cudaMemcpyHostToDevice));
    K1P P = {1.2888e-4f, 2.5776e-4f, 1.2888e-4f, 1.9692583f, -0.9697738f, 1.0f, -1.0f};
    for (int threads = 32; threads <= 512; threads *= 2) {
        for (int variant = 0; variant < 2; variant++) {
            size_t smem = 257 * 16 + (size_t)n * (variant ? 16 : 8);
            if (variant) { CK(cudaFuncSetAttribute(k1_packed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k1_packed<<<148, threads, smem>>>(samp, lut, o2, cyc, n, P, 40503u); }
            else { CK(cudaFuncSetAttribute(k1_scalar, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k1_scalar<<<148, threads, smem>>>(samp, lut, o2, cyc, n, P, 40503u); }
            CK(cudaDeviceSynchronize());
            int nw = 148 * threads / 32; CK(cudaMemcpy(hc, cyc, nw * sizeof(long long), cudaMemcpyDeviceToHost));
            double mx = 0; for (int i = 0; i < nw; i++) if (hc[i] > mx) mx = hc[i];
            printf("K1 %-7s threads/SM=%3d  cycles per sample-step (all warps of the SM) = %.2f  => per-SM ch-samples/cycle = %.3f\n",
                   variant ? "packed" : "scalar", threads, mx / n, threads / (mx / n));
        }
    }
    return 0;
}
