// mcmcn_peaks.cu -- FFMA-only and MUFU-only microbenchmarks.  MEASURED_PEAKS.json holds no
// FP32 / MUFU figure, and the step kernels are bound by exactly these two pipes
// (SURVEY.md section 8d), so the roofline denominators are measured in the same job.
#include <cuda_runtime.h>

#include "mcmcn_host.h"
#include "mcmcn_tc.cuh"

namespace mcmcn {

// The tensor pipe's own limit for the MMA the step kernel issues: back-to-back
// tcgen05.mma kind::tf32, M = 128, N = 208, K = 8, A in tensor memory, B in shared memory,
// two CTAs per SM.  Operand contents are irrelevant to the rate (zeros).
__global__ void __launch_bounds__(128) tf32_peak_kernel(int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned base_s;
    __shared__ unsigned long long mbar_s;
    const unsigned mb = smem_u32(&mbar_s);
    for (int i = threadIdx.x; i < 208 * 8; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.0f;
    if (threadIdx.x == 0) mbar_init(mb, 1);
    if (threadIdx.x < 32) tmem_alloc(&base_s, 256);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tbase = base_s;
    {
        unsigned z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st8(tbase + ((threadIdx.x >> 5) << 21) + 224, z);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        const unsigned long long bdesc = tc_smem_desc(smem_u32(smem), 128, 256);
        const unsigned idesc = tc_idesc(128, 208);
        for (int r = 0; r < reps; ++r) mma_tf32_ts(tbase, tbase + 224, bdesc, idesc, r > 0);
        mma_commit(mb);
    }
    mbar_wait(mb, 0);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_free(tbase, 256);
}

// 16 independent accumulators per thread, one live register per FFMA (a = a*x + y): the operand
// pattern that reaches the FP32 pipe's own limit (124.6 of 128 lanes/clk/SM on B200).  Patterns
// with two fresh register operands per FFMA (the step kernel's x*b[c]+r[c]) top out lower because
// of even/odd register-bank collisions: 98 lanes scalar, 117 with FFMA2 (tools/fma_probe.cu).
// The roofline denominator is the pipe's limit, not the pattern's.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float seed) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i;
    const float x = 1.0f - seed * 1e-7f, y = seed * 0.25f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) mufu_peak_kernel(float* out, int iters, float seed) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + 0.01f * (threadIdx.x & 7) + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678f) out[0] = s;
}

template <typename K>
static int time_kernel(K launch, double work_per_launch, double* out, cudaStream_t stream) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) launch();
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0, stream));
        launch();
        CK(cudaEventRecord(e1, stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = work_per_launch / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CK(cudaGetLastError());
    *out = best;
    return MCMCN_OK;
}

}  // namespace mcmcn

using namespace mcmcn;

extern "C" {

int mcmcn_peak_fp32(double* out_flops, void* stream_) {
    if (!out_flops) { set_error("null output"); return MCMCN_ERR_INVALID; }
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* d = nullptr;
    CK(cudaMalloc(&d, 16));
    const int iters = 4096, blocks = sms * 8, threads = 256;
    const double flops = 2.0 * 16 * 8 * (double)iters * threads * blocks;
    const int rc = time_kernel([&] { ffma_peak_kernel<<<blocks, threads, 0, stream>>>(d, iters, 1.0f); }, flops, out_flops, stream);
    cudaFree(d);
    return rc;
}

int mcmcn_peak_tf32(double* out_flops, void* stream_) {
    if (!out_flops) { set_error("null output"); return MCMCN_ERR_INVALID; }
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int reps = 8000, blocks = sms * 2;
    const double flops = 2.0 * 128 * 208 * 8 * (double)reps * blocks;
    return time_kernel([&] { tf32_peak_kernel<<<blocks, 128, 208 * 32, stream>>>(reps); }, flops, out_flops, stream);
}

int mcmcn_peak_mufu(double* out_ops, void* stream_) {
    if (!out_ops) { set_error("null output"); return MCMCN_ERR_INVALID; }
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* d = nullptr;
    CK(cudaMalloc(&d, 16));
    const int iters = 2048, blocks = sms * 8, threads = 256;
    const double ops = 8.0 * 8 * (double)iters * threads * blocks;
    const int rc = time_kernel([&] { mufu_peak_kernel<<<blocks, threads, 0, stream>>>(d, iters, -1.0f); }, ops, out_ops, stream);
    cudaFree(d);
    return rc;
}

}  // extern "C"
