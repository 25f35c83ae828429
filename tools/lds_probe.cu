// lds_probe.cu -- cost of warp-uniform (broadcast) shared-memory loads on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

template <int VEC>   // VEC = 4: LDS.128, 2: LDS.64, 1: LDS.32
__global__ void __launch_bounds__(512) k(float* out, int iters) {
    __shared__ __align__(16) float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (float)i * 1e-3f;
    __syncthreads();
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const int w = threadIdx.x >> 5;
    int idx = w * 64;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int o = (idx + u * VEC) & 4095 & ~(VEC - 1);
            if (VEC == 4) { const float4 v = *reinterpret_cast<const float4*>(&sm[o]); acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w; }
            if (VEC == 2) { const float2 v = *reinterpret_cast<const float2*>(&sm[o]); acc0 += v.x; acc1 += v.y; }
            if (VEC == 1) { acc0 += sm[o]; }
        }
        idx += 16 * VEC;
    }
    if (acc0 + acc1 + acc2 + acc3 == 1234.5f) out[0] = acc0;
}

template <int VEC>
void run(const char* name, int threads) {
    float* d; cudaMalloc(&d, 16);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<VEC><<<sms, threads>>>(d, iters);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<VEC><<<sms, threads>>>(d, iters); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double loads = 16.0 * iters * (threads / 32);       // warp-level loads per SM
    const double cyc = best * 1e-3 * 1.965e9;
    printf("%-10s warps/SM=%2d  %.2f SM-cycles per warp-load, %.1f B/clk/SM delivered\n", name, threads / 32,
           cyc / loads, loads * 32 * VEC * 4 / cyc);
    cudaFree(d);
}

int main() {
    for (int t : {128, 256, 512}) { run<4>("LDS.128", t); run<2>("LDS.64", t); run<1>("LDS.32", t); }
    return 0;
}
