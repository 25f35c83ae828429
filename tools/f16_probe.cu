// f16_probe.cu -- operand layout of tcgen05.mma kind::f16 with A in tensor memory (experiment behind the
// FP16-split variant of the step kernel): A[r][k] written as packed FP16 pairs into 8 TMEM columns,
// B[n][k] = (n == k) in shared memory (K-major, no swizzle), so D[r][n] = A[r][n] shows which K slot
// every half of every column is.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f16_probe f16_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr, unsigned lbo, unsigned sbo) {
    return (unsigned long long)((addr >> 4) & 0x3FFF) | ((unsigned long long)((lbo >> 4) & 0x3FFF) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__global__ void __launch_bounds__(128) probe(float* out, int variant) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned base_s;
    __shared__ unsigned long long mbar;
    const int tid = threadIdx.x, warp = tid >> 5;
    const unsigned sb = (smem_u32(smem) + 1023u) & ~1023u;
    __half* B = reinterpret_cast<__half*>(smem + (sb - smem_u32(smem)));
    for (int i = tid; i < 16 * 16; i += 128) {
        const int n = i / 16, k = i % 16;
        B[(n / 8) * 128 + (k / 8) * 64 + (n % 8) * 8 + (k % 8)] = __float2half(n == k ? 1.0f : 0.0f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&base_s)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = base_s, tlane = tbase + ((unsigned)warp << 21);
    // A: column c holds (low half = 2c + 1 + 0.5 * (tid & 1), high half = 100 + c): D at columns 0..15, A at 32..39
    unsigned a[8];
    for (int c = 0; c < 8; ++c) {
        const unsigned lo = __half_as_ushort(__float2half(2.0f * c + 1.0f + 0.5f * (tid & 1)));
        const unsigned hi = __half_as_ushort(__float2half(100.0f + c));
        a[c] = lo | (hi << 16);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tlane + 32), "r"(a[0]),
                 "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // variant 0: a_format = b_format = F16 (0); variant 1: BF16 ids (1) for comparison
        const unsigned fmt = variant == 1 ? 1u : 0u;
        const unsigned idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(16 >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        const unsigned long long bd = smem_desc(sb, 128, 256);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tbase),
                     "r"(tbase + 32), "l"(bd), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(tlane));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < 16; ++n) out[tid * 16 + n] = __uint_as_float(v[n]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(64u) : "memory");
}
int main() {
    float* d;
    CK(cudaMalloc(&d, 128 * 16 * 4));
    float h[128 * 16];
    for (int variant = 0; variant < 2; ++variant) {
        CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192));
        probe<<<1, 128, 8192>>>(d, variant);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        for (int r : {0, 1, 37, 127}) {
            printf("variant %d row %3d:", variant, r);
            for (int n = 0; n < 16; ++n) printf(" %g", h[r * 16 + n]);
            printf("\n");
        }
    }
    return 0;
}
