// mcmcn_diag.cu -- batched reduction kernels behind sampleDiagnosis.diagnoseSamples
// (/root/reference/sampleDiagnosis.py:158-255, :419-427, :766-776), all FP64 (1e-10 parity bar).
//
// Input layout: x[key][half-chain j][draw i], contiguous in i.  The reference loops
// keys x m x n^2 in pure Python (:189-208); here every (key, half-chain) or (key, lag)
// pair is one warp / one thread and every sum is taken in a fixed order, so results do
// not depend on the launch geometry.
#include <cuda_runtime.h>

#include <cub/device/device_segmented_sort.cuh>

#include "mcmcn_host.h"

namespace mcmcn {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp per (key, half-chain): mean, then sum of squared deviations (two passes,
// like numpy.var), variance with ddof=1 (:164-166, :174-176).
__global__ void moments_kernel(const double* __restrict__ x, long long rows, int n, double* __restrict__ mean,
                               double* __restrict__ var) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const double* p = x + row * (long long)n;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += p[i];
    const double mu = warp_sum(s) / (double)n;
    double q = 0.0;
    for (int i = lane; i < n; i += 32) { const double d = p[i] - mu; q = fma(d, d, q); }
    q = warp_sum(q);
    if (lane == 0) {
        mean[row] = mu;
        var[row] = q / (double)(n - 1);
    }
}

// Variogram numerators (:189-194): out[key][t] = sum_j sum_{i=t}^{n-1} (x[j][i]-x[j][i-t])^2.
// Block = one key x one slab of lags x one range of half-chains; the half-chains are staged through
// shared memory one at a time and accumulated in half-chain order.  Thread `l` owns lags t0+l and, to
// balance the triangular work, n-1-(t0+l) (lags are paired from both ends).  With several half-chain
// ranges (gridDim.z > 1: a slab of few keys would otherwise leave most SMs idle -- 256 blocks for 128 keys,
// 11 % of the warp slots, profiles/r2_variogram_kernel_ncu_summary.txt) the per-range sums go to
// part[key][range][t] and variogram_fold_kernel adds them in range order: a fixed summation order.
__global__ void variogram_kernel(const double* __restrict__ x, int m, int n, double* __restrict__ out) {
    extern __shared__ double row[];
    const long long key = blockIdx.x;
    const int half = (n + 1) / 2;                       // pairs (t, n-1-t), t < half
    const int t = blockIdx.y * blockDim.x + threadIdx.x;
    const int ta = t, tb = n - 1 - t;
    const bool on = t < half;
    const int nz = gridDim.z, z = blockIdx.z;
    const int j0 = (int)(((long long)m * z) / nz), j1 = (int)(((long long)m * (z + 1)) / nz);
    double sa = 0.0, sb = 0.0;
    for (int j = j0; j < j1; ++j) {
        const double* p = x + (key * m + j) * (long long)n;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) row[i] = p[i];
        __syncthreads();
        if (on) {
            double a = 0.0;
            for (int i = ta; i < n; ++i) { const double d = row[i] - row[i - ta]; a = fma(d, d, a); }
            sa += a;
            if (tb != ta) {
                double b = 0.0;
                for (int i = tb; i < n; ++i) { const double d = row[i] - row[i - tb]; b = fma(d, d, b); }
                sb += b;
            }
        }
    }
    if (on) {
        double* o = out + (key * nz + z) * (long long)n;
        o[ta] = sa;
        if (tb != ta) o[tb] = sb;
    }
}
__global__ void variogram_fold_kernel(const double* __restrict__ part, int nz, int n, long long total, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // key * n + t
    if (i >= total) return;
    const long long key = i / n, t = i - key * n;
    const double* p = part + key * nz * (long long)n + t;
    double s = 0.0;
    for (int z = 0; z < nz; ++z) s += p[(long long)z * n];
    out[i] = s;
}

// Median (numpy.median) and shortest interval of `gap` order statistics (:766-776) of one
// sorted key per block; ties resolve to the smallest index like numpy.where(tmp == min)[0][0].
__global__ void median_hdi_kernel(const double* __restrict__ s, long long len, long long gap, double* __restrict__ out) {
    __shared__ double wbest[32];
    __shared__ long long ibest[32];
    const double* p = s + (long long)blockIdx.x * len;
    double w = __longlong_as_double(0x7ff0000000000000LL);
    long long idx = 0x7fffffffffffffffLL;
    for (long long i = threadIdx.x; i < len - gap; i += blockDim.x) {
        const double d = p[i + gap] - p[i];
        if (d < w || (d == w && i < idx)) { w = d; idx = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w2 = __shfl_xor_sync(0xffffffffu, w, o);
        const long long i2 = __shfl_xor_sync(0xffffffffu, idx, o);
        if (w2 < w || (w2 == w && i2 < idx)) { w = w2; idx = i2; }
    }
    if ((threadIdx.x & 31) == 0) { wbest[threadIdx.x >> 5] = w; ibest[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
            if (wbest[k] < w || (wbest[k] == w && ibest[k] < idx)) { w = wbest[k]; idx = ibest[k]; }
        double* o = out + (long long)blockIdx.x * 3;
        o[0] = (len & 1) ? p[len / 2] : (p[len / 2 - 1] + p[len / 2]) / 2.0;
        o[1] = p[idx];
        o[2] = p[idx + gap];
    }
}

// B, W, vhat, rhat of one key from its half-chain means and variances (:158-187, :216-224).
// One warp per key; sums over the m half-chains are taken lane-strided then shuffled.
__global__ void rhat_kernel(const double* __restrict__ mean, const double* __restrict__ var, long long n_keys, int m,
                            int n, double* __restrict__ out) {
    const long long key = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (key >= n_keys) return;
    const int lane = threadIdx.x & 31;
    const double* mu = mean + key * m;
    const double* vr = var + key * m;
    double s = 0.0, w = 0.0;
    for (int j = lane; j < m; j += 32) { s += mu[j]; w += vr[j]; }
    const double grand = warp_sum(s) / (double)m;
    const double W = warp_sum(w) / (double)m;                        // numpy.mean(numpy.var(rows, ddof=1))
    double q = 0.0;
    for (int j = lane; j < m; j += 32) { const double d = mu[j] - grand; q = fma(d, d, q); }
    const double B = (double)n * (warp_sum(q) / (double)(m - 1));    // n * numpy.var(row means, ddof=1)
    if (lane == 0) {
        const double vhat = W * (double)(n - 1) / (double)n + B / (double)n;
        double* o = out + key * 4;
        o[0] = B; o[1] = W; o[2] = vhat; o[3] = sqrt(vhat / W);
    }
}

// rho_t = 1 - V_t / (2 vhat) for every lag (:196-208), truncation at the first even t with
// rho_{t+1} + rho_{t+2} < 0 (:241-251) and ESS = m n / (1 + 2 sum_{t<=T} rho_t) with the sum
// starting at lag 0 (:253-255).  One block per key; rho is staged in shared memory.
__global__ void ess_kernel(const double* __restrict__ vnum, const double* __restrict__ rh, int m, int n,
                           double* __restrict__ rho_out, double* __restrict__ ess) {
    extern __shared__ double rho[];
    const long long key = blockIdx.x;
    const double vhat = rh[key * 4 + 2];
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const double V = vnum[key * n + t] / ((double)m * (double)(n - t));
        const double r = 1.0 - V / (2.0 * vhat);
        rho[t] = r;
        if (rho_out) rho_out[key * n + t] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int T = n - 1;
        for (int t = 0; t < n - 2; t += 2)
            if (rho[t + 1] + rho[t + 2] < 0.0) { T = t; break; }
        double s = 0.0;
        for (int t = 0; t <= T; ++t) s += rho[t];
        ess[key] = ((double)m * (double)n) / (1.0 + 2.0 * s);
    }
}

// mean of each contiguous row of `len` doubles (Summary, :466-470); one warp per row
__global__ void row_mean_kernel(const double* __restrict__ x, long long rows, long long len, double* __restrict__ out) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const double* p = x + row * len;
    double s = 0.0;
    for (long long i = lane; i < len; i += 32) s += p[i];
    s = warp_sum(s);
    if (lane == 0) out[row] = s / (double)len;
}

// Half-chains of a slab of columns straight out of a sample store ([rows][ncol][S], chain fastest,
// FP32 or FP64): out[k][j0 + 2 c + h][i] = store[h n + i][k0 + k][c] as doubles (sampleDiagnosis.py:118-156:
// every chain's rows split into first / second half); out holds out_m half-chains per column.  32 x 32 tiles through shared memory: reads are
// coalesced along the chains, writes along the draws.  grid = (row tiles, chain tiles, columns).
template <typename TS>
__global__ void halfchains_kernel(const TS* __restrict__ store, int n, long long ncol, long long S, long long k0,
                                  int n_chains, int out_m, int out_j0, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int k = blockIdx.z;
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int two_n = 2 * n;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int r = r0 + dy, c = c0 + threadIdx.x;
        if (r < two_n && c < n_chains) tile[dy][threadIdx.x] = (double)store[((long long)r * ncol + k0 + k) * S + c];
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int c = c0 + dy, r = r0 + threadIdx.x;
        if (r < two_n && c < n_chains) {
            const int h = r >= n ? 1 : 0, i = r - h * n;
            out[((long long)k * out_m + out_j0 + 2 * c + h) * n + i] = tile[threadIdx.x][dy];
        }
    }
}

}  // namespace mcmcn

using namespace mcmcn;

extern "C" {

int mcmcn_diag_halfchains(const void* store, int32_t store_dtype, int32_t n, int64_t ncol, int64_t stride, int64_t k0,
                          int64_t n_keys, int32_t n_chains, int32_t out_m, int32_t out_j0, double* out, void* stream) {
    if (!store || !out || n < 1 || n_keys < 1 || n_keys > 65535 || k0 < 0 || k0 + n_keys > ncol || n_chains < 1 || n_chains > stride ||
        out_j0 < 0 || out_j0 + 2 * n_chains > out_m || (store_dtype != 32 && store_dtype != 64)) {
        set_error("bad diag_halfchains args");
        return MCMCN_ERR_INVALID;
    }
    const dim3 grid((unsigned)((2 * n + 31) / 32), (unsigned)((n_chains + 31) / 32), (unsigned)n_keys);
    if (grid.y > 65535) { set_error("too many chains for one halfchains launch"); return MCMCN_ERR_UNSUPPORTED; }
    const dim3 block(32, 8, 1);
    if (store_dtype == 32)
        halfchains_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)store, n, ncol, stride, k0, n_chains, out_m, out_j0, out);
    else
        halfchains_kernel<double><<<grid, block, 0, (cudaStream_t)stream>>>((const double*)store, n, ncol, stride, k0, n_chains, out_m, out_j0, out);
    CK(cudaGetLastError());
    return MCMCN_OK;
}

int mcmcn_diag_moments(const double* x, int64_t n_keys, int32_t m, int32_t n, double* out_mean, double* out_var,
                       void* stream) {
    if (!x || !out_mean || !out_var || n_keys < 1 || m < 1 || n < 2) { set_error("bad diag_moments args"); return MCMCN_ERR_INVALID; }
    const long long rows = (long long)n_keys * m;
    const int wpb = 8;
    moments_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(x, rows, n, out_mean, out_var);
    CK(cudaGetLastError());
    return MCMCN_OK;
}

int mcmcn_diag_variogram(const double* x, int64_t n_keys, int32_t m, int32_t n, double* out, void* stream_) {
    if (!x || !out || n_keys < 1 || m < 1 || n < 2) { set_error("bad diag_variogram args"); return MCMCN_ERR_INVALID; }
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t smem = sizeof(double) * (size_t)n;
    if (smem > 200 * 1024) { set_error("n=%d draws per half-chain exceed the shared-memory row buffer", n); return MCMCN_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(variogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int half = (n + 1) / 2;
    const int threads = half < 128 ? ((half + 31) & ~31) : 128;
    const unsigned lag_slabs = (unsigned)((half + threads - 1) / threads);
    // ranges of half-chains: a function of m alone, so that a key's sums do not depend on how many keys share the
    // launch (slabs of any size give bit-identical results): 64 half-chains per range, at most 16 ranges
    int nz = m / 64;
    if (nz < 1) nz = 1;
    if (nz > 16) nz = 16;
    const dim3 grid((unsigned)n_keys, lag_slabs, (unsigned)nz);
    if (nz == 1) {
        variogram_kernel<<<grid, threads, smem, stream>>>(x, m, n, out);
        CK(cudaGetLastError());
        return MCMCN_OK;
    }
    double* part = nullptr;
    const long long total = (long long)n_keys * n;
    CK(cudaMallocAsync((void**)&part, sizeof(double) * (size_t)total * nz, stream));
    variogram_kernel<<<grid, threads, smem, stream>>>(x, m, n, part);
    variogram_fold_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(part, nz, n, total, out);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(part, stream);
    CK(e);
    return MCMCN_OK;
}

int mcmcn_diag_rhat(const double* mean, const double* var, int64_t n_keys, int32_t m, int32_t n, double* out, void* stream) {
    if (!mean || !var || !out || n_keys < 1 || m < 2 || n < 2) { set_error("bad diag_rhat args"); return MCMCN_ERR_INVALID; }
    const int wpb = 8;
    rhat_kernel<<<(unsigned)((n_keys + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(mean, var, n_keys, m, n, out);
    CK(cudaGetLastError());
    return MCMCN_OK;
}

int mcmcn_diag_ess(const double* vario, const double* rhat4, int64_t n_keys, int32_t m, int32_t n, double* rho_out,
                   double* ess_out, void* stream) {
    if (!vario || !rhat4 || !ess_out || n_keys < 1 || m < 1 || n < 3) { set_error("bad diag_ess args"); return MCMCN_ERR_INVALID; }
    const size_t smem = sizeof(double) * (size_t)n;
    if (smem > 200 * 1024) { set_error("n=%d draws per half-chain exceed the shared-memory buffer", n); return MCMCN_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(ess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ess_kernel<<<(unsigned)n_keys, 128, smem, (cudaStream_t)stream>>>(vario, rhat4, m, n, rho_out, ess_out);
    CK(cudaGetLastError());
    return MCMCN_OK;
}

int mcmcn_diag_row_mean(const double* x, int64_t rows, int64_t len, double* out, void* stream) {
    if (!x || !out || rows < 1 || len < 1) { set_error("bad diag_row_mean args"); return MCMCN_ERR_INVALID; }
    const int wpb = 8;
    row_mean_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(x, rows, len, out);
    CK(cudaGetLastError());
    return MCMCN_OK;
}

int mcmcn_diag_median_hdi(const double* sorted, int64_t n_keys, int64_t len, int64_t gap, double* out, void* stream) {
    if (!sorted || !out || n_keys < 1 || len < 2 || gap < 1 || gap > len - 1) { set_error("bad diag_median_hdi args"); return MCMCN_ERR_INVALID; }
    median_hdi_kernel<<<(unsigned)n_keys, 256, 0, (cudaStream_t)stream>>>(sorted, len, gap, out);
    CK(cudaGetLastError());
    return MCMCN_OK;
}

int mcmcn_diag_sort_keys(double* x, int64_t n_keys, int64_t len, void* stream_) {
    if (!x || n_keys < 1 || len < 1) { set_error("bad diag_sort args"); return MCMCN_ERR_INVALID; }
    if (n_keys * len > (int64_t)0x7fffffff) { set_error("sort of more than 2^31 items not supported"); return MCMCN_ERR_UNSUPPORTED; }
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n_items = (int)(n_keys * len);
    const int n_seg = (int)n_keys;
    // segment offsets k*len, generated on the host (n_keys is small next to the data)
    long long* off_h = new long long[n_seg + 1];
    for (int k = 0; k <= n_seg; ++k) off_h[k] = (long long)k * len;
    long long* off_d = nullptr;
    double* tmp = nullptr;
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    int rc = MCMCN_OK;
    cudaError_t e = cudaMalloc(&off_d, sizeof(long long) * (n_seg + 1));
    if (e == cudaSuccess) e = cudaMemcpyAsync(off_d, off_h, sizeof(long long) * (n_seg + 1), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, sizeof(double) * (size_t)n_items);
    if (e == cudaSuccess)
        e = cub::DeviceSegmentedSort::SortKeys(nullptr, scratch_bytes, x, tmp, n_items, n_seg, off_d, off_d + 1, stream);
    if (e == cudaSuccess) e = cudaMalloc(&scratch, scratch_bytes ? scratch_bytes : 16);
    if (e == cudaSuccess)
        e = cub::DeviceSegmentedSort::SortKeys(scratch, scratch_bytes, x, tmp, n_items, n_seg, off_d, off_d + 1, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(x, tmp, sizeof(double) * (size_t)n_items, cudaMemcpyDeviceToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { set_error("diag_sort: %s", cudaGetErrorString(e)); rc = MCMCN_ERR_CUDA; }
    cudaFree(scratch);
    cudaFree(tmp);
    cudaFree(off_d);
    delete[] off_h;
    return rc;
}

}  // extern "C"
