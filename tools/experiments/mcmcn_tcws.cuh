// mcmcn_tcws.cuh -- warp-specialised tcgen05 step kernel.
//
// Same arithmetic as sweep_tc_kernel (mcmcn_tc.cuh): residuals of 128 chains x one group as
// 3xTF32 MMAs into tensor memory, read back and squared by the chain's own thread.  What
// changes is who does the rest.  In sweep_tc_kernel every warp executes one long chain of
// dependent FP64 / Philox / MUFU instructions per sweep (one issue per ~8 clk) and tensor memory
// caps the SM at 16 such warps.  Here a CTA has eight warps:
//
//   owners  (warps 0-3, lane = chain = TMEM lane): write the proposal's column of the A operand,
//           issue the MMAs, read back + square the residuals, decide, store.
//   helpers (warps 4-7, same chain mapping): everything of a sweep that does not depend on
//           earlier decisions of the iteration -- state loads, Philox, Box-Muller, the FP64
//           proposal, both log-priors, the sigma terms -- up to two sweeps ahead, handed over
//           through a double-buffered shared-memory record.
//
// Hand-over: named barriers 1/2 = record[b] full (helpers arrive, owners sync), 3/4 = record[b]
// free (owners arrive, helpers sync); barrier 5 = the owners' rendezvous before an MMA issue.
// Still four CTAs per SM (128 TMEM columns each), now 32 warps.
#pragma once

#include "mcmcn_tc.cuh"

namespace mcmcn {

#define MCMCN_WS_THREADS 256
#define MCMCN_WS_FIELDS 6            /* 8-byte fields per chain in a record */
#define MCMCN_WS_RECORD_BYTES (MCMCN_WS_FIELDS * 8 * 128)
#define MCMCN_WS_ONES_BYTES 256      /* one 8-row group of the constant ones operand, reused with SBO = 0 */

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

enum { WS_PROP = 0, WS_U, WS_LPPROP, WS_LPCUR, WS_X0, WS_X1 };   // X0/X1: (wprop, wcur) floats | (m_prop, r_prop) of sigma

__device__ __forceinline__ void ws_issue_chunk(unsigned tbase, unsigned stage, unsigned ones, int np, int c, unsigned mbar) {
    const int row0 = c * MCMCN_TC_CH;
    const int nc = min(MCMCN_TC_CH, np - row0);
    const unsigned idesc = tc_idesc(128, nc);
    const unsigned slab = (unsigned)np * 32u;
    const unsigned base = stage + (unsigned)row0 * 32u;
    const unsigned d = tbase + MCMCN_TC_D;
    mma_tf32_ts(d, tbase + MCMCN_TC_A_HI, tc_smem_desc(base, 128, 256), idesc, 0);               // A_hi . X_hi
    mma_tf32_ts(d, tbase + MCMCN_TC_A_LO, tc_smem_desc(base, 128, 256), idesc, 1);               // A_lo . X_hi
    mma_tf32_ts(d, tbase + MCMCN_TC_A_HI, tc_smem_desc(base + slab, 128, 256), idesc, 1);        // A_hi . X_lo
    mma_tf32_ss(d, tc_smem_desc(ones, 128, 0), tc_smem_desc(base + 2 * slab, 128, 256), idesc, 1);   // 1 . NE (every row group = the same 256 bytes)
    mma_commit(mbar);
}

// grid = (group ranges, chain blocks of 128); block = 256 threads; dynamic shared memory =
// ones (256 B) + 2 records + 2 stages of a.tc_stage_bytes.
template <int F>
__global__ void __launch_bounds__(MCMCN_WS_THREADS, 4) sweep_tcws_kernel(const SweepArgs a) {
    constexpr bool GENERAL = F < 0;
    const bool partial = GENERAL ? (a.partial != 0) : ((F & MCMCN_F_PARTIAL) != 0);
    const bool count = GENERAL ? (a.count != 0) : ((F & MCMCN_F_COUNT) != 0);
    const bool replay = GENERAL && a.tape_z != nullptr;
    const bool trace = GENERAL && a.tr_ll != nullptr;
    const bool forced = GENERAL && a.tape_acc != nullptr;
    const bool override_lp = GENERAL && a.use_override != 0;
    const int P = a.P, K = a.P - 1;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned long long mbar_s[3];          // [0..1] TMA stage full, [2] accumulator full
    __shared__ unsigned tmem_base_s;

    const int nr = gridDim.x;
    const int g0 = (int)(((long long)a.G * blockIdx.x) / nr), g1 = (int)(((long long)a.G * (blockIdx.x + 1)) / nr);
    if (g0 >= g1) return;

    const int tid = threadIdx.x, warp = tid >> 5;
    const bool owner = warp < 4;
    const int t = tid & 127;                                           // chain slot of this thread within the CTA
    const unsigned ones = smem_u32(smem_raw);
    double* record = reinterpret_cast<double*>(smem_raw + MCMCN_WS_ONES_BYTES);
    const unsigned stage0 = ones + MCMCN_WS_ONES_BYTES + 2 * MCMCN_WS_RECORD_BYTES;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&mbar_s[i]), 1);
    }
    if (tid < MCMCN_WS_ONES_BYTES / 16) {   // rows (1, 1, 1, 0 | 0, 0, 0, 0): 8 rows x 16 bytes per K half
        const bool first_half = (tid >> 3) == 0;
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ones + 16u * tid), "f"(first_half ? 1.0f : 0.0f),
                     "f"(first_half ? 1.0f : 0.0f), "f"(first_half ? 1.0f : 0.0f), "f"(0.0f) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor core reads
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, MCMCN_TC_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tbase = tmem_base_s;
    const unsigned tlane = tbase + ((unsigned)(warp & 3) << 21);       // lane field = 32 * (warp % 4)
    const unsigned mb_tma0 = smem_u32(&mbar_s[0]), mb_mma = smem_u32(&mbar_s[2]);

    const int ch = blockIdx.y * 128 + t;
    const bool on = ch < a.n_chains;
    const int chl = min(ch, a.n_chains - 1);                           // lanes past the last chain redo its work, store nothing
    const size_t S = (size_t)a.S;

    if (!owner) {
        // ================================================================= helpers
        TcStash stash;
        stash.z = stash.u = 0.0;
        unsigned n = 0;                                                // sweep counter of this CTA
        for (int g = g0; g < g1; ++g) {
            const int R = a.group_nobs[g];
#pragma unroll 1
            for (int p = 0; p < P; ++p, ++n) {
                const TcInputs in = tc_fetch<GENERAL>(a, p, g, chl, partial, replay, override_lp, stash);
                const double prop = __dadd_rn(in.cur, __dmul_rn(in.sc, in.z));   // numpy.random.normal(value, sd), :304-306
                double lp_prop, lp_cur;
                if (partial) {
                    lp_prop = norm_logpdf_inv(prop, in.h_mu, in.h_isd, in.h_lsd);
                    lp_cur = (GENERAL && override_lp) ? in.lp_cur : norm_logpdf_inv(in.cur, in.h_mu, in.h_isd, in.h_lsd);
                } else {
                    lp_prop = prior_logpdf(a.prior[p], prop);
                    lp_cur = in.lp_cur;
                }
                double x0, x1;
                if (p == K) {                                          // LinReg::aux of the proposed sigma
                    const double sg = (double)(float)prop;
                    if (!(sg > 0.0)) {                                 // scipy: scale <= 0 -> nan
                        x0 = x1 = __longlong_as_double(0x7ff8000000000000LL);
                    } else {
                        const double inv = 1.0 / sg;
                        x0 = -0.5 * inv * inv;
                        x1 = (double)R * (log(sg) + MCMCN_LOG_SQRT_2PI);
                    }
                } else {                                               // centred FP32 coefficient: proposal, current
                    const float wprop = (float)__dsub_rn(prop, in.bbar), wcur = (float)__dsub_rn(in.cur, in.bbar);
                    x0 = __hiloint2double(__float_as_int(wcur), __float_as_int(wprop));
                    x1 = 0.0;
                }
                const int b = (int)(n & 1u);
                if (n >= 2u) named_sync(3 + b, MCMCN_WS_THREADS);      // owners are done with record[b]
                double* rec = record + (size_t)b * (MCMCN_WS_FIELDS * 128) + t;
                rec[WS_PROP * 128] = prop;
                rec[WS_U * 128] = in.u;
                rec[WS_LPPROP * 128] = lp_prop;
                rec[WS_LPCUR * 128] = lp_cur;
                rec[WS_X0 * 128] = x0;
                rec[WS_X1 * 128] = x1;
                __threadfence_block();
                named_arrive(1 + b, MCMCN_WS_THREADS);                 // record[b] full
            }
        }
    } else {
        // ================================================================= owners
        const float* tc = reinterpret_cast<const float*>(a.tc_data);
        auto stage_group = [&](int s, int g) {                         // thread 0 only
            const long long e0 = a.tc_group_off[g], e1 = a.tc_group_off[g + 1];
            const unsigned bytes = (unsigned)((e1 - e0) * 4);
            mbar_expect_tx(mb_tma0 + 8u * s, bytes);
            tma_bulk_g2s(stage0 + (unsigned)s * (unsigned)a.tc_stage_bytes, tc + e0, bytes, mb_tma0 + 8u * s);
        };
        if (tid == 0) {
            stage_group(0, g0);
            if (g0 + 1 < g1) stage_group(1, g0 + 1);
        }
        unsigned tma_phase = 0, mma_phase = 0, n = 0;
        for (int g = g0; g < g1; ++g) {
            const int s = (g - g0) & 1;
            const int R = a.group_nobs[g];
            const int np = max(16, (R + 15) & ~15);                    // padded observation count of the block
            const int nchunks = (np + MCMCN_TC_CH - 1) / MCMCN_TC_CH;
            const unsigned stage = stage0 + (unsigned)s * (unsigned)a.tc_stage_bytes;
            const double* bbar = a.obj_const + (size_t)g * K;
            if (g + 1 < g1) {                                          // next group's state: DRAM -> L2 meanwhile
                for (int k = 0; k < P; ++k)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.theta + ((size_t)k * a.G + g + 1) * S + chl));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ll + (size_t)(g + 1) * S + chl));
            }
            {   // A operand of the current state: centred coefficients (FP32), split hi / lo
                unsigned hi[8], lo[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float bk = k < K ? (float)__dsub_rn(a.theta[((size_t)k * a.G + g) * S + chl], bbar[k]) : 0.0f;
                    hi[k] = tf32_rn(bk);
                    lo[k] = tf32_rn(bk - __uint_as_float(hi[k]));
                }
                tmem_st8(tlane + MCMCN_TC_A_HI, hi);
                tmem_st8(tlane + MCMCN_TC_A_LO, lo);
            }
            double aux_m, aux_r;                                       // LinReg::Aux of the current sigma
            {
                const double sg = (double)(float)a.theta[((size_t)K * a.G + g) * S + chl];
                if (!(sg > 0.0)) {
                    aux_m = aux_r = __longlong_as_double(0x7ff8000000000000LL);
                } else {
                    const double inv = 1.0 / sg;
                    aux_m = -0.5 * inv * inv;
                    aux_r = (double)R * (log(sg) + MCMCN_LOG_SQRT_2PI);
                }
            }
            double ll_cur = a.ll[(size_t)g * S + chl];
            mbar_wait(mb_tma0 + 8u * s, (tma_phase >> s) & 1u);
            tma_phase ^= 1u << s;

#pragma unroll 1
            for (int p = 0; p < P; ++p, ++n) {
                const size_t at = ((size_t)p * a.G + g) * S + chl;
                const bool is_sigma = p == K;
                const int b = (int)(n & 1u);
                const double* rec = record + (size_t)b * (MCMCN_WS_FIELDS * 128) + t;
                named_sync(1 + b, MCMCN_WS_THREADS);                   // record[b] full
                const double x0 = rec[WS_X0 * 128];
                const float wprop = __int_as_float(__double2loint(x0)), wcur = __int_as_float(__double2hiint(x0));
                if (!is_sigma) {                                       // column p of the A operand <- the proposal
                    const unsigned h = tf32_rn(wprop);
                    tmem_st1(tlane + MCMCN_TC_A_HI + p, h);
                    tmem_st1(tlane + MCMCN_TC_A_LO + p, tf32_rn(wprop - __uint_as_float(h)));
                }
                tmem_wait_st();
                tc_fence_before();
                named_sync(5, 128);
                if (tid == 0) {
                    tc_fence_after();
                    ws_issue_chunk(tbase, stage, ones, np, 0, mb_mma);
                }
                double acc = 0.0;
                for (int c = 0; c < nchunks; ++c) {
                    mbar_wait(mb_mma, mma_phase);
                    mma_phase ^= 1u;
                    tc_fence_after();
                    const int nc = min(MCMCN_TC_CH, np - c * MCMCN_TC_CH);
                    acc += tc_sum_squares(tlane + MCMCN_TC_D, nc >> 4);
                    if (c + 1 < nchunks) {                             // the accumulator is free once every lane has read it
                        tc_fence_before();
                        named_sync(5, 128);
                        if (tid == 0) {
                            tc_fence_after();
                            ws_issue_chunk(tbase, stage, ones, np, c + 1, mb_mma);
                        }
                    }
                }

                // Parameter.step decision tree, :334-367
                const double prop = rec[WS_PROP * 128], u = rec[WS_U * 128];
                const double lp_prop = rec[WS_LPPROP * 128], lp_cur = rec[WS_LPCUR * 128];
                const double m_prop = is_sigma ? x0 : aux_m;
                const double r_prop = is_sigma ? rec[WS_X1 * 128] : aux_r;
                named_arrive(3 + b, MCMCN_WS_THREADS);                 // record[b] free (all of it is in registers now)
                const double llp = acc * m_prop - r_prop;
                const double post_prop = lp_prop + llp;
                const double post_cur = lp_cur + ll_cur;
                const double diff = post_prop - post_cur;
                const bool b1 = !finite64(post_cur) && finite64(post_prop);
                const bool test = finite64(llp) && finite64(diff);     // branches 4/5 draw the uniform
                const int fast = log_u_vs_diff_fast(u, diff);
                bool accept = b1 || (test && fast > 0);
                if (!b1 && test && fast == 0) accept = log(u) < diff;  // rare: within 1e-6 of the threshold
                if (GENERAL) {
                    if (trace && on) {
                        a.tr_ll[at] = llp;
                        a.tr_lp[at] = lp_prop;
                        a.tr_diff[at] = diff;
                        a.tr_acc[at] = accept ? 1 : 0;
                    }
                    if (forced) accept = a.tape_acc[at] != 0;
                }
                if (accept) {                                          // :369-378, :608-610
                    if (on) {
                        a.theta[at] = prop;
                        if (!partial) a.lprior[at] = lp_prop;
                    }
                    ll_cur = llp;
                    aux_m = m_prop;
                    aux_r = r_prop;
                }
                if (!is_sigma) {                                       // column p <- the value the chain keeps
                    const float wkeep = accept ? wprop : wcur;
                    const unsigned h = tf32_rn(wkeep);
                    tmem_st1(tlane + MCMCN_TC_A_HI + p, h);
                    tmem_st1(tlane + MCMCN_TC_A_LO + p, tf32_rn(wkeep - __uint_as_float(h)));
                }
                if (count && on) {
                    unsigned cnt = a.counts[at];
                    cnt += accept ? 1u : 0x10000u;
                    if (a.tune) {                                      // Parameter.tune, :385-437
                        const unsigned na = cnt & 0xFFFFu, nrj = cnt >> 16;
                        if (na + nrj) {
                            const double sc = a.scale[at];
                            const double rate = (double)na / (double)(na + nrj);
                            double f = 1.0;
                            if (rate < 0.001) f = 0.1;
                            else if (rate < 0.05) f = 0.5;
                            else if (rate < 0.2) f = 0.9;
                            else if (rate > 0.95) f = 10.0;
                            else if (rate > 0.75) f = 2.0;
                            else if (rate > 0.5) f = 1.1;
                            double ns = __dmul_rn(sc, f);
                            if (ns == 0.0) ns = sc;
                            a.scale[at] = ns;
                            cnt = 0;
                        }
                    }
                    a.counts[at] = cnt;
                }
            }
            if (on) a.ll[(size_t)g * S + chl] = ll_cur;
            // every MMA that read this stage has completed (all owners waited on its mbarrier)
            if (tid == 0 && g + 2 < g1) stage_group(s, g + 2);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tbase, MCMCN_TC_COLS);
}

}  // namespace mcmcn
