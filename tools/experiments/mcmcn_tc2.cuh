// mcmcn_tc2.cuh -- the tcgen05 step kernel with two chain blocks sharing one accumulator.
//
// sweep_tc_kernel (mcmcn_tc.cuh) is bound by parallelism, not by a pipe: each warp executes one
// chain of dependent instructions per sweep, tensor memory caps the SM at four accumulator tiles
// (4 x 128 columns) and with one tile per four warps that is 16 warps, at half the issue slots.
// But a tile is in use only from the MMA issue to the end of the read-back -- about half of a
// sweep; the rest (decision, stores, next proposal, log-priors, Philox) needs no tensor memory.
//
// Here a CTA has eight warps = two blocks of 128 chains, A (warps 0-3) and B (warps 4-7), working
// on the same group (one copy of the observation block in shared memory) and taking turns on ONE
// accumulator tile: while A's MMAs run and A reads them back, B does its scalar work, and vice
// versa.  Same four CTAs per SM, now 32 warps.
//
//   tensor memory (128 columns): accumulator 112 | block A's operand: A_hi 8, A_lo 8
//   shared memory:               block B's operand (A from shared memory: tcgen05.mma SS form),
//                                the ones operand (256 bytes, row-group stride 0), two TMA stages
//   named barriers: 1 / 2 = rendezvous of block A / B before an MMA issue;
//                   3 = accumulator handed to B (A arrives, B syncs), 4 = handed to A.
// Arithmetic, update order, decision tree and random streams are those of sweep_tc_kernel.
#pragma once

#include "mcmcn_tc.cuh"

namespace mcmcn {

#ifndef MCMCN_TC2_CTAS
#define MCMCN_TC2_CTAS 3            /* CTAs per SM the register allocation allows: 3 -> 85 registers per thread */
#endif
#define MCMCN_TC2_THREADS 256
#define MCMCN_TC2_ONES_BYTES 256
#define MCMCN_TC2_BOP_BYTES 8192     /* block B's A operand: [128][8] hi, then [128][8] lo, K-major core matrices */

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void sts32(unsigned addr, unsigned v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// The 4 MMAs of observation chunk c; the chain operand comes from tensor memory (block A) or from
// shared memory (block B).
__device__ __forceinline__ void tc2_issue_chunk(bool block_b, unsigned tbase, unsigned bop, unsigned stage, unsigned ones, int np,
                                                int c, unsigned mbar) {
    const int row0 = c * MCMCN_TC_CH;
    const int nc = min(MCMCN_TC_CH, np - row0);
    const unsigned idesc = tc_idesc(128, nc);
    const unsigned slab = (unsigned)np * 32u;
    const unsigned base = stage + (unsigned)row0 * 32u;
    const unsigned d = tbase + MCMCN_TC_D;
    const unsigned long long x_hi = tc_smem_desc(base, 128, 256), x_lo = tc_smem_desc(base + slab, 128, 256);
    if (block_b) {
        const unsigned long long a_hi = tc_smem_desc(bop, 128, 256), a_lo = tc_smem_desc(bop + 4096, 128, 256);
        mma_tf32_ss(d, a_hi, x_hi, idesc, 0);                                                    // A_hi . X_hi
        mma_tf32_ss(d, a_lo, x_hi, idesc, 1);                                                    // A_lo . X_hi
        mma_tf32_ss(d, a_hi, x_lo, idesc, 1);                                                    // A_hi . X_lo
    } else {
        mma_tf32_ts(d, tbase + MCMCN_TC_A_HI, x_hi, idesc, 0);
        mma_tf32_ts(d, tbase + MCMCN_TC_A_LO, x_hi, idesc, 1);
        mma_tf32_ts(d, tbase + MCMCN_TC_A_HI, x_lo, idesc, 1);
    }
    mma_tf32_ss(d, tc_smem_desc(ones, 128, 0), tc_smem_desc(base + 2 * slab, 128, 256), idesc, 1);   // 1 . NE
    mma_commit(mbar);
}

// grid = (group ranges, chain blocks of 256); block = 256 threads; dynamic shared memory =
// ones (256 B) + block B's operand (8 KB) + 2 stages of a.tc_stage_bytes.
template <int F>
__global__ void __launch_bounds__(MCMCN_TC2_THREADS, MCMCN_TC2_CTAS) sweep_tc2_kernel(const SweepArgs a) {
    constexpr bool GENERAL = F < 0;
    const bool partial = GENERAL ? (a.partial != 0) : ((F & MCMCN_F_PARTIAL) != 0);
    const bool count = GENERAL ? (a.count != 0) : ((F & MCMCN_F_COUNT) != 0);
    const bool replay = GENERAL && a.tape_z != nullptr;
    const bool trace = GENERAL && a.tr_ll != nullptr;
    const bool forced = GENERAL && a.tape_acc != nullptr;
    const bool override_lp = GENERAL && a.use_override != 0;
    const int P = a.P, K = a.P - 1;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ unsigned long long mbar_s[4];          // [0..1] TMA stage full, [2] / [3] MMAs of block A / B done
    __shared__ unsigned tmem_base_s;

    const int nr = gridDim.x;
    const int g0 = (int)(((long long)a.G * blockIdx.x) / nr), g1 = (int)(((long long)a.G * (blockIdx.x + 1)) / nr);
    if (g0 >= g1) return;

    const int tid = threadIdx.x, warp = tid >> 5;
    const bool block_b = warp >= 4;
    const int t = tid & 127;                                           // chain slot = TMEM lane of this thread
    const unsigned ones = smem_u32(smem_raw);
    const unsigned bop = ones + MCMCN_TC2_ONES_BYTES;
    const unsigned stage0 = bop + MCMCN_TC2_BOP_BYTES;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&mbar_s[i]), 1);
    }
    if (tid < MCMCN_TC2_ONES_BYTES / 16) {   // rows (1, 1, 1, 0 | 0, 0, 0, 0): 8 rows x 16 bytes per K half
        const float one = (tid >> 3) == 0 ? 1.0f : 0.0f;
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ones + 16u * tid), "f"(one), "f"(one), "f"(one), "f"(0.0f) : "memory");
        fence_async_smem();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, MCMCN_TC_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tbase = tmem_base_s;
    const unsigned tlane = tbase + ((unsigned)(warp & 3) << 21);       // lane field = 32 * (warp % 4)
    const unsigned mb_tma0 = smem_u32(&mbar_s[0]);
    const unsigned mb_mma = smem_u32(&mbar_s[block_b ? 3 : 2]);
    // this thread's row of block B's shared-memory operand: element k at row_b + (k / 4) * 128 + (k % 4) * 4
    const unsigned row_b = bop + (unsigned)(t >> 3) * 256u + (unsigned)(t & 7) * 16u;
    const int bar_own = block_b ? 2 : 1;                               // rendezvous of this block
    const int bar_take = block_b ? 3 : 4;                              // accumulator handed to this block
    const int bar_give = block_b ? 4 : 3;                              // ... handed to the other block

    const float* tc = reinterpret_cast<const float*>(a.tc_data);
    auto stage_group = [&](int s, int g) {                             // one thread
        const long long e0 = a.tc_group_off[g], e1 = a.tc_group_off[g + 1];
        const unsigned bytes = (unsigned)((e1 - e0) * 4);
        mbar_expect_tx(mb_tma0 + 8u * s, bytes);
        tma_bulk_g2s(stage0 + (unsigned)s * (unsigned)a.tc_stage_bytes, tc + e0, bytes, mb_tma0 + 8u * s);
    };
    if (tid == 0) {
        stage_group(0, g0);
        if (g0 + 1 < g1) stage_group(1, g0 + 1);
    }

    const int ch = blockIdx.y * MCMCN_TC2_THREADS + tid;              // block A: first 128 chains of the pair, B: the next 128
    const bool on = ch < a.n_chains;
    const int chl = min(ch, a.n_chains - 1);                           // lanes past the last chain redo its work, store nothing
    const size_t S = (size_t)a.S;
    unsigned tma_phase = 0, mma_phase = 0, n = 0;
    TcStash stash;
    stash.z = stash.u = 0.0;

    // column p of this chain's operand <- v (split hi / lo)
    auto put_column = [&](int p, float v) {
        const unsigned h = tf32_rn(v), l = tf32_rn(v - __uint_as_float(h));
        if (block_b) {
            const unsigned off = (unsigned)(p >> 2) * 128u + (unsigned)(p & 3) * 4u;
            sts32(row_b + off, h);
            sts32(row_b + 4096u + off, l);
        } else {
            tmem_st1(tlane + MCMCN_TC_A_HI + p, h);
            tmem_st1(tlane + MCMCN_TC_A_LO + p, l);
        }
    };
    // make this thread's operand writes visible to the tensor core, before a barrier
    auto publish = [&]() {
        if (block_b) fence_async_smem();
        else tmem_wait_st();
        tc_fence_before();
    };

    for (int g = g0; g < g1; ++g) {
        const int s = (g - g0) & 1;
        const int R = a.group_nobs[g];
        const int np = max(16, (R + 15) & ~15);                        // padded observation count of the block
        const int nchunks = (np + MCMCN_TC_CH - 1) / MCMCN_TC_CH;
        const unsigned stage = stage0 + (unsigned)s * (unsigned)a.tc_stage_bytes;
        const double* bbar = a.obj_const + (size_t)g * K;

        const size_t GS = (size_t)a.G * S;
        size_t at = (size_t)g * S + chl, hy = (size_t)chl, bb = (size_t)g * K;   // sweep 0; bumped by GS / S / 1 per sweep
        TcInputs in;
        tc_fetch_state<GENERAL>(in, a, 0, at, hy, bb, partial, override_lp);
        tc_fetch_random<GENERAL>(in, a, 0, g, chl, at, replay, stash);
        if (g + 1 < g1) {                                              // next group's state: DRAM -> L2 meanwhile
            for (int k = 0; k < P; ++k)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.theta + ((size_t)k * a.G + g + 1) * S + chl));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ll + (size_t)(g + 1) * S + chl));
        }
        // operand of the current state: centred coefficients (FP32).  The previous group's MMAs of this
        // block have all completed (this block waited on their mbarrier), so the operand may be rewritten.
#pragma unroll
        for (int k = 0; k < 8; ++k)
            put_column(k, k < K ? (float)__dsub_rn(a.theta[((size_t)k * a.G + g) * S + chl], bbar[k]) : 0.0f);
        double aux_m, aux_r;                                           // LinReg::Aux of the current sigma
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
        for (int p = 0; p < P; ++p, at += GS, hy += S, ++bb, ++n) {
            const bool is_sigma = p == K;
            // proposal and log-priors (pure functions of state known before the sweep; the reference
            // evaluates them after the likelihood, :335, :331)
            const double prop = __dadd_rn(in.cur, __dmul_rn(in.sc, in.z));     // numpy.random.normal(value, sd), :304-306
            double lp_prop, lp_cur;
            if (partial) {
                lp_prop = norm_logpdf_inv(prop, in.h_mu, in.h_isd, in.h_lsd);
                lp_cur = (GENERAL && override_lp) ? in.lp_cur : norm_logpdf_inv(in.cur, in.h_mu, in.h_isd, in.h_lsd);
            } else {
                lp_prop = prior_logpdf(a.prior[p], prop);
                lp_cur = in.lp_cur;
            }
            const double u = in.u;
            float wcur = 0.0f, wprop = 0.0f;
            if (!is_sigma) {                                           // column p of the operand <- the proposal
                wcur = (float)__dsub_rn(in.cur, in.bbar);
                wprop = (float)__dsub_rn(prop, in.bbar);
                put_column(p, wprop);
            }
            publish();
            // Take the accumulator: the other block has read its last chunk.  The barrier also gathers this
            // block's 128 operand writes (block A's very first sweep has nobody to wait for).
            if (block_b || n > 0u) named_sync(bar_take, MCMCN_TC2_THREADS);
            else named_sync(bar_own, 128);
            if ((warp & 3) == 0 && elect_one()) {
                tc_fence_after();
                tc2_issue_chunk(block_b, tbase, bop, stage, ones, np, 0, mb_mma);
            }

            double acc = 0.0;
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(mb_mma, mma_phase);
                mma_phase ^= 1u;
                tc_fence_after();
                const int nc = min(MCMCN_TC_CH, np - c * MCMCN_TC_CH);
                acc += tc_sum_squares_any(tlane + MCMCN_TC_D, nc >> 4);   // one 16-column load at a time: the other
                tc_fence_before();                                      // block's warps cover its latency, registers are scarce
                if (c + 1 < nchunks) {                                 // the accumulator is free once every lane has read it
                    named_sync(bar_own, 128);
                    if ((warp & 3) == 0 && elect_one()) {
                        tc_fence_after();
                        tc2_issue_chunk(block_b, tbase, bop, stage, ones, np, c + 1, mb_mma);
                    }
                }
            }
            named_arrive(bar_give, MCMCN_TC2_THREADS);                 // hand the accumulator to the other block

            double m_prop = aux_m, r_prop = aux_r;
            if (is_sigma) {                                            // LinReg::aux of the proposed sigma
                const double sg = (double)(float)prop;
                if (!(sg > 0.0)) {                                     // scipy: scale <= 0 -> nan
                    m_prop = r_prop = __longlong_as_double(0x7ff8000000000000LL);
                } else {
                    const double inv = 1.0 / sg;
                    m_prop = -0.5 * inv * inv;
                    r_prop = (double)R * (log(sg) + MCMCN_LOG_SQRT_2PI);
                }
            }
            // Parameter.step decision tree, :334-367
            const double llp = acc * m_prop - r_prop;
            const double post_prop = lp_prop + llp;
            const double post_cur = lp_cur + ll_cur;
            const double diff = post_prop - post_cur;
            const bool b1 = !finite64(post_cur) && finite64(post_prop);
            const bool test = finite64(llp) && finite64(diff);         // branches 4/5 draw the uniform
            const int fast = log_u_vs_diff_fast(u, diff);
            bool accept = b1 || (test && fast > 0);
            if (!b1 && test && fast == 0) accept = log(u) < diff;      // rare: within 1e-6 of the threshold
            if (GENERAL) {
                if (trace && on) {
                    a.tr_ll[at] = llp;
                    a.tr_lp[at] = lp_prop;
                    a.tr_diff[at] = diff;
                    a.tr_acc[at] = accept ? 1 : 0;
                }
                if (forced) accept = a.tape_acc[at] != 0;
            }
            if (accept) {                                              // :369-378, :608-610
                if (on) {
                    a.theta[at] = prop;
                    if (!partial) a.lprior[at] = lp_prop;
                }
                ll_cur = llp;
                aux_m = m_prop;
                aux_r = r_prop;
            }
            if (!is_sigma) put_column(p, accept ? wprop : wcur);       // column p <- the value the chain keeps
            if (count && on) {
                unsigned cnt = a.counts[at];
                cnt += accept ? 1u : 0x10000u;
                if (a.tune) {                                          // Parameter.tune, :385-437
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
            // State and random numbers of the next sweep.  Fetched here, not behind the MMAs: the other
            // block covers the tensor core's latency, and nothing fetched has to live through the read-back.
            if (p + 1 < P) {
                tc_fetch_state<GENERAL>(in, a, p + 1, at + GS, hy + S, bb + 1, partial, override_lp);
                tc_fetch_random<GENERAL>(in, a, p + 1, g, chl, at + GS, replay, stash);
            }
        }
        if (on) a.ll[(size_t)g * S + chl] = ll_cur;
        // Block B trails block A; when it is through with a group every MMA that read this stage has
        // completed (each block waited on its own), so B's first thread refills it.
        if (tid == 128 && g + 2 < g1) stage_group(s, g + 2);
    }
    if (!block_b) tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tbase, MCMCN_TC_COLS);
}

}  // namespace mcmcn
