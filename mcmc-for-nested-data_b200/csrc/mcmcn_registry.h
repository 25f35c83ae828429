// mcmcn_registry.h -- table of ahead-of-time compiled kernel sets (one per objective shape and
// precision).  The instantiations are spread over several translation units so that nvcc
// compiles them in parallel; mcmcn_kernels.cu looks them up.
#pragma once

#include "mcmcn_device.cuh"

namespace mcmcn {

typedef void (*sweep_fn)(const SweepArgs);
typedef void (*pointwise_fn)(const SweepArgs, const long long*, double*);

struct KernelSet {
    int objective, P, K, precision;
    int c_wide;                 // chains per lane of the wide variant
    sweep_fn sweep_fast[4];     // wide, production (128-thread CTAs, MB per SM): index = MCMCN_F_PARTIAL | MCMCN_F_COUNT
    sweep_fn sweep_wide;        // wide, general (replay tapes, traces, streamed groups)
    sweep_fn sweep_one;         // one chain per lane, general
    sweep_fn eval_wide, eval_one;
    pointwise_fn pointwise;
    int elem_bytes;
    int park_doubles;           // shared-memory parking slots per chain (7 + Aux doubles)
};

// MB = resident CTAs per SM the production variants are compiled for (__launch_bounds__(128, MB): 3 -> 168
// registers, 8 -> 64).  The regression loop wants the registers (3); the Bernoulli-logit loop is short and its
// decision phases are latency-bound, so it runs best with many resident warps (mcmcn_sets_logit.cu).
#define MCMCN_SET_MB(OBJ_ID, OBJ, KK, PREC, T, CW, MB)                                                   \
    {OBJ_ID, OBJ::P, KK, PREC, CW,                                                                       \
     {sweep_kernel<OBJ, CW, T, MB, 0, 128>, sweep_kernel<OBJ, CW, T, MB, 1, 128>,                        \
      sweep_kernel<OBJ, CW, T, MB, 2, 128>, sweep_kernel<OBJ, CW, T, MB, 3, 128>},                       \
     sweep_kernel<OBJ, CW, T, 2, -1>, sweep_kernel<OBJ, 1, T, 1, -1>,                                    \
     eval_kernel<OBJ, CW, T>, eval_kernel<OBJ, 1, T>, pointwise_kernel<OBJ, T>, (int)sizeof(T), 7 + OBJ::AUX_DOUBLES}
#define MCMCN_SET(OBJ_ID, OBJ, KK, PREC, T, CW) MCMCN_SET_MB(OBJ_ID, OBJ, KK, PREC, T, CW, 3)

const KernelSet* sets_linreg_a(int* n);
const KernelSet* sets_linreg_b(int* n);
const KernelSet* sets_linreg_c(int* n);
const KernelSet* sets_linreg_d(int* n);
const KernelSet* sets_linreg_e(int* n);
const KernelSet* sets_linreg_f(int* n);
const KernelSet* sets_linreg_g(int* n);
const KernelSet* sets_linreg_h(int* n);
const KernelSet* sets_linreg_i(int* n);
const KernelSet* sets_logit(int* n);
const KernelSet* sets_gauss(int* n);
const KernelSet* sets_gauss_b(int* n);
sweep_fn tc_sweep_kernel(int f, bool uniform208, int k_blocks);    // mcmcn_sets_tc.cu (one K block), mcmcn_sets_tc2.cu (two)
sweep_fn tc_sweep_kernel_two_blocks(int f);
typedef void (*eval_tc_fn)(const EvalTcArgs);
eval_tc_fn tc_eval_kernel(int k_blocks);             // eval_tc_kernel<1> (mcmcn_sets_tc.cu), <2> (mcmcn_sets_tc2.cu)
eval_tc_fn tc_eval_kernel_two_blocks();

}  // namespace mcmcn
