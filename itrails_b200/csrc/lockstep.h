// lockstep.h — launchers of the lock-step FP64 tensor-core sweeps (lockstep.cu).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "hmm_common.cuh"

namespace itr {

// 32 < K <= 96
bool lockstep_supports(int K);

// scratch words a launch needs (zeroed by the launcher): 512 + one per group of chains
inline size_t lockstep_scratch_words(int64_t n_chains) { return 512 + (size_t)(n_chains + 3) / 4 + 8; }

// per-block log-likelihoods loglik[set * n_blocks + blk]
cudaError_t launch_lockstep_loglik(const ChainSet &cs, const double *A, const double *PI, const double *Et, int K,
                                   int KP, double *loglik, unsigned int *scratch, int sms, cudaStream_t st);

// posterior rows of every block of cs (set 0), written into post[(off[blk] + t) * K + state]
cudaError_t launch_lockstep_posterior(const ChainSet &cs, const double *A, const double *PI, const double *Et, int K,
                                      int KP, double *post, unsigned int *scratch, int sms, cudaStream_t st);

}  // namespace itr
