// hmm_common.cuh — types shared by the recursion kernels (hmm_kernels.cuh, lockstep.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace itr {

constexpr int NSYM = 625;
constexpr unsigned FULL = 0xffffffffu;

struct ChainSet {
    const uint16_t *sym;     // all blocks back to back
    const int64_t *off;      // n_blocks + 1
    const int32_t *order;    // block ids, longest first
    int32_t n_blocks;
    int32_t n_sets;
    unsigned int *queue;     // work counter (zeroed before launch)
};

}  // namespace itr
