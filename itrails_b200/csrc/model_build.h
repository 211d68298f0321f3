// model_build.h — device-side builder of (a, b, pi): replaces trans_emiss_calc
// (reference get_trans_emiss.py:8-170).  Implemented in model_build.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

struct BuilderState;

class ModelBuilder {
  public:
    ModelBuilder();
    ~ModelBuilder();
    // Builds n_sets models on `stream`.  On success *d_a (n_sets x K x K),
    // *d_b (n_sets x K x 256) and *d_pi (n_sets x K) point to device buffers owned by
    // the builder (valid until the next build); hidden (nullable, host) receives the
    // K x 3 sorted hidden-state tuples.  Returns 0 or a negative itr_status with msg.
    int build(cudaStream_t stream, int n_sets, const double *params, int n_int_AB, int n_int_ABC,
              const double *cut_AB, const double *cut_ABC, const double **d_a, const double **d_b,
              const double **d_pi, int32_t *hidden, int64_t *launched, std::string &msg);

  private:
    BuilderState *state_ = nullptr;
};
