// model_plan.h — parameter-independent plan of the model build (host, plain C++).
//
// The reference's trans_emiss_calc (get_trans_emiss.py:8-170) interleaves combinatorics
// (state spaces, omega classes, path keys) with numerics (matrix exponentials, vector
// propagation).  Everything combinatorial depends only on (n_int_AB, n_int_ABC); it is
// computed once here and flattened into index arrays, so that the device only runs
// dense FP64 numerics (model_build.cu).
//
// Reference material restated (paths relative to /root/reference/src/itrails):
//   state spaces / transitions / omega   trans_mat.py:26-194, 269-286, 487-526
//   combine_states                       combine_states.py:5-80
//   path keys of the two-sequence chain  run_markov_chain_AB.py:105-271
//   path keys of the three-seq. chain    run_markov_chain_ABC.py:312-796
//   omega class of a key                 helper_omegas.py:25-87
//   hidden-state order                   get_trans_emiss.py:148-153
// Algebraic restatement used (verified there against the reference):
// the Van Loan sums over omega sub-paths equal sub-blocks of expm(dt * Q[S,S]) on the
// 83-state sets S_xy = {0,x,7} x {0,y,7}, and the last-interval sums are absorption
// probabilities on the transient part of S_xy.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace itr {

using Labels = std::array<uint8_t, 6>;

struct StateSpace {
    int n = 0;        // species in the chain (1, 2 or 3)
    int size = 0;     // 2, 15, 203
    std::vector<Labels> states;                  // restricted-growth labels, length 2n
    std::vector<std::array<int32_t, 3>> trans;   // from, to, kind (1 = coalescence, 2 = recombination)
    std::vector<int> omega_l, omega_r;           // per-locus coalescence class (bitmask over species)
    std::map<Labels, int> index;
    void build(int n_species);
    static Labels canon(const int *labels, int m);
};

// Sparse generator restricted to a subset of states (local indices), with the diagonal
// of the FULL chain (total outgoing coalescence / recombination counts).
struct GenCSR {
    int n = 0;
    std::vector<int32_t> row_ptr, col;
    std::vector<uint8_t> kind;
    std::vector<int32_t> ncoal, nrec;
    std::vector<uint8_t> transient;   // last-interval solve: 1 if some locus is still in class 0
};

enum PlanOpKind : int32_t {
    OP_INIT = 0,     // next[scatter[c + 2 i1 + i2]] = vA[i1] * vB[i2]
    OP_MATVEC = 1,   // next[dst + j] = sum_i cur[src + i] * M[mat][ridx[i] * ld + cidx[j]]
    OP_OUTER = 2,    // next[dst + scatter[c + 2 i + k]] = cur[src + i] * vC[k]
    OP_DOT = 3,      // J[dst] = sum_i cur[src + i] * absorb[mat][ridx[i]]
    OP_SUM = 4       // J[dst] = sum_i cur[src + i]
};

struct PlanOp {
    int32_t kind, src, dst, mat, ridx, cidx, nr, nc;
};

struct PlanStage {
    int32_t op_begin, op_end;
    int32_t next_size;   // doubles in the destination vector buffer (0 for the final stage)
    int32_t zero_next;   // destination must be cleared first (scatter stages)
};

struct EmissionRecipe {      // one hidden state (get_emission_prob_mat.py:803-1033)
    int32_t topo, i, j;
};

struct ModelPlan {
    int n_int_AB = 0, n_int_ABC = 0, K = 0;
    StateSpace ss1, ss2, ss3;
    std::vector<GenCSR> gens;             // 0: one-sequence chain, 1: two-sequence chain, 2..10: S_xy
    std::vector<PlanOp> ops;
    std::vector<PlanStage> stages;
    std::vector<int32_t> idx_pool;        // row / column / scatter index lists
    std::vector<EmissionRecipe> hidden;   // sorted (topology, i, j)
    int n_mats = 0;                       // exponentials per parameter set
    std::vector<int32_t> mat_gen;         // generator id per matrix
    std::vector<int32_t> mat_off;         // offset (doubles) into the per-set matrix pool
    std::vector<int32_t> mat_ld;          // leading dimension
    int64_t mat_pool = 0;                 // doubles per set
    int32_t max_vec = 0;                  // largest vector buffer (doubles)
    int64_t n_keys_max = 0;

    // Throws std::runtime_error on internal inconsistency.
    void build(int n_ab, int n_abc);
};

constexpr int NP1 = 8, NP2 = 16, NP3 = 88;   // padded matrix sizes of the three chains
inline int gen_np(int gen) { return gen == 0 ? NP1 : gen == 1 ? NP2 : NP3; }

}  // namespace itr
