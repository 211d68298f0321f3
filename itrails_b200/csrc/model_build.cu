// model_build.cu — batched device builder of (a, b, pi): replaces trans_emiss_calc
// (reference get_trans_emiss.py:8-170) for n_sets parameter sets at once.
//
// Host: model_plan.cpp flattens everything parameter-independent (state spaces, omega
// classes, path keys) into index arrays, once per (n_int_AB, n_int_ABC).
// Device, per build (all kernels batched over parameter sets):
//   expm_kernel<NT>     one CTA per matrix exponential: generator fill, 1-norm, Pade-13
//                       + scaling and squaring (expm.py:9-167 family); every product is
//                       an FP64 tensor-core contraction (mma.sync m8n8k4.f64, one warp per
//                       8-row tile strip, operands staged in shared memory); the Pade
//                       solve is an in-shared-memory Gauss-Jordan with partial pivoting.
//                       Sizes: 2 (one-sequence chain), 15 (two-sequence chain), 83
//                       (restricted three-sequence generators S_xy).
//   absorb_kernel       last-interval absorption probabilities (deepest_ti.py:215-256
//                       restated as (-Q_TT)^-1 Q_TR 1), one CTA per (set, S_xy).
//   propagate_kernel    one CTA per set walks the plan's stages; one warp per op
//                       (block mat-vec on class-restricted vectors); writes the joint
//                       matrix J, then pi = J 1 and a = J / pi (get_trans_emiss.py:166-168).
//   emission_kernel     one CTA per (hidden state, set): JC69 coalescent tensors and the
//                       4x4 contractions of get_emission_prob_mat.py:47-117,120-424,585-697.
#include "model_build.h"

#include <cmath>
#include <cstdio>
#include <limits>
#include <stdexcept>
#include <vector>

#include "../../include/itrails_b200.h"
#include "model_plan.h"

using namespace itr;

namespace {

// per-set scalar block (coalescent units, get_trans_emiss.py:62-89)
constexpr int SC_TA = 0, SC_TB = 1, SC_TAB = 2, SC_TC = 3, SC_TUP = 4, SC_TOUT = 5, SC_RHO = 6, SC_CAB = 7,
              SC_CABC = 8, SC_MU = 9, SC_CUT = 10;

struct DevGen {
    const int32_t *row_ptr, *col, *ncoal, *nrec;
    const uint8_t *kind, *transient;
    int32_t n, pad;
};

struct ExpmTask {
    double dt, coal, rho;
    int64_t out;     // offset (doubles) into the matrix pool
    int32_t gen, pad;
};

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Pade-13 coefficients (Higham 2005), as in expm.py:21-40
__constant__ double PADE13[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                                  129060195264000.,   10559470521600.,    670442572800.,     33522128640.,
                                  1323241920.,        40840800.,          960960.,           16380.,
                                  182.,               1.};
constexpr double THETA13 = 5.371920351148152;

// ---------------------------------------------------------------------------------
// expm: one CTA (NT warps) per matrix of padded size NP = 8 NT.
// Shared memory: operand buffers X, Y ([NP][LD], LD = NP + 4 keeps the 8-byte fragment
// loads of mma.m8n8k4 conflict-free), pivot row / factor column of the solve.
// Global workspace: seven NP x NP slots per CTA (L2-resident).
// ---------------------------------------------------------------------------------
template <int NT>
struct ExpmCfg {
    static constexpr int NP = 8 * NT, LD = NP + 4, NTHR = 32 * NT;
    static constexpr size_t SMEM = (size_t)(2 * NP * LD + 9 * NP) * sizeof(double) + (4 + 2 * NP) * sizeof(int);
};

template <int NT>
__device__ __forceinline__ void gemm_strip(const double *Xs, const double *Ys,
                                           double (&acc)[NT][2], int warp, int lane) {
    constexpr int NP = 8 * NT, LD = NP + 4;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
    const double *xa = Xs + (8 * warp + g) * LD + t;
    const double *yb = Ys + t * LD + g;
#pragma unroll 2
    for (int kk = 0; kk < NP / 4; ++kk) {
        const double a = xa[4 * kk];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma(acc[nt][0], acc[nt][1], a, yb[4 * kk * LD + 8 * nt]);
    }
}

template <int NT>
__global__ void __launch_bounds__(32 * NT)
expm_kernel(const ExpmTask *__restrict__ tasks, int n_tasks, const DevGen *__restrict__ gens,
            double *__restrict__ ws, double *__restrict__ pool) {
    using C = ExpmCfg<NT>;
    constexpr int NP = C::NP, LD = C::LD, NTHR = C::NTHR, NN = NP * NP;
    extern __shared__ __align__(16) double sm[];
    double *X = sm, *Y = sm + NP * LD, *prow = Y + NP * LD, *fcol = prow + 2 * NP;   // prow[2NP]; fcol: 7 NP doubles (norms, then the solve's buffers)
    int *misc = reinterpret_cast<int *>(fcol + 7 * NP);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double *SA = ws + (size_t)blockIdx.x * 7 * NN, *S2 = SA + NN, *S4 = S2 + NN, *S6 = S4 + NN, *SW = S6 + NN,
           *SU = SW + NN, *SV = SU + NN;
    double acc[NT][2];

    auto load = [&](double *dst, const double *src) {
        for (int e = tid; e < NN; e += NTHR) dst[(e / NP) * LD + (e % NP)] = src[e];
    };
    auto load3 = [&](double *dst, double c6, double c4, double c2) {
        for (int e = tid; e < NN; e += NTHR) dst[(e / NP) * LD + (e % NP)] = c6 * S6[e] + c4 * S4[e] + c2 * S2[e];
    };
    // dst (global, ld NP) = acc [+ c6 S6 + c4 S4 + c2 S2 + cI I]
    auto store = [&](double *dst, bool poly, double c6, double c4, double c2, double cI) {
        const int r = 8 * warp + g;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = 8 * nt + 2 * t + h;
                double v = acc[nt][h];
                if (poly) v += c6 * S6[r * NP + c] + c4 * S4[r * NP + c] + c2 * S2[r * NP + c] + (r == c ? cI : 0.0);
                dst[r * NP + c] = v;
            }
    };

    for (int task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        const ExpmTask tk = tasks[task];
        const DevGen gn = gens[tk.gen];
        // ---- A = dt * Q (trans_mat.py:487-508 on the restricted set, full-chain diagonal)
        for (int e = tid; e < NP * LD; e += NTHR) X[e] = 0.0;
        __syncthreads();
        for (int r = tid; r < gn.n; r += NTHR) {
            X[r * LD + r] = -((double)gn.ncoal[r] * tk.coal + (double)gn.nrec[r] * tk.rho) * tk.dt;
            for (int p = gn.row_ptr[r]; p < gn.row_ptr[r + 1]; ++p)
                X[r * LD + gn.col[p]] = (gn.kind[p] == 2 ? tk.rho : tk.coal) * tk.dt;
        }
        __syncthreads();
        for (int c = tid; c < NP; c += NTHR) {
            double s = 0.0;
            for (int r = 0; r < gn.n; ++r) s += fabs(X[r * LD + c]);
            fcol[c] = s;
        }
        __syncthreads();
        double norm = 0.0;
        for (int c = 0; c < NP; ++c) norm = fmax(norm, fcol[c]);
        int sq = 0;
        if (norm > THETA13) sq = max(0, (int)ceil(log2(norm / THETA13)));
        const double scale = scalbn(1.0, -sq);
        __syncthreads();
        for (int e = tid; e < NP * LD; e += NTHR) X[e] *= scale;
        __syncthreads();
        for (int e = tid; e < NN; e += NTHR) SA[e] = X[(e / NP) * LD + (e % NP)];
        // ---- A2, A4, A6
        gemm_strip<NT>(X, X, acc, warp, lane);
        store(S2, false, 0, 0, 0, 0);
        __syncthreads();
        load(X, S2);
        __syncthreads();
        gemm_strip<NT>(X, X, acc, warp, lane);
        store(S4, false, 0, 0, 0, 0);
        __syncthreads();
        load(Y, S4);
        __syncthreads();
        gemm_strip<NT>(X, Y, acc, warp, lane);
        store(S6, false, 0, 0, 0, 0);
        __syncthreads();
        // ---- U = A (A6 (c13 A6 + c11 A4 + c9 A2) + c7 A6 + c5 A4 + c3 A2 + c1 I)
        load(X, S6);
        load3(Y, PADE13[13], PADE13[11], PADE13[9]);
        __syncthreads();
        gemm_strip<NT>(X, Y, acc, warp, lane);
        store(SW, true, PADE13[7], PADE13[5], PADE13[3], PADE13[1]);
        __syncthreads();
        load(Y, SW);
        load(X, SA);
        __syncthreads();
        gemm_strip<NT>(X, Y, acc, warp, lane);
        store(SU, false, 0, 0, 0, 0);
        __syncthreads();
        load(X, S6);
        load3(Y, PADE13[12], PADE13[10], PADE13[8]);
        __syncthreads();
        gemm_strip<NT>(X, Y, acc, warp, lane);
        store(SV, true, PADE13[6], PADE13[4], PADE13[2], PADE13[0]);
        __syncthreads();
        // ---- solve (V - U) R = (V + U): Gauss-Jordan with partial pivoting on the augmented
        // NP x 2NP matrix [V-U | V+U], held in REGISTERS: thread (rg, cg) = (tid / 16, tid % 16)
        // owns rows 4rg..4rg+3 and columns cg + 16m (m < NT).  Per pivot: the owners of column k
        // publish it; EVERY warp finds the pivot among the rows not used yet (same answer in
        // every warp, so no broadcast and no serial section; implicit pivoting: rows never
        // move, the permutation is applied when the result is written back); the owners of
        // the pivot row publish it as it is; everybody subtracts (column entry / pivot) times
        // that row from its 4 x NT block with register FMAs — the pivot row itself with
        // multiplier 0, so it is never rescaled in place: row where[k] is divided by its
        // pivot once, at the end.  Two barriers per pivot (publications are double buffered).
        {
            double v[4][NT];
            const int rg = tid >> 4, cg = tid & 15;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int m = 0; m < NT; ++m) {
                    const int i = 4 * rg + a, j = cg + 16 * m, e = i * NP + (j < NP ? j : j - NP);
                    const double u = SU[e], w = SV[e];
                    v[a][m] = j < NP ? w - u : w + u;
                }
            double *fc2 = fcol, *pr2 = fcol + 2 * NP, *pv = pr2 + 4 * NP;   // fcol[2][NP], prow[2][2NP], pivot values [NP]
            int *where = misc + 4, *rowof = where + NP;        // where[k] = pivot row of column k
            unsigned usedbits = 0;                             // bit b: row lane + 32 b has been a pivot (same in every warp)
#pragma unroll
            for (int mk = 0; mk < (NP + 15) / 16; ++mk) {
#pragma unroll 1
                for (int kk = 0; kk < 16 && 16 * mk + kk < NP; ++kk) {
                    const int k = 16 * mk + kk;
                    double *fc = fc2 + (k & 1) * NP, *pw = pr2 + (k & 1) * 2 * NP;
                    if (cg == kk) {                        // column k: cg == k % 16, register slot mk
#pragma unroll
                        for (int a = 0; a < 4; ++a) fc[4 * rg + a] = v[a][mk];
                    }
                    __syncthreads();
                    double best = -1.0;
                    int p = 0x7fffffff;
#pragma unroll
                    for (int b = 0; b < (NP + 31) / 32; ++b) {
                        const int i = lane + 32 * b;
                        const double cnd = (i < NP && !((usedbits >> b) & 1u)) ? fabs(fc[i]) : -1.0;
                        if (cnd > best) { best = cnd; p = i; }
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, p, o);
                        if (ob > best || (ob == best && oi < p)) { best = ob; p = oi; }
                    }
                    if ((p & 31) == lane) usedbits |= 1u << (p >> 5);
                    const double pivot = fc[p], inv = 1.0 / pivot;
                    if (tid == 0) { where[k] = p; pv[k] = pivot; }
                    if (rg == (p >> 2)) {                  // the pivot row as it is
                        const int a = p & 3;
#pragma unroll
                        for (int m = 0; m < NT; ++m)
                            pw[cg + 16 * m] = a == 0 ? v[0][m] : a == 1 ? v[1][m] : a == 2 ? v[2][m] : v[3][m];
                    }
                    __syncthreads();
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int i = 4 * rg + a;
                        const double mult = i == p ? 0.0 : fc[i] * inv;
#pragma unroll
                        for (int m = 0; m < NT; ++m) v[a][m] = fma(-mult, pw[cg + 16 * m], v[a][m]);
                    }
                }
            }
            __syncthreads();
            // row where[k], divided by its pivot, is row k of the solution in its right half: Y <- R
            for (int k = tid; k < NP; k += NTHR) rowof[where[k]] = k;
            __syncthreads();
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int kr = rowof[4 * rg + a];
                const double inv = 1.0 / pv[kr];
#pragma unroll
                for (int m = 0; m < NT; ++m) {
                    const int j = cg + 16 * m;
                    if (j >= NP) Y[kr * LD + j - NP] = v[a][m] * inv;
                }
            }
            __syncthreads();
        }
        // ---- squarings: Y <- Y Y
        for (int q = 0; q < sq; ++q) {
            gemm_strip<NT>(Y, Y, acc, warp, lane);
            __syncthreads();
            const int r = 8 * warp + g;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                Y[r * LD + 8 * nt + 2 * t] = acc[nt][0];
                Y[r * LD + 8 * nt + 2 * t + 1] = acc[nt][1];
            }
            __syncthreads();
        }
        double *out = pool + tk.out;
        for (int e = tid; e < NN; e += NTHR) out[e] = Y[(e / NP) * LD + (e % NP)];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// Last interval: w = (-Q_TT)^-1 (Q_TR 1) on the transient states of S_xy
// (restates deepest_ti.py:215-256 + run_markov_chain_ABC.py:519-795).
// grid (9, n_sets), 128 threads.
// ---------------------------------------------------------------------------------
constexpr int AB_MAXT = 72;
__global__ void __launch_bounds__(128)
absorb_kernel(const double *__restrict__ scal, int scal_stride, const DevGen *__restrict__ gens,
              double *__restrict__ absorb) {
    __shared__ double M[AB_MAXT * (AB_MAXT + 1)];
    __shared__ double prow[AB_MAXT + 1], fcol[AB_MAXT];
    __shared__ int rank[NP3], pivot;
    __shared__ int nT_s;
    const int xy = blockIdx.x, set = blockIdx.y, tid = threadIdx.x;
    const DevGen gn = gens[2 + xy];
    const double *sc = scal + (size_t)set * scal_stride;
    const double coal = sc[SC_CABC], rho = sc[SC_RHO];
    if (tid == 0) {
        int n = 0;
        for (int r = 0; r < gn.n; ++r) rank[r] = gn.transient[r] ? n++ : -1;
        nT_s = n;
    }
    __syncthreads();
    const int nT = nT_s, W = nT + 1;
    for (int e = tid; e < nT * W; e += blockDim.x) M[e] = 0.0;
    __syncthreads();
    for (int r = tid; r < gn.n; r += blockDim.x) {
        const int tr = rank[r];
        if (tr < 0) continue;
        M[tr * W + tr] = (double)gn.ncoal[r] * coal + (double)gn.nrec[r] * rho;
        double rhs = 0.0;
        for (int p = gn.row_ptr[r]; p < gn.row_ptr[r + 1]; ++p) {
            const double rate = gn.kind[p] == 2 ? rho : coal;
            const int tc = rank[gn.col[p]];
            if (tc >= 0) M[tr * W + tc] = -rate;
            else rhs += rate;
        }
        M[tr * W + nT] = rhs;
    }
    __syncthreads();
    // Gauss-Jordan on [M | rhs] (nT x (nT + 1)) in registers, as in expm_kernel: thread
    // (rg, cg) = (tid / 16, tid % 16) owns rows 9rg..9rg+8 and columns cg + 16m (m < 5); every
    // warp finds the pivot among the rows not used yet (implicit pivoting), the pivot row is
    // published as it is, two barriers per pivot.  The solution entry of the pivot row of
    // column k is its last column divided by its pivot.
    constexpr int RPT = AB_MAXT / 8, CPT = (AB_MAXT + 1 + 15) / 16;
    double v[RPT][CPT];
    const int rg = tid >> 4, cg = tid & 15, lane = tid & 31;
#pragma unroll
    for (int a = 0; a < RPT; ++a)
#pragma unroll
        for (int m = 0; m < CPT; ++m) {
            const int i = RPT * rg + a, j = cg + 16 * m;
            v[a][m] = (i < nT && j < W) ? M[i * W + j] : 0.0;
        }
    __syncthreads();                                   // M is reused below as publication space
    double *fc2 = M, *pr2 = M + 2 * AB_MAXT, *sol = pr2 + 2 * 16 * CPT, *pvs = sol + AB_MAXT;   // fcol[2][72], prow[2][80], solution [72], pivots [72]
    int *where = reinterpret_cast<int *>(pvs + AB_MAXT);
    unsigned usedbits = 0;
#pragma unroll
    for (int mk = 0; mk < (AB_MAXT + 15) / 16; ++mk) {
#pragma unroll 1
        for (int kk = 0; kk < 16 && 16 * mk + kk < nT; ++kk) {
            const int k = 16 * mk + kk;
            double *fc = fc2 + (k & 1) * AB_MAXT, *pw = pr2 + (k & 1) * 16 * CPT;
            if (cg == kk) {
#pragma unroll
                for (int a = 0; a < RPT; ++a) fc[RPT * rg + a] = v[a][mk];
            }
            __syncthreads();
            double best = -1.0;
            int p = 0x7fffffff;
#pragma unroll
            for (int b = 0; b < (AB_MAXT + 31) / 32; ++b) {
                const int i = lane + 32 * b;
                const double cnd = (i < nT && !((usedbits >> b) & 1u)) ? fabs(fc[i]) : -1.0;
                if (cnd > best) { best = cnd; p = i; }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, p, o);
                if (ob > best || (ob == best && oi < p)) { best = ob; p = oi; }
            }
            if ((p & 31) == lane) usedbits |= 1u << (p >> 5);
            const double inv = 1.0 / fc[p];
            if (tid == 0) { where[k] = p; pvs[k] = fc[p]; }
            if (rg == p / RPT) {
                const int a = p - RPT * rg;
#pragma unroll
                for (int m = 0; m < CPT; ++m) {
                    double x = v[0][m];
#pragma unroll
                    for (int q = 1; q < RPT; ++q) x = a == q ? v[q][m] : x;
                    pw[cg + 16 * m] = x;
                }
            }
            __syncthreads();
#pragma unroll
            for (int a = 0; a < RPT; ++a) {
                const int i = RPT * rg + a;
                const double mult = i == p ? 0.0 : fc[i] * inv;
#pragma unroll
                for (int m = 0; m < CPT; ++m) v[a][m] = fma(-mult, pw[cg + 16 * m], v[a][m]);
            }
        }
    }
    // The pivot row of column k, divided by the pivot it had when it was chosen, is row k of
    // the reduced system: w_k = (its last column) / pivot_k.
    __syncthreads();
    {
        const int mlast = nT / 16, clast = nT % 16;            // the right-hand side: column nT = slot mlast of cg == clast
        double *rhsv = fc2;                                    // [72], indexed by row
#pragma unroll
        for (int m = 0; m < CPT; ++m)
            if (m == mlast && cg == clast) {
#pragma unroll
                for (int a = 0; a < RPT; ++a) rhsv[RPT * rg + a] = v[a][m];
            }
        __syncthreads();
        for (int k = tid; k < nT; k += blockDim.x) sol[k] = rhsv[where[k]] / pvs[k];
        __syncthreads();
    }
    double *out = absorb + ((size_t)set * 9 + xy) * NP3;
    for (int r = tid; r < NP3; r += blockDim.x) out[r] = (r < gn.n && rank[r] >= 0) ? M[2 * AB_MAXT + 2 * 16 * ((AB_MAXT + 1 + 15) / 16) + rank[r]] : 0.0;
}

// ---------------------------------------------------------------------------------
// Propagation of the path-key vectors through the plan's stages; one CTA per set.
// ---------------------------------------------------------------------------------
struct DevPlan {
    const PlanOp *ops;
    const PlanStage *stages;
    const int32_t *idx;
    const int32_t *mat_off, *mat_ld;
    int32_t n_stages, K, max_vec, pad;
    int64_t mat_pool;
};

__global__ void __launch_bounds__(512)
propagate_kernel(DevPlan pl, const double *__restrict__ pool, const double *__restrict__ absorb,
                 double *__restrict__ vec, double *__restrict__ J, double *__restrict__ a_out,
                 double *__restrict__ pi_out) {
    const int set = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const double *mats = pool + (size_t)set * pl.mat_pool;
    const double *ab = absorb + (size_t)set * 9 * NP3;
    double *cur = vec + (size_t)set * 2 * pl.max_vec, *nxt = cur + pl.max_vec;
    double *Js = J + (size_t)set * pl.K * pl.K;
    for (int e = tid; e < pl.K * pl.K; e += blockDim.x) Js[e] = 0.0;
    for (int s = 0; s < pl.n_stages; ++s) {
        const PlanStage st = pl.stages[s];
        if (st.zero_next) {
            for (int e = tid; e < st.next_size; e += blockDim.x) nxt[e] = 0.0;
            __syncthreads();
        }
        for (int o = st.op_begin + warp; o < st.op_end; o += nwarps) {
            const PlanOp op = pl.ops[o];
            switch (op.kind) {
                case OP_INIT: {
                    const double *vA = mats + pl.mat_off[0], *vB = mats + pl.mat_off[1];
                    if (lane < 4) nxt[op.dst + pl.idx[op.cidx + lane]] = vA[lane >> 1] * vB[lane & 1];
                    break;
                }
                case OP_MATVEC: {
                    const double *M = mats + pl.mat_off[op.mat];
                    const int ld = pl.mat_ld[op.mat];
                    const int32_t *ri = pl.idx + op.ridx, *ci = pl.idx + op.cidx;
                    for (int j = lane; j < op.nc; j += 32) {
                        const double *Mc = M + ci[j];
                        double acc = 0.0;
                        for (int i = 0; i < op.nr; ++i) acc = fma(cur[op.src + i], Mc[ri[i] * ld], acc);
                        nxt[op.dst + j] = acc;
                    }
                    break;
                }
                case OP_OUTER: {
                    const double *vC = mats + pl.mat_off[2];
                    for (int e = lane; e < 2 * op.nr; e += 32)
                        nxt[op.dst + pl.idx[op.cidx + e]] = cur[op.src + (e >> 1)] * vC[e & 1];
                    break;
                }
                case OP_DOT: {
                    const double *w = ab + op.mat * NP3;
                    const int32_t *ri = pl.idx + op.ridx;
                    double acc = 0.0;
                    for (int i = lane; i < op.nr; i += 32) acc = fma(cur[op.src + i], w[ri[i]], acc);
#pragma unroll
                    for (int q = 16; q; q >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, q);
                    if (lane == 0) Js[op.dst] = acc;
                    break;
                }
                default: {   // OP_SUM
                    double acc = 0.0;
                    for (int i = lane; i < op.nr; i += 32) acc += cur[op.src + i];
#pragma unroll
                    for (int q = 16; q; q >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, q);
                    if (lane == 0) Js[op.dst] = acc;
                }
            }
        }
        __syncthreads();
        double *tmp = cur;
        cur = nxt;
        nxt = tmp;
    }
    // pi = J 1, a = J / pi                                   get_trans_emiss.py:166-168
    const int K = pl.K;
    for (int r = warp; r < K; r += nwarps) {
        double acc = 0.0;
        for (int c = lane; c < K; c += 32) acc += Js[r * K + c];
#pragma unroll
        for (int q = 16; q; q >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, q);
        if (lane == 0) pi_out[(size_t)set * K + r] = acc;
        for (int c = lane; c < K; c += 32) a_out[((size_t)set * K + r) * K + c] = Js[r * K + c] / acc;
    }
}

// ---------------------------------------------------------------------------------
// Emissions (JC69).  Q4 = 1/4 everywhere, D4 = delta - 1/4; every branch matrix is
// Q4 + e^{-m} D4 (get_emission_prob_mat.py:9-44 with equal rates).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double q4(int) { return 0.25; }
__device__ __forceinline__ double d4(int a, int b) { return (a == b ? 1.0 : 0.0) - 0.25; }
__device__ __forceinline__ double m4(int which, int a, int b) { return which ? d4(a, b) : 0.25; }
__device__ __forceinline__ double int_exp(double lam, double t) {   // int_0^t e^{-lam u} du
    return lam == 0.0 ? t : -expm1(-lam * t) / lam;
}

// F[x1][x2][y] of get_emission_prob_mat.py:47-117 (first coalescence inside an interval
// of length t at rate k, conditioned on happening), by term-wise integration.
__device__ void single_coal_tensor(double *F, double t, double mu, double k, int tid) {
    if (tid < 64) {
        const double norm = -expm1(-k * t);
        const double i1 = k * int_exp(k + mu, t) / norm;
        const double i2 = k * int_exp(k + 2 * mu, t) / norm;
        const double j1 = exp(-mu * t);
        const double j0 = j1 * k * int_exp(k - mu, t) / norm;
        const double j2 = j1 * i1;
        const int a = tid >> 4, b = (tid >> 2) & 3, y = tid & 3;
        double acc = 0.0;
        for (int d = 0; d < 4; ++d) {
            const double Da = d4(a, d), Db = d4(b, d), Dy = d4(d, y);
            acc += 0.25 * 0.25 * 0.25 + i1 * (Da * 0.0625 + 0.0625 * Db) + i2 * Da * Db * 0.25 + j0 * 0.0625 * Dy +
                   j1 * (Da * 0.25 * Dy + 0.25 * Db * Dy) + j2 * Da * Db * Dy;
        }
        F[tid] = acc;
    }
}

// G[x1][x2][x3][y] of get_emission_prob_mat.py:120-424 (both coalescences inside one
// interval of length t, rates 3 then 1, conditioned on both and on the topology).
__device__ void double_coal_tensor(double *G, double t, double mu, int tid) {
    __shared__ double coef[32];
    if (tid < 32) {
        const int na = tid >> 4, nb = (tid >> 3) & 1, ng = (tid >> 2) & 1, nd = (tid >> 1) & 1, ne = tid & 1;
        const double pboth = 1.0 + 0.5 * exp(-3.0 * t) - 1.5 * exp(-t);
        const double p = 2.0 + mu * (na + nb - ng), q = 1.0 + mu * (ng + nd - ne);
        const double integ = (int_exp(p + q, t) - exp(-q * t) * int_exp(p, t)) / q;
        coef[tid] = 3.0 / pboth * exp(-mu * ne * t) * integ;
    }
    __syncthreads();
    if (tid < 256) {
        const int a = tid >> 6, b = (tid >> 4) & 3, c = (tid >> 2) & 3, d = tid & 3;
        double acc = 0.0;
        for (int m = 0; m < 32; ++m) {
            const int na = m >> 4, nb = (m >> 3) & 1, ng = (m >> 2) & 1, nd = (m >> 1) & 1, ne = m & 1;
            double s = 0.0;
            for (int e = 0; e < 4; ++e) {
                const double ab = m4(na, a, e) * m4(nb, b, e);
                for (int f = 0; f < 4; ++f) s += ab * m4(ng, e, f) * m4(nd, c, f) * m4(ne, f, d);
            }
            acc += coef[m] * s;
        }
        G[tid] = acc;
    }
    __syncthreads();
}

// grid (K, n_sets), 256 threads
__global__ void __launch_bounds__(256)
emission_kernel(const EmissionRecipe *__restrict__ hidden, const double *__restrict__ scal, int scal_stride,
                int n_ab, int n_abc, int K, double *__restrict__ b_out) {
    __shared__ double P[5][16];     // Pa, Pb, Pc (third lineage), Pab, Pd
    __shared__ double F1[64], F2[64], G[256], T1[64], T2[64], R3[64], R[64], H[256];
    const int tid = threadIdx.x, set = blockIdx.y;
    const EmissionRecipe h = hidden[blockIdx.x];
    const double *sc = scal + (size_t)set * scal_stride;
    const double *cAB = sc + SC_CUT, *cABC = cAB + n_ab + 1;
    const double mu = sc[SC_MU], t_A = sc[SC_TA], t_B = sc[SC_TB], t_C = sc[SC_TC], t_AB = sc[SC_TAB],
                 t_up = sc[SC_TUP], t_out = sc[SC_TOUT];
    const int n = n_abc, i = h.i, j = h.j;
    auto width = [&](int q) { return q != n - 1 ? cABC[q + 1] - cABC[q] : t_up; };            // :818-820
    auto above = [&](int q) { return q != n - 1 ? t_up + cABC[n - 1] - cABC[q + 1] : 0.0; };  // :822-826
    // branch lengths (rate * time) of the five branch matrices
    double ma, mb, mc, mab = 0.0, md;
    if (h.topo == 0) {
        ma = mu * (t_A + cAB[i]);
        mb = mu * (t_B + cAB[i]);
        mc = mu * (t_C + cABC[j]);
        mab = mu * (t_AB - cAB[i + 1] + cABC[j]);
    } else {
        ma = mu * (t_A + t_AB + cABC[i]);
        mb = mu * (t_B + t_AB + cABC[i]);
        mc = mu * (t_C + cABC[i]);
        if (i != j) mab = mu * (cABC[j] - cABC[i + 1]);
    }
    md = mu * (t_out + above(j));
    // roles: (first, second) coalesce first, `third` joins later
    double m1 = ma, m2 = mb, m3 = mc;
    if (h.topo == 2) { m1 = ma; m2 = mc; m3 = mb; }
    if (h.topo == 3) { m1 = mb; m2 = mc; m3 = ma; }
    if (tid < 80) {
        const int which = tid >> 4, e = tid & 15;
        const double m = which == 0 ? m1 : which == 1 ? m2 : which == 2 ? m3 : which == 3 ? mab : md;
        P[which][e] = 0.25 + exp(-m) * d4(e >> 2, e & 3);
    }
    const bool dbl = (h.topo != 0 && i == j);
    double X = 0.0;     // value at (x1, x2, x3, d) for this thread: tid = 64 x1 + 16 x2 + 4 x3 + d
    const int x1 = tid >> 6, x2 = (tid >> 4) & 3, x3 = (tid >> 2) & 3, xd = tid & 3;
    if (!dbl) {
        if (h.topo == 0) single_coal_tensor(F1, cAB[i + 1] - cAB[i], mu, sc[SC_CAB], tid);
        else single_coal_tensor(F1, cABC[i + 1] - cABC[i], mu, sc[SC_CABC], tid);
        single_coal_tensor(F2, width(j), mu, sc[SC_CABC], tid);
        __syncthreads();
        // einsum("ai,jb,ijk,kl,lmn,mc,nd->abcd") / 4          :585-606
        if (tid < 64) {
            const int a = tid >> 4, b = (tid >> 2) & 3, k = tid & 3;
            double s = 0.0;
            for (int p = 0; p < 4; ++p)
                for (int q = 0; q < 4; ++q) s += P[0][a * 4 + p] * P[1][q * 4 + b] * F1[p * 16 + q * 4 + k];
            T1[tid] = s;
        } else if (tid < 128) {
            const int e = tid - 64, l = e >> 4, c = (e >> 2) & 3, nn = e & 3;
            double s = 0.0;
            for (int m = 0; m < 4; ++m) s += F2[l * 16 + m * 4 + nn] * P[2][m * 4 + c];
            R3[e] = s;
        }
        __syncthreads();
        if (tid < 64) {
            const int a = tid >> 4, b = (tid >> 2) & 3, l = tid & 3;
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += T1[a * 16 + b * 4 + k] * P[3][k * 4 + l];
            T2[tid] = s;
        } else if (tid < 128) {
            const int e = tid - 64, l = e >> 4, c = (e >> 2) & 3, d = e & 3;
            double s = 0.0;
            for (int nn = 0; nn < 4; ++nn) s += R3[l * 16 + c * 4 + nn] * P[4][nn * 4 + d];
            R[e] = s;
        }
        __syncthreads();
        for (int l = 0; l < 4; ++l) X += T2[x1 * 16 + x2 * 4 + l] * R[l * 16 + x3 * 4 + xd];
        X *= 0.25;
    } else {
        double_coal_tensor(G, width(i), mu, tid);
        // einsum("ai,jb,kc,ijkn,nd->abcd") / 4                :681-697
        {
            const int p = tid >> 6, q = (tid >> 4) & 3, r = (tid >> 2) & 3, d = tid & 3;
            double s = 0.0;
            for (int nn = 0; nn < 4; ++nn) s += G[p * 64 + q * 16 + r * 4 + nn] * P[4][nn * 4 + d];
            H[tid] = s;
        }
        __syncthreads();
        for (int p = 0; p < 4; ++p)
            for (int q = 0; q < 4; ++q)
                for (int r = 0; r < 4; ++r)
                    X += P[0][x1 * 4 + p] * P[1][q * 4 + x2] * P[2][r * 4 + x3] * H[p * 64 + q * 16 + r * 4 + xd];
        X *= 0.25;
    }
    // X is indexed by (first, second, third, outgroup) nucleotides; the observed index is
    // 64 A + 16 B + 4 C + D (read_data.py:6-24).           :853-902, :944-984
    int A = x1, B = x2, Cn = x3;
    if (h.topo == 2) { A = x1; Cn = x2; B = x3; }
    if (h.topo == 3) { B = x1; Cn = x2; A = x3; }
    b_out[((size_t)set * K + blockIdx.x) * 256 + 64 * A + 16 * B + 4 * Cn + xd] = X;
}

template <typename T>
cudaError_t upload(T *&d, const std::vector<T> &h) {
    cudaError_t e = cudaMalloc((void **)&d, std::max<size_t>(h.size(), 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

}  // namespace

// ---------------------------------------------------------------------------------
struct BuilderState {
    ModelPlan plan;
    // device copies of the plan
    std::vector<void *> owned;
    DevGen *d_gens = nullptr;
    PlanOp *d_ops = nullptr;
    PlanStage *d_stages = nullptr;
    int32_t *d_idx = nullptr, *d_mat_off = nullptr, *d_mat_ld = nullptr;
    EmissionRecipe *d_hidden = nullptr;
    // per-build buffers
    double *d_scal = nullptr, *d_pool = nullptr, *d_absorb = nullptr, *d_vec = nullptr, *d_J = nullptr,
           *d_a = nullptr, *d_b = nullptr, *d_pi = nullptr, *d_ws = nullptr;
    ExpmTask *d_tasks = nullptr;
    int cap_sets = 0;
    int ws_ctas[3] = {0, 0, 0};
    int sm_count = 0;

    ~BuilderState() { release(); }
    void release() {
        for (void *p : owned) cudaFree(p);
        owned.clear();
        free_build();
    }
    void free_build() {
        for (void *p : {(void *)d_scal, (void *)d_pool, (void *)d_absorb, (void *)d_vec, (void *)d_J, (void *)d_a,
                        (void *)d_b, (void *)d_pi, (void *)d_ws, (void *)d_tasks})
            if (p) cudaFree(p);
        d_scal = d_pool = d_absorb = d_vec = d_J = d_a = d_b = d_pi = d_ws = nullptr;
        d_tasks = nullptr;
        cap_sets = 0;
    }
};

ModelBuilder::ModelBuilder() {}
ModelBuilder::~ModelBuilder() { delete state_; }

#define MB_CK(call)                                                             \
    do {                                                                        \
        cudaError_t e_ = (call);                                                \
        if (e_ != cudaSuccess) {                                                \
            msg = std::string(#call) + " failed: " + cudaGetErrorString(e_);    \
            return e_ == cudaErrorMemoryAllocation ? ITR_ERR_NOMEM : ITR_ERR_CUDA; \
        }                                                                       \
    } while (0)

int ModelBuilder::build(cudaStream_t stream, int n_sets, const double *params, int n_ab, int n_abc,
                        const double *cut_AB, const double *cut_ABC, const double **d_a, const double **d_b,
                        const double **d_pi, int32_t *hidden, int64_t *launched, std::string &msg) {
    if (launched) *launched = 0;
    // ---- plan (cached per discretisation) ---------------------------------------------
    if (!state_ || state_->plan.n_int_AB != n_ab || state_->plan.n_int_ABC != n_abc) {
        delete state_;
        state_ = new BuilderState();
        BuilderState &S = *state_;
        try {
            S.plan.build(n_ab, n_abc);
        } catch (const std::exception &ex) {
            msg = std::string("plan: ") + ex.what();
            delete state_;
            state_ = nullptr;
            return ITR_ERR_ARG;
        }
        const ModelPlan &P = S.plan;
        std::vector<DevGen> hg(P.gens.size());
        for (size_t k = 0; k < P.gens.size(); ++k) {
            const GenCSR &g = P.gens[k];
            int32_t *rp = nullptr, *col = nullptr, *nc = nullptr, *nr = nullptr;
            uint8_t *kind = nullptr, *tr = nullptr;
            MB_CK(upload(rp, g.row_ptr)); S.owned.push_back(rp);
            MB_CK(upload(col, g.col)); S.owned.push_back(col);
            MB_CK(upload(nc, g.ncoal)); S.owned.push_back(nc);
            MB_CK(upload(nr, g.nrec)); S.owned.push_back(nr);
            MB_CK(upload(kind, g.kind)); S.owned.push_back(kind);
            MB_CK(upload(tr, g.transient)); S.owned.push_back(tr);
            hg[k] = DevGen{rp, col, nc, nr, kind, tr, g.n, 0};
        }
        MB_CK(upload(S.d_gens, hg)); S.owned.push_back(S.d_gens);
        MB_CK(upload(S.d_ops, P.ops)); S.owned.push_back(S.d_ops);
        MB_CK(upload(S.d_stages, P.stages)); S.owned.push_back(S.d_stages);
        MB_CK(upload(S.d_idx, P.idx_pool)); S.owned.push_back(S.d_idx);
        MB_CK(upload(S.d_mat_off, P.mat_off)); S.owned.push_back(S.d_mat_off);
        MB_CK(upload(S.d_mat_ld, P.mat_ld)); S.owned.push_back(S.d_mat_ld);
        MB_CK(upload(S.d_hidden, P.hidden)); S.owned.push_back(S.d_hidden);
        int dev = 0;
        MB_CK(cudaGetDevice(&dev));
        MB_CK(cudaDeviceGetAttribute(&S.sm_count, cudaDevAttrMultiProcessorCount, dev));
        MB_CK(cudaFuncSetAttribute(expm_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ExpmCfg<11>::SMEM));
    }
    BuilderState &S = *state_;
    const ModelPlan &P = S.plan;
    const int K = P.K;
    if (hidden)
        for (int k = 0; k < K; ++k) {
            hidden[3 * k] = P.hidden[k].topo;
            hidden[3 * k + 1] = P.hidden[k].i;
            hidden[3 * k + 2] = P.hidden[k].j;
        }

    // ---- per-set scalars and exponential tasks (host, get_trans_emiss.py:62-100) ----------
    const int stride = SC_CUT + (n_ab + 1) + (n_abc + 1);
    std::vector<double> scal((size_t)n_sets * stride);
    std::vector<ExpmTask> tasks[3];
    for (int s = 0; s < n_sets; ++s) {
        const double *p = params + (size_t)s * 9;
        const double N_AB = p[6], N_ABC = p[7], r = p[8];
        for (int q = 0; q < 9; ++q)
            if (!(p[q] > 0.0) || !std::isfinite(p[q])) {
                static const char *names[9] = {"t_A", "t_B", "t_C", "t_2", "t_upper", "t_out", "N_AB", "N_ABC", "r"};
                // (t_upper == 0 passes the reference's workflow validation, which only forbids
                // negative values, but its emission build then divides by the length of the last
                // interval: ZeroDivisionError in get_emission_prob_mat.py — an error there too)
                msg = "parameter set " + std::to_string(s) + ": " + names[q] + " = " + std::to_string(p[q]) +
                      " — every parameter must be positive and finite" +
                      (q == 4 ? " (t_upper == 0, i.e. t_3 equal to the last ABC cutpoint, is a division by zero in the reference as well)" : "");
                return ITR_ERR_ARG;
            }
        const double N_ref = N_ABC;
        double *sc = &scal[(size_t)s * stride];
        sc[SC_TA] = p[0] / N_ref;
        sc[SC_TB] = p[1] / N_ref;
        sc[SC_TC] = p[2] / N_ref;
        sc[SC_TAB] = p[3] / N_ref;
        sc[SC_TUP] = p[4] / N_ref;
        sc[SC_TOUT] = p[5] / N_ref;
        sc[SC_RHO] = N_ref * r;
        sc[SC_CAB] = N_ref / N_AB;
        sc[SC_CABC] = N_ref / N_ABC;
        sc[SC_MU] = N_ref * (4.0 / 3.0);
        double *cab = sc + SC_CUT, *cabc = cab + n_ab + 1;
        {   // cutpoints.py:5-26: truncexpon.ppf(q, b = t_AB / scale, scale = 1 / coal_AB) = -log1p(q expm1(-b)) scale,
            // except q == 1, where scipy's ppf returns the upper end of the support, b * scale
            const double scale = 1.0 / sc[SC_CAB], bb = sc[SC_TAB] / scale;
            for (int q = 0; q <= n_ab; ++q)
                cab[q] = cut_AB ? cut_AB[q] : (q == n_ab ? bb * scale : -std::log1p(((double)q / n_ab) * std::expm1(-bb)) * scale);
        }
        for (int q = 0; q <= n_abc; ++q)     // cutpoints.py:29-45 (expon.ppf), last = +inf
            cabc[q] = cut_ABC ? cut_ABC[q]
                              : (q == n_abc ? std::numeric_limits<double>::infinity()
                                            : -std::log1p(-(double)q / n_abc) / sc[SC_CABC]);
        for (int m = 0; m < P.n_mats; ++m) {
            ExpmTask t{};
            t.gen = P.mat_gen[m];
            t.rho = sc[SC_RHO];
            t.out = (int64_t)s * P.mat_pool + P.mat_off[m];
            int cls;
            if (m < 3) {
                t.dt = m == 0 ? sc[SC_TA] : m == 1 ? sc[SC_TB] : sc[SC_TC];
                t.coal = sc[SC_CAB];            // coal_A = coal_B = coal_C = N_ref / N_AB (:76-80)
                cls = 0;
            } else if (m < 3 + n_ab) {
                t.dt = cab[m - 3 + 1] - cab[m - 3];
                t.coal = sc[SC_CAB];
                cls = 1;
            } else {
                const int iv = (m - 3 - n_ab) / 9;
                t.dt = cabc[iv + 1] - cabc[iv];
                t.coal = sc[SC_CABC];
                cls = 2;
            }
            if (!(t.dt >= 0.0) || !std::isfinite(t.dt)) {
                msg = "parameter set " + std::to_string(s) + ": cutpoints must be increasing and finite";
                return ITR_ERR_ARG;
            }
            tasks[cls].push_back(t);
        }
    }

    // ---- device buffers ----------------------------------------------------------------
    if (n_sets > S.cap_sets) {
        S.free_build();
        MB_CK(cudaMalloc((void **)&S.d_scal, (size_t)n_sets * stride * sizeof(double)));
        MB_CK(cudaMalloc((void **)&S.d_pool, (size_t)n_sets * P.mat_pool * sizeof(double)));
        MB_CK(cudaMalloc((void **)&S.d_absorb, (size_t)n_sets * 9 * NP3 * sizeof(double)));
        MB_CK(cudaMalloc((void **)&S.d_vec, (size_t)n_sets * 2 * std::max(P.max_vec, 1) * sizeof(double)));
        MB_CK(cudaMalloc((void **)&S.d_J, (size_t)n_sets * K * K * sizeof(double)));
        MB_CK(cudaMalloc((void **)&S.d_a, (size_t)n_sets * K * K * sizeof(double)));
        MB_CK(cudaMalloc((void **)&S.d_b, (size_t)n_sets * K * 256 * sizeof(double)));
        MB_CK(cudaMalloc((void **)&S.d_pi, (size_t)n_sets * K * sizeof(double)));
        MB_CK(cudaMalloc((void **)&S.d_tasks, (size_t)n_sets * P.n_mats * sizeof(ExpmTask)));
        const int nps[3] = {NP1, NP2, NP3};
        size_t ws = 0;
        for (int c = 0; c < 3; ++c) {
            const int64_t nt = (int64_t)n_sets * (c == 0 ? 3 : c == 1 ? n_ab : 9 * (n_abc - 1));
            S.ws_ctas[c] = (int)std::max<int64_t>(1, std::min<int64_t>(nt, (int64_t)S.sm_count * (c == 2 ? 1 : 8)));
            ws += (size_t)S.ws_ctas[c] * 7 * nps[c] * nps[c];
        }
        MB_CK(cudaMalloc((void **)&S.d_ws, ws * sizeof(double)));
        S.cap_sets = n_sets;
    }
    MB_CK(cudaMemcpyAsync(S.d_scal, scal.data(), scal.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    size_t toff = 0;
    const ExpmTask *d_t[3];
    for (int c = 0; c < 3; ++c) {
        d_t[c] = S.d_tasks + toff;
        if (!tasks[c].empty())
            MB_CK(cudaMemcpyAsync(S.d_tasks + toff, tasks[c].data(), tasks[c].size() * sizeof(ExpmTask),
                                  cudaMemcpyHostToDevice, stream));
        toff += tasks[c].size();
    }
    // the task vectors must outlive the async copies
    MB_CK(cudaStreamSynchronize(stream));

    // ---- launches ---------------------------------------------------------------------------
    int64_t n_launch = 0;
    double *ws0 = S.d_ws, *ws1 = ws0 + (size_t)S.ws_ctas[0] * 7 * NP1 * NP1,
           *ws2 = ws1 + (size_t)S.ws_ctas[1] * 7 * NP2 * NP2;
    {
        const int n0 = (int)tasks[0].size(), g0 = std::min(n0, S.ws_ctas[0]);
        expm_kernel<1><<<g0, 32, ExpmCfg<1>::SMEM, stream>>>(d_t[0], n0, S.d_gens, ws0, S.d_pool);
        const int n1 = (int)tasks[1].size(), g1 = std::min(n1, S.ws_ctas[1]);
        expm_kernel<2><<<g1, 64, ExpmCfg<2>::SMEM, stream>>>(d_t[1], n1, S.d_gens, ws1, S.d_pool);
        n_launch += 2;
        const int n2 = (int)tasks[2].size(), g2 = std::min(n2, S.ws_ctas[2]);
        if (n2 > 0) {
            expm_kernel<11><<<g2, 352, ExpmCfg<11>::SMEM, stream>>>(d_t[2], n2, S.d_gens, ws2, S.d_pool);
            n_launch += 1;
        }
    }
    absorb_kernel<<<dim3(9, n_sets), 128, 0, stream>>>(S.d_scal, stride, S.d_gens, S.d_absorb);
    DevPlan dp{S.d_ops, S.d_stages, S.d_idx, S.d_mat_off, S.d_mat_ld, (int32_t)P.stages.size(), K,
               std::max(P.max_vec, 1), 0, P.mat_pool};
    propagate_kernel<<<n_sets, 512, 0, stream>>>(dp, S.d_pool, S.d_absorb, S.d_vec, S.d_J, S.d_a, S.d_pi);
    emission_kernel<<<dim3(K, n_sets), 256, 0, stream>>>(S.d_hidden, S.d_scal, stride, n_ab, n_abc, K, S.d_b);
    n_launch += 3;
    MB_CK(cudaGetLastError());
    if (launched) *launched = n_launch;
    *d_a = S.d_a;
    *d_b = S.d_b;
    *d_pi = S.d_pi;
    return ITR_OK;
}
