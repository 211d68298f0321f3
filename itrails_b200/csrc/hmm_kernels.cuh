// hmm_kernels.cuh — forward / Viterbi / posterior recursions for sm_100a.
//
// Replaces (reference paths relative to /root/reference/src/itrails):
//   forward, forward_loglik      optimizer.py:146-188
//   backward, post_prob          optimizer.py:192-238
//   viterbi, backtrack_viterbi   optimizer.py:305-354
//
// Design: an alignment block is a dependent chain, so one warp
// walks one chain.  Lane j owns hidden state j (and j+32, j+64 ... when K > 32): it
// keeps column j of the transition matrix in registers, the state vector is exchanged
// through a double-buffered 256-byte shared-memory line read back with broadcast
// LDS.128, emissions are gathered by symbol from a transposed table E^T[sym][state]
// (one coalesced 256-byte row per column, L1-resident) two columns ahead of use, and
// symbols are read 32 at a time as coalesced uint16.  The forward/backward recursions
// are *scaled*, not log-space: every RESCALE columns the vector is multiplied by an
// exact power of two taken from the warp-max exponent (redux.sync), so no rounding is
// introduced by scaling and the log-normaliser is an integer exponent sum.
// Viterbi does only exactly-rounded FP64 adds and compares on host-provided log
// tables, first-maximum tie-breaking, so paths are bit-identical to the reference.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace itr {

constexpr int NSYM = 625;
constexpr int RESCALE = 8;        // columns between power-of-two rescalings
constexpr int VCHUNK = 256;       // Viterbi traceback chunk (columns)
constexpr unsigned FULL = 0xffffffffu;

struct ChainSet {
    const uint16_t *sym;     // all blocks back to back
    const int64_t *off;      // n_blocks + 1
    const int32_t *order;    // block ids, longest first
    int32_t n_blocks;
    int32_t n_sets;
    unsigned int *queue;     // work counter (zeroed before launch)
};

// ---------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Multiply the warp's non-negative vector by the power of two that brings its
// largest element into [1, 2); returns the exponent removed (0 if the vector is 0).
template <int NS>
__device__ __forceinline__ int rescale_pow2(double (&x)[NS]) {
    unsigned hi = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) hi = max(hi, (unsigned)__double2hiint(x[s]));
    hi = __reduce_max_sync(FULL, hi);
    const int ex = (int)(hi >> 20);            // biased exponent of the maximum (sign 0)
    if (ex == 0 || ex == 0x7ff) return 0;      // zero/denormal or inf/nan: leave alone
    const double sc = __hiloint2double((2046 - ex) << 20, 0);   // 2^(1023-ex)
#pragma unroll
    for (int s = 0; s < NS; ++s) x[s] *= sc;
    return ex - 1023;
}

__device__ __forceinline__ int next_chain(const ChainSet &cs, int lane) {
    unsigned c = 0;
    if (lane == 0) c = atomicAdd(cs.queue, 1u);
    return (int)__shfl_sync(FULL, c, 0);
}

// Column provider for the K x K matrix (a, or log a for Viterbi).  The matrix is
// stored padded on the device as [KP][KP] (pad value 0, or -inf for log a).
//   REGS = true  (K <= 32, NS == 1): lane j keeps column j in KT registers.
//   REGS = false (K  > 32): columns are re-read every step with coalesced,
//                           L1-resident loads (lane j reads A[i][j + 32 s]).
template <int KT, int NS, bool REGS>
struct Cols;

template <int KT>
struct Cols<KT, 1, true> {
    double c[KT];
    __device__ __forceinline__ void load(const double *Ap, int KP, int lane) {
#pragma unroll
        for (int i = 0; i < KT; ++i) c[i] = __ldg(Ap + (size_t)i * KP + lane);
    }
    __device__ __forceinline__ double get(int /*s*/, int i) const { return c[i]; }
};

template <int KT, int NS>
struct Cols<KT, NS, false> {
    const double *p;
    int KP;
    __device__ __forceinline__ void load(const double *Ap, int KP_, int lane) {
        p = Ap + lane;
        KP = KP_;
    }
    __device__ __forceinline__ double get(int s, int i) const {
        return __ldg(p + (size_t)i * KP + 32 * s);
    }
};

// y[s] = sum_i xs[i] * A[i][lane + 32 s].  xs: the exchanged vector in shared
// memory (16-byte aligned, zero beyond K).  KI = number of rows to visit (KT when the
// columns are in registers, K rounded up to 4 otherwise).
template <int KT, int NS, bool REGS>
__device__ __forceinline__ void matvec(const double *xs, const Cols<KT, NS, REGS> &cols, int K4,
                                       double (&y)[NS]) {
    double acc[NS][4];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.0;
    const double2 *x2 = reinterpret_cast<const double2 *>(xs);
    if (REGS) {
#pragma unroll
        for (int i = 0; i < KT; i += 4) {
            const double2 p = x2[i / 2], q = x2[i / 2 + 1];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                acc[s][0] = fma(p.x, cols.get(s, i + 0), acc[s][0]);
                acc[s][1] = fma(p.y, cols.get(s, i + 1), acc[s][1]);
                acc[s][2] = fma(q.x, cols.get(s, i + 2), acc[s][2]);
                acc[s][3] = fma(q.y, cols.get(s, i + 3), acc[s][3]);
            }
        }
    } else {
#pragma unroll 2
        for (int i = 0; i < K4; i += 4) {
            const double2 p = x2[i / 2], q = x2[i / 2 + 1];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                acc[s][0] = fma(p.x, cols.get(s, i + 0), acc[s][0]);
                acc[s][1] = fma(p.y, cols.get(s, i + 1), acc[s][1]);
                acc[s][2] = fma(q.x, cols.get(s, i + 2), acc[s][2]);
                acc[s][3] = fma(q.y, cols.get(s, i + 3), acc[s][3]);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) y[s] = (acc[s][0] + acc[s][1]) + (acc[s][2] + acc[s][3]);
}

// ---------------------------------------------------------------------------------
// Emission table: Et[set][sym][kp] = sum_{n in order[sym]} b[set][k][n]
// (optimizer.py:182 `b[:, order[V[t]]].sum(axis=1)` with read_data.py:46-67).
// digits[sym] packs the four base-5 digits (A,C,T,G,N = 0..4), 3 bits each.
// ---------------------------------------------------------------------------------
__global__ void emission_table_kernel(const double *__restrict__ b, const uint16_t *__restrict__ digits,
                                      double *__restrict__ Et, int K, int KP, int n_sets) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = n_sets * NSYM * KP;
    if (idx >= total) return;
    const int k = idx % KP;
    const int sym = (idx / KP) % NSYM;
    const int set = idx / (KP * NSYM);
    double acc = 0.0;
    if (k < K) {
        const double *row = b + ((size_t)set * K + k) * 256;
        const unsigned d = digits[sym];
        const int d0 = d & 7, d1 = (d >> 3) & 7, d2 = (d >> 6) & 7, d3 = (d >> 9) & 7;
        for (int a0 = (d0 == 4 ? 0 : d0); a0 <= (d0 == 4 ? 3 : d0); ++a0)
            for (int a1 = (d1 == 4 ? 0 : d1); a1 <= (d1 == 4 ? 3 : d1); ++a1)
                for (int a2 = (d2 == 4 ? 0 : d2); a2 <= (d2 == 4 ? 3 : d2); ++a2)
                    for (int a3 = (d3 == 4 ? 0 : d3); a3 <= (d3 == 4 ? 3 : d3); ++a3)
                        acc += row[64 * a0 + 16 * a1 + 4 * a2 + a3];
    }
    Et[idx] = acc;
}

// Transpose a host-provided K x 625 table (Viterbi's log E) into [625][KP], padding
// with `pad`.
__global__ void transpose_table_kernel(const double *__restrict__ src, double *__restrict__ dst,
                                       int K, int KP, double pad) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= NSYM * KP) return;
    const int k = idx % KP, sym = idx / KP;
    dst[idx] = k < K ? src[(size_t)k * NSYM + sym] : pad;
}

// ---------------------------------------------------------------------------------
// Symbol tile: 32 consecutive columns of a chain held one per lane.
// ---------------------------------------------------------------------------------
struct SymTile {
    const uint16_t *base;   // chain start
    int64_t T;
    __device__ __forceinline__ unsigned load(int64_t t0, int lane) const {
        const int64_t t = t0 + lane;
        return (t >= 0 && t < T) ? (unsigned)__ldg(base + t) : 0u;
    }
};

// ---------------------------------------------------------------------------------
// Forward recursion.  MODE 0: log-likelihood only.  MODE 1: also store the scaled
// alpha_t (any per-column power-of-two scale is fine: the posterior is normalised
// per column).
//   x_0 = pi * e(V_0);  x_t = (x_{t-1} @ a) * e(V_t)           optimizer.py:182-187
//   loglik = log(sum x_{T-1}) + ln2 * (sum of removed exponents)  optimizer.py:160-162
// ---------------------------------------------------------------------------------
template <int KT, int NS, bool REGS, int MODE>
__global__ void __launch_bounds__(256)
forward_kernel(ChainSet cs, const double *__restrict__ A, const double *__restrict__ PI,
               const double *__restrict__ Et, int K, int KP, double *__restrict__ loglik,
               double *__restrict__ alpha_out) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *xs = smem + (size_t)warp * 2 * KP;           // double buffer
    const int n_chains = cs.n_sets * cs.n_blocks;

    for (int c = next_chain(cs, lane); c < n_chains; c = next_chain(cs, lane)) {
        const int set = c / cs.n_blocks;
        const int blk = cs.order[c % cs.n_blocks];
        const int64_t beg = cs.off[blk], T = cs.off[blk + 1] - beg;
        const SymTile st{cs.sym + beg, T};
        const double *et = Et + (size_t)set * NSYM * KP;
        const int K4 = (K + 3) & ~3;
        Cols<KT, NS, REGS> acol;
        acol.load(A + (size_t)set * KP * KP, KP, lane);

        unsigned vcur = st.load(0, lane);
        unsigned vnxt = st.load(32, lane);
        double x[NS];
        {
            const unsigned v0 = __shfl_sync(FULL, vcur, 0);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int j = lane + 32 * s;
                x[s] = __ldg(PI + (size_t)set * KP + j) * __ldg(et + (size_t)v0 * KP + j);
            }
        }
        long long shift = 0;
        if (MODE == 1) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int j = lane + 32 * s;
                if (j < K) alpha_out[(size_t)beg * K + j] = x[s];
            }
        }
        // emission rows for columns t+1 and t+2 (software prefetch)
        double e1[NS], e2[NS];
        {
            const unsigned v1 = __shfl_sync(FULL, vcur, 1), v2 = __shfl_sync(FULL, vcur, 2);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                e1[s] = __ldg(et + (size_t)v1 * KP + lane + 32 * s);
                e2[s] = __ldg(et + (size_t)v2 * KP + lane + 32 * s);
            }
        }
        int buf = 0;
        for (int64_t t0 = 0; t0 < T; t0 += 32) {
#pragma unroll
            for (int s32 = 0; s32 < 32; ++s32) {
                const int64_t t = t0 + s32 + 1;        // column being produced
                if (t >= T) break;
                // exchange x
                double *xb = xs + buf * KP;
#pragma unroll
                for (int s = 0; s < NS; ++s) xb[lane + 32 * s] = x[s];
                __syncwarp();
                double y[NS];
                matvec<KT, NS, REGS>(xb, acol, K4, y);
                buf ^= 1;
#pragma unroll
                for (int s = 0; s < NS; ++s) x[s] = y[s] * e1[s];
                // rotate prefetch: column t+2
                const int pos = s32 + 3;               // (t+2) - t0
                const unsigned v = (pos < 32) ? __shfl_sync(FULL, vcur, pos & 31)
                                              : __shfl_sync(FULL, vnxt, pos & 31);
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    e1[s] = e2[s];
                    e2[s] = __ldg(et + (size_t)v * KP + lane + 32 * s);
                }
                if ((s32 & (RESCALE - 1)) == RESCALE - 1) shift += rescale_pow2<NS>(x);
                if (MODE == 1) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        const int j = lane + 32 * s;
                        if (j < K) alpha_out[(size_t)(beg + t) * K + j] = x[s];
                    }
                }
            }
            vcur = vnxt;
            vnxt = st.load(t0 + 64, lane);
        }
        double tot = 0.0;
#pragma unroll
        for (int s = 0; s < NS; ++s) tot += x[s];
        tot = warp_sum(tot);
        if (lane == 0 && loglik)
            loglik[(size_t)set * cs.n_blocks + blk] = log(tot) + (double)shift * 0.6931471805599453094;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Backward recursion fused with the posterior (reference orientation):
//   beta_{T-1} = 1;  beta_t = (beta_{t+1} * e(V_{t+1})) @ a        optimizer.py:205-212
//   post_t = alpha_t * beta_t / sum_j(alpha_t * beta_t)            optimizer.py:231-237
// `post` holds alpha on entry (written by forward_kernel<MODE 1>) and the posterior
// on exit, both (sum T, K) row-major.
// ---------------------------------------------------------------------------------
template <int KT, int NS, bool REGS>
__global__ void __launch_bounds__(256)
backward_posterior_kernel(ChainSet cs, const double *__restrict__ A, const double *__restrict__ Et,
                          int K, int KP, double *__restrict__ post) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *xs = smem + (size_t)warp * 2 * KP;
    const int n_chains = cs.n_blocks;      // set 0 only

    for (int c = next_chain(cs, lane); c < n_chains; c = next_chain(cs, lane)) {
        const int blk = cs.order[c];
        const int64_t beg = cs.off[blk], T = cs.off[blk + 1] - beg;
        const SymTile st{cs.sym + beg, T};
        double *pp = post + (size_t)beg * K;

        const int K4 = (K + 3) & ~3;
        Cols<KT, NS, REGS> acol;
        acol.load(A, KP, lane);
        double beta[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) beta[s] = (lane + 32 * s < K) ? 1.0 : 0.0;

        // walk tiles from the end: tile covers columns [t0, t0+32)
        const int64_t t_last = T - 1;
        int64_t t0 = t_last & ~(int64_t)31;
        unsigned vcur = st.load(t0, lane);
        unsigned vprv = st.load(t0 - 32, lane);
        // t = T-1: posterior = normalised alpha
        {
            double w[NS], tot = 0.0;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int j = lane + 32 * s;
                w[s] = (j < K) ? pp[(size_t)t_last * K + j] * beta[s] : 0.0;
                tot += w[s];
            }
            tot = warp_sum(tot);
            const double r = 1.0 / tot;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int j = lane + 32 * s;
                if (j < K) pp[(size_t)t_last * K + j] = w[s] * r;
            }
        }
        // prefetch: emission row of column t+1 (needed to produce beta_t) and alpha_t
        double e1[NS], a1[NS], a2[NS];
        {
            const unsigned v = __shfl_sync(FULL, vcur, (int)(t_last & 31));
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int j = lane + 32 * s;
                e1[s] = __ldg(Et + (size_t)v * KP + j);
                a1[s] = (j < K && t_last >= 1) ? pp[(size_t)(t_last - 1) * K + j] : 0.0;
                a2[s] = (j < K && t_last >= 2) ? pp[(size_t)(t_last - 2) * K + j] : 0.0;
            }
        }
        int buf = 0, cnt = 0;
        for (int64_t t = t_last - 1; t >= 0; --t) {
            // z = beta_{t+1} * e(V_{t+1})
            double *xb = xs + buf * KP;
#pragma unroll
            for (int s = 0; s < NS; ++s) xb[lane + 32 * s] = beta[s] * e1[s];
            __syncwarp();
            matvec<KT, NS, REGS>(xb, acol, K4, beta);
            buf ^= 1;
            if (((++cnt) & (RESCALE - 1)) == 0) (void)rescale_pow2<NS>(beta);
            // next emission row: column t (to produce beta_{t-1})
            if ((t & 31) == 31) {           // crossed into the previous tile
                t0 -= 32;
                vcur = vprv;
                vprv = st.load(t0 - 32, lane);
            }
            const unsigned v = __shfl_sync(FULL, vcur, (int)(t & 31));
            double w[NS], tot = 0.0;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int j = lane + 32 * s;
                e1[s] = __ldg(Et + (size_t)v * KP + j);
                w[s] = a1[s] * beta[s];
                tot += w[s];
                a1[s] = a2[s];
                a2[s] = (j < K && t >= 2) ? pp[(size_t)(t - 2) * K + j] : 0.0;
            }
            tot = warp_sum(tot);
            const double r = 1.0 / tot;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int j = lane + 32 * s;
                if (j < K) pp[(size_t)t * K + j] = w[s] * r;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Viterbi forward sweep (max-plus), bit-exact recipe:
//   m_ij = (omega_i + LA_ij) + LE_j ; prev_j = first argmax_i ; omega_j = max_i
//                                                               optimizer.py:325-332
// Outputs: backpointers bp[(beg+t)*KP + j] for t >= 1 (uint8), per-chunk composite
// maps comp[chunk][j] (state at the last column of chunk c -> state at the last
// column of chunk c-1) and the final state (first argmax of omega_{T-1},
// optimizer.py:347).
// ---------------------------------------------------------------------------------
template <int KT, int NS, bool REGS>
__global__ void __launch_bounds__(256)
viterbi_forward_kernel(ChainSet cs, const double *__restrict__ LA, const double *__restrict__ LEt,
                       const double *__restrict__ OM0, int K, int KP,
                       uint8_t *__restrict__ bp, uint8_t *__restrict__ comp,
                       const int64_t *__restrict__ chunk_off, int32_t *__restrict__ final_state) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *xs = smem + (size_t)warp * 2 * KP;
    const int n_chains = cs.n_blocks;

    for (int c = next_chain(cs, lane); c < n_chains; c = next_chain(cs, lane)) {
        const int blk = cs.order[c];
        const int64_t beg = cs.off[blk], T = cs.off[blk + 1] - beg;
        const SymTile st{cs.sym + beg, T};
        uint8_t *cmp = comp + (size_t)chunk_off[blk] * KP;

        const int K2 = (K + 1) & ~1;
        Cols<KT, NS, REGS> lacol;
        lacol.load(LA, KP, lane);
        double om[NS];
        int anc[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int j = lane + 32 * s;
            om[s] = __ldg(OM0 + (size_t)blk * KP + j);
            anc[s] = j;
        }
        unsigned vcur = st.load(0, lane);
        unsigned vnxt = st.load(32, lane);
        double e1[NS], e2[NS];
        {
            const unsigned v1 = __shfl_sync(FULL, vcur, 1), v2 = __shfl_sync(FULL, vcur, 2);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                e1[s] = __ldg(LEt + (size_t)v1 * KP + lane + 32 * s);
                e2[s] = __ldg(LEt + (size_t)v2 * KP + lane + 32 * s);
            }
        }
        int buf = 0;
        for (int64_t t0 = 0; t0 < T; t0 += 32) {
#pragma unroll
            for (int s32 = 0; s32 < 32; ++s32) {
                const int64_t t = t0 + s32 + 1;
                if (t >= T) break;
                double *xb = xs + buf * KP;
#pragma unroll
                for (int s = 0; s < NS; ++s) xb[lane + 32 * s] = om[s];
                __syncwarp();
                buf ^= 1;
                const double2 *x2 = reinterpret_cast<const double2 *>(xb);
                int arg[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) { om[s] = -CUDART_INF; arg[s] = 0; }
                auto step2 = [&](int i) {
                    const double2 p = x2[i / 2];
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        const double m0 = __dadd_rn(__dadd_rn(p.x, lacol.get(s, i)), e1[s]);
                        const double m1 = __dadd_rn(__dadd_rn(p.y, lacol.get(s, i + 1)), e1[s]);
                        if (i == 0) { om[s] = m0; arg[s] = 0; }
                        else if (m0 > om[s]) { om[s] = m0; arg[s] = i; }
                        if (m1 > om[s]) { om[s] = m1; arg[s] = i + 1; }
                    }
                };
                if (REGS) {
#pragma unroll
                    for (int i = 0; i < KT; i += 2) step2(i);
                } else {
#pragma unroll 2
                    for (int i = 0; i < K2; i += 2) step2(i);
                }
                // backpointers for column t
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const int j = lane + 32 * s;
                    if (j < KP) bp[(size_t)(beg + t) * KP + j] = (uint8_t)arg[s];
                }
                // chunk composite: anc_t[j] = anc_{t-1}[arg_j]  (reset at chunk start)
                const bool first = (t % VCHUNK) == 0;
                const bool last = ((t % VCHUNK) == VCHUNK - 1) || (t == T - 1);
                int na[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    // gather anc[arg]: arg may live in any lane / slot
                    int g = arg[s];
#pragma unroll
                    for (int q = 0; q < NS; ++q) {
                        const int got = __shfl_sync(FULL, anc[q], arg[s] & 31);
                        if ((arg[s] >> 5) == q) g = got;
                    }
                    na[s] = first ? arg[s] : g;
                }
#pragma unroll
                for (int s = 0; s < NS; ++s) anc[s] = na[s];
                if (last && t >= VCHUNK) {
                    const int64_t ch = t / VCHUNK;
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        const int j = lane + 32 * s;
                        if (j < KP) cmp[(size_t)ch * KP + j] = (uint8_t)anc[s];
                    }
                }
                const int pos = s32 + 3;
                const unsigned v = (pos < 32) ? __shfl_sync(FULL, vcur, pos & 31)
                                              : __shfl_sync(FULL, vnxt, pos & 31);
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    e1[s] = e2[s];
                    e2[s] = __ldg(LEt + (size_t)v * KP + lane + 32 * s);
                }
            }
            vcur = vnxt;
            vnxt = st.load(t0 + 64, lane);
        }
        // first argmax of omega_{T-1}
        double best = -CUDART_INF;
        int bi = 0x7fffffff;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int j = lane + 32 * s;
            if (j < K && (bi == 0x7fffffff || om[s] > best)) { best = om[s]; bi = j; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ob = __shfl_xor_sync(FULL, best, o);
            const int oi = __shfl_xor_sync(FULL, bi, o);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) {
                best = ob; bi = oi;
            }
        }
        if (lane == 0) final_state[blk] = bi;
        __syncwarp();
    }
}

// One thread per block: end state of every traceback chunk.
__global__ void viterbi_boundary_kernel(const int64_t *__restrict__ off, const int64_t *__restrict__ chunk_off,
                                        const uint8_t *__restrict__ comp, const int32_t *__restrict__ final_state,
                                        int KP, int n_blocks, uint8_t *__restrict__ chunk_end) {
    const int blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= n_blocks) return;
    const int64_t T = off[blk + 1] - off[blk];
    const int64_t nch = (T + VCHUNK - 1) / VCHUNK;
    const uint8_t *cmp = comp + (size_t)chunk_off[blk] * KP;
    uint8_t *ce = chunk_end + chunk_off[blk];
    int s = final_state[blk];
    ce[nch - 1] = (uint8_t)s;
    for (int64_t c = nch - 1; c >= 1; --c) {
        s = cmp[(size_t)c * KP + s];
        ce[c - 1] = (uint8_t)s;
    }
}

// One thread per traceback chunk: follow the backpointers inside the chunk from its
// known end state (optimizer.py:349-352) and write the path bytes.
__global__ void viterbi_traceback_kernel(const int64_t *__restrict__ off, const int64_t *__restrict__ chunk_off,
                                         const int32_t *__restrict__ chunk_blk,
                                         const uint8_t *__restrict__ bp, const uint8_t *__restrict__ chunk_end,
                                         int KP, int64_t n_chunks, uint8_t *__restrict__ path) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_chunks) return;
    const int blk = chunk_blk[g];
    const int64_t c = g - chunk_off[blk];
    const int64_t beg = off[blk], T = off[blk + 1] - beg;
    const int64_t ts = c * VCHUNK;
    const int64_t te = min(ts + VCHUNK, T) - 1;
    int s = chunk_end[g];
    uint8_t *p = path + beg;
    p[te] = (uint8_t)s;
    for (int64_t t = te; t > ts; --t) {
        s = bp[(size_t)(beg + t) * KP + s];
        p[t - 1] = (uint8_t)s;
    }
}

}  // namespace itr
