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

#include <type_traits>

#include "hmm_common.cuh"

namespace itr {

constexpr int RESCALE = 8;        // columns between power-of-two rescalings
constexpr int VCHUNK = 256;       // Viterbi traceback chunk (columns)

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
#ifndef ITR_MV_VARIANT
#define ITR_MV_VARIANT 3
#endif
#if ITR_MV_VARIANT == 0
        // compiler-scheduled
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
#else
        // The broadcast loads are fenced into ITR_MV_VARIANT groups with warp barriers
        // (ptxas does not move shared-memory loads across them) so that each group is
        // issued back to back and its ~30-cycle latency is paid once, overlapped with
        // the FMAs of the previous group, instead of once per three loads.
        constexpr int G = ITR_MV_VARIANT;
        constexpr int PER = ((KT / 4 + G - 1) / G) * 4;      // x values per group (multiple of 4)
        double xv[KT];
#pragma unroll
        for (int g = 0; g < G; ++g) {
#pragma unroll
            for (int i = g * PER; i < (g + 1) * PER && i < KT; i += 2) {
                const double2 p = x2[i / 2];
                xv[i] = p.x;
                xv[i + 1] = p.y;
            }
            __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < KT; i += 4) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                acc[s][0] = fma(xv[i + 0], cols.get(s, i + 0), acc[s][0]);
                acc[s][1] = fma(xv[i + 1], cols.get(s, i + 1), acc[s][1]);
                acc[s][2] = fma(xv[i + 2], cols.get(s, i + 2), acc[s][2]);
                acc[s][3] = fma(xv[i + 3], cols.get(s, i + 3), acc[s][3]);
            }
        }
#endif
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

// Symbol of tile position `pos` (0..63 over the current and the next tile).
__device__ __forceinline__ unsigned tile_symbol(unsigned vcur, unsigned vnxt, int pos) {
    const unsigned src = (pos < 32) ? vcur : vnxt;
    return __shfl_sync(FULL, src, pos & 31);
}

#ifndef ITR_UNROLL_FWD
#define ITR_UNROLL_FWD 8
#endif
#ifndef ITR_UNROLL_VIT
#define ITR_UNROLL_VIT 4
#endif
constexpr int UNROLL_FWD = ITR_UNROLL_FWD;   // columns unrolled per loop trip (code must stay in the i-cache)
constexpr int UNROLL_VIT = ITR_UNROLL_VIT;

// ---------------------------------------------------------------------------------
// Forward recursion.  MODE 0: log-likelihood only.  MODE 1: also store the scaled
// alpha_t (any per-column power-of-two scale is fine: the posterior is normalised
// per column).
//   x_0 = pi * e(V_0);  x_t = (x_{t-1} @ a) * e(V_t)           optimizer.py:182-187
//   loglik = log(sum x_{T-1}) + ln2 * (sum of removed exponents)  optimizer.py:160-162
// Per column the dependent chain is: STS x -> broadcast LDS.128 -> 4 FMA chains ->
// 2 adds -> multiply by the emission.  Everything else (symbol broadcast, emission
// gather two columns ahead, alpha store) is issued right after the exchange so its
// latency hides under the matvec.
// ---------------------------------------------------------------------------------
template <int KT, int NS, bool REGS, int MODE>
__global__ void __launch_bounds__(256)
forward_kernel(ChainSet cs, const double *__restrict__ A, const double *__restrict__ PI,
               const double *__restrict__ Et, int K, double *__restrict__ loglik,
               double *__restrict__ alpha_out) {
    constexpr int KP = 32 * NS;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *xs = smem + (size_t)warp * 2 * KP;           // double buffer
    const int n_chains = cs.n_sets * cs.n_blocks;
    const int K4 = (K + 3) & ~3;

    for (int c = next_chain(cs, lane); c < n_chains; c = next_chain(cs, lane)) {
        const int set = c / cs.n_blocks;
        const int blk = cs.order[c % cs.n_blocks];
        const int64_t beg = cs.off[blk], T = cs.off[blk + 1] - beg;
        const SymTile st{cs.sym + beg, T};
        const double *etl = Et + (size_t)set * NSYM * KP + lane;
        Cols<KT, NS, REGS> acol;
        acol.load(A + (size_t)set * KP * KP, KP, lane);

        unsigned vcur = st.load(0, lane);
        unsigned vnxt = st.load(32, lane);
        double x[NS], e1[NS], e2[NS];
        {
            const unsigned v0 = __shfl_sync(FULL, vcur, 0);
            const unsigned v1 = __shfl_sync(FULL, vcur, 1), v2 = __shfl_sync(FULL, vcur, 2);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                x[s] = __ldg(PI + (size_t)set * KP + lane + 32 * s) * __ldg(etl + v0 * KP + 32 * s);
                e1[s] = __ldg(etl + v1 * KP + 32 * s);
                e2[s] = __ldg(etl + v2 * KP + 32 * s);
            }
        }
        long long shift = 0;
        double *ao = (MODE == 1) ? alpha_out + (size_t)beg * K + lane : nullptr;
        if (MODE == 1) {
#pragma unroll
            for (int s = 0; s < NS; ++s)
                if (lane + 32 * s < K) ao[32 * s] = x[s];
        }
        int buf = 0;
        auto column = [&](int s32, int64_t t0) {
            const int64_t t = t0 + s32 + 1;            // column being produced
            double *xb = xs + buf * KP;
#pragma unroll
            for (int s = 0; s < NS; ++s) xb[lane + 32 * s] = x[s];
            __syncwarp();
            buf ^= 1;
            // emission row of column t+2 (independent of the chain)
            const unsigned v = tile_symbol(vcur, vnxt, s32 + 3);
            double e3[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) e3[s] = __ldg(etl + v * KP + 32 * s);
            double y[NS];
            matvec<KT, NS, REGS>(xb, acol, K4, y);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                x[s] = y[s] * e1[s];
                e1[s] = e2[s];
                e2[s] = e3[s];
            }
            if ((s32 & (RESCALE - 1)) == RESCALE - 1) shift += rescale_pow2<NS>(x);
            if (MODE == 1) {
#pragma unroll
                for (int s = 0; s < NS; ++s)
                    if (lane + 32 * s < K) ao[(size_t)t * K + 32 * s] = x[s];
            }
        };
        int64_t t0 = 0;
        for (; t0 + 32 < T; t0 += 32) {                 // whole tile in range
#pragma unroll UNROLL_FWD
            for (int s32 = 0; s32 < 32; ++s32) column(s32, t0);
            vcur = vnxt;
            vnxt = st.load(t0 + 64, lane);
        }
#pragma unroll 1
        for (int s32 = 0; t0 + s32 + 1 < T; ++s32) column(s32, t0);
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
// Run compression for the forward log-likelihood.
//
// In a four-species alignment ~97 % of the columns are invariant (AAAA, CCCC, TTTT,
// GGGG): under JC69 their emission columns are one vector up to rounding, so a run of n
// such columns multiplies the forward vector by (a diag(e))^n.  The kernel below applies
// a run greedily with the precomputed powers (a diag(e))^n, n in {2..8, 16, 24, 32} (2 to 32
// columns per step, at most two steps for a run of up to 39 columns; lane j streams column j of the power from L1), single run columns
// and all other columns one at a time with `a` from registers.  ~6x fewer dependent
// steps per block; the log-likelihood changes by ~1e-14
// relative (symbols are merged into one class only if their emission columns agree to
// 1e-12 relative in every parameter set; stated tolerance 1e-9).
//   symbol_class_kernel    rep[s] = smallest symbol with the same emission column
//   symbol_hist_kernel     symbol counts of the resident alignment (model independent)
//   pick_run_class_kernel  dominant class -> isrun[s], its representative, its share
//   run_power_kernel       the RUN_TABLE powers of a diag(e), each scaled by an exact power of two, per set
// ---------------------------------------------------------------------------------
constexpr int RUN_POWERS = 5;      // table slots 0..4: 2, 4, 8, 16, 32 columns per step
// Slots 5..9 hold the powers 3, 5, 6, 7 and 24, so that a run of n <= 39 columns takes at
// most two steps (one of 8/16/24/32, then the remainder 1..7) instead of one step per set
// bit of n: ~20 % fewer dependent steps per block on four-species alignments.
constexpr int RUN_TABLE = 10;
// columns to take from a run of nrun >= 2 columns, and the table slot of that power
__device__ __forceinline__ int run_take(int nrun) { return nrun >= 8 ? (min(nrun, 32) & ~7) : nrun; }
__device__ __forceinline__ int run_slot(int take) {
    // take < 8: nibble `take` of 0x87615000 (2->0 3->5 4->1 5->6 6->7 7->8);
    // else nibble take/8 of 0x49320 (8->2 16->3 24->9 32->4)
    return take >= 8 ? (0x49320u >> ((take >> 3) * 4)) & 15 : (0x87615000u >> (take * 4)) & 15;
}

__global__ void __launch_bounds__(640)
symbol_class_kernel(const double *__restrict__ Et, int K, int KP, int n_sets, double tol, int32_t *__restrict__ rep) {
    const int s = threadIdx.x;
    if (s >= NSYM) return;
    int c = s;
    const double *mine = Et + (size_t)s * KP;
    for (int o = 0; o < s && c == s; ++o) {
        const double *other = Et + (size_t)o * KP;
        bool same = true;
        for (int k = 0; k < K && same; ++k) {
            const double x = mine[k], y = other[k];
            same = fabs(x - y) <= tol * fmax(fabs(x), fabs(y));
        }
        if (same) c = o;
    }
    if (c != s)      // the equivalence must hold in every parameter set
        for (int q = 1; q < n_sets && c != s; ++q) {
            const double *m2 = mine + (size_t)q * NSYM * KP, *o2 = Et + ((size_t)q * NSYM + c) * KP;
            for (int k = 0; k < K; ++k) {
                const double x = m2[k], y = o2[k];
                if (!(fabs(x - y) <= tol * fmax(fabs(x), fabs(y)))) { c = s; break; }
            }
        }
    rep[s] = c;
}

__global__ void __launch_bounds__(256)
symbol_hist_kernel(const uint16_t *__restrict__ sym, int64_t n, unsigned long long *__restrict__ hist) {
    // bin NSYM counts symbols outside 0..624 (the load is rejected if it is not zero)
    __shared__ unsigned int h[NSYM + 1];
    for (int i = threadIdx.x; i <= NSYM; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned int v = sym[i];
        atomicAdd(&h[v < (unsigned)NSYM ? v : (unsigned)NSYM], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= NSYM; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[i], (unsigned long long)h[i]);
}

// info[0] = representative symbol of the dominant class, info[1] = its column count
__global__ void __launch_bounds__(640)
pick_run_class_kernel(const unsigned long long *__restrict__ hist, const int32_t *__restrict__ rep,
                      uint8_t *__restrict__ isrun, long long *__restrict__ info) {
    __shared__ unsigned long long cnt[NSYM];
    __shared__ int best;
    const int s = threadIdx.x;
    if (s < NSYM) cnt[s] = 0;
    __syncthreads();
    if (s < NSYM) atomicAdd(&cnt[rep[s]], hist[s]);
    __syncthreads();
    if (s == 0) {
        int b = 0;
        for (int i = 1; i < NSYM; ++i)
            if (cnt[i] > cnt[b]) b = i;
        best = b;
        info[0] = b;
        info[1] = (long long)cnt[b];
    }
    __syncthreads();
    if (s < NSYM) isrun[s] = rep[s] == best ? 1 : 0;
}

// One CTA (32 x 32 threads) per parameter set.  Powers are built as M^n = M^(n-1) M for
// n = 2..8, then M^16 = M^8 M^8, M^24 = M^16 M^8, M^32 = M^16 M^16; each is scaled by an
// exact power of two (max element into [1,2)) and its binary exponent recorded in sP.
__global__ void __launch_bounds__(1024)
run_power_kernel(const double *__restrict__ A, const double *__restrict__ Et, const long long *__restrict__ info,
                 int KP, int left, double *__restrict__ P, int32_t *__restrict__ sP, double *__restrict__ ebar) {
    __shared__ double M1[32][33], C[32][33], M8[32][33], M16[32][33];
    __shared__ unsigned int hi_max;
    const int set = blockIdx.x, i = threadIdx.y, j = threadIdx.x;
    const int r = (int)info[0];
    // left = 0: a diag(e) (forward runs); left = 1: diag(e) a (backward runs, reference orientation)
    const double e = Et[((size_t)set * NSYM + r) * KP + (left ? i : j)];
    M1[i][j] = A[((size_t)set * KP + i) * KP + j] * e;
    C[i][j] = M1[i][j];
    if (i == 0 && ebar) ebar[(size_t)set * KP + j] = Et[((size_t)set * NSYM + r) * KP + j];
    __syncthreads();
    // X Y scaled into (value of this thread's element, exponent removed)
    auto product = [&](const double (*X)[33], const double (*Y)[33], int &removed) {
        double acc = 0.0;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) acc = fma(X[i][k], Y[k][j], acc);
        if (i == 0 && j == 0) hi_max = 0;
        __syncthreads();
        atomicMax(&hi_max, (unsigned)__double2hiint(acc));
        __syncthreads();
        const int ex = (int)(hi_max >> 20);
        double sc = 1.0;
        removed = 0;
        if (ex != 0 && ex != 0x7ff) {
            sc = __hiloint2double((2046 - ex) << 20, 0);      // 2^(1023-ex): max element into [1,2)
            removed = ex - 1023;
        }
        return acc * sc;
    };
    auto store = [&](int take, double v, int shift) {
        const int slot = run_slot(take);
        P[(((size_t)set * RUN_TABLE + slot) * KP + i) * KP + j] = v;
        if (i == 0 && j == 0 && sP) sP[set * RUN_TABLE + slot] = shift;
    };
    int shift = 0, shift8 = 0, shift16 = 0, removed;
    for (int n = 2; n <= 8; ++n) {             // C = M^(n-1) -> M^n
        const double v = product(C, M1, removed);
        shift += removed;
        __syncthreads();                        // everyone has read C
        C[i][j] = v;
        if (n == 8) M8[i][j] = v;
        store(n, v, shift);
        __syncthreads();
    }
    shift8 = shift;
    double v = product(M8, M8, removed);
    shift16 = 2 * shift8 + removed;
    M16[i][j] = v;
    store(16, v, shift16);
    __syncthreads();
    v = product(M16, M8, removed);
    store(24, v, shift16 + shift8 + removed);
    __syncthreads();                            // hi_max is reused by the next product
    v = product(M16, M16, removed);
    store(32, v, 2 * shift16 + removed);
}

// Forward log-likelihood with run compression (K <= 32).  Same contract as
// forward_kernel<KT, 1, true, 0>.
template <int KT>
__global__ void __maxnreg__(192)
forward_runs_kernel(ChainSet cs, const double *__restrict__ A, const double *__restrict__ PI,
                    const double *__restrict__ Et, const double *__restrict__ P, const int32_t *__restrict__ sP,
                    const double *__restrict__ ebar, const uint8_t *__restrict__ isrun, int K,
                    double *__restrict__ loglik) {
    constexpr int KP = 32;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *xs = smem + (size_t)warp * 2 * KP;
    const int n_chains = cs.n_sets * cs.n_blocks;

    for (int c = next_chain(cs, lane); c < n_chains; c = next_chain(cs, lane)) {
        const int set = c / cs.n_blocks;
        const int blk = cs.order[c % cs.n_blocks];
        const int64_t beg = cs.off[blk], T = cs.off[blk + 1] - beg;
        const SymTile st{cs.sym + beg, T};
        const double *etl = Et + (size_t)set * NSYM * KP + lane;
        Cols<KT, 1, true> acol;
        acol.load(A + (size_t)set * KP * KP, KP, lane);
        const double *pset = P + (size_t)set * RUN_TABLE * KP * KP;
        const double eb = __ldg(ebar + (size_t)set * KP + lane);
        const int32_t *spset = sP + set * RUN_TABLE;
        // symbol tiles and their run masks are fetched one tile ahead of use
        auto run_mask = [&](unsigned v, int64_t t0) {
            return __ballot_sync(FULL, (t0 + lane < T) && __ldg(isrun + v));
        };
        unsigned vcur = st.load(0, lane), vnxt = st.load(32, lane), vnn = st.load(64, lane);
        unsigned mcur = run_mask(vcur, 0), mnxt = run_mask(vnxt, 32);
        double x[1];
        x[0] = __ldg(PI + (size_t)set * KP + lane) * __ldg(etl + __shfl_sync(FULL, vcur, 0) * KP);
        long long shift = 0;
        int buf = 0, steps = 0;
        int pos = 1;                               // position of the next column inside the current tile
        for (int64_t t0 = 0; t0 < T; t0 += 32) {
            // issued at the top of the tile so that their latency hides under the tile's steps:
            // the run flags of tile + 2 (gathered by symbol) and the symbols of tile + 3
            const bool fnn = (t0 + 64 + lane < T) && __ldg(isrun + vnn);
            const unsigned vn3 = st.load(t0 + 96, lane);
            const unsigned long long win = (unsigned long long)mcur | ((unsigned long long)mnxt << 32);
            const int end = (int)min((int64_t)32, T - t0);
            while (pos < end) {
                double *xb = xs + buf * KP;
                buf ^= 1;
                const int nrun = __ffsll((long long)~(win >> pos)) - 1;   // consecutive run columns from pos
                double y[1];
                // Everything that does not depend on x (which power, its column, the emission
                // row) is issued before x is exchanged, so it overlaps the previous step's tail.
                if (nrun >= 2) {                   // may reach into the next tile
                    const int take = run_take(nrun), k = run_slot(take);
                    Cols<KT, 1, true> pcol;
                    pcol.load(pset + (size_t)k * KP * KP, KP, lane);
                    const int spk = __ldg(spset + k);
                    xb[lane] = x[0];
                    __syncwarp();
                    matvec<KT, 1, true>(xb, pcol, KT, y);
                    x[0] = y[0];
                    shift += spk;
                    pos += take;
                } else if (nrun > 0) {
                    xb[lane] = x[0];
                    __syncwarp();
                    matvec<KT, 1, true>(xb, acol, KT, y);
                    x[0] = y[0] * eb;
                    pos += 1;
                } else {
                    const double e = __ldg(etl + __shfl_sync(FULL, vcur, pos) * KP);
                    xb[lane] = x[0];
                    __syncwarp();
                    matvec<KT, 1, true>(xb, acol, KT, y);
                    x[0] = y[0] * e;
                    pos += 1;
                }
                if ((++steps & 3) == 0) shift += rescale_pow2<1>(x);
            }
            pos -= 32;
            vcur = vnxt;
            vnxt = vnn;
            vnn = vn3;
            mcur = mnxt;
            mnxt = __ballot_sync(FULL, fnn);
        }
        double tot = warp_sum(x[0]);
        if (lane == 0 && loglik)
            loglik[(size_t)set * cs.n_blocks + blk] = log(tot) + (double)shift * 0.6931471805599453094;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Two-pass posterior (K <= 32, alignments with a dominant emission class).
//
// Pass 1 (latency bound, but short): run-compressed forward and backward sweeps, one warp
// per block, that only store *checkpoints* — the forward vector entering every 32-column
// tile and the backward vector at every tile's last column.  Power steps are confined to
// a tile so that every checkpoint is landed on exactly.
// Pass 2 (throughput bound): one warp per tile recomputes the 32 forward vectors of its
// tile from the checkpoint into shared memory, walks the tile backwards from the backward
// checkpoint forming alpha*beta, normalises per column and writes the posterior rows
// with coalesced stores.  HBM traffic is the posterior matrix itself (8K bytes/column)
// plus 2 x 8 bytes/column of checkpoints, instead of alpha + beta + three passes.
// Backward orientation is the reference's: beta_t = (beta_{t+1} * e(V_{t+1})) @ a
// (optimizer.py:210); in a run that is beta_{t+n} (diag(e) a)^n.
// ---------------------------------------------------------------------------------
constexpr int PTILE = 32;
#ifndef ITR_TILES_MINB
#define ITR_TILES_MINB 1
#endif

// DIR 0: forward checkpoints ck[tile] = alpha_{32 m - 1} (state entering tile m; tile 0 unused)
// DIR 1: backward checkpoints ck[tile] = beta_{32 m + 31} (last tile of a block: not stored, it is 1)
// MAXREG 192: no spills, one 8-warp CTA per SM — few chains, latency bound (config 2).
// MAXREG 128: two CTAs per SM — thousands of chains, throughput bound (config 4).
template <int KT, int DIR, int MAXREG>
__global__ void __maxnreg__(MAXREG)
checkpoint_sweep_kernel(ChainSet cs, const double *__restrict__ A, const double *__restrict__ PI,
                        const double *__restrict__ Et, const double *__restrict__ P, const double *__restrict__ ebar,
                        const uint8_t *__restrict__ isrun, const int64_t *__restrict__ tile_off, int K,
                        double *__restrict__ ck) {
    constexpr int KP = 32;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *xs = smem + (size_t)warp * 2 * KP;
    const int n_chains = cs.n_blocks;
    const double *etl = Et + lane;
    Cols<KT, 1, true> acol;
    acol.load(A, KP, lane);
    const double eb = __ldg(ebar + lane);

    for (int c = next_chain(cs, lane); c < n_chains; c = next_chain(cs, lane)) {
        const int blk = cs.order[c];
        const int64_t beg = cs.off[blk], T = cs.off[blk + 1] - beg;
        const SymTile st{cs.sym + beg, T};
        double *ckb = ck + (size_t)tile_off[blk] * KP + lane;
        const int64_t n_tiles = (T + PTILE - 1) / PTILE;
        auto run_mask = [&](unsigned v, int64_t t0) {
            return __ballot_sync(FULL, (t0 + lane >= 0) && (t0 + lane < T) && __ldg(isrun + v));
        };
        double x[1];
        int buf = 0, steps = 0;
        // one step with the power in table slot k (run_slot), or a single column
        auto power_step = [&](int k) {
            Cols<KT, 1, true> pcol;
            pcol.load(P + (size_t)k * KP * KP, KP, lane);
            double *xb = xs + buf * KP;
            buf ^= 1;
            xb[lane] = x[0];
            __syncwarp();
            double y[1];
            matvec<KT, 1, true>(xb, pcol, KT, y);
            x[0] = y[0];
            if ((++steps & 3) == 0) (void)rescale_pow2<1>(x);
        };
        if (DIR == 0) {
            unsigned vcur = st.load(0, lane), vnxt = st.load(PTILE, lane);
            unsigned mcur = run_mask(vcur, 0);
            x[0] = __ldg(PI + lane) * __ldg(etl + __shfl_sync(FULL, vcur, 0) * KP);
            int pos = 1;
            for (int64_t m = 0; m < n_tiles; ++m) {
                const int64_t t0 = m * PTILE;
                if (m > 0) ckb[(size_t)m * KP] = x[0];
                // next tile's run flags and the symbols of the tile after it: issued now, used at the tile's end
                const bool fn = (t0 + PTILE + lane < T) && __ldg(isrun + vnxt);
                const unsigned vn2 = st.load(t0 + 2 * PTILE, lane);
                const int end = (int)min((int64_t)PTILE, T - t0);
                while (pos < end) {
                    const unsigned rest = mcur >> pos;
                    const int nrun = (rest == 0xffffffffu) ? 32 : __ffs(~rest) - 1;
                    if (nrun >= 2) {
                        const int take = run_take(nrun);
                        power_step(run_slot(take));
                        pos += take;
                    } else {
                        const double e = nrun ? eb : __ldg(etl + __shfl_sync(FULL, vcur, pos) * KP);
                        double *xb = xs + buf * KP;
                        buf ^= 1;
                        xb[lane] = x[0];
                        __syncwarp();
                        double y[1];
                        matvec<KT, 1, true>(xb, acol, KT, y);
                        x[0] = y[0] * e;
                        if ((++steps & 3) == 0) (void)rescale_pow2<1>(x);
                        pos += 1;
                    }
                }
                pos = 0;
                vcur = vnxt;
                mcur = __ballot_sync(FULL, fn);
                vnxt = vn2;
            }
        } else {
            // walk tiles from the last to the first; inside a tile from high columns to low.
            // State: beta at column t_cur.  A step down by n consumes the emissions of columns
            // t_cur, t_cur - 1, ..., t_cur - n + 1.
            int64_t m = n_tiles - 1;
            unsigned vcur = st.load(m * PTILE, lane), vnxt = st.load((m - 1) * PTILE, lane);
            unsigned mcur = run_mask(vcur, m * PTILE);
            x[0] = (lane < K) ? 1.0 : 0.0;
            int pos = (int)(T - 1 - m * PTILE);              // position of t_cur inside the tile
            for (; m >= 0; --m) {
                if (m < n_tiles - 1) ckb[(size_t)m * KP] = x[0];       // beta_{32 m + 31}
                const int64_t tn = (m - 1) * PTILE;                     // next (lower) tile
                const bool fn = (tn + lane >= 0) && (tn + lane < T) && __ldg(isrun + vnxt);
                const unsigned vn2 = st.load((m - 2) * PTILE, lane);
                // columns of this tile still to consume: pos, pos-1, ..., 0  (column 0 of the block
                // is never consumed: beta_0 needs e(V_1) only)
                const int lowest = (m == 0) ? 1 : 0;
                while (pos >= lowest) {
                    const unsigned up = mcur << (31 - pos);                  // bit 31 = column `pos`
                    int nrun = (up == 0xffffffffu) ? 32 : __clz(~up);         // run columns pos, pos-1, ...
                    nrun = min(nrun, pos - lowest + 1);
                    if (nrun >= 2) {
                        const int take = run_take(nrun);
                        power_step(run_slot(take));
                        pos -= take;
                    } else {
                        const double e = nrun ? eb : __ldg(etl + __shfl_sync(FULL, vcur, pos) * KP);
                        double *xb = xs + buf * KP;
                        buf ^= 1;
                        xb[lane] = x[0] * e;
                        __syncwarp();
                        double y[1];
                        matvec<KT, 1, true>(xb, acol, KT, y);
                        x[0] = y[0];
                        if ((++steps & 3) == 0) (void)rescale_pow2<1>(x);
                        pos -= 1;
                    }
                }
                pos = PTILE - 1;
                vcur = vnxt;
                mcur = __ballot_sync(FULL, fn);
                vnxt = vn2;
            }
        }
        __syncwarp();
    }
}

// Pass 2: one warp per tile.  al: forward vectors of the tile, stride KP + 1.
template <int KT>
__global__ void __launch_bounds__(128, ITR_TILES_MINB)
posterior_tiles_kernel(const uint16_t *__restrict__ sym, const int64_t *__restrict__ off,
                       const int64_t *__restrict__ tile_off, const int32_t *__restrict__ tile_blk,
                       int64_t g_begin, int64_t g_end, const int64_t *__restrict__ list,
                       const double *__restrict__ A, const double *__restrict__ PI,
                       const double *__restrict__ Et, const double *__restrict__ ck_a,
                       const double *__restrict__ ck_b, int K, double *__restrict__ post) {
    // list == nullptr: the tiles [g_begin, g_end).  Else: the tile ids list[g_begin .. g_end),
    // negative entries skipped (the partial last tiles of a range of blocks, which the
    // tensor-core kernel below leaves to this one).
    constexpr int KP = 32, LD = KP + 1;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    double *xs = smem + (size_t)warp * (2 * KP + PTILE * LD + PTILE);
    double *al = xs + 2 * KP;
    double *rcp = al + PTILE * LD;
    const double *etl = Et + lane;
    Cols<KT, 1, true> acol;
    acol.load(A, KP, lane);
    // y_lane = sum_i xv[i] * a[i][lane]: this kernel is throughput bound (many warps per
    // scheduler), so the loads are left to the compiler's streaming schedule instead of the
    // register-hungry fenced groups of matvec<>
    auto dot = [&](const double *xv) {
        const double2 *x2 = reinterpret_cast<const double2 *>(xv);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int i = 0; i < KT; i += 4) {
            const double2 p = x2[i / 2], q = x2[i / 2 + 1];
            a0 = fma(p.x, acol.c[i], a0);
            a1 = fma(p.y, acol.c[i + 1], a1);
            a2 = fma(q.x, acol.c[i + 2], a2);
            a3 = fma(q.y, acol.c[i + 3], a3);
        }
        return (a0 + a1) + (a2 + a3);
    };

    for (int64_t gi = g_begin + (int64_t)blockIdx.x * nwarps + warp; gi < g_end; gi += (int64_t)gridDim.x * nwarps) {
        const int64_t g = list ? list[gi] : gi;
        if (g < 0) continue;
        const int blk = tile_blk[g];
        const int64_t m = g - tile_off[blk];
        const int64_t beg = off[blk], T = off[blk + 1] - beg;
        const int64_t t0 = m * PTILE;
        const int n = (int)min((int64_t)PTILE, T - t0);
        const unsigned v = (lane < n) ? (unsigned)__ldg(sym + beg + t0 + lane) : 0u;
        int buf = 0;
        // ---- forward through the tile
        double x[1];
        int first = 0;
        if (m == 0) {
            x[0] = __ldg(PI + lane) * __ldg(etl + __shfl_sync(FULL, v, 0) * KP);
            al[lane] = x[0];
            first = 1;
        } else {
            x[0] = __ldg(ck_a + (size_t)g * KP + lane);
        }
        double e_next = __ldg(etl + __shfl_sync(FULL, v, first) * KP);
        for (int i = first; i < n; ++i) {
            const double e = e_next;
            e_next = __ldg(etl + __shfl_sync(FULL, v, (i + 1) & 31) * KP);
            double *xb = xs + buf * KP;
            buf ^= 1;
            xb[lane] = x[0];
            __syncwarp();
            x[0] = dot(xb) * e;
            if ((i & 7) == 7) (void)rescale_pow2<1>(x);
            al[i * LD + lane] = x[0];
        }
        // ---- backward through the tile, forming alpha * beta in place
        double b[1];
        b[0] = (t0 + n == T) ? ((lane < K) ? 1.0 : 0.0) : __ldg(ck_b + (size_t)g * KP + lane);
        for (int i = n - 1; i >= 0; --i) {
            al[i * LD + lane] *= b[0];
            if (i == 0) break;
            const double e = __ldg(etl + __shfl_sync(FULL, v, i) * KP);
            double *xb = xs + buf * KP;
            buf ^= 1;
            xb[lane] = b[0] * e;
            __syncwarp();
            b[0] = dot(xb);
            if ((i & 7) == 0) (void)rescale_pow2<1>(b);
        }
        __syncwarp();
        // ---- normalise per column (lane = column) and write rows with coalesced stores
        if (lane < n) {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += al[lane * LD + k];
            rcp[lane] = 1.0 / s;
        }
        __syncwarp();
        double *out = post + (size_t)(beg + t0) * K;
        for (int e = lane; e < n * K; e += 32) {
            const int col = e / K, k = e - col * K;
            out[e] = al[col * LD + k] * rcp[col];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Pass 2 on the FP64 tensor cores (north_star (3): "batched (B x K)(K x K) contraction when
// many blocks advance in lockstep"; tcgen05 has no FP64 kind, so the FP64 tensor path is
// mma.sync.m8n8k4.f64).
//
// A warp takes FOUR consecutive tiles and walks each from BOTH ends at once: rows 0-3 of
// every 8-row fragment are the forward vectors of the four tiles, rows 4-7 their backward
// vectors — both directions are "row vector times a" in the reference's orientation
// (optimizer.py:187, :210), so one set of DMMAs advances all eight chains.  X (8 x K) times
// a (K x K) is NQ = KT/4 k-chunks by NC = ceil(KT/8) n-chunks of DMMAs with the NQ x NC
// fragments of `a` resident in registers (28 doubles at K = 27, what one column of `a` costs
// the FMA kernel).  States are placed so that NO data movement is needed between columns:
// the k side uses state 4q + c in chunk q for quad lane c (the A-fragment layout), and the
// n side is permuted — position n of chunk nc holds state 8 nc + 4 (n & 1) + (n >> 1) — so
// that the C fragment a thread receives (positions 2c, 2c + 1 of chunk nc) is exactly the
// pair of states 4(2nc) + c and 4(2nc + 1) + c: its A-fragment elements of k-chunks 2nc and
// 2nc + 1 for the next column.  Per step and warp: NQ x NC DMMAs and NQ multiplies by the
// emission; no shared-memory exchange, no shuffles on the chain (the FMA kernel: 1 STS + 14
// broadcast LDS.128 + 28 DFMA per tile-column).
// With f_t = alpha_{t-1} @ a (the forward DMMA's output, before the emission) and
// g_t = beta_t * e_t (the backward DMMA's input), the posterior of column t is f_t * g_t up
// to a factor — so both directions run the SAME recurrence x <- (x @ a) * e(next column in
// my direction), the forward rows handing over the DMMA's output and the backward rows its
// input.  For the first 16 steps every row parks its vector in shared memory, in the
// OUTPUT's layout (tile: 32 rows of K doubles); from step 16 on each row finds the other
// direction's vector waiting in the slot of the column it produces, multiplies, normalises
// inside the quad and leaves the posterior row in place; then the four tiles leave as
// plain coalesced copies.  Shared memory: 4 (32 K + 4) doubles per warp = 27.8 KB at
// K = 27, so EIGHT independent warps fit an SM — two per scheduler, which is what keeps the
// pipe busy: a DMMA blocks its warp's issue slot for its 16 pipe cycles (ptxas pads each
// with a NOP); measured, one warp per scheduler left the pipe 50 % idle.
// Only full 32-column tiles are stored; the last, partial tile of a block goes through
// posterior_tiles_kernel (list mode).
// ---------------------------------------------------------------------------------
// One 64-bit word per 32-column tile, built on the device when blocks are loaded, so that
// pass 2 needs ONE load per tile instead of the chain tile -> block -> offsets:
// bits 0..55 first column of the tile in the concatenated alignment, bit 60 tile starts a
// block, bit 61 tile ends a block, bit 62 tile has all 32 columns.  Also fills tile_blk.
constexpr unsigned long long TILE_FIRST = 1ull << 60, TILE_LAST = 1ull << 61, TILE_FULL = 1ull << 62,
                             TILE_COL_MASK = (1ull << 56) - 1;
__global__ void __launch_bounds__(256)
tile_table_kernel(const int64_t *__restrict__ off, const int64_t *__restrict__ tile_off, int n_blocks,
                  int32_t *__restrict__ tile_blk, unsigned long long *__restrict__ tile_info) {
    for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const int64_t beg = off[b], T = off[b + 1] - beg, g0 = tile_off[b], nt = tile_off[b + 1] - g0;
        for (int64_t m = threadIdx.x; m < nt; m += blockDim.x) {
            const int64_t t0 = m * PTILE;
            unsigned long long w = (unsigned long long)(beg + t0);
            if (m == 0) w |= TILE_FIRST;
            if (t0 + PTILE >= T) w |= TILE_LAST;
            if (T - t0 >= PTILE) w |= TILE_FULL;
            tile_info[g0 + m] = w;
            tile_blk[g0 + m] = b;
        }
    }
}

__device__ __forceinline__ void dmma_884(double &d0, double &d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}

template <int NQ>
__device__ __forceinline__ void quad_rescale(double (&x)[NQ]) {
    unsigned hi = 0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) hi = max(hi, (unsigned)__double2hiint(x[q]));
    hi = max(hi, __shfl_xor_sync(FULL, hi, 1));
    hi = max(hi, __shfl_xor_sync(FULL, hi, 2));
    const int ex = (int)(hi >> 20);
    if (ex != 0 && ex < 0x7ff) {
        const double sc = __hiloint2double((2046 - ex) << 20, 0);
#pragma unroll
        for (int q = 0; q < NQ; ++q) x[q] *= sc;
    }
}

#ifndef ITR_MMA_UNROLL
#define ITR_MMA_UNROLL 4
#endif
constexpr int MMA_UNROLL = ITR_MMA_UNROLL;
#ifndef ITR_MMA_WARPS
#define ITR_MMA_WARPS 4
#endif
// warps per CTA, four tiles each.  Four warps (111 KB of shared memory, 28 k registers): two
// CTAs fill an SM when the kernel is alone, and ONE still fits next to a Viterbi CTA when the
// recursions of a step overlap (an 8-warp CTA needs the whole SM and waits for the Viterbi
// sweep to end).
constexpr int MMA_WARPS = ITR_MMA_WARPS;

template <int KT>
__global__ void __launch_bounds__(32 * MMA_WARPS, 1)
posterior_tiles_mma_kernel(const uint16_t *__restrict__ sym, const unsigned long long *__restrict__ tile_info,
                           int64_t g_begin, int64_t g_end, const double *__restrict__ A,
                           const double *__restrict__ PI, const double *__restrict__ Et,
                           const double *__restrict__ ck_a, const double *__restrict__ ck_b, int K,
                           double *__restrict__ post, unsigned long long *__restrict__ ticket) {
    static_assert(KT % 4 == 0 && KT <= 32, "KT is K rounded up to a multiple of 4");
    constexpr int KP = 32, NQ = KT / 4, NC = (KT + 7) / 8, HALF = PTILE / 2;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = lane >> 2, c = lane & 3;
    const int tl = r & 3;                                 // tile of this row
    const bool bwd = r >= 4;                              // direction of this row
    const int TS = PTILE * K + 4;                         // doubles per tile; TS mod 16 == 4 staggers the tiles over the banks
    double *al = smem + (size_t)warp * 4 * TS;
    double *alr = al + (size_t)tl * TS + c;               // + column * K + 4q: state 4q + c of this row's tile
    // B fragments: element [k = c][n = r] of the (4 x 8) block (q, nc) of `a`
    double B[NQ][NC];
    {
        const int n_state = 4 * (r & 1) + (r >> 1);
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int nc = 0; nc < NC; ++nc) B[q][nc] = __ldg(A + (size_t)(4 * q + c) * KP + 8 * nc + n_state);
    }
    bool live[NQ];                                        // state 4q + c exists
    double pi0[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        live[q] = 4 * q + c < K;
        pi0[q] = __ldg(PI + 4 * q + c);
    }
    const double *etc = Et + c;

    // Groups of four tiles are handed out by a ticket counter (ticket != nullptr): inside a step
    // this kernel starts while the Viterbi sweep still holds most SMs, so its CTAs become
    // resident one by one over tens of milliseconds — with a static round-robin split the CTAs
    // that start last still own a full share and the launch ends a whole share after the
    // sweep (82 ms inside the step against 40 ms alone); with tickets whoever is resident works.
    // The ticket of the group after next is drawn at the top of a group, so the atomic's round
    // trip is hidden behind ~2 000 cycles of work.
    const int64_t stride = (int64_t)gridDim.x * MMA_WARPS * 4;
    auto next_group = [&](int64_t cur) -> int64_t {
        if (ticket == nullptr) return cur + stride;
        unsigned long long tk = 0ull;
        if (lane == 0) tk = atomicAdd(ticket, 1ull);
        tk = __shfl_sync(FULL, tk, 0);
        return g_begin + 4 * (int64_t)tk;
    };
    int64_t g0 = ticket ? next_group(0) : g_begin + ((int64_t)blockIdx.x * MMA_WARPS + warp) * 4;
    int64_t g1 = g0 < g_end ? next_group(g0) : g_end;
    // the tile word of the NEXT group is fetched a whole group ahead, and with it the first
    // symbols of that tile: the chain word -> symbol -> emission row is off the critical path
    auto tile_word = [&](int64_t gg0) { return gg0 + tl < g_end ? __ldg(tile_info + gg0 + tl) : (__ldg(tile_info + g_end - 1) & ~TILE_FULL); };
    unsigned long long w_next = g0 < g_end ? tile_word(g0) : 0ull;
    for (int64_t g2 = g_end; g0 < g_end; g0 = g1, g1 = g2) {
        const unsigned long long w = w_next;
        const int64_t g = min(g0 + tl, g_end - 1);
        g2 = g_end;
        if (g1 < g_end) {
            w_next = tile_word(g1);
            // next group's checkpoint vectors (256 B each, two lines): into L2 now
            if (c < 2) {
                const double *nk = (bwd ? ck_b : ck_a) + (size_t)min(g1 + tl, g_end - 1) * KP + 16 * c;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nk));
            }
            g2 = next_group(g1);
        }
        const bool full = (w & TILE_FULL) != 0;           // tiles that are stored
        const long long col0 = (long long)(w & TILE_COL_MASK);
        const bool last = (w & TILE_LAST) != 0;
        // (symbols may be read up to 33 columns past a short tile: into the next block or the
        // 64 zero columns behind the alignment — valid symbols either way, results discarded)
        const uint16_t *sp = sym + col0;
        const bool first = !bwd && (w & TILE_FIRST) != 0;  // forward row at the start of a block: f_0 := pi
        // Step s: the DMMA turns x into y — forward rows: f_s = alpha_{s-1} @ a; backward rows:
        // beta_{30-s} = g_{31-s} @ a — and x <- y * e(column c), c = s going forward (alpha_s),
        // c = 30 - s going backward (g_{30-s} = beta_{30-s} * e_{30-s}).  Both directions park
        // y for column c and both finish with the NEW x: going forward alpha_c * (parked
        // beta_c), going backward g_c * (parked f_c) = f_c * e_c * beta_c — the same product.
        // Forward parks columns 0..15 and finishes 16..31; backward parks 31 (its start
        // vector, before the loop), 30..16 and finishes 15..0 (nothing left at s = 31); the
        // only same-step hand-over is column 15 at s = 15.
        auto ecol = [&](int s) { return bwd ? max(PTILE - 2 - s, 0) : s; };
        double x[NQ], e[NQ];
        unsigned s_nxt = __ldg(sp + ecol(1));
        {
            const unsigned s0 = __ldg(sp + ecol(0)), s31 = __ldg(sp + PTILE - 1);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                e[q] = __ldg(etc + (size_t)s0 * KP + 4 * q);
                if (bwd) {   // beta_31 parked; g_31 = beta_31 * e_31
                    const double b31 = last ? (live[q] ? 1.0 : 0.0) : __ldg(ck_b + (size_t)g * KP + 4 * q + c);
                    if (live[q]) alr[(PTILE - 1) * K + 4 * q] = b31;
                    x[q] = b31 * __ldg(etc + (size_t)s31 * KP + 4 * q);
                } else {
                    x[q] = __ldg(ck_a + (size_t)g * KP + 4 * q + c);      // alpha_{-1} (tile 0 of a block: unused slot, finite)
                }
            }
        }
        // one step: y = x @ a (handed back for parking), x = y * e
        auto advance = [&](int s, double (&y)[2 * NC]) {
            const unsigned s_n2 = __ldg(sp + ecol(s + 2));
            double en[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) en[q] = __ldg(etc + (size_t)s_nxt * KP + 4 * q);
#pragma unroll
            for (int j = 0; j < 2 * NC; ++j) y[j] = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q)
#pragma unroll
                for (int nc = 0; nc < NC; ++nc) dmma_884(y[2 * nc], y[2 * nc + 1], x[q], B[q][nc]);
            if (s == 0 && first) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) y[q] = pi0[q];
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) x[q] = y[q] * e[q];
            if ((s & 7) == 7) quad_rescale<NQ>(x);
#pragma unroll
            for (int q = 0; q < NQ; ++q) e[q] = en[q];
            s_nxt = s_n2;
        };
        auto park = [&](double *slot, const double (&y)[2 * NC], bool on) {
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                if (on && live[q]) slot[4 * q] = y[q];
        };
        // posterior row of the slot's column: (parked vector of the other direction) * x, normalised
        auto finish = [&](double *slot, bool on) {
            double p[NQ], ps[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                p[q] = (on && live[q]) ? slot[4 * q] * x[q] : 0.0;
                ps[q & 3] += p[q];
            }
            double sum = (ps[0] + ps[1]) + (ps[2] + ps[3]);
            sum += __shfl_xor_sync(FULL, sum, 1);
            sum += __shfl_xor_sync(FULL, sum, 2);
            const double inv = 1.0 / sum;
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                if (on && live[q]) slot[4 * q] = p[q] * inv;
        };
        {
            double *slot = alr + (bwd ? (PTILE - 2) * K : 0);   // column 0 going forward, 30 going backward
            const int dslot = bwd ? -K : K;
#pragma unroll MMA_UNROLL
            for (int s = 0; s < HALF - 1; ++s) {          // park columns 0..14 / 30..16
                double y[2 * NC];
                advance(s, y);
                park(slot, y, true);
                slot += dslot;
            }
            {                                             // s = 15: forward parks column 15, backward finishes it
                double y[2 * NC];
                advance(HALF - 1, y);
                park(slot, y, !bwd);
                __syncwarp();                             // every parked vector is visible from here on
                finish(slot, bwd);
                slot += dslot;
            }
            if (c == 0) {                                 // next group's symbols (its tile word has arrived by now)
                const uint16_t *ns = sym + (w_next & TILE_COL_MASK);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ns));
            }
#pragma unroll MMA_UNROLL
            for (int s = HALF; s < PTILE; ++s) {          // finish columns 16..31 / 14..0
                double y[2 * NC];
                advance(s, y);
                finish(slot, !bwd || s < PTILE - 1);
                slot += dslot;
            }
        }
        __syncwarp();
        // ---- the four tiles leave as plain coalesced copies ----------------------------------------
#pragma unroll 1
        for (int rr = 0; rr < 4; ++rr) {
            if (!__shfl_sync(FULL, (int)full, 4 * rr)) continue;
            double *out = post + (size_t)__shfl_sync(FULL, col0, 4 * rr) * K;
            const double *src = al + (size_t)rr * TS;
            const int n = PTILE * K;
#pragma unroll 9
            for (int j = lane; j < n; j += 32) out[j] = src[j];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Backward recursion (reference orientation, optimizer.py:205-212):
//   beta_{T-1} = 1;  beta_t = (beta_{t+1} * e(V_{t+1})) @ a
// Stores the scaled beta_t as (sum T, K) row-major.  It runs concurrently with the
// alpha-storing forward kernel on a second stream; posterior_combine_kernel then
// forms the posterior.
// ---------------------------------------------------------------------------------
template <int KT, int NS, bool REGS>
__global__ void __launch_bounds__(256)
backward_kernel(ChainSet cs, const double *__restrict__ A, const double *__restrict__ Et, int K,
                double *__restrict__ beta_out) {
    constexpr int KP = 32 * NS;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *xs = smem + (size_t)warp * 2 * KP;
    const int n_chains = cs.n_blocks;      // set 0 only
    const int K4 = (K + 3) & ~3;
    const double *etl = Et + lane;

    for (int c = next_chain(cs, lane); c < n_chains; c = next_chain(cs, lane)) {
        const int blk = cs.order[c];
        const int64_t beg = cs.off[blk], T = cs.off[blk + 1] - beg;
        const SymTile st{cs.sym + beg, T};
        double *bo = beta_out + (size_t)beg * K + lane;
        Cols<KT, NS, REGS> acol;
        acol.load(A, KP, lane);

        double beta[NS], e1[NS], e2[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            beta[s] = (lane + 32 * s < K) ? 1.0 : 0.0;
            if (lane + 32 * s < K) bo[(size_t)(T - 1) * K + 32 * s] = beta[s];
        }
        // Walk backwards: step u = 0,1,... produces column t = T-2-u and needs the
        // emission row of column t+1 = T-1-u.  Tiles are indexed from the end of the
        // block: tile position p of tile u0 <-> column T-1-(u0+p).
        auto load_tile = [&](int64_t u0) -> unsigned {
            const int64_t t = T - 1 - (u0 + lane);
            return (t >= 0 && t < T) ? (unsigned)__ldg(st.base + t) : 0u;
        };
        unsigned vcur = load_tile(0), vnxt = load_tile(32);
        {
            const unsigned v1 = __shfl_sync(FULL, vcur, 0), v2 = __shfl_sync(FULL, vcur, 1);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                e1[s] = __ldg(etl + v1 * KP + 32 * s);     // column T-1
                e2[s] = __ldg(etl + v2 * KP + 32 * s);     // column T-2
            }
        }
        int buf = 0;
        auto column = [&](int s32, int64_t u0) {
            const int64_t t = T - 2 - (u0 + s32);          // column being produced
            double *xb = xs + buf * KP;
#pragma unroll
            for (int s = 0; s < NS; ++s) xb[lane + 32 * s] = beta[s] * e1[s];
            __syncwarp();
            buf ^= 1;
            // emission row for the step after next: column T-1-(u0+s32+2)
            const unsigned v = tile_symbol(vcur, vnxt, s32 + 2);
            double e3[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) e3[s] = __ldg(etl + v * KP + 32 * s);
            matvec<KT, NS, REGS>(xb, acol, K4, beta);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                e1[s] = e2[s];
                e2[s] = e3[s];
            }
            if ((s32 & (RESCALE - 1)) == RESCALE - 1) (void)rescale_pow2<NS>(beta);
#pragma unroll
            for (int s = 0; s < NS; ++s)
                if (lane + 32 * s < K) bo[(size_t)t * K + 32 * s] = beta[s];
        };
        int64_t u0 = 0;
        for (; u0 + 32 < T; u0 += 32) {                     // all 32 columns have t >= 0
#pragma unroll UNROLL_FWD
            for (int s32 = 0; s32 < 32; ++s32) column(s32, u0);
            vcur = vnxt;
            vnxt = load_tile(u0 + 64);
        }
#pragma unroll 1
        for (int s32 = 0; T - 2 - (u0 + s32) >= 0; ++s32) column(s32, u0);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Plain sweeps for 32 < K <= 32 NW (NW = 2, 3): one CTA of NW warps per chain, warp w
// owns states [32 w, 32 w + 32) and keeps their columns of `a` in registers (64 NW
// registers per lane).  Per column: every warp publishes its 32 values, one bar.sync,
// every lane reads all 32 NW values with broadcast LDS.128 and runs its dot product.
// The power-of-two rescaling uses the maximum exponent over all warps, exchanged with
// the values.  DIR 0: forward (MODE 0 log-likelihood, MODE 1 also stores alpha);
// DIR 1: backward in the reference's orientation (stores beta).
// ---------------------------------------------------------------------------------
template <int NW, int DIR, int MODE>
__global__ void __launch_bounds__(32 * NW)
sweep_mw_kernel(ChainSet cs, const double *__restrict__ A, const double *__restrict__ PI,
                const double *__restrict__ Et, int K, double *__restrict__ loglik, double *__restrict__ out) {
    constexpr int KP = 32 * NW;
    __shared__ __align__(16) double xs[2][KP];
    __shared__ unsigned hi_s[2][NW];
    __shared__ int chain_s;
    __shared__ double tot_s[NW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = 32 * warp + lane;                       // owned state
    const int n_chains = cs.n_sets * cs.n_blocks;

    for (;;) {
        if (threadIdx.x == 0) chain_s = (int)atomicAdd(cs.queue, 1u);
        __syncthreads();
        const int c = chain_s;
        __syncthreads();
        if (c >= n_chains) break;
        const int set = c / cs.n_blocks;
        const int blk = cs.order[c % cs.n_blocks];
        const int64_t beg = cs.off[blk], T = cs.off[blk + 1] - beg;
        const uint16_t *symp = cs.sym + beg;
        const double *etj = Et + (size_t)set * NSYM * KP + j;
        double col[KP];                                   // column j of a
#pragma unroll
        for (int i = 0; i < KP; ++i) col[i] = __ldg(A + ((size_t)set * KP + i) * KP + j);

        double x;
        long long shift = 0;
        int buf = 0;
        // x_new[j] = sum_i xs[i] * a[i][j]
        auto dot = [&](const double *xv) {
            const double2 *x2 = reinterpret_cast<const double2 *>(xv);
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int i = 0; i < KP; i += 4) {
                const double2 p = x2[i / 2], q = x2[i / 2 + 1];
                a0 = fma(p.x, col[i], a0);
                a1 = fma(p.y, col[i + 1], a1);
                a2 = fma(q.x, col[i + 2], a2);
                a3 = fma(q.y, col[i + 3], a3);
            }
            return (a0 + a1) + (a2 + a3);
        };
        // publish `v` (and, every 8th column, this warp's maximum exponent); returns the common
        // power-of-two scale to apply to the *next* vector (1.0 when no rescale is due)
        auto exchange = [&](double v, bool with_max) {
            xs[buf][j] = v;
            if (with_max) {
                const unsigned hi = __reduce_max_sync(FULL, (unsigned)__double2hiint(v));
                if (lane == 0) hi_s[buf][warp] = hi;
            }
            __syncthreads();
        };
        auto common_scale = [&](int *removed) {
            unsigned hi = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) hi = max(hi, hi_s[buf][w]);
            const int ex = (int)(hi >> 20);
            if (ex == 0 || ex == 0x7ff) { *removed = 0; return 1.0; }
            *removed = ex - 1023;
            return __hiloint2double((2046 - ex) << 20, 0);
        };
        if (DIR == 0) {
            x = __ldg(PI + (size_t)set * KP + j) * __ldg(etj + (unsigned)__ldg(symp) * KP);
            double *ao = (MODE == 1) ? out + (size_t)beg * K + j : nullptr;
            if (MODE == 1 && j < K) ao[0] = x;
            double e_next = __ldg(etj + (unsigned)__ldg(symp + 1) * KP);
            for (int64_t t = 1; t < T; ++t) {
                const double e = e_next;
                e_next = __ldg(etj + (unsigned)__ldg(symp + t + 1) * KP);   // (64 columns of slack)
                const bool resc = (t & 7) == 0;
                exchange(x, resc);
                double y = dot(xs[buf]);
                if (resc) {
                    int removed;
                    y *= common_scale(&removed);          // scale of the previous vector, applied here
                    shift += removed;
                }
                x = y * e;
                buf ^= 1;
                if (MODE == 1 && j < K) ao[(size_t)t * K] = x;
            }
            const double part = warp_sum(x);
            if (lane == 0) tot_s[warp] = part;
            __syncthreads();
            if (threadIdx.x == 0 && loglik) {
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) tot += tot_s[w];
                loglik[(size_t)set * cs.n_blocks + blk] = log(tot) + (double)shift * 0.6931471805599453094;
            }
        } else {
            double *bo = out + (size_t)beg * K + j;
            x = (j < K) ? 1.0 : 0.0;
            if (j < K) bo[(size_t)(T - 1) * K] = x;
            double e_next = __ldg(etj + (unsigned)__ldg(symp + T - 1) * KP);
            for (int64_t t = T - 2; t >= 0; --t) {
                const double e = e_next;                  // emission of column t + 1
                e_next = __ldg(etj + (unsigned)__ldg(symp + (t > 0 ? t : 0)) * KP);
                const bool resc = (t & 7) == 0;
                exchange(x * e, resc);
                double y = dot(xs[buf]);
                if (resc) {
                    int removed;
                    y *= common_scale(&removed);
                }
                x = y;
                buf ^= 1;
                if (j < K) bo[(size_t)t * K] = x;
            }
        }
        __syncthreads();
    }
}

// post[t][j] = alpha[t][j] * beta[t][j] / sum_j(alpha[t][j] * beta[t][j])
// (optimizer.py:231-237; scales cancel).  `post` holds alpha on entry.  HBM-bound:
// a CTA stages COLS columns through shared memory with coalesced loads and stores.
template <int COLS>
__global__ void __launch_bounds__(COLS)
posterior_combine_kernel(double *__restrict__ post, const double *__restrict__ beta, int K, int64_t n_cols) {
    extern __shared__ __align__(16) double sm[];            // COLS * K products
    const int64_t c0 = (int64_t)blockIdx.x * COLS;
    const int nc = (int)min((int64_t)COLS, n_cols - c0);
    if (nc <= 0) return;
    const int n = nc * K;
    double *p = post + (size_t)c0 * K;
    const double *b = beta + (size_t)c0 * K;
    for (int i = threadIdx.x; i < n; i += COLS) sm[i] = p[i] * b[i];
    __syncthreads();
    __shared__ double rcp[COLS];
    if ((int)threadIdx.x < nc) {
        const double *r = sm + (size_t)threadIdx.x * K;
        double acc = 0.0;
        for (int j = 0; j < K; ++j) acc += r[j];
        rcp[threadIdx.x] = 1.0 / acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += COLS) p[i] = sm[i] * rcp[i / K];
}

// Same, for a group of blocks (blockIdx.y indexes `order`): used when the posterior is
// produced group by group so that the download of finished (shorter) blocks overlaps
// the recursions of the longer ones.
template <int COLS>
__global__ void __launch_bounds__(COLS)
posterior_combine_blocks_kernel(double *__restrict__ post, const double *__restrict__ beta, int K,
                                const int64_t *__restrict__ off, const int32_t *__restrict__ order) {
    extern __shared__ __align__(16) double sm[];
    const int blk = order[blockIdx.y];
    const int64_t c0 = off[blk] + (int64_t)blockIdx.x * COLS;
    const int nc = (int)min((int64_t)COLS, off[blk + 1] - c0);
    if (nc <= 0) return;
    const int n = nc * K;
    double *p = post + (size_t)c0 * K;
    const double *b = beta + (size_t)c0 * K;
    for (int i = threadIdx.x; i < n; i += COLS) sm[i] = p[i] * b[i];
    __syncthreads();
    __shared__ double rcp[COLS];
    if ((int)threadIdx.x < nc) {
        const double *r = sm + (size_t)threadIdx.x * K;
        double acc = 0.0;
        for (int j = 0; j < K; ++j) acc += r[j];
        rcp[threadIdx.x] = 1.0 / acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += COLS) p[i] = sm[i] * rcp[i / K];
}

// ---------------------------------------------------------------------------------
// Viterbi forward sweep (max-plus), bit-exact recipe:
//   m_ij = (omega_i + LA_ij) + LE_j ; prev_j = first argmax_i ; omega_j = max_i
//                                                               optimizer.py:325-332
// The add of LE_j is hoisted out of the scan without changing any decision:
// f(s) = fl(s + LE_j) is monotone, so max_i m_ij = f(max_i s_ij) with
// s_ij = fl(omega_i + LA_ij), and the first maximiser of m is the first maximiser of
// s unless rounding merges s* with a smaller neighbour, i.e. unless
// f(pred(s*)) == f(s*).  That (rare, ~2 % of columns) case is detected per lane and
// the whole warp redoes the column with the literal two-add scan (out of line).
// The first maximiser of s is found with a tournament over (value, index) pairs: all
// K sums are formed first (independent adds), then log2(K) levels of "right wins only
// if strictly greater" — the left operand always holds the smaller indices, so this is
// np.argmax's first-maximum rule — which keeps the dependent chain at 5 compare/select
// links instead of K.  Everything is branch-free selects.
// Outputs: backpointers bp[(beg+t)*KP + j] for t >= 1 (uint8) and the final state
// (first argmax of omega_{T-1}, optimizer.py:347).  The traceback is parallel
// (viterbi_compose_kernel / viterbi_boundary_kernel / viterbi_traceback_kernel).
// ---------------------------------------------------------------------------------

// Literal scan for one output state (column `lac` of the padded log-a matrix, read
// from global memory): m_i = (omega_i + LA_ij) + LE_j, first maximum.
struct ScanResult {
    double best;
    int arg;
};
__device__ __forceinline__ ScanResult viterbi_exact_scan(const double *xb, const double *lac, int KP, int K4, double le) {
    double best = __dadd_rn(__dadd_rn(xb[0], __ldg(lac)), le);
    int arg = 0;
#pragma unroll 1
    for (int i = 1; i < K4; ++i) {
        const double m = __dadd_rn(__dadd_rn(xb[i], __ldg(lac + (size_t)i * KP)), le);
        const bool g = m > best;
        best = g ? m : best;
        arg = g ? i : arg;
    }
    return ScanResult{best, arg};
}

// In-place tournament over v[0..N): afterwards v[0] is the maximum and ix[0] the index
// of its first occurrence.  Compile-time recursion so that every index is static.
template <int N, int CUR>
struct Tournament {
    __device__ __forceinline__ static void run(double (&v)[N], int (&ix)[N]) {
        if constexpr (CUR > 1) {
#pragma unroll
            for (int m = 0; m < CUR / 2; ++m) {
                const bool g = v[2 * m + 1] > v[2 * m];
                v[m] = g ? v[2 * m + 1] : v[2 * m];
                ix[m] = g ? ix[2 * m + 1] : ix[2 * m];
            }
            if constexpr (CUR & 1) {
                v[CUR / 2] = v[CUR - 1];
                ix[CUR / 2] = ix[CUR - 1];
            }
            Tournament<N, (CUR + 1) / 2>::run(v, ix);
        }
    }
};
template <int N>
__device__ __forceinline__ void tournament(double (&v)[N], int (&ix)[N]) {
    Tournament<N, N>::run(v, ix);
}

template <int KT, int NS, bool REGS>
__global__ void __launch_bounds__(256)
viterbi_forward_kernel(ChainSet cs, const double *__restrict__ LA, const double *__restrict__ LEt,
                       const double *__restrict__ OM0, int K,
                       uint8_t *__restrict__ bp, int32_t *__restrict__ final_state) {
    constexpr int KP = 32 * NS;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *xs = smem + (size_t)warp * 2 * KP;
    const int n_chains = cs.n_blocks;
    const int K4 = (K + 3) & ~3;
    const double *etl = LEt + lane;

    for (int c = next_chain(cs, lane); c < n_chains; c = next_chain(cs, lane)) {
        const int blk = cs.order[c];
        const int64_t beg = cs.off[blk], T = cs.off[blk + 1] - beg;
        const SymTile st{cs.sym + beg, T};
        uint8_t *bpl = bp + (size_t)beg * KP + lane;

        Cols<KT, NS, REGS> lacol;
        lacol.load(LA, KP, lane);
        double om[NS], e1[NS], e2[NS];
        unsigned vcur = st.load(0, lane);
        unsigned vnxt = st.load(32, lane);
        {
            const unsigned v1 = __shfl_sync(FULL, vcur, 1), v2 = __shfl_sync(FULL, vcur, 2);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                om[s] = __ldg(OM0 + (size_t)blk * KP + lane + 32 * s);
                e1[s] = __ldg(etl + v1 * KP + 32 * s);
                e2[s] = __ldg(etl + v2 * KP + 32 * s);
            }
        }
        int buf = 0;
        unsigned vpre = tile_symbol(vcur, vnxt, 3);      // symbol of the column two ahead
        uint8_t *bpt = bpl + KP;                          // row of column t = 1
        auto column = [&](int s32, int64_t t0) {
            double *xb = xs + buf * KP;
#pragma unroll
            for (int s = 0; s < NS; ++s) xb[lane + 32 * s] = om[s];
            __syncwarp();
            buf ^= 1;
            // emission row two columns ahead; its symbol was shuffled out one column ago
            double e3[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) e3[s] = __ldg(etl + vpre * KP + 32 * s);
            vpre = tile_symbol(vcur, vnxt, s32 + 4);

            const double2 *x2 = reinterpret_cast<const double2 *>(xb);
            double sstar[NS];
            int arg[NS];
            if (REGS) {
                double sv[KT];
                int ix[KT];
#pragma unroll
                for (int i = 0; i < KT; i += 2) {
                    const double2 p = x2[i / 2];
                    sv[i] = __dadd_rn(p.x, lacol.get(0, i));
                    sv[i + 1] = __dadd_rn(p.y, lacol.get(0, i + 1));
                    ix[i] = i;
                    ix[i + 1] = i + 1;
                }
                tournament<KT>(sv, ix);
                sstar[0] = sv[0];
                arg[0] = ix[0];
            } else {
                // K > 32: four running first-maxima (i mod 4) per owned state, merged with
                // "larger value, then smaller index"
                double b0[NS], b1[NS], b2[NS], b3[NS];
                int i0[NS], i1[NS], i2[NS], i3[NS];
                {
                    const double2 p = x2[0], q = x2[1];
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        b0[s] = __dadd_rn(p.x, lacol.get(s, 0)); i0[s] = 0;
                        b1[s] = __dadd_rn(p.y, lacol.get(s, 1)); i1[s] = 1;
                        b2[s] = __dadd_rn(q.x, lacol.get(s, 2)); i2[s] = 2;
                        b3[s] = __dadd_rn(q.y, lacol.get(s, 3)); i3[s] = 3;
                    }
                }
#pragma unroll 2
                for (int i = 4; i < K4; i += 4) {
                    const double2 p = x2[i / 2], q = x2[i / 2 + 1];
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        const double s0 = __dadd_rn(p.x, lacol.get(s, i));
                        const double s1 = __dadd_rn(p.y, lacol.get(s, i + 1));
                        const double s2 = __dadd_rn(q.x, lacol.get(s, i + 2));
                        const double s3 = __dadd_rn(q.y, lacol.get(s, i + 3));
                        const bool g0 = s0 > b0[s], g1 = s1 > b1[s], g2 = s2 > b2[s], g3 = s3 > b3[s];
                        b0[s] = g0 ? s0 : b0[s]; i0[s] = g0 ? i : i0[s];
                        b1[s] = g1 ? s1 : b1[s]; i1[s] = g1 ? i + 1 : i1[s];
                        b2[s] = g2 ? s2 : b2[s]; i2[s] = g2 ? i + 2 : i2[s];
                        b3[s] = g3 ? s3 : b3[s]; i3[s] = g3 ? i + 3 : i3[s];
                    }
                }
                auto vmerge = [](double &va, int &ia, double vb, int ib) {
                    const bool take = (vb > va) | ((vb == va) & (ib < ia));
                    va = take ? vb : va;
                    ia = take ? ib : ia;
                };
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    vmerge(b0[s], i0[s], b1[s], i1[s]);
                    vmerge(b2[s], i2[s], b3[s], i3[s]);
                    vmerge(b0[s], i0[s], b2[s], i2[s]);
                    sstar[s] = b0[s];
                    arg[s] = i0[s];
                }
            }
            bool slow = false;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const double M = __dadd_rn(sstar[s], e1[s]);
                // pred(s*): next double towards -inf (finite, non-zero s* only)
                const long long bits = __double_as_longlong(sstar[s]);
                const double pred = __longlong_as_double(bits - ((bits >> 63) | 1));
                const bool odd = (sstar[s] == 0.0) | !(fabs(sstar[s]) < CUDART_INF);
                slow |= (lane + 32 * s < K) & (odd | (__dadd_rn(pred, e1[s]) == M));
                om[s] = M;
            }
            if (__builtin_expect(__any_sync(FULL, slow), 0)) {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const ScanResult r = viterbi_exact_scan(xb, LA + lane + 32 * s, KP, K4, e1[s]);
                    om[s] = r.best;
                    arg[s] = r.arg;
                }
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                e1[s] = e2[s];
                e2[s] = e3[s];
                bpt[32 * s] = (uint8_t)arg[s];
            }
            bpt += KP;
        };
        int64_t t0 = 0;
        for (; t0 + 32 < T; t0 += 32) {
#pragma unroll UNROLL_VIT
            for (int s32 = 0; s32 < 32; ++s32) column(s32, t0);
            vcur = vnxt;
            vnxt = st.load(t0 + 64, lane);
        }
#pragma unroll 1
        for (int s32 = 0; t0 + s32 + 1 < T; ++s32) column(s32, t0);
        // first argmax of omega_{T-1}
        double best = -CUDART_INF;
        int bidx = 0x7fffffff;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int j = lane + 32 * s;
            if (j < K && (bidx == 0x7fffffff || om[s] > best)) { best = om[s]; bidx = j; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ob = __shfl_xor_sync(FULL, best, o);
            const int oi = __shfl_xor_sync(FULL, bidx, o);
            if (oi != 0x7fffffff && (bidx == 0x7fffffff || ob > best || (ob == best && oi < bidx))) {
                best = ob; bidx = oi;
            }
        }
        if (lane == 0) final_state[blk] = bidx;
        __syncwarp();
    }
}

__device__ __forceinline__ bool viterbi_hoist_unsafe(double sstar, double le, double M);
// |M| < 2^36 (false for inf and NaN): with a screen-proven margin the hoist is then safe
__device__ __forceinline__ bool viterbi_small_for_hoist(double M) {
    return ((unsigned)__double2hiint(M) & 0x7fffffffu) < 0x42300000u;
}

// ---------------------------------------------------------------------------------
// Viterbi forward sweep for MANY chains (K <= 32): verify the cached pointer first.
//
// With thousands of blocks (config 4: 2 500 chains on 592 schedulers) the sweep is bound
// by issue slots and the FP64 pipe, not by latency, and the K-way arg-max of
// viterbi_forward_kernel costs ~200 instructions per column (a compare and three selects
// per predecessor).  The backpointer vector changes in only ~2 % of the columns, so lane j
// keeps its previous pointer p_j and only CHECKS it: all K sums s_i = fl(omega_i + log a_ij)
// are formed as before, but each is merely tested against s_p (one subtraction whose sign
// bit is OR-accumulated: no selects, no index bookkeeping; p itself is taken out of the
// scan by a -inf in the resident column of log a).  The pointer is kept iff s_p is the strict, unique maximum (no
// other s_i >= s_p: any tie goes to the slow path) and the emission add can be hoisted (same test as viterbi_forward_kernel); then the
// column's result is, by construction, what the full scan returns: same adds, same first
// maximiser.  If any lane fails, the whole warp redoes the column with the exact scan
// (tournament, hoisting test, literal two-add fallback) and refreshes its pointers.
// Bit-identical to viterbi_forward_kernel; ~2.3x fewer instructions per column.
// ---------------------------------------------------------------------------------
// One exact column for the calling warp (all 32 lanes): tournament over all predecessors
// with the column of log a read back from L1, the hoisting test, and the literal two-add
// scan when the test fires.  Cold path of viterbi_check_kernel.
template <int KT>
__device__ __noinline__ ScanResult viterbi_full_column(const double *xb, const double *lac, int K, int K4, double le) {
    constexpr int KP = 32;
    const int lane = threadIdx.x & 31;
    const double2 *x2 = reinterpret_cast<const double2 *>(xb);
    double sv[KT];
    int ix[KT];
#pragma unroll
    for (int i = 0; i < KT; i += 2) {
        const double2 pq = x2[i / 2];
        sv[i] = __dadd_rn(pq.x, __ldg(lac + (size_t)i * KP));
        sv[i + 1] = __dadd_rn(pq.y, __ldg(lac + (size_t)(i + 1) * KP));
        ix[i] = i;
        ix[i + 1] = i + 1;
    }
    tournament<KT>(sv, ix);
    ScanResult r{__dadd_rn(sv[0], le), ix[0]};
    if (__any_sync(FULL, (lane < K) & viterbi_hoist_unsafe(sv[0], le, r.best))) r = viterbi_exact_scan(xb, lac, KP, K4, le);
    return r;
}

template <int KT>
__global__ void __launch_bounds__(128, 3)
viterbi_check_kernel(ChainSet cs, const double *__restrict__ LA, const double *__restrict__ LEt,
                     const double *__restrict__ OM0, int K,
                     uint8_t *__restrict__ bp, int32_t *__restrict__ final_state) {
    constexpr int KP = 32;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *xs = smem + (size_t)warp * 2 * KP;
    const int n_chains = cs.n_blocks;
    const int K4 = (K + 3) & ~3;
    const double *etl = LEt + lane;
    // Column `lane` of log a with the entry of the cached pointer replaced by -inf: the K
    // sums of a column then cover every predecessor EXCEPT p, and "is s_p the strict, unique
    // maximum" is one OR-accumulated compare per predecessor — no counting, no selects.
    double lac[KT];
    auto load_column_without = [&](int p) {
#pragma unroll
        for (int i = 0; i < KT; ++i) lac[i] = (i == p) ? -CUDART_INF : __ldg(LA + (size_t)i * KP + lane);
    };

    for (int c = next_chain(cs, lane); c < n_chains; c = next_chain(cs, lane)) {
        const int blk = cs.order[c];
        const int64_t beg = cs.off[blk], T = cs.off[blk + 1] - beg;
        const SymTile st{cs.sym + beg, T};
        uint8_t *bpt = bp + (size_t)beg * KP + lane + KP;        // row of column t = 1

        int p = lane;                                            // cached pointer: "stay"
        double la_p = __ldg(LA + (size_t)lane * KP + lane);
        load_column_without(p);
        unsigned vcur = st.load(0, lane), vnxt = st.load(32, lane);
        double om = __ldg(OM0 + (size_t)blk * KP + lane);
        double e1 = __ldg(etl + __shfl_sync(FULL, vcur, 1) * KP), e2 = __ldg(etl + __shfl_sync(FULL, vcur, 2) * KP);
        unsigned vpre = tile_symbol(vcur, vnxt, 3);              // symbol of the column two ahead
        int buf = 0;
        auto column = [&](int s32) {
            double *xb = xs + buf * KP;
            xb[lane] = om;
            __syncwarp();
            buf ^= 1;
            const double e3 = __ldg(etl + vpre * KP);
            vpre = tile_symbol(vcur, vnxt, s32 + 4);
            const double s_p = __dadd_rn(xb[p], la_p);
            const double2 *x2 = reinterpret_cast<const double2 *>(xb);
            // "some s_i >= s_p" without compares: with t = pred(s_p) (the next double below
            // s_p), s_i >= s_p <=> s_i > t <=> t - s_i < 0, and the sign of a rounded
            // difference of two doubles is the sign of the exact one (a difference rounds to zero
            // only when it is zero) — so the sign bits of the K differences are OR-ed together,
            // one LOP3 per two predecessors.  (s_p = 0, -inf or NaN goes to the exact path
            // through viterbi_hoist_unsafe below.)
            const long long pb = __double_as_longlong(s_p);
            const double t = __longlong_as_double(pb - ((pb >> 63) | 1));
            unsigned sg[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int i = 0; i < KT; i += 2) {
                const double2 pq = x2[i / 2];
                sg[(i / 2) & 3] |= (unsigned)__double2hiint(__dsub_rn(t, __dadd_rn(pq.x, lac[i]))) |
                                   (unsigned)__double2hiint(__dsub_rn(t, __dadd_rn(pq.y, lac[i + 1])));
            }
            double M = __dadd_rn(s_p, e1);
            const bool beaten = (int)((sg[0] | sg[1]) | (sg[2] | sg[3])) < 0;
            const bool bad = (lane < K) & (beaten | viterbi_hoist_unsafe(s_p, e1, M));
            if (__builtin_expect(__any_sync(FULL, bad), 0)) {
                // exact column (out of line: its K sums and indices would not fit next to
                // the resident column of log a)
                const ScanResult r = viterbi_full_column<KT>(xb, LA + lane, K, K4, e1);
                M = r.best;
                if (r.arg != p) {
                    p = r.arg;
                    la_p = __ldg(LA + (size_t)p * KP + lane);
                    load_column_without(p);
                }
            }
            om = M;
            *bpt = (uint8_t)p;
            bpt += KP;
            e1 = e2;
            e2 = e3;
        };
        int64_t t0 = 0;
        for (; t0 + 32 < T; t0 += 32) {
#pragma unroll 2
            for (int s32 = 0; s32 < 32; ++s32) column(s32);
            vcur = vnxt;
            vnxt = st.load(t0 + 64, lane);
        }
#pragma unroll 1
        for (int s32 = 0; t0 + s32 + 1 < T; ++s32) column(s32);
        // first argmax of omega_{T-1}
        double best = (lane < K) ? om : -CUDART_INF;
        int bidx = (lane < K) ? lane : 0x7fffffff;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ob = __shfl_xor_sync(FULL, best, o);
            const int oi = __shfl_xor_sync(FULL, bidx, o);
            if (oi != 0x7fffffff && (bidx == 0x7fffffff || ob > best || (ob == best && oi < bidx))) {
                best = ob; bidx = oi;
            }
        }
        if (lane == 0) final_state[blk] = bidx;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Check-first sweep with an FP32 SCREEN (K <= 32, many chains).
//
// viterbi_check_kernel spends two FP64 operations per predecessor and column (form the sum,
// test it against s_p) only to learn, 98 % of the time, that the cached pointer still wins
// by several nats.  Here that question is first asked in single precision, on DIFFERENCES:
// every lane publishes d_i = float(omega_i - ref) next to omega_i (ref: omega_0 of the
// previous column, the same double in every lane, so the difference is taken in FP64 and
// only then rounded), and lane j compares s'_p = d_p + float(log a_pj) with the largest of
// the other s'_i = d_i + float(log a_ij) — FADD2 / FMNMX3 on the FP32 pipe, half as many
// shared-memory wavefronts, the column of log a in 28 registers instead of 56.
// Error bound.  The screen is only trusted in a column where every finite d_i and every
// finite log a_ij is below 64 in magnitude (else: exact path).  Then each s'_i is within
// 2^-24 (|d_i| + |log a_ij| + |s'_i|) < 2^-24 * 256 = 1.6e-5 of the real number
// (omega_i - ref) + log a_ij, and the reference's FP64 sum fl(omega_i + log a_ij) within
// 2^-53 |s_i| (< 1e-10) of that number + ref.  So  s'_p - max_i s'_i > 2^-14 = 6.1e-5  PROVES
// that s_p is the strict, unique maximum of the reference's sums — the cached pointer is
// np.argmax's answer — and the column's omega is then formed exactly as the reference does,
// in FP64, from that pointer: (omega_p + log a_pj) + log e_j, with the usual hoisting test.
// The same margin proves that the emission add can be hoisted (HOIST_SMALL): the real sums obey
// s_i <= s_p - (2^-14 - 2 * 1.53e-5 - 2e-10) < s_p - 3.0e-5 for every i != p, and when
// |M| = |fl(s_p + log e_j)| < 2^36 doubles around s_p + log e_j are at most 2^-16 apart (the
// binade below 2^37), so fl(s_i + log e_j) <= s_i + log e_j + 2^-17 < s_p + log e_j - 2^-17 <= M:
// no other predecessor can tie with p after the add, which is all viterbi_hoist_unsafe asks.
// So the hot path tests the exponent of M (two integer instructions) instead of forming
// fl(pred(s_p) + log e_j); the literal test runs out of line with the FP64 check.
// A margin inside the band (ties and near-ties) is decided by the FP64 check of
// viterbi_check_kernel, out of line; a lost pointer, an unsafe hoist, magnitudes beyond the
// bound or NaN by the exact FP64 column (viterbi_full_column).  Bit-identical paths by
// construction; FP64 work per column falls from ~56 to ~7 operations.
// ---------------------------------------------------------------------------------
// FP64 check of the cached pointer for the calling warp (cold path of viterbi_check32_kernel):
// true if some lane's s_p is NOT the strict unique maximum of its K sums.
template <int KT>
__device__ __noinline__ bool viterbi_pointer_beaten(const double *xb, const double *lac, int p, double s_p) {
    constexpr int KP = 32;
    const double2 *x2 = reinterpret_cast<const double2 *>(xb);
    const long long pb = __double_as_longlong(s_p);
    const double t = __longlong_as_double(pb - ((pb >> 63) | 1));         // pred(s_p): s_i >= s_p <=> t - s_i < 0
    unsigned sg = 0u;
#pragma unroll
    for (int i = 0; i < KT; i += 2) {
        const double2 pq = x2[i / 2];
        const double l0 = (i == p) ? -CUDART_INF : __ldg(lac + (size_t)i * KP);
        const double l1 = (i + 1 == p) ? -CUDART_INF : __ldg(lac + (size_t)(i + 1) * KP);
        sg |= (unsigned)__double2hiint(__dsub_rn(t, __dadd_rn(pq.x, l0))) | (unsigned)__double2hiint(__dsub_rn(t, __dadd_rn(pq.y, l1)));
    }
    return (int)sg < 0;
}

// sm_100: two FP32 adds in one instruction (FADD2) and a three-input maximum (FMNMX3)
__device__ __forceinline__ unsigned long long add_f32x2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float max3_f32(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// LAfT[j][i] = float(log a[i][j]) (KP = 32): lane j's column of log a, contiguous, in the screen's precision
__global__ void la_float_transposed_kernel(const double *__restrict__ LA, float *__restrict__ LAfT) {
    const int j = threadIdx.x >> 5, i = threadIdx.x & 31;
    LAfT[j * 32 + i] = (float)LA[i * 32 + j];
}

// maximum of N floats by three-input maxima arranged as a tree
template <int N>
__device__ __forceinline__ float max_tree3(const float (&v)[N]) {
    if constexpr (N == 1) return v[0];
    else if constexpr (N == 2) return fmaxf(v[0], v[1]);
    else {
        constexpr int M = (N + 2) / 3;
        float w[M];
#pragma unroll
        for (int g = 0; g < M; ++g) {
            const int a = 3 * g;
            w[g] = (a + 2 < N) ? max3_f32(v[a], v[a + 1], v[a + 2]) : (a + 1 < N) ? fmaxf(v[a], v[a + 1]) : v[a];
        }
        return max_tree3<M>(w);
    }
}

#ifndef ITR_VCHK32_PAIRS
#define ITR_VCHK32_PAIRS 1      /* column pairs per loop trip (2: same speed, twice the code) */
#endif
constexpr int VCHK32_PAIRS = ITR_VCHK32_PAIRS;
#ifndef ITR_VCHK32_MINB
#define ITR_VCHK32_MINB 3
#endif
template <int KT>
__global__ void __launch_bounds__(128, ITR_VCHK32_MINB)
viterbi_check32_kernel(ChainSet cs, const double *__restrict__ LA, const double *__restrict__ LEt,
                       const double *__restrict__ OM0, int K,
                       uint8_t *__restrict__ bp, int32_t *__restrict__ final_state,
                       const int64_t *__restrict__ chunk_off, uint8_t *__restrict__ comp,
                       const float *__restrict__ LAfT) {
    constexpr int KP = 32;
    constexpr float BAND = 6.103515625e-5f, BOUND = 64.f;      // 2^-14 and the magnitude bound, see above
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *xs = smem + (size_t)warp * 3 * KP;                 // omega, double buffered (2 KP doubles)
    float *fs = reinterpret_cast<float *>(xs + 2 * KP);        // the differences, double buffered (2 KP floats)
    const int n_chains = cs.n_blocks;
    const int K4 = (K + 3) & ~3;
    const double *etl = LEt + lane;
    // column `lane` of log a in single precision, two entries per 64-bit register (the scan
    // uses the packed FADD2 and the three-input FMNMX3 of sm_100: 28 instructions for 28
    // predecessors), the cached pointer's entry replaced by -inf
    unsigned long long laf2[KT / 2];
    // (read from LAfT = float(log a) transposed, la_float_transposed_kernel: seven 16-byte loads per
    // pointer change instead of 28 x (LDG.64 + F2F) — a pointer changes in 2.5 % of the columns)
    auto load_column_without = [&](int p) {
        const float4 *row = reinterpret_cast<const float4 *>(LAfT + (size_t)lane * KP);
        auto pack = [](float lo, float hi) {
            return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
        };
#pragma unroll
        for (int i = 0; i < KT; i += 4) {
            float4 v = __ldg(row + i / 4);
            v.x = (i == p) ? -CUDART_INF_F : v.x;
            v.y = (i + 1 == p) ? -CUDART_INF_F : v.y;
            v.z = (i + 2 == p) ? -CUDART_INF_F : v.z;
            v.w = (i + 3 == p) ? -CUDART_INF_F : v.w;
            laf2[i / 2] = pack(v.x, v.y);
            laf2[i / 2 + 1] = pack(v.z, v.w);
        }
    };
    // is every finite entry of this lane's column of log a inside the bound?  (-inf: that
    // predecessor can never be near the maximum)
    bool la_ok = true;
#pragma unroll
    for (int i = 0; i < KT; ++i) {
        const float l = (float)__ldg(LA + (size_t)i * KP + lane);
        la_ok &= (l == -CUDART_INF_F) | (fabsf(l) < BOUND);
    }

    for (int c = next_chain(cs, lane); c < n_chains; c = next_chain(cs, lane)) {
        const int blk = cs.order[c];
        const int64_t beg = cs.off[blk], T = cs.off[blk + 1] - beg;
        // symbols are read straight from HBM/L1, a few columns ahead (the buffer is padded by 64
        // entries, so the read-ahead past a block's end stays inside it and yields valid symbols):
        // one load per column instead of the five ALU instructions of a register tile + shuffle
        const uint16_t *spt = cs.sym + beg;
        uint8_t *bpt = bp + (size_t)beg * KP + lane + KP;        // row of column t = 1

        int p = lane;                                            // cached pointer: "stay"
        double la_p = __ldg(LA + (size_t)lane * KP + lane);
        float laf_p = la_ok ? (float)la_p : -CUDART_INF_F;      // (a column of log a beyond the bound never passes the screen)
        load_column_without(p);
        double om = __ldg(OM0 + (size_t)blk * KP + lane);
        double ref = __shfl_sync(FULL, om, 0);
        // (one column of read-ahead for the emission row, one more for its symbol: a column takes
        // hundreds of cycles, an L1 hit forty — and with the loop unrolled by two a pipeline of depth
        // one needs no register moves)
        double e1 = __ldg(etl + (size_t)__ldg(spt + 1) * KP);
        unsigned v2 = __ldg(spt + 2);
        spt += 3;
        // omega and its differences are double buffered (one __syncwarp per column); the loop
        // below names the buffers explicitly, so every shared-memory address of a column is a
        // loop constant — including the pointer's own entries, cached per buffer
        const double *xp0 = xs + p, *xp1 = xs + KP + p;
        const float *fq0 = fs + p, *fq1 = fs + KP + p;
        auto column = [&](double *xb, float *fb, const double *xpp, const float *fpp) {
            const float d = (float)__dsub_rn(om, ref);
            xb[lane] = om;
            fb[lane] = d;
            __syncwarp();
            const double e2 = __ldg(etl + v2 * KP);
            v2 = __ldg(spt++);
            // exact part: the reference's two adds for the cached pointer
            const double s_p = __dadd_rn(*xpp, la_p);
            double M = __dadd_rn(s_p, e1);
            // screen: is any other predecessor within the band of s_p?
            const float fp = *fpp + laf_p;
            const ulonglong2 *f4 = reinterpret_cast<const ulonglong2 *>(fb);
            float sv[KT];
#pragma unroll
            for (int i = 0; i < KT; i += 4) {
                const ulonglong2 q = f4[i / 4];
                const unsigned long long r01 = add_f32x2(q.x, laf2[i / 2]), r23 = add_f32x2(q.y, laf2[i / 2 + 1]);
                sv[i] = __uint_as_float((unsigned)r01);
                sv[i + 1] = __uint_as_float((unsigned)(r01 >> 32));
                sv[i + 2] = __uint_as_float((unsigned)r23);
                sv[i + 3] = __uint_as_float((unsigned)(r23 >> 32));
            }
            const float other = max_tree3<KT>(sv);               // (three-input maxima as a tree: depth 4 instead of 7)
            const double ref_next = xb[0];
            // the bound on |d| is a property of the column: every live lane votes on its own d
            const bool d_ok = (lane >= K) | (d == -CUDART_INF_F) | (fabsf(d) < BOUND);
            // A proven margin also proves the hoist (HOIST_SMALL, see the header): two instructions
            // instead of the sixteen of viterbi_hoist_unsafe.  One vote serves both questions: a lane
            // whose d is out of bounds sends the whole warp to the FP64 path, as an unproven lane does.
            const bool proven = (fp - other > BAND) & viterbi_small_for_hoist(M);
            if (__builtin_expect(__any_sync(FULL, ((lane < K) & !proven) | !d_ok), 0)) {
                // band, bound or hoist: the FP64 check decides whether the pointers stand
                // (a pointer that lost by more than the band goes straight to the exact column — that is
                // always right, so the shortcut needs no bound; nearly every column that gets here is one)
                bool redo = __any_sync(FULL, (lane < K) & ((fp - other < -BAND) | viterbi_hoist_unsafe(s_p, e1, M)));
                if (!redo) redo = __any_sync(FULL, (lane < K) & viterbi_pointer_beaten<KT>(xb, LA + lane, p, s_p));
                if (redo) {
                    const ScanResult r = viterbi_full_column<KT>(xb, LA + lane, K, K4, e1);
                    M = r.best;
                    if (r.arg != p) {
                        p = r.arg;
                        la_p = __ldg(LA + (size_t)p * KP + lane);
                        laf_p = la_ok ? (float)la_p : -CUDART_INF_F;
                        load_column_without(p);
                        xp0 = xs + p, xp1 = xs + KP + p;
                        fq0 = fs + p, fq1 = fs + KP + p;
                    }
                }
            }
            om = M;
            ref = ref_next;
            *bpt = (uint8_t)p;
            bpt += KP;
            e1 = e2;
        };
        // The traceback's chunk composites ("state at the last column of the previous chunk, given
        // state j at this chunk's last column", viterbi_compose_kernel) fall out of the sweep:
        // F_t[j] = F_{t-1}[p_t[j]] is one shuffle per column, and the 8 GB re-read of the
        // backpointers by the compose kernel disappears.
        // Column 1 goes first, alone, in buffer 0: after it every chunk but a block's last holds an
        // even number of columns, so the pairs (buffer 1, buffer 0) never straddle a chunk.
        const int64_t cbase = chunk_off[blk];
        int64_t t = 1;
        int F = lane;
        if (t < T) {
            column(xs, fs, xp0, fq0);
            F = __shfl_sync(FULL, F, p);
            ++t;
        }
        for (int64_t c0 = 0; c0 * VCHUNK < T; ++c0) {
            const int64_t tend = min((c0 + 1) * (int64_t)VCHUNK, T);
            if (c0 > 0) F = lane;
            const int n = (int)(tend - t);
#pragma unroll VCHK32_PAIRS
            for (int i = 0; i + 1 < n; i += 2) {
                column(xs + KP, fs + KP, xp1, fq1);
                F = __shfl_sync(FULL, F, p);
                column(xs, fs, xp0, fq0);
                F = __shfl_sync(FULL, F, p);
            }
            if (n & 1) {                                   // (a block's last chunk only)
                column(xs + KP, fs + KP, xp1, fq1);
                F = __shfl_sync(FULL, F, p);
            }
            t = tend;
            if (c0 > 0) comp[(size_t)(cbase + c0) * KP + lane] = (uint8_t)F;
        }
        // first argmax of omega_{T-1}
        double best = (lane < K) ? om : -CUDART_INF;
        int bidx = (lane < K) ? lane : 0x7fffffff;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ob = __shfl_xor_sync(FULL, best, o);
            const int oi = __shfl_xor_sync(FULL, bidx, o);
            if (oi != 0x7fffffff && (bidx == 0x7fffffff || ob > best || (ob == best && oi < bidx))) {
                best = ob; bidx = oi;
            }
        }
        if (lane == 0) final_state[blk] = bidx;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Viterbi forward sweep, four warps per chain (K <= 32).  With few chains (config 2:
// 100 blocks on 148 SMs) the sweep is bound by the per-column latency of one warp's
// ~200-instruction argmax.  Here a CTA of four warps (one per SM sub-partition) walks
// one chain: warp w scans the CPW = KT8/4 predecessors i in [w CPW, (w+1) CPW) for all
// 32 successor lanes, the four (value, index) partials meet in shared memory behind one
// bar.sync, and every warp redundantly finishes the column (merge of the partials in
// i-order with "right wins only if strictly greater" = first maximum, hoisted emission
// add + exactness check exactly as in viterbi_forward_kernel).  Same results bit for
// bit; ~2.5x shorter column latency.
// ---------------------------------------------------------------------------------
// Cold path of viterbi_forward4_kernel (about once per binade crossing of omega): the
// exactness check of the previous column fired.  Every warp redoes that column with
// the literal two-add scan from omega_{t-2} (`prev`), overwrites its copy of
// omega_{t-1} (`cur`) and the backpointer row.  Kept out of line so that the hot loop
// stays straight-line code.
__device__ __noinline__ void viterbi4_repair(const double *prev, double *cur, const double *lac, int K4,
                                             double le, uint8_t *row, int *slow_flag) {
    __syncthreads();                       // everyone has seen the flag
    if (threadIdx.x == 0) *slow_flag = 0;
    const ScanResult r = viterbi_exact_scan(prev, lac, 32, K4, le);
    cur[threadIdx.x & 31] = r.best;
    if (threadIdx.x < 32) *row = (uint8_t)r.arg;
    __syncwarp();
}

__device__ __forceinline__ double2 lds_f64x2(const double *p) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}

template <int CPW>
__global__ void __launch_bounds__(128)
viterbi_forward4_kernel(ChainSet cs, const double *__restrict__ LA, const double *__restrict__ LEt,
                        const double *__restrict__ OM0, int K,
                        uint8_t *__restrict__ bp, int32_t *__restrict__ final_state) {
    constexpr int KP = 32;
    __shared__ __align__(16) double xo[4][2][KP];       // each warp's private copies of omega_{t-1}, omega_t
    __shared__ __align__(16) double pv[2][4][KP];       // partial maxima
    __shared__ int pidx[2][4][KP];                      // partial arg-maxima
    __shared__ int chain_s, slow_flag;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_chains = cs.n_blocks;
    const int K4 = (K + 3) & ~3;
    const double *etl = LEt + lane;
    const int i0 = warp * CPW;
    if (threadIdx.x == 0) slow_flag = 0;

    // rows i0 .. i0+CPW-1 of log a, column `lane`
    double la[CPW];
#pragma unroll
    for (int c = 0; c < CPW; ++c) la[c] = __ldg(LA + (size_t)(i0 + c) * KP + lane);

    for (;;) {
        if (threadIdx.x == 0) chain_s = (int)atomicAdd(cs.queue, 1u);
        __syncthreads();
        const int c = chain_s;
        __syncthreads();
        if (c >= n_chains) break;
        const int blk = cs.order[c];
        const int64_t beg = cs.off[blk], T = cs.off[blk + 1] - beg;
        const SymTile st{cs.sym + beg, T};
        uint8_t *bpl = bp + (size_t)beg * KP + lane;

        unsigned vcur = st.load(0, lane);
        unsigned vnxt = st.load(32, lane);
        double e1, e2, e0 = 0.0;                 // e0: emission row of the column just produced
        {
            const unsigned v1 = __shfl_sync(FULL, vcur, 1), v2 = __shfl_sync(FULL, vcur, 2);
            e1 = __ldg(etl + v1 * KP);
            e2 = __ldg(etl + v2 * KP);
        }
        int cur = 0;                             // xo[warp][cur] holds the newest omega
        xo[warp][0][lane] = __ldg(OM0 + (size_t)blk * KP + lane);
        __syncwarp();
        unsigned vpre = tile_symbol(vcur, vnxt, 3);      // symbol of the column two ahead
        uint8_t *bpt = bpl + KP;                          // row of column t = 1
        int buf = 0;
        // exactness check of the previous column, evaluated lazily: f(pred(s*)) == f(s*)?
        double chk_s = -1.0, chk_m = 0.0;        // s* and M of the previous column (e0 is its emission)

        auto column = [&](int s32, int64_t t0) {
            double mv[4];
            int mi[4];
            for (;;) {
                // ---- phase A: this warp's CPW predecessors of the column after xo[warp][cur]
                const double *xw = &xo[warp][cur][i0];
                double2 px[CPW / 2];
#pragma unroll
                for (int q = 0; q < CPW / 2; ++q) px[q] = lds_f64x2(xw + 2 * q);
                // the previous column's check, off the dependent chain (its inputs are old)
                {
                    const long long bits = __double_as_longlong(chk_s);
                    const double pred = __longlong_as_double(bits - ((bits >> 63) | 1));
                    const bool odd = (chk_s == 0.0) | !(fabs(chk_s) < CUDART_INF);
                    if ((lane < K) & (odd | (__dadd_rn(pred, e0) == chk_m))) slow_flag = 1;
                }
                double sv[CPW];
                int ix[CPW];
#pragma unroll
                for (int q = 0; q < CPW; q += 2) {
                    sv[q] = __dadd_rn(px[q / 2].x, la[q]);
                    sv[q + 1] = __dadd_rn(px[q / 2].y, la[q + 1]);
                    ix[q] = i0 + q;
                    ix[q + 1] = i0 + q + 1;
                }
                tournament<CPW>(sv, ix);
                pv[buf][warp][lane] = sv[0];
                pidx[buf][warp][lane] = ix[0];
                __syncthreads();
                // ---- phase B: the four partials (ascending i ranges) and the check flag
                const int sf = slow_flag;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    mv[w] = pv[buf][w][lane];
                    mi[w] = pidx[buf][w][lane];
                }
                if (__builtin_expect(sf == 0, 1)) break;
                viterbi4_repair(&xo[warp][cur ^ 1][0], &xo[warp][cur][0], LA + lane, K4, e0, bpt - KP, &slow_flag);
                chk_s = -1.0;                     // resolved (a finite negative value never fires)
                chk_m = 0.0;
                __syncthreads();                  // partial buffers may be rewritten
            }
            // emission row two columns ahead; its symbol was shuffled out one column ago
            const double e3 = __ldg(etl + vpre * KP);
            vpre = tile_symbol(vcur, vnxt, s32 + 4);
            buf ^= 1;
            tournament<4>(mv, mi);
            const double sstar = mv[0];
            const double M = __dadd_rn(sstar, e1);
            cur ^= 1;
            xo[warp][cur][lane] = M;
            __syncwarp();
            if (warp == (s32 & 3)) *bpt = (uint8_t)mi[0];    // the four warps take turns
            bpt += KP;
            chk_s = sstar;
            chk_m = M;
            e0 = e1;
            e1 = e2;
            e2 = e3;
        };
        int64_t t0 = 0;
        for (; t0 + 32 < T; t0 += 32) {
#pragma unroll 4
            for (int s32 = 0; s32 < 32; ++s32) column(s32, t0);
            vcur = vnxt;
            vnxt = st.load(t0 + 64, lane);
        }
#pragma unroll 1
        for (int s32 = 0; t0 + s32 + 1 < T; ++s32) column(s32, t0);
        {   // check of the last column
            const long long bits = __double_as_longlong(chk_s);
            const double pred = __longlong_as_double(bits - ((bits >> 63) | 1));
            const bool odd = (chk_s == 0.0) | !(fabs(chk_s) < CUDART_INF);
            const bool slow = (lane < K) & (odd | (__dadd_rn(pred, e0) == chk_m));
            if (__syncthreads_or(slow) && T > 1)
                viterbi4_repair(&xo[warp][cur ^ 1][0], &xo[warp][cur][0], LA + lane, K4, e0, bpt - KP, &slow_flag);
            __syncthreads();
        }
        if (warp == 0) {
            // first argmax of omega_{T-1}
            const double om = xo[0][cur][lane];
            double best = (lane < K) ? om : -CUDART_INF;
            int bidx = (lane < K) ? lane : 0x7fffffff;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double ob = __shfl_xor_sync(FULL, best, o);
                const int oi = __shfl_xor_sync(FULL, bidx, o);
                if (oi != 0x7fffffff && (bidx == 0x7fffffff || ob > best || (ob == best && oi < bidx))) {
                    best = ob; bidx = oi;
                }
            }
            if (lane == 0) final_state[blk] = bidx;
        }
        __syncthreads();
    }
}

// f(pred(s*)) == f(s*)?  (see the header comment of viterbi_forward_kernel)
__device__ __forceinline__ bool viterbi_hoist_unsafe(double sstar, double le, double M) {
    const long long bits = __double_as_longlong(sstar);
    const double pred = __longlong_as_double(bits - ((bits >> 63) | 1));   // next double towards -inf
    const bool odd = (sstar == 0.0) | !(fabs(sstar) < CUDART_INF);
    return odd | (__dadd_rn(pred, le) == M);
}

// ---------------------------------------------------------------------------------
// Viterbi forward sweep by speculation and exact verification, DECOUPLED (K <= 32, few chains).
//
// Measured on the (3,3) model: the vector of backpointers changes in only 1.7 % of the
// columns (mean stable stretch 55 columns; 90 % of the pointers are "stay").  So one warp
// (the runner) advances speculatively with cached pointers — omega_j = (omega_p + log a_pj)
// + log e_j: exactly-rounded adds and one shared-memory gather per column instead of a K-way
// arg-max on the dependent chain — while the other warps verify every column it produced with
// the full exact scan (tournament + hoisting check + literal fallback, as in
// viterbi_forward_kernel) from the speculative omega of the column before.  Columns before
// the first mismatch are committed (verified: same adds, same first maximiser as the
// reference); the mismatching column takes the verifier's result, the pointers are updated
// and the runner resumes right after it: bit-identical to the sequential sweep by induction
// over committed columns.  The runner never waits for verification: it streams columns into
// a ring of STR_R slots and only stops when a verifier raises a mismatch.  (A windowed
// predecessor — speculate 15 columns, verify, repeat — ran at 172 cycles per column;
// DESIGN.md section 7.)  Roles in a CTA of 16 warps:
//   warp 0       runner   omega_t[j] = (max over the candidate pair (lo, hi) of omega_{t-1}[p]
//                         + log a[p][j]) + log e_t[j]; publishes run_t
//   warp 15      feeder   stages the log-emission row of every upcoming column in the ring
//                         (the runner's and the verifiers' lookups become plain indexed loads)
//   other warps  verifiers (11: warps 4, 8, 12 idle so that the runner owns its scheduler), column u
//                         belongs to verifier u mod 11: full exact scan from the
//                         speculative omega_{u-1}; equal pointers -> commit the backpointer
//                         row; else store the verified (arg, omega) and lower fail_t to u
// Epochs: when fail_t is set (a mismatch, or the virtual one at column T that ends the block)
// every verifier still finishes its columns below fail_t, then all warps meet at a barrier;
// by then every column < fail_t is verified (fail_t only decreases), column fail_t takes the
// verifier's result, the runner's candidate pair is updated and everyone resumes at
// fail_t + 1.  A column is always re-verified by the same warp, so a backpointer row written
// from a speculation that was later rolled back is overwritten in program order.
// Ring slots are reused 64 columns later; the runner and the feeder poll the verifiers'
// progress (every 4 columns, consumed 4 columns later) and wait if they would overrun it.
// ---------------------------------------------------------------------------------
#ifndef ITR_STR_ISOLATE
#define ITR_STR_ISOLATE 1
#endif
// ISOLATE: warps 4, 8, 12 stay idle so that the runner has its scheduler (and its FP64
// pipe) to itself; 11 verifiers on the other three schedulers still verify ~2x faster
// than the runner produces.
constexpr int STR_R = 64;
#ifndef ITR_STR_G
#define ITR_STR_G 8
#endif
constexpr int STR_G = ITR_STR_G;                 // columns the runner produces between two looks at the other warps
constexpr int STR_NONE = 0x7fffffff;
#ifndef ITR_STR_SLEEP
#define ITR_STR_SLEEP 200
#endif

// NWARPS = 16: one chain per SM (up to ~1.5 chains per SM in the queue): runner, feeder, 11
//              verifiers, three idle warps; the log-emission table (160 KB) in shared memory.
// NWARPS = 8:  two chains per SM for hundreds of chains (a GPU's share of a chromosome split
//              over 4-8 GPUs): runner, feeder, 6 verifiers; the feeder reads the emission rows
//              through L1 instead (it is off the chain), so a CTA needs only its rings (46 KB)
//              and two fit an SM.  The runner shares its scheduler with a verifier.
template <int NWARPS>
struct StreamCfg {
    static constexpr bool ISOLATE = NWARPS == 16 && ITR_STR_ISOLATE;
    static constexpr bool TABLE = NWARPS == 16;                       // log-emission table in shared memory
    // verifiers: FP32 screen + FP64 check + exact column (as viterbi_check32_kernel) instead of a
    // full arg-max per column (ITR_STR_SCREEN=0 compiles the full arg-max back in).  Measured:
    // N = 8 share of config 4 with 8-warp CTAs: step 26.2 -> 22.9 ms; 100 chains with 16-warp
    // CTAs: sweep 11.5 -> 11.1 ms.
#ifndef ITR_STR_SCREEN
#define ITR_STR_SCREEN 1
#endif
    static constexpr bool SCREEN = ITR_STR_SCREEN != 0;
    static constexpr int NV = ISOLATE ? 11 : NWARPS - 2;
    static constexpr int SLOTS = NWARPS <= 8 ? 8 : 16;                // verifier slots (power of two >= NV)
    static constexpr size_t SMEM = (size_t)((TABLE ? NSYM * 32 : 0) + 2 * STR_R * 32) * sizeof(double) + (size_t)STR_R * 32;
};

#ifndef ITR_STR8_MINB
#define ITR_STR8_MINB 2
#endif
template <int KT, int NWARPS>
__global__ void __launch_bounds__(32 * NWARPS, NWARPS == 8 ? ITR_STR8_MINB : 1)
viterbi_stream_kernel(ChainSet cs, const double *__restrict__ LA, const double *__restrict__ LEt,
                      const double *__restrict__ OM0, int K,
                      uint8_t *__restrict__ bp, int32_t *__restrict__ final_state) {
    using Cfg = StreamCfg<NWARPS>;
    constexpr int STR_NW = NWARPS, STR_NV = Cfg::NV, STR_SL = Cfg::SLOTS;
    constexpr int KP = 32, R = STR_R, NV = STR_NV;
    __shared__ double las[KP][KP];                      // log a, for the runner's pointer lookups
    __shared__ __align__(16) double vrom[STR_SL][KP];   // verified omega of a mismatching column, per verifier
    __shared__ int varg[STR_SL][KP];                    // its verified first arg-maxima
    __shared__ volatile int vfail[STR_SL];              // the column it belongs to (STR_NONE: none)
    __shared__ volatile int ver_next[STR_SL];           // next column each verifier will look at
    __shared__ volatile int run_t, feed_t, fail_t;      // last column produced / staged (exclusive) / first bad column
    __shared__ int chain_s;
    __shared__ __align__(16) float vdiff[STR_SL][KP];   // a verifier's omega differences in single precision
    extern __shared__ __align__(16) double dyn[];
    constexpr int TAB = Cfg::TABLE ? NSYM * KP : 0;
    double *les = dyn;                                                   // [NSYM][KP] log-emission table (NWARPS = 16)
    double (*ring_om)[KP] = reinterpret_cast<double (*)[KP]>(dyn + TAB);          // omega after column t, slot t % R
    double (*ring_ew)[KP] = reinterpret_cast<double (*)[KP]>(dyn + TAB + R * KP); // log e row of column t
    uint8_t (*ring_ptr)[KP] = reinterpret_cast<uint8_t (*)[KP]>(dyn + TAB + 2 * R * KP);  // the runner's choice
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_chains = cs.n_blocks;
    const int K4 = (K + 3) & ~3;
    if (Cfg::TABLE)
        for (int e = threadIdx.x; e < NSYM * KP; e += blockDim.x) les[e] = __ldg(LEt + e);
    // verifiers: column `lane` of log a in single precision, packed two per register, with the
    // entry of the pointer they expect (pc: the runner's choice the last time they looked)
    // replaced by -inf — the FP32 screen of viterbi_check32_kernel
    unsigned long long laf2[KT / 2];
    int pc = -1;
    bool la_ok = true;
    Cols<KT, 1, true> lacol;                            // NWARPS = 16: column `lane` of log a in FP64, full arg-max per column
    auto load_column_without = [&](int p) {
#pragma unroll
        for (int i = 0; i < KT; i += 2) {
            const float l0 = (i == p) ? -CUDART_INF_F : (float)las[i][lane];
            const float l1 = (i + 1 == p) ? -CUDART_INF_F : (float)las[i + 1][lane];
            laf2[i / 2] = (unsigned long long)__float_as_uint(l0) | ((unsigned long long)__float_as_uint(l1) << 32);
        }
    };
    // verifier index of this warp (-1: runner, feeder or idle)
    const int vi = (warp == 0 || warp == STR_NW - 1) ? -1
                   : Cfg::ISOLATE ? ((warp & 3) ? warp - 1 - (warp >> 2) : -1) : warp - 1;
    for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) las[e / KP][e % KP] = __ldg(LA + e);
    __syncthreads();
    if (vi >= 0 && !Cfg::SCREEN) lacol.load(LA, KP, lane);
    if (vi >= 0 && Cfg::SCREEN) {
#pragma unroll
        for (int i = 0; i < KT; ++i) {
            const float l = (float)las[i][lane];
            la_ok &= (l == -CUDART_INF_F) | (fabsf(l) < 64.f);
        }
        pc = lane;
        load_column_without(pc);
    }

    for (;;) {
        if (threadIdx.x == 0) chain_s = (int)atomicAdd(cs.queue, 1u);
        __syncthreads();
        const int c = chain_s;
        __syncthreads();
        if (c >= n_chains) break;
        const int blk = cs.order[c];
        const int64_t beg = cs.off[blk];
        const int T = (int)(cs.off[blk + 1] - beg);     // (block lengths fit in 31 bits; checked by the host)
        const uint16_t *symp = cs.sym + beg;
        uint8_t *bpl = bp + (size_t)beg * KP + lane;

        // runner state: the two most recent predecessors of state `lane`, sorted by index
        int lo = lane, hi = lane;
        double la_lo = las[lane][lane], la_hi = la_lo;
        int t_run = 1;                                  // runner: next column to produce
        int t_feed = 1;                                 // feeder: next column to stage
        int u_ver = (vi >= 0) ? 1 + (vi - 1 + NV) % NV : STR_NONE;      // column u belongs to verifier u % NV
        if (warp == 0) {
            ring_om[0][lane] = __ldg(OM0 + (size_t)blk * KP + lane);
            if (lane == 0) { run_t = 0; feed_t = 1; fail_t = STR_NONE; }
        }
        if (lane == 0 && vi >= 0) { vfail[vi] = STR_NONE; ver_next[vi] = u_ver; }
        __syncthreads();

        for (;;) {                                      // epochs between roll-backs
            if (warp == 0) {
                // ---------------- runner ----------------
                // Probes of the other warps' progress are issued at the top of a group of
                // STR_G columns and consumed at the top of the next one (they only ever lag,
                // which is conservative), so that their latency stays off the column chain.
                constexpr int G = STR_G;
                int p_fail = STR_NONE, p_feed = feed_t, p_min;
                {
                    const int mv = ver_next[lane % NV];
                    p_min = __reduce_min_sync(FULL, mv);
                }
                double xl = ring_om[(t_run - 1) & (R - 1)][lo], xh = ring_om[(t_run - 1) & (R - 1)][hi];
                double ew = 0.0;
                bool have_ew = false, stop = false;
                while (!stop) {
                    if (t_run >= T) { if (lane == 0) atomicMin((int *)&fail_t, T); break; }
                    if (p_fail != STR_NONE) break;
                    // safe to produce columns t_run .. t_run+G-1 (and to read the emission row of t_run+G)?
                    while (p_feed < min(t_run + G + 1, T) || p_min + (R - 1) <= t_run + G - 1) {
                        __nanosleep(64);
                        if (fail_t != STR_NONE) { stop = true; break; }
                        p_feed = feed_t;
                        const int mv = ver_next[lane % NV];
                        p_min = __reduce_min_sync(FULL, mv);
                    }
                    if (stop) break;
                    if (!have_ew) { ew = ring_ew[t_run & (R - 1)][lane]; have_ew = true; }
                    const int q_fail = fail_t, q_feed = feed_t, q_mv = ver_next[lane % NV];
                    int q_min = 0;
                    // one column: two candidate sums, the better one plus the emission, exchange through the ring
#define ITR_STR_COLUMN()                                                                                  \
    do {                                                                                                  \
        const int slot = t_run & (R - 1);                                                                 \
        const double s_l = __dadd_rn(xl, la_lo), s_h = __dadd_rn(xh, la_hi);                              \
        const bool take = s_h > s_l;                         /* (the verifier has the last word) */       \
        const double M = __dadd_rn(take ? s_h : s_l, ew);                                                 \
        ring_om[slot][lane] = M;                                                                          \
        __syncwarp();                                                                                     \
        xl = ring_om[slot][lo];                                                                           \
        xh = ring_om[slot][hi];                                                                           \
        ew = ring_ew[(t_run + 1) & (R - 1)][lane];           /* (staged: p_feed > t_run + G) */           \
        ring_ptr[slot][lane] = (uint8_t)(take ? hi : lo);                                                 \
        ++t_run;                                                                                          \
    } while (0)
                    if (t_run + G <= T) {
#pragma unroll
                        for (int i = 0; i < G; ++i) {
                            ITR_STR_COLUMN();
                            if (i == 1) q_min = __reduce_min_sync(FULL, q_mv);   // (its operand has landed by now)
                        }
                    } else {
                        q_min = __reduce_min_sync(FULL, q_mv);
                        while (t_run < T) ITR_STR_COLUMN();
                    }
#undef ITR_STR_COLUMN
#ifndef ITR_STR_NOFENCE
                    __threadfence_block();
#endif
                    if (lane == 0) run_t = t_run - 1;
                    p_fail = q_fail;
                    p_feed = q_feed;
                    p_min = q_min;
                }
            } else if (warp == STR_NW - 1) {
                // ---------------- feeder ----------------
                for (;;) {
                    if (fail_t != STR_NONE) break;
                    if (t_feed >= T) { __nanosleep(256); continue; }
                    // slot reuse: column t_feed + 31 - R must be verified
                    const int mv = ver_next[lane % NV];
                    const int vmin = __reduce_min_sync(FULL, mv);
                    const int n = min(min(32, T - t_feed), vmin + R - 1 - t_feed);
                    if (n <= 0) { __nanosleep(128); continue; }
                    const int mysym = (int)__ldg(symp + t_feed + lane);     // (64 columns of slack behind the last block)
                    for (int i = 0; i < n; ++i) {
                        const int sy = __shfl_sync(FULL, mysym, i);
                        ring_ew[(t_feed + i) & (R - 1)][lane] = Cfg::TABLE ? les[sy * KP + lane] : __ldg(LEt + sy * KP + lane);
                    }
                    t_feed += n;
                    __threadfence_block();
                    if (lane == 0) feed_t = t_feed;
                }
            } else if (vi >= 0) {
                // ---------------- verifiers ----------------
                for (;;) {
                    const int u = u_ver;
                    bool go = true;
                    while (run_t < u) {
                        if (fail_t <= u) { go = false; break; }
                        __nanosleep(ITR_STR_SLEEP);
                    }
                    if (!go || fail_t <= u) break;
                    __threadfence_block();
                    __syncwarp();
#ifdef ITR_STR_ABL_NOVERIFY   /* ablation: accept every column unverified (results are wrong) */
                    u_ver = u + NV;
                    if (lane == 0) ver_next[vi] = u + NV;
                    continue;
#endif
                    const int myp = ring_ptr[u & (R - 1)][lane];
                    const double le = ring_ew[u & (R - 1)][lane];
                    const double *xin = &ring_om[(u - 1) & (R - 1)][0];
#ifdef ITR_STR_DEBUG
                    if (lane == 0 && (u < 1 || u >= T || run_t < u))
                        printf("STR-DEBUG blk %d verifier %d: column %d outside [1, %d) or ahead of run_t %d\n", blk, vi, u, T, run_t);
#endif
                    double M;
                    int arg;
                    if constexpr (Cfg::SCREEN) {
                        // Screen first (see viterbi_check32_kernel): is the runner's pointer the strict,
                        // unique maximum by more than the FP32 band?  Then the column stands as produced.
                        // Else the FP64 check, and only if that fails too the full arg-max.
                        if (myp != pc) {                     // (rare: the pointer of this lane moved since this warp last looked)
                            pc = myp;
                            load_column_without(pc);
                        }
                        const double xme = xin[lane], ref = xin[0];
                        const float d = (float)__dsub_rn(xme, ref);
                        float *fb = &vdiff[vi][0];
                        __syncwarp();                        // the previous column's readers are done
                        fb[lane] = d;
                        __syncwarp();
                        const double la_p = las[myp][lane];
                        const double s_p = __dadd_rn(xin[myp], la_p);
                        M = __dadd_rn(s_p, le);
                        arg = myp;
                        const float fp = fb[myp] + (float)la_p;
                        const ulonglong2 *f4 = reinterpret_cast<const ulonglong2 *>(fb);
                        float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F;
    #pragma unroll
                        for (int q = 0; q < KT; q += 4) {
                            const ulonglong2 qq = f4[q / 4];
                            const unsigned long long r01 = add_f32x2(qq.x, laf2[q / 2]), r23 = add_f32x2(qq.y, laf2[q / 2 + 1]);
                            m0 = max3_f32(m0, __uint_as_float((unsigned)r01), __uint_as_float((unsigned)(r01 >> 32)));
                            m1 = max3_f32(m1, __uint_as_float((unsigned)r23), __uint_as_float((unsigned)(r23 >> 32)));
                        }
                        const bool d_ok = (lane >= K) | (d == -CUDART_INF_F) | (fabsf(d) < 64.f);
                        const bool proven = la_ok & (fp - fmaxf(m0, m1) > 6.103515625e-5f) & viterbi_small_for_hoist(M);
                        if (__any_sync(FULL, ((lane < K) & !proven) | !d_ok)) {
                            bool redo = __any_sync(FULL, (lane < K) & viterbi_hoist_unsafe(s_p, le, M));
                            if (!redo) redo = __any_sync(FULL, (lane < K) & viterbi_pointer_beaten<KT>(xin, LA + lane, myp, s_p));
                            if (redo) {
                                const ScanResult r = viterbi_full_column<KT>(xin, LA + lane, K, K4, le);
                                M = r.best;
                                arg = r.arg;
                            }
                        }
                    } else {
                        const double2 *x2 = reinterpret_cast<const double2 *>(xin);
                        double sv[KT];
                        int ix[KT];
    #pragma unroll
                        for (int q = 0; q < KT; q += 2) {
                            const double2 pq = x2[q / 2];
                            sv[q] = __dadd_rn(pq.x, lacol.get(0, q));
                            sv[q + 1] = __dadd_rn(pq.y, lacol.get(0, q + 1));
                            ix[q] = q;
                            ix[q + 1] = q + 1;
                        }
                        tournament<KT>(sv, ix);
                        M = __dadd_rn(sv[0], le);
                        arg = ix[0];
                        if (__any_sync(FULL, (lane < K) & viterbi_hoist_unsafe(sv[0], le, M))) {
                            const ScanResult r = viterbi_exact_scan(xin, LA + lane, KP, K4, le);
                            M = r.best;
                            arg = r.arg;
                        }
                    }
                    if (__any_sync(FULL, (lane < K) & (arg != myp))) {
                        vrom[vi][lane] = M;
                        varg[vi][lane] = arg;
                        __threadfence_block();
                        if (lane == 0) { vfail[vi] = u; atomicMin((int *)&fail_t, u); }
                        break;
                    }
                    bpl[(size_t)u * KP] = (uint8_t)myp;
                    u_ver = u + NV;
                    if (lane == 0) ver_next[vi] = u + NV;
                }
            }
            __syncthreads();                            // everyone stopped; every column < fail_t is verified
            const int f = fail_t;
            // Warp 0 picks up the verified result of column f HERE, between the two barriers,
            // while nobody writes: after the second barrier the verifiers reset their slots
            // concurrently with the repair (reading vfail there was a race that only showed
            // under a profiler's timing).
            int np2 = 0, chosen = 0;
            double fix_om = 0.0;
            if (warp == 0 && f < T) {
                const int vf = vfail[lane & (STR_SL - 1)];
                const unsigned who = __ballot_sync(FULL, lane < NV && vf == f);
                const int v = who ? __ffs(who) - 1 : 0;
#ifdef ITR_STR_DEBUG
                if (who == 0 && lane == 0) printf("STR-DEBUG blk %d T %d: fail_t %d but no verifier owns it\n", blk, T, f);
#endif
                np2 = varg[v][lane];
                fix_om = vrom[v][lane];
                chosen = ring_ptr[f & (R - 1)][lane];
            }
            __syncthreads();
            if (f >= T) break;                          // the virtual mismatch at column T: block done
            if (warp == 0) {
                // new candidate pair: the verified predecessor and the most recent other one
                const int other = (chosen != np2) ? chosen : (lo != np2) ? lo : hi;
                lo = min(np2, other);
                hi = max(np2, other);
                la_lo = las[lo][lane];
                la_hi = las[hi][lane];
                bpl[(size_t)f * KP] = (uint8_t)np2;
                ring_om[f & (R - 1)][lane] = fix_om;
                t_run = f + 1;
                if (lane == 0) { run_t = f; fail_t = STR_NONE; }
            } else if (vi >= 0) {
                u_ver = f + 1 + (vi - (f + 1) % NV + NV) % NV;
                if (lane == 0) { vfail[vi] = STR_NONE; ver_next[vi] = u_ver; }
            }
            __syncthreads();
        }
        if (warp == 0) {
            // first argmax of omega_{T-1}
            const double om = ring_om[(T - 1) & (R - 1)][lane];
            double best = (lane < K) ? om : -CUDART_INF;
            int bidx = (lane < K) ? lane : 0x7fffffff;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double ob = __shfl_xor_sync(FULL, best, o);
                const int oi = __shfl_xor_sync(FULL, bidx, o);
                if (oi != 0x7fffffff && (bidx == 0x7fffffff || ob > best || (ob == best && oi < bidx))) {
                    best = ob; bidx = oi;
                }
            }
            if (lane == 0) final_state[blk] = bidx;
        }
        __syncthreads();
    }
}

// Parallel traceback, step 1: one warp per VCHUNK-column chunk.  The chunk's
// backpointer rows are staged in shared memory with coalesced 16-byte loads; lane j
// then follows them from state j at the chunk's last column to the state at the last
// column of the previous chunk: comp[chunk][j].
__global__ void __launch_bounds__(128)
viterbi_compose_kernel(const int64_t *__restrict__ off, const int64_t *__restrict__ chunk_off,
                       const int32_t *__restrict__ chunk_blk, const uint8_t *__restrict__ bp, int KP, int K,
                       int64_t n_chunks, uint8_t *__restrict__ comp) {
    extern __shared__ __align__(16) uint8_t sbp[];            // warps x VCHUNK x KP
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (g >= n_chunks) return;
    const int blk = chunk_blk[g];
    const int64_t c = g - chunk_off[blk];
    if (c == 0) return;                                       // nothing before the first chunk
    const int64_t beg = off[blk], T = off[blk + 1] - beg;
    const int64_t ts = c * VCHUNK, te = min(ts + VCHUNK, T) - 1;
    uint8_t *mine = sbp + (size_t)warp * VCHUNK * KP;
    const int n16 = (int)((te - ts + 1) * KP / 16);
    const uint4 *src = reinterpret_cast<const uint4 *>(bp + (size_t)(beg + ts) * KP);
    uint4 *dst = reinterpret_cast<uint4 *>(mine);
    for (int i = lane; i < n16; i += 32) dst[i] = __ldg(src + i);
    __syncwarp();
    for (int j = lane; j < K; j += 32) {
        int s = j;
        for (int64_t t = te; t >= ts; --t) s = mine[(size_t)(t - ts) * KP + s];
        comp[(size_t)g * KP + j] = (uint8_t)s;
    }
}

// Step 2, one warp per block: end state of every traceback chunk.  The chunk composites
// are staged in shared memory tile by tile (coalesced), so that the dependent chase costs a
// shared-memory load per chunk instead of an L2 round trip (0.33 -> 0.05 ms at config 2).
constexpr int VB_WARPS = 4, VB_BYTES = 8192;      // per-warp staging buffer
__global__ void __launch_bounds__(32 * VB_WARPS)
viterbi_boundary_kernel(const int64_t *__restrict__ off, const int64_t *__restrict__ chunk_off,
                        const uint8_t *__restrict__ comp, const int32_t *__restrict__ final_state,
                        int KP, int n_blocks, uint8_t *__restrict__ chunk_end) {
    __shared__ __align__(16) uint8_t stage[VB_WARPS][VB_BYTES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int blk = blockIdx.x * VB_WARPS + warp;
    if (blk >= n_blocks) return;
    const int64_t T = off[blk + 1] - off[blk];
    const int64_t nch = (T + VCHUNK - 1) / VCHUNK;
    const uint8_t *cmp = comp + (size_t)chunk_off[blk] * KP;
    uint8_t *ce = chunk_end + chunk_off[blk];
    uint8_t *mine = stage[warp];
    const int64_t tile = VB_BYTES / KP;                        // chunks per staging pass (KP <= 256)
    int s = final_state[blk];
    if (lane == 0) ce[nch - 1] = (uint8_t)s;
    for (int64_t c1 = nch; c1 > 1; c1 -= tile) {               // composites of chunks [c0, c1), c0 >= 1
        const int64_t c0 = max((int64_t)1, c1 - tile);
        const int n16 = (int)((c1 - c0) * KP / 16);
        const uint4 *src = reinterpret_cast<const uint4 *>(cmp + (size_t)c0 * KP);
        uint4 *dst = reinterpret_cast<uint4 *>(mine);
        for (int i = lane; i < n16; i += 32) dst[i] = __ldg(src + i);
        __syncwarp();
        if (lane == 0)
            for (int64_t c = c1 - 1; c >= c0; --c) {
                s = mine[(size_t)(c - c0) * KP + s];
                ce[c - 1] = (uint8_t)s;
            }
        s = __shfl_sync(FULL, s, 0);
        __syncwarp();
    }
}

// Step 3, one warp per chunk: stage the chunk's backpointers in shared memory, follow
// them from the chunk's known end state (optimizer.py:349-352) and write the path bytes
// with coalesced stores.
__global__ void __launch_bounds__(128)
viterbi_traceback_kernel(const int64_t *__restrict__ off, const int64_t *__restrict__ chunk_off,
                         const int32_t *__restrict__ chunk_blk, const uint8_t *__restrict__ bp,
                         const uint8_t *__restrict__ chunk_end, int KP, int64_t n_chunks,
                         uint8_t *__restrict__ path) {
    extern __shared__ __align__(16) uint8_t sbp[];            // warps x (VCHUNK x KP + VCHUNK)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (g >= n_chunks) return;
    const int blk = chunk_blk[g];
    const int64_t c = g - chunk_off[blk];
    const int64_t beg = off[blk], T = off[blk + 1] - beg;
    const int64_t ts = c * VCHUNK, te = min(ts + VCHUNK, T) - 1;
    uint8_t *mine = sbp + (size_t)warp * (VCHUNK * KP + VCHUNK);
    uint8_t *out = mine + VCHUNK * KP;
    const int n16 = (int)((te - ts + 1) * KP / 16);
    const uint4 *src = reinterpret_cast<const uint4 *>(bp + (size_t)(beg + ts) * KP);
    uint4 *dst = reinterpret_cast<uint4 *>(mine);
    for (int i = lane; i < n16; i += 32) dst[i] = __ldg(src + i);
    __syncwarp();
    if (lane == 0) {
        int s = chunk_end[g];
        out[te - ts] = (uint8_t)s;
        for (int64_t t = te; t > ts; --t) {
            s = mine[(size_t)(t - ts) * KP + s];
            out[t - 1 - ts] = (uint8_t)s;
        }
    }
    __syncwarp();
    uint8_t *p = path + beg + ts;
    for (int i = lane; i <= te - ts; i += 32) p[i] = out[i];
}

}  // namespace itr
