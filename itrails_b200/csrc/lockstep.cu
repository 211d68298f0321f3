// lockstep.cu — forward log-likelihood and posterior for 32 < K <= 96 when MANY blocks
// advance together: the batched (B x K)(K x K) contraction of north_star (3) on the FP64
// tensor cores (mma.sync.m8n8k4.f64 — tcgen05 has no FP64 kind).
//
// Replaces (reference paths relative to /root/reference/src/itrails):
//   forward, forward_loglik   optimizer.py:146-188
//   backward, post_prob       optimizer.py:192-238
//
// A CTA of NW warps walks EIGHT chains in lock step: the eight rows of the m8n8k4 A
// operand are eight state vectors, so one pass over the fragments of `a` (resident in
// registers, split over the warps by 8-state output chunks) advances all of them.  Every
// step: each warp reads the eight K-vectors from shared memory in A-fragment layout
// (row r = lane / 4, states 4q + lane % 4), runs NQ DMMAs per owned output chunk,
// multiplies by the emission of its rows' next columns and publishes its slice of the new
// vectors; one barrier per step.
//
// MODE 0 (log-likelihood): the eight rows are eight blocks (of the same parameter set).
// MODE 1 (posterior): rows 0-3 walk four blocks forwards and rows 4-7 walk the SAME four
// blocks backwards — both directions are "row vector times a" in the reference's
// orientation (optimizer.py:187, :210).  With f_t = alpha_{t-1} @ a (the forward DMMA's
// output) and g_t = beta_t * e_t (the backward row's next input), the posterior of column
// t is f_t * g_t up to a factor, so both directions run x <- (x @ a) * e(next column in my
// direction).  Until a row reaches the middle of its block it PARKS the DMMA's output in
// the result matrix itself (forward: f_t in columns [0, H); backward: beta_t in [H, T));
// past the middle it finds the other direction's vector waiting in the row it is about to
// produce, multiplies, and the row leaves normalised.  HBM traffic: the result written
// twice and read once (24 K bytes per column) instead of alpha + beta + combine (40 K),
// no second K x T buffer, and the flops of exactly one forward and one backward sweep.
// For an odd block length the backward row starts one step late so that neither direction
// ever reads a row in the step it is written (H = ceil(T / 2)).
// The per-column normaliser needs all K states, which live in NW warps: each warp leaves
// the partial sum of its products next to the exchanged vectors and the row is written one
// step later, after the barrier the exchange needs anyway.
// Scaling: every 8 steps each row is multiplied by the exact power of two that brings its
// largest element into [1, 2) (exponents travel with the exchange), as in the other sweeps.
#include "lockstep.h"

#include <math_constants.h>

namespace itr {

namespace {

__device__ __forceinline__ void dmma_884(double &d0, double &d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}

// Shared state of a CTA (static shared memory of lockstep_kernel).
template <int KT, int NW, int MODE>
struct LockShared {
    static constexpr int LD = KT + 4;             // LD mod 16 in {4, 12}: conflict-free fragment loads
    static constexpr int NCW = (KT / 8 + NW - 1) / NW;
    double xs[2][8][LD];
    unsigned his[8][NW];                          // written at the end of every 8th step, read at the next
    // partial normaliser of every lane: [row][4 warp + quad lane]; the row stride of 4 NW + 2
    // doubles spreads the eight rows' 16-byte reads over all banks (a stride of 4 NW would
    // put them on the same four: an 8-way conflict on every load)
    double psum[2][8][4 * NW + 2];
    double fin[8][NW];
    // staging slots of the other direction's parked vectors (cp.async, one step ahead)
    double pk[MODE == 1 ? 2 : 1][MODE == 1 ? 32 * NW : 1][MODE == 1 ? 2 * NCW : 1];
    int grp;
    unsigned slot;
};

// 1 / x for a normal, positive x: hardware seed (~20 bits) + two Newton steps; no slow path
// (the normalisers are sums of products of vectors that are rescaled into [1, 2) every 8
// steps, far from the denormal range)
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async8(double *dst_shared, const double *src, bool on) {
    asm volatile(
        "{ .reg .pred p; setp.ne.b32 p, %2, 0;\n"
        "  @p cp.async.ca.shared.global [%0], [%1], 8; }" ::"r"((unsigned)__cvta_generic_to_shared(dst_shared)),
        "l"(src), "r"((int)on)
        : "memory");
}
// predicated 8-byte store without a branch (ptxas otherwise wraps groups of predicated
// stores that share a condition into a branch region, which ends the basic block)
__device__ __forceinline__ void st_global_if(double *p, double v, bool on) {
    asm volatile(
        "{ .reg .pred p; setp.ne.b32 p, %2, 0;\n"
        "  @p st.global.f64 [%0], %1; }" ::"l"(p),
        "d"(v), "r"((int)on)
        : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// One group of chains, walked by a warp that owns NCH output chunks (wr, wr + NW, ...): NCH
// is a compile-time constant so that a step is ONE basic block — ptxas then interleaves the
// loads, the epilogue of the previous step's row and the DMMA stream.  Warps of a CTA may
// run different instantiations; they meet at the same hardware barrier (bar.sync 0).
template <int KT, int NW, int MODE, int NCH>
__device__ __forceinline__ void lockstep_group(LockShared<KT, NW, MODE> &sh, const ChainSet &cs,
                                               const double *__restrict__ A, const double *__restrict__ PI,
                                               const double *__restrict__ Et, int K, int KP,
                                               double *__restrict__ loglik, double *__restrict__ post, int g, int gps,
                                               int wr, int &cur_set, double (&B)[KT / 4][NCH], int dbg) {
    constexpr int NQ = KT / 4, LD = KT + 4;
    constexpr int CPG = MODE == 0 ? 8 : 4;        // chains per group
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = lane >> 2, c = lane & 3;
    const int set = g / gps;
    const int ci = (g % gps) * CPG + (MODE == 0 ? r : (r & 3));
    const bool have = ci < cs.n_blocks;
    const int blk = have ? cs.order[ci] : 0;
    const int64_t beg = cs.off[blk];
    const int T = have ? (int)(cs.off[blk + 1] - beg) : 0;
    const bool bwd = MODE == 1 && r >= 4;
    const int delay = (bwd && (T & 1)) ? 1 : 0;
    const int H = (T + 1) >> 1;
    const int nsteps = __reduce_max_sync(FULL, T + delay);   // every warp sees all eight rows
    if (set != cur_set) {                                     // B fragments: a[4q + c][8 nc + r]
        const double *As = A + (size_t)set * KP * KP;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int j = 0; j < NCH; ++j) B[q][j] = __ldg(As + (size_t)(4 * q + c) * KP + 8 * (wr + j * NW) + r);
        cur_set = set;
    }
    const uint16_t *sp = cs.sym + beg;
    const double *ets = Et + (size_t)set * NSYM * KP + 8 * wr + 2 * c;      // + sym * KP + 8 NW j
    const int s_base = 8 * wr + 2 * c;                                      // first owned state; chunk j adds 8 NW j
    bool live0[NCH], live1[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        live0[j] = s_base + 8 * NW * j < K;
        live1[j] = s_base + 8 * NW * j + 1 < K;
    }
    // column of this row at step u (clamped into the block: inactive steps read valid memory)
    const int Tm1 = max(T - 1, 0);
    auto col_of = [&](int u) {
        const int v = min(max(u - delay, 0), Tm1);
        return bwd ? Tm1 - v : v;
    };
    auto emis = [&](unsigned s, double (&e)[NCH][2]) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const double2 v = __ldg(reinterpret_cast<const double2 *>(ets + (size_t)s * KP + 8 * NW * j));
            e[j][0] = v.x;
            e[j][1] = v.y;
        }
    };
    double e_cur[NCH][2];
    emis(__ldg(sp + col_of(0)), e_cur);
    unsigned s_nxt = __ldg(sp + col_of(1));
    double start[NCH][2];                                  // start vector: pi going forward, 1 going backward
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        start[j][0] = bwd ? (live0[j] ? 1.0 : 0.0) : __ldg(PI + (size_t)set * KP + s_base + 8 * NW * j);
        start[j][1] = bwd ? (live1[j] ? 1.0 : 0.0) : __ldg(PI + (size_t)set * KP + s_base + 8 * NW * j + 1);
    }
    // the vectors start at zero (the first step of a row overrides the DMMA's output)
    for (int i = threadIdx.x; i < 2 * 8 * LD; i += 32 * NW) (&sh.xs[0][0][0])[i] = 0.0;
    for (int i = threadIdx.x; i < 2 * 8 * (4 * NW + 2); i += 32 * NW) (&sh.psum[0][0][0])[i] = 1.0;
    if (threadIdx.x < 8 * NW) (&sh.his[0][0])[threadIdx.x] = 0u;
    cta_barrier();
    int buf = 0;
    long long shift = 0;
    bool pend = false;                                     // a finished row waits for its normaliser
    double *pend_row = post;
    double pv[NCH][2];
#pragma unroll
    for (int j = 0; j < NCH; ++j) pv[j][0] = pv[j][1] = 0.0;
    const int first_fin = bwd ? H - 1 : H;                 // first column this row finishes
    if (MODE == 1) cp_async_commit();                      // (keeps the group count of step 0 uniform)

    // A step is written without data-dependent branches (predicated stores, selects) so that
    // it stays one basic block; the two rare events — exponents published every 8th step,
    // a forward row reaching its last column in MODE 0 — sit behind warp-uniform branches.
#pragma unroll 1
    for (int u = 0; u < nsteps; ++u) {
        const int v = u - delay;
        const bool active = v >= 0 && v < T;
        const int col = bwd ? T - 1 - v : v;
        double *row = (MODE == 1) ? post + (size_t)(beg + col) * K + s_base : nullptr;
        const bool parking = MODE == 1 && (bwd ? col >= H : col < H);
        const bool finishing = MODE == 1 && active && !parking;
        // ---- loads whose latency hides under the DMMAs: the vectors, the next column's
        // emission row, the symbol after that
        double xq[NQ];
        {
            const double *xr = &sh.xs[buf][r][c];
#pragma unroll
            for (int q = 0; q < NQ; ++q) xq[q] = xr[4 * q];
        }
        // The other direction's vector for the row finished in the NEXT step goes into this
        // thread's staging slots now (cp.async: no registers, a whole step of latency hidden;
        // it was parked at least three steps ago).  The first finishing step of a row is the
        // exception: its partner was parked only one step ago, so it is fetched in the step
        // itself — as its own (usually empty) copy group, committed first, so that the one
        // wait before the epilogue ("all but the newest group") covers it.
        if (MODE == 1) {
            const bool fin_first = finishing && col == first_fin;
            double *slot0 = &sh.pk[u & 1][threadIdx.x][0];
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                cp_async8(slot0 + 2 * j, row + 8 * NW * j, fin_first && live0[j]);
                cp_async8(slot0 + 2 * j + 1, row + 8 * NW * j + 1, fin_first && live1[j]);
            }
            cp_async_commit();
            const int v1 = v + 1, col1 = bwd ? T - 1 - v1 : v1;
            const bool fin_next = v1 >= 0 && v1 < T && !(bwd ? col1 >= H : col1 < H) && col1 != first_fin && !(dbg & 1);
            const double *row1 = bwd ? row - K : row + K;
            double *slot1 = &sh.pk[(u + 1) & 1][threadIdx.x][0];
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                cp_async8(slot1 + 2 * j, row1 + 8 * NW * j, fin_next && live0[j]);
                cp_async8(slot1 + 2 * j + 1, row1 + 8 * NW * j + 1, fin_next && live1[j]);
            }
            cp_async_commit();
        }
        double e_nxt[NCH][2];
        emis(s_nxt, e_nxt);
        const unsigned s_n2 = __ldg(sp + col_of(u + 2));
        // ---- common power-of-two scale of the vectors read in this step (their exponents
        // were published with them at the end of every 8th step)
        double sc = 1.0;
        if ((u & 7) == 0 && u > 0) {
            unsigned hi = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) hi = max(hi, sh.his[r][w]);
            const int ex = (int)(hi >> 20);
            const bool on = ex != 0 && ex < 0x7ff;
            sc = on ? __hiloint2double((2046 - ex) << 20, 0) : 1.0;
            shift += (on && active) ? ex - 1023 : 0;
        }
        // ---- normaliser of the row finished in the previous step (partial sums of every lane)
        double inv = 1.0;
        if (MODE == 1) {
            const double2 *ps = reinterpret_cast<const double2 *>(&sh.psum[buf][r][0]);
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int w = 0; w < 2 * NW; ++w) {
                const double2 q2 = ps[w];
                t0 += q2.x;
                t1 += q2.y;
            }
            inv = fast_rcp(t0 + t1);
        }
        // ---- per chunk: DMMAs, then the chunk's epilogue (which slides under the next chunk's
        // DMMAs; only the last chunk's is exposed)
        const bool first = v == 0;
        if (MODE == 1) cp_async_wait_all_but_one();         // this step's slots are filled
        const double *slot = &sh.pk[u & 1][threadIdx.x][0];
        // one store per element and step: a parking row leaves the DMMA's output in its row, a
        // finishing row writes the row it finished in the PREVIOUS step, now normalised
        double *dst = parking ? row : pend_row;
        const bool st_on = (parking ? active : pend) && !(dbg & 2);
        double part = 0.0;
        unsigned hmax = 0;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            // y = x @ a for the chunk: two interleaved accumulator pairs, so that consecutive
            // DMMAs never wait for each other
            double ya0 = 0.0, ya1 = 0.0, yb0 = 0.0, yb1 = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; q += 2) {
                dmma_884(ya0, ya1, xq[q], B[q][j]);
                dmma_884(yb0, yb1, xq[q + 1], B[q + 1][j]);
            }
            const double y0 = first ? start[j][0] : (ya0 + yb0) * sc, y1 = first ? start[j][1] : (ya1 + yb1) * sc;
            const double x0 = active ? y0 * e_cur[j][0] : 0.0, x1 = active ? y1 * e_cur[j][1] : 0.0;
            *reinterpret_cast<double2 *>(&sh.xs[buf ^ 1][r][s_base + 8 * NW * j]) = make_double2(x0, x1);
            if (MODE == 1) {
                st_global_if(dst + 8 * NW * j, parking ? y0 : pv[j][0] * inv, st_on && live0[j]);
                st_global_if(dst + 8 * NW * j + 1, parking ? y1 : pv[j][1] * inv, st_on && live1[j]);
                const double2 pk2 = *reinterpret_cast<const double2 *>(slot + 2 * j);
                pv[j][0] = (finishing && live0[j]) ? pk2.x * x0 : 0.0;
                pv[j][1] = (finishing && live1[j]) ? pk2.y * x1 : 0.0;
                part += pv[j][0] + pv[j][1];
            } else {
                part += x0 + x1;
            }
            hmax = max(hmax, max((unsigned)__double2hiint(x0), (unsigned)__double2hiint(x1)));
        }
        if (MODE == 1) {
            sh.psum[buf ^ 1][r][4 * wr + c] = part;
            pend = finishing;
            pend_row = row;
        } else if (__any_sync(FULL, active && v == T - 1)) {
            part += __shfl_xor_sync(FULL, part, 1);
            part += __shfl_xor_sync(FULL, part, 2);
            if (active && v == T - 1 && c == 0) sh.fin[r][wr] = part;   // sum of the last forward vector (this warp's states)
        }
        if (((u + 1) & 7) == 0) {
            hmax = max(hmax, __shfl_xor_sync(FULL, hmax, 1));
            hmax = max(hmax, __shfl_xor_sync(FULL, hmax, 2));
            if (c == 0) sh.his[r][wr] = hmax;
        }
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            e_cur[j][0] = e_nxt[j][0];
            e_cur[j][1] = e_nxt[j][1];
        }
        s_nxt = s_n2;
        cta_barrier();
        buf ^= 1;
    }
    if (MODE == 1) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        const double2 *ps = reinterpret_cast<const double2 *>(&sh.psum[buf][r][0]);
        double t0 = 0.0, t1 = 0.0;                          // same order as inside the loop: copies of a block agree bit for bit
#pragma unroll
        for (int w = 0; w < 2 * NW; ++w) {
            const double2 q2 = ps[w];
            t0 += q2.x;
            t1 += q2.y;
        }
        const double inv = fast_rcp(t0 + t1);
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            if (pend && live0[j]) pend_row[8 * NW * j] = pv[j][0] * inv;
            if (pend && live1[j]) pend_row[8 * NW * j + 1] = pv[j][1] * inv;
        }
    } else if (warp == 0 && c == 0 && have && loglik) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) tot += sh.fin[r][w];
        loglik[(size_t)set * cs.n_blocks + blk] = log(tot) + (double)shift * 0.6931471805599453094;
    }
}

// Scratch of one launch (zeroed before it): sm_slot[256] arrival counters per SM,
// sm_first[256] group taken by the first CTA of an SM (+1), then one "taken" flag per group.
//
// Which CTA walks which group.  Groups are sorted longest first.  Two CTAs share an SM, and a
// pair of co-resident CTAs takes about twice as long per step as a CTA that has the SM to
// itself — so the longest groups should run alone and the others in pairs of equal total
// length.  With S = number of SMs the first CTA to arrive on an SM takes group S - 1 - i (i =
// its arrival order among first CTAs) and the second CTA on that SM the group 2S - 1 minus
// that: pairs (S-1, S), (S-2, S+1), ... have near-constant total length when lengths fall off
// evenly, and groups 0 ... 2S - 1 - n_groups, the longest, stay alone.  Everything else (more
// than 2S groups, or a placement the rule did not foresee) is drained from a longest-first
// queue; a flag per group makes sure each is walked exactly once.
template <int KT, int NW, int MODE>
__global__ void __launch_bounds__(32 * NW, 1)
lockstep_kernel(ChainSet cs, const double *__restrict__ A, const double *__restrict__ PI,
                const double *__restrict__ Et, int K, int KP, double *__restrict__ loglik,
                double *__restrict__ post, unsigned int *__restrict__ scratch, int n_sms, int dbg) {
    static_assert(KT % 8 == 0 && KT <= 96, "KT is K rounded up to a multiple of 8");
    constexpr int NQ = KT / 4, NC = KT / 8, NCW = (NC + NW - 1) / NW;
    constexpr int CPG = MODE == 0 ? 8 : 4;
    static_assert(NC >= NW, "every warp owns at least one chunk");
    __shared__ __align__(16) LockShared<KT, NW, MODE> sh;
    const int warp = threadIdx.x >> 5;
    // Output chunks are dealt to the warps round robin; when NC is not a multiple of NW the
    // first warps carry one chunk more.  The hardware already rotates the warps of CTAs that
    // share an SM over its sub-partitions (tools/warp_map.cu: %warpid 0 1 2 3 for the first
    // CTA, 5 6 7 4 for the second), so the heavier warps of two CTAs do not meet.
    const int wr = warp;
    const bool full = wr + (NCW - 1) * NW < NC;  // NCW chunks (else NCW - 1)
    const int gps = (cs.n_blocks + CPG - 1) / CPG;        // groups per parameter set
    const int n_groups = gps * cs.n_sets;
    unsigned int *sm_slot = scratch, *sm_first = scratch + 256, *taken = scratch + 512;
    int cur_set = -1;
    double B[NQ][NCW];
    bool static_try = !(dbg & 16);
    for (;;) {
        cta_barrier();
        if (threadIdx.x == 0) {
            int g = -1;
            if (static_try) {
                unsigned smid;
                asm("mov.u32 %0, %%smid;" : "=r"(smid));
                smid = min(smid, 254u);
                const unsigned slot = atomicAdd(sm_slot + smid, 1u);
                if (slot == 0) {
                    const int i = (int)atomicAdd(sm_first + 255, 1u);      // arrival order among first CTAs
                    g = n_sms - 1 - i;
                    if (g < 0 || g >= n_groups) g = -1;
                    atomicExch(sm_first + smid, (unsigned)(g + 2));         // 1: none, g + 2: group g
                } else if (slot == 1) {
                    unsigned f;
                    while ((f = atomicAdd(sm_first + smid, 0u)) == 0u) { }
                    g = f >= 2u ? 2 * n_sms - 1 - (int)(f - 2u) : -1;
                    if (g < 0 || g >= n_groups) g = -1;
                }
                if (g >= 0 && atomicCAS(taken + g, 0u, 1u) != 0u) g = -1;
            }
            while (g < 0) {
                g = (int)atomicAdd(cs.queue, 1u);
                if (g >= n_groups) break;
                if (atomicCAS(taken + g, 0u, 1u) != 0u) g = -1;
            }
            sh.grp = g;
        }
        static_try = false;
        cta_barrier();
        const int g = sh.grp;
        if (g >= n_groups) break;
        if (NC % NW == 0 || full) {
            lockstep_group<KT, NW, MODE, NCW>(sh, cs, A, PI, Et, K, KP, loglik, post, g, gps, wr, cur_set, B, dbg);
        } else if constexpr (NC % NW != 0) {
            lockstep_group<KT, NW, MODE, NCW - 1>(sh, cs, A, PI, Et, K, KP, loglik, post, g, gps, wr, cur_set,
                                                  reinterpret_cast<double (&)[NQ][NCW - 1]>(B), dbg);
        }
    }
}

__global__ void zero_slots_kernel(unsigned int *p, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0u;
}

template <int KT, int NW, int MODE>
cudaError_t launch_t(const ChainSet &cs, const double *A, const double *PI, const double *Et, int K, int KP,
                     double *loglik, double *post, unsigned int *scratch, int sms, cudaStream_t st) {
    constexpr int CPG = MODE == 0 ? 8 : 4;
    const int64_t n_groups = (int64_t)((cs.n_blocks + CPG - 1) / CPG) * cs.n_sets;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_groups, (int64_t)sms * 2));
    const char *dbg = getenv("ITR_LOCKSTEP_DBG");                             // experiments
    const int n_zero = 512 + (int)n_groups;
    zero_slots_kernel<<<(n_zero + 1023) / 1024, 1024, 0, st>>>(scratch, n_zero);
    lockstep_kernel<KT, NW, MODE><<<grid, 32 * NW, 0, st>>>(cs, A, PI, Et, K, KP, loglik, post, scratch, sms, dbg ? atoi(dbg) : 0);
    return cudaGetLastError();
}

template <int MODE>
cudaError_t dispatch(const ChainSet &cs, const double *A, const double *PI, const double *Et, int K, int KP,
                     double *loglik, double *post, unsigned int *scratch, int sms, cudaStream_t st) {
#define LS(KT, NW) return launch_t<KT, NW, MODE>(cs, A, PI, Et, K, KP, loglik, post, scratch, sms, st)
    switch ((K + 7) / 8) {
        case 5: LS(40, 4);
        case 6: LS(48, 4);
        case 7: LS(56, 4);
        case 8: LS(64, 4);
        case 9: LS(72, 4);
        case 10: LS(80, 8);
        case 11: LS(88, 8);
        case 12: LS(96, 8);
        default: return cudaErrorInvalidValue;
    }
#undef LS
}

}  // namespace

bool lockstep_supports(int K) { return K > 32 && K <= 96; }

cudaError_t launch_lockstep_loglik(const ChainSet &cs, const double *A, const double *PI, const double *Et, int K,
                                   int KP, double *loglik, unsigned int *scratch, int sms, cudaStream_t st) {
    return dispatch<0>(cs, A, PI, Et, K, KP, loglik, nullptr, scratch, sms, st);
}

cudaError_t launch_lockstep_posterior(const ChainSet &cs, const double *A, const double *PI, const double *Et, int K,
                                      int KP, double *post, unsigned int *scratch, int sms, cudaStream_t st) {
    return dispatch<1>(cs, A, PI, Et, K, KP, nullptr, post, scratch, sms, st);
}

}  // namespace itr
