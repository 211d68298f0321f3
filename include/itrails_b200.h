/*
 * itrails_b200.h — C ABI of libitrails_b200.so (hand-written CUDA, sm_100a).
 *
 * The reference (trails-phylogeny/itrails) is pure Python and has no FFI layer; its
 * boundary for the coalescent-HMM hot path is a set of Python functions.  Each entry
 * point below names the reference function it replaces (paths relative to
 * /root/reference/src/itrails).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add to call these instead of the Python/numba
 * implementations.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types cross the boundary.
 *   - Every function returns 0 on success or a negative itr_status; the message is
 *     available from itr_last_error(ctx) (never NULL).  No exceptions cross the ABI.
 *   - Host buffers are caller-owned.  Device buffers live inside the opaque context
 *     and are released by itr_destroy.  One context drives one GPU; a context is not
 *     thread-safe.  Multi-GPU = one process (and context) per GPU; the only exchange
 *     is the scalar sum of log-likelihood partials, which the host side performs with
 *     an NCCL all-reduce (itrails_b200/distributed.py).
 *   - All floating point is IEEE binary64.  Matrices are row-major.
 *   - Symbols are the reference's observed-state indices 0..624
 *     (read_data.py:6-24: nucleotide order A,C,T,G; 256.. contain at least one N).
 *   - There is NO CPU fallback: every compute entry point fails with
 *     ITR_ERR_CUDA when no sm_100-class device is usable.
 */
#ifndef ITRAILS_B200_H
#define ITRAILS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct itr_ctx itr_ctx;

enum itr_status {
    ITR_OK = 0,
    ITR_ERR_ARG = -1,      /* invalid argument (message says which)            */
    ITR_ERR_STATE = -2,    /* call order violated (e.g. loglik before set_model) */
    ITR_ERR_CUDA = -3,     /* CUDA runtime error / no usable device            */
    ITR_ERR_NOMEM = -4,    /* host or device allocation failed                 */
    ITR_ERR_UNSUPPORTED = -5,
    ITR_ERR_IO = -6        /* a file could not be opened / written              */
};

#define ITR_N_SYMBOLS 625   /* read_data.py:6-24 */
#define ITR_N_PLAIN 256
#define ITR_MAX_STATES 255  /* uint8 Viterbi backpointers / paths */

/* ---- lifetime ------------------------------------------------------------------ */

/* Create a context bound to CUDA device `device` (ordinal as seen by this process). */
int itr_create(int device, itr_ctx **out);
void itr_destroy(itr_ctx *ctx);
/* Message of the last failing call on this context ("" if none).  With ctx == NULL,
 * returns the message of the last failing itr_create in this thread. */
const char *itr_last_error(const itr_ctx *ctx);
/* Library/ABI version (major*1000 + minor). */
int itr_version(void);

/* ---- alignment blocks ------------------------------------------------------------
 * Replaces the `V_lst` argument every reference wrapper takes (list of int64 arrays,
 * one per MAF block, from read_data.py:94-117 maf_parser).  `sym` holds the blocks
 * back to back as uint16; block b is sym[block_offsets[b] .. block_offsets[b+1]).
 * One host->device copy; blocks are scheduled longest-first on the device.
 * Every block must have at least one column. */
int itr_load_blocks(itr_ctx *ctx, const uint16_t *sym, const int64_t *block_offsets,
                    int64_t n_blocks);
/* Same, from int64 symbols exactly as maf_parser returns them (values checked). */
int itr_load_blocks_i64(itr_ctx *ctx, const int64_t *sym, const int64_t *block_offsets,
                        int64_t n_blocks);

/* ---- MAF ingest (host only; needs no GPU) ------------------------------------------
 * Replaces maf_parser (read_data.py:94-117) and parse_coordinates (read_data.py:146-220).
 * species: the four names in the reference's sp_lst order (A, B sister; C; D outgroup);
 * ref: species whose coordinates polarise the columns, or NULL for symbols only;
 * n_threads <= 0 uses every host core.  On success *out owns:
 *   symbols      uint16 observed-state indices of every kept block back to back
 *                (kept = all four species present), with num_blocks + 1 offsets — exactly
 *                the arguments of itr_load_blocks;
 *   coordinates  (ref != NULL) int64 per column of every block whose rows hold exactly four
 *                listed species, -9 at gaps / when ref is absent, with their own offsets.
 * A character outside A,C,G,T,N,- (any case) in a kept block is an error, as
 * list.index raises in the reference.  err (nullable) receives the message. */
typedef struct itr_maf itr_maf;
int itr_maf_read(const char *path, const char *const species[4], const char *ref, int n_threads,
                 itr_maf **out, char *err, int err_cap);
void itr_maf_free(itr_maf *m);
int64_t itr_maf_num_blocks(const itr_maf *m);
int64_t itr_maf_num_columns(const itr_maf *m);
/* Copies the symbols of all kept blocks (itr_maf_num_columns values) and / or the coordinates
 * (itr_maf_coord_offsets()[itr_maf_num_coord_blocks] values) into the caller's arrays, with
 * n_threads host threads (0: all); either pointer may be NULL.  Prefer it to the pointer
 * accessors below, which concatenate inside the handle on first use. */
int itr_maf_export(const itr_maf *m, uint16_t *sym_out, int64_t *coord_out, int n_threads);
const uint16_t *itr_maf_symbols(const itr_maf *m);
const int64_t *itr_maf_offsets(const itr_maf *m);
int64_t itr_maf_num_coord_blocks(const itr_maf *m);
const int64_t *itr_maf_coordinates(const itr_maf *m);
const int64_t *itr_maf_coord_offsets(const itr_maf *m);

/* ---- model -----------------------------------------------------------------------
 * Accepts (a, b, pi) as returned by the reference's trans_emiss_calc
 * (get_trans_emiss.py:8-170): a is n_sets x K x K, b is n_sets x K x 256,
 * pi is n_sets x K.  The library expands b to the K x 625 emission table
 * E[:, s] = sum over the N-free symbols s marginalises (optimizer.py:182 with
 * read_data.py:46-67) on the device.  n_sets > 1 evaluates several parameter sets
 * over the same blocks in one launch (log-likelihood only). */
int itr_set_model(itr_ctx *ctx, int n_sets, int K, const double *a, const double *b,
                  const double *pi);

/* Build (a, b, pi) on the GPU for n_sets parameter sets.
 * Replaces trans_emiss_calc (get_trans_emiss.py:8-170).  params is n_sets x 9:
 * {t_A, t_B, t_C, t_2, t_upper, t_out, N_AB, N_ABC, r} in the reference's scaled
 * units.  cut_AB (n_int_AB+1) / cut_ABC (n_int_ABC+1, last = +inf) are normalised
 * cutpoints shared by all sets, or NULL for the reference's "standard" quantile
 * cutpoints (cutpoints.py:5-45).  Outputs (nullable): a n_sets x K x K,
 * b n_sets x K x 256, pi n_sets x K, hidden K x 3 (topology, i, j) in sorted order.
 * The model is also left installed on the device as if by itr_set_model. */
int itr_build_model(itr_ctx *ctx, int n_sets, const double *params, int n_int_AB,
                    int n_int_ABC, const double *cut_AB, const double *cut_ABC,
                    double *a, double *b, double *pi, int32_t *hidden);
/* Number of hidden states for a discretisation
 * (get_emission_prob_mat.py:789-791). */
int itr_num_states(int n_int_AB, int n_int_ABC);
/* Parameter-independent plan of the model build for a discretisation (host only, needs
 * no GPU): number of hidden states, matrix exponentials per parameter set, block
 * mat-vec operations and peak number of live path keys
 * (run_markov_chain_ABC.py:350-518 bookkeeping), and the sorted hidden-state tuples
 * (get_trans_emiss.py:148-153; hidden is K x 3, nullable).  Returns 0 or ITR_ERR_ARG. */
int itr_plan_info(int n_int_AB, int n_int_ABC, int32_t *K, int32_t *n_mats, int64_t *n_ops,
                  int64_t *n_keys, int32_t *hidden);

/* ---- recursions ------------------------------------------------------------------ */

/* Forward log-likelihood.  Replaces loglik_wrapper / loglik_wrapper_par
 * (optimizer.py:40-116): total[s] = sum over blocks (in block order) of
 * forward_loglik (optimizer.py:146-162).  per_block (nullable) is n_sets x n_blocks. */
int itr_loglik(itr_ctx *ctx, double *total, double *per_block);

/* Viterbi decoding of every block with parameter set 0.  Replaces viterbi_wrapper
 * (optimizer.py:357-377).  To be bit-exact the caller supplies the three tables
 * whose FP64 values decide the path, computed with NumPy exactly as the reference
 * does (optimizer.py:323-330): log_a = log(a) (K x K), log_E = log(E) (K x 625),
 * omega0 = log(pi * E[:, V0]) per block (n_blocks x K).  path receives sum(T) state
 * indices (uint8), blocks back to back.  path may be NULL to leave the result on the
 * device (see itr_viterbi_fetch). */
int itr_viterbi(itr_ctx *ctx, const double *log_a, const double *log_E,
                const double *omega0, uint8_t *path);
int itr_viterbi_fetch(itr_ctx *ctx, uint8_t *path);
/* Columns [col0, col0 + n_cols) of the concatenated path kept on the device: lets a
 * writer stream a chromosome-scale result block by block (workflow_viterbi.py:692-743). */
int itr_viterbi_fetch_range(itr_ctx *ctx, int64_t col0, int64_t n_cols, uint8_t *path);

/* Posterior decoding of every block with parameter set 0.  Replaces
 * post_prob_wrapper (optimizer.py:241-262), including the reference's backward
 * orientation (optimizer.py:210, row vector times a).  post receives sum(T) x K
 * doubles, blocks back to back; NULL leaves the result on the device
 * (see itr_posterior_fetch). */
int itr_posterior(itr_ctx *ctx, double *post);
int itr_posterior_fetch(itr_ctx *ctx, double *post);
/* Rows [col0, col0 + n_cols) of the posterior matrix kept on the device (n_cols x K
 * doubles): the 250 Mb x 27 result is 54 GB and is consumed block by block by the
 * CSV writer (workflow_posterior.py:697-716). */
int itr_posterior_fetch_range(itr_ctx *ctx, int64_t col0, int64_t n_cols, double *post);

/* Posterior decoding streamed to the host (replaces post_prob_wrapper for results that
 * do not fit in host memory: 250 Mb x 27 states = 54 GB).  The posterior is computed on
 * the device in contiguous ranges of blocks and downloaded behind the computation in
 * pieces of at most slot_cols columns (a piece never spans more than the ring slot) into
 * ring[(piece % n_slots) * slot_cols * K ...]; ring holds n_slots * slot_cols * K doubles
 * and should be page-locked.  sink (nullable) is called from the calling thread, in
 * ascending column order, once per piece: rows = n_cols x K doubles for the columns
 * [col0, col0 + n_cols) of the concatenated alignment; the slot is reused after it
 * returns; a non-zero return stops the stream (ITR_ERR_IO).  With sink == NULL the pieces
 * are only delivered into the ring.  The full result also stays on the device
 * (itr_posterior_fetch_range).  In deferred mode (itr_set_async) a preceding
 * itr_posterior(ctx, NULL) is the computation that gets drained: enqueue the posterior
 * first, the other recursions behind it, then call this to download behind the kernels. */
typedef int (*itr_rows_sink)(void *user, int64_t col0, int64_t n_cols, const double *rows);
int itr_posterior_stream(itr_ctx *ctx, double *ring, int64_t slot_cols, int n_slots,
                         itr_rows_sink sink, void *user);

/* ---- result writers (host C++, thread pool) --------------------------------------
 * Write `{prefix}.posterior.csv` exactly as the reference's csv.writer loop does
 * (workflow_posterior.py:697-716): header `alignment_block_idx,position_idx,prob_state_0…`,
 * one row per column, "\r\n" line ends, integers in decimal, probabilities as Python's
 * repr(float) (shortest round-trip digits; exponent form iff the decimal exponent is
 * < -4 or >= 16) — byte-identical files.  positions (nullable) holds one int64 per
 * column of the loaded alignment (the reference coordinates of parse_coordinates);
 * NULL writes 0..T-1 within every block.  n_threads <= 0 uses every host core.
 * itr_posterior_write_csv streams the posterior kept on the device by the last
 * itr_posterior block by block (download of block i+1 overlaps formatting of block i), so
 * a chromosome-scale result never has to exist in host memory. */
int itr_posterior_write_csv(itr_ctx *ctx, const char *path, const int64_t *positions, int n_threads);
/* Sharded form (one process per GPU, each holding an LPT share of the blocks;
 * workflow_posterior.py:693-716 writes ONE file with global block indices): block_ids
 * (nullable, one per loaded block) is the index printed for each block, write_header == 0
 * leaves the header line out (a part file), block_bytes (nullable, one per loaded block)
 * receives the bytes each block's rows took, so that the parts can be spliced in global
 * block order. */
int itr_posterior_write_csv_ex(itr_ctx *ctx, const char *path, const int64_t *positions,
                               const int64_t *block_ids, int write_header, int64_t *block_bytes,
                               int n_threads);
/* The same writer on a host matrix (no GPU context needed): post is sum(T) x K. */
int itr_csv_posterior_host(const char *path, int K, int64_t n_blocks, const int64_t *offsets,
                           const int64_t *positions, const double *post, int n_threads);
/* Host-matrix writer with the sharded options of itr_posterior_write_csv_ex. */
int itr_csv_posterior_host_ex(const char *path, int K, int64_t n_blocks, const int64_t *offsets,
                              const int64_t *positions, const double *post, const int64_t *block_ids,
                              int write_header, int64_t *block_bytes, int n_threads);
/* repr(float) of one value into out (cap >= 32, NUL terminated); returns the length. */
int itr_csv_format_double(double x, char *out, int cap);

/* ---- overlapping the recursions ---------------------------------------------------
 * Every recursion runs on its own CUDA stream.  By default each call returns when its
 * result is complete.  With itr_set_async(ctx, 1), itr_loglik / itr_viterbi /
 * itr_posterior only enqueue their work and return; independent recursions then run
 * concurrently on the device, and all host outputs (total, per_block, path, post) are
 * valid after itr_sync(ctx).  Device-to-host copies overlap only into page-locked
 * buffers; into pageable memory they complete before the call returns.  Changing the
 * blocks or the model always waits for enqueued work first. */
/* In deferred mode itr_build_model with a == b == pi == NULL also only enqueues the build. */
int itr_set_async(itr_ctx *ctx, int on);
int itr_sync(itr_ctx *ctx);

/* ---- introspection --------------------------------------------------------------- */

enum itr_phase {
    ITR_PH_LOGLIK = 0,
    ITR_PH_VITERBI_FWD = 1,
    ITR_PH_VITERBI_TRACE = 2,
    ITR_PH_POST_FWD = 3,
    ITR_PH_POST_BWD = 4,
    ITR_PH_MODEL = 5,
    ITR_PH_EMIT_TABLE = 6,
    ITR_PH_POST_COMBINE = 7,
    ITR_PH_POST_TOTAL = 8,   /* forward || backward, then combine (wall on the device) */
    ITR_PH_COUNT = 9
};
/* Device time (CUDA events on the launching stream) of the most recent run of a
 * phase, in milliseconds; negative if the phase has not run. */
double itr_phase_ms(itr_ctx *ctx, int phase);
/* Kernels launched by this context since creation. */
int64_t itr_launch_count(const itr_ctx *ctx);
/* Launches of the lock-step FP64 tensor-core sweeps (32 < K <= 96 with many blocks:
 * optimizer.py:146-238 for eight chains per CTA) since creation — lets a caller or a test
 * see which path a call took. */
int64_t itr_lockstep_launch_count(const itr_ctx *ctx);
int64_t itr_total_columns(const itr_ctx *ctx);
int64_t itr_num_blocks(const itr_ctx *ctx);
/* Fills name (NUL-terminated, <= cap bytes), SM count and compute capability. */
int itr_device_info(itr_ctx *ctx, char *name, int cap, int *sm_count, int *cc_major,
                    int *cc_minor);

#ifdef __cplusplus
}
#endif
#endif /* ITRAILS_B200_H */
