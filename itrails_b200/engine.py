"""Engine: one GPU context (itr_ctx) with resident alignment blocks and model."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib as L


class Engine:
    """Owns one ``itr_ctx`` (one GPU).  Typical use::

        eng = Engine(device=0)
        eng.load_blocks(V_lst)              # list of int arrays, values 0..624
        eng.set_model(a, b, pi)             # from trans_emiss_calc (or build_model)
        ll = eng.loglik()
        paths = eng.viterbi(log_a, log_E, omega0)
        post = eng.posterior()
    """

    def __init__(self, device=0):
        self._lib = L.load()
        ctx = ctypes.c_void_p()
        rc = self._lib.itr_create(int(device), ctypes.byref(ctx))
        if rc != 0:
            L.check(self._lib, None, rc)
        self._ctx = ctx
        self.device = int(device)
        self._offsets = None
        self.K = 0
        self.n_sets = 0
        self._pending = []      # host arrays a deferred call will still write into (alive until sync)

    # -- lifetime ----------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.itr_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        L.check(self._lib, self._ctx, rc)

    # -- data --------------------------------------------------------------------
    @staticmethod
    def pack_blocks(V_lst):
        """list of per-block symbol arrays -> (uint16 symbols, int64 offsets).
        Values outside 0..624 raise ValueError, like ``list.index`` does in the
        reference's maf_parser (read_data.py:113-115)."""
        if len(V_lst) == 0:
            raise ValueError("V_lst is empty")
        off = np.zeros(len(V_lst) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(v) for v in V_lst])
        if np.any(np.diff(off) <= 0):
            raise ValueError("every alignment block must have at least one column")
        cat = np.concatenate([np.asarray(v) for v in V_lst])
        if cat.dtype.kind not in "iu":
            raise ValueError("symbols must be integers")
        if cat.size and (cat.min() < 0 or cat.max() > 624):
            raise ValueError("symbol outside 0..624")
        return np.ascontiguousarray(cat, dtype=np.uint16), off

    def load_blocks(self, V_lst):
        sym, off = self.pack_blocks(V_lst)
        self.load_packed(sym, off)

    def load_packed(self, sym, off):
        sym = np.ascontiguousarray(sym, dtype=np.uint16)
        off = np.ascontiguousarray(off, dtype=np.int64)
        self._ck(self._lib.itr_load_blocks(self._ctx, L.as_ptr(sym, ctypes.c_uint16),
                                           L.as_ptr(off, ctypes.c_int64), len(off) - 1))
        self._offsets = off.copy()

    @property
    def n_blocks(self):
        return int(self._lib.itr_num_blocks(self._ctx))

    @property
    def n_columns(self):
        return int(self._lib.itr_total_columns(self._ctx))

    # -- model -------------------------------------------------------------------
    def set_model(self, a, b, pi):
        a, b, pi = L.c_f64(a), L.c_f64(b), L.c_f64(pi)
        if a.ndim == 2:
            a, b, pi = a[None], b[None], pi[None]
        n_sets, K = a.shape[0], a.shape[1]
        if a.shape != (n_sets, K, K) or b.shape != (n_sets, K, 256) or pi.shape != (n_sets, K):
            raise ValueError(f"inconsistent model shapes a{a.shape} b{b.shape} pi{pi.shape}")
        self._ck(self._lib.itr_set_model(self._ctx, n_sets, K, L.as_ptr(a, ctypes.c_double),
                                         L.as_ptr(b, ctypes.c_double), L.as_ptr(pi, ctypes.c_double)))
        self.K, self.n_sets = K, n_sets

    def build_model(self, params, n_int_AB, n_int_ABC, cut_AB=None, cut_ABC=None, fetch=True):
        """Batched GPU model build (replaces trans_emiss_calc).  ``params`` is
        (n_sets, 9) = t_A,t_B,t_C,t_2,t_upper,t_out,N_AB,N_ABC,r (scaled units).
        Returns (a, b, pi, hidden) as arrays (or None, None, None, hidden if not fetch)."""
        params = np.atleast_2d(L.c_f64(params))
        if params.shape[1] != 9:
            raise ValueError("params must have 9 columns")
        n_sets = params.shape[0]
        K = self._lib.itr_num_states(int(n_int_AB), int(n_int_ABC))
        if K < 1:
            raise ValueError("n_int_AB and n_int_ABC must be >= 1")
        cab = None if cut_AB is None else L.c_f64(cut_AB)
        cabc = None if cut_ABC is None else L.c_f64(cut_ABC)
        if cab is not None and cab.shape != (n_int_AB + 1,):
            raise ValueError("cut_AB must have n_int_AB+1 entries")
        if cabc is not None and cabc.shape != (n_int_ABC + 1,):
            raise ValueError("cut_ABC must have n_int_ABC+1 entries")
        a = np.empty((n_sets, K, K)) if fetch else None
        b = np.empty((n_sets, K, 256)) if fetch else None
        pi = np.empty((n_sets, K)) if fetch else None
        hidden = np.empty((K, 3), dtype=np.int32)
        self._ck(self._lib.itr_build_model(
            self._ctx, n_sets, L.as_ptr(params, ctypes.c_double), int(n_int_AB), int(n_int_ABC),
            L.as_ptr(cab, ctypes.c_double), L.as_ptr(cabc, ctypes.c_double),
            L.as_ptr(a, ctypes.c_double), L.as_ptr(b, ctypes.c_double), L.as_ptr(pi, ctypes.c_double),
            L.as_ptr(hidden, ctypes.c_int32)))
        self.K, self.n_sets = K, n_sets
        return a, b, pi, hidden

    # -- recursions --------------------------------------------------------------
    def loglik(self, per_block=False):
        """Summed forward log-likelihood per parameter set (and per block).  In deferred
        mode (:meth:`set_async`) the returned arrays are filled at :meth:`sync`; the engine
        keeps them alive until then, so dropping the result cannot leave the library with a
        dangling pointer."""
        tot = np.empty(self.n_sets)
        pb = np.empty((self.n_sets, self.n_blocks)) if per_block else None
        self._pending = [tot, pb]       # (a second deferred loglik completes the first one inside the library)
        self._ck(self._lib.itr_loglik(self._ctx, L.as_ptr(tot, ctypes.c_double),
                                      L.as_ptr(pb, ctypes.c_double)))
        return (tot, pb) if per_block else tot

    def viterbi(self, log_a, log_E, omega0, out=None, fetch=True):
        """Returns the concatenated uint8 state path (or None if not fetch)."""
        K = self.K
        log_a, log_E, omega0 = L.c_f64(log_a), L.c_f64(log_E), L.c_f64(omega0)
        if log_a.shape != (K, K) or log_E.shape != (K, 625) or omega0.shape != (self.n_blocks, K):
            raise ValueError("viterbi table shapes do not match the installed model/blocks")
        if fetch and out is None:
            out = np.empty(self.n_columns, dtype=np.uint8)
        self._ck(self._lib.itr_viterbi(self._ctx, L.as_ptr(log_a, ctypes.c_double),
                                       L.as_ptr(log_E, ctypes.c_double),
                                       L.as_ptr(omega0, ctypes.c_double),
                                       L.as_ptr(out, ctypes.c_uint8) if fetch else None))
        return out if fetch else None

    def posterior(self, out=None, fetch=True):
        """Returns the (sum T, K) posterior matrix (or None if not fetch)."""
        if fetch and out is None:
            out = np.empty((self.n_columns, self.K))
        self._ck(self._lib.itr_posterior(self._ctx, L.as_ptr(out, ctypes.c_double) if fetch else None))
        return out if fetch else None

    def posterior_block(self, i, out=None):
        """Posterior matrix (T_i, K) of block ``i`` from the result kept on the device."""
        off = self._offsets
        c0, n = int(off[i]), int(off[i + 1] - off[i])
        if out is None:
            out = np.empty((n, self.K))
        self._ck(self._lib.itr_posterior_fetch_range(self._ctx, c0, n, L.as_ptr(out, ctypes.c_double)))
        return out

    def write_posterior_csv(self, path, positions=None, n_threads=0, block_ids=None, header=True):
        """``{prefix}.posterior.csv`` (workflow_posterior.py:697-716) written by the native
        writer straight from the posterior kept on the device; ``positions`` = one int64
        per column (reference coordinates) or None for 0..T-1 within each block.
        ``block_ids`` (one per loaded block) replaces the printed block index and
        ``header=False`` leaves the header out: the part file of one rank of a sharded
        run.  Returns the bytes written per block."""
        if positions is not None:
            positions = np.ascontiguousarray(positions, dtype=np.int64)
            if positions.shape != (self.n_columns,):
                raise ValueError("positions must hold one entry per alignment column")
        if block_ids is not None:
            block_ids = np.ascontiguousarray(block_ids, dtype=np.int64)
            if block_ids.shape != (self.n_blocks,):
                raise ValueError("block_ids must hold one entry per loaded block")
        nbytes = np.zeros(self.n_blocks, dtype=np.int64)
        self._ck(self._lib.itr_posterior_write_csv_ex(
            self._ctx, os.fsencode(path), L.as_ptr(positions, ctypes.c_int64), L.as_ptr(block_ids, ctypes.c_int64),
            1 if header else 0, L.as_ptr(nbytes, ctypes.c_int64), int(n_threads)))
        return nbytes

    def posterior_stream(self, ring, slot_cols, n_slots, sink=None):
        """Posterior decoding streamed through a bounded host ring (itr_posterior_stream):
        ``ring`` is a C-contiguous float64 array of at least n_slots * slot_cols * K values
        (page-locked for full PCIe speed); ``sink(col0, rows)`` — optional — receives every
        piece as an (n, K) view into the ring, in ascending column order, and must be done
        with it when it returns.  The full result also stays on the device."""
        ring = np.asarray(ring)
        if ring.dtype != np.float64 or not ring.flags.c_contiguous or ring.size < n_slots * slot_cols * self.K:
            raise ValueError("ring must be a C-contiguous float64 array of n_slots * slot_cols * K values")
        K = self.K
        err = []
        cb = None
        if sink is not None:
            def _cb(_user, col0, n, rows):
                try:
                    sink(int(col0), np.ctypeslib.as_array(rows, shape=(int(n), K)))
                    return 0
                except BaseException as e:      # noqa: BLE001 - must not unwind through the C frame
                    err.append(e)
                    return 1
            cb = L.ROWS_SINK(_cb)
        rc = self._lib.itr_posterior_stream(self._ctx, L.as_ptr(ring, ctypes.c_double), int(slot_cols), int(n_slots),
                                            ctypes.cast(cb, ctypes.c_void_p) if cb is not None else None, None)
        if err:
            raise err[0]
        self._ck(rc)

    def viterbi_block(self, i, out=None):
        """uint8 state path of block ``i`` from the result kept on the device."""
        off = self._offsets
        c0, n = int(off[i]), int(off[i + 1] - off[i])
        if out is None:
            out = np.empty(n, dtype=np.uint8)
        self._ck(self._lib.itr_viterbi_fetch_range(self._ctx, c0, n, L.as_ptr(out, ctypes.c_uint8)))
        return out

    def split(self, flat):
        """Concatenated per-column result -> list of per-block views."""
        off = self._offsets
        return [flat[off[i]:off[i + 1]] for i in range(len(off) - 1)]

    # -- overlap ------------------------------------------------------------------
    def set_async(self, on=True):
        """With ``on``, loglik / viterbi / posterior only enqueue their work (each on its
        own CUDA stream, so independent recursions overlap on the device); host results
        are valid after :meth:`sync`."""
        self._ck(self._lib.itr_set_async(self._ctx, 1 if on else 0))

    def sync(self):
        self._ck(self._lib.itr_sync(self._ctx))
        self._pending = []

    # -- introspection -----------------------------------------------------------
    def phase_ms(self, name):
        return float(self._lib.itr_phase_ms(self._ctx, L.PHASES[name]))

    @property
    def launch_count(self):
        return int(self._lib.itr_launch_count(self._ctx))

    @property
    def lockstep_launch_count(self):
        """Launches of the lock-step tensor-core sweeps (32 < K <= 96, many blocks)."""
        return int(self._lib.itr_lockstep_launch_count(self._ctx))

    def device_info(self):
        name = ctypes.create_string_buffer(256)
        sm, ma, mi = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._ck(self._lib.itr_device_info(self._ctx, name, 256, ctypes.byref(sm),
                                           ctypes.byref(ma), ctypes.byref(mi)))
        return {"name": name.value.decode(), "sm_count": sm.value, "cc": (ma.value, mi.value)}
