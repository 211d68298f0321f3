"""Data ingest — mirrors the reference's read_data.py (same function names/results).

    get_obs_state_dct   read_data.py:6-24
    get_idx_state       read_data.py:46-67
    maf_parser          read_data.py:94-117
    parse_coordinates   read_data.py:146-220

The reference parses MAF through Biopython's ``AlignIO.parse(file, "maf")`` and
resolves every alignment column with an O(625) ``list.index``.  Here the MAF text is
parsed directly (format: blocks opened by an ``a`` line, rows
``s src start size strand srcSize text``, closed by a blank line) and columns are
converted with two 256/625-entry lookup tables.
"""
from __future__ import annotations

import numpy as np

NUC = "ACTG"  # the reference's nucleotide order (read_data.py:13), NOT "ACGT"

# ---------------------------------------------------------------------------
# alphabet
# ---------------------------------------------------------------------------
_NAMES = None
_CODE_TO_INDEX = None      # base-5 code (A,C,T,G,N digits) -> reference symbol index
_ORDER = None


def _build_tables():
    global _NAMES, _CODE_TO_INDEX
    ext = NUC + "N"
    plain, rest = [], []
    code_to_index = np.full(625, -1, dtype=np.int64)
    for a in range(5):
        for b in range(5):
            for c in range(5):
                for d in range(5):
                    s = ext[a] + ext[b] + ext[c] + ext[d]
                    code = ((a * 5 + b) * 5 + c) * 5 + d
                    if max(a, b, c, d) < 4:
                        code_to_index[code] = 64 * a + 16 * b + 4 * c + d
                        plain.append(s)
                    else:
                        code_to_index[code] = 256 + len(rest)
                        rest.append(s)
    _NAMES = plain + rest
    _CODE_TO_INDEX = code_to_index


def get_obs_state_dct():
    """The 625 observed four-species column strings in index order
    (read_data.py:6-24): 256 N-free strings (index 64a+16b+4c+d with A,C,T,G=0..3)
    followed by the 369 strings containing ``N``."""
    if _NAMES is None:
        _build_tables()
    return list(_NAMES)


def get_idx_state(state):
    """Indices of the N-free symbols that observed symbol ``state`` marginalises
    over, ascending (read_data.py:46-67)."""
    global _ORDER
    if _ORDER is None:
        names = get_obs_state_dct()
        _ORDER = []
        for s in names:
            idx = [0]
            for ch in s:
                if ch == "N":
                    idx = [4 * i + k for i in idx for k in range(4)]
                else:
                    k = NUC.index(ch)
                    idx = [4 * i + k for i in idx]
            _ORDER.append(np.array(idx, dtype=np.int64))
    return _ORDER[int(state)].copy()


def order_lists():
    """``[get_idx_state(i) for i in range(625)]`` (optimizer.py:54) — cached."""
    get_idx_state(0)
    return _ORDER


# ---------------------------------------------------------------------------
# MAF parsing (C++ reader in libitrails_b200: csrc/maf_reader.cpp)
# ---------------------------------------------------------------------------
def read_maf(file, sp_lst, ref=None, n_threads=0):
    """Parse a MAF file with the native reader.  Returns ``(sym, off, coord, coord_off)``:
    uint16 symbols of all kept blocks back to back with their int64 offsets (the arguments
    of ``Engine.load_packed``) and, when ``ref`` is given, int64 coordinates with their own
    offsets (else ``None, None``)."""
    import ctypes

    from . import _lib as L
    if len(sp_lst) != 4:
        raise ValueError("sp_lst must name four species")
    lib = L.load()
    names = (ctypes.c_char_p * 4)(*[s.encode() for s in sp_lst])
    handle = ctypes.c_void_p()
    err = ctypes.create_string_buffer(512)
    rc = lib.itr_maf_read(str(file).encode(), names, ref.encode() if ref is not None else None,
                          int(n_threads), ctypes.byref(handle), err, 512)
    if rc != 0:
        msg = err.value.decode("utf-8", "replace")
        if "cannot open" in msg or "cannot stat" in msg:
            raise FileNotFoundError(msg)
        raise ValueError(msg)
    try:
        nb, nc = lib.itr_maf_num_blocks(handle), lib.itr_maf_num_columns(handle)
        off = np.ctypeslib.as_array(lib.itr_maf_offsets(handle), shape=(nb + 1,)).copy()
        sym = np.empty(nc, np.uint16)                 # filled (and first touched) by the reader's threads
        coord = coord_off = None
        if ref is not None:
            ncb = lib.itr_maf_num_coord_blocks(handle)
            coord_off = np.ctypeslib.as_array(lib.itr_maf_coord_offsets(handle), shape=(ncb + 1,)).copy()
            coord = np.empty(int(coord_off[-1]), np.int64)
        rc = lib.itr_maf_export(handle, L.as_ptr(sym, ctypes.c_uint16) if nc else None,
                                L.as_ptr(coord, ctypes.c_int64) if coord is not None and coord.size else None,
                                int(n_threads))
        if rc != 0:
            raise RuntimeError(f"itr_maf_export failed ({rc})")
    finally:
        lib.itr_maf_free(handle)
    return sym, off, coord, coord_off


def maf_parser(file, sp_lst):
    """MAF file -> list of int64 arrays of observed-symbol indices, one per alignment
    block that contains all four species of ``sp_lst`` (read_data.py:94-117).
    Species = text before the first ``.`` of the source name; gaps count as ``N``;
    a character outside A,C,G,T,N,- raises ValueError (as ``list.index`` does)."""
    sym, off, _, _ = read_maf(file, sp_lst)
    return [sym[off[i]:off[i + 1]].astype(np.int64) for i in range(len(off) - 1)]


def parse_coordinates(file, sp_lst, ref):
    """Per kept block, the reference-species coordinate of every column, -9 at
    gaps / when the reference species is absent (read_data.py:146-220)."""
    _, _, coord, coff = read_maf(file, sp_lst, ref=ref)
    return [coord[coff[i]:coff[i + 1]].tolist() for i in range(len(coff) - 1)]
