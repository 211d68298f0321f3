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
# MAF parsing
# ---------------------------------------------------------------------------
_BYTE_TO_DIGIT = np.full(256, 255, dtype=np.uint8)
for _i, _ch in enumerate(NUC):
    _BYTE_TO_DIGIT[ord(_ch)] = _i
    _BYTE_TO_DIGIT[ord(_ch.lower())] = _i
_BYTE_TO_DIGIT[ord("N")] = 4
_BYTE_TO_DIGIT[ord("n")] = 4
_BYTE_TO_DIGIT[ord("-")] = 4          # read_data.py:109: gaps become N


def _maf_blocks(file):
    """Yield one list of (src, start, size, strand, srcSize, text) per ``a`` block."""
    rows, in_block = [], False
    with open(file, "rb") as fh:
        for raw in fh:
            line = raw.strip()
            if not line:
                if in_block:
                    yield rows
                rows, in_block = [], False
                continue
            tag = line[:1]
            if tag == b"#":
                continue
            if tag == b"a":
                if in_block:
                    yield rows
                rows, in_block = [], True
            elif tag == b"s" and in_block:
                f = line.split()
                if len(f) != 7:
                    raise ValueError(f"malformed MAF sequence line: {line[:60]!r}")
                strand = 1 if f[4] == b"+" else -1
                rows.append((f[1].decode(), int(f[2]), int(f[3]), strand, int(f[5]), f[6]))
        if in_block:
            yield rows


def maf_parser(file, sp_lst):
    """MAF file -> list of int64 arrays of observed-symbol indices, one per alignment
    block that contains all four species of ``sp_lst`` (read_data.py:94-117).
    Species = text before the first ``.`` of the source name; gaps count as ``N``;
    a character outside A,C,G,T,N,- raises ValueError (as ``list.index`` does)."""
    if _CODE_TO_INDEX is None:
        _build_tables()
    total = []
    for rows in _maf_blocks(file):
        dct = {}
        length = None
        for src, _start, _size, _strand, _srcsize, text in rows:
            if length is None:
                length = len(text)
            elif len(text) != length:
                raise ValueError("sequences in a MAF block must have equal length")
            sp = src.split(".")[0]
            if sp in sp_lst:
                dct[sp] = text
        if len(dct) == 4:
            code = np.zeros(length, dtype=np.int64)
            for sp in sp_lst:
                d = _BYTE_TO_DIGIT[np.frombuffer(dct[sp], dtype=np.uint8)]
                if d.size and d.max() == 255:
                    bad = chr(dct[sp][int(np.argmax(d == 255))])
                    raise ValueError(f"'{bad}' is not a valid nucleotide in a MAF column")
                code = code * 5 + d
            total.append(_CODE_TO_INDEX[code])
    return total


def parse_coordinates(file, sp_lst, ref):
    """Per kept block, the reference-species coordinate of every column, -9 at
    gaps / when the reference species is absent (read_data.py:146-220)."""
    tot = []
    for rows in _maf_blocks(file):
        acc, length = 0, 0
        hit = None
        for src, start, _size, strand, srcsize, text in rows:
            sp = src.split(".")[0]
            if sp in sp_lst:
                length = len(text)
                acc += 1
            if sp == ref:
                hit = (start, strand, srcsize, text)
        if acc != 4:
            continue
        if hit is None:
            tot.append([-9] * length)
            continue
        start, strand, srcsize, text = hit
        st = start if strand == 1 else srcsize - start
        present = np.frombuffer(text, dtype=np.uint8) != ord("-")
        coords = np.full(len(text), -9, dtype=np.int64)
        coords[present] = st + strand * np.arange(int(present.sum()))
        tot.append(coords.tolist())
    return tot
