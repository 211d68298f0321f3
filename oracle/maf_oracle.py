"""CPU restatement (NumPy) of the reference's MAF ingest.

TEST INFRASTRUCTURE ONLY.  Nothing under ``itrails_b200/`` may import this module.
Restates read_data.py:94-117 (``maf_parser``) and read_data.py:146-220
(``parse_coordinates``) of the reference without Biopython (absent from this image), from
the MAF format itself: blocks opened by an ``a`` line, rows
``s src start size strand srcSize text``, closed by a blank line.  Parity unpinned by the
reference (it has no tests and Biopython cannot be installed here); pinned by hand-written
MAF text with hand-computed expectations in tests/test_host.py.
"""
import numpy as np

from hmm_oracle import NUC, code_to_index_table

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
    _CODE_TO_INDEX = code_to_index_table()
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
