"""CPU restatement (NumPy) of the reference's MAF ingest.

TEST INFRASTRUCTURE ONLY.  Nothing under ``itrails_b200/`` may import this module.
Restates read_data.py:94-117 (``maf_parser``) and read_data.py:146-220
(``parse_coordinates``) of the reference, and — because those go through
``Bio.AlignIO.parse(file, "maf")`` — the block iterator of Biopython 1.84
(``Bio/AlignIO/MafIO.py``, ``MafIterator``; the version the reference pins, absent from
this image and not installable here), restated from its published algorithm:

* outside a block: a line starting with ``a`` opens one (its ``key=value`` words are
  checked: as many words as ``=`` signs, else ``ValueError``); every other line —
  ``##maf``, ``#`` comments, ``track`` lines, blank lines, anything — is skipped;
* inside a block: a line starting with ``s`` is a sequence row and must have exactly 7
  whitespace-separated fields (``ValueError`` otherwise); strand ``+`` is 1, ``-`` is -1,
  anything else 1; a ``.`` in the text means "same letter as the FIRST row of the block";
  lines starting with ``i``, ``e``, ``q``, ``#`` are skipped; a blank line (or the end of
  the file) closes the block; ANY other line — including an ``a`` line, i.e. two blocks
  without a blank line between them, or an indented row — is a ``ValueError``;
* closing a block builds a ``MultipleSeqAlignment``: rows of unequal length are a
  ``ValueError``; a block without rows is an empty alignment (skipped by both callers);
* the file is read in text mode with universal newlines, so CRLF files parse alike.

Parity is UNPINNED by the reference (it has no tests; Biopython cannot run here): the pins
are the hand-computed cases of tests/test_maf_cases.py, one per quirk, each citing the
read_data.py line it exercises.
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
    """Yield one list of (src, start, size, strand, srcSize, text) per ``a`` block
    (MafIterator of Biopython 1.84, see the module docstring)."""
    rows, in_block = [], False

    def close(rows):
        if rows and any(len(r[5]) != len(rows[0][5]) for r in rows):
            raise ValueError("Sequences must all be the same length")      # MultipleSeqAlignment
        return rows

    with open(file, "r", newline=None) as fh:           # universal newlines, like AlignIO
        for line in fh:
            if in_block:
                if line.startswith("s"):
                    f = line.strip().split()
                    if len(f) != 7:
                        raise ValueError("Error parsing alignment - 's' line must have 7 fields")
                    strand = -1 if f[4] == "-" else 1
                    text = f[6]
                    if "." in text:
                        if not rows:
                            raise ValueError("Found dot/period in first sequence of alignment")
                        text = "".join(r if c == "." else c for c, r in zip(text, rows[0][5]))
                    rows.append((f[1], int(f[2]), int(f[3]), strand, int(f[5]), text))
                elif line[:1] in ("i", "e", "q", "#"):
                    pass
                elif not line.strip():
                    yield close(rows)
                    rows, in_block = [], False
                else:
                    raise ValueError(f"Error parsing alignment - unexpected line:\n{line}")
            elif line.startswith("a"):
                words = line.strip().split()[1:]
                if len(words) != line.count("="):
                    raise ValueError("Error parsing alignment - invalid key in 'a' line")
                rows, in_block = [], True
        if in_block:
            yield close(rows)


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
                raw = np.frombuffer(dct[sp].encode("latin-1", "replace"), dtype=np.uint8)
                d = _BYTE_TO_DIGIT[raw]
                if d.size and d.max() == 255:
                    bad = dct[sp][int(np.argmax(d == 255))]
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
        present = np.frombuffer(text.encode("latin-1", "replace"), dtype=np.uint8) != ord("-")
        coords = np.full(len(text), -9, dtype=np.int64)
        coords[present] = st + strand * np.arange(int(present.sum()))
        tot.append(coords.tolist())
    return tot
