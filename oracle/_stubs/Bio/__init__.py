"""Stub of the Biopython namespace so that the reference's read_data.py imports
without Biopython installed (test infrastructure only; see oracle/README.md)."""


class _AlignIO:
    @staticmethod
    def parse(*_a, **_k):
        raise RuntimeError("Biopython is not installed; stub AlignIO cannot parse")


AlignIO = _AlignIO()
