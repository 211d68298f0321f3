"""Console entry point itrails-viterbi (reference: workflow_viterbi.py main)."""
from .workflows import viterbi_main as main  # noqa: F401

if __name__ == "__main__":
    main()
