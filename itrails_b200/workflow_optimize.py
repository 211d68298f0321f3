"""Console entry point itrails-optimize (reference: workflow_optimize.py main)."""
from .workflows import optimize_main as main  # noqa: F401

if __name__ == "__main__":
    main()
