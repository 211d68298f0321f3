"""Console entry point itrails-posterior (reference: workflow_posterior.py main)."""
from .workflows import posterior_main as main  # noqa: F401

if __name__ == "__main__":
    main()
