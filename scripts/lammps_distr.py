#!/usr/bin/env python
"""Entry point with the reference's script name and flags (scripts/lammps_distr.py upstream); the RDF histogram runs
in neuralmelting_b200's CUDA kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralmelting_b200.distr import main  # noqa: E402

if __name__ == "__main__":
    main(sys.argv[1:])
