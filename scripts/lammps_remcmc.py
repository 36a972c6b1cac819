#!/usr/bin/env python
"""Entry point with the reference's script name and flags (scripts/lammps_remcmc.py upstream); the work is done by
neuralmelting_b200.remcmc on the GPU. Multi-GPU: torchrun --nproc-per-node N scripts/lammps_remcmc.py ..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralmelting_b200.remcmc import main  # noqa: E402

if __name__ == "__main__":
    main(sys.argv[1:])
