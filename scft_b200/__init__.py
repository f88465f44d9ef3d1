"""scft_b200 — B200-native SCFT propagator engine (hot path of giantsda/SCFT).

The product is the C-ABI shared library scft_b200/lib/libscft_b200.so (include/scft_b200.h), built
from scft_b200/csrc by `make -C scft_b200/csrc` (or __graft_entry__.build()).  This package is the
thin Python host mirror used by tests/ and bench.py; it fails loudly if the library is missing —
there is no CPU fallback.
"""
from .engine import (Engine, lib, IE_ROWSCALE, IE_CONSISTENT, IRK4_CONSISTENT, QUAD_ROMBERG,  # noqa: F401
                     QUAD_TRAPEZOID, TAU_REF, L_REF, ScftError, launch_count, LIB_PATH, AndersonBatch,
                     spline, refine_mesh, refine_mesh_adaptive, write_solution, write_detailed_solution, read_solution, read_res, Engine2D, nccl_unique_id,
                     PrecondAndersonBatch, padm_batch, free_energy_weights, SweepSolver)
