"""Toy nonlinear systems used by the reference's own solver checks."""
import numpy as np


def toyF(v):
    return np.array([np.cos(v[1]) - v[0], np.sin(v[0]) * 0.5 - v[1], 0.3 * v[0] - v[2] + 1])


def toy3(v):  # DEALII_SCFT/test_ADM.c:8-18
    x, y, z = v
    return np.array([x * y * z - 12., x * x + y * y - 8., x + y + z - 511.])


def toy2(v):  # 1D_FEM.c:372-379 (myfun): fixed point (-4, 6)
    return np.array([v[0] * 0.5 - 2., v[1] * 0.5 + 3.])


def fp4(v):
    return np.array([np.cos(v[1]), 0.5 * np.sin(v[0]) + 0.1 * v[2], 0.3 * v[0] + 1, 0.2 * v[3] + 0.1 * v[0] * v[1]])
