"""Shared helpers of the parity tests (test infrastructure)."""
import json
import os

import numpy as np

from oracle import coeff, cpu, mesh


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def oracle_backend(hier, eta_q, averaging, ps, bc):
    """C matrix-free oracle backend on the oracle's own mesh restatement."""
    op = {"none": "none", "harmonic average": "harmonic", "harmonic average only viscosity": "harmonic",
          "arithmetic average": "arithmetic", "geometric average": "geometric"}[averaging]
    active, levels, emin, emax = coeff.viscosity_tables(hier, eta_q, op)
    return cpu.backend_from_hierarchy(hier, active, levels, ps, bc), active, levels, emin, emax


def seeded(n, seed):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, size=n)


_GOLDEN = None


def golden(test):
    """Pinned values of the reference's own test `test` (tests/golden/reference_screen_output.json, written by
    tools/extract_goldens.py from /root/reference/tests/<test>/screen-output)."""
    global _GOLDEN
    if _GOLDEN is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_screen_output.json")) as f:
            _GOLDEN = json.load(f)
    return _GOLDEN[test]
