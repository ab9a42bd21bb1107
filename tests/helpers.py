"""Shared helpers of the parity tests."""
import numpy as np

from warpx_b200 import abi, workloads  # noqa: F401


def rel_linf(a, b):
    """max|a-b| / max|b| (fields with heavy cancellation are compared against their scale)."""
    s = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / s) if s > 0 else float(np.max(np.abs(a - b)))


def random_fields(orc, box_lo, box_hi, ng, seed, comps=range(6), scale=None):
    """HostFab list Ex..jz with random values everywhere (guards included)."""
    rng = np.random.default_rng(seed)
    out = []
    for c in comps:
        f = orc.HostFab(box_lo, box_hi, ng, abi.YEE_STAG[c])
        s = 1.0 if scale is None else scale[c]
        f.a[...] = s * rng.standard_normal(f.a.shape)
        out.append(f)
    return out


def make_species_in_box(n_cell, prob_lo, prob_hi, ppc, u_th, seed, perturbation=0.0):
    wl = workloads.uniform_plasma_3d(n_cell=n_cell, ppc=ppc, u_th=u_th, seed=seed,
                                     lx=prob_hi[0] - prob_lo[0], perturbation=perturbation)
    return wl


def lower_corner(prob_lo, dx, box_lo, ng):
    lo = [box_lo[d] - ng[d] for d in range(3)]
    return [prob_lo[d] + dx[d] * lo[d] for d in range(3)], lo
