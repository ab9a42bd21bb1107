"""TEST INFRASTRUCTURE -- reads the plotfile layout warpx_b200/diagnostics.py writes (the AMReX plotfile format the
reference produces through FlushFormatPlotfile) and computes the reference's regression checksums from it the way
Regression/Checksum/checksum.py:62-217 does with yt: per field sum(|Q|) over the covering grid of the level, per
species sum(|q|) of every particle quantity.  Stands in for yt, which is not in this image."""
import os
import re

import numpy as np


def _box(text):
    m = re.match(r"\(\((-?\d+),(-?\d+),(-?\d+)\) \((-?\d+),(-?\d+),(-?\d+)\) \((\d+),(\d+),(\d+)\)\)", text.strip())
    v = [int(x) for x in m.groups()]
    return tuple(v[0:3]), tuple(v[3:6])


def read_fields(root):
    """{name: array [nz, ny, nx] over the whole domain}, plus the header dictionary."""
    with open(os.path.join(root, "Header")) as f:
        lines = [ln.rstrip("\n") for ln in f]
    assert lines[0] == "HyperCLaw-V1.1"
    ncomp = int(lines[1])
    names = lines[2:2 + ncomp]
    p = 2 + ncomp
    assert int(lines[p]) == 3
    time = float(lines[p + 1])
    assert int(lines[p + 2]) == 0                      # finest level
    prob_lo = [float(v) for v in lines[p + 3].split()]
    prob_hi = [float(v) for v in lines[p + 4].split()]
    dom_lo, dom_hi = _box(lines[p + 6])
    step = int(lines[p + 7].split()[0])
    dx = [float(v) for v in lines[p + 8].split()]
    level, ngrids, _ = lines[p + 11].split()
    ngrids = int(ngrids)
    assert lines[p + 13 + 3 * ngrids] == "Level_0/Cell"
    n = [dom_hi[d] - dom_lo[d] + 1 for d in range(3)]
    for d in range(3):
        assert abs((prob_hi[d] - prob_lo[d]) / n[d] - dx[d]) <= 1e-12 * dx[d]
    lev = os.path.join(root, "Level_0")
    with open(os.path.join(lev, "Cell_H")) as f:
        h = [ln.rstrip("\n") for ln in f]
    assert int(h[2]) == ncomp
    nb = int(h[4].split()[0][1:])
    boxes = [_box(h[5 + b]) for b in range(nb)]
    assert h[5 + nb] == ")" and int(h[6 + nb]) == nb
    fabs = []
    for b in range(nb):
        _, fname, off = h[7 + nb + b].split()
        fabs.append((fname, int(off)))
    out = {name: np.zeros((n[2], n[1], n[0])) for name in names}
    covered = np.zeros((n[2], n[1], n[0]), dtype=bool)
    for (lo, hi), (fname, off) in zip(boxes, fabs):
        with open(os.path.join(lev, fname), "rb") as f:
            f.seek(off)
            head = f.readline().decode()
            assert head.startswith("FAB ((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))")
            blo, bhi = _box(head[head.index(")))") + 3:head.rindex(" ")])
            assert (blo, bhi) == (lo, hi) and int(head.split()[-1]) == ncomp
            m = [hi[d] - lo[d] + 1 for d in range(3)]
            for name in names:
                a = np.frombuffer(f.read(8 * m[0] * m[1] * m[2]), dtype="<f8").reshape(m[2], m[1], m[0])
                out[name][lo[2] - dom_lo[2]:hi[2] - dom_lo[2] + 1, lo[1] - dom_lo[1]:hi[1] - dom_lo[1] + 1,
                          lo[0] - dom_lo[0]:hi[0] - dom_lo[0] + 1] = a
        covered[lo[2] - dom_lo[2]:hi[2] - dom_lo[2] + 1, lo[1] - dom_lo[1]:hi[1] - dom_lo[1] + 1,
                lo[0] - dom_lo[0]:hi[0] - dom_lo[0] + 1] = True
    assert covered.all()
    return out, dict(time=time, step=step, prob_lo=prob_lo, prob_hi=prob_hi, n_cell=n, dx=dx, names=names, ngrids=nb)


def read_species(root, name):
    """{position_x, ..., weight, momentum_x, ...: array over all particles of the species}"""
    with open(os.path.join(root, name, "Header")) as f:
        h = [ln.strip() for ln in f]
    assert h[0] == "Version_Two_Dot_Zero_double" and int(h[1]) == 3
    nreal = int(h[2])
    rnames = h[3:3 + nreal]
    p = 3 + nreal
    assert int(h[p]) == 0 and int(h[p + 1]) == 0            # no integer components, not a checkpoint
    ntot, ngrids = int(h[p + 2]), int(h[p + 5])
    cols = ["position_x", "position_y", "position_z"] + rnames
    parts = []
    for g in range(ngrids):
        which, count, where = (int(v) for v in h[p + 6 + g].split())
        with open(os.path.join(root, name, "Level_0", "DATA_%05d" % which), "rb") as f:
            f.seek(where)
            parts.append(np.frombuffer(f.read(8 * count * len(cols)), dtype="<f8").reshape(count, len(cols)))
    rec = np.concatenate(parts) if parts else np.zeros((0, len(cols)))
    assert len(rec) == ntot
    return {c: rec[:, k] for k, c in enumerate(cols)}


def checksums(root, species=()):
    """The dictionary Regression/Checksum/checksum.py builds from a plotfile: {"lev=0": {field: sum|Q|},
    species: {"particle_" + quantity: sum|q|}}."""
    fields, _ = read_fields(root)
    data = {"lev=0": {k: float(np.sum(np.abs(v))) for k, v in fields.items()}}
    for s in species:
        P = read_species(root, s)
        data[s] = {"particle_" + k: float(np.sum(np.abs(v))) for k, v in P.items()}
    return data
