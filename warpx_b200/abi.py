"""ctypes mirror of include/pic_b200.h (the C ABI).  Pure declarations: no compute, no CUDA."""
import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class pic_fab(C.Structure):
    _fields_ = [("p", C.c_void_p), ("lo", C.c_int * 3), ("hi", C.c_int * 3),
                ("ng", C.c_int * 3), ("stag", C.c_int * 3)]

    @property
    def shape(self):  # numpy (z, y, x) shape of the allocated region
        return tuple(self.hi[d] - self.lo[d] + 1 for d in (2, 1, 0))

    @property
    def size(self):
        return int(np.prod(self.shape))

    def valid_slices(self):
        """numpy slices (z, y, x) of the valid region inside the allocated array."""
        return tuple(slice(self.ng[d], self.hi[d] - self.lo[d] + 1 - self.ng[d]) for d in (2, 1, 0))


class pic_soa(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p), ("w", C.c_void_p),
                ("ux", C.c_void_p), ("uy", C.c_void_p), ("uz", C.c_void_p),
                ("idcpu", C.c_void_p), ("np", C.c_long)]


class pic_stencil(C.Structure):
    _fields_ = [("algo", C.c_int), ("cx", C.c_double * 5), ("cy", C.c_double * 5),
                ("cz", C.c_double * 5)]


class pic_bins(C.Structure):
    _fields_ = [("cell_start", C.c_void_p), ("box_lo", C.c_int * 3), ("box_hi", C.c_int * 3),
                ("tile", C.c_int * 3), ("np_binned", C.c_long)]


class pic_escape_list(C.Structure):
    _fields_ = [("idx", C.c_void_p), ("count", C.c_void_p), ("capacity", C.c_int),
                ("lo", C.c_double * 3), ("hi", C.c_double * 3)]


class pic_geom(C.Structure):
    _fields_ = [("n_cell", C.c_int * 3), ("prob_lo", C.c_double * 3), ("prob_hi", C.c_double * 3),
                ("periodic", C.c_int * 3)]


SOLVER_YEE, SOLVER_CKC = 0, 1
PIC_ERR_ABORT, PIC_ERR_RETURN = 0, 1
PUSHER_BORIS, PUSHER_VAY, PUSHER_HC = 0, 1, 2

# Yee staggering of WarpX (Source/WarpX.cpp:2117-2125): 1 = nodal.  Order Ex Ey Ez Bx By Bz jx jy jz.
YEE_STAG = ((0, 1, 1), (1, 0, 1), (1, 1, 0),
            (1, 0, 0), (0, 1, 0), (0, 0, 1),
            (0, 1, 1), (1, 0, 1), (1, 1, 0))
COMP_NAMES = ("Ex", "Ey", "Ez", "Bx", "By", "Bz", "jx", "jy", "jz")


def make_fab(ptr, box_lo, box_hi, ng, stag):
    """Descriptor of one component on the cell box [box_lo, box_hi] (inclusive)."""
    f = pic_fab()
    f.p = ptr
    for d in range(3):
        f.stag[d] = stag[d]
        f.ng[d] = ng[d]
        f.lo[d] = box_lo[d] - ng[d]
        f.hi[d] = box_hi[d] + stag[d] + ng[d]
    return f


def make_geom(n_cell, prob_lo, prob_hi, periodic=(1, 1, 1)):
    g = pic_geom()
    for d in range(3):
        g.n_cell[d] = int(n_cell[d])
        g.prob_lo[d] = float(prob_lo[d])
        g.prob_hi[d] = float(prob_hi[d])
        g.periodic[d] = int(periodic[d])
    return g


def int3(v):
    return (C.c_int * 3)(*[int(a) for a in v])


def dbl3(v):
    return (C.c_double * 3)(*[float(a) for a in v])
