// TEST INFRASTRUCTURE (host harness): exposes the __host__ __device__ shape-factor templates of
// warpx_b200/csrc/pic_common.cuh to the CPU test-suite.  Never part of the product library.
#include "pic_common.cuh"

extern "C" int pic_host_shape(int order, double x, double* s) {
    switch (order) {
        case 0: return pic::shape_factor<0>(s, x);
        case 1: return pic::shape_factor<1>(s, x);
        case 2: return pic::shape_factor<2>(s, x);
        case 3: return pic::shape_factor<3>(s, x);
        case 4: return pic::shape_factor<4>(s, x);
    }
    return -999;
}
extern "C" int pic_host_shifted_shape(int order, double x_old, int i_new, double* s /* order+3, pre-zeroed */) {
    switch (order) {
        case 1: return pic::shifted_shape_factor<1>(s, x_old, i_new);
        case 2: return pic::shifted_shape_factor<2>(s, x_old, i_new);
        case 3: return pic::shifted_shape_factor<3>(s, x_old, i_new);
        case 4: return pic::shifted_shape_factor<4>(s, x_old, i_new);
    }
    return -999;
}
