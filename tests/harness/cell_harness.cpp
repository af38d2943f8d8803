// cell_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Host build of the PRODUCT's cell-search header (libcpab_b200/csrc/cpab_cell.cuh) so that the
// exact-arithmetic claims in that header can be swept against the oracle on the CPU, with no GPU.
// Built by tests/test_cell_host.py with `g++ -O2 -ffp-contract=off` (fmaf() is the correctly
// rounded libm/hardware FMA, like the device FFMA).  Never loaded by the product.
#include "../../libcpab_b200/csrc/cpab_cell.cuh"

extern "C" {

void harness_findcellidx_f32(int ndim, const int* nc, const float* pts, long nP, int* out)
{
    const cpab::Geom g = cpab::make_geom(ndim, nc);
    for (long i = 0; i < nP; ++i) {
        float p[3] = {0.f, 0.f, 0.f};
        for (int j = 0; j < ndim; ++j) p[j] = pts[i + (long)j * nP];
        out[i] = ndim == 1 ? cpab::find_cell<1>(p, g)
               : ndim == 2 ? cpab::find_cell<2>(p, g) : cpab::find_cell<3>(p, g);
    }
}

void harness_findcellidx_f64(int ndim, const int* nc, const double* pts, long nP, int* out)
{
    const cpab::Geom g = cpab::make_geom(ndim, nc);
    for (long i = 0; i < nP; ++i) {
        double p[3] = {0., 0., 0.};
        for (int j = 0; j < ndim; ++j) p[j] = pts[i + (long)j * nP];
        out[i] = ndim == 1 ? cpab::find_cell<1>(p, g)
               : ndim == 2 ? cpab::find_cell<2>(p, g) : cpab::find_cell<3>(p, g);
    }
}

}  // extern "C"
