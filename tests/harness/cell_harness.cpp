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

// Development aid: how often does the lean 3-D fast path hand a point to the complete search along
// forward trajectories?  Integrates nsteps of p <- Trels[cell] [p;1] (float, unfused, the oracle's
// order) and counts, per step, the points for which find_cell_3d_lean returns true; `why` gets
// three counters: out2-style (coordinates outside), estimate too large, inside the plane band.
extern "C" void harness_lean3d_rare_along_flow(const int* nc, const float* pts, long nP, const float* trels, int nsteps,
                                               long* rare_per_step, long* why)
{
    const cpab::Geom g = cpab::make_geom(3, nc);
    for (int s = 0; s < nsteps; ++s) rare_per_step[s] = 0;
    why[0] = why[1] = why[2] = 0;
    for (long i = 0; i < nP; ++i) {
        float p[3] = {pts[i], pts[i + nP], pts[i + 2 * nP]};
        for (int s = 0; s < nsteps; ++s) {
            int cell; float dist;
            const bool rare = cpab::find_cell_3d_lean<true>(p[0], p[1], p[2], g, 12582912.0f, cell, dist);
            if (rare) {
                rare_per_step[s]++;
                const float dx = fabsf(p[0] - 0.5f), dy = fabsf(p[1] - 0.5f), dz = fabsf(p[2] - 0.5f);
                if (fminf(fminf(dx, dy), dz) > 0.5f) why[0]++;
                cpab::Geom g2 = g; g2.band3 = 0.0f;
                int c2; float d2;
                if (!cpab::find_cell_3d_lean<true>(p[0], p[1], p[2], g2, 12582912.0f, c2, d2)) why[2]++;   // only the band
                else why[1]++;                                                                             // something else
            }
            const int c = cpab::find_cell<3>(p, g);
            const float* T = trels + (long)c * 12;
            float q[3];
            for (int r = 0; r < 3; ++r) {
                float acc = T[4 * r] * p[0];
                acc = acc + T[4 * r + 1] * p[1];
                acc = acc + T[4 * r + 2] * p[2];
                q[r] = acc + T[4 * r + 3];
            }
            p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
        }
    }
}

extern "C" void harness_lean3d_flags(const int* nc, const float* pts, long nP, unsigned char* rare)
{
    const cpab::Geom g = cpab::make_geom(3, nc);
    for (long i = 0; i < nP; ++i) {
        int cell; float dist;
        rare[i] = cpab::find_cell_3d_lean<true>(pts[i], pts[i + nP], pts[i + 2 * nP], g, 12582912.0f, cell, dist) ? 1 : 0;
    }
}
