// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// extern "C" doors onto the reference's own C++ core so that ctypes can reach it without
// guessing mangled names.  It is compiled TOGETHER WITH the unmodified reference source
// /root/reference/libcpab/core/cpab_ops.cpp (never copied into this repo) into
// oracle/_ref/libcpab_ref.so by oracle/Makefile.  The declarations below are the reference's
// public prototypes, libcpab/core/cpab_ops.h:4-20.
int findcellidx(int ndim, const float* p, const int* nc);
void cpab_forward_op(float* newpoints, const float* points, const float* trels,
                     const int* nstepsolver, const int* nc, const int ndim, const int nP,
                     const int batch_size, const int broadcast);
void cpab_backward_op(float* grad, const float* points, const float* As, const float* Bs,
                      const int* nstepsolver, const int* nc, const int n_theta, const int d,
                      const int ndim, const int nP, const int nC, const int broadcast);

extern "C" {

int cpab_ref_abi(void) { return 1; }

// planar [ndim,nP] points -> cell index per point, through the reference's findcellidx
void cpab_ref_findcellidx(int ndim, const int* nc, const float* points, long nP, int* out) {
    for (long i = 0; i < nP; ++i) {
        float pt[3] = {0.f, 0.f, 0.f};
        for (int j = 0; j < ndim; ++j) pt[j] = points[i + (long)j * nP];
        out[i] = findcellidx(ndim, pt, nc);
    }
}

void cpab_ref_forward(float* newpoints, const float* points, const float* trels, int nsteps,
                      const int* nc, int ndim, int nP, int n_theta, int broadcast) {
    cpab_forward_op(newpoints, points, trels, &nsteps, nc, ndim, nP, n_theta, broadcast);
}

// grad must be zero-initialised by the caller (the reference allocates it with torch::zeros,
// libcpab/pytorch/transformer.cpp:50, and reads q back from it every step).
void cpab_ref_backward(float* grad, const float* points, const float* As, const float* Bs,
                       int nsteps, const int* nc, int n_theta, int d, int ndim, int nP, int nC,
                       int broadcast) {
    cpab_backward_op(grad, points, As, Bs, &nsteps, nc, n_theta, d, ndim, nP, nC, broadcast);
}

}  // extern "C"
