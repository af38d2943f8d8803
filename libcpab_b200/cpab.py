"""`Cpab` -- the user-facing class, API-compatible with libcpab.cpab.Cpab (libcpab/cpab.py:16-539)
for the one configuration this package implements: backend='pytorch', device='gpu'.

Constructor arguments, method names, argument meaning, shapes and assertion behaviour follow the
reference so that code written against libcpab runs unchanged; the matplotlib visualisation
helpers (cpab.py:349-476) are not part of the hot path and are not provided.
"""
from __future__ import annotations

import os

import numpy as np

from . import functions as _backend
from .tessellation import Tessellation


class params:   # same bag-of-attributes object as libcpab/core/utility.py:20-24
    def __repr__(self):
        return str({k: v for k, v in self.__dict__.items() if not k.startswith("_")})


class Cpab(object):
    def __init__(self, tess_size, backend="pytorch", device="gpu", zero_boundary=True,
                 volume_perservation=False, override=False, basis=None):
        self._check_input(tess_size, backend, device, zero_boundary, volume_perservation, override)
        p = self.params = params()
        p.nc = list(tess_size)
        p.ndim = len(tess_size)
        p.Ashape = [p.ndim, p.ndim + 1]
        p.valid_outside = not zero_boundary
        p.zero_boundary = zero_boundary
        p.volume_perservation = volume_perservation
        p.domain_max = [1 for _ in p.nc]
        p.domain_min = [0 for _ in p.nc]
        p.inc = [(p.domain_max[i] - p.domain_min[i]) / p.nc[i] for i in range(p.ndim)]
        p.nstepsolver = 50
        p.numeric_grad = False
        p.use_slow = False
        p.fast_math = False        # extension: FMA-contracted forward (not bit-exact with the CPU ref)
        p.points_grad = False      # extension: return dL/dpoints (reference returns None)
        p.closed_form = False      # extension: exact hit-time integration instead of nstepsolver steps
        # transform_data as one forward + one backward kernel (identical results, 5 launches instead
        # of 8): None = automatic (1-D, where the fused accesses stay unit-stride, and launch-bound
        # sizes; measured break-even ~8M pairs in 2-D/3-D, profiles/r01_fused_vs_unfused.txt),
        # True / False force it.  LIBCPAB_B200_FUSED=0/1 overrides the default.
        env = os.environ.get("LIBCPAB_B200_FUSED")
        p.fused_transform_data = None if env is None else env != "0"
        p.nC = int({1: 1, 2: 4, 3: 5}[p.ndim] * np.prod(p.nc))
        p.params_pr_cell = p.ndim * (p.ndim + 1)

        self.tesselation = Tessellation(p.nc, p.domain_min, p.domain_max, zero_boundary,
                                        volume_perservation, override=override)
        p.constrain_mat = self.tesselation.L
        # `basis=` lets a caller share the exact B of another implementation (null(L) is unique
        # only up to a rotation of its columns) -- used by the parity tests
        p.basis = self.tesselation.B if basis is None else np.ascontiguousarray(basis, dtype=np.float64)
        p.D, p.d = p.basis.shape
        assert p.D == p.nC * p.params_pr_cell, "basis does not match the tessellation"

        self.backend_name = backend
        self.backend = _backend
        self.device = device.lower()
        self.backend.assert_version()

    # ------------------------------------------------------------------ accessors
    def get_theta_dim(self):
        return self.params.d

    def get_params(self):
        return self.params

    def get_basis(self):
        return self.params.basis

    def set_solver_params(self, nstepsolver=50, numeric_grad=False, use_slow=False):
        assert nstepsolver > 0, "nstepsolver must be a positive number"
        assert type(nstepsolver) == int, "nstepsolver must be integer"
        assert type(numeric_grad) == bool, "numeric_grad must be bool"
        assert type(use_slow) == bool, "use_slow must be bool"
        if numeric_grad or use_slow:
            raise NotImplementedError("libcpab_b200 implements the fast analytic path only "
                                      "(no slow / numeric-gradient fallback)")
        self.params.nstepsolver = nstepsolver
        self.params.numeric_grad = numeric_grad
        self.params.use_slow = use_slow

    # ------------------------------------------------------------------ sampling helpers
    def uniform_meshgrid(self, n_points):
        return self.backend.uniform_meshgrid(self.params.ndim, self.params.domain_min,
                                             self.params.domain_max, n_points, self.device)

    def sample_transformation(self, n_sample=1, mean=None, cov=None):
        if mean is not None:
            self._check_type(mean); self._check_device(mean)
        if cov is not None:
            self._check_type(cov); self._check_device(cov)
        samples = self.backend.sample_transformation(self.params.d, n_sample, mean, cov, self.device)
        return self.backend.to(samples, device=self.device)

    def sample_transformation_with_prior(self, n_sample=1, mean=None, length_scale=0.1,
                                         output_variance=1):
        """Smooth prior over theta (libcpab/cpab.py:192-241); the O(nC^2) python block loop of the
        reference is replaced by one Kronecker-structured expression with the same values."""
        import torch
        centers = self.backend.to(self.tesselation.get_cell_centers(), device=self.device)
        dist = self.backend.pdist(centers)
        ppc = self.params.params_pr_cell
        big = 100 * self.backend.maximum(dist)
        eye = torch.eye(ppc, device=dist.device, dtype=dist.dtype)
        cov_init = torch.kron(dist, eye) + torch.kron(torch.ones_like(dist), big * (1 - eye))
        cov_avees = output_variance ** 2 * self.backend.exp(-(cov_init / (2 * length_scale ** 2)))
        B = self.backend.to(self.params.basis, device=self.device)
        cov_theta = self.backend.matmul(self.backend.transpose(B), self.backend.matmul(cov_avees, B))
        return self.sample_transformation(n_sample, mean=mean, cov=cov_theta)

    def identity(self, n_sample=1, epsilon=0):
        return self.backend.identity(self.params.d, n_sample, epsilon, self.device)

    # ------------------------------------------------------------------ the hot path
    def transform_grid(self, grid, theta):
        self._check_type(grid); self._check_device(grid)
        self._check_type(theta); self._check_device(theta)
        if len(grid.shape) == 3:
            assert grid.shape[0] == theta.shape[0], \
                "When passing a 3D grid, expects the first dimension to be of same length as " \
                "the first dimension of theta"
        return self.backend.transformer(grid, theta, self.params)

    def interpolate(self, data, grid, outsize):
        self._check_type(data); self._check_device(data)
        self._check_type(grid); self._check_device(grid)
        return self.backend.interpolate(self.params.ndim, data, grid, outsize)

    def transform_data(self, data, theta, outsize):
        self._check_type(data); self._check_device(data)
        self._check_type(theta); self._check_device(theta)
        # on theta's own device (one process may drive several GPUs)
        grid = self.backend.uniform_meshgrid(self.params.ndim, self.params.domain_min,
                                             self.params.domain_max, outsize, theta.device, _share=True)
        if grid.dtype != theta.dtype:
            grid = grid.to(theta.dtype)          # float64 check mode
        p = self.params
        fused = getattr(p, "fused_transform_data", None)
        if fused is None:
            fused = p.ndim == 1 or theta.shape[0] * grid.shape[1] <= (1 << 23)
        if (fused and not p.closed_form and data.dim() == p.ndim + 2
                and data.shape[0] == theta.shape[0] and data.dtype == theta.dtype == grid.dtype):
            from .transformer import fused_transform_data
            return fused_transform_data(data, theta, grid, p, outsize)
        grid_t = self.transform_grid(grid, theta)
        return self.interpolate(data, grid_t, outsize)

    def calc_vectorfield(self, grid, theta):
        self._check_type(grid); self._check_device(grid)
        self._check_type(theta); self._check_device(theta)
        return self.backend.calc_vectorfield(grid, theta, self.params)

    def findcellidx(self, grid):
        """Cell index per grid point (what visualize_tesselation plots in the reference)."""
        self._check_type(grid); self._check_device(grid)
        return self.backend.findcellidx(self.params.ndim, grid, self.params.nc)

    # ------------------------------------------------------------------ checks (cpab.py:479-520)
    def _check_input(self, tess_size, backend, device, zero_boundary, volume_perservation, override):
        assert len(tess_size) > 0 and len(tess_size) <= 3, "Transformer only supports 1D, 2D or 3D"
        assert type(tess_size) == list or type(tess_size) == tuple, \
            "Argument tess_size must be a list or tuple"
        assert all([type(e) == int for e in tess_size]), "All elements of tess_size must be integers"
        assert all([e > 0 for e in tess_size]), "All elements of tess_size must be positive"
        assert backend in ["numpy", "tensorflow", "pytorch"], \
            "Unknown backend, choose between 'numpy', 'tensorflow' or 'pytorch' "
        assert device in ["cpu", "gpu"], "Unknown device, choose between 'cpu' or 'gpu' "
        if backend != "pytorch" or device != "gpu":
            raise NotImplementedError("libcpab_b200 implements backend='pytorch', device='gpu' only "
                                      "(no multi-backend dispatch, no CPU fallback)")
        assert type(zero_boundary) == bool, "Argument zero_boundary must be True or False"
        assert type(volume_perservation) == bool, "Argument volume_perservation must be True or False"
        assert type(override) == bool, "Argument override must be True or False "

    def _check_type(self, x):
        assert isinstance(x, self.backend.backend_type()), \
            " Input has type {0} but expected type {1} ".format(type(x), self.backend.backend_type())

    def _check_device(self, x):
        assert self.backend.check_device(x, self.device), \
            "Input is placed on device {0} but the class expects it to be on device {1}".format(
                str(x.device), self.device)

    def __repr__(self):
        p = self.params
        return ("\n        CPAB transformer class (libcpab_b200).\n            Parameters:\n"
                f"                Tesselation size:           {p.nc}\n"
                f"                Total number of cells:      {p.nC}\n"
                f"                Theta size:                 {p.d}\n"
                f"                Domain lower bound:         {p.domain_min}\n"
                f"                Domain upper bound:         {p.domain_max}\n"
                f"                Zero Boundary:              {p.zero_boundary}\n"
                f"                Volume perservation:        {p.volume_perservation}\n"
                f"            Backend:                        {self.backend_name}\n")
