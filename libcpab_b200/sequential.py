"""`CpabSequential` -- chain of Cpab warps (libcpab/sequential.py:15-143): the grid is pushed
through every transformer in turn and the data are interpolated once with the last one."""
from __future__ import annotations

from .cpab import Cpab


class CpabSequential(object):
    def __init__(self, *cpab):
        self.n_cpab = len(cpab)
        self.cpab = cpab
        for i in range(self.n_cpab):
            assert isinstance(self.cpab[i], Cpab), \
                " Class {0} is not a member of the cpab core class ".format(i)
        self.ndim = self.cpab[0].params.ndim
        for i in range(1, self.n_cpab):
            assert self.ndim == self.cpab[i].params.ndim, \
                "Mismatching dimensionality of transformers. Transformer 1 have dimensionality " \
                "{0} but transformer {1} have dimensionality {2}".format(
                    self.ndim, i + 1, self.cpab[i].params.ndim)
        self.backend = self.cpab[0].backend
        self.backend_name = self.cpab[0].backend_name

    def get_theta_dim(self):
        return [c.get_theta_dim() for c in self.cpab]

    def get_params(self):
        return [c.get_params() for c in self.cpab]

    def get_basis(self):
        return [c.get_basis() for c in self.cpab]

    def uniform_meshgrid(self, n_points):
        return self.cpab[0].uniform_meshgrid(n_points)

    def sample_transformation(self, n_sample, means=None, covs=None):
        means = self.n_cpab * [None] if means is None else means
        covs = self.n_cpab * [None] if covs is None else covs
        assert len(means) == self.n_cpab, \
            "The number of supplied means should be equal to the number of transformations"
        assert len(covs) == self.n_cpab, \
            "The number of supplied covariances should be equal to the number of transformations"
        return [c.sample_transformation(n_sample, m, v) for c, m, v in zip(self.cpab, means, covs)]

    def identity(self, n_sample, epsilon=0):
        return [c.identity(n_sample, epsilon) for c in self.cpab]

    def transform_grid(self, grid, thetas, output_all=False):
        self._assert_theta_shape(thetas)
        if not output_all:
            for i in range(self.n_cpab):
                grid = self.cpab[i].transform_grid(grid, thetas[i])
            return grid
        grids = [self.cpab[0].transform_grid(grid, thetas[0])]
        for i in range(1, self.n_cpab):
            grids.append(self.cpab[i].transform_grid(grids[-1], thetas[i]))
        return grids

    def transform_data(self, data, thetas, outsize, output_all=False):
        self._assert_theta_shape(thetas)
        grid = self.uniform_meshgrid(outsize)
        grid_t = self.transform_grid(grid, thetas, output_all=output_all)
        if not output_all:
            return self.cpab[-1].interpolate(data, grid_t, outsize)
        return [self.cpab[i].interpolate(data, grid_t[i], outsize) for i in range(self.n_cpab)]

    def _assert_theta_shape(self, thetas):
        n_theta = len(thetas)
        assert n_theta == self.n_cpab, " Number of parametrizations needed are {0}".format(self.n_cpab)
        batch_size = thetas[0].shape[0]
        for i in range(1, n_theta):
            assert batch_size == thetas[i].shape[0], " Batch size should be the same for all theta's "

    def __repr__(self):
        return "\n".join("======= Transformer {0} ======= \n{1}".format(i + 1, c)
                         for i, c in enumerate(self.cpab))
