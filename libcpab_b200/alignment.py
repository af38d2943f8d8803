"""`CpabAligner` -- sampling- and gradient-based alignment (libcpab/alignment.py:10-106), the
pure-python caller of the hot path in which, for multi-GPU runs, the theta-gradient all-reduce
hooks in (see libcpab_b200.distributed)."""
from __future__ import annotations

import torch

from .cpab import Cpab


class CpabAligner(object):
    def __init__(self, cpab_class):
        assert isinstance(cpab_class, Cpab), \
            "The input class needs to be an instance of the core cpab class "
        self.T = cpab_class
        self.backend = cpab_class.backend

    def alignment_by_sampling(self, x1, x2, maxiter=100):
        self.T._check_type(x1)
        self.T._check_type(x2)
        assert x1.shape == x2.shape, " Two data points does not have the same shape "
        outsize = tuple(x2.shape[2:])
        current_sample = self.T.identity(1)
        current_error = self.backend.norm(x1 - x2)
        accept = 0
        for _ in range(maxiter):
            theta = 1e-1 * self.T.sample_transformation(1, mean=current_sample.flatten())
            x1_trans = self.T.transform_data(x1, theta, outsize=outsize)
            new_error = self.backend.norm(x1_trans - x2)
            if new_error < current_error:
                current_sample, current_error = theta, new_error
                accept += 1
        self.accept_ratio = accept / max(maxiter, 1)
        return current_sample

    def alignment_by_gradient(self, x1, x2, maxiter=100, lr=1e-2, grad_hook=None):
        """Adam on theta (alignment.py:60-87).  `grad_hook(theta)` runs between backward() and the
        optimiser step; multi-GPU callers pass libcpab_b200.distributed.allreduce_grad_ there."""
        self.T._check_type(x1)
        self.T._check_type(x2)
        assert x1.shape == x2.shape, " Two data points does not have the same shape "
        theta = self.T.identity(1, epsilon=1e-6).requires_grad_(True)
        optimizer = torch.optim.Adam([theta], lr=lr)
        self.losses = []
        for _ in range(maxiter):
            optimizer.zero_grad()
            x1_trans = self.T.transform_data(x1, theta, outsize=x1.shape[2:])
            loss = self.backend.norm(x1_trans - x2)
            loss.backward()
            if grad_hook is not None:
                grad_hook(theta)
            optimizer.step()
            self.losses.append(loss.detach())
        return theta
