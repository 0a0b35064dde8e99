"""Preprocess transforms (reference: components/transforms.py:4-21)."""
import torch as th


class Transform:
    def transform(self, tensor):
        raise NotImplementedError

    def infer_output_info(self, vshape_in, dtype_in):
        raise NotImplementedError


class OneHot(Transform):
    """integer index [..., 1] -> float32 one-hot [..., out_dim]."""

    def __init__(self, out_dim):
        self.out_dim = out_dim

    def transform(self, tensor):
        out = th.zeros(*tensor.shape[:-1], self.out_dim, dtype=th.float32, device=tensor.device)
        out.scatter_(-1, tensor.long(), 1.0)
        return out

    def infer_output_info(self, vshape_in, dtype_in):
        return (self.out_dim,), th.float32
