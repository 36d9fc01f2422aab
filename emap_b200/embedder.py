"""Positional-encoding descriptor (host-side mirror of ``src/models/embedder.py``).

The encoding itself is evaluated inside the MLP kernels' input stage (emap_b200/csrc/mlp_tc.cu);
this module only keeps the reference's ``get_embedder(multires, input_dims) -> (embed_fn, out_dim)``
contract so code that introspects ``UDFNetwork.embed_fn_fine`` keeps working.  ``embed_fn`` returns
the kernel-computed encoding (it is the PE output of ``emap_udf_forward``).
"""
from __future__ import annotations


class Embedder:
    """gamma(x) = [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]  (embedder.py:5-35)."""

    def __init__(self, multires: int, input_dims: int = 3):
        if input_dims != 3:
            raise NotImplementedError("emap_b200 embeds 3-d points only")
        self.multires = int(multires)
        self.input_dims = input_dims
        self.out_dim = input_dims * (1 + 2 * self.multires)
        self.freq_bands = [float(2 ** j) for j in range(self.multires)]

    def embed(self, inputs):
        from . import ops
        return ops.positional_encoding(inputs, self.multires)


def get_embedder(multires, input_dims=3):
    eo = Embedder(multires, input_dims)
    return eo.embed, eo.out_dim
