"""Drop-in for the reference's `causal_conv1d` package (requirements/Mamba/causal-conv1d/causal_conv1d/__init__.py),
backed by mmunet_b200's sm_100a kernels.  `causal_conv1d_update` is the single-token decode op, which MM-UNet never
calls; it is exported (the TFM mamba_simple.py imports the name, :13-16) and raises if used."""
from mmunet_b200.ops import causal_conv1d_fn  # noqa: F401

__version__ = "1.0.0"


def causal_conv1d_update(x, conv_state, weight, bias=None, activation=None):
    raise NotImplementedError("causal_conv1d_update (decode step) is outside the MM-UNet training hot path")
