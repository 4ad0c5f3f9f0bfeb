"""`Mlp` (flash_attn/modules/mlp.py:13-30): fc1 -> activation -> fc2, un-fused variant."""
import torch.nn as nn
import torch.nn.functional as F

from ..ops.fused_dense import FusedDenseGeluDense  # noqa: F401  (re-exported like the reference does)


class Mlp(nn.Module):

    def __init__(self, in_features, hidden_features=None, out_features=None, activation=F.gelu,
                 return_residual=False, device=None, dtype=None):
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.return_residual = return_residual
        self.fc1 = nn.Linear(in_features, hidden_features, **factory_kwargs)
        self.activation = activation
        self.fc2 = nn.Linear(hidden_features, out_features, **factory_kwargs)

    def forward(self, x):
        y = self.fc2(self.activation(self.fc1(x)))
        return y if not self.return_residual else (y, x)
