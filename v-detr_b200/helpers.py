"""Small building blocks of the decoder (drop-in for /root/reference/models/helpers.py:17-145).

Same class names, constructor arguments and state_dict keys as the reference so that checkpoints load
unchanged; these layers are plain PyTorch (cuBLAS / cuDNN) -- SURVEY.md 8(a) rows a7-a10 keep them outside
the hand-written kernels.
"""
from __future__ import annotations

import copy
from functools import partial

import torch.nn as nn
import torch.nn.functional as F


def pointwise_tokens(seq, x):
    """Apply a Conv1d(k=1)/BatchNorm1d/activation/Dropout stack to token-major features x [T, C_in] -> [T, C_out].

    A kernel-size-1 Conv1d over [B, C, N] is the Linear map of every token, and BatchNorm1d over [B, C, N] and over
    [B*N, C] normalises the same B*N samples per channel, so the same parameters / running statistics give the same
    result (up to summation order) -- but as plain GEMMs on contiguous rows instead of cuDNN convolutions on a
    permuted copy (the cuDNN weight-gradient kernels for the 1/3/18-channel output convolutions alone cost ~80 us
    per call on a B200).  Used for the box heads (models/helpers.py:74-141) and the query-position MLP (:17-33)."""
    from . import ops
    mods, i = list(seq), 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Conv1d):
            assert m.kernel_size == (1,) and m.stride == (1,) and m.groups == 1
            x = ops.linear(x, m.weight.squeeze(-1), m.bias)
        elif (type(m) is nn.BatchNorm1d and i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
              and ops.bn_relu_train_supported(x, m)):
            # batch statistics + running-stat update + ReLU (+ the Dropout behind it) fused (csrc/batchnorm.cu)
            drop = mods[i + 2] if i + 2 < len(mods) and type(mods[i + 2]) is nn.Dropout and mods[i + 2].training else None
            x = ops.bn_relu_train(x, m, drop.p if drop is not None else 0.0)
            i += 1 if drop is None else 2
        else:
            x = m(x)
        i += 1
    return x


class PositionEmbeddingLearned(nn.Module):
    """Learned absolute position embedding: Conv1d - BN - ReLU - Conv1d on [B, N, C_in] (helpers.py:17-33)."""

    def __init__(self, input_channel, num_pos_feats=288):
        super().__init__()
        self.position_embedding_head = nn.Sequential(
            nn.Conv1d(input_channel, num_pos_feats, kernel_size=1),
            nn.BatchNorm1d(num_pos_feats),
            nn.ReLU(inplace=True),
            nn.Conv1d(num_pos_feats, num_pos_feats, kernel_size=1))

    def forward(self, xyz):
        return self.position_embedding_head(xyz.transpose(1, 2).contiguous())

    def forward_tokens(self, xyz):
        """[B, N, C_in] -> [N, B, C_out]: what callers build with forward(xyz).permute(2, 0, 1), computed token-major."""
        B, N, C = xyz.shape
        out = pointwise_tokens(self.position_embedding_head, xyz.reshape(B * N, C))
        return out.view(B, N, -1).permute(1, 0, 2)


class BatchNormDim1Swap(nn.BatchNorm1d):
    """BatchNorm over the channel dim of an [HW, N, C] tensor (helpers.py:36-53)."""

    def forward(self, x):
        return super().forward(x.permute(1, 2, 0)).permute(2, 0, 1)


class Linear(nn.Linear):
    """nn.Linear (same parameters / state_dict) routed through ops.linear (bias gradient by the library's kernel)."""

    def forward(self, x):
        from . import ops
        return ops.linear(x, self.weight, self.bias)


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm (same parameters / state_dict) whose fp32 CUDA path runs the library's warp-per-row kernels
    (csrc/layernorm.cu); anything else (CPU, other dtypes, widths the kernels do not cover) is stock PyTorch."""

    def forward(self, x):
        from . import ops
        if len(self.normalized_shape) == 1 and ops.layer_norm_supported(x, self.weight, self.bias):
            return ops.layer_norm(x, self.weight, self.bias, self.eps)
        return super().forward(x)


NORM_DICT = {"bn": BatchNormDim1Swap, "bn1d": nn.BatchNorm1d, "id": nn.Identity, "ln": LayerNorm}
ACTIVATION_DICT = {"relu": nn.ReLU, "gelu": nn.GELU, "leakyrelu": partial(nn.LeakyReLU, negative_slope=0.1)}
WEIGHT_INIT_DICT = {"xavier_uniform": nn.init.xavier_uniform_}


class GenericMLP(nn.Module):
    """helpers.py:74-141: [Linear|Conv1d(k=1) - norm - act - dropout] x len(hidden_dims) + output layer."""

    def __init__(self, input_dim, hidden_dims, output_dim, norm_fn_name=None, activation="relu", use_conv=False,
                 dropout=None, hidden_use_bias=False, output_use_bias=True, output_use_activation=False,
                 output_use_norm=False, weight_init_name=None):
        super().__init__()
        act = ACTIVATION_DICT[activation]
        norm = NORM_DICT[norm_fn_name] if norm_fn_name is not None else None
        if norm_fn_name == "ln" and use_conv:
            norm = lambda c: nn.GroupNorm(1, c)  # noqa: E731  (LayerNorm over channels of a conv feature map)
        if dropout is not None and not isinstance(dropout, list):
            dropout = [dropout] * len(hidden_dims)

        def dense(i, o, bias):
            return nn.Conv1d(i, o, 1, bias=bias) if use_conv else nn.Linear(i, o, bias=bias)

        mods, width = [], input_dim
        for li, h in enumerate(hidden_dims):
            mods.append(dense(width, h, hidden_use_bias))
            if norm:
                mods.append(norm(h))
            mods.append(act())
            if dropout is not None:
                mods.append(nn.Dropout(p=dropout[li]))
            width = h
        mods.append(dense(width, output_dim, output_use_bias))
        if output_use_norm:
            mods.append(norm(output_dim))
        if output_use_activation:
            mods.append(act())
        self.layers = nn.Sequential(*mods)
        if weight_init_name is not None:
            self.do_weight_init(weight_init_name)

    def do_weight_init(self, weight_init_name):
        init = WEIGHT_INIT_DICT[weight_init_name]
        for _, p in self.named_parameters():
            if p.dim() > 1:
                init(p)

    def forward(self, x):
        return self.layers(x)

    def forward_tokens(self, x):
        """Token-major evaluation of a use_conv=True stack: x [T, C_in] -> [T, C_out] (see pointwise_tokens)."""
        return pointwise_tokens(self.layers, x)

    @property
    def supports_tokens(self):
        return all(isinstance(m, (nn.Conv1d, nn.BatchNorm1d, nn.ReLU, nn.GELU, nn.LeakyReLU, nn.Dropout, nn.Identity))
                   and not isinstance(m, BatchNormDim1Swap) for m in self.layers)


def get_clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])
