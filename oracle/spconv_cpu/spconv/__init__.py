"""CPU ORACLE of the ``spconv`` 1.x Python API that second.pytorch imports.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).
The product package is second.pytorch_b200/spconv (CUDA); it never imports this one.

PARITY UNPINNED at this boundary: the reference holds no tests or golden vectors for spconv
 and spconv itself is absent.  What this restates, and the reference call sites:

* ``SparseConvTensor``  second/pytorch/models/middle.py:199-200,206 ; resnet.py:54-64
* ``SparseSequential``/``SparseModule``  middle.py:145 ; resnet.py:32,69
* ``SubMConv3d``/``SparseConv3d``  middle.py:146-189 (bias handled through
  torchplus/tools.py:11-46 ``change_default_args``: ``bias`` must be a named parameter)
Importable as ``spconv`` by putting ``oracle/spconv_cpu`` on ``sys.path``.
"""
import math

import numpy as np
import torch
from torch import nn

from . import ops  # noqa: F401
from . import utils  # noqa: F401

__oracle__ = True


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self.features = features
        self.indices = indices
        if self.indices.dtype != torch.int32:
            self.indices = self.indices.int()
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = {}
        self.grid = grid

    @property
    def spatial_size(self):
        return int(np.prod(self.spatial_shape))

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key, None)

    def dense(self, channels_first=True):
        shape = [self.batch_size] + list(self.spatial_shape) + [self.features.shape[1]]
        out = torch.zeros(*shape, dtype=self.features.dtype, device=self.features.device)
        idx = self.indices.long()
        out[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]] = self.features
        if not channels_first:
            return out
        return out.permute(0, 4, 1, 2, 3).contiguous()

    @property
    def sparity(self):
        return self.indices.shape[0] / np.prod(self.spatial_shape) / self.batch_size


class SparseModule(nn.Module):
    """marker base class: modules that take / return a SparseConvTensor."""
    pass


def is_spconv_module(module):
    return isinstance(module, SparseModule)


class SparseSequential(SparseModule):
    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], dict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError("index {} is out of range".format(idx))
        if idx < 0:
            idx += len(self)
        return list(self._modules.values())[idx]

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        for module in self._modules.values():
            if is_spconv_module(module):
                assert isinstance(input, SparseConvTensor)
                input = module(input)
            elif isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    input.features = module(input.features)
            else:
                input = module(input)
        return input


def _triple(v, ndim=3):
    if isinstance(v, (list, tuple, np.ndarray)):
        assert len(v) == ndim
        return [int(x) for x in v]
    return [int(v)] * ndim


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, subm=False, output_padding=0, transposed=False, inverse=False,
                 indice_key=None):
        super().__init__()
        assert groups == 1
        assert ndim == 3
        self.ndim = ndim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _triple(kernel_size)
        self.conv1x1 = int(np.prod(self.kernel_size)) == 1
        self.stride = _triple(stride)
        self.padding = _triple(padding)
        self.dilation = _triple(dilation)
        self.transposed = transposed
        self.inverse = inverse
        self.output_padding = _triple(output_padding)
        self.groups = groups
        self.subm = subm
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.Tensor(*self.kernel_size, in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, input):
        assert isinstance(input, SparseConvTensor)
        features = input.features
        indices = input.indices
        spatial_shape = input.spatial_shape
        batch_size = input.batch_size
        if not self.subm:
            out_spatial_shape = ops.get_conv_output_size(spatial_shape, self.kernel_size, self.stride,
                                                         self.padding, self.dilation)
        else:
            out_spatial_shape = spatial_shape
        if self.conv1x1:
            feats = torch.mm(input.features, self.weight.view(self.in_channels, self.out_channels))
            if self.bias is not None:
                feats = feats + self.bias
            out_tensor = SparseConvTensor(feats, input.indices, input.spatial_shape, input.batch_size)
            out_tensor.indice_dict = input.indice_dict
            out_tensor.grid = input.grid
            return out_tensor
        datas = input.find_indice_pair(self.indice_key)
        if self.indice_key is not None and datas is not None:
            outids, _, indice_pairs, indice_pair_num, _ = datas
        else:
            outids, indice_pairs, indice_pair_num = ops.get_indice_pairs(
                indices, batch_size, spatial_shape, self.kernel_size, self.stride, self.padding,
                self.dilation, self.output_padding, self.subm, self.transposed, grid=input.grid)
            input.indice_dict[self.indice_key] = (outids, indices, indice_pairs, indice_pair_num,
                                                  spatial_shape)
        out_features = ops.indice_conv(features, self.weight, indice_pairs, indice_pair_num,
                                       outids.shape[0], False, self.subm)
        if self.bias is not None:
            out_features = out_features + self.bias
        out_tensor = SparseConvTensor(out_features, outids, out_spatial_shape, batch_size)
        out_tensor.indice_dict = input.indice_dict
        out_tensor.grid = input.grid
        return out_tensor


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         bias, indice_key=indice_key)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         bias, True, indice_key=indice_key)


class ToDense(SparseModule):
    def forward(self, x):
        return x.dense()
