"""TEST INFRASTRUCTURE ONLY — stand-ins for the five third-party symbols the reference imports.

The reference (hyllios/CGAT) imports `torch_scatter` and `torch_geometric`, neither of which is
installed here.  This module registers minimal stand-in modules in `sys.modules`
so that the UNMODIFIED reference files under /root/reference can be imported and run on CPU.
It is used only by `oracle/make_golden.py` (in the build container, where /root/reference exists)
to generate the committed fixtures under tests/golden/.  Nothing in the product imports it.

Semantics reproduced (documented behaviour of the pinned third-party versions, README.md:7-8):
  torch_scatter 2.0.x   scatter_add / scatter_max / scatter_mean   (call sites CGAT.py:6,60;
                        roost_message.py:27,280,307,311,315)
  torch_geometric 2.0.x MessagePassing(aggr='add', flow='source_to_target').propagate
                        (CGAT.py:233,275,313), utils.softmax (CGAT.py:59,323),
                        data.Data/Batch (data.py:1,140; lightning_module.py:21,200)
"""
import sys
import types

import torch


def _bcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    if dim < 0:
        dim = src.dim() + dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    return res.scatter_add(dim, _bcast(index, src, dim), src)


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    if dim < 0:
        dim = src.dim() + dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    idx = _bcast(index, src, dim)
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    # include_self=False: empty segments keep the fill value 0 (torch_scatter behaviour)
    res = res.scatter_reduce(dim, idx, src, reduce='amax', include_self=False)
    # argmax is returned by torch_scatter but never consumed by the reference
    return res, None


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    s = scatter_add(src, index, dim, None, dim_size)
    cnt = scatter_add(torch.ones_like(src), index, dim, None, dim_size).clamp_(min=1)
    return s / cnt


def pyg_softmax(src, index, ptr=None, num_nodes=None, dim=0):
    """torch_geometric.utils.softmax (2.0.x): exp(src - segmax) / (segsum + 1e-16) along dim 0."""
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    idx = _bcast(index, src, 0)
    mx = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    mx = mx.scatter_reduce(0, idx, src.detach(), reduce='amax', include_self=False)
    out = (src - mx.index_select(0, index)).exp()
    den = scatter_add(out, index, dim=0, dim_size=n).index_select(0, index)
    return out / (den + 1e-16)


class MessagePassing(torch.nn.Module):
    """aggr='add', flow='source_to_target': x_j = x[edge_index[0]], x_i = x[edge_index[1]],
    aggregation index = edge_index[1]."""

    def __init__(self, aggr='add', flow='source_to_target', node_dim=-2, **kwargs):
        super().__init__()
        assert aggr == 'add' and flow == 'source_to_target'
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        import inspect
        x = kwargs['x']
        n = x.size(self.node_dim)
        msg_params = inspect.signature(self.message).parameters
        upd_params = inspect.signature(self.update).parameters
        margs = {}
        for name in msg_params:
            if name.endswith('_i') and name[:-2] in kwargs:
                margs[name] = kwargs[name[:-2]].index_select(self.node_dim, edge_index[1])
            elif name.endswith('_j') and name[:-2] in kwargs:
                margs[name] = kwargs[name[:-2]].index_select(self.node_dim, edge_index[0])
            elif name == 'edge_index_i':
                margs[name] = edge_index[1]
            elif name == 'edge_index_j':
                margs[name] = edge_index[0]
            elif name == 'size_i':
                margs[name] = n
            elif name in kwargs:
                margs[name] = kwargs[name]
        out = self.message(**margs)
        out = scatter_add(out, edge_index[1], dim=self.node_dim, dim_size=n)
        uargs = {k: kwargs[k] for k in upd_params if k in kwargs}
        return self.update(out, **uargs)


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kw):
        self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return self.x.shape[0]

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class Batch(Data):
    @staticmethod
    def from_data_list(data_list):
        xs, eis, eas, ys, bs = [], [], [], [], []
        off = 0
        for i, d in enumerate(data_list):
            n = d.x.shape[0]
            xs.append(d.x)
            eis.append(d.edge_index + off)
            eas.append(d.edge_attr)
            ys.append(d.y)
            bs.append(torch.full((n,), i, dtype=torch.long))
            off += n
        b = Batch(x=torch.cat(xs), edge_index=torch.cat(eis, dim=1), edge_attr=torch.cat(eas),
                  y=torch.cat(ys))
        b.batch = torch.cat(bs)
        return b


def install():
    if 'torch_scatter' in sys.modules and getattr(sys.modules['torch_scatter'], '_cgat_standin', False):
        return
    ts = types.ModuleType('torch_scatter')
    ts.scatter_add, ts.scatter_max, ts.scatter_mean = scatter_add, scatter_max, scatter_mean
    ts._cgat_standin = True
    tg = types.ModuleType('torch_geometric')
    tgnn = types.ModuleType('torch_geometric.nn')
    tgnn.MessagePassing = MessagePassing
    tgu = types.ModuleType('torch_geometric.utils')
    tgu.softmax = pyg_softmax
    tgd = types.ModuleType('torch_geometric.data')
    tgd.Data, tgd.Batch = Data, Batch
    tg.nn, tg.utils, tg.data = tgnn, tgu, tgd
    sys.modules.update({'torch_scatter': ts, 'torch_geometric': tg, 'torch_geometric.nn': tgnn,
                        'torch_geometric.utils': tgu, 'torch_geometric.data': tgd})


def import_reference(root='/root/reference'):
    """Import the unmodified reference modules. Only possible where `root` exists."""
    install()
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    mods = {}
    for name in ('CGAT.message_changed', 'CGAT.roost_message', 'CGAT.Hypernetworksmp', 'CGAT.CGAT'):
        mods[name.split('.')[-1]] = importlib.import_module(name)
    return mods
