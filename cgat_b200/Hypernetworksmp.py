"""Hypernetwork node update of CGAtNet, parameter-compatible with the reference's
CGAT/Hypernetworksmp.py (same class names, module tree and initialisers).

Per atom n the reference predicts a full dense layer from a hyper-input h_n
(HyperLinear.forward, reference Hypernetworksmp.py:243-254):
    p_n = Linear(128->16512)( [Linear+Tanh]x4 (h_n) );  W_n = p_n[:16384].view(128,128);  b_n = p_n[16384:]
and applies it to that atom's aggregated message (BatchLinear.forward, :205-209), four times, the
first three followed by LayerNorm(no affine)+Tanh (:103-107).  As written this materialises
66 KB per atom per hyper-layer in HBM.  Here `HyperLinear.apply_to` contracts the predicted weights
with the activation inside one op (ops.hyper_linear) so the (N,16512) tensor never has to be stored
for backward; the module-returning API of the reference (HyperLinear.forward -> BatchLinear) is kept
for completeness but is not on the hot path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class FCLayer(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(in_features, out_features), nn.Tanh())

    def forward(self, input):
        return self.net(input)


def _kaiming(m):
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight, a=0.0, nonlinearity="leaky_relu", mode="fan_in")


def last_hyper_layer_init(m):
    """reference Hypernetworksmp.py:212-219"""
    if isinstance(m, nn.Linear):
        _kaiming(m)
        m.weight.data *= 1e-1


class FCBlock(nn.Module):
    """[Linear+Tanh] x (1+num_hidden_layers) then Linear (outermost_linear) — reference :36-83."""

    def __init__(self, hidden_ch, num_hidden_layers, in_features, out_features, outermost_linear=False):
        super().__init__()
        layers = [FCLayer(in_features, hidden_ch)]
        layers += [FCLayer(hidden_ch, hidden_ch) for _ in range(num_hidden_layers)]
        layers.append(nn.Linear(hidden_ch, out_features) if outermost_linear else FCLayer(hidden_ch, out_features))
        self.net = nn.Sequential(*layers)
        self.net.apply(_kaiming)

    def __getitem__(self, item):
        return self.net[item]

    def trunk(self, h):
        """All layers but the last: the (N, hidden) code the big last Linear expands."""
        for layer in list(self.net)[:-1]:
            h = layer(h)
        return h

    def forward(self, input):
        return self.net(input)


class BatchLinear(nn.Module):
    """Per-sample dense layer with explicit (batch,out,in) weights — reference :188-209.
    API parity only; the hot path never materialises these weights."""

    def __init__(self, weights, biases):
        super().__init__()
        self.weights, self.biases = weights, biases

    def forward(self, input):
        return input.matmul(self.weights.transpose(-1, -2)) + self.biases


class HyperLinear(nn.Module):
    """Hypernetwork predicting one linear layer — reference :222-254."""

    def __init__(self, in_ch, out_ch, hyper_in_ch, hyper_num_hidden_layers, hyper_hidden_ch):
        super().__init__()
        self.in_ch, self.out_ch = in_ch, out_ch
        self.hypo_params = FCBlock(in_features=hyper_in_ch, hidden_ch=hyper_hidden_ch,
                                   num_hidden_layers=hyper_num_hidden_layers,
                                   out_features=in_ch * out_ch + out_ch, outermost_linear=True)
        self.hypo_params[-1].apply(last_hyper_layer_init)

    def forward(self, hyper_input):
        p = self.hypo_params(hyper_input)
        w = p[..., : self.in_ch * self.out_ch].reshape(*p.shape[:-1], self.out_ch, self.in_ch)
        b = p[..., self.in_ch * self.out_ch:].reshape(*p.shape[:-1], 1, self.out_ch)
        return BatchLinear(weights=w, biases=b)

    def trunk_spec(self):
        """([(W, b) of the tanh layers], (last weight, last bias, rows of the predicted matrix)) for ops.hyper_trunks."""
        net = list(self.hypo_params.net)
        last = net[-1]
        return [(l.net[0].weight, l.net[0].bias) for l in net[:-1]], (last.weight, last.bias, self.in_ch * self.out_ch)

    def apply_to(self, hyper_input, y, z=None, e=None):
        """y_out[n] = W_n y[n] + b_n with (W_n, b_n) predicted from hyper_input[n]; fused.  z / e: this layer's
        trunk output and bias tail when HyperFC has already run all trunks of the node layer in one kernel."""
        if z is None:
            z = self.hypo_params.trunk(hyper_input)
        last = self.hypo_params[-1]
        return ops.hyper_linear(z, last.weight, last.bias, y, self.out_ch, e=e)


class HyperLayer(nn.Module):
    """HyperLinear + LayerNorm(no affine) + Tanh — reference :86-114."""

    def __init__(self, in_ch, out_ch, hyper_in_ch, hyper_num_hidden_layers, hyper_hidden_ch):
        super().__init__()
        self.hyper_linear = HyperLinear(in_ch, out_ch, hyper_in_ch, hyper_num_hidden_layers, hyper_hidden_ch)
        self.norm_nl = nn.Sequential(nn.LayerNorm([out_ch], elementwise_affine=False), nn.Tanh())

    def forward(self, hyper_input):
        return nn.Sequential(self.hyper_linear(hyper_input), self.norm_nl)

    def apply_to(self, hyper_input, y, z=None, e=None):
        y = self.hyper_linear.apply_to(hyper_input, y, z=z, e=e)
        return torch.tanh(F.layer_norm(y, (y.shape[-1],), eps=1e-5))


class HyperFC(nn.Module):
    """Hypernetwork predicting a whole MLP — reference :117-185."""

    def __init__(self, hyper_in_ch, hyper_num_hidden_layers, hyper_hidden_ch, hidden_ch, num_hidden_layers,
                 in_ch, out_ch, outermost_linear=False):
        super().__init__()
        kw = dict(hyper_in_ch=hyper_in_ch, hyper_num_hidden_layers=hyper_num_hidden_layers,
                  hyper_hidden_ch=hyper_hidden_ch)
        self.layers = nn.ModuleList([HyperLayer(in_ch=in_ch, out_ch=hidden_ch, **kw)])
        self.layers.extend(HyperLayer(in_ch=hidden_ch, out_ch=hidden_ch, **kw) for _ in range(num_hidden_layers))
        last = HyperLinear if outermost_linear else HyperLayer
        self.layers.append(last(in_ch=hidden_ch, out_ch=out_ch, **kw))

    def forward(self, hyper_input):
        return nn.Sequential(*[layer(hyper_input) for layer in self.layers])

    def apply_to(self, hyper_input, y):
        # all trunks see the same hyper-input: one chained kernel for the whole node layer (ops.hyper_trunks)
        specs = [(l.hyper_linear if isinstance(l, HyperLayer) else l).trunk_spec() for l in self.layers]
        zs, es = ops.hyper_trunks(hyper_input, [s[0] for s in specs], [s[1] for s in specs])
        for layer, z, e in zip(self.layers, zs, es):
            y = layer.apply_to(hyper_input, y, z=z, e=e)
        return y


class H_Net_0(nn.Module):
    """First-layer node update: hyper-input = current atom features — reference :257-285."""

    def __init__(self, hyper_in_ch, hyper_num_hidden_layers, hyper_hidden_ch, hidden_ch, num_hidden_layers,
                 in_ch, out_ch, outermost_linear=True):
        super().__init__()
        self.Hyper = HyperFC(hyper_in_ch, hyper_num_hidden_layers, hyper_hidden_ch, hidden_ch,
                             num_hidden_layers, in_ch, out_ch, outermost_linear=True)
        self.out_ch = out_ch

    def forward(self, h_0, x):
        return self.Hyper.apply_to(h_0, x)


class H_Net(nn.Module):
    """Later-layer node update: hyper-input = d*h_0 + (1-d)*x with d = clamp(damping,0,1); h_t is
    accepted and ignored exactly like the reference (:288-313).  The in-place clamp of
    `damping.data` on every forward is replicated."""

    def __init__(self, hyper_in_ch, hyper_num_hidden_layers, hyper_hidden_ch, hidden_ch, num_hidden_layers,
                 in_ch, out_ch, outermost_linear=True):
        super().__init__()
        self.Hyper = HyperFC(hyper_in_ch, hyper_num_hidden_layers, hyper_hidden_ch, hidden_ch,
                             num_hidden_layers, in_ch, out_ch, outermost_linear=True)
        self.damping = nn.Parameter(torch.rand(1))
        self.out_ch = out_ch

    def forward(self, h_0, h_t, x):
        with torch.no_grad():
            self.damping.data.clamp_(0.0, 1.0)
        return self.Hyper.apply_to(self.damping * h_0 + (1 - self.damping) * x, x)
