"""CPU oracle for MAP-ConvNeXt (backbone + MAP attention-pooling head)  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Functional, state-dict driven restatement in plain PyTorch of `/root/reference/MAP/models/map_convnext.py` and
`/root/reference/MAP/models/map.py` for the two registered MAP-ConvNeXt variants.  Parity status: PINNED -- checked
against the unmodified reference (imported through oracle/timm_shim) by tests/golden/make_golden.py, whose outputs are
committed under tests/golden/map_convnext_model.pt.  Known-answer tests: 47 833 760 / 82 837 664 parameters
(MAP/README.MD:308,373).  Dropout layers (attn_drop/drop 0.05, map.py:149) are identity here: the parity contract is
eval mode, and train mode with every drop rate set to 0.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from .ga_convnext_oracle import batchnorm, layernorm2d

Tensor = torch.Tensor
State = Dict[str, Tensor]


@dataclass(frozen=True)
class MAPSpec:
    """map_convnext.py:199-239 constructor arguments."""
    depths: Tuple[int, ...]
    dims: Tuple[int, ...] = (96, 192, 384, 768)
    last_dim: int = 384
    n_groups: int = 4
    n_tokens: int = 2
    gram_group: int = 24
    bp_dim: int = 384
    ca_dim: int = 384
    num_heads: int = 12
    mlp_ratio: int = 4
    mlp_groups: int = 2
    num_classes: int = 1000

    @property
    def tri(self):
        return self.bp_dim * (self.bp_dim + 1) // 2

    @property
    def all_tokens(self):
        return self.n_tokens + 1      # + the self-distillation (mean) token, map.py:245-247


SPECS = {
    'map_convnext_tiny': MAPSpec((3, 3, 9, 3), n_tokens=2, gram_group=24),
    'map_convnext_small': MAPSpec((3, 3, 27, 3), n_tokens=3, gram_group=16),
}
PARAM_COUNTS = {'map_convnext_tiny': 47833760, 'map_convnext_small': 82837664}


def state_shapes(spec: MAPSpec) -> Dict[str, Tuple[Tuple[int, ...], str]]:
    S: Dict[str, Tuple[Tuple[int, ...], str]] = {}

    def conv(name, cout, cin_g, k, bias=True):
        S[name + '.weight'] = ((cout, cin_g, k, k), 'w')
        if bias:
            S[name + '.bias'] = ((cout,), 'b')

    def lin(name, cout, cin, bias=True):
        S[name + '.weight'] = ((cout, cin), 'w')
        if bias:
            S[name + '.bias'] = ((cout,), 'b')

    def ln(name, c):
        S[name + '.weight'] = ((c,), 'g')
        S[name + '.bias'] = ((c,), 'b')

    def bn(name, c):
        ln(name, c)
        S[name + '.running_mean'] = ((c,), 'rm')
        S[name + '.running_var'] = ((c,), 'rv')
        S[name + '.num_batches_tracked'] = ((), 'n')

    d = spec.dims
    conv('downsample_layers.0.0', d[0], 3, 4)
    ln('downsample_layers.0.1', d[0])
    for i in range(3):
        ln(f'downsample_layers.{i + 1}.0', d[i])
        conv(f'downsample_layers.{i + 1}.1', d[i + 1], d[i], 2)
    for i in range(4):
        for j in range(spec.depths[i]):
            p = f'stages.{i}.{j}.'
            S[p + 'gamma'] = ((d[i],), 'g')
            conv(p + 'dwconv', d[i], 1, 7)
            ln(p + 'norm', d[i])
            lin(p + 'pwconv1', 4 * d[i], d[i])
            lin(p + 'pwconv2', d[i], 4 * d[i])
    L_, T = spec.last_dim, spec.n_tokens
    cat = d[0] + sum(d)
    conv('head.mmcap.multi_scale.concat_conv.0', L_, cat, 1, bias=False)
    bn('head.mmcap.multi_scale.concat_conv.1', L_)
    for g in range(spec.n_groups):
        c = f'head.mmcap.mmcap.{g}.'
        a = c + 'attention.0.'
        ln(a + 'norm2', L_)
        lin(a + 'attn.proj', L_, spec.ca_dim)
        lin(a + 'attn.q', spec.ca_dim, L_)
        lin(a + 'attn.k', spec.ca_dim, L_)
        lin(a + 'attn.v', spec.ca_dim, L_)
        conv(a + 'mlp.fc1', L_ * spec.mlp_ratio, L_ // spec.mlp_groups, 1)
        conv(a + 'mlp.fc2', L_, L_ * spec.mlp_ratio // spec.mlp_groups, 1)
        ln(a + 'norm1', L_)
        t = c + 'gram_token_extraction.'
        S[t + 'bp_index'] = ((spec.tri,), 'idx')
        conv(t + 'ch_reduction.0', spec.bp_dim, L_, 1, bias=False)
        bn(t + 'ch_reduction.1', spec.bp_dim)
        conv(t + 'bp_reduction.0', L_ * T, spec.tri // spec.gram_group, 1, bias=False)
        bn(t + 'bp_reduction.1', L_ * T)
    for g in range(spec.n_groups):
        ln(f'head.heads.{g}.norm', L_ * T)
        lin(f'head.heads.{g}.head', spec.num_classes, L_ * T)
    for g in range(spec.n_groups):
        ln(f'head.self_dt_heads.{g}.norm', L_)
        lin(f'head.self_dt_heads.{g}.head', spec.num_classes, L_)
    return S


def make_state(spec: MAPSpec, seed: int = 0, profile: str = 'sensitised') -> State:
    """Sensitised deterministic state (see ga_convnext_oracle.make_state); profile 'trained' draws the ConvNeXt-block layer
    scales (`*.gamma`) from U(0.1, 0.3) instead of U(0.5, 1.5)."""
    P: State = {}
    for name, (shape, kind) in state_shapes(spec).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        if kind == 'w':
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            P[name] = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        elif kind == 'b':
            P[name] = torch.randn(shape, generator=g) * 0.1
        elif kind in ('g', 'rv'):
            P[name] = 0.5 + torch.rand(shape, generator=g)
            if profile == 'trained' and name.endswith('.gamma'):
                P[name] = P[name] * 0.2
        elif kind == 'rm':
            P[name] = torch.randn(shape, generator=g) * 0.1
        elif kind == 'idx':
            r, q = torch.triu_indices(spec.bp_dim, spec.bp_dim)
            P[name] = r * spec.bp_dim + q
        else:
            P[name] = torch.zeros(shape, dtype=torch.long)
    return P


# ----------------------------------------------------------------------------- backbone
def block(P: State, pre: str, x: Tensor) -> Tensor:
    """map_convnext.Block.forward (:27-40): gamma is applied in NHWC before the permute; same maths as GA's block."""
    C = x.shape[1]
    y = F.conv2d(x, P[pre + 'dwconv.weight'], P[pre + 'dwconv.bias'], padding=3, groups=C)
    y = F.layer_norm(y.permute(0, 2, 3, 1), (C,), P[pre + 'norm.weight'], P[pre + 'norm.bias'], 1e-6)
    y = F.linear(F.gelu(F.linear(y, P[pre + 'pwconv1.weight'], P[pre + 'pwconv1.bias'])), P[pre + 'pwconv2.weight'],
                 P[pre + 'pwconv2.bias'])
    return x + (P[pre + 'gamma'] * y).permute(0, 3, 1, 2)


def features(P: State, spec: MAPSpec, x: Tensor) -> List[Tensor]:
    """ConvNeXt.forward_features (:124-135): [stem output, stage 0..3 outputs]."""
    feats = []
    for i in range(4):
        pre = f'downsample_layers.{i}.'
        if i == 0:
            x = F.conv2d(x, P[pre + '0.weight'], P[pre + '0.bias'], stride=4)
            x = layernorm2d(x, P[pre + '1.weight'], P[pre + '1.bias'])
            feats.append(x)
        else:
            x = layernorm2d(x, P[pre + '0.weight'], P[pre + '0.bias'])
            x = F.conv2d(x, P[pre + '1.weight'], P[pre + '1.bias'], stride=2)
        for j in range(spec.depths[i]):
            x = block(P, f'stages.{i}.{j}.', x)
        feats.append(x)
    return feats


# ----------------------------------------------------------------------------- MAP head
def multi_scale(P: State, feats: List[Tensor], training: bool, level: int = 3) -> Tensor:
    """MultiScale.forward (map.py:322-333): smaller maps are enlarged by adaptive_avg_pool2d (replication), larger
    maps shrunk by NON-antialiased bilinear interpolation; concat; 1x1 conv + BN + GELU."""
    H, W = feats[level].shape[2:]
    out = []
    for f in feats:
        if H > f.shape[2]:
            f = F.adaptive_avg_pool2d(f, (H, W))
        elif H < f.shape[2]:
            f = F.interpolate(f, size=(H, W), mode='bilinear')
        out.append(f)
    pre = 'head.mmcap.multi_scale.concat_conv.'
    y = F.conv2d(torch.cat(out, 1), P[pre + '0.weight'])
    return F.gelu(batchnorm(P, pre + '1', y, training))


def gram_tokens(P: State, spec: MAPSpec, pre: str, x: Tensor, training: bool) -> Tensor:
    """GramToken.forward (map.py:210-234) -> [B, n_tokens, last_dim]."""
    x = batchnorm(P, pre + 'ch_reduction.1', F.conv2d(x, P[pre + 'ch_reduction.0.weight']), training)
    b, c, h, w = x.shape
    x = x.reshape(b, c, h * w) / (h * w)
    a = (x @ x.transpose(-1, -2)).reshape(b, c * c)[:, P[pre + 'bp_index']]
    a = F.normalize(a, dim=-1)
    a = a.reshape(b, -1, spec.n_tokens, 1, 1).permute(0, 2, 1, 3, 4).reshape(b, spec.tri, 1, 1)   # token interleave
    t = F.conv2d(a, P[pre + 'bp_reduction.0.weight'], groups=spec.gram_group)
    t = batchnorm(P, pre + 'bp_reduction.1', t, training)
    return t.reshape(b, spec.last_dim, spec.n_tokens).permute(0, 2, 1)


def class_attention(P: State, spec: MAPSpec, pre: str, u: Tensor, n_q: int) -> Tensor:
    """ClassAttention.forward, equal-dim branch (map.py:117-144): queries = first n_q rows, keys/values = all rows."""
    B, N, _ = u.shape
    E, H = spec.ca_dim, spec.num_heads
    hd = E // H
    q = F.linear(u[:, :n_q], P[pre + 'q.weight'], P[pre + 'q.bias']).reshape(B, n_q, H, hd).permute(0, 2, 1, 3) * hd ** -0.5
    k = F.linear(u, P[pre + 'k.weight'], P[pre + 'k.bias']).reshape(B, N, H, hd).permute(0, 2, 1, 3)
    v = F.linear(u, P[pre + 'v.weight'], P[pre + 'v.bias']).reshape(B, N, H, hd).permute(0, 2, 1, 3)
    a = torch.softmax(q @ k.transpose(-2, -1), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, n_q, E)
    return F.linear(o, P[pre + 'proj.weight'], P[pre + 'proj.bias'])


def group_conv_mlp(P: State, pre: str, t: Tensor, groups: int, relu_mask=None) -> Tensor:
    """map.GroupConvMlp (map.py:56-66) on tokens [B, n, C]: grouped 1x1 -> ReLU -> channel shuffle -> grouped 1x1.
    relu_mask: optional 0/1 tensor [B, hidden, n, 1] replacing the ReLU decisions (mask-pinned gradient check, see
    ga_convnext_oracle._relu)."""
    B, n, C = t.shape
    x = t.permute(0, 2, 1).unsqueeze(-1)
    x = F.conv2d(x, P[pre + 'fc1.weight'], P[pre + 'fc1.bias'], groups=groups)
    x = F.relu(x) if relu_mask is None else x * relu_mask.to(x.dtype)
    hid = x.shape[1]
    x = x.reshape(B, hid // groups, groups, n, 1).permute(0, 2, 1, 3, 4).reshape(B, hid, n, 1)
    x = F.conv2d(x, P[pre + 'fc2.weight'], P[pre + 'fc2.bias'], groups=groups)
    return x.squeeze(-1).permute(0, 2, 1)


def cap(P: State, spec: MAPSpec, g: int, x: Tensor, training: bool, relu_masks=None) -> Tensor:
    """CAP.forward (map.py:262-278): gram tokens + their mean token -> CABlock -> [B, all_tokens*last_dim]."""
    pre = f'head.mmcap.mmcap.{g}.'
    cls = gram_tokens(P, spec, pre + 'gram_token_extraction.', x, training)
    B, C, H, W = x.shape
    tok = x.reshape(B, C, H * W).permute(0, 2, 1)
    cls = torch.cat([cls, cls.mean(dim=1, keepdim=True)], dim=1)
    a = pre + 'attention.0.'
    u = torch.cat((cls, tok), dim=1)
    u = F.layer_norm(u, (C,), P[a + 'norm1.weight'], P[a + 'norm1.bias'], 1e-6)
    cls = cls + class_attention(P, spec, a + 'attn.', u, spec.all_tokens)
    h = F.layer_norm(cls, (C,), P[a + 'norm2.weight'], P[a + 'norm2.bias'], 1e-6)
    cls = cls + group_conv_mlp(P, a + 'mlp.', h, spec.mlp_groups, None if relu_masks is None else relu_masks.get(f'mlp{g}'))
    return cls.reshape(B, -1)


def norm_head(P: State, pre: str, x: Tensor) -> Tensor:
    """NormHead.forward (map.py:402-412): LayerNorm (eps 1e-5) -> Linear."""
    x = F.layer_norm(x, (x.shape[-1],), P[pre + 'norm.weight'], P[pre + 'norm.bias'], 1e-5)
    return F.linear(x, P[pre + 'head.weight'], P[pre + 'head.bias'])


def forward(P: State, spec: MAPSpec, x: Tensor, training: bool = False, relu_masks=None):
    """map_convnext ConvNeXt.forward with global_pool='mmcap' (:137-140) + MAPHead.forward (map.py:512-539).
    eval: list of n_groups logits; train: list of [main logits, self-distillation logits] pairs."""
    f = multi_scale(P, features(P, spec, x), training)
    out = []
    w = spec.last_dim * spec.n_tokens
    for g in range(spec.n_groups):
        pool = cap(P, spec, g, f, training, relu_masks)
        main = norm_head(P, f'head.heads.{g}.', pool[:, :w])
        if training:
            out.append([main, norm_head(P, f'head.self_dt_heads.{g}.', pool[:, w:])])
        else:
            out.append(main)
    return out


def map_loss(outputs, target, dec_lam: float = 0.0, loss_fn=F.cross_entropy) -> Tensor:
    """multi_group_loss, distill_tokens == 0 branch (MAP/train.py:792-839):
    sum_g [ L(y_g, t) + KL_sum(logsm(ymean_g) || logsm(y_g).detach()) / numel ] + dec_lam * sum_g KL_mean(logsm(y_g) || logsm(mean_g y).detach())."""
    loss = 0
    mains = []
    for o in outputs:
        if isinstance(o, (list, tuple)):
            y_hat, y_mean = o
            loss = loss + loss_fn(y_hat, target) + F.kl_div(F.log_softmax(y_mean, dim=1), F.log_softmax(y_hat, dim=1).detach(),
                                                            reduction='sum', log_target=True) / y_hat.numel()
        else:
            y_hat = o
            loss = loss + loss_fn(y_hat, target)
        mains.append(y_hat)
    if len(outputs) > 1:
        ref = F.log_softmax(sum(m.detach() for m in mains) / len(mains), dim=1)
        for m in mains:
            loss = loss + F.kl_div(F.log_softmax(m, dim=1), ref, reduction='mean', log_target=True) * dec_lam
    return loss
