"""CPU oracle for the GA-ConvNeXt hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (state-dict driven, module-free) fp32/fp64 restatement in plain PyTorch of what
`/root/reference/GA/ga_convnext.py` computes.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product package never does.

Parity status: PINNED.  `tests/golden/make_golden.py` runs the unmodified reference modules (imported
from /root/reference through `oracle/timm_shim`) on the same state dict and inputs, asserts this file
agrees with them, and commits the reference's outputs as fixtures under `tests/golden/`.

Every function cites the reference lines it follows.  Tensors are NCHW like the reference sees them.
`P` is a flat `{state_dict key: tensor}` mapping using the reference's own key names, which is the
on-disk contract.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]


@dataclass(frozen=True)
class GASpec:
    """Constructor arguments of one registered GA-ConvNeXt variant (ga_convnext.py:572-613)."""
    depths: Tuple[int, ...]
    dims: Tuple[int, ...]
    dim_embed: int
    naggre: int
    gram_dim: int = 192
    branches: int = 5
    embed_groups: int = 8
    heads: int = 8
    mlp_groups: int = 4
    num_classes: int = 1000
    se_width: Optional[int] = None  # derived

    @property
    def cat_width(self) -> int:  # ga_convnext.py:372
        return sum(self.dims[:-1]) + self.dims[2] * self.naggre

    @property
    def tri(self) -> int:  # ga_convnext.py:418
        return (self.gram_dim + 1) * self.gram_dim // 2


SPECS: Dict[str, GASpec] = {
    'ga_convnext_tiny_688': GASpec((3, 3, 9, 3, 1), (96, 192, 384, 688, 688), 168, 2),
    'ga_convnext_tiny_768': GASpec((3, 3, 9, 3, 1), (96, 192, 384, 768, 768), 192, 2),
    'ga_convnext_small_688': GASpec((3, 3, 27, 3, 1), (96, 192, 384, 688, 688), 168, 4),
    'ga_convnext_small_768': GASpec((3, 3, 27, 3, 1), (96, 192, 384, 768, 768), 192, 4),
    'ga_convnext_base_976': GASpec((3, 3, 27, 3, 1), (128, 256, 512, 976, 976), 240, 4),
    'ga_convnext_base_1024': GASpec((3, 3, 27, 3, 1), (128, 256, 512, 1024, 1024), 256, 4),
}


def se_reduced(width: int) -> int:
    """timm make_divisible(width/4, 8, round_limit=0) as used by create_attn('se', rd_ratio=1/4)."""
    return max(8, int(width * 0.25 + 4) // 8 * 8)


# ----------------------------------------------------------------------------- state layout

def state_shapes(spec: GASpec) -> Dict[str, Tuple[Tuple[int, ...], str]]:
    """Every state_dict entry of GA_ConvNeXt(spec) -> (shape, kind).  kind drives make_state()."""
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
        S[name + '.weight'] = ((c,), 'g')
        S[name + '.bias'] = ((c,), 'b')
        S[name + '.running_mean'] = ((c,), 'rm')
        S[name + '.running_var'] = ((c,), 'rv')
        S[name + '.num_batches_tracked'] = ((), 'n')

    def block(pre, c):
        conv(pre + 'conv_dw', c, 1, 7)
        ln(pre + 'norm', c)
        lin(pre + 'mlp.fc1', 4 * c, c)
        lin(pre + 'mlp.fc2', c, 4 * c)
        S[pre + 'gamma'] = ((c,), 'g')

    d = spec.dims
    conv('stem.0', d[0], 3, 4)
    ln('stem.1', d[0])
    prev = d[0]
    for i in range(4):
        if i > 0:
            ln(f'stages.{i}.downsample.0', prev)
            conv(f'stages.{i}.downsample.1', d[i], prev, 2)
        for j in range(spec.depths[i]):
            block(f'stages.{i}.blocks.{j}.', d[i])
        prev = d[i]
    # stage 4 = Bottleneck (ga_convnext.py:376, 251-289)
    cin, cout = spec.cat_width, d[4]
    w = cout // 4
    conv('stages.4.downsample.0', cout, cin, 1)
    bn('stages.4.downsample.1', cout)
    conv('stages.4.conv1', w, cin, 1, bias=False)
    bn('stages.4.bn1', w)
    conv('stages.4.conv2', w, w, 3, bias=False)
    bn('stages.4.bn2', w)
    conv('stages.4.se.fc1', se_reduced(w), w, 1)
    conv('stages.4.se.fc2', w, se_reduced(w), 1)
    conv('stages.4.conv3', cout, w, 1, bias=False)
    bn('stages.4.bn3', cout)
    g = spec.gram_dim
    for k in range(spec.branches):
        conv(f'gram_contraction.{k}.0', g, cout, 1)
        bn(f'gram_contraction.{k}.1', g)
        block(f'gram_layer.{k}.blocks.0.', g)
        conv(f'gram_embedding.{k}.0', cout, spec.tri // spec.embed_groups, 1)
        bn(f'gram_embedding.{k}.1', cout)
        a = f'ga.{k}.'
        ln(a + 'norm1', cout)
        lin(a + 'attn.q', spec.dim_embed, cout, bias=False)
        lin(a + 'attn.k', spec.dim_embed, cout, bias=False)
        lin(a + 'attn.v', spec.dim_embed, cout, bias=False)
        lin(a + 'attn.proj', cout, spec.dim_embed)
        ln(a + 'norm2', cout)
        conv(a + 'mlp.fc1', 4 * cout, cout // spec.mlp_groups, 1)
        conv(a + 'mlp.fc2', cout, 4 * cout // spec.mlp_groups, 1)
        S[a + 'gamma_1'] = ((cout,), 'g')
        S[a + 'gamma_2'] = ((cout,), 'g')
        lin(f'fc.{k}', spec.num_classes, cout)
    return S


def make_state(spec: GASpec, seed: int = 0, dtype=torch.float32, profile: str = 'sensitised') -> State:
    """A *sensitised* deterministic state dict: O(1) layer scales, random BN statistics.

    profile 'trained': the ConvNeXt-block layer scales (`*.gamma`) are drawn from U(0.1, 0.3) instead of U(0.5, 1.5), the
    magnitude they have in a trained network; 18 residual blocks then no longer multiply every rounding error by O(1)
    per block, which is what the bf16 fixtures need to be well conditioned.  Everything else is identical.

    The reference's default init (layer-scale 1e-6, GA gamma 1e-4; ga_convnext.py:95,241-242) makes
    the logits insensitive to almost every kernel, so parity runs use this instead.
    Each tensor has its own generator seeded from (seed, crc32(name)) -> independent of key order.
    """
    P: State = {}
    for name, (shape, kind) in state_shapes(spec).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        if kind == 'w':
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        elif kind == 'b':
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == 'g':
            t = 0.5 + torch.rand(shape, generator=g)
            if profile == 'trained' and name.endswith('.gamma'):
                t = t * 0.2
        elif kind == 'rm':
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == 'rv':
            t = 0.5 + torch.rand(shape, generator=g)
        else:
            t = torch.zeros(shape, dtype=torch.long)
        P[name] = t if kind == 'n' else t.to(dtype)
    return P


# ----------------------------------------------------------------------------- building blocks

def layernorm2d(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """LayerNorm over C of an NCHW tensor (ga_convnext.py:51-67; both branches are the same maths)."""
    var, mean = torch.var_mean(x, dim=1, unbiased=False, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * w[None, :, None, None] + b[None, :, None, None]


def convnext_block(P: State, pre: str, x: Tensor, path_scale: Optional[Tensor] = None) -> Tensor:
    """dw7x7 -> LN -> fc1 -> GELU(erf) -> fc2 -> *gamma -> drop-path -> +x   (ga_convnext.py:98-112).

    path_scale: optional per-sample multiplier [B] standing in for DropPath's bernoulli/keep mask.
    """
    C = x.shape[1]
    y = F.conv2d(x, P[pre + 'conv_dw.weight'], P[pre + 'conv_dw.bias'], padding=3, groups=C)
    y = F.layer_norm(y.permute(0, 2, 3, 1), (C,), P[pre + 'norm.weight'], P[pre + 'norm.bias'], 1e-6)
    y = F.gelu(F.linear(y, P[pre + 'mlp.fc1.weight'], P[pre + 'mlp.fc1.bias']))
    y = F.linear(y, P[pre + 'mlp.fc2.weight'], P[pre + 'mlp.fc2.bias'])
    y = y.permute(0, 3, 1, 2) * P[pre + 'gamma'].reshape(1, -1, 1, 1)
    if path_scale is not None:
        y = y * path_scale.reshape(-1, 1, 1, 1)
    return y + x


def batchnorm(P: State, pre: str, x: Tensor, training: bool, momentum: float = 0.1,
              eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm2d semantics; updates P[pre+running_*] in place when training."""
    if training:
        P[pre + '.num_batches_tracked'] += 1
    return F.batch_norm(x, P[pre + '.running_mean'], P[pre + '.running_var'], P[pre + '.weight'],
                        P[pre + '.bias'], training, momentum, eps)


def stage(P: State, spec: GASpec, i: int, x: Tensor) -> Tuple[Tensor, List[Tensor]]:
    """ConvNeXtStage.forward incl. the tap rule (ga_convnext.py:139-150)."""
    pre = f'stages.{i}.'
    if i > 0:
        x = layernorm2d(x, P[pre + 'downsample.0.weight'], P[pre + 'downsample.0.bias'])
        x = F.conv2d(x, P[pre + 'downsample.1.weight'], P[pre + 'downsample.1.bias'], stride=2)
    depth = spec.depths[i]
    taps: List[Tensor] = []
    for j in range(depth):
        x = convnext_block(P, f'{pre}blocks.{j}.', x)
        if depth > 5 and (j + 1) % (depth // (spec.naggre + 1)) == 0 and len(taps) < spec.naggre:
            taps.append(x)
    return x, taps


def _relu(z: Tensor, masks, key: str) -> Tensor:
    """ReLU, or -- for the mask-pinned gradient check of tests/ -- multiplication by a given 0/1 decision tensor.

    A ReLU network's gradient is discontinuous in its pre-activations: the reference run twice with inputs that differ by
    1e-7 relative flips one of the 2.2 M decisions of the last Bottleneck ReLU and its own upstream gradients move by 1e-3
    (measured, tests/golden/make_golden.py).  Comparing gradients of two implementations therefore needs the same decisions
    on both sides; `masks[key]` carries the implementation's, and masks=None is the reference's behaviour."""
    if masks is None or key not in masks:
        return F.relu(z)
    return z * masks[key].to(z.dtype)


def bottleneck(P: State, x: Tensor, training: bool, relu_masks=None) -> Tensor:
    """stages.4: 1x1+BN+ReLU, 3x3+BN+ReLU, SE, 1x1+BN, + (1x1+BN shortcut), ReLU (ga_convnext.py:294-318).

    DropPath on the residual branch (:310-311) is identity in eval / at rate 0, the parity contract.
    relu_masks: optional {'bn1', 'bn2', 'out', 'se'} 0/1 tensors (NCHW; 'se' [B, R, 1, 1]) replacing the ReLU decisions (see _relu).
    """
    pre = 'stages.4.'
    y = _relu(batchnorm(P, pre + 'bn1', F.conv2d(x, P[pre + 'conv1.weight']), training), relu_masks, 'bn1')
    y = _relu(batchnorm(P, pre + 'bn2', F.conv2d(y, P[pre + 'conv2.weight'], padding=1), training), relu_masks, 'bn2')
    # timm SEModule: mean over HW -> fc1 -> ReLU -> fc2 -> sigmoid gate
    s = y.mean((2, 3), keepdim=True)
    s = _relu(F.conv2d(s, P[pre + 'se.fc1.weight'], P[pre + 'se.fc1.bias']), relu_masks, 'se')
    s = F.conv2d(s, P[pre + 'se.fc2.weight'], P[pre + 'se.fc2.bias'])
    y = y * torch.sigmoid(s)
    y = batchnorm(P, pre + 'bn3', F.conv2d(y, P[pre + 'conv3.weight']), training)
    sc = F.conv2d(x, P[pre + 'downsample.0.weight'], P[pre + 'downsample.0.bias'])
    sc = batchnorm(P, pre + 'downsample.1', sc, training)
    return _relu(y + sc, relu_masks, 'out')


def triu_index(c: int) -> Tensor:
    """Row-major i<=j gather list (ga_convnext.py:424-430)."""
    r, q = torch.triu_indices(c, c)
    return r * c + q


def gram_vector(x: Tensor, training: bool) -> Tensor:
    """get_gram (ga_convnext.py:452-467): X/H, [fp64 if training and B<128], X Xt / HW, triu, L2 norm."""
    B, C, H, W = x.shape
    x = x / H
    if training and B < 128:
        x = x.to(torch.float64)
    x = x.reshape(B, C, H * W)
    g = torch.bmm(x, x.transpose(1, 2)) / (H * W)
    g = g.reshape(B, C * C)[:, triu_index(C)]
    g = F.normalize(g)  # L2 over dim 1, eps 1e-12
    return g.float().reshape(B, -1, 1, 1)


def channel_shuffle(x: Tensor, groups: int) -> Tensor:
    """ga_convnext.py:557-566 on a [B, C] token: view C as (C/g, g) and transpose."""
    B, C = x.shape
    return x.reshape(B, C // groups, groups).transpose(1, 2).reshape(B, C)


def class_attention(P: State, pre: str, u: Tensor, heads: int) -> Tensor:
    """ClassAttn.forward (ga_convnext.py:170-187): one query (token 0) against all N tokens."""
    B, N, _ = u.shape
    E = P[pre + 'q.weight'].shape[0]
    hd = E // heads
    q = F.linear(u[:, 0], P[pre + 'q.weight']).reshape(B, heads, 1, hd) * hd ** -0.5
    k = F.linear(u, P[pre + 'k.weight']).reshape(B, N, heads, hd).permute(0, 2, 1, 3)
    v = F.linear(u, P[pre + 'v.weight']).reshape(B, N, heads, hd).permute(0, 2, 1, 3)
    a = torch.softmax(q @ k.transpose(-2, -1), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, E)
    return F.linear(o, P[pre + 'proj.weight'], P[pre + 'proj.bias'])


def group_conv_mlp(P: State, pre: str, t: Tensor, groups: int) -> Tensor:
    """GroupConvMlp on one token [B, C] (ga_convnext.py:208-222); act is GELU (passed at :239)."""
    h = F.conv2d(t[:, :, None, None], P[pre + 'fc1.weight'], P[pre + 'fc1.bias'], groups=groups)
    h = F.gelu(h)[:, :, 0, 0]
    h = channel_shuffle(h, groups)
    return F.conv2d(h[:, :, None, None], P[pre + 'fc2.weight'], P[pre + 'fc2.bias'], groups=groups)[:, :, 0, 0]


def ga_block(P: State, spec: GASpec, k: int, tokens: Tensor, cls: Tensor) -> Tensor:
    """LayerScaleBlockClassAttn.forward (ga_convnext.py:244-248). tokens [B,N,C], cls [B,C]."""
    pre = f'ga.{k}.'
    C = cls.shape[1]
    u = torch.cat((cls[:, None, :], tokens), dim=1)
    u = F.layer_norm(u, (C,), P[pre + 'norm1.weight'], P[pre + 'norm1.bias'], 1e-5)
    cls = cls + P[pre + 'gamma_1'] * class_attention(P, pre + 'attn.', u, spec.heads)
    h = F.layer_norm(cls, (C,), P[pre + 'norm2.weight'], P[pre + 'norm2.bias'], 1e-5)
    return cls + P[pre + 'gamma_2'] * group_conv_mlp(P, pre + 'mlp.', h, spec.mlp_groups)


def aggregate(spec: GASpec, feats: List[Tensor], taps: List[Tensor], pool: int = 14) -> Tensor:
    """forward_features' multi-scale concat (ga_convnext.py:479-483): pool-to-14, taps, x2, bilinear x2.
    pool: the reference hard-codes AdaptiveAvgPool2d(14) (:397), which is H/16 at 224; other input sizes only run with the pool
    target set to H/16 (make_golden.py patches the module ATTRIBUTE of the unmodified reference for the 384 fixture)."""
    x0, x1, x2, x3 = feats
    return torch.cat((F.adaptive_avg_pool2d(x0, pool), F.adaptive_avg_pool2d(x1, pool), *taps, x2,
                      F.interpolate(x3, scale_factor=2, mode='bilinear')), dim=1)


def forward_features(P: State, spec: GASpec, x: Tensor, training: bool, relu_masks=None) -> Tensor:
    """stem -> 4 stages -> aggregation -> Bottleneck (ga_convnext.py:469-485)."""
    x = F.conv2d(x, P['stem.0.weight'], P['stem.0.bias'], stride=4)
    x = layernorm2d(x, P['stem.1.weight'], P['stem.1.bias'])
    feats, taps = [], []
    for i in range(4):
        x, t = stage(P, spec, i, x)
        feats.append(x)
        taps += t
    return bottleneck(P, aggregate(spec, feats, taps, pool=feats[2].shape[-1]), training, relu_masks)


def branch(P: State, spec: GASpec, k: int, f: Tensor, training: bool) -> Tensor:
    """One GA branch (ga_convnext.py:491-504) -> logits [B, num_classes]."""
    B, C, H, W = f.shape
    g = F.conv2d(f, P[f'gram_contraction.{k}.0.weight'], P[f'gram_contraction.{k}.0.bias'])
    g = batchnorm(P, f'gram_contraction.{k}.1', g, training)
    g = convnext_block(P, f'gram_layer.{k}.blocks.0.', g)
    g = gram_vector(g, training)
    c = F.conv2d(g, P[f'gram_embedding.{k}.0.weight'], P[f'gram_embedding.{k}.0.bias'], groups=spec.embed_groups)
    c = batchnorm(P, f'gram_embedding.{k}.1', c, training)[:, :, 0, 0]
    tokens = f.reshape(B, C, H * W).permute(0, 2, 1)
    c = ga_block(P, spec, k, tokens, c)
    return F.linear(c, P[f'fc.{k}.weight'], P[f'fc.{k}.bias'])


def forward(P: State, spec: GASpec, x: Tensor, training: bool = False, relu_masks=None) -> List[Tensor]:
    """GA_ConvNeXt.forward (ga_convnext.py:487-505): list of `branches` logits tensors."""
    f = forward_features(P, spec, x, training, relu_masks)
    return [branch(P, spec, k, f, training) for k in range(spec.branches)]


# ----------------------------------------------------------------------------- loss / metrics

def ga_loss(outputs: List[Tensor], target: Tensor, lam: float = 0.0, loss_fn=F.cross_entropy) -> Tensor:
    """GA/train.py:735-745: sum_k L(out_k, y) + lam * sum_k KL_mean(logsm(out_k) || logsm(mean_k out).detach())."""
    loss = sum(loss_fn(o, target) for o in outputs)
    mean = sum(o.detach() for o in outputs) / len(outputs)
    ref = F.log_softmax(mean, dim=-1)
    for o in outputs:
        loss = loss + F.kl_div(F.log_softmax(o, dim=-1), ref, reduction='mean', log_target=True) * lam
    return loss


def topk_correct(logits: Tensor, target: Tensor, ks=(1, 5)) -> List[Tensor]:
    """timm.utils.accuracy as used at GA/train.py:859: percentage of rows whose target is in the top-k."""
    _, pred = logits.topk(max(ks), 1, True, True)
    hit = pred.t().eq(target.reshape(1, -1))
    return [hit[:k].reshape(-1).float().sum() * 100.0 / target.shape[0] for k in ks]
