"""CPU oracle for the GA-CSWin hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Functional (state-dict driven) fp32 restatement in plain PyTorch of `/root/reference/GA/ga_cswin.py`
(SURVEY.md section 8 rows a14-a18).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may
import it.

Parity status: PINNED.  `tests/golden/make_golden.py cswin` builds the unmodified reference
`GA_CSWinTransformer` (imported through `oracle/timm_shim`), loads `make_state(...)` with strict=True, asserts
this file agrees with it on forward, loss and every parameter gradient, and commits the reference's outputs
under `tests/golden/ga_cswin_*.pt`.

The reference file registers no factory (its `default_cfgs` names `ga_CSWin_64_12211_tiny_224` only); the
constructor arguments in SPECS are the ones SURVEY.md section 8 row a18 derives from GA/README.md's parameter
count (43.4 M vs 42.0 M published) -- an assumption, stated as such in DESIGN.md.

Tokens are [B, L, C] like the reference sees them; `P` uses the reference's state_dict key names.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from .ga_convnext_oracle import (State, Tensor, batchnorm, class_attention, ga_loss, group_conv_mlp,  # noqa: F401
                                 topk_correct, triu_index)


@dataclass(frozen=True)
class CSWinSpec:
    """Constructor arguments of GA_CSWinTransformer (ga_cswin.py:448-453)."""
    depth: Tuple[int, ...] = (1, 2, 21, 1)
    split_size: Tuple[int, ...] = (1, 2, 7, 7, 7)
    num_heads: Tuple[int, ...] = (2, 4, 8, 16, 16)
    dims: Tuple[int, ...] = (64, 128, 256, 512)
    embed_dim: int = 64
    img_size: int = 224
    naggre: int = 4
    gram_dim: int = 192
    gram_heads: int = 6          # hard-coded at ga_cswin.py:568
    branches: int = 5
    groups: int = 8              # contraction / embedding conv groups (ga_cswin.py:560, 586)
    heads: int = 8               # class-attention heads (:589)
    expansion: int = 4           # ClassAttn width reduction (:273)
    mlp_groups: int = 2          # ga_mlp_groups (:451)
    mlp_ratio: float = 4.
    num_classes: int = 1000

    @property
    def cat_width(self) -> int:  # ga_cswin.py:528
        return sum(self.dims) + self.dims[2] * self.naggre

    @property
    def tri(self) -> int:
        return (self.gram_dim + 1) * self.gram_dim // 2


SPECS: Dict[str, CSWinSpec] = {
    'ga_CSWin_64_12211_tiny_224': CSWinSpec(),
    # reduced widths / depth for fast tests: same 224 geometry and code paths, head_dim 32 everywhere.  The reference
    # hard-codes 6 heads for gram_layer (:568), so gram_dim stays 192.
    'ga_cswin_test': CSWinSpec(depth=(1, 1, 5, 1), num_heads=(2, 2, 4, 4, 4), dims=(64, 64, 128, 128), embed_dim=32,
                               num_classes=100),
}
PARAM_COUNTS = {'ga_CSWin_64_12211_tiny_224': 43431816}   # 43.43 M (README: 42.0 M)


def block_branches(reso: int, split: int, last_stage: bool = False) -> int:
    """ga_cswin.py:153-158."""
    return 1 if (last_stage or reso == split) else 2


# ----------------------------------------------------------------------------- state layout

def block_shapes(S, pre: str, dim: int, reso: int, split: int, last_stage: bool, mlp_ratio: float = 4.):
    def lin(name, cout, cin):
        S[name + '.weight'] = ((cout, cin), 'w')
        S[name + '.bias'] = ((cout,), 'b')

    def ln(name, c):
        S[name + '.weight'] = ((c,), 'g')
        S[name + '.bias'] = ((c,), 'b')
    lin(pre + 'qkv', 3 * dim, dim)
    ln(pre + 'norm1', dim)
    lin(pre + 'proj', dim, dim)
    nb = block_branches(reso, split, last_stage)
    for i in range(nb):
        S[f'{pre}attns.{i}.get_v.weight'] = ((dim // nb, 1, 3, 3), 'w')
        S[f'{pre}attns.{i}.get_v.bias'] = ((dim // nb,), 'b')
    hid = int(dim * mlp_ratio)
    lin(pre + 'mlp.fc1', hid, dim)
    lin(pre + 'mlp.fc2', dim, hid)
    ln(pre + 'norm2', dim)


def state_shapes(spec: CSWinSpec) -> Dict[str, Tuple[Tuple[int, ...], str]]:
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

    d, e, r = spec.dims, spec.embed_dim, spec.img_size
    # deep stem (ga_cswin.py:462-477); Sequential indices skip the parameter-free Rearrange / GELU entries
    conv('stage1_conv_embed.0', e, 3, 3, bias=False)
    ln('stage1_conv_embed.2', e)
    conv('stage1_conv_embed.5', e, e, 3, bias=False)
    ln('stage1_conv_embed.7', e)
    conv('stage1_conv_embed.10', d[0], e, 3, bias=False)
    ln('stage1_conv_embed.12', d[0])
    for s in range(4):
        reso = r // (4 << s)
        for j in range(spec.depth[s]):
            block_shapes(S, f'stage{s + 1}.{j}.', d[s], reso, spec.split_size[s] if s < 3 else spec.split_size[-1],
                         last_stage=(s == 3), mlp_ratio=spec.mlp_ratio)
        if s < 3:
            conv(f'merge{s + 1}.conv', d[s + 1], d[s], 3)
            ln(f'merge{s + 1}.norm', d[s + 1])
    c = d[3]
    conv('stage5.1.conv', c, spec.cat_width, 1)
    ln('stage5.1.norm', c)
    block_shapes(S, 'stage5.2.', c, r // 16, spec.split_size[4], False, spec.mlp_ratio)
    g = spec.gram_dim
    for k in range(spec.branches):
        conv(f'gram_contraction.{k}.0', g, c // spec.groups, 1)
        bn(f'gram_contraction.{k}.1', g)
        block_shapes(S, f'gram_layer.{k}.1.', g, r // 16, spec.split_size[4], False)
        conv(f'gram_embedding.{k}.0', c, spec.tri // spec.groups, 1)
        bn(f'gram_embedding.{k}.1', c)
        a = f'ga.{k}.'
        E = c // spec.expansion
        ln(a + 'norm1', c)
        lin(a + 'attn.q', E, c, bias=False)
        lin(a + 'attn.k', E, c, bias=False)
        lin(a + 'attn.v', E, c, bias=False)
        lin(a + 'attn.proj', c, E)
        ln(a + 'norm2', c)
        conv(a + 'mlp.fc1', 4 * c, c // spec.mlp_groups, 1)
        conv(a + 'mlp.fc2', c, 4 * c // spec.mlp_groups, 1)
        S[a + 'gamma_1'] = ((c,), 'g')
        S[a + 'gamma_2'] = ((c,), 'g')
        lin(f'fc.{k}', spec.num_classes, c)
    return S


def make_state(spec: CSWinSpec, seed: int = 0, dtype=torch.float32, profile: str = 'sensitised') -> State:
    """Sensitised deterministic state (same recipe as ga_convnext_oracle.make_state).  CSWin blocks have no layer scale; profile
    'trained' scales the weights of every residual branch's last Linear (attention `proj`, `mlp.fc2`) by 0.3, the role the layer
    scale plays in the ConvNeXt fixtures: 31 unscaled residual branches otherwise let every rounding error through at full size."""
    import math
    import zlib
    P: State = {}
    for name, (shape, kind) in state_shapes(spec).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        if kind == 'w':
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
            if profile == 'trained' and (name.endswith('.proj.weight') or name.endswith('.mlp.fc2.weight')) and 'ga.' not in name:
                t = t * 0.3
        elif kind in ('b', 'rm'):
            t = torch.randn(shape, generator=g) * 0.1
        elif kind in ('g', 'rv'):
            t = 0.5 + torch.rand(shape, generator=g)
        else:
            t = torch.zeros(shape, dtype=torch.long)
        P[name] = t if kind == 'n' else t.to(dtype)
    return P


# ----------------------------------------------------------------------------- building blocks

def to_image(x: Tensor, H: int, W: int) -> Tensor:
    B, L, C = x.shape
    return x.transpose(1, 2).reshape(B, C, H, W)


def to_tokens(x: Tensor) -> Tensor:
    B, C, H, W = x.shape
    return x.reshape(B, C, H * W).transpose(1, 2)


def windows(x: Tensor, H: int, W: int, hs: int, ws: int) -> Tensor:
    """[B, H*W, C] -> [B * H/hs * W/ws, hs*ws, C]  (img2windows, ga_cswin.py:215-223)."""
    B, L, C = x.shape
    x = x.reshape(B, H // hs, hs, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, hs * ws, C)


def unwindows(x: Tensor, B: int, H: int, W: int, hs: int, ws: int) -> Tensor:
    """inverse of windows() (windows2img, ga_cswin.py:225-234) -> [B, H*W, C]."""
    C = x.shape[-1]
    x = x.reshape(B, H // hs, W // ws, hs, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B, H * W, C)


def lepe_attention(P: State, pre: str, q: Tensor, k: Tensor, v: Tensor, reso: int, hs: int, ws: int, heads: int) -> Tensor:
    """LePEAttention.forward (ga_cswin.py:105-136): per-stripe multi-head softmax attention plus a depthwise 3x3
    of V evaluated INSIDE each stripe (zero padding at stripe borders, get_lepe :92-103)."""
    B, L, C = q.shape
    hd = C // heads
    n = hs * ws

    def split(t):   # -> [B', heads, n, hd]
        return windows(t, reso, reso, hs, ws).reshape(-1, n, heads, hd).permute(0, 2, 1, 3)
    qw, kw, vw = split(q), split(k), split(v)
    vimg = windows(v, reso, reso, hs, ws).transpose(1, 2).reshape(-1, C, hs, ws)
    lepe = F.conv2d(vimg, P[pre + 'get_v.weight'], P[pre + 'get_v.bias'], padding=1, groups=C)
    lepe = lepe.reshape(-1, heads, hd, n).permute(0, 1, 3, 2)
    a = torch.softmax((qw * hd ** -0.5) @ kw.transpose(-2, -1), dim=-1)
    o = a @ vw + lepe
    o = o.transpose(1, 2).reshape(-1, n, C)
    return unwindows(o, B, reso, reso, hs, ws)


def cswin_block(P: State, pre: str, x: Tensor, reso: int, split: int, heads: int, last_stage: bool = False) -> Tensor:
    """CSWinBlock.forward (ga_cswin.py:187-212); drop rates 0."""
    B, L, C = x.shape
    img = F.layer_norm(x, (C,), P[pre + 'norm1.weight'], P[pre + 'norm1.bias'], 1e-5)
    qkv = F.linear(img, P[pre + 'qkv.weight'], P[pre + 'qkv.bias'])
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    if block_branches(reso, split, last_stage) == 2:
        h = C // 2
        a0 = lepe_attention(P, pre + 'attns.0.', q[..., :h], k[..., :h], v[..., :h], reso, reso, split, heads // 2)
        a1 = lepe_attention(P, pre + 'attns.1.', q[..., h:], k[..., h:], v[..., h:], reso, split, reso, heads // 2)
        att = torch.cat((a0, a1), dim=2)
    else:
        att = lepe_attention(P, pre + 'attns.0.', q, k, v, reso, reso, reso, heads)
    x = x + F.linear(att, P[pre + 'proj.weight'], P[pre + 'proj.bias'])
    h = F.layer_norm(x, (C,), P[pre + 'norm2.weight'], P[pre + 'norm2.bias'], 1e-5)
    h = F.gelu(F.linear(h, P[pre + 'mlp.fc1.weight'], P[pre + 'mlp.fc1.bias']))
    return x + F.linear(h, P[pre + 'mlp.fc2.weight'], P[pre + 'mlp.fc2.bias'])


def merge_block(P: State, pre: str, x: Tensor, reso: int, k: int) -> Tensor:
    """Merge_Block (3x3 s2 p1, ga_cswin.py:253-268) / Merge_Block_LCF (1x1, :236-251) + LayerNorm on tokens."""
    img = to_image(x, reso, reso)
    if k == 3:
        img = F.conv2d(img, P[pre + 'conv.weight'], P[pre + 'conv.bias'], stride=2, padding=1)
    else:
        img = F.conv2d(img, P[pre + 'conv.weight'], P[pre + 'conv.bias'])
    t = to_tokens(img)
    return F.layer_norm(t, (t.shape[-1],), P[pre + 'norm.weight'], P[pre + 'norm.bias'], 1e-5)


def deep_stem(P: State, x: Tensor) -> Tensor:
    """stage1_conv_embed (ga_cswin.py:462-477): 3x3 s2 -> LN -> GELU -> 3x3 -> LN -> GELU -> 3x3 s2 -> LN."""
    pre = 'stage1_conv_embed.'

    def ln_img(t, name, act):
        tok = to_tokens(t)
        tok = F.layer_norm(tok, (tok.shape[-1],), P[pre + name + '.weight'], P[pre + name + '.bias'], 1e-5)
        return F.gelu(to_image(tok, t.shape[2], t.shape[3])) if act else tok
    x = ln_img(F.conv2d(x, P[pre + '0.weight'], stride=2, padding=1), '2', True)
    x = ln_img(F.conv2d(x, P[pre + '5.weight'], stride=1, padding=1), '7', True)
    return ln_img(F.conv2d(x, P[pre + '10.weight'], stride=2, padding=1), '12', False)


def forward_features(P: State, spec: CSWinSpec, x: Tensor, training: bool = False) -> Tensor:
    """ga_cswin.py:636-671 -> [B, C, r/16, r/16]."""
    r = spec.img_size
    x = deep_stem(P, x)
    xs: List[Tensor] = []
    for j in range(spec.depth[0]):
        x = cswin_block(P, f'stage1.{j}.', x, r // 4, spec.split_size[0], spec.num_heads[0])
    xs.append(to_image(x, r // 4, r // 4))
    for s in (1, 2, 3):
        reso = r // (4 << s)
        x = merge_block(P, f'merge{s}.', x, reso * 2, 3)
        n = spec.depth[s]
        for j in range(n):
            x = cswin_block(P, f'stage{s + 1}.{j}.', x, reso, spec.split_size[s] if s < 3 else spec.split_size[-1],
                            spec.num_heads[s], last_stage=(s == 3))
            if s == 2 and (j + 1) % (n // (spec.naggre + 1)) == 0 and len(xs) < spec.naggre + 2:   # :659
                xs.append(to_image(x, reso, reso))
        xs.append(to_image(x, reso, reso))
    po = r // 16
    # the reference pools to 14 (:546) = img_size // 16 at 224
    cat = torch.cat((F.adaptive_avg_pool2d(xs[0], po), F.adaptive_avg_pool2d(xs[1], po), *xs[2:-1],
                     F.interpolate(xs[-1], scale_factor=2, mode='bilinear')), dim=1)
    t = merge_block(P, 'stage5.1.', to_tokens(cat), po, 1)
    t = cswin_block(P, 'stage5.2.', t, po, spec.split_size[4], spec.num_heads[4])
    return to_image(t, po, po)


def gram_vector(x: Tensor) -> Tensor:
    """get_gram (ga_cswin.py:624-634): no fp64 branch here, unlike ga_convnext."""
    B, C, H, W = x.shape
    x = (x / H).reshape(B, C, H * W)
    g = torch.bmm(x, x.transpose(1, 2)) / (H * W)
    g = g.reshape(B, C * C)[:, triu_index(C)]
    return F.normalize(g).float().reshape(B, -1, 1, 1)


def branch(P: State, spec: CSWinSpec, k: int, f: Tensor, training: bool) -> Tensor:
    """One GA branch (ga_cswin.py:676-692)."""
    B, C, H, W = f.shape
    g = F.conv2d(f, P[f'gram_contraction.{k}.0.weight'], P[f'gram_contraction.{k}.0.bias'], groups=spec.groups)
    g = batchnorm(P, f'gram_contraction.{k}.1', g, training)
    g = cswin_block(P, f'gram_layer.{k}.1.', to_tokens(g), H, spec.split_size[4], spec.gram_heads)
    g = gram_vector(to_image(g, H, W))
    c = F.conv2d(g, P[f'gram_embedding.{k}.0.weight'], P[f'gram_embedding.{k}.0.bias'], groups=spec.groups)
    c = batchnorm(P, f'gram_embedding.{k}.1', c, training)[:, :, 0, 0]
    pre = f'ga.{k}.'
    u = torch.cat((c[:, None, :], to_tokens(f)), dim=1)
    u = F.layer_norm(u, (C,), P[pre + 'norm1.weight'], P[pre + 'norm1.bias'], 1e-5)
    c = c + P[pre + 'gamma_1'] * class_attention(P, pre + 'attn.', u, spec.heads)
    h = F.layer_norm(c, (C,), P[pre + 'norm2.weight'], P[pre + 'norm2.bias'], 1e-5)
    c = c + P[pre + 'gamma_2'] * group_conv_mlp(P, pre + 'mlp.', h, spec.mlp_groups)
    return F.linear(c, P[f'fc.{k}.weight'], P[f'fc.{k}.bias'])


def forward(P: State, spec: CSWinSpec, x: Tensor, training: bool = False) -> List[Tensor]:
    f = forward_features(P, spec, x, training)
    return [branch(P, spec, k, f, training) for k in range(spec.branches)]
