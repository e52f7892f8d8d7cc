"""GA-ConvNeXt on sm_100a kernels: drop-in for `GA/ga_convnext.py` (same factories, attribute names, state_dict).

The module tree below only *holds* parameters (so `state_dict()`, `load_state_dict(strict=True)`, `deepcopy`,
`.cuda()`, DDP and optimizer factories see exactly the reference layout, 407 keys for tiny_688); the arithmetic
is done by `ops.*` on NHWC row matrices through libga_sm100.so.  fp32 inputs run the fp32 kernels; under
`torch.autocast(..., dtype=torch.bfloat16)` (or `model.compute_dtype = torch.bfloat16`) activations are bf16 with
fp32 accumulation.  There is no PyTorch/CPU fallback: calling forward on CPU tensors raises.

Reference lines are cited per method; construction order mirrors ga_convnext.py:335-432 so that
`torch.manual_seed(s); create_model(name)` yields bit-identical initial weights.
"""
from __future__ import annotations

import os
import warnings
from typing import List, Optional

import weakref

import torch
import torch.nn as nn

from . import ops
from .lib import ACT_GELU
from .registry import build_model_with_cfg, register_model

__all__ = ['GA_ConvNeXt']

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def _cfg(url='', **kwargs):
    return {'url': url, 'num_classes': 1000, 'input_size': (3, 224, 224), 'pool_size': (7, 7), 'crop_pct': 0.875,
            'interpolation': 'bicubic', 'mean': IMAGENET_DEFAULT_MEAN, 'std': IMAGENET_DEFAULT_STD,
            'first_conv': 'stem.0', 'classifier': 'head.fc', **kwargs}


default_cfgs = dict(ga_convnext_tiny=_cfg(), ga_convnext_small=_cfg(), ga_convnext_base=_cfg())


def _params(mod: nn.Module) -> dict:
    """Flat {relative name: tensor} of a holder module (parameters and buffers)."""
    d = dict(mod.named_parameters())
    d.update(dict(mod.named_buffers()))
    return d


_KEEP_CACHE = {}


def _path_scales(probs, training: bool, batch: int, device):
    """DropPath factors of several blocks from ONE uniform draw (was a bernoulli_ + div_ pair of launches per block):
    floor(u + keep) / keep per sample, None where the rate is 0 or the module is in eval mode."""
    probs = [float(p) for p in probs]
    live = [i for i, p in enumerate(probs) if p > 0. and training]
    out = [None] * len(probs)
    if live:
        key = (tuple(probs[i] for i in live), str(device))
        keep = _KEEP_CACHE.get(key)
        if keep is None:              # made once (eager warm-up): a host-to-device copy cannot be captured in a CUDA graph
            keep = _KEEP_CACHE[key] = torch.tensor([1.0 - p for p in key[0]], dtype=torch.float32).clamp_min(1e-12).to(device)
        m = torch.rand(len(live), batch, dtype=torch.float32, device=device).add_(keep[:, None]).floor_().div_(keep[:, None])
        for j, i in enumerate(live):
            out[i] = m[j]
    return out


def _path_scale(p: float, training: bool, batch: int, device) -> Optional[torch.Tensor]:
    """timm DropPath: per-sample bernoulli(keep)/keep multiplier on the residual branch (None = identity)."""
    if p == 0. or not training:
        return None
    keep = 1.0 - p
    m = torch.empty(batch, dtype=torch.float32, device=device).bernoulli_(keep)
    if keep > 0.0:
        m.div_(keep)
    return m


def _block_params(blk):
    """ConvNeXt block parameters under the reference's key names, `gamma` absent when the block has no layer scale."""
    p = _params(blk)
    if getattr(blk, 'gamma', None) is None:
        p.pop('gamma', None)
    return p


# model -> ops.BlockWeights, held outside the module so deepcopy (EMA) and pickling see plain nn.Modules
_BLOCK_WEIGHTS = weakref.WeakKeyDictionary()


def block_weights(model, blocks, T):
    """Refresh and return the prepared operands of `blocks` (all ConvNeXt blocks of `model`), or None when unsupported."""
    if not blocks or not ops.BlockWeights.supported([_block_params(b) for b in blocks], T):
        return None
    bw = _BLOCK_WEIGHTS.get(model)
    if bw is None:
        bw = _BLOCK_WEIGHTS[model] = ops.BlockWeights(lambda: [_block_params(b) for b in blocks])
    bw.refresh()
    return [bw.get(j) for j in range(len(blocks))]


class LayerNorm2d(nn.LayerNorm):
    """Parameter holder for the channel LayerNorm of stem / downsample (ga_convnext.py:51-67), eps 1e-6."""

    def __init__(self, normalized_shape, eps=1e-6):
        super().__init__(normalized_shape, eps=eps)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class ConvNeXtBlock(nn.Module):
    """ga_convnext.py:70-112.  One fused autograd node: K1 + two tcgen05 GEMMs (see ops.ConvNeXtBlockFn)."""

    def __init__(self, dim, drop_path=0., ls_init_value=1e-6, mlp_ratio=4):
        super().__init__()
        self.conv_dw = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, int(mlp_ratio * dim))
        self.gamma = nn.Parameter(ls_init_value * torch.ones(dim)) if ls_init_value > 0 else None
        self.drop_prob = float(drop_path)

    def run(self, x, xs, geom, T, ps=False, ps_prev=None, prep=None):
        """x: residual stream (fp32 under bf16 compute, as in the autocast reference); xs: its bf16 shadow or None.
        ps: this block's DropPath factors (drawn here when not given); ps_prev: those of the previous block; prep: this block's
        prepared operands (ops.BlockWeights) or None."""
        p = _params(self)
        if self.gamma is None:
            p['gamma'] = torch.ones_like(p['norm.weight'])
        if ps is False:
            ps = _path_scale(self.drop_prob, self.training, geom[0], x.device)
        return ops.convnext_block(x, p, geom, ps, torch.is_grad_enabled(), xs=xs, T=T, ps_prev=ps_prev, prep=prep)


class ConvNeXtStage(nn.Module):
    """ga_convnext.py:115-150: optional LN2d + 2x2/2 conv, block sequence, taps for deep stages."""

    def __init__(self, in_chs, out_chs, stride=2, depth=2, dp_rates=None, ls_init_value=1.0, stage3_naggre=2):
        super().__init__()
        self.stage3_naggre = stage3_naggre
        self.stride = stride
        if in_chs != out_chs or stride > 1:
            self.downsample = nn.Sequential(LayerNorm2d(in_chs), nn.Conv2d(in_chs, out_chs, kernel_size=stride, stride=stride))
        else:
            self.downsample = nn.Identity()
        dp_rates = dp_rates or [0.] * depth
        self.blocks = nn.Sequential(*[ConvNeXtBlock(out_chs, drop_path=dp_rates[j], ls_init_value=ls_init_value)
                                      for j in range(depth)])

    def run(self, x, xs, geom, T, RT, scales=None, preps=None):
        """x / xs: residual stream (dtype RT) and its compute-dtype (T) shadow (None when RT == T).
        scales / preps: per-block DropPath factors and prepared operands when the model made them for all stages at once.
        Returns (x, xs, geom, taps); taps hold the compute-dtype view of the tapped activations."""
        Bn, H, W = geom
        if not isinstance(self.downsample, nn.Identity):
            ln, conv = self.downsample[0], self.downsample[1]
            k = self.stride
            cin, cout = conv.in_channels, conv.out_channels
            h = ops.layernorm(xs if xs is not None else x, ln.weight, ln.bias, ln.eps)
            if k > 1:
                h = ops.patchify(h, (Bn, H, W, cin), k)
                H, W = H // k, W // k
            x = ops.linear(h, conv.weight.permute(0, 2, 3, 1).reshape(cout, k * k * cin), conv.bias, out_dtype=RT)
            xs = ops.to_dtype(x, T) if RT != T else None
        geom = (Bn, H, W)
        taps: List[torch.Tensor] = []
        n = len(self.blocks)
        if scales is None:
            scales = _path_scales([blk.drop_prob for blk in self.blocks], self.training, Bn, x.device)
        for i, blk in enumerate(self.blocks):
            x, xs = blk.run(x, xs, geom, T, ps=scales[i], ps_prev=scales[i - 1] if i > 0 else None,
                            prep=preps[i] if preps is not None else None)
            if n > 5 and (i + 1) % (n // (self.stage3_naggre + 1)) == 0 and len(taps) < self.stage3_naggre:
                taps.append(xs if xs is not None else x)
        return x, xs, geom, taps


class _SE(nn.Module):
    """timm SEModule parameter layout (fc1/fc2 1x1 convs); reduce width = make_divisible(C/4, 8, round_limit=0)."""

    def __init__(self, channels, rd_ratio=0.25):
        super().__init__()
        rd = max(8, int(channels * rd_ratio + 4) // 8 * 8)
        self.fc1 = nn.Conv2d(channels, rd, kernel_size=1, bias=True)
        self.fc2 = nn.Conv2d(rd, channels, kernel_size=1, bias=True)


class Bottleneck(nn.Module):
    """ga_convnext.py:251-318: 1x1+BN+ReLU, 3x3+BN+ReLU, SE, 1x1+BN, DropPath, + (1x1+BN shortcut), ReLU."""

    def __init__(self, inplanes, planes, outplanes, drop_path=0.):
        super().__init__()
        self.downsample = nn.Sequential(nn.Conv2d(inplanes, outplanes, kernel_size=1, stride=1), nn.BatchNorm2d(outplanes))
        width = planes
        self.conv1 = nn.Conv2d(inplanes, width, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.se = _SE(width)
        self.conv3 = nn.Conv2d(width, outplanes, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(outplanes)
        self.drop_prob = float(drop_path or 0.)

    def run(self, x, geom):
        Bn, H, W = geom
        tr = self.training
        cin, width, cout = self.conv1.in_channels, self.conv1.out_channels, self.conv3.out_channels
        y = ops.linear(x, self.conv1.weight.reshape(width, cin))
        y = ops.batchnorm(y, _params(self.bn1), tr, relu=True)
        col = ops.im2col3(y, geom)
        y = ops.linear(col, self.conv2.weight.permute(0, 2, 3, 1).reshape(width, 9 * width))
        y = ops.batchnorm(y, _params(self.bn2), tr, relu=True)
        y = ops.se_gate(y, self.se.fc1.weight, self.se.fc1.bias, self.se.fc2.weight, self.se.fc2.bias, Bn, H * W)
        y = ops.linear(y, self.conv3.weight.reshape(cout, width))
        sc = ops.linear(x, self.downsample[0].weight.reshape(cout, cin), self.downsample[0].bias)
        ps = _path_scale(self.drop_prob, tr, Bn, x.device)
        if ps is None:
            return ops.batchnorm(y, _params(self.bn3), tr, relu=True, xb=sc, bnb=_params(self.downsample[1]))
        y = ops.batchnorm(y, _params(self.bn3), tr)
        y = ops.scale_rows(y, ps, H * W)
        return ops.batchnorm(sc, _params(self.downsample[1]), tr, relu=True, xb=y)


class ClassAttn(nn.Module):
    """ga_convnext.py:153-187 parameter holder (q from the class token only; k, v from all tokens)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, dim_embed=128):
        super().__init__()
        self.dim_embed, self.num_heads = dim_embed, num_heads
        self.scale = (dim_embed // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim_embed, bias=qkv_bias)
        self.k = nn.Linear(dim, dim_embed, bias=qkv_bias)
        self.v = nn.Linear(dim, dim_embed, bias=qkv_bias)
        self.proj = nn.Linear(dim_embed, dim)


class GroupConvMlp(nn.Module):
    """ga_convnext.py:190-222 parameter holder: grouped 1x1 -> act -> channel_shuffle -> grouped 1x1."""

    def __init__(self, in_features, hidden_features, groups):
        super().__init__()
        self.groups = groups
        self.fc1 = nn.Conv2d(in_features, hidden_features, kernel_size=1, bias=True, groups=groups)
        self.fc2 = nn.Conv2d(hidden_features, in_features, kernel_size=1, bias=True, groups=groups)

    ACT = ACT_GELU

    def run(self, t, T=torch.float32, drop: float = 0.0):
        """t [B, C] fp32 token; T: operand dtype (bf16 under autocast, like the reference's convs).  Output fp32 [B, C].
        drop: dropout on the hidden activation after the nonlinearity (map.GroupConvMlp.drop, map.py:61; 0 for GA).
        The channel shuffle between the two grouped convs is a re-striding of the hidden activation; with bf16 operands both
        GEMM inputs are laid out group-major with 16-byte pitches so they run on the tcgen05 path (M is only the batch)."""
        Bn, Cc = t.shape
        g = self.groups
        hid = self.fc1.out_channels
        cg, hg = Cc // g, hid // g
        if T == torch.float32:
            a3 = t.view(Bn, g, cg).transpose(0, 1)
            h = ops.grouped_linear(a3, self.fc1.weight.view(g, hg, cg), self.fc1.bias, act=self.ACT)
            if drop > 0.0:
                h = torch.nn.functional.dropout(h, drop, True)
            # channel_shuffle: shuffled[a*(hid/g) + b] = h[b*g + a]  (ga_convnext.py:557-566)
            a3 = h.view(Bn, hg, g).permute(2, 0, 1)
            return ops.grouped_linear(a3, self.fc2.weight.view(g, cg, hg), self.fc2.bias)
        a3 = torch.empty(g, Bn, ops.pad8(cg), dtype=T, device=t.device)[:, :, :cg]
        a3 = a3.copy_(t.view(Bn, g, cg).transpose(0, 1))
        h = ops.grouped_linear(a3, self.fc1.weight.view(g, hg, cg), self.fc1.bias, act=self.ACT)
        if drop > 0.0:
            h = torch.nn.functional.dropout(h, drop, True)
        a3 = torch.empty(g, Bn, ops.pad8(hg), dtype=T, device=t.device)[:, :, :hg]
        a3 = a3.copy_(h.view(Bn, hg, g).permute(2, 0, 1))
        return ops.grouped_linear(a3, self.fc2.weight.view(g, cg, hg), self.fc2.bias, out_dtype=torch.float32)


class LayerScaleBlockClassAttn(nn.Module):
    """ga_convnext.py:225-248 parameter holder; evaluated branch-batched in GA_ConvNeXt._heads."""

    def __init__(self, dim, num_heads, mlp_ratio=4., mlp_block_groups=2, init_values=1e-4, dim_embed=128):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = ClassAttn(dim, num_heads=num_heads, dim_embed=dim_embed)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = GroupConvMlp(dim, int(dim * mlp_ratio), mlp_block_groups)
        self.gamma_1 = nn.Parameter(init_values * torch.ones(dim))
        self.gamma_2 = nn.Parameter(init_values * torch.ones(dim))


def _init_weights(module):
    """ga_convnext.py:508-519: trunc_normal(std .02) on every Conv2d / Linear weight, zero bias."""
    if isinstance(module, (nn.Conv2d, nn.Linear)):
        nn.init.trunc_normal_(module.weight, std=.02, a=-2., b=2.)
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)


def _apply_children_first(fn, module):
    """timm.named_apply(depth_first=True, include_root=False) traversal order."""
    for child in module.children():
        _apply_children_first(fn, child)
        fn(child)


class GA_ConvNeXt(nn.Module):
    """ga_convnext.py:320-505."""

    def __init__(self, in_chans=3, num_classes=1000, output_stride=32, patch_size=4, depths=(3, 3, 9, 3, 1),
                 dims=(96, 192, 384, 768, 768), ls_init_value=1e-6, conv_mlp=False, head_init_scale=1., norm_layer=None,
                 drop_rate=0., drop_path_rate=0., branches=5, gram_embedding_gropus=8, dim_embed=128, stage3_naggre=2,
                 gram_dim=192, gram_layer=True, **unused):
        super().__init__()
        assert output_stride == 32 and in_chans == 3 and not conv_mlp and norm_layer is None and gram_layer
        depths, dims = tuple(depths), tuple(dims)
        self.num_classes, self.drop_rate = num_classes, drop_rate
        self.patch_size, self.branches, self.gram_dim = patch_size, branches, gram_dim
        self.embed_groups, self.naggre, self.dims = gram_embedding_gropus, stage3_naggre, dims
        self.compute_dtype = None   # None: follow autocast; else torch.float32 / torch.bfloat16
        self.default_cfg = self.pretrained_cfg = default_cfgs['ga_convnext_tiny']

        self.stem = nn.Sequential(nn.Conv2d(in_chans, dims[0], kernel_size=patch_size, stride=patch_size), LayerNorm2d(dims[0]))
        dp_rates = [x.tolist() for x in torch.linspace(0, drop_path_rate, sum(depths)).split(depths)]
        stages, prev = [], dims[0]
        for i in range(len(dims)):
            if i == 4:
                prev = sum(dims[:-1]) + dims[2] * stage3_naggre
                stages.append(Bottleneck(prev, dims[i] // 4, dims[i], drop_path=drop_path_rate))
            else:
                stages.append(ConvNeXtStage(prev, dims[i], stride=2 if i > 0 else 1, depth=depths[i], dp_rates=dp_rates[i],
                                            ls_init_value=ls_init_value, stage3_naggre=stage3_naggre))
            prev = dims[i]
        self.stages = nn.Sequential(*stages)
        self.num_features = prev

        self.gram_contraction, self.gram_layer = nn.ModuleList(), nn.ModuleList()
        self.gram_embedding, self.ga, self.fc = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        tri = (gram_dim + 1) * gram_dim // 2
        for _ in range(branches):
            self.gram_contraction.append(nn.Sequential(nn.Conv2d(dims[-1], gram_dim, kernel_size=1), nn.BatchNorm2d(gram_dim)))
            self.gram_layer.append(ConvNeXtStage(gram_dim, gram_dim, stride=1, depth=1, dp_rates=dp_rates[-1],
                                                 ls_init_value=ls_init_value))
            self.gram_embedding.append(nn.Sequential(nn.Conv2d(tri, dims[-1], kernel_size=1, groups=gram_embedding_gropus),
                                                     nn.BatchNorm2d(dims[-1])))
            self.ga.append(LayerScaleBlockClassAttn(dims[-1], num_heads=8, mlp_block_groups=4, dim_embed=dim_embed))
            self.fc.append(nn.Linear(dims[-1], num_classes))
        _apply_children_first(_init_weights, self)

    # -- reference API surface kept for callers -------------------------------------------------------------
    @torch.jit.ignore
    def no_weight_decay(self):
        return set()

    def get_classifier(self):
        return self.fc

    # -- execution ------------------------------------------------------------------------------------------------
    def _dtype(self, x):
        if self.compute_dtype is not None:
            return self.compute_dtype
        if torch.is_autocast_enabled():
            dt = torch.get_autocast_dtype('cuda')
            if dt == torch.float16:
                warnings.warn('fp16 autocast requested: the sm_100a kernels compute in bf16 (fp32 accumulate) instead', stacklevel=3)
            return torch.bfloat16
        return torch.float32

    def forward_features(self, x):
        """stem -> 4 stages -> aggregation -> Bottleneck (ga_convnext.py:469-485); returns ([B*14*14, C] rows, geom)."""
        if not x.is_cuda:
            raise ops.L.GaError('GA_ConvNeXt runs on CUDA (sm_100a) tensors only: there is no CPU path')
        T = self._dtype(x)
        Bn, _, H, W = x.shape
        k = self.patch_size
        stem_conv, stem_ln = self.stem[0], self.stem[1]
        with torch.autocast('cuda', enabled=False):
            RT = torch.float32                         # residual stream stays fp32 (autocast semantics of the reference)
            rows = ops.stem_patchify(x.float(), k, T)
            y = ops.linear(rows, stem_conv.weight.permute(0, 2, 3, 1).reshape(stem_conv.out_channels, -1), stem_conv.bias,
                           out_dtype=RT)
            y = ops.layernorm(y, stem_ln.weight, stem_ln.bias, stem_ln.eps)
            ys = ops.to_dtype(y, T) if RT != T else None
            geom = (Bn, H // k, W // k)
            feats, taps = [], []
            # per-forward work shared by all 18 blocks: one DropPath draw, one operand-preparation launch
            blocks = [blk for i in range(4) for blk in self.stages[i].blocks]
            n_stage_blocks = len(blocks)
            blocks += [blk for gl in self.gram_layer for blk in gl.blocks]      # the heads' gram layers ride along
            scales = _path_scales([blk.drop_prob for blk in blocks], self.training, Bn, x.device)
            preps = block_weights(self, blocks, T)
            self.__dict__['_gram_layer_aux'] = (scales[n_stage_blocks:], preps[n_stage_blocks:] if preps is not None else None)
            off = 0
            for i in range(4):
                n = len(self.stages[i].blocks)
                y, ys, geom, t = self.stages[i].run(y, ys, geom, T, RT, scales=scales[off:off + n],
                                                    preps=preps[off:off + n] if preps is not None else None)
                off += n
                feats.append((ys if ys is not None else y, geom))
                taps += [(tt, geom) for tt in t]
            (x0, g0), (x1, g1), (x2, g2), (x3, g3) = feats
            Ho, Wo = g2[1], g2[2]     # the reference pools to 14 = H/16 at 224 (ga_convnext.py:397); generalised to H/16
            items = [(g0[1], g0[2], x0.shape[1], 0), (g1[1], g1[2], x1.shape[1], 0)]
            items += [(g[1], g[2], t.shape[1], 1) for t, g in taps]
            items += [(g2[1], g2[2], x2.shape[1], 1), (g3[1], g3[2], x3.shape[1], 2)]
            cat = ops.aggregate((Bn, Ho, Wo, items), [x0, x1, *[t for t, _ in taps], x2, x3])
            f = self.stages[4].run(cat, (Bn, Ho, Wo))
        return f, (Bn, Ho, Wo)

    def _gram_contractions(self, f):
        """The `branches` 1x1 contraction convs as ONE GEMM of width branches * gram_dim on the shared feature rows: the
        features are read once instead of per branch, and the data gradient is one K = branches * gram_dim GEMM instead of
        five GEMMs whose [M, C] results autograd has to add up.  Returns the per-branch column slices (strided views)."""
        w = torch.cat([gc[0].weight.reshape(self.gram_dim, f.shape[1]) for gc in self.gram_contraction], 0)
        b = torch.cat([gc[0].bias for gc in self.gram_contraction], 0)
        return ops.linear(f, w, b).split(self.gram_dim, dim=1)

    def _gram_features(self, k, f, geom, g=None):
        """Branch k's contraction conv (unless its output `g` is given) + BN + gram layer on the shared feature rows -> [B*HW, gram_dim]."""
        conv, bn = self.gram_contraction[k][0], self.gram_contraction[k][1]
        if g is None:
            g = ops.linear(f, conv.weight.reshape(self.gram_dim, f.shape[1]), conv.bias)
        g = ops.batchnorm(g, _params(bn), self.training)
        aux = self.__dict__.get('_gram_layer_aux')
        nblk = len(self.gram_layer[k].blocks)
        if aux is not None and g.dtype == torch.bfloat16 and len(aux[0]) == nblk * len(self.gram_layer):
            sc = aux[0][k * nblk:(k + 1) * nblk]
            pr = aux[1][k * nblk:(k + 1) * nblk] if aux[1] is not None else None
            g, _, _, _ = self.gram_layer[k].run(g, None, geom, g.dtype, g.dtype, scales=sc, preps=pr)
        else:
            g, _, _, _ = self.gram_layer[k].run(g, None, geom, g.dtype, g.dtype)
        return g

    def _heads(self, f, geom):
        """The `branches` GA heads (ga_convnext.py:491-504; shared with GA-CSWin, ga_cswin.py:673-693).  Token-side work is batched over branches: one shared
        normalisation of the 196 tokens (norm1's affine folded into k/v), one k/v projection GEMM of width
        branches*2E and one attention-pooling pass; class-token work ([B, C] rows) stays fp32."""
        Bn, H, W = geom
        HW = H * W
        nb = self.branches
        tr = self.training
        Cc = f.shape[1]
        T = f.dtype
        E = self.ga[0].attn.dim_embed
        heads = self.ga[0].attn.num_heads
        fhat = ops.layernorm(f, None, None, self.ga[0].norm1.eps)
        # norm1's affine folded into the token-side k / v projection of every branch: W diag(w) and W b.  Elementwise + row sums on
        # the stacked [nb, 2E, C] weights (no library GEMV on the path); autograd carries the four parameter gradients
        w_kv = torch.stack([torch.cat((g.attn.k.weight, g.attn.v.weight), 0) for g in self.ga])          # [nb, 2E, C]
        n1w = torch.stack([g.norm1.weight for g in self.ga]).unsqueeze(1)
        n1b = torch.stack([g.norm1.bias for g in self.ga]).unsqueeze(1)
        wkv = (w_kv * n1w).view(nb * 2 * E, Cc)
        bkv = (w_kv * n1b).sum(-1).view(nb * 2 * E)
        # k / v of the 196 tokens are kept in fp32 (0.3 GB more traffic per step at B=256): the softmax backward subtracts
        # nearly equal terms (dP - sum P dP), which turns a bf16 rounding of k / v into 2e-2 on the q / k weight gradients
        kv_tok = ops.linear(fhat, wkv, bkv, out_dtype=torch.float32)               # [B*HW, nb*2E]
        batched = self._can_batch_heads(T)
        gpre = None
        if batched and type(self)._gram_features is GA_ConvNeXt._gram_features and self.gram_dim % 8 == 0:
            gpre = self._gram_contractions(f)
        cls, qs, kvcs = [], [], []
        for k in range(nb):
            g = self._gram_features(k, f, geom, gpre[k]) if gpre is not None else self._gram_features(k, f, geom)
            emb, ebn = self.gram_embedding[k][0], self.gram_embedding[k][1]
            G = self.embed_groups
            glen = emb.weight.shape[1]
            c = ops.gram_embed(g, emb.weight.view(G, Cc // G, glen), emb.bias, Bn, HW, float(H))   # get_gram + embedding, [B, C] fp32
            c = ops.batchnorm(c, _params(ebn), tr)
            cls.append(c)
            if batched:
                continue
            blk = self.ga[k]
            cn = ops.layernorm(c, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
            cn_t = ops.to_dtype(cn, T)       # like autocast: fp32 LayerNorm output, bf16 operands for the Linear
            qs.append(ops.linear(cn_t, blk.attn.q.weight * blk.attn.scale, out_dtype=torch.float32))
            kvcs.append(ops.linear(cn_t, torch.cat((blk.attn.k.weight, blk.attn.v.weight), 0), out_dtype=torch.float32))
        cst = None
        if batched:
            # norm1 + q / k / v of the class tokens for all branches at once (see _heads_tail_batched)
            blks = list(self.ga)
            cst = torch.stack(cls)                                                     # [nb, B, C] fp32
            cn = ops.layernorm(cst.view(nb * Bn, Cc), None, None, blks[0].norm1.eps).view(nb, Bn, Cc)
            cn = torch.addcmul(torch.stack([b.norm1.bias for b in blks]).unsqueeze(1), cn,
                               torch.stack([b.norm1.weight for b in blks]).unsqueeze(1))
            cn_t = ops.to_dtype(cn.view(nb * Bn, Cc), T).view(nb, Bn, Cc)
            q = ops.stacked_linear(cn_t, [b.attn.q.weight for b in blks], out_dtype=torch.float32, alpha=blks[0].attn.scale).view(nb, Bn, 1, E)
            kc = ops.stacked_linear(cn_t, [b.attn.k.weight for b in blks], out_dtype=torch.float32)
            vc = ops.stacked_linear(cn_t, [b.attn.v.weight for b in blks], out_dtype=torch.float32)
            kvc = torch.cat((kc, vc), -1).view(nb, Bn, 1, 2 * E)
        else:
            q = torch.stack(qs).view(nb, Bn, 1, E)
            kvc = torch.stack(kvcs).view(nb, Bn, 1, 2 * E)
        o = ops.attnpool(q, kvc, kv_tok, HW, heads)                                # [nb, B, 1, E]
        if batched:
            return self._heads_tail_batched(o, cst, T)
        outs = []
        for k in range(nb):
            blk = self.ga[k]
            c = cls[k] + blk.gamma_1 * ops.linear(ops.to_dtype(o[k, :, 0], T), blk.attn.proj.weight, blk.attn.proj.bias,
                                                  out_dtype=torch.float32)
            h = ops.layernorm(c, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
            c = c + blk.gamma_2 * blk.mlp.run(h, T)
            outs.append(ops.linear(ops.to_dtype(c, T), self.fc[k].weight, self.fc[k].bias, out_dtype=torch.float32))
        return outs

    def _can_batch_heads(self, T):
        """The branch-batched class-token tail needs bf16 operands and 16-byte group pitches; GA_BATCH_HEADS=0 keeps the per-branch
        form (the fp32 parity path always uses it)."""
        if T != torch.bfloat16 or os.environ.get('GA_BATCH_HEADS', '1') != '1' or self.branches < 2:
            return False
        m = self.ga[0].mlp
        g = m.groups
        # group widths: 4-element multiples for the vector epilogues (172 = 688 / 4 is padded to a 16-byte pitch below)
        return (m.fc1.in_channels // g) % 4 == 0 and (m.fc1.out_channels // g) % 8 == 0 and self.ga[0].attn.dim_embed % 8 == 0

    def _heads_tail_batched(self, o, cls, T):
        """Everything after the attention pooling (ga_convnext.py:240-248, 503) for all branches at once: the class tokens are
        [B, C] rows, so per branch every Linear is a 12-CTA GEMM; stacked over the branches each layer is one grouped launch
        forward, one for the data gradient and one for the weight gradient (ops.StackedLinearFn).  Per-branch parameters
        (layer scales, norm2 affine) are stacked by torch and applied to the [nb, B, C] tensor."""
        nb, Bn = o.shape[0], o.shape[1]
        blks = list(self.ga)
        c = cls                                                                    # [nb, B, C] fp32, the stacked class tokens
        Cc = c.shape[2]
        E = o.shape[-1]
        t = ops.stacked_linear(ops.to_dtype(o.reshape(nb * Bn, E), T).view(nb, Bn, E), [b.attn.proj.weight for b in blks],
                               [b.attn.proj.bias for b in blks], out_dtype=torch.float32)
        c = torch.addcmul(c, torch.stack([b.gamma_1 for b in blks]).unsqueeze(1), t)
        eps = blks[0].norm2.eps
        h = ops.layernorm(c.view(nb * Bn, Cc), None, None, eps).view(nb, Bn, Cc)
        h = torch.addcmul(torch.stack([b.norm2.bias for b in blks]).unsqueeze(1), h, torch.stack([b.norm2.weight for b in blks]).unsqueeze(1))
        # GroupConvMlp: grouped 1x1 -> GELU -> channel shuffle -> grouped 1x1; groups of all branches side by side
        g = blks[0].mlp.groups
        hid = blks[0].mlp.fc1.out_channels
        cg, hg = Cc // g, hid // g
        a3 = torch.empty(nb * g, Bn, ops.pad8(cg), dtype=T, device=c.device)[:, :, :cg]
        a3.view(nb, g, Bn, cg).copy_(h.view(nb, Bn, g, cg).permute(0, 2, 1, 3))
        h1 = ops.stacked_linear(a3, [b.mlp.fc1.weight for b in blks], [b.mlp.fc1.bias for b in blks], act=GroupConvMlp.ACT, sub=g)
        # channel_shuffle (ga_convnext.py:557-566): the second conv's group a reads hidden channels {b * g + a}
        hrow = h1.view(nb, g, Bn, hg).permute(0, 2, 1, 3).reshape(nb, Bn, hg, g)   # rows [B, hid] per branch, viewed as (hg, g)
        a3 = torch.empty(nb, g, Bn, hg, dtype=T, device=c.device).copy_(hrow.permute(0, 3, 1, 2)).view(nb * g, Bn, hg)
        m = ops.stacked_linear(a3, [b.mlp.fc2.weight for b in blks], [b.mlp.fc2.bias for b in blks], out_dtype=torch.float32, sub=g)
        m = m.view(nb, g, Bn, cg).permute(0, 2, 1, 3).reshape(nb, Bn, Cc)          # back to rows [B, C] (group-major columns)
        c = torch.addcmul(c, torch.stack([b.gamma_2 for b in blks]).unsqueeze(1), m)
        out = ops.stacked_linear(ops.to_dtype(c.view(nb * Bn, Cc), T).view(nb, Bn, Cc), [fc.weight for fc in self.fc], [fc.bias for fc in self.fc],
                                 out_dtype=torch.float32)
        return list(out.unbind(0))

    def forward(self, x):
        with ops.collect_bn_counters():
            f, geom = self.forward_features(x)
            with torch.autocast('cuda', enabled=False):
                return self._heads(f, geom)


def _create_convnext(variant, pretrained=False, **kwargs):
    return build_model_with_cfg(GA_ConvNeXt, variant, pretrained, **kwargs)


@register_model
def ga_convnext_tiny_688(pretrained=False, **kwargs):
    args = dict(depths=[3, 3, 9, 3, 1], dims=[96, 192, 384, 688, 688], gram_embedding_gropus=8, **kwargs)
    return _create_convnext('ga_convnext_tiny', pretrained=pretrained, dim_embed=168, stage3_naggre=2, gram_dim=192, **args)


@register_model
def ga_convnext_tiny_768(pretrained=False, **kwargs):
    args = dict(depths=[3, 3, 9, 3, 1], dims=[96, 192, 384, 768, 768], gram_embedding_gropus=8, **kwargs)
    return _create_convnext('ga_convnext_tiny', pretrained=pretrained, dim_embed=192, stage3_naggre=2, gram_dim=192, **args)


@register_model
def ga_convnext_small_688(pretrained=False, **kwargs):
    args = dict(depths=[3, 3, 27, 3, 1], dims=[96, 192, 384, 688, 688], gram_embedding_gropus=8, **kwargs)
    return _create_convnext('ga_convnext_small', pretrained=pretrained, dim_embed=168, stage3_naggre=4, gram_dim=192, **args)


@register_model
def ga_convnext_small_768(pretrained=False, **kwargs):
    args = dict(depths=[3, 3, 27, 3, 1], dims=[96, 192, 384, 768, 768], gram_embedding_gropus=8, **kwargs)
    return _create_convnext('ga_convnext_small', pretrained=pretrained, dim_embed=192, stage3_naggre=4, gram_dim=192, **args)


@register_model
def ga_convnext_base_976(pretrained=False, **kwargs):
    args = dict(depths=[3, 3, 27, 3, 1], dims=[128, 256, 512, 976, 976], gram_embedding_gropus=8, dim_embed=240,
                stage3_naggre=4, gram_dim=192, **kwargs)
    return _create_convnext('ga_convnext_base', pretrained=pretrained, **args)


@register_model
def ga_convnext_base_1024(pretrained=False, **kwargs):
    args = dict(depths=[3, 3, 27, 3, 1], dims=[128, 256, 512, 1024, 1024], gram_embedding_gropus=8, dim_embed=256,
                stage3_naggre=4, gram_dim=192, **kwargs)
    return _create_convnext('ga_convnext_base', pretrained=pretrained, **args)
