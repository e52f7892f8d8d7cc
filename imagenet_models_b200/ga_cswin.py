"""GA-CSWin on sm_100a kernels: drop-in for `GA/ga_cswin.py` (same class name, constructor arguments, attribute names
and state_dict).

As in ga_convnext.py the module tree only holds parameters in the reference's layout -- parameter-free entries of the
reference's `nn.Sequential`s (einops Rearrange, GELU) are `nn.Identity` placeholders so the state_dict indices match --
and the arithmetic runs on token rows [B*H*W, C] through libga_sm100.so: K6 stripe attention (csrc/attn.cu), the tcgen05
GEMM with fused bias / GELU / residual epilogues, LayerNorm, strided 3x3 patch gathers.  No PyTorch/CPU fallback.

Construction order mirrors ga_cswin.py:448-608 so that `torch.manual_seed(s); GA_CSWinTransformer(...)` yields
bit-identical initial weights.  The reference file registers no factory; `ga_CSWin_64_12211_tiny_224` below uses the
constructor arguments SURVEY.md section 8 (a18) infers for the published GA-CSWin-T (43.4 M parameters).
"""
from __future__ import annotations

import warnings

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .ga_convnext import (GA_ConvNeXt, LayerScaleBlockClassAttn, _apply_children_first, _params, _path_scale)
from .registry import register_model

__all__ = ['GA_CSWinTransformer']

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def _cfg(url='', **kwargs):
    return {'url': url, 'num_classes': 1000, 'input_size': (3, 224, 224), 'pool_size': None, 'crop_pct': .9,
            'interpolation': 'bicubic', 'mean': IMAGENET_DEFAULT_MEAN, 'std': IMAGENET_DEFAULT_STD,
            'first_conv': 'patch_embed.proj', 'classifier': 'head', **kwargs}


default_cfgs = dict(ga_CSWin_64_12211_tiny_224=_cfg(), ga_CSWin_64_24322_small_224=_cfg())


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class LePEAttention(nn.Module):
    """ga_cswin.py:59-88 parameter holder: the depthwise 3x3 `get_v` of one branch."""

    def __init__(self, dim, resolution, idx, split_size, num_heads):
        super().__init__()
        self.dim, self.resolution, self.split_size, self.num_heads, self.idx = dim, resolution, split_size, num_heads, idx
        self.get_v = nn.Conv2d(dim, dim, kernel_size=3, stride=1, padding=1, groups=dim)


class CSWinBlock(nn.Module):
    """ga_cswin.py:139-212.  One fused autograd node (ops.CSWinBlockFn)."""

    def __init__(self, dim, reso, num_heads, split_size=7, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, last_stage=False, mlp_groups=1):
        super().__init__()
        assert qkv_bias and qk_scale is None and drop == 0. and attn_drop == 0. and mlp_groups == 1, \
            'sm_100a CSWinBlock: qkv_bias=True, default scale, no dropout, plain Mlp (the published configuration)'
        assert dim // num_heads == 32, f'stripe-attention kernel is built for 32-wide heads (dim={dim}, heads={num_heads})'
        self.dim, self.num_heads, self.patches_resolution, self.split_size = dim, num_heads, reso, split_size
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.norm1 = nn.LayerNorm(dim)
        if reso == split_size:
            last_stage = True
        self.branch_num = 1 if last_stage else 2
        self.proj = nn.Linear(dim, dim)
        if last_stage:
            self.attns = nn.ModuleList([LePEAttention(dim, reso, -1, split_size, num_heads)])
        else:
            self.attns = nn.ModuleList([LePEAttention(dim // 2, reso, i, split_size, num_heads // 2) for i in range(2)])
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.norm2 = nn.LayerNorm(dim)
        self.drop_prob = float(drop_path)

    def run(self, x, xs, Bn, T):
        p = _params(self)
        tr = self.training
        ps1 = _path_scale(self.drop_prob, tr, Bn, x.device)
        ps2 = _path_scale(self.drop_prob, tr, Bn, x.device)
        return ops.cswin_block(x, p, (Bn, self.patches_resolution, self.split_size, self.branch_num), ps1, ps2,
                               torch.is_grad_enabled(), xs=xs, T=T)


class Merge_Block(nn.Module):
    """ga_cswin.py:253-268: 3x3 stride-2 conv + LayerNorm on token rows."""

    def __init__(self, dim, dim_out):
        super().__init__()
        self.conv = nn.Conv2d(dim, dim_out, 3, 2, 1)
        self.norm = nn.LayerNorm(dim_out)

    def run(self, h, geom, RT):
        Bn, H, W = geom
        cin, cout = self.conv.in_channels, self.conv.out_channels
        col = ops.im2col3(h, geom, 2)
        y = ops.linear(col, self.conv.weight.permute(0, 2, 3, 1).reshape(cout, 9 * cin), self.conv.bias, out_dtype=RT)
        return ops.layernorm(y, self.norm.weight, self.norm.bias, self.norm.eps)


class Merge_Block_LCF(nn.Module):
    """ga_cswin.py:236-251: 1x1 conv + LayerNorm on token rows."""

    def __init__(self, dim, dim_out):
        super().__init__()
        self.conv = nn.Conv2d(dim, dim_out, kernel_size=1, stride=1, padding=0)
        self.norm = nn.LayerNorm(dim_out)

    def run(self, h, RT):
        cin, cout = self.conv.in_channels, self.conv.out_channels
        y = ops.linear(h, self.conv.weight.reshape(cout, cin), self.conv.bias, out_dtype=RT)
        return ops.layernorm(y, self.norm.weight, self.norm.bias, self.norm.eps)


def _init_weights(m):
    """ga_cswin.py:610-617: trunc_normal(.02) on Linear only (convs keep PyTorch's default init); norms to (1, 0)."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02, a=-2., b=2.)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, (nn.LayerNorm, nn.BatchNorm2d)):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)


class GA_CSWinTransformer(nn.Module):
    """ga_cswin.py:445-693."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=64, depth=[2, 2, 6, 2],
                 split_size=[3, 5, 7], num_heads=12, mlp_ratio=4., mlp_ratio_stage4=4., mlp_ratio_stage5=4., qkv_bias=True,
                 qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0., norm_layer=nn.LayerNorm, use_chk=False,
                 dims=[64, 128, 256, 512], stage3_naggre=4, ga_mlp_groups=2, ga_layer_mlp_groups=1, branches=5, gram_dim=192,
                 deep_stem=True, stage5='CSWin', stage5_mlp_groups=1, ga_layer=True):
        super().__init__()
        assert deep_stem and stage5 == 'CSWin' and ga_layer and in_chans == 3 and norm_layer is nn.LayerNorm, \
            'sm_100a GA-CSWin covers the published configuration: deep stem, CSWin stage 5, gram layers'
        assert ga_layer_mlp_groups == 1 and stage5_mlp_groups == 1 and drop_rate == 0. and attn_drop_rate == 0.
        assert img_size % 32 == 0 and len(num_heads) == 5 and len(split_size) == 5     # ga_cswin.py:535-536 reads [4]
        if use_chk:
            warnings.warn('use_chk (activation checkpointing) is ignored: activations stay resident in HBM', stacklevel=2)
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.stage3_naggre = self.naggre = stage3_naggre
        self.img_size, self.branches, self.gram_dim, self.embed_groups = img_size, branches, gram_dim, 8
        self.compute_dtype = None
        self.default_cfg = self.pretrained_cfg = default_cfgs['ga_CSWin_64_12211_tiny_224']
        heads = num_heads
        ident = nn.Identity
        self.stage1_conv_embed = nn.Sequential(
            nn.Conv2d(in_chans, embed_dim, 3, stride=2, padding=1, bias=False), ident(), nn.LayerNorm(embed_dim), ident(), ident(),
            nn.Conv2d(embed_dim, embed_dim, 3, stride=1, padding=1, bias=False), ident(), nn.LayerNorm(embed_dim), ident(), ident(),
            nn.Conv2d(embed_dim, dims[0], 3, stride=2, padding=1, bias=False), ident(), nn.LayerNorm(dims[0]))
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, int(np.sum(depth)))]

        def blocks(s, n, off, ratio, last=False):
            return nn.ModuleList([CSWinBlock(dim=dims[s], num_heads=heads[s], reso=img_size // (4 << s), mlp_ratio=ratio,
                                             qkv_bias=qkv_bias, qk_scale=qk_scale, split_size=split_size[s] if not last else split_size[-1],
                                             drop_path=dpr[off + i], last_stage=last) for i in range(n)])
        self.stage1 = blocks(0, depth[0], 0, mlp_ratio)
        self.merge1 = Merge_Block(dims[0], dims[1])
        self.stage2 = blocks(1, depth[1], int(np.sum(depth[:1])), mlp_ratio)
        self.merge2 = Merge_Block(dims[1], dims[2])
        self.stage3 = blocks(2, depth[2], int(np.sum(depth[:2])), mlp_ratio)
        self.merge3 = Merge_Block(dims[2], dims[3])
        self.stage4 = blocks(3, depth[3], int(np.sum(depth[:3])), mlp_ratio_stage4, last=True)
        curr = dims[3]
        aggre_dim = sum(dims) + dims[2] * stage3_naggre
        self.merge4 = None
        self.stage5 = nn.Sequential(ident(), Merge_Block_LCF(aggre_dim, curr),
                                    CSWinBlock(dim=curr, num_heads=heads[4], reso=img_size // 16, mlp_ratio=mlp_ratio_stage5,
                                               qkv_bias=qkv_bias, qk_scale=qk_scale, split_size=split_size[4], drop_path=dpr[-1]),
                                    ident())
        self.gram_contraction, self.gram_layer = nn.ModuleList(), nn.ModuleList()
        self.gram_embedding, self.ga, self.fc = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        tri = (gram_dim + 1) * gram_dim // 2
        for _ in range(branches):
            self.gram_contraction.append(nn.Sequential(nn.Conv2d(curr, gram_dim, kernel_size=1, groups=8), nn.BatchNorm2d(gram_dim)))
            self.gram_layer.append(nn.Sequential(ident(), CSWinBlock(dim=gram_dim, num_heads=6, reso=img_size // 16, qkv_bias=qkv_bias,
                                                                     qk_scale=qk_scale, split_size=split_size[4], drop_path=dpr[-1]),
                                                 ident()))
            self.gram_embedding.append(nn.Sequential(nn.Conv2d(tri, curr, kernel_size=1, groups=8), nn.BatchNorm2d(curr)))
            self.ga.append(LayerScaleBlockClassAttn(curr, num_heads=8, mlp_block_groups=ga_mlp_groups, dim_embed=curr // 4))
            self.fc.append(nn.Linear(curr, num_classes))
        _apply_children_first(_init_weights, self)

    # -- reference API surface ----------------------------------------------------------------------------------------
    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    def get_classifier(self):
        return self.fc

    # -- execution ----------------------------------------------------------------------------------------------------
    _dtype = GA_ConvNeXt._dtype
    _heads = GA_ConvNeXt._heads
    _can_batch_heads = GA_ConvNeXt._can_batch_heads
    _heads_tail_batched = GA_ConvNeXt._heads_tail_batched

    def _gram_features(self, k, f, geom):
        """Grouped (g=8) 1x1 contraction + BN + one CSWinBlock at gram_dim (ga_cswin.py:557-576, 676-677)."""
        Bn, H, W = geom
        conv, bn = self.gram_contraction[k][0], self.gram_contraction[k][1]
        M, Cc = f.shape
        G = conv.groups
        a3 = f.view(M, G, Cc // G).transpose(0, 1)
        g = ops.grouped_linear(a3, conv.weight.view(G, self.gram_dim // G, Cc // G), conv.bias)
        g = ops.batchnorm(g, _params(bn), self.training)
        g, _ = self.gram_layer[k][1].run(g.contiguous(), None, Bn, g.dtype)
        return g

    def forward_features(self, x):
        """deep stem -> 4 stages with Merge_Blocks -> aggregation -> stage 5 (ga_cswin.py:636-671); returns (rows, geom)."""
        if not x.is_cuda:
            raise ops.L.GaError('GA_CSWinTransformer runs on CUDA (sm_100a) tensors only: there is no CPU path')
        T = self._dtype(x)
        Bn, _, H, W = x.shape
        assert H == self.img_size and W == self.img_size, 'flatten img_tokens has wrong size'
        with torch.autocast('cuda', enabled=False):
            RT = torch.float32
            st = self.stage1_conv_embed
            e = st[0].out_channels
            w0 = torch.nn.functional.pad(st[0].weight.permute(0, 2, 3, 1).reshape(e, 27), (0, 5))
            y = ops.linear(ops.stem_im2col3(x.float(), 2, T), w0)
            y = ops.gelu(ops.layernorm(y, st[2].weight, st[2].bias, st[2].eps))
            g1 = (Bn, H // 2, W // 2)
            y = ops.linear(ops.im2col3(y, g1, 1), st[5].weight.permute(0, 2, 3, 1).reshape(e, 9 * e))
            y = ops.gelu(ops.layernorm(y, st[7].weight, st[7].bias, st[7].eps))
            d0 = st[10].out_channels
            y = ops.linear(ops.im2col3(y, g1, 2), st[10].weight.permute(0, 2, 3, 1).reshape(d0, 9 * e), out_dtype=RT)
            y = ops.layernorm(y, st[12].weight, st[12].bias, st[12].eps)
            ys = ops.to_dtype(y, T) if RT != T else None
            R = H // 4
            feats = []
            for blk in self.stage1:
                y, ys = blk.run(y, ys, Bn, T)
            feats.append((ys if ys is not None else y, R))
            for li, (pre, blocks) in enumerate(((self.merge1, self.stage2), (self.merge2, self.stage3), (self.merge3, self.stage4))):
                y = pre.run(ys if ys is not None else y, (Bn, R, R), RT)
                ys = ops.to_dtype(y, T) if RT != T else None
                R //= 2
                n = len(blocks)
                for bi, blk in enumerate(blocks):
                    y, ys = blk.run(y, ys, Bn, T)
                    if li == 1 and (bi + 1) % (n // (self.stage3_naggre + 1)) == 0 and len(feats) < self.stage3_naggre + 2:
                        feats.append((ys if ys is not None else y, R))
                feats.append((ys if ys is not None else y, R))
            Ro = H // 16
            items = [(r, r, t.shape[1], 0) for t, r in feats[:2]] + [(r, r, t.shape[1], 1) for t, r in feats[2:-1]]
            items.append((feats[-1][1], feats[-1][1], feats[-1][0].shape[1], 2))
            cat = ops.aggregate((Bn, Ro, Ro, items), [t for t, _ in feats])
            y = self.stage5[1].run(cat, RT)
            ys = ops.to_dtype(y, T) if RT != T else None
            y, ys = self.stage5[2].run(y, ys, Bn, T)
        return (ys if ys is not None else y), (Bn, Ro, Ro)

    def forward(self, x):
        with ops.collect_bn_counters():
            f, geom = self.forward_features(x)
            with torch.autocast('cuda', enabled=False):
                return self._heads(f, geom)


@register_model
def ga_CSWin_64_12211_tiny_224(pretrained=False, **kwargs):
    """GA-CSWin-T (GA/README.md: 42.0 M / 6.1 G): constructor arguments per SURVEY.md section 8 row a18."""
    kwargs.pop('pretrained_cfg', None)
    kwargs.pop('pretrained_cfg_overlay', None)
    args = dict(patch_size=4, embed_dim=64, depth=[1, 2, 21, 1], split_size=[1, 2, 7, 7, 7], num_heads=[2, 4, 8, 16, 16],
                dims=[64, 128, 256, 512], stage3_naggre=4)
    args.update(kwargs)
    return GA_CSWinTransformer(**args)
