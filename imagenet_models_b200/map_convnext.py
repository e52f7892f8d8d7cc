"""MAP-ConvNeXt on sm_100a kernels: drop-in for `MAP/models/map_convnext.py` + `MAP/models/map.py` (MAPHead).

Same factories (`map_convnext_tiny`, `map_convnext_small`), attribute names and state_dict keys (332 entries for tiny,
incl. `head.mmcap.mmcap.*.gram_token_extraction.bp_index`); the module tree only holds parameters, the arithmetic runs
through `ops.*` (libga_sm100.so) on NHWC row matrices.  train mode returns `[main logits, self-distillation logits]`
pairs per group, eval mode the main logits (map.py:519-537).  In train mode the CABlock dropouts of the reference are
applied with its defaults (attn_drop = drop = 0.05, map.py:149): on the attention probabilities inside the attention-pooling
kernel (mask from torch's Philox generator), on the proj output and on the MLP hidden activation; `MAPHead.drop` /
`MAPHead.attn_drop` = 0 gives the deterministic path the parity tests use.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from . import ops
from .ga_convnext import GroupConvMlp as _GAGroupConvMlp
from .ga_convnext import _BLOCK_WEIGHTS, _apply_children_first, _init_weights, _path_scale, _path_scales
from .lib import ACT_RELU
from .registry import register_model

__all__ = ['ConvNeXt', 'MAPHead']


class LayerNorm(nn.Module):
    """map_convnext.LayerNorm (:145-170) parameter holder (channels_last and channels_first are the same row LN here)."""

    def __init__(self, normalized_shape, eps=1e-6, data_format='channels_last'):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps, self.data_format = eps, data_format


class Block(nn.Module):
    """map_convnext.Block (:14-40): same fused kernels as GA's ConvNeXtBlock, reference key names dwconv/pwconv1/pwconv2."""

    def __init__(self, dim, drop_path=0., layer_scale_init_value=1e-6):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None
        self.drop_prob = float(drop_path)

    def params(self):
        """Parameters under the key names ops.convnext_block / ops.BlockWeights use (`gamma` absent without layer scale)."""
        p = {'conv_dw.weight': self.dwconv.weight, 'conv_dw.bias': self.dwconv.bias, 'norm.weight': self.norm.weight,
             'norm.bias': self.norm.bias, 'mlp.fc1.weight': self.pwconv1.weight, 'mlp.fc1.bias': self.pwconv1.bias,
             'mlp.fc2.weight': self.pwconv2.weight, 'mlp.fc2.bias': self.pwconv2.bias}
        if self.gamma is not None:
            p['gamma'] = self.gamma
        return p

    def run(self, x, xs, geom, T, ps=False, ps_prev=None, prep=None):
        p = self.params()
        if self.gamma is None:
            p['gamma'] = torch.ones_like(self.norm.weight)
        if ps is False:
            ps = _path_scale(self.drop_prob, self.training, geom[0], x.device)
        return ops.convnext_block(x, p, geom, ps, torch.is_grad_enabled(), xs=xs, T=T, ps_prev=ps_prev, prep=prep)


class ClassAttention(nn.Module):
    """map.ClassAttention (:69-98), equal-dim branch: parameter holder."""

    def __init__(self, dim, num_heads, embed_dim, qkv_bias=True):
        super().__init__()
        self.num_heads, self.embed_dim = num_heads, embed_dim
        self.scale = (embed_dim // num_heads) ** -0.5
        self.proj = nn.Linear(embed_dim, dim)
        self.q = nn.Linear(dim, embed_dim, bias=qkv_bias)
        self.k = nn.Linear(dim, embed_dim, bias=qkv_bias)
        self.v = nn.Linear(dim, embed_dim, bias=qkv_bias)


class GroupConvMlp(_GAGroupConvMlp):
    """map.GroupConvMlp (:43-66): ReLU instead of GELU, otherwise the GA layout (grouped 1x1, shuffle, grouped 1x1)."""

    ACT = ACT_RELU


class CABlock(nn.Module):
    """map.CABlock (:147-184) parameter holder (registration order norm2, attn, mlp, norm1 as in the reference)."""

    def __init__(self, dim, num_heads, mlp_ratio, groups, ca_dim):
        super().__init__()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = ClassAttention(dim, num_heads, ca_dim)
        self.mlp = GroupConvMlp(dim, int(dim * mlp_ratio), groups)
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)


class GramToken(nn.Module):
    """map.GramToken (:187-208) parameter holder."""

    def __init__(self, ch_dim, num_groups, num_tokens, bp_dim, out_dim):
        super().__init__()
        tri = torch.triu_indices(bp_dim, bp_dim)
        self.register_buffer('bp_index', tri[0] * bp_dim + tri[1])
        self.num_groups, self.num_tokens, self.bp_dim, self.out_dim = num_groups, num_tokens, bp_dim, out_dim
        self.gram_dim = bp_dim * (bp_dim + 1) // 2
        self.ch_reduction = nn.Sequential(nn.Conv2d(ch_dim, bp_dim, 1, bias=False), nn.BatchNorm2d(bp_dim))
        self.bp_reduction = nn.Sequential(nn.Conv2d(self.gram_dim, out_dim * num_tokens, 1, bias=False, groups=num_groups),
                                          nn.BatchNorm2d(out_dim * num_tokens))


class CAP(nn.Module):
    def __init__(self, last_dim, num_heads, mlp_ratio, mlp_groups, n_tokens, gram_group, bp_dim, ca_dim):
        super().__init__()
        self.T = n_tokens
        self.dim = last_dim * (n_tokens + 1)
        self.attention = nn.Sequential(CABlock(last_dim, num_heads, mlp_ratio, mlp_groups, ca_dim))
        self.gram_token_extraction = GramToken(last_dim, gram_group, n_tokens, bp_dim, last_dim)


class MultiScale(nn.Module):
    def __init__(self, channels, out_dim):
        super().__init__()
        self.concat_conv = nn.Sequential(nn.Conv2d(sum(channels), out_dim, 1, bias=False), nn.BatchNorm2d(out_dim), nn.GELU())


class MAP(nn.Module):
    def __init__(self, channels, last_dim, num_heads, mlp_ratio, mlp_groups, n_tokens, n_groups, gram_group, bp_dim, ca_dim):
        super().__init__()
        self.mmcap = nn.ModuleList([CAP(last_dim, num_heads, mlp_ratio, mlp_groups, n_tokens, gram_group, bp_dim, ca_dim)
                                    for _ in range(n_groups)])
        self.multi_scale = MultiScale(channels, last_dim)


class NormHead(nn.Module):
    """map.NormHead (:393-412): LayerNorm (eps 1e-5) -> Linear."""

    def __init__(self, ch, num_classes):
        super().__init__()
        self.norm = nn.LayerNorm(ch)
        self.head = nn.Linear(ch, num_classes)

    def run(self, x, T):
        h = ops.layernorm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        return ops.linear(ops.to_dtype(h, T), self.head.weight, self.head.bias, out_dtype=torch.float32)


class MAPHead(nn.Module):
    """map.MAPHead (:462-539) for the ConvNeXt configuration: multi-scale level 3, gram tokens, self-distillation token."""

    def __init__(self, channels, last_dim=384, num_heads=12, n_tokens=2, n_groups=4, gram_group=24, bp_dim=384, ca_dim=384,
                 mlp_ratio=4, mlp_groups=2, num_classes=1000):
        super().__init__()
        self.n_groups, self.n_tokens, self.last_dim = n_groups, n_tokens, last_dim
        self.out_ch = last_dim * n_tokens
        self.mmcap = MAP(channels, last_dim, num_heads, mlp_ratio, mlp_groups, n_tokens, n_groups, gram_group, bp_dim, ca_dim)
        self.heads = nn.ModuleList([NormHead(last_dim * n_tokens, num_classes) for _ in range(n_groups)])
        self.self_dt_heads = nn.ModuleList([NormHead(last_dim, num_classes) for _ in range(n_groups)])
        self.drop, self.attn_drop = 0.05, 0.05          # CABlock defaults (map.py:149); train mode only

    def run(self, feats, geoms, T, training):
        """feats: compute-dtype row matrices [stem, s0, s1, s2, s3] with geoms (B,H,W).  -> list of logits (pairs in train)."""
        Bn, Ho, Wo = geoms[3]
        HW = Ho * Wo
        items = []
        for f, (b, h, w) in zip(feats, geoms):
            mode = 1 if h == Ho else (3 if h > Ho else 4)        # MultiScale.forward, map.py:322-331
            items.append((h, w, f.shape[1], mode))
        cat = ops.aggregate((Bn, Ho, Wo, items), list(feats))
        ms = self.mmcap.multi_scale.concat_conv
        L_ = self.last_dim
        f = ops.linear(cat, ms[0].weight.reshape(L_, -1))
        f = ops.gelu(ops.batchnorm(f, _bn(ms[1]), training))
        caps: List[CAP] = list(self.mmcap.mmcap)
        nb, nt, nq = len(caps), self.n_tokens, self.n_tokens + 1
        blk0 = caps[0].attention[0]
        E, heads = blk0.attn.embed_dim, blk0.attn.num_heads
        # token side, batched over groups: shared normalisation (norm1's affine folded into k/v) and one k/v projection
        fhat = ops.layernorm(f, None, None, blk0.norm1.eps)
        wkv, bkv = [], []
        for c in caps:
            a = c.attention[0]
            w = torch.cat((a.attn.k.weight, a.attn.v.weight), 0)
            wkv.append(w * a.norm1.weight[None, :])
            bkv.append((w * a.norm1.bias[None, :]).sum(-1) + torch.cat((a.attn.k.bias, a.attn.v.bias), 0))   # W b without a library GEMV
        kv_tok = ops.linear(fhat, torch.cat(wkv, 0), torch.cat(bkv, 0), out_dtype=torch.float32)   # [B*HW, nb*2E]; fp32 like GA_ConvNeXt._heads
        cls_all, qs, kvcs = [], [], []
        for c in caps:
            gt, a = c.gram_token_extraction, c.attention[0]
            r = ops.linear(f, gt.ch_reduction[0].weight.reshape(gt.bp_dim, L_))
            r = ops.batchnorm(r, _bn(gt.ch_reduction[1]), training)
            G, glen = gt.num_groups, gt.bp_reduction[0].weight.shape[1]
            # x/(hw) then X X^T (map.py:217-218): alpha = 1/(hw)^2 = 1/(div^2 * hw) with div = sqrt(hw)
            t = ops.gram_embed(r, gt.bp_reduction[0].weight.view(G, L_ * nt // G, glen), None, Bn, HW, float(HW) ** 0.5, interleave=nt)
            t = ops.batchnorm(t, _bn(gt.bp_reduction[1]), training)               # [B, L*nt] fp32, channel c*nt + j = token j
            tok = t.view(Bn, L_, nt).permute(0, 2, 1)
            cls = torch.cat((tok, tok.mean(dim=1, keepdim=True)), dim=1).contiguous()   # + self-distillation token
            cn = ops.layernorm(cls.view(Bn * nq, L_), a.norm1.weight, a.norm1.bias, a.norm1.eps)
            cn_t = ops.to_dtype(cn, T)
            qs.append(ops.linear(cn_t, a.attn.q.weight * a.attn.scale, a.attn.q.bias * a.attn.scale, out_dtype=torch.float32))
            kvcs.append(ops.linear(cn_t, torch.cat((a.attn.k.weight, a.attn.v.weight), 0),
                                   torch.cat((a.attn.k.bias, a.attn.v.bias), 0), out_dtype=torch.float32))
            cls_all.append(cls)
        q = torch.stack(qs).view(nb, Bn, nq, E)
        kvc = torch.stack(kvcs).view(nb, Bn, nq, 2 * E)
        p_att = self.attn_drop if training else 0.0
        p_drop = self.drop if training else 0.0
        mask = ops.dropout_mask((nb, Bn, heads, nq, nq + HW), p_att, f.device) if p_att > 0.0 else None
        o = ops.attnpool(q, kvc, kv_tok, HW, heads, mask)                          # [nb, B, nq, E]
        out = []
        for g, c in enumerate(caps):
            a = c.attention[0]
            cls = cls_all[g].view(Bn * nq, L_)
            pr = ops.linear(ops.to_dtype(o[g].reshape(Bn * nq, E), T), a.attn.proj.weight, a.attn.proj.bias, out_dtype=torch.float32)
            if p_drop > 0.0:
                pr = torch.nn.functional.dropout(pr, p_drop, True)                 # proj_drop (map.py:141)
            cls = cls + pr
            h = ops.layernorm(cls, a.norm2.weight, a.norm2.bias, a.norm2.eps)
            cls = cls + a.mlp.run(h, T, drop=p_drop)
            pool = cls.view(Bn, nq * L_)
            main = self.heads[g].run(pool[:, :self.out_ch].contiguous(), T)
            if training:
                out.append([main, self.self_dt_heads[g].run(pool[:, self.out_ch:].contiguous(), T)])
            else:
                out.append(main)
        return out


def _bn(m: nn.BatchNorm2d) -> dict:
    return {'weight': m.weight, 'bias': m.bias, 'running_mean': m.running_mean, 'running_var': m.running_var,
            'num_batches_tracked': m.num_batches_tracked}


class ConvNeXt(nn.Module):
    """map_convnext.ConvNeXt (:43-140) with global_pool='mmcap'."""

    def __init__(self, in_chans=3, num_classes=1000, depths=(3, 3, 9, 3), dims=(96, 192, 384, 768), drop_path_rate=0.,
                 layer_scale_init_value=1e-6, head_init_scale=1., global_pool='mmcap', last_dim=384, n_groups=4, n_tokens=3,
                 gram_group=8, bp_dim=192, ca_dim=128, num_heads=8, **unused):
        super().__init__()
        assert global_pool == 'mmcap' and in_chans == 3, 'only the MAP-head configuration is on the B200 path'
        depths, dims = list(depths), list(dims)
        self.num_classes = num_classes
        self.compute_dtype = None
        self.default_cfg = self.pretrained_cfg = {'input_size': (3, 224, 224), 'num_classes': num_classes, 'crop_pct': 0.875,
                                                  'interpolation': 'bicubic', 'mean': (0.485, 0.456, 0.406),
                                                  'std': (0.229, 0.224, 0.225)}
        self.downsample_layers = nn.ModuleList()
        self.downsample_layers.append(nn.Sequential(nn.Conv2d(in_chans, dims[0], kernel_size=4, stride=4),
                                                    LayerNorm(dims[0], eps=1e-6, data_format='channels_first')))
        for i in range(3):
            self.downsample_layers.append(nn.Sequential(LayerNorm(dims[i], eps=1e-6, data_format='channels_first'),
                                                        nn.Conv2d(dims[i], dims[i + 1], kernel_size=2, stride=2)))
        self.stages = nn.ModuleList()
        dp = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        cur = 0
        for i in range(4):
            self.stages.append(nn.Sequential(*[Block(dims[i], dp[cur + j], layer_scale_init_value) for j in range(depths[i])]))
            cur += depths[i]
        self.global_pool = global_pool
        self.norm = nn.Identity()
        self.head = MAPHead([dims[0]] + dims, last_dim=last_dim, num_heads=num_heads, n_tokens=n_tokens, n_groups=n_groups,
                            gram_group=gram_group, bp_dim=bp_dim, ca_dim=ca_dim, num_classes=num_classes)
        _apply_children_first(_init_weights, self)

    def _dtype(self):
        if self.compute_dtype is not None:
            return self.compute_dtype
        return torch.bfloat16 if torch.is_autocast_enabled() else torch.float32

    def forward_features(self, x, T=None):
        """-> ([stem, s0, s1, s2, s3] compute-dtype rows, their (B,H,W) geometries), map_convnext.py:124-135."""
        if not x.is_cuda:
            raise ops.L.GaError('MAP-ConvNeXt runs on CUDA (sm_100a) tensors only: there is no CPU path')
        T, RT = (T or self._dtype()), torch.float32
        Bn, _, H, W = x.shape
        feats, geoms = [], []
        y = ys = None
        geom = None
        # per-forward work shared by every block: one DropPath draw, one operand-preparation launch
        all_blocks = [blk for i in range(4) for blk in self.stages[i]]
        all_scales = _path_scales([blk.drop_prob for blk in all_blocks], self.training, Bn, x.device)
        all_preps = None
        if T == torch.bfloat16 and all(blk.norm.weight.numel() % 8 == 0 for blk in all_blocks):
            bw = _BLOCK_WEIGHTS.get(self)
            if bw is None:
                bw = _BLOCK_WEIGHTS[self] = ops.BlockWeights(lambda: [blk.params() for blk in all_blocks])
            bw.refresh()
            all_preps = [bw.get(j) for j in range(len(all_blocks))]
        off = 0
        for i in range(4):
            ds = self.downsample_layers[i]
            if i == 0:
                conv, ln = ds[0], ds[1]
                rows = ops.stem_patchify(x.float(), 4, T)
                y = ops.linear(rows, conv.weight.permute(0, 2, 3, 1).reshape(conv.out_channels, -1), conv.bias, out_dtype=RT)
                y = ops.layernorm(y, ln.weight, ln.bias, ln.eps)
                geom = (Bn, H // 4, W // 4)
                ys = ops.to_dtype(y, T) if RT != T else None
                feats.append(ys if ys is not None else y)
                geoms.append(geom)
            else:
                ln, conv = ds[0], ds[1]
                h = ops.layernorm(ys if ys is not None else y, ln.weight, ln.bias, ln.eps)
                h = ops.patchify(h, (geom[0], geom[1], geom[2], conv.in_channels), 2)
                geom = (Bn, geom[1] // 2, geom[2] // 2)
                y = ops.linear(h, conv.weight.permute(0, 2, 3, 1).reshape(conv.out_channels, -1), conv.bias, out_dtype=RT)
                ys = ops.to_dtype(y, T) if RT != T else None
            blocks = list(self.stages[i])
            scales = all_scales[off:off + len(blocks)]                                              # DropPath factors of the stage
            for j, blk in enumerate(blocks):
                y, ys = blk.run(y, ys, geom, T, ps=scales[j], ps_prev=scales[j - 1] if j > 0 else None,
                                prep=all_preps[off + j] if all_preps is not None else None)
            off += len(blocks)
            feats.append(ys if ys is not None else y)
            geoms.append(geom)
        return feats, geoms

    def forward(self, x):
        T = self._dtype()                                   # read the autocast state before disabling it for the glue ops
        with torch.autocast('cuda', enabled=False), ops.collect_bn_counters():
            feats, geoms = self.forward_features(x, T)
            return self.head.run(feats, geoms, T, self.training)


@register_model
def map_convnext_tiny(pretrained=False, in_22k=False, **kwargs):
    kwargs.pop('pretrained_cfg', None)
    kwargs.pop('pretrained_cfg_overlay', None)
    assert not pretrained, 'release checkpoints are not reachable from this image; use checkpoint_path'
    return ConvNeXt(depths=[3, 3, 9, 3], dims=[96, 192, 384, 768], global_pool='mmcap', last_dim=384, n_groups=4, n_tokens=2,
                    gram_group=24, bp_dim=384, ca_dim=384, num_heads=12, **kwargs)


@register_model
def map_convnext_small(pretrained=False, in_22k=False, **kwargs):
    kwargs.pop('pretrained_cfg', None)
    kwargs.pop('pretrained_cfg_overlay', None)
    assert not pretrained, 'release checkpoints are not reachable from this image; use checkpoint_path'
    return ConvNeXt(depths=[3, 3, 27, 3], dims=[96, 192, 384, 768], global_pool='mmcap', last_dim=384, n_groups=4, n_tokens=3,
                    gram_group=16, bp_dim=384, ca_dim=384, num_heads=12, **kwargs)
