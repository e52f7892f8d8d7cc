"""Kernel parity AT THE BENCHMARK SHAPES (BASELINE config 2: batch 256 per GPU): one ConvNeXt block forward + backward at
stage 0 (56 x 56 x 96, M = 802 816 rows: 6 272 GEMM m-tiles, one-wave split-K weight gradients, the fused LayerNorm-backward
epilogue, the tile plans the bench actually runs) and at stage 3 (7 x 7 x 688), against an fp32 PyTorch evaluation of the same
block (GA/ga_convnext.py:98-112: dw7x7 -> LayerNorm -> fc1 -> GELU -> fc2 -> gamma -> + x).  No ReLU on this path, so outputs and
every gradient are compared directly; the north-star bound for bf16 is 2e-2."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from imagenet_models_b200 import ops  # noqa: E402

DEV = 'cuda'


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize('C,HW', [(96, 56), (688, 7)])
def test_convnext_block_at_bench_batch_vs_fp32_torch(C, HW):
    B = 256
    M = B * HW * HW
    g = torch.Generator().manual_seed(100 + C)

    def rn(*shape, s=1.0):
        return (torch.randn(*shape, generator=g) * s).to(DEV)
    P = {'conv_dw.weight': rn(C, 1, 7, 7, s=0.15), 'conv_dw.bias': rn(C, s=0.1), 'norm.weight': 1.0 + rn(C, s=0.1), 'norm.bias': rn(C, s=0.1),
         'mlp.fc1.weight': rn(4 * C, C, s=C ** -0.5), 'mlp.fc1.bias': rn(4 * C, s=0.1), 'mlp.fc2.weight': rn(C, 4 * C, s=(4 * C) ** -0.5),
         'mlp.fc2.bias': rn(C, s=0.1), 'gamma': 0.2 + rn(C, s=0.05)}
    P = {k: v.requires_grad_(True) for k, v in P.items()}
    # images differ in scale and offset (i.i.d. noise of one scale would make every row statistically alike)
    x4 = torch.randn(B, HW, HW, C, generator=g) * (0.5 + torch.rand(B, 1, 1, 1, generator=g)) + 0.3 * torch.randn(B, 1, 1, C, generator=g)
    x = x4.reshape(M, C).to(DEV).contiguous().requires_grad_(True)
    gy = rn(M, C)
    xs = x.detach().to(torch.bfloat16)
    y, ys = ops.convnext_block(x, P, (B, HW, HW), None, True, xs=xs, T=torch.bfloat16)
    assert y.dtype == torch.float32 and ys.dtype == torch.bfloat16
    y.backward(gy)
    torch.cuda.synchronize()
    got = {k: v.grad.detach().clone() for k, v in P.items()}
    got_dx, got_y = x.grad.detach().clone(), y.detach()
    # ---- fp32 PyTorch evaluation of the same block on the same operands
    Pr = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    xr = x.detach().clone().requires_grad_(True)
    h = F.conv2d(xr.view(B, HW, HW, C).permute(0, 3, 1, 2), Pr['conv_dw.weight'], Pr['conv_dw.bias'], padding=3, groups=C)
    h = F.layer_norm(h.permute(0, 2, 3, 1).reshape(M, C), (C,), Pr['norm.weight'], Pr['norm.bias'], 1e-6)
    h = F.gelu(F.linear(h, Pr['mlp.fc1.weight'], Pr['mlp.fc1.bias']))
    yr = xr + Pr['gamma'] * F.linear(h, Pr['mlp.fc2.weight'], Pr['mlp.fc2.bias'])
    yr.backward(gy)
    torch.cuda.synchronize()
    TOL = 2e-2
    errs = {'branch': rel(got_y - x.detach(), yr.detach() - xr.detach()), 'dx_branch': rel(got_dx - gy, xr.grad - gy)}
    for k in P:
        errs[k] = rel(got[k], Pr[k].grad)
    errs['y'], errs['shadow'] = rel(got_y, yr.detach()), rel(ys.float(), yr.detach())
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    if os.path.isdir(out):                                               # measured values for profiles/ (scratch directory of a GPU run)
        json.dump({'B': B, 'H': HW, 'C': C, 'rows': M, 'rel_err_vs_fp32_torch': errs}, open(os.path.join(out, f'bench_shape_block_C{C}.json'), 'w'), indent=1)
    assert errs['y'] <= 1e-2, errs
    assert errs['shadow'] <= 1e-2, errs                                   # the bf16 shadow of the stream
    bad = {k: v for k, v in errs.items() if k not in ('y', 'shadow') and not v <= TOL}
    assert not bad, (bad, errs)


@pytest.mark.parametrize('R,split,C', [(14, 7, 256), (56, 1, 64)])
def test_cswin_stripe_attention_at_bench_batch_vs_oracle(R, split, C):
    """K6 (LePE stripe attention, ga_cswin.py:110-136) at GA-CSWin-T's batch 128: stage 3 (14 x 14 tokens, 7-wide stripes, 256
    channels; the tcgen05 forward) and stage 1 (56 x 56, 1-wide stripes, 64 channels), both branches, forward + backward in bf16
    against the fp32 oracle evaluated on the same (bf16-rounded) operands on the GPU."""
    from oracle import ga_cswin_oracle as CO
    B, nbr = 128, 2
    g = torch.Generator().manual_seed(R * 100 + C)
    qkv = torch.randn(B, R * R, 3 * C, generator=g).to(DEV)
    dy = torch.randn(B, R * R, C, generator=g).to(DEV)
    cb = C // nbr
    P = {f'{i}.get_v.weight': (torch.randn(cb, 1, 3, 3, generator=g) * 0.3).to(DEV) for i in range(nbr)}
    P.update({f'{i}.get_v.bias': (torch.randn(cb, generator=g) * 0.1).to(DEV) for i in range(nbr)})
    P = {k: v.requires_grad_(True) for k, v in P.items()}
    qo = qkv.to(torch.bfloat16).float().requires_grad_(True)
    q, k, v = qo[..., :C], qo[..., C:2 * C], qo[..., 2 * C:]
    h = C // 2
    o = torch.cat((CO.lepe_attention(P, '0.', q[..., :h], k[..., :h], v[..., :h], R, R, split, h // 32),
                   CO.lepe_attention(P, '1.', q[..., h:], k[..., h:], v[..., h:], R, split, R, h // 32)), 2)
    o.backward(dy.to(torch.bfloat16).float())
    lw = torch.cat([P[f'{i}.get_v.weight'].detach().reshape(cb, 9) for i in range(nbr)]).requires_grad_(True)
    lb = torch.cat([P[f'{i}.get_v.bias'].detach() for i in range(nbr)]).requires_grad_(True)
    qg = qkv.reshape(-1, 3 * C).to(torch.bfloat16).requires_grad_(True)
    og = ops.cswin_attention(qg, lw, lb, B, R, split, nbr)
    og.backward(dy.reshape(-1, C).to(torch.bfloat16))
    torch.cuda.synchronize()
    dw = torch.cat([P[f'{i}.get_v.weight'].grad.reshape(cb, 9) for i in range(nbr)])
    db = torch.cat([P[f'{i}.get_v.bias'].grad for i in range(nbr)])
    errs = {'out': rel(og.detach().float().reshape(o.shape), o.detach()), 'dqkv': rel(qg.grad.float().reshape(qo.shape), qo.grad),
            'dlepe_w': rel(lw.grad, dw), 'dlepe_b': rel(lb.grad, db)}
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    if os.path.isdir(out):
        json.dump({'B': B, 'R': R, 'split': split, 'C': C, 'rel_err_vs_fp32_oracle': errs}, open(os.path.join(out, f'bench_shape_attn_R{R}.json'), 'w'), indent=1)
    bad = {k_: v_ for k_, v_ in errs.items() if not v_ <= 2e-2}
    assert not bad, (bad, errs)
