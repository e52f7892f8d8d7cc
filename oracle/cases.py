"""Seeded parity cases shared by tests/golden/make_golden.py and tests/ (TEST INFRASTRUCTURE).

Inputs and weights are regenerated from these seeds on both sides, so the committed fixtures only hold
the reference's outputs.  CPU torch RNG is deterministic for a fixed torch build (same image on the GPU box).
"""
import math
import zlib

import torch

STATE_SEED = 7
GA_LAM = -0.8          # --GA_lam used by the published recipe (GA/README.md:26)
GA_MODEL_CASES = [('ga_convnext_tiny_688', 2)]
MAP_MODEL_CASES = [('map_convnext_tiny', 2)]
MAP_DEC_LAM = -0.8
PARAM_COUNTS = {
    'ga_convnext_tiny_688': 47821324, 'ga_convnext_tiny_768': 54354584,
    'ga_convnext_small_688': 70116364, 'ga_convnext_small_768': 76726424,
    'ga_convnext_base_976': 123688396, 'ga_convnext_base_1024': 128839176,
}

# name -> (C, H(=W), B)
BLOCK_CASES = {'c32_h9': (32, 9, 2), 'c96_h14': (96, 14, 2), 'c192_h14': (192, 14, 1), 'c688_h7': (688, 7, 1)}
GRAM_CASES = {'c192_h14': (192, 14, 3), 'c24_h5': (24, 5, 2)}
# GA-CSWin: name -> (dim, reso, split, heads, last_stage, B): every stripe geometry of the T configuration
CSWIN_BLOCK_CASES = {'s1_c64_r56_sp1': (64, 56, 1, 2, False, 1), 's2_c128_r28_sp2': (128, 28, 2, 4, False, 2),
                     's3_c256_r14_sp7': (256, 14, 7, 8, False, 2), 's4_c512_r7_last': (512, 7, 7, 16, True, 2),
                     'gram_c192_r14_sp7': (192, 14, 7, 6, False, 3)}
CSWIN_MODEL_CASES = [('ga_cswin_test', 2), ('ga_CSWin_64_12211_tiny_224', 2)]
# name -> (C, dim_embed, N tokens, B)
CLASSATTN_CASES = {'c688_e168': (688, 168, 196, 2), 'c64_e32': (64, 32, 10, 3)}


def _gen(seed, name):
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))


def _fill(shapes, seed):
    P = {}
    for name, (shape, kind) in shapes.items():
        g = _gen(seed, name)
        if kind == 'w':
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            P[name] = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        elif kind == 'b':
            P[name] = torch.randn(shape, generator=g) * 0.1
        else:
            P[name] = 0.5 + torch.rand(shape, generator=g)
    return P


def ga_inputs(B, seed=42, size=224):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, size, size, generator=g), torch.randint(0, 1000, (B,), generator=g)


def ga_inputs_diverse(B, seed=42, size=224):
    """Images that differ from one another (per-image sinusoid patterns, colour offsets, noise levels): with i.i.d. noise
    images every sample has the same statistics, train-mode BatchNorm over the batch then divides by a vanishing standard
    deviation and amplifies rounding ~25x (gram_embedding BN over [B, C, 1, 1]); real images are not like that."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, size), torch.linspace(-1, 1, size), indexing='ij')
    imgs = []
    for _ in range(B):
        f = torch.rand(3, 2, generator=g) * 12 + 1
        ph = torch.rand(3, generator=g) * 6.28
        amp = torch.rand(3, generator=g) * 1.5 + 0.2
        mean = torch.randn(3, generator=g) * 0.7
        noise = torch.randn(3, size, size, generator=g) * (0.1 + torch.rand(1, generator=g) * 0.9)
        pat = torch.stack([amp[c] * torch.sin(f[c, 0] * xx * 3.14 + f[c, 1] * yy * 3.14 + ph[c]) for c in range(3)])
        imgs.append(pat + mean[:, None, None] + noise)
    return torch.stack(imgs), torch.randint(0, 1000, (B,), generator=g)


# whole-model parity cases at the BASELINE.json shapes: key -> (model, B, state profile, input kind)
#   config1: BASELINE config 1 (fp32 training step, batch 8; the fp64 Gram branch of get_gram is active: training and B < 128)
#   bf16:    the bf16 training fixture, batch 16
# Both use mutually different images and trained-magnitude layer scales (see ga_inputs_diverse / make_state(profile='trained')).
GA_PARITY_CASES = {'config1': ('ga_convnext_tiny_688', 8, 'trained', 'diverse'),
                   'bf16': ('ga_convnext_tiny_688', 16, 'trained', 'diverse')}
# parameters whose gradient does not pass through any Bottleneck ReLU (everything after stages.4): compared with the reference
# directly; the rest is compared with the oracle evaluated at the implementation's own ReLU decisions (oracle _relu docstring)
GA_TAIL_PREFIXES = ('gram_contraction.', 'gram_layer.', 'gram_embedding.', 'ga.', 'fc.')


MAP_PARITY_CASES = {'map': ('map_convnext_tiny', 8, 'trained', 'diverse')}
CSWIN_PARITY_CASES = {'cswin': ('ga_CSWin_64_12211_tiny_224', 8, 'trained', 'diverse')}
MAP_TAIL_PREFIXES = ('head.heads.', 'head.self_dt_heads.')      # after the CABlock MLP's ReLU
CSWIN_TAIL_PREFIXES = ('',)                                      # GA-CSWin has no ReLU at all: every gradient is compared raw


def parity_inputs(kind, B, size=224):
    return ga_inputs_diverse(B, size=size) if kind == 'diverse' else ga_inputs(B, size=size)


def block_state(C, seed=STATE_SEED):
    return _fill({'conv_dw.weight': ((C, 1, 7, 7), 'w'), 'conv_dw.bias': ((C,), 'b'),
                  'norm.weight': ((C,), 'g'), 'norm.bias': ((C,), 'b'),
                  'mlp.fc1.weight': ((4 * C, C), 'w'), 'mlp.fc1.bias': ((4 * C,), 'b'),
                  'mlp.fc2.weight': ((C, 4 * C), 'w'), 'mlp.fc2.bias': ((C,), 'b'),
                  'gamma': ((C,), 'g')}, seed)


def block_inputs(C, H, B, seed=11):
    g = torch.Generator().manual_seed(seed + C * 131 + H)
    return torch.randn(B, C, H, H, generator=g), torch.randn(B, C, H, H, generator=g)


def cswin_block_inputs(dim, reso, B, seed=19):
    g = torch.Generator().manual_seed(seed + dim * 131 + reso)
    return torch.randn(B, reso * reso, dim, generator=g), torch.randn(B, reso * reso, dim, generator=g)


def gram_input(C, H, B, seed=13):
    g = torch.Generator().manual_seed(seed + C)
    return torch.randn(B, C, H, H, generator=g)


def ga_block_state(C, E, seed=STATE_SEED, groups=4):
    return _fill({'norm1.weight': ((C,), 'g'), 'norm1.bias': ((C,), 'b'),
                  'attn.q.weight': ((E, C), 'w'), 'attn.k.weight': ((E, C), 'w'), 'attn.v.weight': ((E, C), 'w'),
                  'attn.proj.weight': ((C, E), 'w'), 'attn.proj.bias': ((C,), 'b'),
                  'norm2.weight': ((C,), 'g'), 'norm2.bias': ((C,), 'b'),
                  'mlp.fc1.weight': ((4 * C, C // groups, 1, 1), 'w'), 'mlp.fc1.bias': ((4 * C,), 'b'),
                  'mlp.fc2.weight': ((C, 4 * C // groups, 1, 1), 'w'), 'mlp.fc2.bias': ((C,), 'b'),
                  'gamma_1': ((C,), 'g'), 'gamma_2': ((C,), 'g')}, seed)


def ga_block_inputs(C, N, B, seed=17):
    g = torch.Generator().manual_seed(seed + C)
    return (torch.randn(B, N, C, generator=g), torch.randn(B, C, generator=g), torch.randn(B, C, generator=g))


def digest(t, n=2048):
    """(L2 norm, strided sample) of a tensor: the committed form of large reference gradients."""
    flat = t.detach().reshape(-1)
    stride = max(1, flat.numel() // n)
    return (flat.double().norm().item(), flat[::stride][:n].clone().float(), stride)


def digest_close(t, dig, rtol, atol=3e-4):
    """Compare a tensor with a digest made by digest(): norm and sampled entries.

    atol absorbs parameters whose true gradient is exactly zero (conv biases feeding a train-mode
    BatchNorm): both sides then hold rounding noise only.
    """
    norm, sample, stride = dig
    flat = t.detach().reshape(-1).float().cpu()
    mine = flat[::stride][:sample.numel()]
    err = (mine.double() - sample.double()).norm().item()
    ok_sample = err <= rtol * sample.double().norm().item() + atol
    ok_norm = abs(flat.double().norm().item() - norm) <= rtol * norm + atol
    return ok_sample and ok_norm


def digest_rel_err(t, dig):
    """Relative L2 error of a tensor against the sampled entries of a digest made by digest()."""
    norm, sample, stride = dig
    flat = t.detach().reshape(-1).float().cpu()
    mine = flat[::stride][:sample.numel()]
    return (mine.double() - sample.double()).norm().item() / max(sample.double().norm().item(), 1e-30)
