"""Flat-buffer training state: fused AdamW + EMA step (K7) and bucketed gradient all-reduce.

Reference behaviour being replaced (GA/train.py:466,499,505-515,758-761):
  * `create_optimizer_v2(model, opt='adamw', lr, weight_decay, filter_bias_and_bn=True)` -> torch AdamW over 365 tensors,
    no decay on 1-D parameters / biases;
  * `ModelEmaV2(model, decay)` -> python loop `ema = d*ema + (1-d)*model` over all 407 state_dict entries;
  * `DistributedDataParallel` -> bucketed NCCL all-reduce(avg) of gradients overlapped with backward.
Here every parameter (and its gradient, Adam moments and EMA copy) is a view into ONE flat fp32 buffer, so the
optimizer + EMA is a single kernel launch and gradient buckets are contiguous slices handed to NCCL as they fill.
"""
from __future__ import annotations

import copy
from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import lib as L

SEG_SHIFT = 4           # weight-decay flags are per 16-element segment; tensors are padded to that
SEG = 1 << SEG_SHIFT


def _round_up(n, m):
    return (n + m - 1) // m * m


class FlatState:
    """Re-homes a module's parameters into one flat fp32 buffer (+ flat grad)."""

    def __init__(self, model: nn.Module):
        params = [p for p in model.parameters() if p.requires_grad]
        assert params and all(p.dtype == torch.float32 for p in params), 'fp32 parameters expected'
        dev = params[0].device
        self.names = [n for n, p in model.named_parameters() if p.requires_grad]
        self.params = params
        self.offsets, off = [], 0
        for p in params:
            self.offsets.append(off)
            off += _round_up(p.numel(), SEG)
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self._tables = {}
        for p, o in zip(params, self.offsets):
            n = p.numel()
            self.flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + n].view(p.shape)
            p.grad = None
        # float buffers (BatchNorm running statistics) re-homed the same way, for a one-launch EMA of non-parameter state
        self.bufflat = _flatten_buffers(model, dev)
        # bf16 shadow of every parameter at the same offsets: written by the optimizer kernel together with the fp32 update,
        # so no forward or backward GEMM casts a weight again (ops.cast_like looks the parameter up)
        self.flat16 = None
        if dev.type == 'cuda':
            self.flat16 = torch.empty(off, dtype=torch.bfloat16, device=dev)
            self.refresh_shadows()

    def refresh_shadows(self):
        """Re-cast every parameter and re-register the shadows (construction; after parameters were written from outside the
        optimizer, e.g. load_state_dict, whose version bump had un-registered them)."""
        if self.flat16 is None:
            return
        from . import ops
        L.check(L.load().ga_cast_bf16(L.ptr(self.flat), L.ptr(self.flat16), L.ll(self.numel), L.stream()), 'ga_cast_bf16')
        ops.register_weight_shadows(self.params, self.offsets, self.flat16)

    def shadows_current(self):
        if self.flat16 is None:
            return True
        from . import ops
        return all((e := ops._weight_shadows.get(p.data_ptr())) is not None and e[1] == p._version for p in self.params)

    def decay_flags(self, no_decay_1d: bool = True) -> torch.Tensor:
        """timm filter_bias_and_bn: 1-D tensors and biases get no weight decay."""
        flags = torch.zeros(self.numel >> SEG_SHIFT, dtype=torch.uint8)
        for name, p, o in zip(self.names, self.params, self.offsets):
            decay = not (no_decay_1d and (p.ndim <= 1 or name.endswith('.bias')))
            if decay:
                flags[o >> SEG_SHIFT:(o + _round_up(p.numel(), SEG)) >> SEG_SHIFT] = 1
        return flags.to(self.flat.device)

    def zero_grad(self):
        """Drop every .grad: autograd then *adopts* the gradient tensors our backward kernels produce instead of launching
        one accumulate kernel per parameter; gather() collects them into the flat buffer in one launch."""
        for p in self.params:
            p.grad = None
        from . import ops
        ops.clear_shadows()
        ops.ZEROS.begin(self.flat.device)       # one zero fill for every gradient accumulator of the coming backward

    def _table(self, members):
        key = tuple(members)
        ent = self._tables.get(key)
        if ent is None:
            import numpy as np
            rec = np.zeros((len(members), 4), dtype=np.int64)
            chunk = 0
            for r, idx in enumerate(members):
                n = self.params[idx].numel()
                rec[r, 1], rec[r, 2], rec[r, 3] = self.offsets[idx], n, chunk
                chunk += (n + 4095) // 4096
            ent = {'rec': rec, 'chunks': chunk, 'dev': torch.empty(len(members) * 4, dtype=torch.int64, device=self.flat.device),
                   'graph_host': torch.empty(rec.size, dtype=torch.int64, pin_memory=True)}
            self._tables[key] = ent
        return ent

    def gather(self, members=None):
        """p.grad of `members` (default: all) -> their slices of the flat gradient (ga_gather_grads); absent grads give zeros."""
        members = range(len(self.params)) if members is None else members
        ent = self._table(members)
        rec = ent['rec']
        for r, idx in enumerate(members):
            g = self.params[idx].grad
            if g is None:
                rec[r, 0] = 0
                continue
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous()
                self.params[idx].grad = g
            rec[r, 0] = g.data_ptr()
        if torch.cuda.is_current_stream_capturing():
            host = ent['graph_host']                    # the captured copy node re-reads this buffer at every replay
        else:
            host = torch.empty(rec.size, dtype=torch.int64, pin_memory=True)   # caching host allocator: reuse is stream-safe
        host.numpy()[:] = rec.reshape(-1)
        ent['dev'].copy_(host, non_blocking=True)
        L.check(L.load().ga_gather_grads(L.ptr(ent['dev']), len(rec), L.ll(ent['chunks']), L.ptr(self.grad), L.stream()),
                'ga_gather_grads')


def _flatten_buffers(model: nn.Module, dev) -> torch.Tensor:
    bufs = [b for b in model.buffers() if b.is_floating_point()]
    offs, off = [], 0
    for b in bufs:
        offs.append(off)
        off += _round_up(b.numel(), 4)
    flat = torch.zeros(max(off, 4), dtype=torch.float32, device=dev)
    for b, o in zip(bufs, offs):
        flat[o:o + b.numel()].copy_(b.data.reshape(-1))
        b.data = flat[o:o + b.numel()].view(b.shape)
    return flat


class FusedAdamWEma:
    """torch.optim.AdamW semantics + timm ModelEmaV2, one kernel over the flat state (ga_adamw_ema)."""

    def __init__(self, model: nn.Module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05, ema_decay: Optional[float] = None,
                 filter_bias_and_bn=True):
        self.model = model
        self.state = FlatState(model)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.m = torch.zeros_like(self.state.flat)
        self.v = torch.zeros_like(self.state.flat)
        self.flags = self.state.decay_flags(filter_bias_and_bn) if weight_decay else None
        self.step_count = 0
        self.ema_decay = ema_decay
        self.ema_flat = self.ema_model = None
        if ema_decay is not None:
            self.ema_flat = self.state.flat.clone()
            self.ema_model = copy.deepcopy(model).eval()          # same layout -> same offsets
            for p in self.ema_model.parameters():
                p.requires_grad_(False)
            ema_named = dict(self.ema_model.named_parameters())      # same names as the model; frozen parameters are not in the flat state
            for name, o in zip(self.state.names, self.state.offsets):
                p = ema_named[name]
                p.data = self.ema_flat[o:o + p.numel()].view(p.shape)
            self.ema_bufflat = _flatten_buffers(self.ema_model, self.ema_flat.device)
        self.clip_grad = None                                     # global-norm clip (set by TrainEngine(clip_grad=...))
        self._clip_scratch = torch.zeros(2, dtype=torch.float32, device=self.state.flat.device)
        self.param_groups = [{'lr': lr}]                          # what the reference's logging / schedulers read
        self.hyper = torch.zeros(4, dtype=torch.float32, device=self.state.flat.device) if self.state.flat.is_cuda else None

    def zero_grad(self, set_to_none=False):
        self.state.zero_grad()

    def state_dict(self):
        """Flat moments + step count + lr (what resume needs; timm's resume_checkpoint(optimizer=...) restores the same things)."""
        return {'m': self.m.detach().cpu(), 'v': self.v.detach().cpu(), 'step_count': self.step_count,
                'param_groups': [dict(g) for g in self.param_groups], 'names': list(self.state.names),
                'offsets': list(self.state.offsets)}

    def load_state_dict(self, sd):
        assert list(sd['names']) == list(self.state.names) and list(sd['offsets']) == list(self.state.offsets), 'optimizer state of another model'
        self.m.copy_(sd['m'])
        self.v.copy_(sd['v'])
        self.step_count = int(sd['step_count'])
        self.param_groups[0].update(sd['param_groups'][0])

    def push_hyper(self, grad_scale: float = 1.0):
        """Advance the step count and upload {lr, 1-b1^t, 1-b2^t, grad_scale} for a graph-captured step(device_hyper=True)."""
        self.step_count += 1
        t = self.step_count
        b1, b2 = self.betas
        host = torch.empty(4, dtype=torch.float32, pin_memory=True)
        host[0], host[1], host[2], host[3] = self.param_groups[0]['lr'], 1 - b1 ** t, 1 - b2 ** t, grad_scale
        self.hyper.copy_(host, non_blocking=True)

    def step(self, grad_scale: float = 1.0, gathered: bool = False, device_hyper: bool = False):
        """gathered: the flat gradient already holds this step's gradients (GradBuckets did it bucket by bucket).
        device_hyper: read lr / bias corrections / grad_scale from self.hyper (see push_hyper) instead of scalar arguments."""
        if not gathered:
            self.state.gather()
        b1, b2 = self.betas
        lib = L.load()
        ema_d = self.ema_decay if self.ema_decay is not None else 0.0
        if self.clip_grad is not None and not device_hyper:       # clipping rescales hyper[3] on the device: take that path
            self.push_hyper(grad_scale)
            device_hyper = True
        if self.clip_grad is not None:
            L.check(lib.ga_grad_clip_scale(L.ptr(self.state.grad), L.ll(self.state.numel), L.f(self.clip_grad), L.ptr(self.hyper),
                                           L.ptr(self._clip_scratch), L.stream()), 'ga_grad_clip_scale')
        if device_hyper:
            L.check(lib.ga_adamw_ema_dev(L.ptr(self.state.flat), L.ptr(self.state.grad), L.ptr(self.m), L.ptr(self.v),
                                         L.ptr(self.ema_flat), L.ptr(self.state.flat16), L.ptr(self.flags), SEG_SHIFT, L.ll(self.state.numel),
                                         L.ptr(self.hyper), L.f(b1), L.f(b2), L.f(self.eps), L.f(self.wd), L.f(ema_d), L.stream()),
                    'ga_adamw_ema_dev')
        else:
            self.step_count += 1
            t = self.step_count
            lr = self.param_groups[0]['lr']
            L.check(lib.ga_adamw_ema(L.ptr(self.state.flat), L.ptr(self.state.grad), L.ptr(self.m), L.ptr(self.v),
                                     L.ptr(self.ema_flat), L.ptr(self.state.flat16), L.ptr(self.flags), SEG_SHIFT, L.ll(self.state.numel), L.f(lr),
                                     L.f(b1), L.f(b2), L.f(self.eps), L.f(self.wd), L.f(1 - b1 ** t), L.f(1 - b2 ** t), L.f(ema_d),
                                     L.f(grad_scale), L.stream()), 'ga_adamw_ema')
        if self.ema_model is not None:
            L.check(lib.ga_ema_lerp(L.ptr(self.ema_bufflat), L.ptr(self.state.bufflat), L.ll(self.ema_bufflat.numel()),
                                    L.f(self.ema_decay), L.stream()), 'ga_ema_lerp')


class FusedLambEma(FusedAdamWEma):
    """timm.optim.Lamb semantics (the optimizer of the published recipes, GA/README.md:26) + ModelEmaV2 over the flat state:
    global-norm clip, Adam moments, per-tensor trust ratio on the decayed tensors; three launches (ga_lamb_ema).
    Same interface as FusedAdamWEma (step / push_hyper / zero_grad / param_groups)."""

    def __init__(self, model: nn.Module, lr=5e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.05, ema_decay: Optional[float] = None,
                 filter_bias_and_bn=True, max_grad_norm: float = 1.0):
        super().__init__(model, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, ema_decay=ema_decay,
                         filter_bias_and_bn=filter_bias_and_bn)
        import numpy as np
        st = self.state
        rec = np.zeros((len(st.params), 4), dtype=np.int64)
        chunk = 0
        for r, (name, p_, o) in enumerate(zip(st.names, st.params, st.offsets)):
            decay = bool(weight_decay) and not (filter_bias_and_bn and (p_.ndim <= 1 or name.endswith('.bias')))
            rec[r] = (o, p_.numel(), chunk, int(decay))
            chunk += (p_.numel() + 4095) // 4096
        self.max_grad_norm = max_grad_norm
        self._chunks = chunk
        self._table = torch.from_numpy(rec.reshape(-1)).to(st.flat.device)
        self._scratch = torch.zeros(1 + 2 * len(st.params), dtype=torch.float32, device=st.flat.device)

    def step(self, grad_scale: float = 1.0, gathered: bool = False, device_hyper: bool = False):
        if not gathered:
            self.state.gather()
        if not device_hyper:
            self.push_hyper(grad_scale)
        b1, b2 = self.betas
        L.check(L.load().ga_lamb_ema(L.ptr(self.state.flat), L.ptr(self.state.grad), L.ptr(self.m), L.ptr(self.v), L.ptr(self.ema_flat),
                                     L.ptr(self._table), len(self.state.params), L.ll(self._chunks), L.ll(self.state.numel),
                                     L.ptr(self._scratch), L.ptr(self.hyper), L.f(b1), L.f(b2), L.f(self.eps), L.f(self.wd),
                                     L.f(self.max_grad_norm), L.f(self.ema_decay if self.ema_decay is not None else 0.0), L.stream()),
                'ga_lamb_ema')
        if self.state.flat16 is not None:          # bf16 shadows of the updated parameters (the AdamW kernel writes them itself)
            L.check(L.load().ga_cast_bf16(L.ptr(self.state.flat), L.ptr(self.state.flat16), L.ll(self.state.numel), L.stream()),
                    'ga_cast_bf16')
        if self.ema_model is not None:
            L.check(L.load().ga_ema_lerp(L.ptr(self.ema_bufflat), L.ptr(self.state.bufflat), L.ll(self.ema_bufflat.numel()),
                                         L.f(self.ema_decay), L.stream()), 'ga_ema_lerp')


class GradBuckets:
    """Overlapped data-parallel gradient all-reduce (mean) over contiguous slices of the flat gradient.

    Parameters are bucketed in reverse registration order (the order backward produces them); a post-accumulate
    hook counts a bucket down and launches its all-reduce on a side stream as soon as it is complete.
    """

    def __init__(self, state: FlatState, process_group=None, bucket_mb: float = 25.0, overlap: bool = True):
        self.state, self.pg, self.overlap = state, process_group, overlap
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        cap = int(bucket_mb * (1 << 20) / 4)
        self.buckets: List[dict] = []
        cur = None
        for idx in reversed(range(len(state.params))):
            o, n = state.offsets[idx], _round_up(state.params[idx].numel(), SEG)
            if cur is None or cur['hi'] - o > cap:
                cur = {'lo': o, 'hi': o + n, 'members': [], 'pending': 0, 'work': None}
                self.buckets.append(cur)
            cur['lo'] = o
            cur['members'].append(idx)
        self.bucket_of = {}
        for b in self.buckets:
            for idx in b['members']:
                self.bucket_of[idx] = b
        self.enabled = True
        self._cuda = state.flat.is_cuda
        self.side = torch.cuda.Stream() if (self._cuda and overlap) else None
        if self.world > 1:
            for idx, p in enumerate(state.params):
                p.register_post_accumulate_grad_hook(self._make_hook(idx))

    def _make_hook(self, idx):
        def hook(param):
            if not self.enabled:
                return
            b = self.bucket_of[idx]
            b['pending'] -= 1
            if b['pending'] == 0:
                self._launch(b)
        return hook

    def prepare(self):
        for b in self.buckets:
            b['pending'], b['work'] = len(b['members']), None

    def _launch(self, b):
        if self._cuda:
            self.state.gather(b['members'])        # adopt the bucket's gradients into its contiguous slice (one launch)
        else:
            for idx in b['members']:               # CPU (gloo tests): plain copies
                p, o = self.state.params[idx], self.state.offsets[idx]
                dst = self.state.grad[o:o + p.numel()]
                dst.zero_() if p.grad is None else dst.copy_(p.grad.reshape(-1))
        buf = self.state.grad[b['lo']:b['hi']]
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.side):
                b['work'] = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        else:
            b['work'] = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def finish(self) -> float:
        """Wait for every bucket; returns the 1/world factor to fold into the optimizer's grad_scale."""
        if self.world == 1:
            return 1.0
        for b in self.buckets:
            if b['work'] is None:           # parameter without gradient this step: reduce the (zero) slice anyway
                self._launch(b)
        for b in self.buckets:
            b['work'].wait()
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        return 1.0 / self.world
