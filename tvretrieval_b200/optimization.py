"""`BertAdam` whose `step()` is ONE fused multi-tensor CUDA update (xmlb_bert_adam_step, csrc/train.cu).

Mirror of the optimizer surface that reference baselines/crossmodal_moment_localization/train.py:150-164,84 uses from
baselines/crossmodal_moment_localization/optimization.py:219-338: constructor arguments, parameter groups with
per-group `weight_decay` / `lr`, `get_lr()`, `zero_grad()`, `step()`, and the per-parameter state keys `step`,
`next_m`, `next_v` (so `state_dict()` round-trips with the reference's).  Semantics (optimization.py:296-330):
per-PARAMETER gradient-norm clipping to `max_grad_norm` (rescales the stored gradient, like clip_grad_norm_), Adam
moments without bias correction, decoupled weight decay, lr multiplied by the warm-up schedule at the parameter's
own step count.  Parameters and gradients must be CUDA fp32: there is no CPU path.
"""
import math

import numpy as np
import torch
from torch.optim import Optimizer

from . import _lib, ops

CHUNK = 8192  # elements of one tensor per CTA


class _LRSchedule:
    """Learning-rate multiplier as a function of training progress = step / t_total (optimization.py:31-76)."""
    warn_t_total = False

    def __init__(self, warmup=0.002, t_total=-1, **kw):
        if not 0.0 <= warmup < 1.0 and not warmup == -1:
            raise ValueError("Invalid warmup: {} - should be in [0.0, 1.0[ or -1".format(warmup))
        self.warmup, self.t_total = float(max(warmup, 0.0)), float(t_total)

    def get_lr(self, step, nowarn=False):
        if self.t_total < 0:
            return 1.0
        return self.get_lr_(float(step) / self.t_total)

    def get_lr_(self, progress):
        return 1.0


class ConstantLR(_LRSchedule):
    pass


class WarmupConstantSchedule(_LRSchedule):
    def get_lr_(self, progress):
        return progress / self.warmup if progress < self.warmup else 1.0


class WarmupLinearSchedule(_LRSchedule):
    warn_t_total = True

    def get_lr_(self, progress):
        if progress < self.warmup:
            return progress / self.warmup
        return max((progress - 1.0) / (self.warmup - 1.0), 0.0)


class WarmupCosineSchedule(_LRSchedule):
    warn_t_total = True

    def __init__(self, warmup=0.002, t_total=-1, cycles=0.5, **kw):
        super().__init__(warmup=warmup, t_total=t_total, **kw)
        self.cycles = cycles

    def get_lr_(self, progress):
        if progress < self.warmup:
            return progress / self.warmup
        progress = (progress - self.warmup) / (1 - self.warmup)
        return 0.5 * (1.0 + math.cos(math.pi * self.cycles * 2 * progress))


SCHEDULES = {None: ConstantLR, "none": ConstantLR, "warmup_cosine": WarmupCosineSchedule,
             "warmup_constant": WarmupConstantSchedule, "warmup_linear": WarmupLinearSchedule}


class BertAdam(Optimizer):
    def __init__(self, params, lr, warmup=-1, t_total=-1, schedule="warmup_linear", b1=0.9, b2=0.999, e=1e-6,
                 weight_decay=0.01, max_grad_norm=1.0, **kwargs):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not isinstance(schedule, _LRSchedule) and schedule not in SCHEDULES:
            raise ValueError("Invalid schedule parameter: {}".format(schedule))
        if not 0.0 <= b1 < 1.0:
            raise ValueError("Invalid b1 parameter: {} - should be in [0.0, 1.0[".format(b1))
        if not 0.0 <= b2 < 1.0:
            raise ValueError("Invalid b2 parameter: {} - should be in [0.0, 1.0[".format(b2))
        if not e >= 0.0:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(e))
        if not isinstance(schedule, _LRSchedule):
            schedule = SCHEDULES[schedule](warmup=warmup, t_total=t_total)
        super().__init__(params, dict(lr=lr, schedule=schedule, b1=b1, b2=b2, e=e, weight_decay=weight_decay,
                                      max_grad_norm=max_grad_norm))

    def get_lr(self):
        lr = []
        for group in self.param_groups:
            for p in group["params"]:
                state = self.state[p]
                if len(state) == 0:
                    return [0]
                lr.append(group["lr"] * group["schedule"].get_lr(state["step"]))
        return lr

    def zero_grad(self, set_to_none=False):
        """Gradients are zeroed in place by default: their buffers are reused by the next backward pass."""
        super().zero_grad(set_to_none=set_to_none)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:  # one launch pair per hyper-parameter set (normally: one)
            self._step_group(group)
        ops.invalidate_weight_caches()  # the kernels wrote the parameters behind torch's version counters
        return loss

    def _step_group(self, group):
        live, lrs = [], []
        for p in group["params"]:
            if p.grad is None:
                continue
            if p.grad.is_sparse:
                raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise _lib.XmlbError("BertAdam: parameters must be contiguous CUDA float32 tensors (no CPU path)")
            if not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
            state = self.state[p]
            if len(state) == 0:
                state["step"] = 0
                state["next_m"] = torch.zeros_like(p)
                state["next_v"] = torch.zeros_like(p)
            live.append(p)
            lrs.append(group["lr"] * group["schedule"].get_lr(state["step"]))
            state["step"] += 1
        if not live:
            return
        dev = live[0].device
        # The chunk table only depends on the buffer addresses, which stay the same from step to step when
        # zero_grad() keeps the gradient buffers: build and upload it once, then refresh only the per-tensor
        # (lr, weight decay) rows -- a few hundred bytes -- every step.
        key = tuple((p.data_ptr(), p.grad.data_ptr(), p.numel()) for p in live)
        tables = self.__dict__.setdefault("_xmlb_tables", {})  # not in param_groups: state_dict() stays reference-like
        cache = tables.get(id(group))
        if cache is None or cache["key"] != key:
            chunks, spans = [], []
            for t, p in enumerate(live):
                st = self.state[p]
                ptrs = (p.data_ptr(), p.grad.data_ptr(), st["next_m"].data_ptr(), st["next_v"].data_ptr())
                n, first = p.numel(), len(chunks)
                for off in range(0, n, CHUNK):
                    chunks.append((ptrs[0] + 4 * off, ptrs[1] + 4 * off, ptrs[2] + 4 * off, ptrs[3] + 4 * off,
                                   min(CHUNK, n - off), t))
                spans.append((first, len(chunks) - first))
            tens_np = np.zeros((len(live), 4), dtype=np.int32)
            tens_np[:, :2] = np.asarray(spans, dtype=np.int32)
            cache = dict(key=key, n_chunks=len(chunks), tens_np=tens_np,
                         chunk_dev=torch.from_numpy(np.asarray(chunks, dtype=np.int64)).to(dev),
                         tens_host=torch.empty((len(live), 4), dtype=torch.int32).pin_memory(),
                         tens_dev=torch.empty((len(live), 4), dtype=torch.int32, device=dev),
                         partial=torch.empty(len(chunks), device=dev, dtype=torch.float32), copied=None)
            tables[id(group)] = cache
        tens_np = cache["tens_np"]
        tens_np[:, 2] = np.asarray(lrs, dtype=np.float32).view(np.int32)
        tens_np[:, 3] = np.float32(group["weight_decay"]).view(np.int32)
        if cache["copied"] is not None:
            cache["copied"].synchronize()  # the previous step's upload has left the pinned staging buffer
        cache["tens_host"].numpy()[:] = tens_np
        cache["tens_dev"].copy_(cache["tens_host"], non_blocking=True)
        cache["copied"] = torch.cuda.Event()
        cache["copied"].record(torch.cuda.current_stream(dev))
        rc = _lib.lib().xmlb_bert_adam_step(cache["chunk_dev"].data_ptr(), cache["n_chunks"],
                                            cache["tens_dev"].data_ptr(), len(live), cache["partial"].data_ptr(),
                                            group["b1"], group["b2"], group["e"], group["max_grad_norm"],
                                            torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "xmlb_bert_adam_step")
