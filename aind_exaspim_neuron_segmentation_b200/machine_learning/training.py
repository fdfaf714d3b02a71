"""Training-mode forward and backward of the U-Net on the native engine.

The reference trains with plain torch autograd (machine_learning/train.py:123-157, 200-223:
``model.train(); hat_y = model(x); loss = criterion(hat_y, y); scaler.scale(loss).backward()``).
The drop-in keeps that calling convention: in training mode ``UNet3D.forward`` returns logits
that carry a ``grad_fn``; its backward hands ``dLoss/dlogits`` to the native trainer
(C ABI ``exa_train_forward`` / ``exa_train_backward``) and gives every parameter its gradient,
so the reference's ``Trainer`` (optimizer, GradScaler, scheduler) runs unchanged on top.

What runs natively: the convolutions and their data gradients on the tcgen05 kernels of the
inference path, BatchNorm3d with batch statistics (running statistics are updated in place like
``nn.BatchNorm3d`` does), the backward of BatchNorm/LeakyReLU/MaxPool3d/Upsample/head and the
weight gradients (a tcgen05 kernel, csrc/train_wgrad.cu; ``EXA_WGRAD=mma`` selects the earlier
warp-level MMA kernel).  Nothing here falls back to torch operators.
"""

import ctypes
import weakref

import torch

from .. import _native


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class TrainEngine:
    """Owns an ``exa_trainer`` bound to the parameter storage of one module."""

    def __init__(self, module, precision="bf16"):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        device = next(module.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("exaspim_b200 trains on CUDA devices only (no CPU fallback); got "
                               f"device '{device}'")
        self.device = device
        self.precision = precision
        self._lib = _native.lib()
        handle = ctypes.c_void_p()
        code = self._lib.exa_train_create(
            device.index if device.index is not None else torch.cuda.current_device(),
            _native.PRECISION_BF16 if precision == "bf16" else _native.PRECISION_FP32,
            ctypes.byref(handle))
        if code < 0:
            msg = self._lib.exa_train_last_error(None)
            raise RuntimeError(f"exaspim_b200 exa_train_create failed ({code}): "
                               f"{msg.decode() if msg else 'unknown error'}")
        self._h = handle
        self._fingerprint = None
        self._module = weakref.ref(module)
        self._bn_counters = []
        self.rebind(module)

    def _check(self, code, what):
        if code < 0:
            msg = self._lib.exa_train_last_error(self._h)
            raise RuntimeError(f"exaspim_b200 {what} failed ({code}): "
                               f"{msg.decode() if msg else 'unknown error'}")
        return code

    @staticmethod
    def fingerprint(module):
        return tuple((k, v.data_ptr(), tuple(v.shape)) for k, v in module.state_dict().items())

    def rebind(self, module):
        """Bind the device storage of every float32 state_dict entry (parameters are read on every
        forward; BatchNorm running statistics are updated in place)."""
        fp = self.fingerprint(module)
        if fp == self._fingerprint:
            return
        self._bn_counters = []
        for name, t in module.state_dict(keep_vars=True).items():
            if t.dtype == torch.int64:
                self._bn_counters.append(t)  # num_batches_tracked: incremented here, in Python
                continue
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                raise RuntimeError(f"state_dict entry {name} must be a contiguous float32 tensor "
                                   f"on {self.device}")
            shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
            self._check(self._lib.exa_train_bind(self._h, name.encode(),
                                                 ctypes.c_void_p(t.data_ptr()), shape, t.dim()),
                        f"exa_train_bind({name})")
        n = ctypes.c_int64()
        self._check(self._lib.exa_train_grad_elems(self._h, ctypes.byref(n)), "exa_train_grad_elems")
        self.grad_elems = n.value
        self.out_channels = self._check(self._lib.exa_train_out_channels(self._h),
                                        "exa_train_out_channels")
        # parameters in module.named_parameters() order with their gradient slots
        self._slots = []
        for name, p in module.named_parameters():
            off, numel = ctypes.c_int64(), ctypes.c_int64()
            self._check(self._lib.exa_train_grad_slot(self._h, name.encode(), ctypes.byref(off),
                                                      ctypes.byref(numel)),
                        f"exa_train_grad_slot({name})")
            if numel.value != p.numel():
                raise RuntimeError(f"gradient slot of {name} has {numel.value} elements, the "
                                   f"parameter {p.numel()}")
            self._slots.append((name, off.value, numel.value, tuple(p.shape)))
        self._fingerprint = fp

    def forward(self, x):
        """(B,1,D,H,W) float32 cuda -> logits (B,C,D,H,W) float32; batch statistics, running
        statistics updated (unet3d.py:77-105 in train() mode)."""
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError("expected input of shape (B, 1, D, H, W)")
        x = x.to(torch.float32).contiguous()
        b, _, d, h, w = x.shape
        logits = torch.empty((b, self.out_channels, d, h, w), dtype=torch.float32, device=x.device)
        patch = (ctypes.c_int32 * 3)(d, h, w)
        self._check(self._lib.exa_train_forward(self._h, ctypes.c_void_p(x.data_ptr()), b, patch,
                                                ctypes.c_void_p(logits.data_ptr()),
                                                _stream_ptr(x.device)), "exa_train_forward")
        for counter in self._bn_counters:
            counter += 1
        return x, logits

    def backward(self, x, grad_logits):
        """Gradients of every parameter for the most recent forward: a list of tensors in
        ``named_parameters()`` order (views of one flat buffer)."""
        g = grad_logits.to(torch.float32).contiguous()
        flat = torch.empty(self.grad_elems, dtype=torch.float32, device=x.device)
        self._check(self._lib.exa_train_backward(self._h, ctypes.c_void_p(x.data_ptr()),
                                                 ctypes.c_void_p(g.data_ptr()),
                                                 ctypes.c_void_p(flat.data_ptr()),
                                                 _stream_ptr(x.device)), "exa_train_backward")
        return [flat[off:off + numel].view(shape) for _, off, numel, shape in self._slots]

    PROFILE_CATEGORIES = ("pack", "fprop", "bn_fwd", "misc_fwd", "bn_bwd", "wgrad", "dgrad",
                          "head_bwd", "wgrad_stem", "wgrad_reduce", "upsample_bwd", "pool_bwd")

    def profile_begin(self):
        self._check(self._lib.exa_train_profile_begin(self._h), "exa_train_profile_begin")

    def profile_end(self):
        """{category: (device ms, launches)} since profile_begin (synchronises)."""
        n = len(self.PROFILE_CATEGORIES)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int64 * n)()
        self._check(self._lib.exa_train_profile_end(self._h, ms, cnt, n), "exa_train_profile_end")
        return {name: (ms[i], cnt[i]) for i, name in enumerate(self.PROFILE_CATEGORIES)}

    @property
    def launch_count(self):
        return int(self._lib.exa_train_launch_count(self._h))

    @property
    def workspace_bytes(self):
        return int(self._lib.exa_train_workspace_bytes(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.exa_train_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# {module: {precision: TrainEngine}} outside the module's __dict__ (handles cannot be pickled)
_TRAINERS = weakref.WeakKeyDictionary()


def trainer_for_module(module, precision):
    per_module = _TRAINERS.setdefault(module, {})
    eng = per_module.get(precision)
    if eng is None:
        eng = per_module[precision] = TrainEngine(module, precision)
    else:
        eng.rebind(module)  # no-op unless the parameter storage moved (load_state_dict keeps it)
    return eng


def release(module):
    """Free the native trainers of ``module`` (activation workspace: about 1.1 GiB per 96^3 patch of
    the batch).  They are rebuilt on the next ``train()``-mode forward."""
    for eng in _TRAINERS.pop(module, {}).values():
        eng.close()


class _TrainStep(torch.autograd.Function):
    """logits = UNet3D(x) in training mode; backward returns the parameter gradients."""

    @staticmethod
    def forward(ctx, x, engine, *params):
        x, logits = engine.forward(x)
        ctx.engine = engine
        ctx.generation = engine.generation = getattr(engine, "generation", 0) + 1
        ctx.save_for_backward(x)
        return logits

    @staticmethod
    def backward(ctx, grad_logits):
        engine = ctx.engine
        if ctx.generation != engine.generation:
            raise RuntimeError("backward of a stale forward: the native trainer keeps the "
                               "activations of its most recent forward only")
        (x,) = ctx.saved_tensors
        grads = engine.backward(x, grad_logits)
        return (None, None, *grads)


def train_forward(module, x, precision):
    """Training-mode ``module(x)`` with a grad_fn (the input itself gets no gradient)."""
    engine = trainer_for_module(module, precision)
    params = [p for _, p in module.named_parameters()]
    return _TrainStep.apply(x, engine, *params)


def bce_with_logits(logits, target, grad_scale=1.0, with_grad=True):
    """``nn.BCEWithLogitsLoss()(logits, target)`` (train.py:76,222) on the native kernel.

    Returns ``(loss, grad)``: the mean loss as a 0-d float32 tensor and, when ``with_grad``,
    ``grad_scale * dLoss/dlogits``."""
    lib = _native.lib()
    logits = logits.to(torch.float32).contiguous()
    target = target.to(torch.float32).contiguous()
    if logits.shape != target.shape or not logits.is_cuda:
        raise ValueError("logits and target must be CUDA tensors of the same shape")
    n = logits.numel()
    loss_sum = torch.zeros(1, dtype=torch.float64, device=logits.device)
    grad = torch.empty_like(logits) if with_grad else None
    code = lib.exa_bce_with_logits(
        ctypes.c_void_p(logits.data_ptr()), ctypes.c_void_p(target.data_ptr()), n,
        float(grad_scale), ctypes.c_void_p(loss_sum.data_ptr()),
        ctypes.c_void_p(grad.data_ptr() if with_grad else 0), _stream_ptr(logits.device))
    if code < 0:
        msg = lib.exa_train_last_error(None)
        raise RuntimeError(f"exa_bce_with_logits failed ({code}): {msg.decode() if msg else ''}")
    return (loss_sum / n).to(torch.float32).reshape(()), grad
