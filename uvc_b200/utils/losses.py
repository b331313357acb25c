"""Knowledge-distillation loss of the UVC loops, fused on the device.

Mirror of the reference's `utils/losses.py` (`DistillationLoss(base_criterion, teacher_model, distillation_type, alpha,
tau)(inputs, outputs, labels)`, :10-65).  With the base criteria the UVC scripts use (timm SoftTargetCrossEntropy /
LabelSmoothingCrossEntropy / nn.CrossEntropyLoss) the base loss, the soft KL term and d(loss)/d(logits) come out of ONE
kernel launch (`uvc_distill_loss`), and that gradient is what the classifier-head backward GEMM consumes; the teacher
forward runs through the same engine in inference mode (no activations kept).  Any other base criterion is evaluated
with its own torch code and only the distillation term uses the kernel.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .mixup import LabelSmoothingCrossEntropy, SoftTargetCrossEntropy, one_hot


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, teacher_logits, targets, alpha, T):
        out, dl = ops.distill_loss(logits.contiguous(), teacher_logits, targets.contiguous(), alpha, T, 1.0, want_grad=True)
        ctx.save_for_backward(dl)
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, g, _g_parts):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None, None, None


class DistillationLoss(nn.Module):
    def __init__(self, base_criterion, teacher_model, distillation_type, alpha, tau):
        super().__init__()
        assert distillation_type in ['none', 'soft', 'hard']
        self.base_criterion = base_criterion
        self.teacher_model = teacher_model
        if teacher_model is not None and hasattr(teacher_model, "_engine_forward"):
            teacher_model.weights_frozen = True      # never trained (utils/losses.py:47-49 runs it under no_grad): the engine keeps its converted weights
        self.distillation_type = distillation_type
        self.alpha = alpha
        self.tau = tau
        self.last_parts = None      # device tensor (loss, base, kd) of the last fused call
        self._side = None           # side stream of prefetch_teacher()
        self._pref = None

    def prefetch_teacher(self, inputs):
        """Optional: start the teacher forward of `inputs` NOW, on a side stream, so that it runs next to the student forward the caller issues
        after this call (the two are independent until the loss; utils/losses.py:47-49 runs the teacher inside the criterion, after the student).
        Every hot kernel is persistent with a static share of the tiles, so a kernel's last wave leaves SMs idle (8.03 tile waves run as 9, 5.2
        attention waves as 6); CTAs of the other stream's kernels move onto SMs as they are vacated.  `forward` picks the result up when it is called
        with the same `inputs` tensor, and falls back to computing it itself otherwise.  UVC_TEACHER_STREAM=0 disables it."""
        import os
        self._pref = None
        if self.distillation_type == 'none' or self.teacher_model is None or not inputs.is_cuda or os.environ.get("UVC_TEACHER_STREAM", "1") == "0":
            return
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=inputs.device)
        self._side.wait_stream(cur)                                  # the (mixed) inputs are ready
        with torch.cuda.stream(self._side), torch.no_grad():
            out, _ = self.teacher_model(inputs)
        self._pref = (inputs.data_ptr(), inputs._version, out)

    def _teacher(self, inputs):
        pref, self._pref = self._pref, None
        if pref is not None and pref[0] == inputs.data_ptr() and pref[1] == inputs._version:
            cur = torch.cuda.current_stream()
            cur.wait_stream(self._side)
            pref[2].record_stream(cur)                               # allocated on the side stream, consumed here
            return pref[2]
        with torch.no_grad():
            out, _ = self.teacher_model(inputs)
        return out

    def _targets(self, outputs, labels):
        """soft-target matrix equivalent to the base criterion, or None if the criterion is not one of the fusable ones"""
        nc = outputs.shape[1]
        bc = self.base_criterion
        if isinstance(bc, SoftTargetCrossEntropy) or type(bc).__name__ == "SoftTargetCrossEntropy":
            return labels
        if isinstance(bc, LabelSmoothingCrossEntropy) or type(bc).__name__ == "LabelSmoothingCrossEntropy":
            off = bc.smoothing / nc
            return one_hot(labels, nc, bc.confidence + off, off)
        if isinstance(bc, nn.CrossEntropyLoss) and bc.weight is None and bc.reduction == "mean" and labels.dtype == torch.long:
            ls = getattr(bc, "label_smoothing", 0.0)
            return one_hot(labels, nc, 1.0 - ls + ls / nc, ls / nc)
        return None

    def forward(self, inputs, outputs, labels):
        outputs_kd = None
        if not isinstance(outputs, torch.Tensor):
            outputs, outputs_kd = outputs
        if self.distillation_type != 'none' and outputs_kd is None:
            raise ValueError("When knowledge distillation is enabled, the model is expected to return a Tuple[Tensor, Tensor] with the "
                             "output of the class_token and the dist_token")
        targets = self._targets(outputs, labels)
        fused = targets is not None and (self.distillation_type == 'none' or outputs_kd is outputs)
        teacher_outputs = None
        if self.distillation_type != 'none':
            teacher_outputs = self._teacher(inputs)
        if fused:
            T = float(self.tau)
            if self.distillation_type == 'soft':
                loss, parts = _FusedLoss.apply(outputs, teacher_outputs.contiguous(), targets, float(self.alpha), T)
            elif self.distillation_type == 'hard':
                # (1-a) * CE(y) + a * CE(onehot(argmax teacher)) is one soft-target cross entropy
                mixed = targets * (1 - self.alpha) + one_hot(teacher_outputs.argmax(dim=1), outputs.shape[1]) * self.alpha
                loss, parts = _FusedLoss.apply(outputs, None, mixed, 0.0, 1.0)
            else:
                loss, parts = _FusedLoss.apply(outputs, None, targets, 0.0, 1.0)
            self.last_parts = parts
            return loss
        # generic composition (unknown base criterion or a separate distillation head)
        base_loss = self.base_criterion(outputs, labels)
        if self.distillation_type == 'none':
            return base_loss
        if self.distillation_type == 'soft':
            T = self.tau
            kd = F.kl_div(F.log_softmax(outputs_kd / T, dim=1), F.log_softmax(teacher_outputs / T, dim=1), reduction='sum',
                          log_target=True) * (T * T) / outputs_kd.numel()
        else:
            kd = F.cross_entropy(outputs_kd, teacher_outputs.argmax(dim=1))
        return base_loss * (1 - self.alpha) + kd * self.alpha
