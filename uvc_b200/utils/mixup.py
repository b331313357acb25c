"""Batch mixup / cutmix and the soft-target criteria the UVC loops take from timm (un-vendored, unpinned dependency of the
reference: joint_train.py:924-944, post_train.py:618-632).  Re-stated from timm's public semantics (timm.data.Mixup in
'batch' mode, timm.loss.SoftTargetCrossEntropy / LabelSmoothingCrossEntropy); numpy's global RNG drives the draws exactly
as timm does, so `np.random.seed(args.seed)` reproduces the same lambda / box sequence.

On CUDA batches the mixing of the images AND the smoothed mixed targets are ONE kernel launch (`uvc_mixup`, input_pipeline.cu) instead of ~10
elementwise torch launches; lambda and the CutMix box are still drawn here on the host, from numpy's RNG, as timm does.  The plain-torch
arithmetic below is what runs for host (CPU) tensors -- a loader that mixes before the copy, and the CPU tests of these semantics."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def one_hot(x, num_classes, on_value=1., off_value=0.):
    x = x.long().view(-1, 1)
    return torch.full((x.size(0), num_classes), off_value, device=x.device, dtype=torch.float32).scatter_(1, x, on_value)


def mixup_target(target, num_classes, lam=1., smoothing=0.0):
    off = smoothing / num_classes
    on = 1. - smoothing + off
    y1 = one_hot(target, num_classes, on, off)
    y2 = one_hot(target.flip(0), num_classes, on, off)
    return y1 * lam + y2 * (1. - lam)


def rand_bbox(img_shape, lam, margin=0.):
    ratio = np.sqrt(1 - lam)
    img_h, img_w = img_shape[-2:]
    cut_h, cut_w = int(img_h * ratio), int(img_w * ratio)
    margin_y, margin_x = int(margin * cut_h), int(margin * cut_w)
    cy = np.random.randint(0 + margin_y, img_h - margin_y)
    cx = np.random.randint(0 + margin_x, img_w - margin_x)
    yl = np.clip(cy - cut_h // 2, 0, img_h); yh = np.clip(cy + cut_h // 2, 0, img_h)
    xl = np.clip(cx - cut_w // 2, 0, img_w); xh = np.clip(cx + cut_w // 2, 0, img_w)
    return int(yl), int(yh), int(xl), int(xh)


def cutmix_bbox_and_lam(img_shape, lam, correct_lam=True):
    yl, yu, xl, xu = rand_bbox(img_shape, lam)
    if correct_lam:
        lam = 1. - (yu - yl) * (xu - xl) / float(img_shape[-2] * img_shape[-1])
    return (yl, yu, xl, xu), lam


class Mixup:
    def __init__(self, mixup_alpha=1., cutmix_alpha=0., cutmix_minmax=None, prob=1.0, switch_prob=0.5, mode='batch', correct_lam=True,
                 label_smoothing=0.1, num_classes=1000):
        assert cutmix_minmax is None and mode == 'batch', "UVC uses batch-mode mixup without cutmix_minmax"
        self.mixup_alpha, self.cutmix_alpha = mixup_alpha, cutmix_alpha
        self.mix_prob, self.switch_prob = prob, switch_prob
        self.label_smoothing, self.num_classes = label_smoothing, num_classes
        self.correct_lam = correct_lam
        self.mixup_enabled = True

    def _params_per_batch(self):
        lam, use_cutmix = 1., False
        if self.mixup_enabled and np.random.rand() < self.mix_prob:
            if self.mixup_alpha > 0. and self.cutmix_alpha > 0.:
                use_cutmix = np.random.rand() < self.switch_prob
                lam_mix = np.random.beta(self.cutmix_alpha, self.cutmix_alpha) if use_cutmix else np.random.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.mixup_alpha > 0.:
                lam_mix = np.random.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.cutmix_alpha > 0.:
                use_cutmix = True
                lam_mix = np.random.beta(self.cutmix_alpha, self.cutmix_alpha)
            else:
                raise ValueError("one of mixup_alpha > 0., cutmix_alpha > 0. must be set")
            lam = float(lam_mix)
        return lam, use_cutmix

    def _mix_batch(self, x):
        lam, use_cutmix = self._params_per_batch()
        if lam == 1.:
            return 1.
        if use_cutmix:
            (yl, yh, xl, xh), lam = cutmix_bbox_and_lam(x.shape, lam, correct_lam=self.correct_lam)
            x[:, :, yl:yh, xl:xh] = x.flip(0)[:, :, yl:yh, xl:xh]
        else:
            x_flipped = x.flip(0).mul_(1. - lam)
            x.mul_(lam).add_(x_flipped)
        return lam

    def __call__(self, x, target):
        assert len(x) % 2 == 0, 'Batch size should be even when using this'
        if x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous() and (x[0].numel() % 4 == 0):
            from .. import ops
            lam, use_cutmix = self._params_per_batch()       # same draws, in the same order, as the host path below
            box = None
            if lam != 1. and use_cutmix:
                box, lam = cutmix_bbox_and_lam(x.shape, lam, correct_lam=self.correct_lam)
            return x, ops.mixup_(x, target.long().contiguous(), self.num_classes, lam, self.label_smoothing, box)
        lam = self._mix_batch(x)
        return x, mixup_target(target, self.num_classes, lam, self.label_smoothing)


class SoftTargetCrossEntropy(nn.Module):
    """mean_b sum_c -target * log_softmax(x)"""

    def forward(self, x, target):
        return torch.sum(-target * F.log_softmax(x, dim=-1), dim=-1).mean()


class LabelSmoothingCrossEntropy(nn.Module):
    def __init__(self, smoothing=0.1):
        super().__init__()
        assert smoothing < 1.0
        self.smoothing, self.confidence = smoothing, 1. - smoothing

    def forward(self, x, target):
        logprobs = F.log_softmax(x, dim=-1)
        nll = -logprobs.gather(dim=-1, index=target.unsqueeze(1)).squeeze(1)
        smooth = -logprobs.mean(dim=-1)
        return (self.confidence * nll + self.smoothing * smooth).mean()
