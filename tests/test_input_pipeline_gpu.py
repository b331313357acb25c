"""GPU parity of the on-device input pipeline (uvc_mixup: timm batch-mode Mixup / CutMix + smoothed mixed targets in one launch) against the
plain-torch formulas of timm's public semantics (timm is not in the reference tree: parity unpinned by the reference, pinned here), and of the
Mixup class on CUDA against the same class on the host with the same numpy seed (same lambda / box draws)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref_targets(y, nc, lam, smoothing):
    off = smoothing / nc
    on = 1. - smoothing + off
    oh = lambda t: torch.full((t.numel(), nc), off).scatter_(1, t.view(-1, 1), on)
    return oh(y) * lam + oh(y.flip(0)) * (1. - lam)


@pytest.mark.parametrize("B,C,H,W", [(8, 3, 224, 224), (2, 3, 32, 32), (6, 1, 20, 36)])
def test_mixup_kernel_matches_torch_formula(B, C, H, W):
    from uvc_b200 import ops
    g = torch.Generator().manual_seed(B)
    x0 = torch.randn(B, C, H, W, generator=g); y = torch.randint(0, 1000, (B,), generator=g)
    lam = 0.37
    x = x0.cuda().clone()
    t = ops.mixup_(x, y.cuda(), 1000, lam, 0.1)
    want = x0 * lam + x0.flip(0) * (1. - lam)
    assert torch.equal(x.cpu(), x0.clone().mul_(lam).add_(x0.flip(0).mul_(1. - lam))) or (x.cpu() - want).abs().max() < 1e-6
    torch.testing.assert_close(t.cpu(), ref_targets(y, 1000, lam, 0.1), rtol=1e-6, atol=1e-7)
    # cutmix: only the box is exchanged with the flipped batch
    box = (H // 4, H // 4 + H // 2, 3, W - 5)
    x = x0.cuda().clone()
    t = ops.mixup_(x, y.cuda(), 1000, 0.6, 0.0, box=box)
    want = x0.clone(); yl, yh, xl, xh = box
    want[:, :, yl:yh, xl:xh] = x0.flip(0)[:, :, yl:yh, xl:xh]
    assert torch.equal(x.cpu(), want)
    torch.testing.assert_close(t.cpu(), ref_targets(y, 1000, 0.6, 0.0), rtol=1e-6, atol=1e-7)
    # lam == 1: images untouched, targets = smoothed one-hot
    x = x0.cuda().clone()
    t = ops.mixup_(x, y.cuda(), 1000, 1.0, 0.1)
    assert torch.equal(x.cpu(), x0)
    torch.testing.assert_close(t.cpu(), ref_targets(y, 1000, 1.0, 0.1), rtol=1e-6, atol=1e-7)


def test_mixup_class_on_cuda_equals_host_path_with_the_same_seed():
    from uvc_b200.utils.mixup import Mixup
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(10, 3, 64, 64, generator=g); y = torch.randint(0, 100, (10,), generator=g)
    for trial in range(6):          # covers mixup, cutmix and the "no mix" draw
        mix = Mixup(mixup_alpha=0.8, cutmix_alpha=1.0, prob=0.8, switch_prob=0.5, label_smoothing=0.1, num_classes=100)
        np.random.seed(100 + trial)
        xh, th = mix(x0.clone(), y.clone())
        np.random.seed(100 + trial)
        xd, td = mix(x0.cuda().clone(), y.cuda())
        assert (xd.cpu() - xh).abs().max() < 1e-6 and (td.cpu() - th).abs().max() < 1e-6, trial
