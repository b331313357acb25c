"""Data loaders of the UVC loops (mirror of the reference's utils/data_utils.py:13-105: CIFAR / ImageNet folders through
torchvision with a DistributedSampler), plus `--dataset synthetic`: ImageNet-shaped random tensors generated once on the
device, for machines without a dataset (the benchmark box).  I/O is outside the accelerated hot path."""
import os

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, DistributedSampler, RandomSampler, SequentialSampler


class SyntheticLoader:
    """`steps` batches of N(0,1) images and uniform labels; the same two device tensors are reused (no host traffic)."""

    def __init__(self, batch_size, steps, img_size=224, num_classes=1000, device="cuda", seed=730):
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.x = torch.randn(batch_size, 3, img_size, img_size, generator=g).to(device)
        self.y = torch.randint(0, num_classes, (batch_size,), generator=g).to(device)
        self.steps = steps

    def __len__(self):
        return self.steps

    def __iter__(self):
        for _ in range(self.steps):
            yield self.x.clone(), self.y       # mixup works in place on x


def get_loader(args):
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if args.dataset == "synthetic":
        nc = getattr(args, "num_classes", 1000)
        steps = getattr(args, "synthetic_steps", 100)
        return (SyntheticLoader(args.train_batch_size, steps, args.img_size, nc, args.device, seed=args.seed),
                SyntheticLoader(args.eval_batch_size, max(1, steps // 10), args.img_size, nc, args.device, seed=args.seed + 1))
    from torchvision import datasets, transforms
    if args.local_rank not in [-1, 0] and world > 1:
        dist.barrier()       # rank 0 prepares the dataset first (reference :16)
    if args.dataset in ("cifar10", "cifar100"):
        norm = transforms.Normalize(mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5])
        ttrain = transforms.Compose([transforms.RandomResizedCrop((args.img_size, args.img_size), scale=(0.05, 1.0)), transforms.ToTensor(), norm])
        ttest = transforms.Compose([transforms.Resize((args.img_size, args.img_size)), transforms.ToTensor(), norm])
        DS = datasets.CIFAR10 if args.dataset == "cifar10" else datasets.CIFAR100
        trainset = DS(root="./data", train=True, download=True, transform=ttrain)
        testset = DS(root="./data", train=False, download=True, transform=ttest)
    else:
        norm = transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
        ttrain = transforms.Compose([transforms.RandomResizedCrop(args.img_size), transforms.RandomHorizontalFlip(), transforms.ToTensor(), norm])
        ttest = transforms.Compose([transforms.Resize(int(args.img_size * 256 / 224)), transforms.CenterCrop(args.img_size), transforms.ToTensor(), norm])
        trainset = datasets.ImageFolder(os.path.join(args.data_dir, 'train'), ttrain)
        testset = datasets.ImageFolder(os.path.join(args.data_dir, 'val'), ttest)
    if args.local_rank == 0 and world > 1:
        dist.barrier()
    train_sampler = RandomSampler(trainset) if args.local_rank == -1 or world == 1 else DistributedSampler(trainset)
    train_loader = DataLoader(trainset, sampler=train_sampler, batch_size=args.train_batch_size, num_workers=args.num_workers, pin_memory=True)
    test_loader = DataLoader(testset, sampler=SequentialSampler(testset), batch_size=args.eval_batch_size, num_workers=args.num_workers, pin_memory=True)
    return train_loader, test_loader
