"""Hot-path subset of the reference's source/utils.py."""
import torch


def unfold(tensor, mode):
    """Mode-`mode` unfolding, modes starting at 0 (reference source/utils.py:60-74): a view/copy
    made by torch; the solver's own unfoldings are produced once per layer by admmq_unfold3."""
    return torch.reshape(torch.moveaxis(tensor, mode, 0), (tensor.shape[mode], -1))


def bncalibrate_model(model, dataset_loader, num_samples=1000, device='cuda'):
    """Re-estimate the batch-norm statistics of a (factorized) model (reference source/utils.py:134-155): parameters
    frozen, BatchNorm / LayerNorm modules in train mode, forward passes over `dataset_loader` (an iterable of
    (images, labels) batches with a `batch_size` attribute, or of plain batches) until more than `num_samples`
    images have been seen, no gradients."""
    from torch import nn
    model.eval()
    for param in model.parameters():
        param.requires_grad = False
    for m in model.modules():
        if isinstance(m, (nn.BatchNorm2d, nn.LayerNorm)):
            m.train()
    count = 0
    for batch in dataset_loader:
        if count > num_samples:
            break
        x = batch[0] if isinstance(batch, (tuple, list)) else batch
        with torch.no_grad():
            model(x.to(device))
        count += getattr(dataset_loader, "batch_size", None) or x.shape[0]
    return model


class SyntheticImages:
    """Deterministic stand-in for the ImageNet loaders of the reference (there is no dataset offline): batches of
    standard-normal images; `labels_from` (a model) makes the labels the arg-max of that model's logits, so its own
    top-1 is 100 % by construction and a compressed model's top-1 is its agreement with it."""

    def __init__(self, n_batches, batch_size, image_size=64, seed=0, device="cpu", labels_from=None):
        self.n_batches, self.batch_size, self.image_size = n_batches, batch_size, image_size
        self.seed, self.device, self.teacher = seed, device, labels_from

    def __len__(self):
        return self.n_batches

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        for _ in range(self.n_batches):
            x = torch.randn(self.batch_size, 3, self.image_size, self.image_size, generator=g).to(self.device)
            if self.teacher is None:
                y = torch.zeros(self.batch_size, dtype=torch.long, device=self.device)
            else:
                with torch.no_grad():
                    y = self.teacher(x).argmax(dim=1)
            yield x, y


def top1_accuracy(model, loader, device="cuda"):
    """Top-1 accuracy in percent over a loader of (images, labels)."""
    model.eval()
    hit = total = 0
    with torch.no_grad():
        for x, y in loader:
            pred = model(x.to(device)).argmax(dim=1)
            hit += int((pred == y.to(device)).sum())
            total += int(y.numel())
    return 100.0 * hit / max(total, 1)
