"""Hot-path subset of the reference's source/utils.py."""
import torch


def unfold(tensor, mode):
    """Mode-`mode` unfolding, modes starting at 0 (reference source/utils.py:60-74): a view/copy
    made by torch; the solver's own unfoldings are produced once per layer by admmq_unfold3."""
    return torch.reshape(torch.moveaxis(tensor, mode, 0), (tensor.shape[mode], -1))
