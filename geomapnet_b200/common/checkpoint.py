"""Checkpoint interchange with the reference.

Host-side mirror of the three places the reference touches checkpoints:

* ``load_state_dict(model, state_dict)``   -- /root/reference/common/train.py:22-53: loads a
  state dict whose keys carry a different module prefix than the model's (``mapnet.`` when a
  PoseNet checkpoint goes into a MapNet, ``module.`` from nn.DataParallel, or the reverse);
  the prefix is inferred from the FIRST parameter name of each side, a mismatch raises KeyError.
* ``save_checkpoint(...)``                 -- common/train.py:198-204: the dict
  ``{epoch, model_state_dict, optim_state_dict, criterion_state_dict}`` under
  ``epoch_%03d.pth.tar``.
* ``load_checkpoint(...)``                 -- common/train.py:160-177 (Trainer resume) and
  scripts/eval.py:85-90: model weights always, optimizer / epoch / criterion scalars only when
  ``resume_optim``; criterion parameters absent from the checkpoint are taken as 0.0.

The product modules expose the reference's 222 state_dict keys, so the authors' released
``epoch_*.pth.tar`` files load unchanged.  One difference is bridged here: the product's
BatchNorm nodes are plain parameter containers, not ``nn.BatchNorm2d``, so the version shim
that lets torch load pre-0.4.1 checkpoints (no ``num_batches_tracked`` entries) does not run
for them -- ``load_state_dict`` supplies zeros for exactly those keys, which is what the shim does.
"""
import os
from collections import OrderedDict

import torch

__all__ = ["load_state_dict", "save_checkpoint", "load_checkpoint", "checkpoint_filename"]


def _prefix_fix(model_first, state_first):
    """(add, strip): what turns a state-dict key into the model's key.  The two FIRST parameter names must be equal
    up to a leading module prefix on one side (``mapnet.`` / ``module.``); anything else is a KeyError."""
    if model_first.endswith(state_first):
        return model_first[:len(model_first) - len(state_first)], ""
    if state_first.endswith(model_first):
        return "", state_first[:len(state_first) - len(model_first)]
    raise KeyError("Could not find the correct prefixes between %s and %s" % (model_first, state_first))


def load_state_dict(model, state_dict):
    """Prefix-tolerant load (behaviour of /root/reference/common/train.py:22-53).  Returns the re-keyed dict."""
    model_first = next((n for n, _ in model.named_parameters()), None)
    state_first = next(iter(state_dict.keys()), None)
    if model_first is None or state_first is None:
        raise KeyError("load_state_dict: empty model or state dict")
    add, strip = _prefix_fix(model_first, state_first)
    rekeyed = OrderedDict((add + (k[len(strip):] if strip and k.startswith(strip) else k), v)
                          for k, v in state_dict.items())
    # checkpoints written before BatchNorm tracked its batch count (torch < 0.4.1) carry no num_batches_tracked
    # entries; nn.BatchNorm2d fills them with zeros on load, the product's plain BN containers get the same here
    for k, v in model.state_dict().items():
        if k.endswith(".num_batches_tracked"):
            rekeyed.setdefault(k, torch.zeros_like(v))
    model.load_state_dict(rekeyed)
    return rekeyed


def checkpoint_filename(logdir, epoch):
    return os.path.join(logdir, "epoch_{:03d}.pth.tar".format(epoch))


def save_checkpoint(logdir, epoch, model, optimizer, criterion):
    """common/train.py:198-204.  `optimizer` is the reference-style wrapper (``.learner``) or a
    bare torch optimizer.  Returns the file name."""
    learner = optimizer.learner if hasattr(optimizer, "learner") else optimizer
    filename = checkpoint_filename(logdir, epoch)
    checkpoint_dict = {"epoch": epoch, "model_state_dict": model.state_dict(),
                       "optim_state_dict": learner.state_dict(),
                       "criterion_state_dict": criterion.state_dict()}
    torch.save(checkpoint_dict, filename)
    return filename


def load_checkpoint(checkpoint_file, model, optimizer=None, criterion=None, resume_optim=False, map_location=None):
    """common/train.py:160-177.  Returns the epoch to start from (0 unless `resume_optim`)."""
    if not os.path.isfile(checkpoint_file):
        raise IOError("checkpoint %s not found" % checkpoint_file)
    if map_location is None and not torch.cuda.is_available():
        map_location = lambda storage, loc: storage      # noqa: E731  (train.py:163)
    checkpoint = torch.load(checkpoint_file, map_location=map_location, weights_only=False)
    load_state_dict(model, checkpoint["model_state_dict"])
    start_epoch = 0
    if resume_optim:
        if optimizer is None:
            raise ValueError("resume_optim needs the optimizer")
        learner = optimizer.learner if hasattr(optimizer, "learner") else optimizer
        learner.load_state_dict(checkpoint["optim_state_dict"])
        start_epoch = checkpoint["epoch"]
        if "criterion_state_dict" in checkpoint and criterion is not None:
            c_state = checkpoint["criterion_state_dict"]
            append_dict = {k: torch.Tensor([0.0]) for k, _ in criterion.named_parameters() if k not in c_state}
            c_state.update(append_dict)
            criterion.load_state_dict(c_state)
    return start_epoch
