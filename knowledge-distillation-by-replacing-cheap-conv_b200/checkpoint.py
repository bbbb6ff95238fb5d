"""Checkpoint layout of the reference, kept byte-compatible.

    state = {'arch', 'epoch', 'state_dict', 'optimizer', 'monitor_best', 'config'}      base/base_trainer.py:170-177
    resume = replay the surgery plan for epochs 1..epoch, then forgiving_state_restore  trainer/layerwise_trainer.py:404-427

`forgiving_state_restore` follows models/__init__.py:61-88: tensors whose name or size does not match are skipped, and
a state dict saved from an nn.DataParallel wrapper (every key starts with 'module.') loads into the bare module.  The
reference wraps the net in DataParallel to do that; here the prefix is simply stripped (no wrapper, works under DDP /
one process per GPU).  `save_checkpoint` writes from rank 0 only and puts a barrier behind it.
"""
import logging
import os

import torch
import torch.distributed as dist


def forgiving_state_restore(net, loaded_dict):
    """Partial load: keep every tensor of `loaded_dict` whose key and size match `net`; returns `net`."""
    keys = list(loaded_dict.keys())
    is_parallel = len(keys) > 0 and all(k.startswith('module.') for k in keys)
    if is_parallel:
        loaded_dict = {k[len('module.'):]: v for k, v in loaded_dict.items()}
    net_state_dict = net.state_dict()
    new_loaded_dict = {}
    for k in net_state_dict:
        if k in loaded_dict and net_state_dict[k].size() == loaded_dict[k].size():
            new_loaded_dict[k] = loaded_dict[k]
        else:
            logging.info("Skipped loading parameter %s", k)
    net_state_dict.update(new_loaded_dict)
    net.load_state_dict(net_state_dict)
    return net


def restore_snapshot(net, optimizer, snapshot, restore_optimizer_bool=False):
    """models/__init__.py:46-59: load `snapshot` (a path or an already loaded dict) into net (and optimizer)."""
    checkpoint = torch.load(snapshot, map_location=torch.device('cpu'), weights_only=False) if isinstance(snapshot, (str, os.PathLike)) else snapshot
    if optimizer is not None and 'optimizer' in checkpoint and restore_optimizer_bool:
        optimizer.load_state_dict(checkpoint['optimizer'])
    forgiving_state_restore(net, checkpoint['state_dict'] if 'state_dict' in checkpoint else checkpoint)
    return net, optimizer


def checkpoint_state(model, optimizer, epoch, monitor_best, config):
    """The dict base/base_trainer.py:170-177 pickles (same keys; `config` is whatever object the trainer carries)."""
    return {'arch': type(model).__name__, 'epoch': epoch, 'state_dict': model.state_dict(),
            'optimizer': optimizer.state_dict() if optimizer is not None else None,
            'monitor_best': monitor_best, 'config': config}


def save_checkpoint(path, model, optimizer, epoch, monitor_best=None, config=None, save_best=False):
    """Rank 0 writes `path` (and model_best.pth next to it when save_best); every rank waits behind a barrier."""
    distributed = dist.is_available() and dist.is_initialized()
    if not distributed or dist.get_rank() == 0:
        state = checkpoint_state(model, optimizer, epoch, monitor_best, config)
        torch.save(state, path)
        if save_best:
            torch.save(state, os.path.join(os.path.dirname(os.path.abspath(path)), 'model_best.pth'))
    if distributed:
        dist.barrier()
    return path


def resume(model, optimizer, path, make_optimizer, optimizer_args=None, optimizer_type=None, step=None, pruning=None):
    """trainer/layerwise_trainer.py:404-427 for a kdcc.DepthwiseStudent.

    The saved config's own "pruning" section is replayed through `kdcc.trainer.prepare_train_epoch` for every epoch
    1..saved epoch -- exactly what the reference does with `self.prepare_train_epoch(i, checkpoint['config'])`: its
    replace / hint / unfreeze lists are honoured as written, epoch 1 builds a fresh optimizer over the trainable student
    parameters (`make_optimizer(params)`), later epochs add one param group per unfrozen layer.  The tensors are then
    restored forgivingly and the optimizer state is loaded unless the optimizer type changed (`optimizer_type`, the
    current config's optimizer "type", is compared with the checkpoint's; None skips the comparison) -- the only case
    in which the reference drops it (:420-425).  A state that does not fit the rebuilt param groups raises, as in the
    reference.  `pruning` overrides the section taken from the checkpoint (a checkpoint whose `config` was not saved);
    a LayerwiseStep passed as `step` receives the optimizer and rebuilds its gradient bucket.
    Returns (saved epoch, optimizer, monitor_best)."""
    from .trainer import prepare_train_epoch
    checkpoint = torch.load(path, map_location=torch.device('cpu'), weights_only=False)
    epoch = int(checkpoint['epoch'])
    config = checkpoint.get('config')
    if pruning is None:
        if config is None:
            raise KeyError("checkpoint %s carries no config; pass pruning=" % path)
        pruning = config['pruning']
    for i in range(1, epoch + 1):
        optimizer = prepare_train_epoch(model, pruning, i, optimizer, make_optimizer, optimizer_args, step=step)
    forgiving_state_restore(model, checkpoint['state_dict'])
    saved_type = None
    try:
        saved_type = config['optimizer']['type']
    except (TypeError, KeyError):
        pass
    if optimizer is not None and checkpoint.get('optimizer') is not None:
        if optimizer_type is not None and saved_type is not None and saved_type != optimizer_type:
            logging.warning("optimizer type in %s (%s) differs from the configured one (%s); optimizer state not resumed",
                            path, saved_type, optimizer_type)
        else:
            optimizer.load_state_dict(checkpoint['optimizer'])
    if step is not None:
        step.optimizer = optimizer
        step.rebuild_bucket()
    return epoch, optimizer, checkpoint.get('monitor_best')
