"""PeCLR pre-training entrypoint (mirror of src/experiments/peclr_training.py:27-96 of the reference):
same flags, same config merging, same model construction and hook sequence -- with the reference's dataset /
Comet plumbing replaced by the synthetic two-view generator and a JSONL logger, and the Lightning Trainer by
peclr_b200.lightning.Trainer.

    python src/experiments/peclr_training.py --rotate --crop -resnet_size 50 -epochs 1 -batch_size 128 \
        -accumulate_grad_batches 1 -save_top_k 1 -save_period 1 -num_workers 8

Multi-GPU: launch one process per GPU with torchrun; NT-Xent then sees the global batch (fused all-gather) and
gradients are summed across ranks.
"""
import json
import os
from pprint import pformat

import torch

from .easydict import EasyDict as edict
from .experiments_utils import (get_callbacks, get_general_args, get_model, prepare_name, save_experiment_key,
                                update_model_params, update_train_params)
from .lightning import JsonlLogger, Trainer, seed_everything

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config")
TRAINING_CONFIG_PATH = os.path.join(CONFIG_DIR, "training_config.json")
HYBRID2_CONFIG = os.path.join(CONFIG_DIR, "hybrid2_config.json")


def read_json(path):
    with open(path, "r") as f:
        return json.load(f)


def _init_distributed():
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world


def main(argv=None):
    experiment_type = "hybrid2"
    args = get_general_args("Hybrid model 2 training script.", argv)
    train_param = update_train_params(args, edict(read_json(TRAINING_CONFIG_PATH)))
    model_param = edict(read_json(HYBRID2_CONFIG))
    world = _init_distributed()
    rank = int(os.environ.get("RANK", "0"))
    if rank == 0:
        print(f"Train parameters {pformat(dict(train_param))}")
    seed_everything(train_param.seed)

    from .synthetic import SyntheticTwoViewDataset, get_train_val_split

    num_samples = args.num_samples or 64 * train_param.batch_size * world
    data = SyntheticTwoViewDataset(num_samples, args.image_size, seed=train_param.seed + rank,
                                   rotate=bool(train_param.augmentation_flags.get("rotate")),
                                   train_ratio=train_param.train_ratio)
    train_loader, val_loader = get_train_val_split(
        data, batch_size=train_param.batch_size, num_workers=train_param.num_workers, pin_memory=True,
        drop_last=True, persistent_workers=train_param.num_workers > 0)
    experiment_name = args.experiment_name or prepare_name(f"{experiment_type}_", train_param, hybrid_naming=False)

    model_param = update_model_params(model_param, args, len(data) * world, train_param)
    model_param.augmentation = [k for k, v in train_param.augmentation_flags.items() if v]
    if rank == 0:
        print(f"Model parameters {pformat(dict(model_param))}")
    model = get_model(experiment_type="hybrid2", heatmap_flag=args.heatmap, denoiser_flag=args.denoiser)(
        config=model_param)
    callbacks = get_callbacks(logging_interval=args.log_interval, experiment_type="hybrid2",
                              save_top_k=args.save_top_k, period=args.save_period)
    save_dir = os.environ.get("SAVED_META_INFO_PATH")
    logger = JsonlLogger(save_dir, experiment_name) if rank == 0 else None
    if args.meta_file is not None and rank == 0:
        save_experiment_key(experiment_name, logger.get_key(), args.meta_file)
    # checkpoints: $SAVED_MODELS_BASE_PATH/<experiment key>/checkpoints/epoch=N.ckpt when that variable is set (the
    # layout get_latest_checkpoint / restore_model read, src/models/utils.py:189-206), else <root dir>/checkpoints;
    # -experiment_key [-checkpoint epoch=N.ckpt] resumes from such a directory (weights, optimiser, schedule, epoch)
    models_base = os.environ.get("SAVED_MODELS_BASE_PATH")
    if models_base and logger is not None:
        callbacks["checkpoint_callback"].dirpath = os.path.join(models_base, logger.get_key(), "checkpoints")
    resume = None
    if args.experiment_key:
        from .model_utils import get_latest_checkpoint

        resume = get_latest_checkpoint(args.experiment_key, args.checkpoint)
    trainer = Trainer(accumulate_grad_batches=train_param.accumulate_grad_batches, gpus="0", logger=logger,
                      max_epochs=train_param.epochs, precision=train_param.precision, amp_backend="native",
                      limit_train_batches=args.limit_train_batches, resume_from_checkpoint=resume, **callbacks)
    trainer.fit(model, train_loader, val_loader if len(val_loader) > 0 else None)
    if rank == 0:
        loss = float(model.train_metrics_epoch.get("loss", float("nan")))
        print(json.dumps({"epoch_loss": loss, "images_per_sec": trainer.images_per_sec, "world": world}))
    return model, trainer


if __name__ == "__main__":
    main()
