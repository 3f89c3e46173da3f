"""Flag surface and config merging of the reference's pre-training entrypoint (mirror of the parts of
src/experiments/utils.py that peclr_training.py uses: get_general_args :29-163, update_train_params :276-314,
update_param :317-332, get_model :564-574, get_callbacks :587-605, update_model_params :608-615), plus the
flags this build adds for synthetic data and multi-GPU runs."""
import argparse
import os
from typing import List

import torch

from .easydict import EasyDict as edict
from .hybrid2_model import Hybrid2Model
from .lightning import LearningRateMonitor, ModelCheckpoint
from .simclr_model import SimCLR

AUGMENTATION_FLAGS = ["color_drop", "color_jitter", "crop", "cut_out", "flip", "gaussian_blur", "random_crop",
                      "resize", "rotate", "sobel_filter", "gaussian_noise"]


def get_general_args(description: str = "Script for training baseline supervised model", argv=None) -> argparse.Namespace:
    parser = argparse.ArgumentParser(description=description)
    for flag, text in (("color_drop", "random color drop"), ("color_jitter", "random jitter"), ("crop", "cropping"),
                       ("cut_out", "random cut out"), ("flip", "random flipping"), ("gaussian_blur", "gaussian blur"),
                       ("rotate", "random rotation"), ("random_crop", "random cropping"), ("resize", "resizing"),
                       ("sobel_filter", "sobel filtering"), ("gaussian_noise", "gaussian noise")):
        parser.add_argument("--" + flag, action="store_true", help="To enable " + text)
    parser.add_argument("-tag", action="append", help="Tag for the run", default=[])
    parser.add_argument("-batch_size", type=int, help="Batch size")
    parser.add_argument("-epochs", type=int, help="Number of epochs")
    parser.add_argument("-seed", type=int, help="To add seed")
    parser.add_argument("-num_workers", type=int, help="Number of workers for Dataloader.")
    parser.add_argument("-train_ratio", type=float, help="Ratio of train:validation split.")
    parser.add_argument("-accumulate_grad_batches", type=int, help="Number of batches to accumulate gradient.")
    parser.add_argument("-lr", type=float, help="learning rate", default=None)
    parser.add_argument("-optimizer", type=str, help="Select optimizer", default=None, choices=["LARS", "adam"])
    parser.add_argument("--denoiser", action="store_true", default=False)
    parser.add_argument("--heatmap", action="store_true", default=False)
    parser.add_argument("-sources", action="append", default=[], choices=["freihand", "interhand", "mpii", "youtube"])
    parser.add_argument("-log_interval", type=str, default="epoch", choices=["step", "epoch"])
    parser.add_argument("-experiment_key", type=str, default=None)
    parser.add_argument("-checkpoint", type=str, default="")
    parser.add_argument("-meta_file", type=str, default=None)
    parser.add_argument("-experiment_name", type=str, default="")
    parser.add_argument("-save_period", type=int, default=1)
    parser.add_argument("-save_top_k", type=int, default=3)
    parser.add_argument("--encoder_trainable", action="store_true", default=False)
    parser.add_argument("-resnet_size", type=str, default="18", choices=["18", "34", "50", "101", "152"])
    parser.add_argument("-lr_max_epochs", type=int, default=None)
    parser.add_argument("--use_palm", action="store_true", default=False)
    # additions of this build (the reference's data loaders are bypassed)
    parser.add_argument("--synthetic", action="store_true", default=True,
                        help="Synthetic two-view data (the only data source of this build).")
    parser.add_argument("-image_size", type=int, default=224, help="Synthetic image side.")
    parser.add_argument("-num_samples", type=int, default=None, help="Synthetic samples per epoch.")
    parser.add_argument("-limit_train_batches", type=int, default=None)
    return parser.parse_args(argv)


def update_param(args: argparse.Namespace, config: edict, params: List[str]) -> edict:
    given = vars(args)
    for name in params:
        if given.get(name) is not None:
            config[name] = given[name]
    return config


def update_train_params(args: argparse.Namespace, train_param: edict) -> edict:
    if args.train_ratio is not None:
        train_param.train_ratio = (args.train_ratio * 100 % 100) / 100.0
    train_param.update(update_param(args, train_param,
                                    ["batch_size", "epochs", "train_ratio", "num_workers", "seed", "use_palm"]))
    train_param.augmentation_flags = update_param(args, train_param.augmentation_flags, AUGMENTATION_FLAGS)
    if args.accumulate_grad_batches is not None:
        train_param.accumulate_grad_batches = args.accumulate_grad_batches
    return train_param


def update_model_params(model_param: edict, args, data_length: int, train_param: edict) -> edict:
    model_param = update_param(args, model_param, ["optimizer", "lr", "resnet_size", "lr_max_epochs"])
    model_param.num_samples = data_length
    model_param.batch_size = train_param.batch_size
    model_param.num_of_mini_batch = train_param.accumulate_grad_batches
    return model_param


_NAME_CODES = {"color_drop": "CD", "color_jitter": "CJ", "crop": "C", "cut_out": "CO", "flip": "F",
               "gaussian_blur": "GB", "random_crop": "RC", "resize": "Re", "rotate": "Ro", "sobel_filter": "SF",
               "gaussian_noise": "GN"}


def prepare_name(prefix: str, train_param: edict, hybrid_naming: bool = False) -> str:
    """src/experiments/utils.py:335-393: experiment name = prefix + batch size + sorted augmentation codes."""
    def codes(flags):
        return "_".join(sorted(_NAME_CODES[k] for k, v in flags.items() if v))

    if hybrid_naming:
        return (f"{prefix}{train_param.batch_size}_rel_{codes(train_param.pairwise.augmentation_flags)}"
                f"_con_{codes(train_param.contrastive.augmentation_flags)}")
    return f"{prefix}{train_param.batch_size}{codes(train_param.augmentation_flags)}"


def save_experiment_key(experiment_name: str, experiment_key: str, filename="default.csv"):
    """src/experiments/utils.py:396-409: appends "<name>,<key>" to $SAVED_META_INFO_PATH/<filename>."""
    with open(os.path.join(os.environ.get("SAVED_META_INFO_PATH", "."), filename), "a") as f:
        f.write(f"{experiment_name},{experiment_key}\n")


def restore_model(model, experiment_key: str, checkpoint: str = ""):
    """src/experiments/utils.py:535-546: loads the newest (or the named) checkpoint's state_dict into `model`."""
    from .model_utils import get_latest_checkpoint

    path = get_latest_checkpoint(experiment_key, checkpoint)
    print(f"Restoring {path}")
    model.load_state_dict(torch.load(path, map_location="cpu")["state_dict"])
    return model


def get_checkpoints(experiment_key: str, number: int = 3) -> List[str]:
    """src/experiments/utils.py:549-561: the last `number` checkpoint file names (reverse lexicographic order)."""
    base = os.environ.get("SAVED_MODELS_BASE_PATH", "")
    return sorted(os.listdir(os.path.join(base, experiment_key, "checkpoints")))[::-1][:number]


def get_model(experiment_type: str, heatmap_flag: bool, denoiser_flag: bool):
    if heatmap_flag:
        raise NotImplementedError("heat-map models are not part of the pre-training path")
    if experiment_type == "simclr":
        return SimCLR
    if experiment_type == "hybrid2":
        return Hybrid2Model
    raise NotImplementedError(experiment_type)


def get_callbacks(logging_interval: str, experiment_type: str, save_top_k: int = 1, period: int = 1,
                  monitor: str = "checkpoint_saving_loss"):
    return {
        "callbacks": [LearningRateMonitor(logging_interval=logging_interval)],
        "checkpoint_callback": ModelCheckpoint(save_top_k=save_top_k, period=period, monitor=monitor),
    }
