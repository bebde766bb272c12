#!/usr/bin/env python
"""Entry point with the reference's CLI (`main.py:10-68`):

    python main.py --config configs/stanford.ini --log ./log [--override "k=v,k2=v2"]

Parses the .ini (flattened), applies overrides, saves the effective config to <log>/config.ini, opens a
TensorBoard SummaryWriter on <log> and runs the dataset driver on the CUDA path."""
import argparse
import os

from piccolo_b200 import localize
from piccolo_b200.parse_utils import apply_override, parse_ini, parse_override, save_effective_config


def main():
    cli = argparse.ArgumentParser(description="PICCOLO sampling-loss localisation on the B200 CUDA path")
    cli.add_argument("--config", type=str, default=None, help=".ini file (configs/stanford.ini, configs/stanford_parallel.ini, configs/omniscenes.ini)")
    cli.add_argument("--log", type=str, default="./log", help="output directory: config.ini, *_results.csv, TensorBoard events, results/*.png")
    cli.add_argument("--override", default=None, help='config overrides, "key=value,key2=value2"')
    args = cli.parse_args()
    cfg = parse_ini(args.config)
    os.makedirs(args.log, exist_ok=True)
    from torch.utils.tensorboard import SummaryWriter
    writer = SummaryWriter(args.log)
    if args.override is not None:
        cfg = apply_override(cfg, parse_override(args.override))
    save_effective_config(cfg, os.path.join(args.log, "config.ini"))
    if cfg.dataset == "Stanford2D-3D-S":
        localize.localize_stanford(cfg, writer, args.log)
    elif cfg.dataset == "OmniScenes":
        localize.localize_omniscenes(cfg, writer, args.log)
    else:
        raise ValueError
    writer.close()


if __name__ == "__main__":
    main()
