"""`python -m nerf_tiny_b200.main --conf=lego` — the reference's main.py call surface (main.py:10-56) with tolerant
config handling: the shipped .ini files define EPOCH but not TOTAL_ITER / RESULTS_PATH /
CONTINUE, and LR_MILESTONE is parsed as a real list."""
from __future__ import annotations

import argparse
import ast
from configparser import ConfigParser

CONF_DIR = "./conf/"


def read_conf(path, section):
    conf = ConfigParser()
    if not conf.read(path):
        raise FileNotFoundError(path)
    get = lambda key, default=None: conf.get(section, key, fallback=default)
    total_iter = get("TOTAL_ITER", get("EPOCH", "100000"))
    return dict(
        gpu=int(get("GPU", "0")), img_dir=get("IMG_DIR"), results_path=get("RESULTS_PATH", "./results/"),
        ckpt_path=get("CKPT_PATH", "./checkpoint/"), low_res=int(get("LOW_RES", "1")), total_iter=int(total_iter),
        batch_ray=int(get("BATCH_RAY", "400")), learning=float(get("LEARNING", "1e-3")), lr_gamma=float(get("LR_GAMMA", "0.1")),
        lr_milestone=list(ast.literal_eval(get("LR_MILESTONE", "[10, 200]"))), n_coarse=int(get("N_COARSE", "64")),
        n_fine=int(get("N_FINE", "128")), data_type=get("DATA_TYPE", "sync"), step=int(get("STEP", "100")),
        decay_end=float(get("DECAY_END", "200000")), sched=get("SCHED", "EXP"),
        continue_=bool(ast.literal_eval(get("CONTINUE", "False"))))


def main(argv=None):
    ap = argparse.ArgumentParser(description="NeRF argument parser.")
    ap.add_argument("--conf", type=str, default="lego")
    ap.add_argument("--conf-dir", type=str, default=CONF_DIR)
    args = ap.parse_args(argv)
    kw = read_conf(args.conf_dir + args.conf + ".ini", args.conf)
    from .nerf import NeRFRunner
    runner = NeRFRunner(**kw)
    runner.trainer("train")        # the reference calls trainer() without its required `mode` (main.py:55)
    runner.display()


if __name__ == "__main__":
    main()
