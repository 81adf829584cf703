"""Registry and factory with the reference's names (`mqe/envs/utils.py:38-134`, `mqe/utils/helpers.py:81-261`)."""
from __future__ import annotations

import argparse
import os
import random

import numpy as np

from . import configs as C
from .go1 import Go1, Go1FootballDefender, Go1Object, Go1Sheep
from .wrappers import (EmptyWrapper, Go1FootballDefenderWrapper, Go1FootballGameWrapper, Go1GateWrapper, Go1PushboxWrapper, Go1RotationWrapper,
                       Go1SeesawWrapper, Go1SheepWrapper, Go1WrestlingWrapper, Go1BridgeWrapper, Go1TugWrapper)

ENV_DICT = {
    "go1plane": {"class": Go1, "config": C.Go1PlaneCfg, "wrapper": EmptyWrapper},
    "go1gate": {"class": Go1, "config": C.Go1GateCfg, "wrapper": Go1GateWrapper},
    "go1sheep-easy": {"class": Go1Sheep, "config": C.SingleSheepCfg, "wrapper": Go1SheepWrapper},
    "go1sheep-hard": {"class": Go1Sheep, "config": C.NineSheepCfg, "wrapper": Go1SheepWrapper},
    "go1football-defender": {"class": Go1FootballDefender, "config": C.Go1FootballDefenderCfg, "wrapper": Go1FootballDefenderWrapper},
    "go1football-1vs1": {"class": Go1Object, "config": C.Go1Football1vs1Cfg, "wrapper": Go1FootballGameWrapper},
    "go1football-2vs2": {"class": Go1Object, "config": C.Go1Football2vs2Cfg, "wrapper": Go1FootballGameWrapper},
    "go1seesaw": {"class": Go1Object, "config": C.Go1SeesawCfg, "wrapper": Go1SeesawWrapper},
    "go1pushbox": {"class": Go1Object, "config": C.Go1PushboxCfg, "wrapper": Go1PushboxWrapper},
    "go1revolvingdoor": {"class": Go1Object, "config": C.Go1RotationCfg, "wrapper": Go1RotationWrapper},
    "go1tug": {"class": Go1Object, "config": C.Go1TugCfg, "wrapper": Go1TugWrapper},
    "go1wrestling": {"class": Go1Object, "config": C.Go1WrestlingCfg, "wrapper": Go1WrestlingWrapper},
    "go1bridge": {"class": Go1Object, "config": C.Go1BridgeCfg, "wrapper": Go1BridgeWrapper},
}
# every task of the reference registry (mqe/envs/utils.py:38-109) is built; kept for callers that probe it
NOT_YET = ()


def set_seed(seed):
    """helpers.py:81-91"""
    import torch
    if seed == -1:
        seed = np.random.randint(0, 10000)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    return seed


def get_args(argv=None):
    """The argument surface of helpers.py:168-194 + gymutil.parse_arguments that make_env consumes."""
    p = argparse.ArgumentParser(description="RL Policy")
    p.add_argument("--task", type=str, default="go1gate")
    p.add_argument("--headless", action="store_true", default=False)
    p.add_argument("--horovod", action="store_true", default=False)
    p.add_argument("--rl_device", type=str, default="cuda:0")
    p.add_argument("--sim_device", type=str, default="cuda:0")
    p.add_argument("--num_envs", type=int, default=None)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--max_iterations", type=int, default=None)
    p.add_argument("--record_video", action="store_true", default=False)
    p.add_argument("--num_threads", type=int, default=0)
    p.add_argument("--subscenes", type=int, default=0)
    p.add_argument("--physics_engine", type=str, default="mqe_b200")
    p.add_argument("--use_gpu", action="store_true", default=True)
    p.add_argument("--use_gpu_pipeline", action="store_true", default=True)
    args, _ = p.parse_known_args(argv)
    return args


def make_env(task_class, env_cfg, args=None, **engine_kw):
    """helpers.py:245-261"""
    if args is None:
        args = get_args()
    if getattr(args, "num_envs", None) is not None:
        env_cfg.env.num_envs = args.num_envs
    seed = set_seed(getattr(args, "seed", 0))
    sim_params = {"sim": C.class_to_dict(env_cfg.sim)}
    env = task_class(cfg=env_cfg, sim_params=sim_params, physics_engine=getattr(args, "physics_engine", "mqe_b200"),
                     sim_device=getattr(args, "sim_device", "cuda:0"), headless=getattr(args, "headless", True),
                     seed=seed, **engine_kw)
    return env, env_cfg


def make_mqe_env(env_name: str, args=None, custom_cfg=None, **engine_kw):
    """utils.py:111-121 -> (wrapped env, env_cfg)"""
    if env_name not in ENV_DICT:
        if env_name in NOT_YET:
            raise NotImplementedError(f"task '{env_name}' is in the reference registry but outside this build's hot-path scope (SURVEY.md 8(f))")
        raise KeyError(env_name)
    env_dict = ENV_DICT[env_name]
    cfg = env_dict["config"]()
    if callable(custom_cfg):
        cfg = custom_cfg(cfg)
    env, env_cfg = make_env(env_dict["class"], cfg, args, **engine_kw)
    env = env_dict["wrapper"](env)
    return env, env_cfg


def custom_cfg(args):
    """utils.py:123-134"""
    def fn(cfg):
        if getattr(args, "num_envs", None) is not None:
            cfg.env.num_envs = args.num_envs
        cfg.env.record_video = getattr(args, "record_video", False)
        return cfg
    return fn
