"""isaacgym.gymutil: argument parsing helpers (restated from the public Preview-4 package; openrl_ws/utils.py:157-228 carries
the same logic)."""
import argparse

from . import gymapi


def parse_device_str(device_str):
    device, device_id = "cpu", 0
    if device_str in ("cpu", "cuda"):
        device = device_str
    else:
        parts = device_str.split(":")
        assert len(parts) == 2 and parts[0] == "cuda", f'Invalid device string "{device_str}"'
        device, device_id = parts[0], int(parts[1])
    return device, device_id


def parse_arguments(description="Isaac Gym Example", headless=False, no_graphics=False, custom_parameters=()):
    parser = argparse.ArgumentParser(description=description)
    if headless:
        parser.add_argument("--headless", action="store_true")
    if no_graphics:
        parser.add_argument("--nographics", action="store_true")
    parser.add_argument("--sim_device", type=str, default="cuda:0")
    parser.add_argument("--pipeline", type=str, default="gpu")
    parser.add_argument("--graphics_device_id", type=int, default=0)
    group = parser.add_mutually_exclusive_group()
    group.add_argument("--flex", action="store_true")
    group.add_argument("--physx", action="store_true")
    parser.add_argument("--num_threads", type=int, default=0)
    parser.add_argument("--subscenes", type=int, default=0)
    parser.add_argument("--slices", type=int)
    for a in custom_parameters:
        if "name" in a and ("type" in a or "action" in a):
            kw = {"help": a.get("help", "")}
            if "type" in a:
                kw["type"] = a["type"]
                if "default" in a:
                    kw["default"] = a["default"]
            else:
                kw["action"] = a["action"]
            parser.add_argument(a["name"], **kw)
    args, _ = parser.parse_known_args()
    args.sim_device_type, args.compute_device_id = parse_device_str(args.sim_device)
    pipeline = args.pipeline.lower()
    assert pipeline in ("cpu", "gpu", "cuda")
    args.use_gpu_pipeline = pipeline in ("gpu", "cuda")
    args.physics_engine = gymapi.SIM_FLEX if args.flex else gymapi.SIM_PHYSX
    args.use_gpu = args.sim_device_type == "cuda"
    if no_graphics and args.nographics:
        args.headless = True
    if args.slices is None:
        args.slices = args.subscenes
    return args
