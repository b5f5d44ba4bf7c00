"""`isaacgym.gymutil` shim: argument / sim-config parsing used by the reference's
utils/config.py (get_args :234-330, parse_sim_params :198-231)."""
import argparse

from . import gymapi


def parse_device_str(device_str):
    device, device_id = 'cpu', 0
    if device_str == 'cpu' or device_str == 'cuda':
        device = device_str
    else:
        parts = device_str.split(':')
        assert len(parts) == 2 and parts[0] == 'cuda', f'Invalid device string "{device_str}"'
        device, device_id = parts[0], int(parts[1])
    return device, device_id


def parse_arguments(description="Isaac Gym Example", headless=False, no_graphics=False, custom_parameters=[]):
    parser = argparse.ArgumentParser(description=description)
    if headless:
        parser.add_argument('--headless', action='store_true', help='Run headless')
    if no_graphics:
        parser.add_argument('--nographics', action='store_true')
    parser.add_argument('--sim_device', type=str, default="cuda:0")
    parser.add_argument('--pipeline', type=str, default="gpu")
    parser.add_argument('--graphics_device_id', type=int, default=0)
    physics = parser.add_mutually_exclusive_group()
    physics.add_argument('--flex', action='store_true')
    physics.add_argument('--physx', action='store_true')
    parser.add_argument('--num_threads', type=int, default=0)
    parser.add_argument('--subscenes', type=int, default=0)
    parser.add_argument('--slices', type=int)
    for argument in custom_parameters:
        if ("name" in argument) and ("type" in argument or "action" in argument):
            help_str = argument.get("help", "")
            if "type" in argument:
                if "default" in argument:
                    parser.add_argument(argument["name"], type=argument["type"], default=argument["default"], help=help_str)
                else:
                    parser.add_argument(argument["name"], type=argument["type"], help=help_str)
            elif "action" in argument:
                parser.add_argument(argument["name"], action=argument["action"], help=help_str)
    args = parser.parse_args()
    args.sim_device_type, args.compute_device_id = parse_device_str(args.sim_device)
    pipeline = args.pipeline.lower()
    assert pipeline in ('cpu', 'gpu', 'cuda'), f"Invalid pipeline '{args.pipeline}'"
    args.use_gpu_pipeline = pipeline in ('gpu', 'cuda')
    if args.sim_device_type != 'cuda' and args.flex:
        args.sim_device, args.sim_device_type, args.compute_device_id = 'cuda:0', 'cuda', 0
    if args.sim_device_type != 'cuda' and pipeline == 'gpu':
        args.pipeline, args.use_gpu_pipeline = 'CPU', False
    args.physics_engine = gymapi.SIM_PHYSX
    args.use_gpu = (args.sim_device_type == 'cuda')
    if args.flex:
        args.physics_engine = gymapi.SIM_FLEX
    if no_graphics and args.nographics:
        args.headless = True
    if args.slices is None:
        args.slices = args.subscenes
    return args


def parse_bool(v):
    return bool(v)


def parse_vec3(v):
    return gymapi.Vec3(*v)


def _set_attrs(obj, cfg):
    for k, v in cfg.items():
        if not hasattr(obj, k):
            continue
        if isinstance(getattr(obj, k), gymapi.Vec3):
            setattr(obj, k, gymapi.Vec3(*v))
        else:
            setattr(obj, k, v)


def parse_sim_config(sim_cfg, sim_options):
    opts = ["dt", "substeps", "gravity", "use_gpu_pipeline", "num_client_threads", "up_axis"]
    for k in opts:
        if k in sim_cfg:
            v = sim_cfg[k]
            if k == "gravity":
                v = gymapi.Vec3(*v)
            setattr(sim_options, k, v)
    if "physx" in sim_cfg:
        _set_attrs(sim_options.physx, sim_cfg["physx"])
    if "flex" in sim_cfg:
        _set_attrs(sim_options.flex, sim_cfg["flex"])


def get_property_setter_map(gym):
    return {}


def get_property_getter_map(gym):
    return {}


def get_default_setter_args(gym):
    return {}


def apply_random_samples(*a, **k):
    raise RuntimeError("isaacgym shim: domain randomisation is out of scope (SURVEY.md §2 row 1)")


def check_buckets(*a, **k):
    raise RuntimeError("isaacgym shim: domain randomisation is out of scope")


def generate_random_samples(*a, **k):
    raise RuntimeError("isaacgym shim: domain randomisation is out of scope")
