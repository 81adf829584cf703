"""The C-ABI library loads and exports every symbol include/mqe_b200.h declares (no compute calls: CPU box)."""
import ctypes
import os
import re

from mqe_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mqe_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mqe_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    import __graft_entry__ as g
    if not os.path.exists(E.LIB_PATH):
        g.build()
    lib = E.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in mqe_b200.h but not exported"
    assert sorted(E.EXPORTED_SYMBOLS) == syms
    assert lib.mqe_abi_version() == E.ABI_VERSION


def test_struct_mirrors_match_header_sizes():
    """ctypes mirrors vs the C compiler's view of the structs (compiled with gcc from the header)."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "mqe_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(MqeRobotModel), sizeof(MqeWeights), sizeof(MqeSimDesc), sizeof(MqeStepResultLayout), sizeof(MqeWrapperDesc));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    from mqe_b200.model import RobotModelC
    assert [int(x) for x in out] == [ctypes.sizeof(RobotModelC), ctypes.sizeof(E.WeightsC), ctypes.sizeof(E.SimDescC),
                                     ctypes.sizeof(E.StepResultLayoutC), ctypes.sizeof(E.WrapperDescC)]


def test_no_cpu_fallback():
    """Without a GPU the engine refuses to construct (MQE_ERR_NO_DEVICE); nothing routes through the oracle."""
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mqe_b200 import scene as S
    from mqe_b200.envs import configs as C
    cfg = C.Go1GateCfg(); cfg.env.num_envs = 2
    np.random.seed(0)
    sc = S.build_scene(cfg)
    with pytest.raises(E.EngineError) as ei:
        E.Engine(sc.desc)
    assert "no CPU fallback" in str(ei.value) or "-3" in str(ei.value)
    src = "".join(open(os.path.join(ROOT, "multiagent-quadruped-environment_b200", f)).read()
                  for f in ("engine.py", "scene.py", os.path.join("envs", "go1.py"), os.path.join("envs", "wrappers.py")))
    assert "import oracle" not in src and "from oracle" not in src


def test_substep_launch_plan_residency():
    """Host-side launch plan of k_substeps (physics.cu `substeps_plan`): shared memory per CTA must fit the 227 KB opt-in limit for every
    task shape, and C2 (4096 envs x 2 robots) must come out as 7 resident warps of 4 envs -- the one-wave design point of DESIGN.md 3.1."""
    import __graft_entry__ as g
    if not os.path.exists(E.LIB_PATH):
        g.build()
    lib = E.load_library()
    fn = lib.mqe_substeps_smem_bytes
    fn.restype = ctypes.c_size_t
    fn.argtypes = [ctypes.c_int] * 5
    LIMIT = 227 * 1024
    shapes = {  # task: (A, NPCs that own a lane, pair budget)
        "go1gate": (2, 0, 8), "go1sheep-hard": (2, 9, 16), "go1sheep-easy": (2, 1, 16), "go1seesaw": (2, 1, 16),
        "go1football-defender": (3, 1, 16), "go1football-2vs2": (4, 1, 16), "go1plane": (1, 0, 8), "go1wrestling": (2, 0, 16),
    }
    for task, (A, Pd, maxpair) in shapes.items():
        lanes = 4 * A + Pd
        Eenv = 32 // lanes
        for N in (1, 7, 4096, 32768):
            b = fn(N, A, Pd, Eenv, maxpair)
            assert 0 < b <= LIMIT, (task, N, b)
    assert fn(4096, 2, 0, 4, 8) == 230944          # 8.3 KB header + 7 warps x 4 envs x 7.9 KB
    assert fn(4, 2, 0, 4, 8) < 48 * 1024           # a single warp of envs needs no opt-in at all
