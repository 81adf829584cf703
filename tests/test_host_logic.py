"""Host-side logic of the Python mirror that needs neither a GPU nor the oracle."""
import numpy as np


def test_adapter_infos_and_dones_helpers():
    """mqe_openrl_wrapper hands out the same list of empty info dicts every step and the same two dones arrays in turn (openrl_ws/utils.py:61-65
    allocates both per step): the dicts must come back empty whatever a caller did to them, dones must be [N, agent_num] bool copies."""
    from mqe_b200.openrl_adapter import mqe_openrl_wrapper
    w = mqe_openrl_wrapper.__new__(mqe_openrl_wrapper)
    w.agent_num = 3
    a = w._empty_infos(64)
    assert len(a) == 64 and all(isinstance(d, dict) and not d for d in a) and len({id(d) for d in a}) == 64
    a[7]["x"] = 1
    a[9].update(y=2)
    a[11].setdefault("z", 3)
    a[13] |= {"k": 1}
    b = w._empty_infos(64)
    assert b is a and not any(b)
    b[3]["t"] = 0
    b[3].pop("t")
    assert not any(w._empty_infos(64)) and len(w._empty_infos(32)) == 32
    done = np.zeros(64, dtype=bool)
    done[5] = True
    d1 = w._dones(done)
    d2 = w._dones(~done)
    d3 = w._dones(done)
    assert d1.shape == (64, 3) and d1.dtype == bool and d1.flags.c_contiguous and d3 is d1 and d2 is not d1
    assert d1[5].all() and not d1[4].any() and d2[4].all() and not d2[5].any()
