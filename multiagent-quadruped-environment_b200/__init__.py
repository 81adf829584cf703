"""mqe-b200: B200-native replacement for the Go1.step() hot path of MQE.

Import as `mqe_b200` (see ../mqe_b200/__init__.py).  Only the hot path of SURVEY.md section 8
lives here: csrc/ (CUDA kernels + C-ABI), engine.py (ctypes binding), envs/ (the mqe VecEnv
surface: make_mqe_env / Go1 / wrappers / configs), terrain/ (BarrierTrack restatement), model.py.
"""
__version__ = "0.1.0"
