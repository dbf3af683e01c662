"""smartstartcontinuous_b200 -- B200-native (sm_100a) SmartStart start-state selection
(Gaussian KDE + UCB + argmax) and NND_MB random-shooting MPC navigation, behind the
unchanged SmartStartContinuous / NND_MB_agent interfaces of darren-huang/SmartStartContinuous.

Python host code -> ctypes -> C ABI (include/ss_b200.h) -> hand-written CUDA kernels.
No CPU fallback: importing the agents works anywhere, using them needs libss_b200.so and a B200.
"""
from .agents_abstract_classes import (NavigationRLAgent, ReplayBufferRLAgent, RLAgent,  # noqa: F401
                                      ValueFuncRLAgent)
from .replay_buffer import ReplayBuffer  # noqa: F401

__all__ = ["Engine", "SmartStartContinuous", "NND_MB_agent", "ReplayBuffer", "RLAgent",
           "NavigationRLAgent", "ValueFuncRLAgent", "ReplayBufferRLAgent"]


def __getattr__(name):
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "SmartStartContinuous":
        from .smart_start import SmartStartContinuous
        return SmartStartContinuous
    if name == "NND_MB_agent":
        from .nnd_mb_agent import NND_MB_agent
        return NND_MB_agent
    raise AttributeError(name)
