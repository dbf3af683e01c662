"""The agent plugin API that ``rlTrain`` drives (unchanged names and meaning).

Mirrors smartstart/reinforcementLearningCore/agents_abstract_classes.py:6-91:
RLAgent (get_action / observe / render / start_new_episode / end_episode /
get_param_dict), NavigationRLAgent.start_new_episode_plan,
ValueFuncRLAgent.get_state_value, ReplayBufferRLAgent.  A base agent written against
the reference's classes can subclass these instead without edits.
"""
from __future__ import annotations

import abc


class RLAgent(metaclass=abc.ABCMeta):
    @abc.abstractmethod
    def get_action(self, state):
        """Action to take in ``state``."""

    @abc.abstractmethod
    def observe(self, state, action, reward, new_state, done):
        """Digest one transition."""

    @abc.abstractmethod
    def render(self, env, **kwargs):
        """Render the environment the agent lives in."""

    def start_new_episode(self, state):
        pass

    def end_episode(self):
        pass

    def get_param_dict(self):
        raise NotImplementedError(
            "Agent hasn't overridden get_param_dict, if you don't wish to implement it just return None")


class NavigationRLAgent(RLAgent):
    """Agent that plans along a list of previously visited states at episode start."""

    def start_new_episode_plan(self, state, path_to_follow):
        pass


class ValueFuncRLAgent(RLAgent):
    @abc.abstractmethod
    def get_state_value(self, state):
        """V(s) (max_a Q(s, a)) for one state or a batch [m, d] -> [m, 1]."""


class ReplayBufferRLAgent(RLAgent):
    def __init__(self):
        self.replay_buffer = None

    def set_replay_buffer_main_agent(self, new_main_agent):
        self.replay_buffer.set_main_agent(new_main_agent)
