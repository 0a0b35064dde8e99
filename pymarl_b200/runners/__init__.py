"""Rollout side of the path: a vectorised runner that keeps the env batch, the episode batch and
the agent on the device (reference: runners/parallel_runner.py:88-204, which talks to one env process per pipe)."""
from .vector_runner import VectorRunner, SyntheticVectorEnv

REGISTRY = {"vector": VectorRunner}

__all__ = ["VectorRunner", "SyntheticVectorEnv", "REGISTRY"]
