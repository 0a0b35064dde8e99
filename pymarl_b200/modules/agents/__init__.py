REGISTRY = {}

from .rnn_agent import RNNAgent  # noqa: E402

REGISTRY["rnn"] = RNNAgent
