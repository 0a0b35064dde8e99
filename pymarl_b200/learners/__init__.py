from .q_learner import QLearner

REGISTRY = {}

REGISTRY["q_learner"] = QLearner
from .coma_learner import COMALearner

REGISTRY["coma_learner"] = COMALearner
