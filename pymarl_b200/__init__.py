"""pymarl_b200 - a B200-native (sm_100a CUDA) implementation of PyMARL's data-parallel hot
path: the QMIX / VDN / IQL ``QLearner.train`` step and the batched ``BasicMAC`` forward /
``select_actions``, behind the reference's registry and call surface.

    from pymarl_b200 import le_REGISTRY, mac_REGISTRY      # same keys as the reference
    learner = le_REGISTRY["q_learner"](mac, scheme, logger, args)

``install_into_reference()`` swaps these classes into an importable reference tree so the
reference's own ``run.py`` drives them.
"""
from .learners import REGISTRY as le_REGISTRY
from .controllers import REGISTRY as mac_REGISTRY
from .modules.agents import REGISTRY as agent_REGISTRY
from .components.action_selectors import REGISTRY as action_REGISTRY
from .components.episode_buffer import EpisodeBatch, ReplayBuffer, IndexedEpisodeBatch
from .modules.mixers.qmix import QMixer
from .modules.mixers.vdn import VDNMixer

__all__ = ["le_REGISTRY", "mac_REGISTRY", "agent_REGISTRY", "action_REGISTRY", "EpisodeBatch", "ReplayBuffer",
           "QMixer", "VDNMixer", "install_into_reference"]


def install_into_reference():
    """Register the CUDA-backed classes under the reference's registry keys.  Requires the
    reference's ``src`` directory on sys.path (its modules use top-level absolute imports)."""
    import learners as ref_learners
    import controllers as ref_controllers
    import modules.agents as ref_agents
    import components.action_selectors as ref_selectors
    ref_learners.REGISTRY["q_learner"] = le_REGISTRY["q_learner"]
    ref_learners.REGISTRY["coma_learner"] = le_REGISTRY["coma_learner"]
    ref_selectors.REGISTRY["multinomial"] = action_REGISTRY["multinomial"]
    ref_controllers.REGISTRY["basic_mac"] = mac_REGISTRY["basic_mac"]
    ref_agents.REGISTRY["rnn"] = agent_REGISTRY["rnn"]
    ref_selectors.REGISTRY["epsilon_greedy"] = action_REGISTRY["epsilon_greedy"]
