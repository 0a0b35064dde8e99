"""VDN mixer: q_tot = sum over agents (reference: modules/mixers/vdn.py:5-10)."""
import torch as th
import torch.nn as nn


class VDNMixer(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, agent_qs, batch):
        # parameter-free; the learner's fused path (pmb_mixer_fwd, VDN variant) does not go through here
        return th.sum(agent_qs, dim=2, keepdim=True)
