"""QMIX monotonic mixing network (reference: modules/mixers/qmix.py:7-47).

Parameters keep the reference names (hyper_w_1, hyper_w_final, hyper_b_1, V.0, V.2);
forward() runs pmb_mixer_fwd: one GEMM for the four state hypernets + a fused
abs / ELU / dot-product mixing kernel.  Inference only (the learner owns the backward)."""
import ctypes as C

import numpy as np
import torch as th
import torch.nn as nn

from ... import _lib, flat as _flat


class QMixer(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.n_agents = args.n_agents
        self.state_dim = int(np.prod(args.state_shape))
        self.embed_dim = args.mixing_embed_dim
        self.hyper_w_1 = nn.Linear(self.state_dim, self.embed_dim * self.n_agents)
        self.hyper_w_final = nn.Linear(self.state_dim, self.embed_dim)
        self.hyper_b_1 = nn.Linear(self.state_dim, self.embed_dim)
        self.V = nn.Sequential(nn.Linear(self.state_dim, self.embed_dim), nn.ReLU(), nn.Linear(self.embed_dim, 1))

    @th.no_grad()
    def forward(self, agent_qs, states):
        """agent_qs [B, T', N], states [B, T', S] -> q_tot [B, T', 1]."""
        _lib.require_cuda(agent_qs, "agent_qs")
        bs, tp = agent_qs.shape[0], agent_qs.shape[1]
        qs = agent_qs.detach().to(th.float32).contiguous()
        st = states.to(th.float32).reshape(bs, tp, self.state_dim).contiguous()
        # T = tp + 1 so that the kernel's [B, T-1] row space is exactly these rows
        dims = _lib.make_dims(B=bs, T=tp + 1, N=self.n_agents, O=1, S=self.state_dim, A=1, H=16,
                              E=self.embed_dim, mixer="qmix")
        flat_mixer = C.c_void_p(_flat.ensure_block(self, "mixer", dims))
        batch = _lib.Batch()
        batch.state = st.data_ptr()
        batch.state_sb = tp * self.state_dim
        raw = th.empty(bs * tp, (self.n_agents + 3) * self.embed_dim, dtype=th.float32, device=qs.device)
        q_tot = th.empty(bs, tp, 1, dtype=th.float32, device=qs.device)
        _lib.check(_lib.lib().pmb_mixer_fwd(C.byref(dims), C.byref(batch), flat_mixer,
                                            _lib.ptr(qs), 0, _lib.ptr(raw), _lib.ptr(q_tot),
                                            _lib.stream_ptr(qs.device)), "pmb_mixer_fwd")
        return q_tot
