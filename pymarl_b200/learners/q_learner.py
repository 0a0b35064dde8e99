"""QLearner: QMIX / VDN / IQL learner (reference: learners/q_learner.py:9-143).

train(batch, t_env, episode_num) issues ONE C-ABI call (pmb_qlearner_train_step) that runs
the whole step on the GPU: both agent unrolls, double-Q target selection, both mixers, the
masked TD loss, the hand-written backward (mixer backward + GRU BPTT), grad-norm clipping,
RMSprop and the hard target sync.  Parameters, gradients, RMSprop state and target
parameters live in four flat fp32 buffers; the nn.Parameters (reference names) are views.

Data parallel: with torch.distributed initialised (one process per GPU, NCCL) every rank
trains on its shard of the episodes; the un-normalised gradients and the five loss sums are
all-reduced once per step and every rank applies the identical update.
"""
import copy
import ctypes as C

import torch as th

from .. import _lib, flat as _flat, data_parallel
from ..modules.mixers.qmix import QMixer
from ..modules.mixers.vdn import VDNMixer


class FusedRMSprop:
    """State holder with torch.optim.RMSprop's state_dict format (square_avg, step) so that
    opt.th checkpoints interchange with the reference (q_learner.py:30,135,143).  The update
    itself is the fused clip + RMSprop kernel (pmb_clip_rmsprop_update)."""

    def __init__(self, params, lr, alpha, eps):
        self.params = list(params)
        self.defaults = dict(lr=lr, momentum=0, alpha=alpha, eps=eps, centered=False, weight_decay=0,
                             capturable=False, foreach=None, maximize=False, differentiable=False)
        self.square_avg = [th.zeros_like(p.data) for p in self.params]       # re-pointed into flat_sq by the learner
        self.step_count = 0

    def zero_grad(self):
        pass                                                                   # gradients are overwritten each step

    def state_dict(self):
        state = {i: {"step": th.tensor(float(self.step_count)), "square_avg": sq.detach().clone()}
                 for i, sq in enumerate(self.square_avg)} if self.step_count > 0 else {}
        group = dict(self.defaults)
        group["params"] = list(range(len(self.params)))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        for i, st in sd.get("state", {}).items():
            self.square_avg[int(i)].copy_(st["square_avg"])
            self.step_count = int(st.get("step", self.step_count))
        for g in sd.get("param_groups", [])[:1]:
            for k in ("lr", "alpha", "eps"):
                if k in g:
                    self.defaults[k] = g[k]


class QLearner:
    def __init__(self, mac, scheme, logger, args):
        self.args = args
        self.mac = mac
        self.logger = logger

        self.params = list(mac.parameters())
        self.last_target_update_episode = 0

        self.mixer = None
        if args.mixer is not None:
            if args.mixer == "vdn":
                self.mixer = VDNMixer()
            elif args.mixer == "qmix":
                self.mixer = QMixer(args)
            else:
                raise ValueError("Mixer {} not recognised.".format(args.mixer))
            self.params += list(self.mixer.parameters())
            self.target_mixer = copy.deepcopy(self.mixer)

        self.optimiser = FusedRMSprop(self.params, lr=args.lr, alpha=args.optim_alpha, eps=args.optim_eps)
        self.target_mac = copy.deepcopy(mac)
        self.log_stats_t = -self.args.learner_log_interval - 1

        self.precision = getattr(args, "precision", "fp32")
        self._flat = None               # dict(p, g, sq, target, layout, key)
        self._workspace = None
        self._stats = None
        self.last_stats = None          # device tensor [16] float64 of the latest step

    # ---- flat storage -----------------------------------------------------------------------
    def _layout_dims(self, B=1, T=2, O=None, S=None):
        a = self.args
        d_in = self.mac.agent.d_in
        O = d_in - (a.n_actions if a.obs_last_action else 0) - (a.n_agents if a.obs_agent_id else 0) if O is None else O
        S = int(getattr(self.mixer, "state_dim", 1)) if S is None else S
        return _lib.make_dims(B=B, T=T, N=a.n_agents, O=O, S=S, A=a.n_actions, H=a.rnn_hidden_dim,
                              E=getattr(a, "mixing_embed_dim", 1) if a.mixer == "qmix" else 1,
                              obs_last_action=a.obs_last_action, obs_agent_id=a.obs_agent_id, mixer=a.mixer,
                              double_q=a.double_q, precision=self.precision)

    def _ensure_flat(self):
        """(Re)build the four flat buffers when the parameters are not views of them (first
        call, or after cuda()/.to() moved the modules)."""
        dev = self.mac.agent.fc1.weight.device
        _lib.require_cuda(self.mac.agent.fc1.weight, "learner parameters (call learner.cuda())")
        dims = self._layout_dims()
        layout = _lib.flat_layout(dims)
        has_mixer_params = self.args.mixer == "qmix"
        f = self._flat
        if f is not None and f["p"].device == dev:
            ok = _flat.block_ptr(self.mac.agent, "agent", layout) == f["p"].data_ptr()
            ok = ok and _flat.block_ptr(self.target_mac.agent, "agent", layout) == f["target"].data_ptr()
            if has_mixer_params:
                ok = ok and _flat.block_ptr(self.mixer, "mixer", layout) == f["p"].data_ptr() + layout.n_agent * 4
                ok = ok and _flat.block_ptr(self.target_mixer, "mixer", layout) == f["target"].data_ptr() + layout.n_agent * 4
            if ok:
                return f
        n = layout.n_total
        new = dict(p=th.zeros(n, dtype=th.float32, device=dev), g=th.zeros(n, dtype=th.float32, device=dev),
                   sq=th.zeros(n, dtype=th.float32, device=dev), target=th.zeros(n, dtype=th.float32, device=dev),
                   layout=layout)
        _flat.bind(new["p"], layout, self.mac.agent, "agent", grad=new["g"])
        _flat.bind(new["target"], layout, self.target_mac.agent, "agent")
        if has_mixer_params:
            _flat.bind(new["p"], layout, self.mixer, "mixer", grad=new["g"])
            _flat.bind(new["target"], layout, self.target_mixer, "mixer")
        # RMSprop state: carry over existing values, then view into flat_sq (self.params order)
        base, sq_views = new["p"].data_ptr(), []
        for p, old in zip(self.params, self.optimiser.square_avg):
            off = (p.data_ptr() - base) // 4
            view = new["sq"][off:off + p.numel()].view(p.shape)
            view.copy_(old.to(dev))
            sq_views.append(view)
        self.optimiser.square_avg = sq_views
        self._flat = new
        return new

    def _ensure_workspace(self, dims, dev):
        need = _lib.lib().pmb_learner_workspace_bytes(C.byref(dims))
        if need < 0:
            _lib.check(1, "pmb_learner_workspace_bytes")
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != dev:
            self._workspace = None
            self._workspace = th.empty(need, dtype=th.uint8, device=dev)
        if self._stats is None or self._stats.device != dev:
            self._stats = th.zeros(_lib.STATS_LEN, dtype=th.float64, device=dev)
        return need

    def workspace_views(self, dims):
        """Named tensors over the step's workspace (tests / debugging)."""
        v = _lib.WsViews()
        _lib.check(_lib.lib().pmb_learner_workspace_views(C.byref(dims), _lib.ptr(self._workspace),
                                                          self._workspace.numel(), C.byref(v)), "workspace_views")
        base = self._workspace.data_ptr()
        f32 = self._workspace.view(th.float32)
        B, T, N, A, H, E = dims.B, dims.T, dims.N, dims.A, dims.H, dims.E
        R, M = B * N, B * (T - 1)
        W = N if dims.mixer == 0 else 1
        shapes = dict(x_on=(T, R, H), x_tg=(T, R, H), h_stash=(T + 1, R, H), gates=(T, R, 4 * H), q_on=(T, R, A),
                      q_tg=(T, R, A), chosen=(B, T - 1, N), tmax=(B, T - 1, N), q_tot=(B, T - 1, W),
                      t_tot=(B, T - 1, W), g=(B, T - 1, W), d_chosen=(B, T - 1, N))
        if dims.mixer == 2:
            shapes["raw_on"] = (M, (N + 3) * E)
        out = {}
        for k, shp in shapes.items():
            off = (getattr(v, k) - base) // 4
            n = 1
            for s in shp:
                n *= s
            out[k] = f32[off:off + n].view(shp)
        return out

    # ---- the step ---------------------------------------------------------------------------
    def train(self, batch, t_env: int, episode_num: int):
        a = self.args
        f = self._ensure_flat()
        dev = f["p"].device
        keep = []
        fields = {}
        names = ["obs", "actions", "avail_actions", "reward", "terminated", "filled"]
        if a.mixer == "qmix":
            names.append("state")
        ep_index = None
        if hasattr(batch, "ep_ids") and hasattr(batch, "buffer"):
            # zero-copy replay sample (IndexedEpisodeBatch): the kernels read the buffer's episodes in place
            ep_index = batch.ep_ids
            for k in names:
                fields[k] = batch.buffer.data.transition_data[k]
        else:
            for k in names:
                t = batch[k]
                fields[k] = t if t.is_cuda else t.to(dev, non_blocking=True)   # host batch: H2D here (run.py:214)
        obs = fields["obs"]
        B, T, N, O = obs.shape
        if ep_index is not None:
            B, T = batch.batch_size, batch.max_seq_length
        S = fields["state"].shape[-1] if "state" in fields else 1
        dims = self._layout_dims(B=B, T=T, O=O, S=S)
        pb = _lib.make_batch(fields, need_state=(a.mixer == "qmix"), keep=keep, ep_index=ep_index)
        need = self._ensure_workspace(dims, dev)

        do_sync = (episode_num - self.last_target_update_episode) / a.target_update_interval >= 1.0
        dp = data_parallel.is_active() and getattr(a, "data_parallel", True)
        hp = _lib.HParams(a.gamma, a.lr, a.optim_alpha, a.optim_eps, a.grad_norm_clip, int(do_sync), int(dp),
                          int(bool(getattr(a, "keep_q", False))))
        L = _lib.lib()
        s = _lib.stream_ptr(dev)
        _lib.check(L.pmb_qlearner_train_step(C.byref(dims), C.byref(pb), C.byref(hp), _lib.ptr(f["p"]),
                                             _lib.ptr(f["g"]), _lib.ptr(f["sq"]), _lib.ptr(f["target"]),
                                             _lib.ptr(self._workspace), need, _lib.ptr(self._stats), s),
                   "pmb_qlearner_train_step")
        if dp:
            # one exchange per step: gradients of sum((td*mask)^2) and the five loss sums
            data_parallel.allreduce_step(f["g"], self._stats[:5])
            scratch = self._workspace[:4096 * 4].view(th.float32)
            _lib.check(L.pmb_clip_rmsprop_update(f["layout"].n_total, _lib.ptr(f["p"]), _lib.ptr(f["g"]),
                                                 _lib.ptr(f["sq"]), _lib.ptr(f["target"]), int(do_sync),
                                                 _lib.ptr(self._stats), a.lr, a.optim_alpha, a.optim_eps,
                                                 a.grad_norm_clip, _lib.ptr(scratch), s), "pmb_clip_rmsprop_update")
        self.optimiser.step_count += 1
        self.last_stats = self._stats
        self._last_dims = dims

        if do_sync:
            self.last_target_update_episode = episode_num
            self.logger.console_logger.info("Updated target network")

        if t_env - self.log_stats_t >= a.learner_log_interval:
            st = self._stats.tolist()                                            # one D2H sync
            mask_elems = st[_lib.STAT_IDS["mask_sum"]]
            self.logger.log_stat("loss", st[_lib.STAT_IDS["td2_sum"]] / mask_elems, t_env)
            self.logger.log_stat("grad_norm", self._stats[_lib.STAT_IDS["grad_norm"]].float(), t_env)
            self.logger.log_stat("td_error_abs", st[_lib.STAT_IDS["tdabs_sum"]] / mask_elems, t_env)
            self.logger.log_stat("q_taken_mean", st[_lib.STAT_IDS["qtaken_sum"]] / (mask_elems * a.n_agents), t_env)
            self.logger.log_stat("target_mean", st[_lib.STAT_IDS["target_sum"]] / (mask_elems * a.n_agents), t_env)
            self.log_stats_t = t_env

    def stats(self):
        """The 5 logged scalars + grad_norm of the latest step as python floats (syncs)."""
        st = self.last_stats.tolist()
        m = st[0]
        n = self.args.n_agents
        return dict(loss=st[1] / m, grad_norm=st[5], td_error_abs=st[2] / m, q_taken_mean=st[3] / (m * n),
                    target_mean=st[4] / (m * n), mask_sum=m, clip_coef=st[7])

    def _update_targets(self):
        self.target_mac.load_state(self.mac)
        if self.mixer is not None:
            self.target_mixer.load_state_dict(self.mixer.state_dict())
        self.logger.console_logger.info("Updated target network")

    def cuda(self):
        self.mac.cuda()
        self.target_mac.cuda()
        if self.mixer is not None:
            self.mixer.cuda()
            self.target_mixer.cuda()
        self._ensure_flat()

    def save_models(self, path):
        self.mac.save_models(path)
        if self.mixer is not None:
            th.save(self.mixer.state_dict(), "{}/mixer.th".format(path))
        th.save(self.optimiser.state_dict(), "{}/opt.th".format(path))

    def load_models(self, path):
        self.mac.load_models(path)
        # Like the reference (q_learner.py:137-143): the target MAC loads the online weights
        # and the target mixer is left as it is.
        self.target_mac.load_models(path)
        if self.mixer is not None:
            self.mixer.load_state_dict(th.load("{}/mixer.th".format(path), map_location=lambda storage, loc: storage))
        self.optimiser.load_state_dict(th.load("{}/opt.th".format(path), map_location=lambda storage, loc: storage))
