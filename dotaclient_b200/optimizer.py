"""``DotaOptimizer`` -- drop-in for the hot path of the reference's ``optimizer.py``.

Keeps the reference's module surface (``DotaOptimizer``, ``Sequence``, ``MessageQueue``,
``advantage_returns``, ``discount``, ``init_distribution``, ``main``, the CLI flags) while one
optimizer step runs as:  unit-encoder kernel chain + tcgen05 3xTF32 GEMMs -> hand-written recurrence kernels ->
fused PPO loss+grad kernel -> autograd backward through the same kernels -> ONE NCCL all-reduce of
a flat gradient buffer -> fused count-divide / grad-norm / clip / Adam kernel -- replayed from a CUDA graph
when the batch is device-resident.  CUDA only.

Line references are to TimZaman/dotaclient ``optimizer.py`` @ 8615b90.
"""
import argparse
import io
import logging
import math
import os
import pickle
import queue
import re
import socket
import threading
import time
from datetime import datetime

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, ops
from .distributed import DistributedDataParallelSparseParamCPU
from .flat import FlatParameterSpace, VALUE_SLOT
from .policy import Policy, REWARD_KEYS

logging.basicConfig(format='%(asctime)s %(levelname)-8s %(message)s')
logger = logging.getLogger(__name__)
logger.setLevel(logging.INFO)

eps = np.finfo(np.float32).eps.item()                                  # :38
GAMMA, LAMBDA = 0.98, 0.97                                              # :421


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("dotaclient_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device('cuda', torch.cuda.current_device())


def is_distributed():                                                   # :42-43
    return dist.is_available() and dist.is_initialized()


def is_master():                                                        # :46-50
    return dist.get_rank() == 0 if is_distributed() else True


# ------------------------------------------------------------------------------------------ GAE
def advantage_returns(rewards, values, gamma, lam):
    """GAE advantages and rewards-to-go (:57-64) on the GPU scan kernel.

    ``rewards`` / ``values`` carry one trailing bootstrap element like the reference's
    (:417-420, both 0 for terminated rollouts).  numpy in -> numpy out; torch in -> torch (cuda) out.
    """
    as_numpy = isinstance(rewards, np.ndarray)
    dev = _device()
    r = torch.as_tensor(rewards, dtype=torch.float32).to(dev)
    v = torch.as_tensor(values, dtype=torch.float32).to(dev)
    n = r.numel() - 1
    seg = torch.tensor([0, n], dtype=torch.int64, device=dev)
    adv, ret = ops.gae_scan(r[:n], v[:n], seg, gamma=gamma, lam=lam, boot_value=v[n:n + 1].contiguous(),
                            boot_reward=r[n:n + 1].contiguous())
    if as_numpy:
        return adv.cpu().numpy(), ret.cpu().numpy()
    return adv, ret


def discount(x, gamma):
    """Reverse discounted cumulative sum (:53-54), float64 accumulate, fp32 result."""
    as_numpy = isinstance(x, np.ndarray)
    dev = _device()
    xt = torch.as_tensor(x, dtype=torch.float32).to(dev).contiguous()
    seg = torch.tensor([0, xt.numel()], dtype=torch.int64, device=dev)
    _, ret = ops.gae_scan(xt, torch.zeros_like(xt), seg, gamma=gamma, lam=1.0)
    return ret.cpu().numpy() if as_numpy else ret


# ------------------------------------------------------------------------------------------ broker
class _Delivery:
    def __init__(self, tag):
        self.delivery_tag = tag


class _Broker:
    """In-process stand-in for the RabbitMQ broker: one 'experience' queue, one 'model' slot."""
    _registry = {}
    _lock = threading.Lock()

    def __init__(self):
        self.experience = queue.Queue()
        self.model = None            # (body, headers): x-recent-history exchange of length 1 (:118-122)
        self.model_lock = threading.Lock()
        self.tag = 0

    @classmethod
    def get(cls, host, port):
        with cls._lock:
            return cls._registry.setdefault((host, port), _Broker())


class MessageQueue:
    """Same interface as the reference's pika client (:67-174), backed by an in-process broker.

    The real AMQP transport is out of scope (no broker, no pika here); actors running in the same
    process publish with ``publish_experience`` and read weights with ``latest_model``.
    """
    EXPERIENCE_QUEUE_NAME = 'experience'
    MODEL_EXCHANGE_NAME = 'model'
    MAX_RETRIES = 10

    def __init__(self, host, port, prefetch_count, use_model_exchange):
        self.host, self.port = host, port
        self.prefetch_count = prefetch_count
        self.use_model_exchange = use_model_exchange
        self._broker = None

    def connect(self):
        if self._broker is None:
            self._broker = _Broker.get(self.host, self.port)

    @property
    def xp_queue_size(self):
        return self._broker.experience.qsize() if self._broker else None

    def process_events(self):
        pass

    def process_data_events(self):          # heartbeat in the reference (:132-137)
        pass

    def publish_model(self, msg, hdr):
        with self._broker.model_lock:
            self._broker.model = (msg, dict(hdr))

    def consume_xp(self, timeout=None):
        body = self._broker.experience.get(timeout=timeout)
        self._broker.tag += 1
        return _Delivery(self._broker.tag), None, body

    def close(self):
        self._broker = None

    # -- actor side of the in-process broker ------------------------------------------------
    def publish_experience(self, body):
        self._broker.experience.put(body)

    def latest_model(self):
        with self._broker.model_lock:
            return self._broker.model


# ------------------------------------------------------------------------------------------ records
class Sequence:
    """One ``seq_len`` chunk of a rollout (:176-190).

    ``log_probs_sel`` (compact per-head vectors, reference form) is derived lazily from the dense
    ``old_logp [S, 5]`` the kernels produce, so building a Sequence never syncs the device.
    """

    def __init__(self, game_id, weight_version, team_id, observations, actions, masks, values, rewards, hidden,
                 log_probs_sel=None, old_logp=None):
        self.game_id = game_id
        self.weight_version = weight_version
        self.team_id = team_id
        self.observations = observations
        self.actions = actions
        self.masks = masks
        self.rewards = rewards
        self.values = values
        self.hidden = hidden
        self._log_probs_sel = log_probs_sel
        self.old_logp = old_logp
        self.advantages = None
        self.returns = None

    @property
    def log_probs_sel(self):
        if self._log_probs_sel is None and self.old_logp is not None:
            self._log_probs_sel = {}
            for h, key in enumerate(ops.HEAD_KEYS):
                step = self.actions[key].bool().any(dim=-1)
                self._log_probs_sel[key] = self.old_logp[:, h][step]
        return self._log_probs_sel

    def dense_old_logp(self):
        if self.old_logp is None:
            ref = self.actions['enum']
            dense = torch.zeros((ref.shape[0], 5), dtype=torch.float32, device=ref.device)
            for h, key in enumerate(ops.HEAD_KEYS):
                step = self.actions[key].bool().any(dim=-1)
                dense[step, h] = self._log_probs_sel[key].to(ref.device)
            self.old_logp = dense
        return self.old_logp


class ExperienceBatch:
    """A training batch stacked time-major ``[S, B, ...]`` -- the layout the kernels consume.

    ``DotaOptimizer.train`` accepts either a list of ``Sequence`` (reference API) or one of these.
    Tensors may live in pinned host memory; ``to(device)`` issues the asynchronous H2D copies.
    """
    FIELDS = ("advantages", "returns", "old_logp", "h0", "c0")

    def __init__(self, observations, masks, actions, old_logp, advantages, returns, h0, c0=None):
        self.observations, self.masks, self.actions = observations, masks, actions
        self.old_logp, self.advantages, self.returns, self.h0, self.c0 = old_logp, advantages, returns, h0, c0

    def __del__(self):
        # a batch that was uploaded (prefetched) but never trained on must not leave its ready-events behind: a later tensor
        # allocated at the same address would otherwise match a stale event
        try:
            for ptr in getattr(self, "_h2d_ptrs", ()):
                ops.H2D_EVENTS.pop(ptr, None)
        except Exception:                       # interpreter shutdown: module globals may already be gone
            pass

    @property
    def seq_len(self):
        return self.advantages.shape[0]

    @property
    def batch_size(self):
        return self.advantages.shape[1]

    def tensors(self):
        for d in (self.observations, self.masks, self.actions):
            for k, v in d.items():
                yield d, k, v
        for f in self.FIELDS:
            v = getattr(self, f)
            if v is not None:
                yield self, f, v

    def nbytes(self):
        return sum(v.numel() * v.element_size() for _, _, v in self.tensors())

    def to(self, device, non_blocking=True, prefetch=False):
        """Host -> device.  From pinned memory the copies are issued on a side stream in the order the step consumes
        them (env, states, unit groups, then the loss inputs) and every tensor gets an event: ``train`` makes the
        compute stream wait per tensor right before first use, so the PCIe transfer of unit group g+1 overlaps the
        encoder kernels of group g and the loss inputs arrive during forward/backward."""
        out = ExperienceBatch({}, {}, {}, None, None, None, None, None)
        overlap = non_blocking and self.advantages.is_pinned()
        compute = torch.cuda.current_stream(device)
        side = _copy_stream(device) if overlap else None
        if overlap and not prefetch:
            side.wait_stream(compute)            # prefetch=True: the copies start now, next to whatever the compute stream is doing
        def priority(item):                      # copy order == order of first use in the step
            holder, k, _ = item
            if holder is self.observations:
                return 0 if k == 'env' else 3 + Policy.INPUT_KEYS.index(k)
            if not isinstance(holder, dict) and k in ('h0', 'c0'):
                return 1
            return 100
        items = sorted(self.tensors(), key=priority)
        out._h2d_ptrs = []
        for holder, k, v in items:
            if overlap:
                with torch.cuda.stream(side):
                    moved = v.to(device, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(side)
                moved.record_stream(compute)
                ops.H2D_EVENTS[moved.data_ptr()] = ev
                out._h2d_ptrs.append(moved.data_ptr())
            else:
                moved = v.to(device, non_blocking=non_blocking)
            if isinstance(holder, dict):
                target = out.observations if holder is self.observations else out.masks if holder is self.masks else out.actions
                target[k] = moved
            else:
                setattr(out, k, moved)
        return out

    def pin_memory(self):
        out = ExperienceBatch({}, {}, {}, None, None, None, None, None)
        for holder, k, v in self.tensors():
            moved = v.cpu().pin_memory()
            if isinstance(holder, dict):
                target = out.observations if holder is self.observations else out.masks if holder is self.masks else out.actions
                target[k] = moved
            else:
                setattr(out, k, moved)
        return out

    @staticmethod
    def from_sequences(experiences, device):
        """Stacks ``Sequence`` records along dim 1 (the reference stacks along dim 0, :587-615)."""
        def stack(ts):
            return torch.stack([t.to(device) for t in ts], dim=1)
        obs = {k: stack([e.observations[k] for e in experiences]) for k in Policy.INPUT_KEYS}
        masks = {k: stack([e.masks[k] for e in experiences]) for k in Policy.OUTPUT_KEYS}
        actions = {k: stack([e.actions[k] for e in experiences]) for k in Policy.OUTPUT_KEYS}
        old = stack([e.dense_old_logp() for e in experiences])
        adv = stack([torch.as_tensor(e.advantages) for e in experiences])
        ret = stack([torch.as_tensor(e.returns) for e in experiences])
        if isinstance(experiences[0].hidden, tuple):
            h0 = torch.cat([e.hidden[0].to(device) for e in experiences], dim=1)
            c0 = torch.cat([e.hidden[1].to(device) for e in experiences], dim=1)
        else:
            h0, c0 = torch.cat([e.hidden.to(device) for e in experiences], dim=1), None      # :591
        return ExperienceBatch(obs, masks, actions, old, adv, ret, h0.detach(), None if c0 is None else c0.detach())


_copy_streams = {}


def _copy_stream(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=device)
    return _copy_streams[key]


def all_gather(t):                                                        # :193-196 (unused by the reference too)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.cat(out)


# ------------------------------------------------------------------------------------------ optimizer
class DotaOptimizer:
    MODEL_FILENAME_FMT = "model_%09d.pt"
    ADAM_FILENAME_FMT = "adam_%09d.state"         # extension: Adam moments of the same iteration (torch.optim.Adam layout)
    ADAM_FILES_KEPT = 3
    BUCKET_NAME = 'dotaservice'
    MODEL_HISTOGRAM_FREQ = 128
    MAX_GRAD_NORM = 0.5
    SPEED_KEY = 'steps per s'
    ADAM_BETAS = (0.9, 0.999)       # torch.optim.Adam defaults (:275)
    ADAM_EPS = 1e-8

    def __init__(self, rmq_host, rmq_port, epochs, min_seq_per_epoch, seq_len,
                 learning_rate, checkpoint, pretrained_model, mq_prefetch_count, log_dir,
                 entropy_coef, vf_coef, run_local, *, hidden_size=256, cell="gru", mq=None, iterations=100000,
                 rollout_prefetch=0):
        self.rmq_host, self.rmq_port = rmq_host, rmq_port
        self.epochs = epochs
        self.min_seq_per_epoch = min_seq_per_epoch
        self.seq_len = seq_len
        self.learning_rate = learning_rate
        self.checkpoint = checkpoint
        self.mq_prefetch_count = mq_prefetch_count
        self.iteration_start = 1
        self.log_dir = log_dir
        self.entropy_coef = entropy_coef
        self.vf_coef = vf_coef
        self.run_local = run_local
        self.iterations = iterations        # :226
        self.model_upload_freq = 10         # :227
        self.e_clip = 0.1                   # :229
        self.device = _device()
        _lib.load()                         # fail loudly, up front, if the CUDA library is missing

        with torch.random.fork_rng(devices=[]):     # :34 seeds torch with 7 at import and Policy() init depends on it;
            torch.manual_seed(7)                    # forked so that constructing an optimizer leaves the caller's RNG alone
            self.policy_base = Policy(hidden_size=hidden_size, cell=cell)

        if self.checkpoint:
            logger.info('Checkpointing to: {}'.format(self.log_dir))
            os.makedirs(self.log_dir, exist_ok=True)
            if not self.run_local:
                logger.warning('GCS is out of scope for dotaclient_b200; checkpoints stay in %s', self.log_dir)
            latest_model = self.get_latest_model(prefix=self.log_dir)
            if latest_model is not None:
                logger.info('Found a latest model in pretrained dir: {}'.format(latest_model))
                if pretrained_model is not None:
                    logger.warning('Overriding pretrained model by latest model.')
                pretrained_model = os.path.join(self.log_dir, latest_model)
            if pretrained_model is not None:
                self.iteration_start = self.iteration_from_model_filename(filename=pretrained_model) + 1   # :253
        if pretrained_model is not None:
            self.policy_base.load_state_dict(torch.load(pretrained_model, map_location='cpu'), strict=False)  # :263-266

        self.policy_base.to(self.device)
        self.flat = FlatParameterSpace(self.policy_base, self.device)
        if is_distributed():
            self.policy = DistributedDataParallelSparseParamCPU(self.policy_base, flat_space=self.flat)   # :268-269
        else:
            self.policy = self.policy_base

        # Adam state lives next to the flat parameter buffer; `optimizer` keeps a torch-like handle (:275).
        self.exp_avg = torch.zeros_like(self.flat.param)
        self.exp_avg_sq = torch.zeros_like(self.flat.param)
        self.adam_steps = torch.zeros(self.flat.n_seg, dtype=torch.int32, device=self.device)
        self.optimizer = _FusedAdamHandle(self)
        if self.checkpoint and pretrained_model is not None:
            adam_file = os.path.join(os.path.dirname(pretrained_model), self.ADAM_FILENAME_FMT % (self.iteration_start - 1))
            if os.path.isfile(adam_file):
                logger.info('Restoring Adam state from {}'.format(adam_file))
                self.optimizer.load_state_dict(torch.load(adam_file, map_location='cpu'))
        self._sync_resume_state()
        self._n_actions = torch.zeros(8, dtype=torch.int32, device=self.device)
        self._n_actions[VALUE_SLOT] = 1 if vf_coef > 0 else 0
        self._metrics = torch.zeros(4, dtype=torch.float32, device=self.device)
        self._finish_ws = torch.zeros(_lib.FINISH_WORKSPACE_BYTES, dtype=torch.uint8, device=self.device)
        self._host_result = torch.zeros(_lib.LOSS_SLOTS + 4, dtype=torch.float32).pin_memory()
        self.last_step_launch_estimate = 0
        self._staging, self._staging_event = {}, None    # pinned host staging of the batched experience prep
        # rollout_prefetch > 0: a background thread pulls and unpickles up to that many rollouts ahead (same order, same
        # rollouts as the reference's one-at-a-time loop, :448-466) while the GPU trains on the current iteration
        self.rollout_prefetch = int(rollout_prefetch)
        self._rollout_q, self._prefetch_thread = None, None
        self.use_cuda_graph = True          # replay device-resident batches of a known shape from a captured graph of the step
        self._graphs = {}
        self._input_slots = {}              # batch shape -> up to two sets of static device input buffers (prefetch + graph)
        self._last_iteration_shape = None
        self.time_last_it = time.time()

        self.mq = mq if mq is not None else MessageQueue(host=self.rmq_host, port=self.rmq_port,
                                                        prefetch_count=mq_prefetch_count,
                                                        use_model_exchange=self.checkpoint)
        self.mq.connect()
        self.upload_model(version=self.iteration_start)                  # :284

    def close(self):
        """Releases the captured step graphs (they hold NCCL work when data-parallel: destroy them BEFORE
        ``dist.destroy_process_group()``, which otherwise waits forever) and the pinned staging buffers."""
        self._graphs.clear()
        self._input_slots.clear()
        self._staging.clear()
        import gc
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def _sync_resume_state(self):
        """Data-parallel resume: only the master scans ``log_dir`` and restores (``checkpoint = is_master()``, :751), so the
        restored Adam moments, step counters and ``iteration_start`` are broadcast from rank 0 -- otherwise the replicas
        would apply different Adam updates to the same all-reduced gradient and silently diverge.  (The weights themselves
        are broadcast by the wrapper's ``sync_parameters``, distributed.py:71-74.)"""
        if not is_distributed():
            return
        dist.broadcast(self.exp_avg, 0)
        dist.broadcast(self.exp_avg_sq, 0)
        dist.broadcast(self.adam_steps, 0)
        it = torch.tensor([self.iteration_start], dtype=torch.int64, device=self.device)
        dist.broadcast(it, 0)
        self.iteration_start = int(it.item())

    # -- checkpoints (:287-308, :697-723) --------------------------------------------------------
    @staticmethod
    def iteration_from_model_filename(filename):
        return int(re.search(r'(\d+)(?=.pt)', filename).group(0))

    def get_latest_model(self, prefix):
        """Lexicographically latest ``*.pt`` in ``prefix`` (local listing; the reference's local branch
        is broken -- ``os.path.isfile`` relative to cwd and ``.name`` on a str, :296,302 -- this is the intent)."""
        if not os.path.isdir(prefix):
            return None
        names = sorted(f for f in os.listdir(prefix) if f.endswith('.pt') and os.path.isfile(os.path.join(prefix, f)))
        return names[-1] if names else None

    def upload_model(self, version):
        if not is_master():
            return
        buffer = io.BytesIO()
        state_dict = {k: v.detach().cpu().clone() for k, v in self.policy_base.state_dict().items()}
        torch.save(obj=state_dict, f=buffer)                              # same bytes-format as :705-709
        state_dict_b = buffer.getvalue()
        if self.checkpoint:
            with open(os.path.join(self.log_dir, self.MODEL_FILENAME_FMT % version), 'wb') as f:
                f.write(state_dict_b)
            # extension: the Adam moments next to the weights, in torch.optim.Adam's own format.  The name
            # does not end in .pt, so the reference's "latest *.pt" scan and its agents never see it.
            torch.save(self.optimizer.state_dict(), os.path.join(self.log_dir, self.ADAM_FILENAME_FMT % version))
            stale = sorted(f for f in os.listdir(self.log_dir) if re.fullmatch(r'adam_\d{9}\.state', f))[:-self.ADAM_FILES_KEPT]
            for f in stale:                                                 # resume only ever needs the newest: bound the disk growth
                os.remove(os.path.join(self.log_dir, f))
        self.mq.publish_model(msg=state_dict_b, hdr={'version': version})   # :716

    # -- experience intake (:314-430) -------------------------------------------------------------
    def get_rollout(self):
        method, properties, body = self.mq.consume_xp()
        return self._describe_rollout(pickle.loads(body))

    @staticmethod
    def _describe_rollout(data):
        rollout_len = data['rewards'].shape[0]
        subrewards = data['rewards'].sum(axis=0)
        return data, subrewards, rollout_len, data['weight_version'], data.get('canvas')

    def _next_rollout(self):
        """``get_rollout()``, optionally served by the decode-ahead thread -- same rollouts, same order as the reference's
        one-at-a-time loop (:448-466).  (A pool of decoder PROCESSES was measured and dropped: ``pickle.loads`` of a 2.7 MB
        rollout is 4 ms here, shipping the bytes out and the tensors back through shared memory costs 9x that.)"""
        if self.rollout_prefetch <= 0:
            return self.get_rollout()
        if self._prefetch_thread is None:
            self._rollout_q = queue.Queue(maxsize=self.rollout_prefetch)

            def pump():
                while True:
                    self._rollout_q.put(self.get_rollout())
            self._prefetch_thread = threading.Thread(target=pump, daemon=True, name="dc-rollout-decode")
            self._prefetch_thread.start()
        return self._rollout_q.get()

    def experiences_from_rollout(self, data):
        """Rollout -> list of ``Sequence`` (:328-430), in ONE padded pass instead of a per-chunk loop.

        Running the recurrence over the whole zero-padded rollout from the zero state is exactly the
        reference's chunk-by-chunk forward with carried hidden (:340,384-385); the hidden state entering
        chunk i is read back from the recurrence's state buffer.  Old log-probs come from the fused
        selected-log-prob kernel, advantages/returns from the GAE scan kernel (padding inside the scan).
        """
        S, dev = self.seq_len, self.device
        L = data['rewards'].shape[0]
        n_chunks = (L + S - 1) // S
        Lp = n_chunks * S

        def padded(t):
            t = torch.as_tensor(t).to(dev)
            if Lp == L:
                return t
            out = torch.zeros((Lp,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)      # :367-382
            out[:L] = t
            return out

        obs = {k: padded(v) for k, v in data['observations'].items()}
        masks = {k: padded(v).bool() for k, v in data['masks'].items()}
        actions = {k: padded(v).bool() for k, v in data['actions'].items()}
        rewards_np = np.asarray(data['rewards'], dtype=np.float32)
        if Lp != L:
            rewards_np = np.pad(rewards_np, ((0, Lp - L), (0, 0)), mode='constant')
        rewards = torch.from_numpy(rewards_np).to(dev)
        with torch.no_grad():
            pol = self.policy_base
            x, unit_embedding = pol._encode(obs['env'].unsqueeze(1),
                                            [obs[k].unsqueeze(1) for k in Policy.INPUT_KEYS[1:]])
            hidden = pol.init_hidden()
            hidden = tuple(h.to(dev) for h in hidden) if isinstance(hidden, tuple) else hidden.to(dev)
            r = pol.rnn
            h0 = hidden[0][0] if pol.cell == "lstm" else hidden[0]
            c0 = hidden[1][0] if pol.cell == "lstm" else None
            ybuf, cbuf = ops.rnn_forward_states(x.contiguous(), r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0,
                                                r.bias_hh_l0, h0, c0, pol.cell)
            logits, values = pol._heads(ybuf[1:], unit_embedding)
            keys = ops.HEAD_KEYS
            old_logp = ops.selected_logp([logits[k] for k in keys], [masks[k] for k in keys],
                                         [actions[k] for k in keys])                         # :387-390
            seg = torch.tensor([0, Lp], dtype=torch.int64, device=dev)
            adv, ret = ops.gae_scan(rewards, values.reshape(-1), seg, gamma=GAMMA, lam=LAMBDA)   # :417-421
        sequences = []
        for i in range(n_chunks):
            sl = slice(i * S, (i + 1) * S)
            if pol.cell == "lstm":
                hid = (ybuf[i * S].unsqueeze(0), cbuf[i * S].unsqueeze(0))
            else:
                hid = ybuf[i * S].unsqueeze(0)
            seq = Sequence(game_id=data.get('game_id'), weight_version=data.get('weight_version'),
                           team_id=data.get('team_id'),
                           observations={k: v[sl] for k, v in obs.items()},
                           actions={k: v[sl] for k, v in actions.items()},
                           masks={k: v[sl] for k, v in masks.items()},
                           values=values[sl].reshape(1, S, 1), rewards=rewards_np[sl], hidden=hid,
                           old_logp=old_logp[sl])
            seq.advantages = adv[sl]
            seq.returns = ret[sl]
            sequences.append(seq)
        return sequences

    def _prepare_rollouts(self, datas):
        """The batched no-grad half of an iteration: all rollouts become the batch dimension of ONE
        time-major ``[L_max, R, ...]`` pass -- encoder chain, recurrence from the zero state, heads, selected log-probs
        (:384-390) -- followed by ONE segmented GAE scan over every rollout's own padded length (:417-421).  Returns the raw
        device tensors; ``experiences_from_rollouts`` / ``batch_from_rollouts`` slice them."""
        S, dev, pol = self.seq_len, self.device, self.policy_base
        R = len(datas)
        Ls = [int(d['rewards'].shape[0]) for d in datas]
        Lps = [(L + S - 1) // S * S for L in Ls]
        Lmax = max(Lps)
        same = all(L == Lmax for L in Ls)

        if self._staging_event is not None:
            self._staging_event.synchronize()          # the previous iteration's uploads have left the staging buffers

        def batched(group, key, dtype):
            """Rollouts -> one time-major ``[Lmax, R, ...]`` device tensor.  The host side only does contiguous per-rollout
            copies into a cached PINNED ``[R, Lmax, ...]`` staging buffer (memcpy speed; stacking time-major on the host is a
            768-byte-granular scatter, 3x slower) and uploads asynchronously; the transposition to time-major runs on the GPU."""
            srcs = [torch.as_tensor(d[group][key]) for d in datas]
            shape = (R, Lmax) + tuple(srcs[0].shape[1:])
            buf = self._staging.get((group, key))
            if buf is None or buf.shape != shape or buf.dtype != dtype:
                buf = torch.empty(shape, dtype=dtype).pin_memory()
                self._staging[(group, key)] = buf
            for i, t in enumerate(srcs):
                buf[i, :Ls[i]].copy_(t)
                if Ls[i] < Lmax:
                    buf[i, Ls[i]:].zero_()                                             # zero padding (:367-382)
            return buf.to(dev, non_blocking=True).transpose(0, 1).contiguous()

        obs = {k: batched('observations', k, torch.float32) for k in Policy.INPUT_KEYS}
        masks = {k: batched('masks', k, torch.bool) for k in Policy.OUTPUT_KEYS}
        actions = {k: batched('actions', k, torch.bool) for k in Policy.OUTPUT_KEYS}
        self._staging_event = torch.cuda.Event()
        self._staging_event.record()
        rewards_np = np.zeros((R, Lmax, len(REWARD_KEYS)), dtype=np.float32)
        for i, d in enumerate(datas):
            rewards_np[i, :Ls[i]] = np.asarray(d['rewards'], dtype=np.float32)
        with torch.no_grad():
            x, unit_embedding = pol._encode(obs['env'], [obs[k] for k in Policy.INPUT_KEYS[1:]])
            r = pol.rnn
            h0 = torch.zeros((R, pol.hidden_size), dtype=torch.float32, device=dev)
            c0 = torch.zeros_like(h0) if pol.cell == "lstm" else None
            ybuf, cbuf = ops.rnn_forward_states(x.contiguous(), r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0,
                                                h0, c0, pol.cell)
            logits, values = pol._heads(ybuf[1:], unit_embedding)
            keys = ops.HEAD_KEYS
            old_logp = ops.selected_logp([logits[k] for k in keys], [masks[k] for k in keys],
                                         [actions[k] for k in keys]).view(Lmax, R, 5)        # :387-390
            # GAE per rollout over ITS padded length: back-to-back segments, rollout-major
            values_lr = values.reshape(Lmax, R)
            if same:
                vals_c = values_lr.t().reshape(-1)
                rew_c = torch.from_numpy(rewards_np.reshape(R * Lmax, -1)).to(dev, non_blocking=True)
            else:
                vals_c = torch.cat([values_lr[:Lps[i], i] for i in range(R)])
                rew_c = torch.from_numpy(np.concatenate([rewards_np[i, :Lps[i]] for i in range(R)])).to(dev)
            seg = torch.tensor(np.concatenate([[0], np.cumsum(Lps)]), dtype=torch.int64, device=dev)
            adv_c, ret_c = ops.gae_scan(rew_c, vals_c, seg, gamma=GAMMA, lam=LAMBDA)          # :417-421
        return dict(obs=obs, masks=masks, actions=actions, rewards_np=rewards_np, old_logp=old_logp, values_lr=values_lr,
                    adv_c=adv_c, ret_c=ret_c, ybuf=ybuf, cbuf=cbuf, Ls=Ls, Lps=Lps, Lmax=Lmax, same=same)

    def experiences_from_rollouts(self, datas):
        """``experiences_from_rollout`` (:328-430) for all rollouts of an iteration at once: per rollout the result equals the
        per-rollout path -- own padding to a multiple of ``seq_len``, own terminal bootstrap, chunks beyond its padded
        length are not emitted -- but the work is one batched pass instead of ``R`` batch-1 passes."""
        S, pol = self.seq_len, self.policy_base
        p = self._prepare_rollouts(datas)
        obs, masks, actions, ybuf, cbuf, Lps = p['obs'], p['masks'], p['actions'], p['ybuf'], p['cbuf'], p['Lps']
        out = []
        for i, d in enumerate(datas):
            base = int(sum(Lps[:i]))
            sequences = []
            for j in range(Lps[i] // S):
                sl = slice(j * S, (j + 1) * S)
                if pol.cell == "lstm":
                    hid = (ybuf[j * S, i].reshape(1, 1, -1), cbuf[j * S, i].reshape(1, 1, -1))
                else:
                    hid = ybuf[j * S, i].reshape(1, 1, -1)
                seq = Sequence(game_id=d.get('game_id'), weight_version=d.get('weight_version'), team_id=d.get('team_id'),
                               observations={k: v[sl, i] for k, v in obs.items()},
                               actions={k: v[sl, i] for k, v in actions.items()},
                               masks={k: v[sl, i] for k, v in masks.items()},
                               values=p['values_lr'][sl, i].reshape(1, S, 1), rewards=p['rewards_np'][i, sl], hidden=hid,
                               old_logp=p['old_logp'][sl, i])
                seq.advantages = p['adv_c'][base + j * S: base + (j + 1) * S]
                seq.returns = p['ret_c'][base + j * S: base + (j + 1) * S]
                sequences.append(seq)
            out.append(sequences)
        return out

    def batch_from_rollouts(self, datas):
        """Rollouts -> one stacked, time-major ``ExperienceBatch`` (what ``train`` consumes; :587-615 stacks the same
        sequences batch-first), in the order ``experiences_from_rollouts`` emits them (rollout by rollout, chunk by chunk).
        Built from the prepared ``[L_max, R, ...]`` tensors with a handful of tensor ops per ROLLOUT -- no per-sequence
        Python objects, no per-sequence re-stacking (an iteration of the stream has ~1000 sequences of 16 steps)."""
        S, pol = self.seq_len, self.policy_base
        p = self._prepare_rollouts(datas)
        R, Lps = len(datas), p['Lps']
        n_chunks = [lp // S for lp in Lps]

        def chunked(t):                       # [Lmax, R, ...] -> [S, sum(n_chunks), ...]
            if p['same'] and p['Lmax'] == S:
                return t
            parts = [t[:Lps[i], i].reshape((n_chunks[i], S) + tuple(t.shape[2:])).transpose(0, 1) for i in range(R)]
            return torch.cat(parts, dim=1)
        obs = {k: chunked(v) for k, v in p['obs'].items()}
        masks = {k: chunked(v) for k, v in p['masks'].items()}
        actions = {k: chunked(v) for k, v in p['actions'].items()}
        old_logp = chunked(p['old_logp'])
        B = sum(n_chunks)
        adv = p['adv_c'].view(B, S).t().contiguous()                  # the GAE outputs are rollout-major back-to-back segments
        ret = p['ret_c'].view(B, S).t().contiguous()
        # hidden state entering chunk j of rollout i = state buffer slot j*S (:340,384-385: carried, not re-zeroed)
        t_idx = torch.tensor([j * S for i in range(R) for j in range(n_chunks[i])], dtype=torch.int64, device=self.device)
        r_idx = torch.tensor([i for i in range(R) for _ in range(n_chunks[i])], dtype=torch.int64, device=self.device)
        h0 = p['ybuf'][t_idx, r_idx].unsqueeze(0)
        c0 = p['cbuf'][t_idx, r_idx].unsqueeze(0) if pol.cell == "lstm" else None
        return ExperienceBatch(obs, masks, actions, old_logp, adv, ret, h0, c0)

    @staticmethod
    def list_of_dicts_to_dict_of_lists(x):
        return {k: torch.stack([torch.as_tensor(d[k]) for d in x]) for k in x[0]}

    # -- one optimizer step (:581-689) -------------------------------------------------------------
    def train(self, experiences):
        """One PPO/Adam step on a list of ``Sequence`` (or an ``ExperienceBatch``).

        Returns the reference's three dicts: losses, per-head entropies, grad norms (CPU scalars).
        A device-resident batch of a shape seen before is replayed from a CUDA graph of the whole step (forward, loss,
        backward, all-reduce, finish: ~80 kernel launches -> one graph launch); batches still in flight from the host
        (``prefetch``) run the same kernels launch by launch so that the upload overlaps them.
        """
        if isinstance(experiences, ExperienceBatch):
            batch = experiences if experiences.advantages.is_cuda else experiences.to(self.device)
        else:
            batch = ExperienceBatch.from_sequences(experiences, self.device)
        t_enter = time.perf_counter()
        slot = getattr(batch, "_slot", None)
        if slot is not None:                       # uploaded by prefetch() straight into a graph's static input buffers
            torch.cuda.current_stream().wait_event(slot["ready"])
            out, metrics = self._replay_step(batch, static=batch) if not ops.PROFILE.enabled else self._enqueue_step(batch)
            slot["done"].record()
            slot["busy"] = False
        elif self.use_cuda_graph and not ops.H2D_EVENTS and not ops.PROFILE.enabled:
            out, metrics = self._replay_step(batch)
        else:
            out, metrics = self._enqueue_step(batch)
        host = self._host_result
        host[:_lib.LOSS_SLOTS].copy_(out, non_blocking=True)
        host[_lib.LOSS_SLOTS:].copy_(metrics, non_blocking=True)
        self.host_enqueue_s = time.perf_counter() - t_enter   # host time to launch the step (the GPU runs behind it)
        torch.cuda.current_stream().synchronize()      # the step's single host sync (result read-back)
        res = host.clone()
        keys = ops.HEAD_KEYS
        if res[_lib.LOSS_SLOTS + 3] != 0:               # :667-669, :678-679 (parameters were left untouched)
            if math.isnan(float(res[0])):
                raise ValueError('loss={}, policy_loss={}, entropy_loss={}, value_loss={}'.format(
                    float(res[0]), float(res[1]), float(res[2]), float(res[3])))
            raise ValueError('grad_norm={}'.format(float(res[_lib.LOSS_SLOTS])))
        losses = {'loss': res[0], 'policy_loss': res[1], 'entropy_loss': res[2], 'value_loss': res[3]}
        entropies = {k: res[4 + h] for h, k in enumerate(keys)}
        return losses, entropies, {'unclipped': res[_lib.LOSS_SLOTS], 'clipped': res[_lib.LOSS_SLOTS + 1]}

    def _enqueue_step(self, batch):
        """Launches one optimizer step (:581-689) on the current stream; returns the device result vectors (loss slots, metrics)."""
        keys = ops.HEAD_KEYS
        self.flat.zero_grad_detached()                                    # :671 (grads gathered into the flat buffer below)
        hidden = (batch.h0, batch.c0) if self.policy_base.cell == "lstm" else batch.h0
        ddp = self.policy if isinstance(self.policy, DistributedDataParallelSparseParamCPU) else None
        if ddp is not None:
            ddp.auto_reduce = False        # the count-divide is fused into the finish kernel below
        ops.wait_h2d(batch.observations['env'], batch.h0, batch.c0)
        logits, values, _ = self.policy.forward_time_major(batch.observations, hidden)   # :619
        ops.wait_h2d(batch.old_logp, batch.advantages, batch.returns, *batch.masks.values(), *batch.actions.values(),
                     *batch.observations.values())
        packed = self.policy_base._packed_heads     # small heads + value are column ranges of one packed GEMM output
        out, n_actions, d_packed, d_tu = ops.ppo_loss_packed(
            packed, logits['target_unit'], [batch.masks[k] for k in keys], [batch.actions[k] for k in keys],
            batch.old_logp, batch.advantages, batch.returns, self.e_clip, self.entropy_coef, self.vf_coef)
        self._n_actions[:5].copy_(n_actions)
        torch.autograd.backward([packed, logits['target_unit']], [d_packed, d_tu])                     # :672
        # drop every reference into this step's autograd graph: a graph kept alive until the next forward keeps its saved
        # activations (GBs) AND the parameters' AccumulateGrad nodes, whose stream then mismatches a later graph capture
        self.policy_base._packed_heads = None
        del logits, values, packed, d_packed, d_tu
        self.flat.gather_grads()
        # distributed.py:29-57 -> flags + ONE all-reduce; divide fused into the finish kernel
        ops.grad_flags(self.flat.grad_full, self.flat.total, self.flat.seg_head, self._n_actions)
        if ddp is not None:
            ddp.allreduce_gradients(divide=False, flags_ready=True)
            ddp.needs_reduction = False
            ddp.auto_reduce = True
        ops.grad_finish(self.flat.param, self.flat.grad_full, self.exp_avg, self.exp_avg_sq, self.adam_steps,
                        self.flat.seg_lo, self.flat.seg_hi, self.flat.seg_head, self.flat.total, self.learning_rate, self.ADAM_BETAS,
                        self.ADAM_EPS, self.MAX_GRAD_NORM, out, self._metrics, self._finish_ws)     # :674-681
        return out, self._metrics

    # -- CUDA graph of the step ----------------------------------------------------------------------
    def _replay_step(self, batch, static=None):
        """Replays the captured step for this batch shape (captures it the second time the shape is seen: the first call of
        a shape runs launch by launch, which also warms every kernel up).  Inputs are copied into the graph's static
        buffers (device to device) -- or, for a batch that ``prefetch`` uploaded into an input slot (``static`` = the batch
        itself), are already there; parameters, gradients, Adam state and step counters are the same device buffers the
        eager path uses, so eager and graphed steps can be mixed freely."""
        key = (batch.seq_len, batch.batch_size) if static is None else (batch.seq_len, batch.batch_size, id(static._slot))
        entry = self._graphs.get(key)
        if entry is None:
            self._graphs[key] = "seen"
            return self._enqueue_step(batch)
        if entry == "seen":
            for k in [k for k, v in self._graphs.items() if isinstance(v, tuple)][:-1]:
                del self._graphs[k]            # at most two captured shapes alive: a graph pins its step's activations
            entry = self._capture_step(batch, static)
            self._graphs[key] = entry
            if entry == "eager":
                return self._enqueue_step(batch)
        elif entry == "eager":
            return self._enqueue_step(batch)
        graph_static, graph, out = entry
        if static is None:
            srcs = [v for _, _, v in batch.tensors()]
            dsts = [v for _, _, v in graph_static.tensors()]
            torch._foreach_copy_(dsts, srcs)
        graph.replay()
        return out, self._metrics

    def _capture_step(self, batch, static=None):
        if static is None:
            static = ExperienceBatch({}, {}, {}, None, None, None, None, None)
            for holder, k, v in batch.tensors():
                c = v.detach().clone()
                if isinstance(holder, dict):
                    (static.observations if holder is batch.observations else static.masks if holder is batch.masks else static.actions)[k] = c
                else:
                    setattr(static, k, c)
        graph = torch.cuda.CUDAGraph()
        try:
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                out, _ = self._enqueue_step(static)
        except Exception as e:                      # same kernels either way: fall back to launch-by-launch for this shape
            logger.warning('CUDA graph capture of the step failed (%s); this shape keeps running launch by launch', e)
            torch.cuda.synchronize()
            self.flat.rebind()
            return "eager"
        return static, graph, out

    def prefetch(self, experiences):
        """Starts the asynchronous upload of a pinned-host ``ExperienceBatch`` on the copy stream and returns the device batch
        at once (every tensor carries its own ready-event; ``train`` waits per tensor at first use).  Called while the previous
        step is still computing, it hides the PCIe transfer behind that step (double buffering) -- raw rollout data does not
        depend on the weights, so an optimizer fed from the experience queue can always upload one batch ahead."""
        if not isinstance(experiences, ExperienceBatch):
            experiences = ExperienceBatch.from_sequences(experiences, torch.device("cpu")).pin_memory()
        if experiences.advantages.is_cuda:
            return experiences
        if self.use_cuda_graph and experiences.advantages.is_pinned():
            staged = self._stage_into_slot(experiences)
            if staged is not None:
                return staged
        return experiences.to(self.device, prefetch=True)

    def _stage_into_slot(self, host):
        """Double-buffered graph inputs: every batch shape owns TWO sets of static device input buffers; ``prefetch`` copies
        the pinned host batch into the set that is not being trained on (copy stream, behind the replay that last read that
        set) and ``train`` replays the graph captured over that set -- so the upload of step k+1 overlaps the graph of step k.
        Returns None (caller falls back to per-tensor uploads) when both sets are still waiting to be trained on."""
        key = (host.seq_len, host.batch_size)
        slots = self._input_slots.setdefault(key, [])
        slot = next((sl for sl in slots if not sl["busy"]), None)
        if slot is None:
            if len(slots) >= 2:
                return None
            dev_batch = ExperienceBatch({}, {}, {}, None, None, None, None, None)
            for holder, k, v in host.tensors():
                c = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                if isinstance(holder, dict):
                    (dev_batch.observations if holder is host.observations else dev_batch.masks if holder is host.masks else dev_batch.actions)[k] = c
                else:
                    setattr(dev_batch, k, c)
            slot = {"batch": dev_batch, "busy": False, "ready": torch.cuda.Event(), "done": torch.cuda.Event()}
            slot["done"].record()
            dev_batch._slot = slot
            slots.append(slot)
        side = _copy_stream(self.device)
        side.wait_event(slot["done"])                    # the replay that last read these buffers has finished
        with torch.cuda.stream(side):
            dsts = [v for _, _, v in slot["batch"].tensors()]
            srcs = [v for _, _, v in host.tensors()]
            for d, src in zip(dsts, srcs):
                d.copy_(src, non_blocking=True)
            slot["ready"].record(side)
        slot["busy"] = True
        return slot["batch"]

    def mean_gradient_norm(self):
        """Mean per-tensor L2 norm over parameters that got a gradient in the last step (:691-695)."""
        has = self.flat.flags > 0
        norms = torch.stack([self.flat.grad[lo:hi].norm(2) for lo, hi in
                             zip(self.flat.starts, self.flat.ends)])
        return norms[has].mean()

    # -- iteration driver (:436-579) ----------------------------------------------------------------
    def run(self):
        for it in range(self.iteration_start, self.iterations):
            self.run_iteration(it)

    def run_iteration(self, it):
        logger.info('iteration {}/{}'.format(it, self.iterations))
        experiences, subrewards, rollout_lens, weight_ages = [], [], [], []
        start_xp = time.time()
        xp_waits = 0
        # The reference pulls and prepares rollouts one at a time until it holds min_seq_per_epoch sequences (:448-466).  The
        # number of sequences a rollout yields is known from its length alone, so the SAME rollouts are pulled here first and
        # then prepared together in one batched pass (experiences_from_rollouts).
        rollouts, n_seq = [], 0
        while n_seq < self.min_seq_per_epoch:                             # :448
            start_xp_wait = time.time()
            rollout, rollout_subrewards, rollout_len, weight_version, _ = self._next_rollout()
            xp_waits += time.time() - start_xp_wait
            rollouts.append(rollout)
            n_seq += (rollout_len + self.seq_len - 1) // self.seq_len
            subrewards.append(rollout_subrewards)
            rollout_lens.append(rollout_len)
            weight_ages.append(it - weight_version)
        batch = self.batch_from_rollouts(rollouts)                        # prepared + stacked once, reused by every epoch
        time_xp = time.time() - start_xp
        # a stream of rollouts gives every iteration its own batch size: capturing a graph per shape would cost more than the
        # `epochs` replays return, so the graph path is used only while consecutive iterations keep the same shape
        shape = (batch.seq_len, batch.batch_size)
        graph_setting, self.use_cuda_graph = self.use_cuda_graph, self.use_cuda_graph and shape == self._last_iteration_shape
        self._last_iteration_shape = shape

        losses, entropies, grad_norms = [], [], []
        start_optimizing = time.time()
        try:
            for ep in range(self.epochs):                                  # :469
                self.mq.process_data_events()
                loss_d, entropy_d, grad_norm_d = self.train(experiences=batch)
                losses.append(loss_d)
                entropies.append(entropy_d)
                grad_norms.append(grad_norm_d)
        finally:
            self.use_cuda_graph = graph_setting
        time_optimizing = time.time() - start_optimizing

        losses = self.list_of_dicts_to_dict_of_lists(losses)
        entropies = self.list_of_dicts_to_dict_of_lists(entropies)
        grad_norms = self.list_of_dicts_to_dict_of_lists(grad_norms)
        n_steps = batch.batch_size * self.seq_len                          # :486 (len(experiences) * seq_len)
        subrewards_per_sec = np.stack(subrewards) / n_steps * Policy.OBSERVATIONS_PER_SECOND
        reward_dict = dict(zip(REWARD_KEYS, subrewards_per_sec.sum(axis=0)))
        time_it = time.time() - self.time_last_it
        self.time_last_it = time.time()
        metrics = {
            self.SPEED_KEY: n_steps / time_it,                             # :501,505 (environment steps per second)
            'reward_per_sec/sum': subrewards_per_sec.sum(axis=1).sum(),
            'loss/sum': losses['loss'].mean(),
            'loss/policy': losses['policy_loss'].mean(),
            'loss/entropy': losses['entropy_loss'].mean(),
            'loss/value': losses['value_loss'].mean(),
            'entropy': torch.stack(list(entropies.values())).sum(dim=0).mean(),
            'avg_rollout_len': torch.tensor(rollout_lens, dtype=torch.float32).mean(),
            'avg_weight_age': torch.tensor(weight_ages, dtype=torch.float32).mean(),
            'timing/it': time_it, 'timing/xp_total': time_xp, 'timing/xp_mq_wait': xp_waits,
            'timing/optimizer': time_optimizing,
        }
        for k, v in entropies.items():
            metrics['entropy/{}'.format(k)] = v.mean()
        for k, v in grad_norms.items():
            metrics['grad_norm/{}'.format(k)] = v.mean()
        for k, v in reward_dict.items():
            metrics['reward_per_sec/{}'.format(k)] = v
        logger.info('steps_per_s={:.2f}, avg_weight_age={:.1f}, loss={:.4f}, entropy={:.3f}'.format(
            metrics[self.SPEED_KEY], float(metrics['avg_weight_age']), float(metrics['loss/sum']), float(metrics['entropy'])))
        if self.checkpoint:
            self.upload_model(version=it)                                  # :575
        self.last_metrics = metrics
        return metrics


class _FusedAdamHandle:
    """Minimal ``optimizer``-attribute stand-in: the Adam update itself is fused into ``dc_grad_finish``."""

    def __init__(self, owner):
        self._owner = owner
        self.defaults = {'lr': owner.learning_rate, 'betas': owner.ADAM_BETAS, 'eps': owner.ADAM_EPS, 'weight_decay': 0}

    @property
    def param_groups(self):
        return [dict(self.defaults, lr=self._owner.learning_rate, params=list(self._owner.flat.params))]

    def zero_grad(self, set_to_none=False):
        self._owner.flat.zero_grad()

    def state_dict(self):
        """``torch.optim.Adam.state_dict()`` layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq`` keyed by the
        parameter's index in ``named_parameters()`` order; tensors that never received a gradient have no entry, as in torch),
        so the checkpoint loads into a stock ``torch.optim.Adam`` over the same module and vice versa."""
        o = self._owner
        steps = o.adam_steps.cpu()
        state = {}
        for i, (p, lo, hi) in enumerate(zip(o.flat.params, o.flat.starts, o.flat.ends)):
            if int(steps[i]) > 0:
                state[i] = {'step': torch.tensor(float(steps[i])),
                            'exp_avg': o.exp_avg[lo:hi].view(p.shape).detach().cpu().clone(),
                            'exp_avg_sq': o.exp_avg_sq[lo:hi].view(p.shape).detach().cpu().clone()}
        group = dict(self.defaults, lr=o.learning_rate, amsgrad=False, maximize=False, foreach=None, capturable=False,
                     differentiable=False, fused=None, decoupled_weight_decay=False, params=list(range(o.flat.n_seg)))
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        o = self._owner
        if 'state' not in sd:                      # round-1 flat layout
            o.exp_avg.copy_(sd['exp_avg']); o.exp_avg_sq.copy_(sd['exp_avg_sq']); o.adam_steps.copy_(sd['step'])
            return
        o.exp_avg.zero_(); o.exp_avg_sq.zero_(); o.adam_steps.zero_()
        steps = torch.zeros(o.flat.n_seg, dtype=torch.int32)
        for i, st in sd['state'].items():
            i = int(i)
            lo, hi = o.flat.starts[i], o.flat.ends[i]
            o.exp_avg[lo:hi].copy_(st['exp_avg'].reshape(-1))
            o.exp_avg_sq[lo:hi].copy_(st['exp_avg_sq'].reshape(-1))
            steps[i] = int(st['step'])
        o.adam_steps.copy_(steps)


# ------------------------------------------------------------------------------------------ process entry
def init_distribution(backend='nccl'):
    """``env://`` rendezvous (:726-734); NCCL over NVLink instead of the reference's gloo over TCP."""
    assert 'WORLD_SIZE' in os.environ
    world_size = int(os.environ['WORLD_SIZE'])
    if world_size < 2:
        logger.warning('skipping distribution: world size too small ({})'.format(world_size))
        return
    if backend == 'nccl':
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    dist.init_process_group(backend=backend)
    logger.info("Distribution initialized.")


def main(rmq_host, rmq_port, epochs, min_seq_per_epoch, seq_len, learning_rate,
         pretrained_model, mq_prefetch_count, log_dir, entropy_coef, vf_coef, run_local,
         hidden_size=256, cell="gru"):
    if dist.is_available() and 'WORLD_SIZE' in os.environ:
        init_distribution()
    dota_optimizer = DotaOptimizer(
        rmq_host=rmq_host, rmq_port=rmq_port, epochs=epochs, min_seq_per_epoch=min_seq_per_epoch, seq_len=seq_len,
        learning_rate=learning_rate, checkpoint=is_master(), pretrained_model=pretrained_model,
        mq_prefetch_count=mq_prefetch_count, log_dir=log_dir, entropy_coef=entropy_coef, vf_coef=vf_coef,
        run_local=run_local, hidden_size=hidden_size, cell=cell)
    if isinstance(dota_optimizer.mq, MessageQueue):
        logger.warning('the built-in MessageQueue is an IN-PROCESS broker (the AMQP transport is out of scope): with no producer '
                       'thread publishing to it in this process run() will wait forever; pass mq=<your pika-backed queue> to '
                       'DotaOptimizer for a RabbitMQ deployment (--ip/--port are accepted for CLI compatibility only)')
    dota_optimizer.run()


def default_log_dir():
    return '{}_{}'.format(datetime.now().strftime('%b%d_%H-%M-%S'), socket.gethostname())


def build_arg_parser():
    """The reference's flags and defaults (:777-794) plus ``--hidden-size`` and ``--cell``."""
    p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--log-dir", type=str, help="log and job dir name", default=default_log_dir())
    p.add_argument("--ip", type=str, help="mq ip", default='127.0.0.1')
    p.add_argument("--port", type=int, help="mq port", default=5672)
    p.add_argument("--epochs", type=int, help="amount of epochs", default=4)
    p.add_argument("--min-seq-per-epoch", type=int, help="minimum amount of sequences per epoch", default=1024)
    p.add_argument("--seq-len", type=int, help="sequence length (truncated BPTT window)", default=16)
    p.add_argument("--learning-rate", type=float, help="learning rate", default=5e-5)
    p.add_argument("--entropy-coef", type=float, help="entropy coef", default=5e-4)
    p.add_argument("--vf-coef", type=float, help="value fn coef", default=0.5)
    p.add_argument("--pretrained-model", type=str, help="pretrained model file", default=None)
    p.add_argument("--mq-prefetch-count", type=int, help="experience messages to prefetch from mq", default=1)
    p.add_argument("-l", "--log", dest="log_level", help="Set the logging level",
                   choices=['DEBUG', 'INFO', 'WARNING', 'ERROR', 'CRITICAL'], default='INFO')
    p.add_argument("--run-local", type=bool, help="set to true to run locally (not using GCP)", default=False)
    p.add_argument("--hidden-size", type=int, help="recurrent width (reference: 256)", default=256)
    p.add_argument("--cell", type=str, choices=['gru', 'lstm'], help="recurrent cell (reference: gru)", default='gru')
    return p


if __name__ == '__main__':
    args = build_arg_parser().parse_args()
    logger.setLevel(args.log_level)
    try:
        main(rmq_host=args.ip, rmq_port=args.port, epochs=args.epochs, min_seq_per_epoch=args.min_seq_per_epoch,
             seq_len=args.seq_len, learning_rate=args.learning_rate, pretrained_model=args.pretrained_model,
             mq_prefetch_count=args.mq_prefetch_count, log_dir=args.log_dir, entropy_coef=args.entropy_coef,
             vf_coef=args.vf_coef, run_local=args.run_local, hidden_size=args.hidden_size, cell=args.cell)
    except KeyboardInterrupt:
        pass
