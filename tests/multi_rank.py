"""Two-rank data-parallel check shared by ``tests/test_gpu_multi.py`` and ``__graft_entry__.smoke()`` (TEST INFRASTRUCTURE):
every rank runs the product's DotaOptimizer under NCCL on its own GPU; the parent compares with the N-rank oracle
(``oracle/ref_distributed.py``, itself pinned to the reference's ``distributed.py`` under gloo)."""
import os
import socket
import tempfile

import numpy as np
import torch

S, H, CELL, WORLD = 8, 128, "lstm", 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make(rank, log_dir, checkpoint=False):
    from dotaclient_b200.optimizer import DotaOptimizer
    return DotaOptimizer(rmq_host="multi%s" % log_dir, rmq_port=rank, epochs=1, min_seq_per_epoch=1, seq_len=S, learning_rate=5e-5,
                         checkpoint=checkpoint, pretrained_model=None, mq_prefetch_count=1, log_dir=log_dir,
                         entropy_coef=5e-4, vf_coef=0.5, run_local=True, hidden_size=H, cell=CELL)


def _init(rank, world, port):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    return dist


def step_worker(rank, world, port, out_dir):
    from dotaclient_b200.distributed import DistributedDataParallelSparseParamCPU
    from dotaclient_b200.synthetic import make_rollout
    dist = _init(rank, world, port)
    opt = _make(rank, tempfile.mkdtemp())
    assert isinstance(opt.policy, DistributedDataParallelSparseParamCPU)
    xs = opt.experiences_from_rollout(make_rollout(24, 300 + rank))     # prep works through the wrapper-era API
    recs = []
    for _ in range(3):                                                  # launch by launch, then the captured graph (with NCCL inside)
        l, e, g = opt.train(xs)
        recs.append(([float(l[k]) for k in ("loss", "policy_loss", "entropy_loss", "value_loss")],
                     float(g["unclipped"]), float(g["clipped"])))
    torch.save({"recs": recs, "sd": {k: v.cpu() for k, v in opt.policy_base.state_dict().items()}},
               os.path.join(out_dir, "rank%d.pt" % rank))
    opt.close()                                                         # graphs with NCCL work die before the process group
    dist.barrier()
    dist.destroy_process_group()


def resume_worker(rank, world, port, out_dir):
    """Train, checkpoint (master only), build a NEW optimizer on the same log_dir (resume), train again."""
    from dotaclient_b200.optimizer import is_master
    from dotaclient_b200.synthetic import make_rollout
    dist = _init(rank, world, port)
    log_dir = os.path.join(out_dir, "ckpt")
    opt = _make(rank, log_dir, checkpoint=is_master())
    xs = opt.experiences_from_rollout(make_rollout(24, 500 + rank))
    for _ in range(2):
        opt.train(xs)
    opt.upload_model(version=5)                                         # master writes model_000000005.pt + adam_000000005.state
    dist.barrier()
    opt2 = _make(rank, log_dir, checkpoint=is_master())                 # only the master finds + restores the checkpoint
    xs2 = opt2.experiences_from_rollout(make_rollout(24, 500 + rank))
    opt2.train(xs2)
    torch.save({"iteration_start": opt2.iteration_start, "steps": opt2.adam_steps.cpu(), "exp_avg": opt2.exp_avg.cpu(),
                "param": opt2.flat.param.cpu(), "exp_avg_before": opt.exp_avg.cpu()}, os.path.join(out_dir, "resume%d.pt" % rank))
    opt.close()
    opt2.close()
    dist.barrier()
    dist.destroy_process_group()


def run_two_rank_step_check(out_dir):
    """Spawns the two ranks and compares with the two-rank oracle.  Raises AssertionError on mismatch."""
    import torch.multiprocessing as mp
    from oracle import ref_distributed, ref_optimizer as RO
    from oracle.ref_policy import RefPolicy
    from dotaclient_b200.synthetic import make_rollout
    mp.spawn(step_worker, args=(WORLD, _free_port(), str(out_dir)), nprocs=WORLD, join=True)
    opts = []
    for _ in range(WORLD):
        torch.manual_seed(7)
        opts.append(RO.RefOptimizer(RefPolicy(H, CELL), seq_len=S))
    shards = [opts[r].experiences_from_rollout(make_rollout(24, 300 + r)) for r in range(WORLD)]
    oracle = [ref_distributed.train_ranks(opts, shards) for _ in range(3)]
    got = [torch.load(os.path.join(str(out_dir), "rank%d.pt" % r)) for r in range(WORLD)]
    for k in got[0]["sd"]:
        assert torch.equal(got[0]["sd"][k], got[1]["sd"][k]), k            # replicas stay bit-identical
    for r in range(WORLD):
        for ep in range(3):
            l, e, g = oracle[ep][r]
            want = [float(l[k]) for k in ("loss", "policy_loss", "entropy_loss", "value_loss")]
            np.testing.assert_allclose(got[r]["recs"][ep][0], want, rtol=2e-4, atol=2e-6)
            np.testing.assert_allclose(got[r]["recs"][ep][1], float(g["unclipped"]), rtol=2e-3)
            np.testing.assert_allclose(got[r]["recs"][ep][2], float(g["clipped"]), rtol=2e-3)
    torch.manual_seed(7)
    init = RefPolicy(H, CELL).state_dict()
    dm = torch.cat([(got[0]["sd"][k] - init[k]).flatten() for k in init])
    do = torch.cat([(opts[0].policy_base.state_dict()[k] - init[k]).flatten() for k in init])
    assert torch.nn.functional.cosine_similarity(dm, do, dim=0) > 0.995
    return got
