"""2-GPU test of the data-parallel step: NCCL all-reduce of the flat gradient buffer + fused finish, against
the N-rank oracle (``oracle/ref_distributed.py``, itself pinned to the reference's ``distributed.py`` under gloo)."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

S, H, CELL, WORLD = 8, 128, "lstm", 2


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from dotaclient_b200.optimizer import DotaOptimizer
    from dotaclient_b200.distributed import DistributedDataParallelSparseParamCPU
    from dotaclient_b200.synthetic import make_rollout
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    opt = DotaOptimizer(rmq_host="multi", rmq_port=rank, epochs=1, min_seq_per_epoch=1, seq_len=S, learning_rate=5e-5,
                        checkpoint=False, pretrained_model=None, mq_prefetch_count=1, log_dir=tempfile.mkdtemp(),
                        entropy_coef=5e-4, vf_coef=0.5, run_local=True, hidden_size=H, cell=CELL)
    assert isinstance(opt.policy, DistributedDataParallelSparseParamCPU)
    xs = opt.experiences_from_rollout(make_rollout(24, 300 + rank))     # prep works through the wrapper-era API
    recs = []
    for _ in range(2):
        l, e, g = opt.train(xs)
        recs.append(([float(l[k]) for k in ("loss", "policy_loss", "entropy_loss", "value_loss")],
                     float(g["unclipped"]), float(g["clipped"])))
    torch.save({"recs": recs, "sd": {k: v.cpu() for k, v in opt.policy_base.state_dict().items()}},
               os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_step_matches_nrank_oracle(tmp_path):
    import torch.multiprocessing as mp
    from oracle import ref_distributed, ref_optimizer as RO
    from oracle.ref_policy import RefPolicy
    from dotaclient_b200.synthetic import make_rollout
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(WORLD, port, str(tmp_path)), nprocs=WORLD, join=True)
    opts = []
    for _ in range(WORLD):
        torch.manual_seed(7)
        opts.append(RO.RefOptimizer(RefPolicy(H, CELL), seq_len=S))
    shards = [opts[r].experiences_from_rollout(make_rollout(24, 300 + r)) for r in range(WORLD)]
    oracle = [ref_distributed.train_ranks(opts, shards) for _ in range(2)]
    got = [torch.load(tmp_path / ("rank%d.pt" % r)) for r in range(WORLD)]
    for k in got[0]["sd"]:
        assert torch.equal(got[0]["sd"][k], got[1]["sd"][k]), k            # replicas stay bit-identical
    for r in range(WORLD):
        for ep in range(2):
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
