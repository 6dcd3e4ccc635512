"""2-rank NCCL test of the data-parallel step's only collective (needs 2 GPUs: `gpurun --gpus 2`; skipped on one):
after FlatTrainer.allreduce() every rank holds the SUM of the two ranks' flat gradients -- for torch.distributed's NCCL
all_reduce and for the C ABI's own communicator (cf_comm_*), on the gradients of a real X3D Bottleneck (tcgen05 weight-
gradient kernels accumulating straight into the flat buffer) with a different clip per rank."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _block(seed_weights=7):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from synth import synth_state_dict
    from coarse_fine_networks_b200 import x3d_fine
    blk = x3d_fine.Bottleneck(24, (54, 24), stride=1, downsample=None, index=0, base_bn_splits=1)
    blk.load_state_dict(synth_state_dict(blk.state_dict(), seed_weights))
    return blk.cuda().train()


def _grads_of(blk, trainer, seed):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from synth import synth_tensor
    trainer.zero_grad()
    x = synth_tensor((2, 24, 4, 28, 28), 100 + seed).cuda()
    out = blk(x)
    (out * synth_tensor(tuple(out.shape), 200 + seed).cuda()).sum().backward()
    return trainer.flat_g.clone()


def _worker(rank, world, port, native, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    dist.barrier()
    from coarse_fine_networks_b200 import train
    blk = _block()
    tr = train.FlatTrainer([blk], lr=0.01, native_comm=native)
    assert tr.world == world and (tr.comm is not None) == native
    mine = _grads_of(blk, tr, rank)
    others = [_grads_of(blk, tr, r) for r in range(world)]          # every rank's gradient, recomputed locally
    _grads_of(blk, tr, rank)                                        # leave this rank's own gradient in the flat buffer
    tr.allreduce()
    torch.cuda.synchronize()
    want = sum(others)
    err = float((tr.flat_g - want).abs().max() / want.abs().max())
    same = float((others[rank] - mine).abs().max())                # the backward is reproducible enough to compare sums
    q.put((rank, err, same, float(want.abs().max())))
    if tr.comm is not None:
        tr.comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("native", [False, True])
def test_two_rank_flat_gradient_allreduce(native):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000 + (1 if native else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, native, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    for rank, err, same, scale in res:
        # fp32 atomics order the weight-gradient partial sums differently from run to run: 1e-5 relative covers it
        assert err <= 1e-5 and same <= 1e-5 * scale, (rank, err, same, scale)
