"""N > 1 host logic on CPU with the gloo backend (world size 2): the one collective of the design — rank 0's
packed ViT weight buffer broadcast to every rank — and per-rank pair assignment."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from splice_b200.dino_init import random_dino_state_dict
    from splice_b200.engine import pack_vit_weights, packed_key_order

    sd = random_dino_state_dict("dino_vits16", seed=1234)
    n = sum(v.numel() for v in sd.values())
    packed = pack_vit_weights(sd, "cpu") if rank == 0 else torch.zeros(n)
    dist.broadcast(packed, src=0)
    ref = pack_vit_weights(sd, "cpu")
    same = bool(torch.equal(packed, ref))
    # every rank optimises its own pair: seeds 1000+2r / 1001+2r
    from bench import synth_image

    a = synth_image(1000 + 2 * rank, 32, 8)
    sums = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(sums, a.sum().reshape(1))
    q.put((rank, same, [float(s) for s in sums], len(packed_key_order())))
    dist.barrier()
    dist.destroy_process_group()


def test_weight_broadcast_and_pair_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(same for _, same, _, _ in res)
    assert res[0][2] == res[1][2] and res[0][2][0] != res[0][2][1]   # two different pairs, consistent view
    assert res[0][3] == 150
