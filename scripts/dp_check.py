"""Multi-GPU check of the data-parallel path (run under torchrun, one rank per GPU):
N-rank training of one model must reproduce the single-GPU loss sequence within float32
tolerance, and every rank must end with identical parameters.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/dp_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_training_b200 import dp, init  # noqa: E402
from vae_training_b200.packed import RunSpec  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h = [256, 256, 256]
    spec = RunSpec("sphere", 32, 32, 16, h, h, batch_size=4096, learning_rate=1e-4, epsilon=-3.0,
                   tunable_decoder_var=True, dataset_seed=69)
    st = init.initial_state(spec)
    steps = 12
    model = dp.DataParallelVAE(spec, st.params, st.model_key, st.dataset_key, st.A)
    losses = []
    for _ in range(steps):
        losses.append(model.train_step().clone())
    losses = torch.cat(losses).cpu().numpy()
    # parameters identical on every rank
    ref = model.params.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(ref, model.params))
    flags = torch.tensor([int(same)], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok = bool(flags.item())
    if rank == 0:
        single = dp.DataParallelVAE(spec, st.params, st.model_key, st.dataset_key, st.A, rank=0, world=1)
        l1 = torch.cat([single.train_step().clone() for _ in range(steps)]).cpu().numpy()
        rel = float(np.max(np.abs(losses - l1) / np.abs(l1)))
        pdiff = float((single.params - model.params).abs().max())
        print(f"dp_check world={world}: loss rel diff vs 1 GPU {rel:.3e}, max |dparam| {pdiff:.3e}, "
              f"replicas identical: {ok}", flush=True)
        assert rel <= 1e-5 and ok and pdiff <= 0.5 * steps * spec.learning_rate
        print("DP CHECK OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
